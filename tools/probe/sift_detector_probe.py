"""Probe (build container, needs cv2): how far a restatement of cv2.SIFT_create().detectAndCompute gets for the DETECTOR
(scale space from cv2 primitives, extrema, adjustLocalExtrema) — groundwork for a SIFT front-end, not an oracle yet.

Findings (OpenCV 4.13 in this image: baseline SSE3, dispatched AVX2 + FMA3 / AVX512_SKX; this host takes a dispatched path):
  * the keypoint list is deterministic and independent of the thread count (it is sorted by removeDuplicatedSorted);
  * structure right at the first try: 498 / 500 distinct positions bit-equal with plain fp32 arithmetic;
  * the scalar code of sift.simd.hpp is built with FMA contraction.  contr = fma(img, img_scale, t / 2) with
    t = fma(dD2, xi, fma(dD1, xr, dD0 * xc)) takes the response mismatches from 113 to 3; contracting the 3 x 3 solve
    (Matx_FastSolveOp) as p*q - r*s -> fma(p, q, -(r*s)) and X*a - Y*b + Z*c -> fma(Z, c, fma(X, a, -(Y*b))), with the first
    minor of x(2) taken as the NEGATED third minor of x(0) (common-subexpression reuse: fma(-r, s, p*q)), makes all 500
    (pt, size, response, octave) records bit-equal (defaults: M=1AA M2=BAAA);
  * consequence: SIFT's low-order bits depend on which dispatched build the host CPU selects (contraction exists only in
    the AVX2 / AVX-512 objects, and calcOrientationHist / the descriptor have vector bodies with contracted scalar tails
    whose split depends on the vector width).  Bit-exact SIFT parity is therefore not a property of the reference but of
    the reference on one CPU; the realistic bar for a GPU SIFT front-end is: same keypoint set with positions to ~1e-4 px,
    angles to ~1e-3 degrees, descriptor entries within +-1;
  * not started: calcOrientationHist (hal::exp32f / fastAtan2 / magnitude32f), duplicate removal, the descriptor.
usage: python tools/probe/sift_detector_probe.py      (M / M2 select other contraction hypotheses)"""
import cv2, numpy as np
F=np.float32
def fma(a,b,c): return np.float32(np.float64(a)*np.float64(b)+np.float64(c))
import os, ctypes
MODE=os.environ.get('M','1AA')
_libm=ctypes.CDLL('libm.so.6'); _libm.powf.restype=ctypes.c_float; _libm.powf.argtypes=[ctypes.c_float,ctypes.c_float]
def powf(a,b): return np.float32(_libm.powf(float(a),float(b)))
_libm.exp2f.restype=ctypes.c_float; _libm.exp2f.argtypes=[ctypes.c_float]
def exp2f(b): return np.float32(_libm.exp2f(float(b)))
def build(img, nOctaveLayers=3, sigma=1.6):
    gray=img.astype(np.float32)
    sig_diff=np.sqrt(max(F(sigma)*F(sigma)-F(0.5)*F(0.5)*4,F(0.01))).astype(np.float32)
    dbl=cv2.resize(gray,(gray.shape[1]*2,gray.shape[0]*2),interpolation=cv2.INTER_LINEAR)
    basei=cv2.GaussianBlur(dbl,(0,0),sigmaX=float(sig_diff),sigmaY=float(sig_diff))
    nOct=int(np.rint(np.log(float(min(basei.shape)))/np.log(2.0)-2))-(-1)   # firstOctave=-1 -> base already doubled
    # OpenCV: nOctaves = cvRound(log(min(base.cols,base.rows))/log(2) - 2) - firstOctave, with base = doubled image
    nOct=int(np.rint(np.log(float(min(basei.shape)))/np.log(2.0)-2))+1
    sig=[sigma]; k=2.0**(1.0/nOctaveLayers)
    for i in range(1,nOctaveLayers+3):
        sp=(k**(i-1))*sigma; st=sp*k; sig.append(np.sqrt(st*st-sp*sp))
    pyr=[]
    for o in range(nOct):
        for i in range(nOctaveLayers+3):
            if o==0 and i==0: d=basei
            elif i==0:
                src=pyr[(o-1)*(nOctaveLayers+3)+nOctaveLayers]
                d=cv2.resize(src,(src.shape[1]//2,src.shape[0]//2),interpolation=cv2.INTER_NEAREST)
            else:
                d=cv2.GaussianBlur(pyr[-1],(0,0),sigmaX=sig[i],sigmaY=sig[i])
            pyr.append(d)
    dog=[]
    for o in range(nOct):
        for i in range(nOctaveLayers+2):
            dog.append(cv2.subtract(pyr[o*(nOctaveLayers+3)+i+1],pyr[o*(nOctaveLayers+3)+i]))
    return pyr,dog,nOct
X2MODE=os.environ.get('M2','BAAA')
def m2(p,q,r,s_,ov=None):
    """p*q - r*s_ as GCC contracts it (MODE[1]: A = fma(p,q,-(r*s)), B = fma(-r,s,p*q), 0 = none)"""
    c=ov if ov else (MODE[1] if len(MODE)>1 else '0')
    if c=='A': return fma(p,q,-(r*s_))
    if c=='B': return fma(-r,s_,p*q)
    return p*q-r*s_
def comb(X,a,Y,b,Z,c_,ov=None):
    """X*a - Y*b + Z*c"""
    c=ov if ov else (MODE[2] if len(MODE)>2 else '0')
    if c=='C': return fma(X,a,fma(Z,c_,-(Y*b)))        # X*a + (Z*c - Y*b)
    if c=='D': return fma(-Y,b,fma(Z,c_,X*a))
    if c=='A': return fma(Z,c_,fma(X,a,-(Y*b)))
    if c=='B': return fma(Z,c_,fma(-Y,b,X*a))
    return X*a-Y*b+Z*c_
def solve3(H,b):
    a=H
    d=comb(a[0,0],m2(a[1,1],a[2,2],a[2,1],a[1,2]),a[0,1],m2(a[1,0],a[2,2],a[2,0],a[1,2]),a[0,2],m2(a[1,0],a[2,1],a[2,0],a[1,1]))
    d=F(d)
    if d==0: return None
    d=F(1)/d
    x0=d*comb(b[0],m2(a[1,1],a[2,2],a[1,2],a[2,1]),a[0,1],m2(b[1],a[2,2],a[1,2],b[2]),a[0,2],m2(b[1],a[2,1],a[1,1],b[2]))
    x1=d*comb(a[0,0],m2(b[1],a[2,2],a[1,2],b[2]),b[0],m2(a[1,0],a[2,2],a[1,2],a[2,0]),a[0,2],m2(a[1,0],b[2],b[1],a[2,0]))
    o1=X2MODE[0] if X2MODE else None; o2=X2MODE[1] if len(X2MODE)>1 else None; o3=X2MODE[2] if len(X2MODE)>2 else o1
    x2=d*comb(a[0,0],m2(a[1,1],b[2],b[1],a[2,1],o1),a[0,1],m2(a[1,0],b[2],b[1],a[2,0],(X2MODE[3] if len(X2MODE)>3 else o1)),b[0],m2(a[1,0],a[2,1],a[1,1],a[2,0],o3),o2)
    return np.array([x0,x1,x2],np.float32)
def adjust(dog,octv,layer,r,c,nOL=3,contr_thr=F(0.04),edge_thr=F(10.0),sigma=F(1.6)):
    img_scale=F(1.0)/F(255); deriv_scale=img_scale*F(0.5); sds=img_scale; cds=img_scale*F(0.25)
    xi=xr=xc=F(0)
    for it in range(5):
        idx=octv*(nOL+2)+layer
        img=dog[idx];prev=dog[idx-1];nxt=dog[idx+1]
        dD=np.array([(img[r,c+1]-img[r,c-1])*deriv_scale,(img[r+1,c]-img[r-1,c])*deriv_scale,(nxt[r,c]-prev[r,c])*deriv_scale],np.float32)
        v2=img[r,c]*F(2)
        dxx=(img[r,c+1]+img[r,c-1]-v2)*sds; dyy=(img[r+1,c]+img[r-1,c]-v2)*sds; dss=(nxt[r,c]+prev[r,c]-v2)*sds
        dxy=(img[r+1,c+1]-img[r+1,c-1]-img[r-1,c+1]+img[r-1,c-1])*cds
        dxs=(nxt[r,c+1]-nxt[r,c-1]-prev[r,c+1]+prev[r,c-1])*cds
        dys=(nxt[r+1,c]-nxt[r-1,c]-prev[r+1,c]+prev[r-1,c])*cds
        H=np.array([[dxx,dxy,dxs],[dxy,dyy,dys],[dxs,dys,dss]],np.float32)
        X=solve3(H,dD)
        if X is None: X=np.zeros(3,np.float32)
        xi=-X[2];xr=-X[1];xc=-X[0]
        if abs(xi)<0.5 and abs(xr)<0.5 and abs(xc)<0.5: break
        if abs(xi)>2**31/3 or abs(xr)>2**31/3 or abs(xc)>2**31/3: return None
        c+=int(np.rint(xc)); r+=int(np.rint(xr)); layer+=int(np.rint(xi))
        if layer<1 or layer>nOL or c<5 or c>=img.shape[1]-5 or r<5 or r>=img.shape[0]-5: return None
    else:
        return None
    idx=octv*(nOL+2)+layer
    img=dog[idx];prev=dog[idx-1];nxt=dog[idx+1]
    dD=np.array([(img[r,c+1]-img[r,c-1])*deriv_scale,(img[r+1,c]-img[r-1,c])*deriv_scale,(nxt[r,c]-prev[r,c])*deriv_scale],np.float32)
    if MODE[0]=='0':
        t=dD[0]*xc+dD[1]*xr+dD[2]*xi
        contr=img[r,c]*img_scale+t*F(0.5)
    else:
        t=fma(dD[2],xi,fma(dD[1],xr,dD[0]*xc))
        contr=fma(img[r,c],img_scale,t*F(0.5))
    if abs(contr)*nOL<contr_thr: return None
    v2=img[r,c]*F(2)
    dxx=(img[r,c+1]+img[r,c-1]-v2)*sds; dyy=(img[r+1,c]+img[r-1,c]-v2)*sds
    dxy=(img[r+1,c+1]-img[r+1,c-1]-img[r-1,c+1]+img[r-1,c-1])*cds
    tr=dxx+dyy; det=dxx*dyy-dxy*dxy
    if det<=0 or tr*tr*edge_thr>=(edge_thr+1)*(edge_thr+1)*det: return None
    ptx=(F(c)+xc)*F(1<<octv); pty=(F(r)+xr)*F(1<<octv)
    octave=octv+(layer<<8)+(int(np.rint((np.float64(xi)+0.5)*255))<<16)
    size=sigma*exp2f((F(layer)+xi)/F(nOL))*F(1<<octv)*F(2)
    return (ptx,pty,octave,size,abs(contr),r,c,layer)
def detect(img):
    pyr,dog,nOct=build(img)
    nOL=3; thr=int(np.floor(0.5*0.04/nOL*255))
    out=[]
    for o in range(nOct):
        for i in range(1,nOL+1):
            idx=o*(nOL+2)+i
            cur=dog[idx];prev=dog[idx-1];nxt=dog[idx+1]
            H,W=cur.shape
            if H<=10 or W<=10: continue
            core=cur[5:H-5,5:W-5]
            stack=[]
            for im in (prev,cur,nxt):
                for dy in (-1,0,1):
                    for dx in (-1,0,1):
                        stack.append(im[5+dy:H-5+dy,5+dx:W-5+dx])
            st=np.stack(stack,0)
            mx=st.max(0); mn=st.min(0)
            cand=(np.abs(core)>thr)&(((core>0)&(core>=mx))|((core<0)&(core<=mn)))
            rs,cs=np.nonzero(cand)
            for r,c in zip(rs+5,cs+5):
                k=adjust(dog,o,i,int(r),int(c))
                if k is not None: out.append(k)
    return out,pyr,dog
if __name__=="__main__":
    rng=np.random.default_rng(5)
    base=rng.integers(0,256,(30,50),dtype=np.uint8)
    img=cv2.resize(base,(400,240),interpolation=cv2.INTER_CUBIC)
    img=cv2.add(img,rng.integers(0,12,img.shape,dtype=np.uint8))
    kps,_=cv2.SIFT_create().detectAndCompute(img,None)
    # cv2 keypoints after firstOctave=-1 adjustment: pt*0.5,size*0.5, octave = (octave & ~255) | ((octave + firstOctave) & 255)
    ref=set()
    for k in kps: ref.add((np.float32(k.pt[0]),np.float32(k.pt[1]),np.float32(k.size),np.float32(k.response),k.octave))
    out,pyr,dog=detect(img)
    got=set()
    for (x,y,octave,size,resp,r,c,layer) in out:
        oc=(octave&~255)|((octave-1)&255)
        got.add((np.float32(x*np.float32(0.5)),np.float32(y*np.float32(0.5)),np.float32(size*np.float32(0.5)),np.float32(resp),oc))
    print(len(ref),len(got),len(ref&got))
    # looser: positions only
    rp={(a,b) for a,b,_,_,_ in ref}; gp={(a,b) for a,b,_,_,_ in got}
    print("pos match",len(rp&gp),len(rp),len(gp))
    rd={(a,b):(s,r,o) for a,b,s,r,o in ref}; gd={(a,b):(s,r,o) for a,b,s,r,o in got}
    ns=nr=no=0; ex=[]
    for k in rd:
        if k in gd:
            ns+=rd[k][0]!=gd[k][0]; nr+=rd[k][1]!=gd[k][1]; no+=rd[k][2]!=gd[k][2]
            if rd[k]!=gd[k] and len(ex)<6: ex.append((k,rd[k],gd[k]))
    print("size diff",ns,"resp diff",nr,"octave diff",no)
    for e in ex: print(e)
    print([k for k in rd if k not in gd][:3],[k for k in gd if k not in rd][:3])
