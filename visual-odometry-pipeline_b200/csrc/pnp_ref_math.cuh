// Arithmetic of the REFERENCE-SAMPLER PnP-RANSAC ("Mode R"): what cv2.solvePnPRansac(iterationsCount=100,
// reprojectionError=1.5) does inside VisualOdometry.computepose_3D_2D (VisualOdometry_Stereo.py:120-135), restated from the
// published algorithms so that it compiles for the device and for the host (tests/host_math_shim.cpp):
//   * the sample table: OpenCV's multiply-with-carry RNG, re-seeded with 2^64 - 1 on every call, five distinct indices per
//     iteration (calib3d ptsetreg.cpp RANSACPointSetRegistrator::getSubset) — a function of the point count only;
//   * the minimal solver: EPnP on five points (Lepetit, Moreno-Noguer, Fua 2009; calib3d epnp.cpp): PCA control points,
//     barycentric coordinates, the 10 x 12 system M, the four eigenvectors of M^T M with the smallest eigenvalues, beta
//     initialisations for N = 1, 2, 3, five Gauss-Newton steps on the betas each, absolute orientation, lowest mean
//     reprojection error wins;
//   * the inlier rule: cv2.projectPoints in double, rounded to float, squared pixel distance in float, err <= (float)(1.5 * 1.5);
//   * the adaptive iteration count of RANSACUpdateNumIters (confidence 0.99, model points 5).
// Two deliberate differences from OpenCV's binary, both forced: with five points M^T M has a two-dimensional null
// space, and OpenCV takes whatever basis of it its SVD returns — rounding noise, different per build and CPU
// (tools/probe/epnp_basis_probe.py: the per-hypothesis poses of two implementations of the same algorithm differ by
// millimetres).  Here the null-space basis is made canonical (see canonical_null_basis), so that this header, the numpy
// restatement of the test suite (LAPACK eigen-solver) and the kernels agree to ~1e-9; and the absolute orientation uses Horn's
// quaternion form (identical to U V^T whenever det(U V^T) > 0; a proper rotation in the reflected, degenerate case too).
// Parity with the reference's own call is therefore exact for the sampler / scoring / stopping rule (tests/test_oracle_pnp_ref.py
// drives the same control flow with cv2's minimal solver and reproduces cv2.solvePnPRansac bit for bit) and statistical for the
// minimal solver (tests/test_gpu_trajectory_long.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define VO_RHD __host__ __device__ __forceinline__
#define VO_RHDN __host__ __device__ __noinline__
#else
#define VO_RHD inline
#define VO_RHDN inline
#endif

namespace vo {
namespace refpnp {

constexpr int EP_N = 5;          // model points of solvePnPRansac's minimal kernel (SOLVEPNP_EPNP)

// ---- OpenCV's RNG and the RANSAC sample table -------------------------------------------------------------------------
VO_RHD uint32_t mwc_next(uint64_t &s) {
    s = (uint64_t)(uint32_t)s * 4164903690ull + (s >> 32);
    return (uint32_t)s;
}
// table[it][0..4]: the five distinct indices iteration `it` of a solvePnPRansac call draws from n points (n >= 5)
VO_RHD void mwc_table(int n, int iters, int32_t *table) {
    uint64_t s = 0xffffffffffffffffull;
    for (int it = 0; it < iters; ++it) {
        int32_t *row = table + it * EP_N;
        int i = 0;
        while (i < EP_N) {
            const int32_t v = (int32_t)(mwc_next(s) % (uint32_t)n);
            bool dup = false;
            for (int j = 0; j < i; ++j) dup = dup || (row[j] == v);
            if (dup) continue;
            row[i++] = v;
        }
    }
}

// ---- small dense helpers (double) ---------------------------------------------------------------------------------------
// Cyclic Jacobi eigenvalue iteration on a symmetric N x N matrix: A is destroyed, V's COLUMNS are the eigenvectors, d the
// eigenvalues; sorted by descending eigenvalue afterwards.
template <int N>
VO_RHDN void jacobi_eig(double (&A)[N][N], double (&V)[N][N], double (&d)[N]) {
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < N; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {          // A <- J^T A J, rows / columns p and q
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < N; ++i) d[i] = A[i][i];
    for (int i = 0; i < N - 1; ++i) {          // selection sort, descending
        int m = i;
        for (int j = i + 1; j < N; ++j) m = (d[j] > d[m]) ? j : m;
        if (m != i) {
            const double td = d[i]; d[i] = d[m]; d[m] = td;
            for (int k = 0; k < N; ++k) { const double tv = V[k][i]; V[k][i] = V[k][m]; V[k][m] = tv; }
        }
    }
}

// The 12 x 12 case of EPnP in PARALLEL ORDER: a sweep is 11 rounds of 6 rotations on disjoint index pairs (round-robin
// tournament: index 11 stays, the others rotate), and a round is applied in two phases — every pair's angle from the matrix as it
// stands (disjoint pairs do not touch each other's a_pp, a_qq, a_pq), then all column updates A <- A J, V <- V J, then all row
// updates A <- J^T A.  The phases make the arithmetic of every element independent of the order in which the six pairs are
// visited, so a warp can do them side by side (pnp.cu: jacobi12_warp) and lands on the same bits as this serial loop.
// Converges like the cyclic-by-row order (one sweep more at most); same stopping rule and sorting as jacobi_eig.
VO_RHD void jacobi12_pair(int round, int k, int &p, int &q) {
    const int a = (k == 0) ? 11 : (round + k) % 11, b = (k == 0) ? round : (round + 11 - k) % 11;
    p = a < b ? a : b;
    q = a < b ? b : a;
}
VO_RHD void jacobi12_angle(double app, double aqq, double apq, double &c, double &s) {
    if (apq == 0.0) { c = 1.0; s = 0.0; return; }
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    c = 1.0 / sqrt(t * t + 1.0);
    s = t * c;
}
VO_RHDN void jacobi_eig12_rr(double (&A)[12][12], double (&V)[12][12], double (&d)[12]) {
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < 12; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < 12; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-30 * diag || off == 0.0) break;
        for (int round = 0; round < 11; ++round) {
            int P[6], Q[6];
            double C[6], S[6];
            for (int k = 0; k < 6; ++k) {
                jacobi12_pair(round, k, P[k], Q[k]);
                jacobi12_angle(A[P[k]][P[k]], A[Q[k]][Q[k]], A[P[k]][Q[k]], C[k], S[k]);
            }
            for (int k = 0; k < 6; ++k)
                for (int i = 0; i < 12; ++i) {      // columns p, q of A and V
                    const double akp = A[i][P[k]], akq = A[i][Q[k]];
                    A[i][P[k]] = C[k] * akp - S[k] * akq;
                    A[i][Q[k]] = S[k] * akp + C[k] * akq;
                    const double vkp = V[i][P[k]], vkq = V[i][Q[k]];
                    V[i][P[k]] = C[k] * vkp - S[k] * vkq;
                    V[i][Q[k]] = S[k] * vkp + C[k] * vkq;
                }
            for (int k = 0; k < 6; ++k)
                for (int j = 0; j < 12; ++j) {      // rows p, q of A
                    const double apk = A[P[k]][j], aqk = A[Q[k]][j];
                    A[P[k]][j] = C[k] * apk - S[k] * aqk;
                    A[Q[k]][j] = S[k] * apk + C[k] * aqk;
                }
        }
    }
    for (int i = 0; i < 12; ++i) d[i] = A[i][i];
    for (int i = 0; i < 11; ++i) {          // selection sort, descending
        int m = i;
        for (int j = i + 1; j < 12; ++j) m = (d[j] > d[m]) ? j : m;
        if (m != i) {
            const double td = d[i]; d[i] = d[m]; d[m] = td;
            for (int k = 0; k < 12; ++k) { const double tv = V[k][i]; V[k][i] = V[k][m]; V[k][m] = tv; }
        }
    }
}

// Least squares min |A x - b| for an R x C system (R >= C) by Householder QR; a column that is numerically zero gets x = 0.
template <int R, int C>
VO_RHDN void lsq_qr(double (&A)[R][C], double (&b)[R], double (&x)[C]) {
    double rdiag[C];
    for (int k = 0; k < C; ++k) {
        double nrm = 0.0;
        for (int i = k; i < R; ++i) nrm += A[i][k] * A[i][k];
        nrm = sqrt(nrm);
        if (!(nrm > 1e-300)) { rdiag[k] = 0.0; continue; }
        const double alpha = (A[k][k] > 0.0) ? -nrm : nrm;
        A[k][k] -= alpha;                                   // v = column k below the diagonal, in place
        double vv = 0.0;
        for (int i = k; i < R; ++i) vv += A[i][k] * A[i][k];
        if (vv > 0.0) {
            for (int j = k + 1; j < C; ++j) {
                double s = 0.0;
                for (int i = k; i < R; ++i) s += A[i][k] * A[i][j];
                s = 2.0 * s / vv;
                for (int i = k; i < R; ++i) A[i][j] -= s * A[i][k];
            }
            double s = 0.0;
            for (int i = k; i < R; ++i) s += A[i][k] * b[i];
            s = 2.0 * s / vv;
            for (int i = k; i < R; ++i) b[i] -= s * A[i][k];
        }
        rdiag[k] = alpha;
    }
    for (int k = C - 1; k >= 0; --k) {
        if (rdiag[k] == 0.0) { x[k] = 0.0; continue; }
        double s = b[k];
        for (int j = k + 1; j < C; ++j) s -= A[k][j] * x[j];
        x[k] = s / rdiag[k];
    }
}

VO_RHD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Sign convention of an eigenvector: the component of largest magnitude (lowest index on ties) is positive.
template <int N>
VO_RHD void fix_sign(double (&v)[N]) {
    int m = 0;
    for (int k = 1; k < N; ++k) m = (fabs(v[k]) > fabs(v[m])) ? k : m;
    if (v[m] < 0.0)
        for (int k = 0; k < N; ++k) v[k] = -v[k];
}

// Canonical basis of the two-dimensional null space spanned by the orthonormal pair (a, b): the projector P = a a^T + b b^T
// does not depend on the pair; w1 = P e_k / |P e_k| for the coordinate k with the largest P_kk, w0 = the unit vector of the
// space orthogonal to w1, sign-fixed.  (w1 plays ut[11], w0 plays ut[10] of epnp.cpp.)
VO_RHD void canonical_null_basis(double (&a)[12], double (&b)[12]) {
    int k = 0;
    double best = -1.0;
    for (int i = 0; i < 12; ++i) {
        const double p = a[i] * a[i] + b[i] * b[i];
        if (p > best) { best = p; k = i; }
    }
    const double ca = a[k], cb = b[k], nrm = sqrt(ca * ca + cb * cb);
    if (!(nrm > 0.0)) return;
    double w1[12], w0[12];
    for (int i = 0; i < 12; ++i) {
        w1[i] = (ca * a[i] + cb * b[i]) / nrm;
        w0[i] = (-cb * a[i] + ca * b[i]) / nrm;
    }
    fix_sign(w0);
    for (int i = 0; i < 12; ++i) { a[i] = w1[i]; b[i] = w0[i]; }
}

struct Pose {
    double R[9], t[3];
};

// ---- EPnP on five points ----------------------------------------------------------------------------------------------
// X[5][3]: object points (the float32 values, widened); uv[5][2]: image points (float32, widened); intrinsics in double.
// Follows epnp.cpp function by function (choose_control_points, compute_barycentric_coordinates, fill_M, compute_L_6x10,
// compute_rho, find_betas_approx_{1,2,3}, gauss_newton, compute_R_and_t, reprojection_error).  Returns false when the
// result is not finite.
// Split in three so that the device can run the 12 x 12 eigen-decomposition warp-wide between the two serial parts
// (csrc/pnp.cu ref_epnp_kernel); epnp5() below is the plain composition the host build and the tests use.
struct EpnpState {
    double us[EP_N][2], cws[4][3], al[EP_N][4], ks[3], cut, fu, fv, uc, vc;
    bool planar;      // coplanar points: three effective control points (see epnp5_basis_planar)
};

// image points -> `us`, control points, barycentric coordinates, M^T M
VO_RHDN void epnp5_prepare(const double (&X)[EP_N][3], const double (&uv_in)[EP_N][2], double fu, double fv, double uc, double vc,
                           EpnpState &S, double (&MtM)[12][12]) {
    constexpr int n = EP_N;
    double (&us)[EP_N][2] = S.us;
    double (&cws)[4][3] = S.cws;
    double (&al)[EP_N][4] = S.al;
    double (&ks)[3] = S.ks;
    S.fu = fu; S.fv = fv; S.uc = uc; S.vc = vc;
    // solvePnP(EPNP) undistorts the image points first: normalised coordinates stored as float32 (no distortion: a pure
    // change of variables), which epnp::init_points maps back with fu, fv, uc, vc in double
    for (int i = 0; i < n; ++i) {
        const float xn = (float)((uv_in[i][0] - uc) * (1.0 / fu)), yn = (float)((uv_in[i][1] - vc) * (1.0 / fv));
        us[i][0] = (double)xn * fu + uc;
        us[i][1] = (double)yn * fv + vc;
    }
    // control points: centroid + principal directions scaled by sqrt(eigenvalue / n)
    for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += X[i][j];
        cws[0][j] = s / n;
    }
    double axes[3][3];                              // principal directions (orthonormal, sign-fixed); ks = their scales
    {
        double C[3][3], V[3][3], d[3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double s = 0.0;
                for (int i = 0; i < n; ++i) s += (X[i][a] - cws[0][a]) * (X[i][b] - cws[0][b]);
                C[a][b] = s;
            }
        jacobi_eig<3>(C, V, d);
        for (int i = 0; i < 3; ++i) {
            double v[3] = {V[0][i], V[1][i], V[2][i]};
            fix_sign(v);
            ks[i] = sqrt(fmax(d[i], 0.0) / n);
            for (int j = 0; j < 3; ++j) { axes[i][j] = v[j]; cws[i + 1][j] = cws[0][j] + ks[i] * v[j]; }
        }
    }
    // barycentric coordinates: the control vectors are k_i * (orthonormal axis i), so the inverse of [c1-c0 c2-c0 c3-c0] is
    // axis_i / k_i row by row; a vanishing k_i (planar or collinear points) gives a zero row — the pseudo-inverse
    // cvInvert(CV_SVD) returns in epnp.cpp (singular values <= 2 eps * sum are dropped, SVD::backSubst)
    const double cut = 2.0 * 2.220446049250313e-16 * (ks[0] + ks[1] + ks[2]);
    S.cut = cut;
    S.planar = !(ks[2] > cut) && ks[1] > cut;
    {
        for (int i = 0; i < n; ++i) {
            const double p[3] = {X[i][0] - cws[0][0], X[i][1] - cws[0][1], X[i][2] - cws[0][2]};
            for (int j = 0; j < 3; ++j) al[i][1 + j] = (ks[j] > cut) ? dot3(axes[j], p) / ks[j] : 0.0;
            al[i][0] = 1.0 - al[i][1] - al[i][2] - al[i][3];
        }
    }
    // M^T M (12 x 12) accumulated row by row of M
    for (int a = 0; a < 12; ++a)
        for (int b = 0; b < 12; ++b) MtM[a][b] = 0.0;
    for (int i = 0; i < n; ++i) {
        double m1[12], m2[12];
        for (int j = 0; j < 4; ++j) {
            m1[3 * j] = al[i][j] * fu; m1[3 * j + 1] = 0.0; m1[3 * j + 2] = al[i][j] * (uc - us[i][0]);
            m2[3 * j] = 0.0; m2[3 * j + 1] = al[i][j] * fv; m2[3 * j + 2] = al[i][j] * (vc - us[i][1]);
        }
        for (int a = 0; a < 12; ++a)
            for (int b = 0; b < 12; ++b) MtM[a][b] += m1[a] * m1[b] + m2[a] * m2[b];
    }
}

// v[0] = ut[11] ... v[3] = ut[8] of epnp.cpp from the eigenvectors of M^T M (columns of V, descending eigenvalues): the canonical
// basis of the two-dimensional null space, sign-fixed partners
VO_RHD void epnp5_basis_from_eig(const double (&V)[12][12], double (&v)[4][12]) {
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 12; ++k) v[i][k] = V[k][11 - i];
    canonical_null_basis(v[0], v[1]);
    fix_sign(v[2]);
    fix_sign(v[3]);
}
// Coplanar points: the fourth control point coincides with the centroid, its three columns of M vanish and e9, e10, e11 span an
// exactly-null eigenspace (OpenCV gets an arbitrary basis of it from its SVD).  Canonical choice: those unit vectors as
// v[0..2], and the weakest direction of the 9 x 9 block of the three real control points as v[3].
VO_RHDN void epnp5_basis_planar(const double (&MtM)[12][12], double (&v)[4][12]) {
    double B[9][9], V9[9][9], d9[9];
    for (int a = 0; a < 9; ++a)
        for (int b = 0; b < 9; ++b) B[a][b] = MtM[a][b];
    jacobi_eig<9>(B, V9, d9);
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 12; ++k) v[i][k] = 0.0;
    v[0][11] = 1.0; v[1][10] = 1.0; v[2][9] = 1.0;
    for (int k = 0; k < 9; ++k) v[3][k] = V9[k][8];
    fix_sign(v[3]);
}

// One of the three candidates of compute_pose (cand = 0, 1, 2: find_betas_approx_1 / _2 / _3): the distance constraints L, rho,
// the beta initialisation, five Gauss-Newton steps, absolute orientation, mean reprojection error.  Returns false when the result
// is not finite.  The candidates are independent of each other: the device gives each its own lane.
VO_RHDN bool epnp5_candidate(const double (&X)[EP_N][3], const EpnpState &S, const double (&v)[4][12], int cand, Pose &p, double &err_out) {
    constexpr int n = EP_N;
    const double (&us)[EP_N][2] = S.us;
    const double (&cws)[4][3] = S.cws;
    const double (&al)[EP_N][4] = S.al;
    const double fu = S.fu, fv = S.fv, uc = S.uc, vc = S.vc;
    // L (6 x 10) and rho
    double L[6][10], rho[6];
    {
        double dv[4][6][3];
        for (int i = 0; i < 4; ++i) {
            int a = 0, b = 1;
            for (int j = 0; j < 6; ++j) {
                for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * a + k] - v[i][3 * b + k];
                if (++b > 3) { ++a; b = a + 1; }
            }
        }
        for (int i = 0; i < 6; ++i) {
            L[i][0] = dot3(dv[0][i], dv[0][i]);
            L[i][1] = 2.0 * dot3(dv[0][i], dv[1][i]);
            L[i][2] = dot3(dv[1][i], dv[1][i]);
            L[i][3] = 2.0 * dot3(dv[0][i], dv[2][i]);
            L[i][4] = 2.0 * dot3(dv[1][i], dv[2][i]);
            L[i][5] = dot3(dv[2][i], dv[2][i]);
            L[i][6] = 2.0 * dot3(dv[0][i], dv[3][i]);
            L[i][7] = 2.0 * dot3(dv[1][i], dv[3][i]);
            L[i][8] = 2.0 * dot3(dv[2][i], dv[3][i]);
            L[i][9] = dot3(dv[3][i], dv[3][i]);
        }
        const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
        for (int i = 0; i < 6; ++i) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) { const double e = cws[pa[i]][k] - cws[pb[i]][k]; s += e * e; }
            rho[i] = s;
        }
    }
    {
        double be[4] = {0.0, 0.0, 0.0, 0.0};
        if (cand == 0) {                                     // find_betas_approx_1: B11 B12 B13 B14
            double A[6][4], b[6], x[4];
            for (int i = 0; i < 6; ++i) { A[i][0] = L[i][0]; A[i][1] = L[i][1]; A[i][2] = L[i][3]; A[i][3] = L[i][6]; b[i] = rho[i]; }
            lsq_qr<6, 4>(A, b, x);
            if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
            else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
        } else if (cand == 1) {                              // find_betas_approx_2: B11 B12 B22
            double A[6][3], b[6], x[3];
            for (int i = 0; i < 6; ++i) { A[i][0] = L[i][0]; A[i][1] = L[i][1]; A[i][2] = L[i][2]; b[i] = rho[i]; }
            lsq_qr<6, 3>(A, b, x);
            if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0; }
            else { be[0] = sqrt(x[0]); be[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0; }
            if (x[1] < 0) be[0] = -be[0];
        } else {                                             // find_betas_approx_3: B11 B12 B22 B13 B23
            double A[6][5], b[6], x[5];
            for (int i = 0; i < 6; ++i) { for (int k = 0; k < 5; ++k) A[i][k] = L[i][k]; b[i] = rho[i]; }
            lsq_qr<6, 5>(A, b, x);
            if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0; }
            else { be[0] = sqrt(x[0]); be[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0; }
            if (x[1] < 0) be[0] = -be[0];
            be[2] = x[3] / be[0];
        }
        for (int it = 0; it < 5; ++it) {                     // gauss_newton on the four betas
            double A[6][4], b[6], x[4];
            for (int i = 0; i < 6; ++i) {
                const double *r = L[i];
                A[i][0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3];
                A[i][1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3];
                A[i][2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3];
                A[i][3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3];
                b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] + r[3] * be[0] * be[2] +
                                 r[4] * be[1] * be[2] + r[5] * be[2] * be[2] + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] +
                                 r[8] * be[2] * be[3] + r[9] * be[3] * be[3]);
            }
            lsq_qr<6, 4>(A, b, x);
            for (int k = 0; k < 4; ++k) be[k] += x[k];
        }
        // compute_R_and_t: control points and points in the camera frame, sign, absolute orientation
        double ccs[4][3], pcs[n][3];
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 3; ++k) ccs[j][k] = be[0] * v[0][3 * j + k] + be[1] * v[1][3 * j + k] + be[2] * v[2][3 * j + k] + be[3] * v[3][3 * j + k];
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) pcs[i][k] = al[i][0] * ccs[0][k] + al[i][1] * ccs[1][k] + al[i][2] * ccs[2][k] + al[i][3] * ccs[3][k];
        if (pcs[0][2] < 0.0)
            for (int i = 0; i < n; ++i)
                for (int k = 0; k < 3; ++k) pcs[i][k] = -pcs[i][k];
        double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 3; ++k) { pc0[k] += pcs[i][k]; pw0[k] += X[i][k]; }
        for (int k = 0; k < 3; ++k) { pc0[k] /= n; pw0[k] /= n; }
        double S[3][3];                                      // S[a][b] = sum pw_a pc_b  (Horn: world -> camera)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double s = 0.0;
                for (int i = 0; i < n; ++i) s += (X[i][a] - pw0[a]) * (pcs[i][b] - pc0[b]);
                S[a][b] = s;
            }
        double Nq[4][4] = {{S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
                           {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
                           {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
                           {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
        double Vq[4][4], dq[4];
        jacobi_eig<4>(Nq, Vq, dq);
        const double qw = Vq[0][0], qx = Vq[1][0], qy = Vq[2][0], qz = Vq[3][0];
        p.R[0] = qw * qw + qx * qx - qy * qy - qz * qz; p.R[1] = 2 * (qx * qy - qw * qz); p.R[2] = 2 * (qx * qz + qw * qy);
        p.R[3] = 2 * (qx * qy + qw * qz); p.R[4] = qw * qw - qx * qx + qy * qy - qz * qz; p.R[5] = 2 * (qy * qz - qw * qx);
        p.R[6] = 2 * (qx * qz - qw * qy); p.R[7] = 2 * (qy * qz + qw * qx); p.R[8] = qw * qw - qx * qx - qy * qy + qz * qz;
        for (int k = 0; k < 3; ++k) p.t[k] = pc0[k] - dot3(p.R + 3 * k, pw0);
        double err = 0.0;                                    // reprojection_error: mean pixel distance over the five points
        for (int i = 0; i < n; ++i) {
            const double Xc = dot3(p.R, X[i]) + p.t[0], Yc = dot3(p.R + 3, X[i]) + p.t[1], iz = 1.0 / (dot3(p.R + 6, X[i]) + p.t[2]);
            const double du = us[i][0] - (uc + fu * Xc * iz), dw = us[i][1] - (vc + fv * Yc * iz);
            err += sqrt(du * du + dw * dw);
        }
        err /= n;
        bool finite = err == err && err < 1e300;
        for (int k = 0; k < 9; ++k) finite = finite && p.R[k] == p.R[k];
        for (int k = 0; k < 3; ++k) finite = finite && p.t[k] == p.t[k] && fabs(p.t[k]) < 1e300;
        err_out = err;
        return finite;
    }
}

// the candidate with the lowest mean reprojection error (strict <: the earlier candidate keeps ties)
VO_RHDN bool epnp5_finish(const double (&X)[EP_N][3], const EpnpState &S, const double (&v)[4][12], Pose &out) {
    double best_err = 0.0;
    bool have = false;
    for (int cand = 0; cand < 3; ++cand) {
        Pose p;
        double err;
        if (epnp5_candidate(X, S, v, cand, p, err) && (!have || err < best_err)) { best_err = err; out = p; have = true; }
    }
    return have;
}

VO_RHDN bool epnp5(const double (&X)[EP_N][3], const double (&uv_in)[EP_N][2], double fu, double fv, double uc, double vc, Pose &out) {
    EpnpState S;
    double MtM[12][12], v[4][12];
    epnp5_prepare(X, uv_in, fu, fv, uc, vc, S, MtM);
    if (S.planar) {
        epnp5_basis_planar(MtM, v);
    } else {
        double V[12][12], d[12];
        jacobi_eig12_rr(MtM, V, d);
        epnp5_basis_from_eig(V, v);
    }
    return epnp5_finish(X, S, v, out);
}

// ---- inlier rule of PnPRansacCallback::computeError + RANSACPointSetRegistrator::findInliers ----------------------------
// cv2.projectPoints in double (no distortion), stored as float32; squared distance to the float32 image point in float32.
VO_RHD float reproj_err2(const Pose &p, double fx, double fy, double cx, double cy, float Xf, float Yf, float Zf, float uf, float vf) {
    const double X = Xf, Y = Yf, Z = Zf;
    const double x = p.R[0] * X + p.R[1] * Y + p.R[2] * Z + p.t[0];
    const double y = p.R[3] * X + p.R[4] * Y + p.R[5] * Z + p.t[1];
    double z = p.R[6] * X + p.R[7] * Y + p.R[8] * Z + p.t[2];
    z = z ? 1.0 / z : 1.0;
    const float pu = (float)(x * z * fx + cx), pv = (float)(y * z * fy + cy);
    const float du = uf - pu, dv = vf - pv;
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
#else
    return du * du + dv * dv;
#endif
}

// ---- the iteration loop of RANSACPointSetRegistrator::run, replayed over the counts of all `iters` hypotheses -----------
// counts[h] < 0 marks a hypothesis whose minimal solve failed (no model: the iteration is spent, nothing is compared).
// Returns the index of the model OpenCV would return (-1: none) and the iterations it would have run.
VO_RHD int ransac_scan(const int32_t *counts, int n, int iters, double confidence, int *iters_run, int *best_count) {
    int niters = iters, best = -1, max_good = 0, it = 0;
    for (; it < niters; ++it) {
        const int good = counts[it];
        if (good < 0) continue;
        if (good > (max_good > EP_N - 1 ? max_good : EP_N - 1)) {
            max_good = good;
            best = it;
            // RANSACUpdateNumIters(confidence, (n - good) / n, model points, niters)
            double ep = (double)(n - good) / (double)n;
            ep = ep < 0.0 ? 0.0 : (ep > 1.0 ? 1.0 : ep);
            double num = 1.0 - confidence;
            num = num > 2.2250738585072014e-308 ? num : 2.2250738585072014e-308;
            double denom = 1.0 - pow(1.0 - ep, (double)EP_N);
            if (denom < 2.2250738585072014e-308) { niters = 0; }
            else {
                num = log(num);
                denom = log(denom);
                if (!(denom >= 0.0 || -num >= niters * (-denom))) niters = (int)rint(num / denom);
            }
        }
    }
    *iters_run = it;
    *best_count = max_good;
    return max_good > 0 ? best : -1;
}

}  // namespace refpnp
}  // namespace vo
