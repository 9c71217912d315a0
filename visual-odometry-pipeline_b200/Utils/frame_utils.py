"""Frame / FramePair containers with the field names of the reference's Utils/frame_utils.py (Frame :31-58,
FramePair :3-29).  No arithmetic here.  Unlike the reference, default poses are created per instance (the
reference shares one mutable `SE3()` default between all instances, frame_utils.py:4,32)."""
import numpy as np

from Utils.SE3_utils import SE3


class Frame:
    def __init__(self, id, img, kps, desc, fil, pose=None, seg_img=None, depth=None, image_arr=None):
        self.id = id
        self.image = img
        self.keypoints = kps
        self.descriptors = desc
        self.filename = fil
        self.depth = np.zeros(1) if depth is None else depth
        self.pose = SE3() if pose is None else pose
        self.image_arr = image_arr
        self.seg_img = seg_img
        self.glob_pose = None
        self.global_pose = None
        self.tracked_kps = None
        self.kps_index = np.arange(len(kps))
        self.device_cache = {}  # device-resident copies (depth map) kept while the frame is a keyframe

    def getitems(self):
        return self.image, self.keypoints, self.descriptors, self.filename

    def get_kp_desc(self):
        return self.keypoints, self.descriptors

    def get_image(self):
        return self.image

    def get_file(self):
        return self.filename


class FramePair:
    def __init__(self, f1, f2, matches_no, left_kp, right_kp, frame1_idx=None, cheirality_pts_ct=0, inlier_pts_ct=0,
                 pose=None):
        self.frame1, self.frame2 = f1, f2
        self.left_kp, self.right_kp = left_kp, right_kp
        self.matches_no = matches_no
        self.frame_index = frame1_idx
        self.cheirality_pts_ct = cheirality_pts_ct
        self.inlier_pts_ct = inlier_pts_ct
        self.pose = SE3() if pose is None else pose
        self.avg_optical_flow = 0
        self.ess_mat = None

    def getpose(self):
        out = SE3()
        out.t = self.pose.t.copy()
        out.R = self.pose.R.copy()
        return out

    def getkeypts(self):
        return self.left_kp, self.right_kp

    def getchecks(self):
        return self.cheirality_pts_ct, self.inlier_pts_ct
