"""4x4 rigid-pose wrapper with the attribute surface of the reference's Utils/SE3_utils.py:5-51
(`pose`, `inv_pose`, `R`, `t`, plus the `_pose` array the VO loop touches directly).  Host-side fp64."""
import numpy as np


class SE3:
    __slots__ = ("_pose",)

    def __init__(self, np_arr=None):
        self._pose = np.eye(4) if np_arr is None else np_arr

    # full matrix -----------------------------------------------------------------------------
    def _get_pose(self):
        return self._pose

    def _set_pose(self, value):
        self._pose = value

    pose = property(_get_pose, _set_pose, doc="4x4 camera pose")

    # inverse (setting it stores the inverse of the given matrix) --------------------------------
    def _get_inv(self):
        return np.linalg.inv(self._pose)

    def _set_inv(self, value):
        self._pose = np.linalg.inv(value)

    inv_pose = property(_get_inv, _set_inv, doc="inverse of the 4x4 pose")

    # blocks: views into _pose, so in-place edits propagate as in the reference ------------------
    def _get_R(self):
        return self._pose[:3, :3]

    def _set_R(self, value):
        self._pose[:3, :3] = value

    R = property(_get_R, _set_R, doc="3x3 rotation block")

    def _get_t(self):
        return self._pose[:3, 3:]

    def _set_t(self, value):
        self._pose[:3, 3:] = value

    t = property(_get_t, _set_t, doc="3x1 translation block")

    def __repr__(self):
        return f"SE3(t={self._pose[:3, 3]})"
