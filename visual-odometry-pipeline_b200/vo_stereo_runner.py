"""Offline RGB-D driver — interface of the reference's vo_stereo_runner.py: `vo_offline_data(cam_intr, img_path,
output_filename)` (:27-60) walks sorted `*.png` / `*_depth.npy` pairs, feeds VisualOdometry.process_frame and saves
the (N,4,4) float64 global poses with np.save."""
import time

import numpy as np

from VisualOdometry_Stereo import VisualOdometry
from vo_b200.frame_io import FramePrefetcher

midpoints = [(100, 100)]
numpyseeds = [8214]
for _seed in numpyseeds:           # kept for parity of the global numpy stream (vo_stereo_runner.py:20-24)
    np.random.seed(_seed)


def vo_offline_data(cam_intr, img_path, output_filename):
    vo = VisualOdometry(cam_intr, seq=0)
    poses = []
    t_start = time.time()
    # same files, same order, same BGR->RGB conversion as the reference loop (:38-50); decoding runs ahead in threads
    for index, frame, depth in FramePrefetcher(img_path):
        if index == 1:
            t_start = time.time()
        pose = vo.process_frame(frame, depth, midpoints[0], index)
        print("Time taken:" + str(time.time() - t_start))
        print("frame_pose.t.T" + str(pose.t.T))
        poses.append(pose.pose)
    if poses:
        print("Average time per frame:", (time.time() - t_start) / len(poses))
    np.save(output_filename, np.asarray(poses))
    return np.asarray(poses)
