"""Offline RGB-D driver — interface of the reference's vo_stereo_runner.py: `vo_offline_data(cam_intr, img_path,
output_filename)` (:27-60) walks sorted `*.png` / `*_depth.npy` pairs, feeds VisualOdometry.process_frame and saves
the (N,4,4) float64 global poses with np.save."""
import glob
import time

import cv2
import numpy as np

from VisualOdometry_Stereo import VisualOdometry

midpoints = [(100, 100)]
numpyseeds = [8214]
for _seed in numpyseeds:           # kept for parity of the global numpy stream (vo_stereo_runner.py:20-24)
    np.random.seed(_seed)


def vo_offline_data(cam_intr, img_path, output_filename):
    vo = VisualOdometry(cam_intr, seq=0)
    images = sorted(glob.glob(img_path + "/*.png"))
    depths = sorted(glob.glob(img_path + "/*_depth.npy"))
    poses = []
    t_start = time.time()
    for index, (image_file, depth_file) in enumerate(zip(images, depths)):
        if index == 1:
            t_start = time.time()
        frame = cv2.imread(image_file)
        if frame is None:
            raise FileNotFoundError(image_file)
        depth = np.load(depth_file)
        frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
        pose = vo.process_frame(frame, depth, midpoints[0], index)
        print("Time taken:" + str(time.time() - t_start))
        print("frame_pose.t.T" + str(pose.t.T))
        poses.append(pose.pose)
    if poses:
        print("Average time per frame:", (time.time() - t_start) / len(poses))
    np.save(output_filename, np.asarray(poses))
    return np.asarray(poses)
