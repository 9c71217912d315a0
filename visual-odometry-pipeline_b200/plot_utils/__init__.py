"""Trajectory text formats and the KITTI odometry evaluator with the reference's interface (plot_utils/)."""
