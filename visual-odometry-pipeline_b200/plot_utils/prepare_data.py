"""Pose file formats (interface of the reference's plot_utils/prepare_data.py).

  * `prepare_data(file_name)`          (:8-15)  (N,4,4) .npy of global poses -> `<file_name>.txt`, one pose per line
                                                 as 16 floats in numpy's `savetxt` format ('%.18e', space separated);
  * `prepare_kitti_gt_data(gt_file)`   (:18-27) KITTI ground truth (12 floats per line) -> `<stem>_modified.txt` with
                                                 the homogeneous row "0.00 0.00 0.00 1.00" appended to every line.

Written for long sequences: one vectorised `savetxt` / one join instead of a Python loop per pose.  Plus the readers
and writers the throughput mode needs (`load_kitti_poses`, `write_kitti_poses`).
"""
import numpy as np
import yaml


def prepare_data(file_name):
    data = np.load(file_name)
    flat = np.asarray(data, dtype=np.float64).reshape(len(data), 16)
    np.savetxt(file_name + ".txt", flat)          # identical bytes to the reference's per-pose savetxt calls
    print("Data processed")


def prepare_kitti_gt_data(gt_file):
    with open(gt_file, "r") as f:
        lines = f.read().split("\n")
    op_filename = gt_file.split(".txt")[0] + "_modified.txt"
    tail = " 0.00 0.00 0.00 1.00\n"
    with open(op_filename, "w") as f:
        f.write("".join(line + tail for line in lines))
    print("Data processed")


def load_kitti_poses(file_name):
    """KITTI pose text -> (N,4,4) float64.  Accepts 12 floats (3x4), 16 floats (4x4) or 13 (index + 3x4) per line;
    blank lines are skipped.  Returns (poses, indices)."""
    rows, idx = [], []
    with open(file_name, "r") as f:
        for cnt, line in enumerate(f):
            vals = line.split()
            if not vals:
                continue
            v = np.array(vals, dtype=np.float64)
            if v.size == 13:
                idx.append(int(v[0]))
                v = v[1:]
            else:
                idx.append(len(idx))
            if v.size < 12:
                raise ValueError(f"{file_name}:{cnt + 1}: expected 12, 13 or 16 numbers, got {v.size}")
            rows.append(v[:12])
    P = np.tile(np.eye(4), (len(rows), 1, 1))
    if rows:
        P[:, :3, :] = np.asarray(rows).reshape(-1, 3, 4)
    return P, np.asarray(idx, dtype=np.int64)


def write_kitti_poses(file_name, poses, homogeneous_row=False):
    """(N,4,4) poses -> KITTI text, 12 floats per line (or 16 with the homogeneous row)."""
    P = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    flat = P.reshape(len(P), 16) if homogeneous_row else P[:, :3, :].reshape(len(P), 12)
    np.savetxt(file_name, flat)


if __name__ == "__main__":
    with open("../config/vo_params.yaml") as f:
        vo_params = yaml.load(f, Loader=yaml.FullLoader)
    prepare_kitti_gt_data(vo_params["gt_txt_file_path"])
    prepare_data(vo_params["poses_file_path"])
