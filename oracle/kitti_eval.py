"""Trajectory acceptance metrics — TEST INFRASTRUCTURE ONLY (see oracle.py header).

numpy restatement of the reference evaluator's eval() (plot_utils/kittievalodom.py:513-570): re-anchor both
trajectories to frame 0 (:534-539; the '6dof' branch applies no alignment, :543-551), then
  ATE  = SUM_i || (x,z)_gt - (x,z)_pred ||          (compute_ATE :392-427 — a sum, not an RMSE, :424)
  RPE  = mean_i trans_err(rel_err_i) / ||t_gt_rel||  with translation error ignoring y (:152-164, :429-469)
  rot  = SUM_i rotation_error(rel_err_i) in degrees  (:138-150, :469)
and returns (ATE/dist, RPE_trans, rot/dist, dist) exactly like :570.
Pinned by tests/golden/kitti03_eval.npz (the reference evaluator run on its own shipped KITTI-03 data).
"""
import numpy as np


def _to4(P12):
    T = np.tile(np.eye(4), (P12.shape[0], 1, 1))
    T[:, :3, :] = P12.reshape(-1, 3, 4)
    return T


def rotation_error(E):
    d = 0.5 * (E[0, 0] + E[1, 1] + E[2, 2] - 1.0)
    return float(np.arccos(max(min(d, 1.0), -1.0)))


def translation_error(E):
    return float(np.sqrt(E[0, 3] ** 2 + E[2, 3] ** 2))  # y ignored, kittievalodom.py:161


def evaluate(gt, pred):
    """gt, pred: (N,12) KITTI rows, (N,4,4) or (N,16).  Returns the reference's 4-tuple."""
    gt, pred = np.asarray(gt, np.float64), np.asarray(pred, np.float64)
    gt = _to4(gt.reshape(len(gt), -1)[:, :12]) if gt.ndim == 2 else gt
    pred = _to4(pred.reshape(len(pred), -1)[:, :12]) if pred.ndim == 2 else pred
    n = len(pred)
    gt = np.linalg.inv(gt[0]) @ gt[:n]
    pred = np.linalg.inv(pred[0]) @ pred
    ate = float(np.sum(np.linalg.norm(gt[:, [0, 2], 3] - pred[:, [0, 2], 3], axis=1)))
    trans, rot, dist = [], [], 0.0
    for i in range(n - 1):
        g = np.linalg.inv(gt[i]) @ gt[i + 1]
        p = np.linalg.inv(pred[i]) @ pred[i + 1]
        e = np.linalg.inv(g) @ p
        ld = np.linalg.norm(g[:3, 3])
        dist += ld
        trans.append(translation_error(e) / ld)
        rot.append(rotation_error(e))
    return ate / dist, float(np.mean(trans)), float(np.sum(rot) * 180 / np.pi) / dist, dist


def sequence_errors(gt, pred, lengths=(100, 200, 300, 400, 500, 600, 700, 800), step=10):
    """Plain-loop restatement of calc_sequence_errors (plot_utils/kittievalodom.py:181-233) with its helpers
    trajectory_distances (:118-136) and last_frame_from_segment_length (:166-179): for every 10th first frame and
    every segment length, the relative pose error between first frame and the first frame further than `length`
    along the ground-truth path.  gt / pred: (N,4,4).  Returns rows [first, r_err/len, t_err/len, len, speed]."""
    gt, pred = np.asarray(gt, np.float64), np.asarray(pred, np.float64)
    dist = [0.0]
    for i in range(len(gt) - 1):
        d = gt[i, :3, 3] - gt[i + 1, :3, 3]
        dist.append(dist[i] + np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2))
    err = []
    for first in range(0, len(gt), step):
        for ln in lengths:
            last = -1
            for i in range(first, len(dist)):
                if dist[i] > dist[first] + ln:
                    last = i
                    break
            if last == -1 or last >= len(pred) or first >= len(pred):
                continue
            dg = np.dot(np.linalg.inv(gt[first]), gt[last])
            dr = np.dot(np.linalg.inv(pred[first]), pred[last])
            e = np.dot(np.linalg.inv(dr), dg)
            err.append([first, rotation_error(e) / ln, translation_error(e) / ln, ln, ln / (0.1 * (last - first + 1.0))])
    return err
