"""KITTI odometry evaluator with the interface of the reference's plot_utils/kittievalodom.py (class KittiEvalOdom,
same method names, arguments and return values), written for 100k-pose sequences: the per-pose Python loops
(`np.linalg.inv` + `@` per frame in compute_RPE :429-469, the linear scan of last_frame_from_segment_length :166-179
inside the double loop of calc_sequence_errors :181-233) become batched 4x4 algebra and one `searchsorted`.

Definitions kept exactly (they are what the reference reports, quirks included):
  * translation error ignores y (:152-164); rotation error = arccos of the clamped trace formula (:138-150);
  * ATE is a SUM of x-z distances, not an RMSE (:392-427); compute_RPE returns (mean relative translation error,
    SUM of rotation errors in degrees, path length) (:429-469);
  * segment lengths 100..800 m, first frames every 10 poses, speed assumes 10 FPS (:181-233);
  * eval() re-anchors both trajectories to the first predicted frame and, for alignment in {"scale_7dof", "7dof",
    "6dof"}, collects the positions but applies NO alignment (:543-551); only "scale" rescales (:541-542).
Plots need matplotlib, imported lazily by the two plotting methods only.
"""
import copy

import numpy as np
import yaml


def scale_lse_solver(X, Y):
    """Scale s minimising |s X - Y| (:14-25)."""
    return np.sum(X * Y) / np.sum(X ** 2)


def umeyama_alignment(x, y, with_scale=False):
    """Similarity transform (r, t, c) minimising sum |y_i - (c r x_i + t)|^2 over two m x n point sets (columns are points) —
    the closed form of Umeyama (1991), interface of the reference's helper (plot_utils/kittievalodom.py:28-77).  Written from
    the paper: cross-covariance of the centred sets, its SVD, the reflection guard on the last singular direction, scale from
    the guarded singular values over the variance of x."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    if x.shape != y.shape:
        raise AssertionError("x.shape not equal to y.shape")
    dim, count = x.shape
    cx, cy = x.mean(axis=1, keepdims=True), y.mean(axis=1, keepdims=True)
    dx, dy = x - cx, y - cy
    cross = dy @ dx.T / count                                 # E[(y - cy)(x - cx)^T]
    left, sing, right_t = np.linalg.svd(cross)
    guard = np.ones(dim)
    if np.linalg.det(left) * np.linalg.det(right_t) < 0.0:    # a reflection would fit better: flip the weakest direction
        guard[-1] = -1.0
    r = (left * guard) @ right_t
    c = float((sing * guard).sum() / ((dx * dx).sum() / count)) if with_scale else 1.0
    t = (cy - c * (r @ cx)).ravel()
    return r, t, c


def _stack(poses, keys):
    return np.stack([poses[k] for k in keys]) if len(keys) else np.zeros((0, 4, 4))


def _rot_err(E):
    d = 0.5 * (E[..., 0, 0] + E[..., 1, 1] + E[..., 2, 2] - 1.0)
    return np.arccos(np.clip(d, -1.0, 1.0))


def _trans_err(E):
    return np.sqrt(E[..., 0, 3] ** 2 + E[..., 2, 3] ** 2)


class KittiEvalOdom():
    """Evaluate an odometry result.  vo_eval = KittiEvalOdom(); vo_eval.eval(result_dir, flag, alignment="6dof")."""

    def __init__(self):
        self.lengths = [100, 200, 300, 400, 500, 600, 700, 800]
        self.num_lengths = len(self.lengths)
        self.step_size = 10

    # ------------------------------------------------------------------ I/O
    def load_poses_from_txt(self, file_name):
        """{idx: 4x4} from KITTI text: 12 numbers per line, 13 with a leading frame index, or 16 (4x4) (:85-116)."""
        poses = {}
        with open(file_name, "r") as f:
            lines = f.readlines()
        for cnt, line in enumerate(lines):
            vals = [float(i) for i in line.split(" ") if i != ""]
            with_idx = len(vals) == 13
            P = np.eye(4)
            P[:3, :] = np.asarray(vals[with_idx:with_idx + 12]).reshape(3, 4)
            poses[vals[0] if with_idx else cnt] = P
        return poses

    # ------------------------------------------------------------------ per-pose metrics
    def trajectory_distances(self, poses):
        """Path length from frame 0 to every pose (:118-136)."""
        keys = sorted(poses.keys())
        xyz = _stack(poses, keys)[:, :3, 3]
        d = xyz[:-1] - xyz[1:]
        step = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2)
        return [0] + list(np.cumsum(step))

    def rotation_error(self, pose_error):
        return float(_rot_err(np.asarray(pose_error)))

    def translation_error(self, pose_error):
        return float(_trans_err(np.asarray(pose_error)))

    def last_frame_from_segment_length(self, dist, first_frame, length):
        """First index i >= first_frame with dist[i] > dist[first_frame] + length, or -1 (:166-179).  `dist` is
        non-decreasing, so this is one binary search."""
        dist = np.asarray(dist, dtype=np.float64)
        i = int(np.searchsorted(dist, dist[first_frame] + length, side="right"))
        i = max(i, first_frame)
        return i if i < len(dist) else -1

    def calc_sequence_errors(self, poses_gt, poses_result):
        """[[first_frame, r_err/len, t_err/len, len, speed], ...] over all (first frame, segment length) (:181-233)."""
        dist = np.asarray(self.trajectory_distances(poses_gt), dtype=np.float64)
        n = len(poses_gt)
        self.step_size = 10
        firsts = np.arange(0, n, self.step_size)
        lengths = np.asarray(self.lengths, dtype=np.float64)
        if len(firsts) == 0:
            return []
        last = np.searchsorted(dist, dist[firsts][:, None] + lengths[None, :], side="right")     # (F, L)
        last = np.maximum(last, firsts[:, None])
        ff = np.repeat(firsts, len(lengths))
        ll = np.tile(lengths, len(firsts))
        lf = last.reshape(-1)
        ok = lf < n
        res_keys = set(poses_result.keys())
        ok &= np.array([(int(a) in res_keys) and (int(b) in res_keys) if o else False for a, b, o in zip(ff, lf, ok)], dtype=bool)
        ff, ll, lf = ff[ok], ll[ok], lf[ok]
        if len(ff) == 0:
            return []
        need = sorted(set(ff.tolist()) | set(lf.tolist()))
        pos = {k: i for i, k in enumerate(need)}
        G = _stack(poses_gt, need)
        R = _stack(poses_result, need)
        a = np.array([pos[int(k)] for k in ff]); b = np.array([pos[int(k)] for k in lf])
        delta_gt = np.linalg.inv(G[a]) @ G[b]
        delta_res = np.linalg.inv(R[a]) @ R[b]
        E = np.linalg.inv(delta_res) @ delta_gt
        r_err, t_err = _rot_err(E), _trans_err(E)
        speed = ll / (0.1 * (lf - ff + 1.0))
        return [[int(f), float(r / l), float(t / l), (int(l) if float(l).is_integer() else float(l)), float(s)]
                for f, r, t, l, s in zip(ff, r_err, t_err, ll, speed)]

    def save_sequence_errors(self, err, file_name):
        with open(file_name, "w") as fp:
            fp.writelines(" ".join(str(j) for j in i) + "\n" for i in err)

    def compute_overall_err(self, seq_err):
        """(average translation error, average rotation error) over all segments (:247-270)."""
        if len(seq_err) == 0:
            return 0, 0
        e = np.asarray([[i[1], i[2]] for i in seq_err], dtype=np.float64)
        return float(np.sum(e[:, 1]) / len(e)), float(np.sum(e[:, 0]) / len(e))

    def compute_segment_error(self, seq_errs):
        """{length: [avg_t_err, avg_r_err]} (empty list for a length without segments) (:361-390)."""
        out = {}
        e = np.asarray([[i[3], i[2], i[1]] for i in seq_errs], dtype=np.float64).reshape(-1, 3)
        for len_ in self.lengths:
            sel = e[e[:, 0] == len_]
            out[len_] = [float(np.mean(sel[:, 1])), float(np.mean(sel[:, 2]))] if len(sel) else []
        return out

    def compute_ATE(self, gt, pred):
        """SUM over predicted frames of the x-z distance between ground truth and prediction (:392-427)."""
        keys = list(pred.keys())
        g = _stack(gt, keys)[:, [0, 2], 3]
        p = _stack(pred, keys)[:, [0, 2], 3]
        return float(np.sum(np.linalg.norm(g - p, axis=1)))

    def compute_RPE(self, gt, pred):
        """(mean relative translation error, SUM of rotation errors in degrees, path length) (:429-469)."""
        keys = list(pred.keys())[:-1]
        if not keys:
            return float("nan"), 0.0, 0
        nxt = [k + 1 for k in keys]
        g_rel = np.linalg.inv(_stack(gt, keys)) @ _stack(gt, nxt)
        p_rel = np.linalg.inv(_stack(pred, keys)) @ _stack(pred, nxt)
        E = np.linalg.inv(g_rel) @ p_rel
        local = np.linalg.norm(g_rel[:, :3, 3], axis=1)
        trans = _trans_err(E) / local
        rot = _rot_err(E)
        total = 0
        for v in local:           # the reference accumulates left to right (:461); keep its rounding
            total += v
        return float(np.mean(trans)), float(np.sum(rot) * 180 / np.pi), float(total)

    def scale_optimization(self, gt, pred):
        """Rescale the predicted translations by the least-squares factor (:471-493)."""
        pred_updated = copy.deepcopy(pred)
        keys = list(pred.keys())
        scale = scale_lse_solver(_stack(pred, keys)[:, :3, 3], _stack(gt, keys)[:, :3, 3])
        for i in pred_updated:
            pred_updated[i][:3, 3] *= scale
        return pred_updated

    def write_result(self, f, seq, errs):
        ave_t_err, ave_r_err, ate, rpe_trans, rpe_rot = errs
        f.writelines(["Sequence: \t {} \n".format(seq),
                      "Trans. err. (%): \t {:.3f} \n".format(ave_t_err * 100),
                      "Rot. err. (deg/100m): \t {:.3f} \n".format(ave_r_err / np.pi * 180 * 100),
                      "ATE (m): \t {:.3f} \n".format(ate),
                      "RPE (m): \t {:.3f} \n".format(rpe_trans),
                      "RPE (deg): \t {:.3f} \n\n".format(rpe_rot * 180 / np.pi)])

    # ------------------------------------------------------------------ plots (matplotlib imported on demand)
    def plot_trajectory(self, poses_gt, poses_result, seq):
        from matplotlib import pyplot as plt
        fig = plt.figure()
        plt.gca().set_aspect("equal")
        keys = sorted(poses_result.keys())
        for label, poses in (("Ground Truth", poses_gt), ("Ours", poses_result)):
            xz = _stack(poses, keys)[:, [0, 2], 3]
            plt.plot(xz[:, 0], xz[:, 1], label=label)
        plt.legend(loc="upper right", prop={"size": 20})
        plt.xlabel("x (m)", fontsize=20)
        plt.ylabel("z (m)", fontsize=20)
        fig.set_size_inches(10, 10)
        plt.savefig(self.plot_path_dir + "/sequence_{:02}.pdf".format(seq), bbox_inches="tight", pad_inches=0)
        plt.close(fig)

    def plot_error(self, avg_segment_errs, seq):
        from matplotlib import pyplot as plt
        for col, scale, ylabel, name in ((0, 100.0, "Translation Error (%)", "trans_err"),
                                         (1, 100.0 * 180.0 / np.pi, "Rotation Error (deg/100m)", "rot_err")):
            y = [avg_segment_errs[l][col] * scale if len(avg_segment_errs[l]) > 0 else 0 for l in self.lengths]
            fig = plt.figure()
            plt.plot(self.lengths, y, "bs-", label=ylabel.split(" (")[0])
            plt.ylabel(ylabel, fontsize=10)
            plt.xlabel("Path Length (m)", fontsize=10)
            plt.legend(loc="upper right", prop={"size": 10})
            fig.set_size_inches(5, 5)
            plt.savefig(self.plot_error_dir + "/{}_{:02}.pdf".format(name, seq), bbox_inches="tight", pad_inches=0)
            plt.close(fig)

    # ------------------------------------------------------------------ the reference's entry point
    def eval_poses(self, poses_gt, poses_result, alignment=None):
        """eval() on in-memory {idx: 4x4} dictionaries (what eval() does after reading its two files)."""
        poses_gt, poses_result = dict(poses_gt), dict(poses_result)
        idx_0 = sorted(list(poses_result.keys()))[0]
        inv_p0, inv_g0 = np.linalg.inv(poses_result[idx_0]), np.linalg.inv(poses_gt[idx_0])
        for cnt in poses_result:
            poses_result[cnt] = inv_p0 @ poses_result[cnt]
            poses_gt[cnt] = inv_g0 @ poses_gt[cnt]
        if alignment == "scale":
            poses_result = self.scale_optimization(poses_gt, poses_result)
        # "scale_7dof" / "7dof" / "6dof": the reference gathers the positions and applies nothing (:543-551)
        seq_err = self.calc_sequence_errors(poses_gt, poses_result)
        self.avg_segment_errs = self.compute_segment_error(seq_err)
        self.ave_t_err, self.ave_r_err = self.compute_overall_err(seq_err)
        ate_error = self.compute_ATE(poses_gt, poses_result)
        rep_error, rot_error, total_distance = self.compute_RPE(poses_gt, poses_result)
        return (ate_error / total_distance, rep_error, rot_error / total_distance, total_distance)

    def eval(self, result_dir, flag, alignment=None, seqs=None):
        with open("../config/vo_params.yaml") as f:
            vo_params = yaml.load(f, Loader=yaml.FullLoader)
        poses_result = self.load_poses_from_txt(vo_params["poses_file_path"] + ".txt")
        gt_filename = vo_params["gt_txt_file_path"]
        poses_gt = self.load_poses_from_txt(gt_filename.split(".txt")[0] + "_modified.txt")
        self.result_file_name = result_dir + "resultall.txt"
        return self.eval_poses(poses_gt, poses_result, alignment)


if __name__ == "__main__":
    keo = KittiEvalOdom()
    print(keo.eval("resdir", 1, alignment="6dof"))
