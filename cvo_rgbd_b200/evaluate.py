"""Trajectory evaluation after the hot path: absolute trajectory error and relative pose
error against a TUM ground-truth file, with the semantics of the reference's vendored TUM tools
(data/rgbd_dataset/rgbd_benchmark_tools/associate.py:72-108, evaluate_ate.py:47-79,
evaluate_rpe.py:204-297).  Vectorised NumPy; host-side only."""
import numpy as np


def associate(stamps_a, stamps_b, offset=0.0, max_difference=0.02):
    """Greedy closest-timestamp matching: every (a, b) with |a - (b + offset)| < max_difference is a candidate;
    candidates are taken in order of increasing difference, each stamp used at most once; result sorted by a."""
    a = np.asarray(sorted(stamps_a), np.float64)
    b = np.asarray(sorted(stamps_b), np.float64)
    cand = []
    j0 = 0
    for i, ta in enumerate(a):
        while j0 < len(b) and b[j0] + offset <= ta - max_difference:
            j0 += 1
        j = j0
        while j < len(b) and b[j] + offset < ta + max_difference:
            cand.append((abs(ta - (b[j] + offset)), i, j))
            j += 1
    cand.sort()
    used_a, used_b, out = set(), set(), []
    for _, i, j in cand:
        if i not in used_a and j not in used_b:
            used_a.add(i)
            used_b.add(j)
            out.append((a[i], b[j]))
    out.sort()
    return out


def horn_align(model, data):
    """Rigid (rotation + translation, no scale) least-squares alignment of 3 x n `model` onto `data` (Horn / Kabsch).
    Returns (R, t, per-point translational error)."""
    model, data = np.asarray(model, np.float64), np.asarray(data, np.float64)
    mc, dc = model.mean(1, keepdims=True), data.mean(1, keepdims=True)
    W = (model - mc) @ (data - dc).T
    U, _, Vt = np.linalg.svd(W.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    t = dc - R @ mc
    err = np.linalg.norm(R @ model + t - data, axis=0)
    return R, t, err


def absolute_trajectory_error(traj_gt, traj_est, offset=0.0, max_difference=0.02, scale=1.0):
    """ATE statistics (rmse, mean, median, std, min, max, pairs) after Horn alignment of the matched positions."""
    matches = associate(traj_gt.keys(), traj_est.keys(), offset, max_difference)
    if len(matches) < 2:
        raise ValueError("fewer than 2 matching timestamps between ground truth and estimate")
    gt = np.array([traj_gt[a][:3, 3] for a, _ in matches]).T
    est = np.array([traj_est[b][:3, 3] * scale for _, b in matches]).T
    _, _, err = horn_align(est, gt)
    return dict(pairs=len(err), rmse=float(np.sqrt(np.mean(err * err))), mean=float(err.mean()),
                median=float(np.median(err)), std=float(err.std()), min=float(err.min()), max=float(err.max()))


def _rot_angle(T):
    return float(np.arccos(np.clip((np.trace(T[:3, :3]) - 1) / 2, -1.0, 1.0)))


def relative_pose_error(traj_gt, traj_est, delta=1.0, delta_unit="s", offset=0.0, scale=1.0):
    """RPE over pairs at a fixed delta (evaluate_rpe.py --fixed_delta): for each estimated pose i the pose j whose index
    (time in 's', frame number in 'f') is closest to index_i + delta, skipping j = last; ground-truth poses are the
    closest in time and a pair is dropped when either is farther than twice the median ground-truth interval.
    Returns translational / rotational error statistics."""
    sg = np.array(sorted(traj_gt.keys()))
    se = np.array(sorted(traj_est.keys()))
    if delta_unit == "s":
        index = se
    elif delta_unit == "f":
        index = np.arange(len(se), dtype=np.float64)
    else:
        raise ValueError("delta_unit must be 's' or 'f'")

    def closest(arr, t):
        k = int(np.searchsorted(arr, t))
        if k <= 0:
            return 0
        if k >= len(arr):
            return len(arr) - 1
        return k if arr[k] - t < t - arr[k - 1] else k - 1

    max_dt = 2 * float(np.median(np.diff(sg)))
    trans, rot = [], []
    for i in range(len(se)):
        j = closest(index, index[i] + delta)
        if j == len(se) - 1:
            continue
        g0, g1 = sg[closest(sg, se[i] + offset)], sg[closest(sg, se[j] + offset)]
        if abs(g0 - (se[i] + offset)) > max_dt or abs(g1 - (se[j] + offset)) > max_dt:
            continue
        # ominus(a, b) = inv(a) b with a = the LATER pose (evaluate_rpe.py:138-149, 285-287)
        d_est = np.linalg.inv(traj_est[se[j]]) @ traj_est[se[i]]
        d_est[:3, 3] *= scale
        d_gt = np.linalg.inv(traj_gt[g1]) @ traj_gt[g0]
        E = np.linalg.inv(d_est) @ d_gt
        trans.append(float(np.linalg.norm(E[:3, 3])))
        rot.append(_rot_angle(E))
    if len(trans) < 2:
        raise ValueError("could not find matching timestamp pairs between ground truth and estimate")
    trans, rot = np.array(trans), np.array(rot)
    return dict(pairs=len(trans), trans_rmse=float(np.sqrt(np.mean(trans ** 2))), trans_mean=float(trans.mean()),
                trans_median=float(np.median(trans)), trans_max=float(trans.max()),
                rot_rmse=float(np.sqrt(np.mean(rot ** 2))), rot_mean=float(rot.mean()),
                rot_median=float(np.median(rot)), rot_max=float(rot.max()))
