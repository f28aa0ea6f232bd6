"""GPU-vs-oracle pose-error DISTRIBUTION -> profiles/r02_parity_distribution.json.

For each mode (cfg2 fixed ell / stock cvo / stock acvo) N seeded 3000 x 3000 pairs are aligned (a) on the GPU in one
batch through the C ABI, (b) by the oracle port (brute-force ball, no FMA contraction), (c) by the oracle built with
the reference's nanoflann and GCC's default contraction (oracle/_ref).  Reported: quantiles of the GPU-vs-port pose
error beside port-vs-ref (the same restatement compiled two ways: the noise floor of the algorithm itself), and the
fraction of pairs inside north_star's 1e-4 rad / 1e-4 m.

Second part: the (y, y) self list of acvo keeps the squared distances of its BUILD pose while the reference recomputes
them from the freshly transformed cloud every iteration (src/adaptive_cvo.cpp:160).  For a few full acvo runs the
GPU's per-iteration nnz_yy is compared with the oracle evaluated at the GPU's own state of that iteration:
max |delta nnz_yy| is the measured cost of that deviation (plus the usual ulp-level boundary flips).

    python scripts/parity_distribution.py [N]        (GPU box; ~3 min of host CPU for N = 200)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import pose_diff  # noqa: E402
from cvo_rgbd_b200 import capi, synth  # noqa: E402
from oracle import cvo_oracle as O  # noqa: E402


def quantiles(a):
    a = np.asarray(a, float)
    return {k: float(np.quantile(a, q)) for k, q in (("p50", .5), ("p90", .9), ("p95", .95), ("p99", .99), ("max", 1.0))}


def summarize(errs):
    errs = np.asarray(errs, float)
    return {"rot_rad": quantiles(errs[:, 0]), "trans_m": quantiles(errs[:, 1]),
            "frac_within_1e-4": float(((errs[:, 0] < 1e-4) & (errs[:, 1] < 1e-4)).mean()),
            "frac_within_3e-4": float(((errs[:, 0] < 3e-4) & (errs[:, 1] < 3e-4)).mean())}


def run_mode(name, kind, cfg, n_pairs, fixed):
    prs = [synth.config_pair(cfg, 5000 + i) for i in range(n_pairs)]  # seeds disjoint from the tests' and the bench's
    n = 3000
    hx = np.stack([p["x_pos"] for p in prs]).astype(np.float32)
    hfx = np.stack([p["x_feat"] for p in prs]).astype(np.float32)
    hy = np.stack([p["y_pos"] for p in prs]).astype(np.float32)
    hfy = np.stack([p["y_feat"] for p in prs]).astype(np.float32)
    counts = np.full(n_pairs, n, np.int32)
    gp = capi.default_params(kind)
    if fixed:
        gp.ell_policy, gp.ell_init, gp.fixed_iters = capi.ELL_FIXED, 0.10, 100
    with capi.Context(0, max_points=n + 72, max_slots=n_pairs) as ctx:
        ctx.set_pairs(np.arange(n_pairs, dtype=np.int32), hx, hfx, counts, hy, hfy, counts)
        g = ctx.align(np.arange(n_pairs, dtype=np.int32), gp)
        G = ctx.last_cluster_size
    out = {}
    poses = {}
    for variant in ("port", "ref"):
        O.load(variant)
        op = O.default_params(kind, variant)
        if fixed:
            op.ell_policy, op.ell_init, op.fixed_iters = O.ELL_FIXED, 0.10, 100
        t0 = time.time()
        poses[variant] = [O.align(p["x_pos"], p["x_feat"], p["y_pos"], p["y_feat"], op, variant=variant) for p in prs]
        out["oracle_%s_seconds" % variant] = time.time() - t0
    e_gp = [pose_diff(g["transform"][i], poses["port"][i]["transform"]) for i in range(n_pairs)]
    e_gr = [pose_diff(g["transform"][i], poses["ref"][i]["transform"]) for i in range(n_pairs)]
    e_pr = [pose_diff(poses["port"][i]["transform"], poses["ref"][i]["transform"]) for i in range(n_pairs)]
    out.update({"pairs": n_pairs, "kind": kind, "cfg": cfg, "fixed_ell_100_iters": fixed, "ctas_per_pair": G,
                "gpu_vs_oracle_port": summarize(e_gp), "gpu_vs_oracle_ref_nanoflann": summarize(e_gr),
                "oracle_port_vs_oracle_ref": summarize(e_pr),
                "iters_gpu_mean": float(np.mean(g["iters"])),
                "iters_oracle_port_mean": float(np.mean([p["iters"] for p in poses["port"]]))})
    print(name, json.dumps(out["gpu_vs_oracle_port"]), "| oracle vs itself:", json.dumps(out["oracle_port_vs_oracle_ref"]))
    return out


def stale_yy(n_runs):
    """max |nnz_yy(GPU list pass, iteration k) - nnz_yy(oracle at the GPU's state of iteration k)| over full acvo runs."""
    op = O.default_params("acvo")
    gp = capi.default_params("acvo")
    worst, worst_xx, worst_xy, total_iters, nonzero = 0, 0, 0, 0, 0
    with capi.Context(0, max_points=3072, max_slots=1) as ctx:
        for r in range(n_runs):
            pr = synth.config_pair(3, 6000 + r)
            ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
            g = ctx.align_trace(0, gp, trace_cap=2048)
            R, T = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
            for k, rec in enumerate(g["trace"]):
                o = O.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], R, T, rec["ell"], op)
                d = abs(rec["nnz_yy"] - o["nnz_yy"])
                worst = max(worst, d)
                worst_xx = max(worst_xx, abs(rec["nnz_xx"] - o["nnz_xx"]))
                worst_xy = max(worst_xy, abs(rec["nnz"] - o["nnz"]))
                nonzero += d != 0
                total_iters += 1
                R, T = rec["R"], rec["T"]
    return {"acvo_runs": n_runs, "iterations": total_iters, "max_abs_delta_nnz_yy": int(worst),
            "iterations_with_delta": int(nonzero), "max_abs_delta_nnz_xx": int(worst_xx), "max_abs_delta_nnz_xy": int(worst_xy),
            "note": "oracle evaluated at the GPU's own (R, T, ell) of every iteration; nnz_yy of the order of 1e5"}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    out = {"what": __doc__.split("\n\n")[0], "modes": {}}
    only = sys.argv[2] if len(sys.argv) > 2 else None  # one mode only (tuning experiments)
    if only in (None, "cfg2"):
        out["modes"]["cfg2_fixed_ell_0.10_100_iters"] = run_mode("cfg2", "cvo", 2, n, True)
    if only in (None, "cvo"):
        out["modes"]["stock_cvo"] = run_mode("cvo", "cvo", 2, n, False)
    if only in (None, "acvo"):
        out["modes"]["stock_acvo"] = run_mode("acvo", "acvo", 3, n, False)
    if only is not None:
        return
    out["stale_yy_list"] = stale_yy(8)
    print(json.dumps(out["stale_yy_list"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "r02_parity_distribution.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
