"""Evidence: device time of the workloads beside the cfg-2 headline -- stock cvo / acvo batches (2 x #SMs pairs), the single
cfg-2 pair (latency mode) and BASELINE config 5 (10 000 x 10 000 points, fixed ell 0.10, 20 and 100 iterations, whole-GPU
mode).  usage: gpu_other_configs.py all | cvo | acvo | cfg5   (one workload, three launches: for ncu -s 2 -c 1)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
what = sys.argv[1] if len(sys.argv) > 1 else "all"
P = 296

def batch(kind, cfg):
    ctx = capi.Context(0, max_points=3072, max_slots=P)
    for s in range(P):
        pr = synth.config_pair(cfg, s)
        ctx.set_pair(s, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    gp = capi.default_params(kind)
    for rep in range(3):
        r = ctx.align(list(range(P)), gp)
    print("stock %-4s %d pairs: %.3f ms = %.0f pairs/s  (iterations %.1f mean, %d max; sweeps %.2f, filters %.2f per pair; %d CTAs per pair)"
          % (kind, P, ctx.last_kernel_ms, P / ctx.last_kernel_ms * 1e3, r["iters"].mean(), r["iters"].max(),
             ctx.last_list_builds / P, ctx.last_list_refines / P, ctx.last_cluster_size), flush=True)
    ctx.close()

def single(cfg, iters):
    ctx = capi.Context(0, max_points=10240, max_slots=1)
    pr = synth.config_pair(cfg)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    gp = capi.default_params("cvo"); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = iters
    for rep in range(3):
        ctx.align([0], gp)
    n = len(pr["x_pos"])
    print("cfg%d single pair %d x %d, %d iterations: %.3f ms = %.1f us per iteration  (%d clusters x %d CTAs, %d sweeps; algorithmic %.2f MB per iteration -> %.1f GB/s)"
          % (cfg, n, len(pr["y_pos"]), iters, ctx.last_kernel_ms, ctx.last_kernel_ms / iters * 1e3, ctx.last_group_clusters,
             ctx.last_cluster_size, ctx.last_list_builds, (64 * 2 * n + 96) / 1e6, (64 * 2 * n + 96) * iters / ctx.last_kernel_ms / 1e6), flush=True)
    ctx.close()

if what in ("all", "cvo"): batch("cvo", 2)
if what in ("all", "acvo"): batch("acvo", 3)
if what == "all": single(2, 100)
if what in ("all", "cfg5"): single(5, 20)
if what == "all": single(5, 100)
