"""Tuning aid: device time of the cfg2 batch launch (296 pairs, fixed ell 0.10, 100 iterations) for the library named by
CVO_B200_LIB (scripts/build_variants.py), plus stock cvo / acvo batches.  One line per workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
tag = os.path.basename(os.environ.get("CVO_B200_LIB", "default"))
P = 296
ctx = capi.Context(0, max_points=3072, max_slots=P)
for s in range(P):
    pr = synth.config_pair(2, s)
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
best = 1e9
for rep in range(4):
    r = ctx.align(list(range(P)), gp)
    best = min(best, ctx.last_kernel_ms)
print("%-28s cfg2 %.3f ms = %.0f pairs/s  (sweeps %.2f, filters %.2f per pair)" % (tag, best, P / best * 1e3, ctx.last_list_builds / P, ctx.last_list_refines / P), flush=True)
if "--all" in sys.argv:
    gp = capi.default_params('cvo')
    for rep in range(3):
        r = ctx.align(list(range(P)), gp)
    print("%-28s stock cvo %.3f ms = %.0f pairs/s (iters %.1f; sweeps %.2f, filters %.2f per pair)" % (tag, ctx.last_kernel_ms, P / ctx.last_kernel_ms * 1e3, r['iters'].mean(), ctx.last_list_builds / P, ctx.last_list_refines / P), flush=True)
    for s in range(P):
        pr = synth.config_pair(3, s)
        ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    gp = capi.default_params('acvo')
    for rep in range(3):
        r = ctx.align(list(range(P)), gp)
    print("%-28s stock acvo %.3f ms = %.0f pairs/s (iters %.1f)" % (tag, ctx.last_kernel_ms, P / ctx.last_kernel_ms * 1e3, r['iters'].mean()), flush=True)
