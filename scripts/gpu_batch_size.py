import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from cvo_rgbd_b200 import capi, synth
prs = [synth.config_pair(2, s) for s in range(296)]
for P in (148, 296, 444, 592, 888, 1184):
    ctx = capi.Context(0, max_points=3072, max_slots=P)
    for s in range(P):
        pr = prs[s % 296]
        ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
    best = 1e9
    for rep in range(3):
        ctx.align(list(range(P)), gp); best = min(best, ctx.last_kernel_ms)
    print("P=%4d cfg2 %.3f ms = %.0f pairs/s" % (P, best, P / best * 1e3), flush=True)
    ctx.close()
