"""Tuning aid / evidence: whole-GPU mode (cvo_b200_set_group_clusters) on single pairs -- BASELINE config 2 (3000 x 3000,
fixed ell 0.10, 100 iterations) and config 5 (10000 x 10000, fixed ell 0.10, 20 iterations): device time per iteration
by (CTAs per cluster, clusters per pair), and the pose against the one-cluster run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth

def pose_diff(A, B):
    D = np.linalg.inv(np.asarray(A, float)) @ np.asarray(B, float)
    S = (D[:3, :3] - D[:3, :3].T) / 2
    return float(np.linalg.norm([S[2, 1], S[0, 2], S[1, 0]])), float(np.linalg.norm(D[:3, 3]))

ctx = capi.Context(0, max_points=10240, max_slots=2)
for cfg, iters, combos in ((2, 100, [(16, 1), (16, 2), (16, 3), (16, 4), (16, 8), (8, 6), (8, 12), (8, 18), (16, 0)]),
                           (5, 20, [(16, 1), (16, 2), (16, 4), (16, 8), (8, 12), (8, 18), (4, 36), (16, 0)]),
                           (5, 100, [(16, 1), (16, 8), (8, 18), (16, 0)])):
    pr = synth.config_pair(cfg)
    ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = iters
    ref = None
    for G, H in combos:
        ctx.set_cluster_size(G); ctx.set_group_clusters(H)
        best = 1e9
        for rep in range(3):
            r = ctx.align([0], gp)
            best = min(best, ctx.last_kernel_ms)
        if ref is None: ref = r['transform'][0]
        rot, tr = pose_diff(ref, r['transform'][0])
        print("cfg%d %4d iters  G=%2d clusters=%3d (asked %2d) -> %3d CTAs  %8.3f ms  %6.1f us/iter  builds %d  vs one cluster: %.1e rad %.1e m"
              % (cfg, iters, ctx.last_cluster_size, ctx.last_group_clusters, H, ctx.last_cluster_size * ctx.last_group_clusters,
                 best, best / iters * 1e3, ctx.last_list_builds, rot, tr), flush=True)
