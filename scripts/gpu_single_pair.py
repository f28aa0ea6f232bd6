import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from cvo_rgbd_b200 import capi, synth
tag = os.path.basename(os.environ.get("CVO_B200_LIB", "default"))
ctx = capi.Context(0, max_points=3072, max_slots=2)
pr = synth.config_pair(2)
ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
for G in (16, 8, 4, 2, 1):
    ctx.set_cluster_size(G); ctx.set_group_clusters(1)
    for rep in range(3): ctx.align([0], gp)
    print(tag, "G", G, "ms", ctx.last_kernel_ms, flush=True)
