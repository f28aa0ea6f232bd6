"""Scratch: a short align workload for ncu (1 GPU). usage: prof_run.py P G iters"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
P, G, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = capi.Context(0, max_points=3072, max_slots=P)
for s in range(P):
    pr = synth.config_pair(2, s)
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = iters
ctx.set_cluster_size(G)
for rep in range(2):
    ctx.align(list(range(P)), gp)
print('kernel_ms', ctx.last_kernel_ms, 'list builds', ctx.last_list_builds, 'fill (entries, slots)', ctx.last_list_fill)
