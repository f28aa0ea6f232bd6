"""Scratch: timing sweep for cfg2 (fixed ell 0.1, 100 iters) + stock cvo + acvo + cfg5."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
ctx = capi.Context(0, max_points=10240, max_slots=296)
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
pairs = [synth.config_pair(2, i) for i in range(296)]
for s, pr in enumerate(pairs):
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
ctx.sync()
for P, G in ((1, 16), (1, 8), (1, 1), (18, 8), (74, 2), (148, 1), (296, 1), (296, 2)):
    ctx.set_cluster_size(G)
    for rep in range(2):
        r = ctx.align(list(range(P)), gp)
    ms = ctx.last_kernel_ms
    print(f"cfg2 P={P} G={G} ncl={ctx.last_num_clusters} kernel_ms={ms:.3f} pairs/s={P/ms*1e3:.1f} us/iter/pair-slot={ms*1e3/100/max(1,-(-P//ctx.last_num_clusters)):.1f}")
ctx.set_cluster_size(0)
gp = capi.default_params('cvo')
for rep in range(2):
    r = ctx.align(list(range(296)), gp)
print(f"cvo stock P=296 kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={296/ctx.last_kernel_ms*1e3:.1f} iters mean={r['iters'].mean():.1f}")
prs = [synth.config_pair(3, i) for i in range(148)]
for s, pr in enumerate(prs):
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('acvo')
for rep in range(2):
    r = ctx.align(list(range(148)), gp)
print(f"acvo stock P=148 kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={148/ctx.last_kernel_ms*1e3:.1f} iters mean={r['iters'].mean():.1f}")
pr = synth.config_pair(5)
ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 20
for G in (16, 8):
    ctx.set_cluster_size(G)
    for rep in range(2):
        r = ctx.align([0], gp)
    print(f"cfg5 10k G={G} kernel_ms={ctx.last_kernel_ms:.3f} per-iter={ctx.last_kernel_ms/20*1e3:.1f} us")
