"""Scratch: short timing set for kernel variants (CVO_B200_LIB selects the library)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
tag = os.path.basename(os.environ.get("CVO_B200_LIB", "default")) + (" nolists" if os.environ.get("CVO_B200_NO_LISTS") == "1" else " skin=" + os.environ.get("CVO_B200_LIST_SKIN", "0.08"))
ctx = capi.Context(0, max_points=10240, max_slots=296)
pairs = [synth.config_pair(2, i) for i in range(296)]
for s, pr in enumerate(pairs):
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
ctx.sync()
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
for P, G in ((296, 1), (1, 16)):
    ctx.set_cluster_size(G)
    for rep in range(2):
        r = ctx.align(list(range(P)), gp)
    print(f"[{tag}] cfg2 P={P} G={G} kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={P/ctx.last_kernel_ms*1e3:.1f} list_builds/pair={ctx.last_list_builds/P:.1f}")
ctx.set_cluster_size(0)
gp = capi.default_params('cvo')
for rep in range(2):
    r = ctx.align(list(range(296)), gp)
print(f"[{tag}] cvo stock P=296 kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={296/ctx.last_kernel_ms*1e3:.1f} iters mean={r["iters"].mean():.1f} max={r["iters"].max()} builds/pair={ctx.last_list_builds/len(r["iters"]):.1f} refines/pair={ctx.last_list_refines/len(r["iters"]):.1f}")
prs = [synth.config_pair(3, i) for i in range(148)]
for s, pr in enumerate(prs):
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('acvo')
for rep in range(2):
    r = ctx.align(list(range(148)), gp)
print(f"[{tag}] acvo stock P=148 kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={148/ctx.last_kernel_ms*1e3:.1f} iters mean={r["iters"].mean():.1f} max={r["iters"].max()} builds/pair={ctx.last_list_builds/len(r["iters"]):.1f} refines/pair={ctx.last_list_refines/len(r["iters"]):.1f}")
pr = synth.config_pair(5)
ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 20
ctx.set_cluster_size(16)
for rep in range(2):
    r = ctx.align([0], gp)
print(f"[{tag}] cfg5 10k G=16 kernel_ms={ctx.last_kernel_ms:.3f} per-iter={ctx.last_kernel_ms/20*1e3:.1f} us")
