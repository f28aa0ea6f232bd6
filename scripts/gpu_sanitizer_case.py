"""A small workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): neighbour lists with sweeps,
wide-list filters, self-list narrowing, quad passes, cluster and whole-GPU reductions -- cvo and acvo, 700 x 800 points,
a dozen iterations each, on 1, 2 and 2 x 2 CTAs per pair; plus three iterations of a 3300 x 3200 pair on one CTA (several row
rounds and column chunks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
ctx = capi.Context(0, max_points=1024, max_slots=2)
for kind in ("cvo", "acvo"):
    pr = synth.make_pair(7, 700, 800, kind)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    for G, H in ((1, 1), (2, 1), (2, 2)):
        ctx.set_cluster_size(G); ctx.set_group_clusters(H)
        gp = capi.default_params(kind); gp.max_iter = 12
        r = ctx.align([0], gp)
        print(kind, G, H, r["transform"][0][:3, 3], int(r["iters"][0]), "sweeps", ctx.last_list_builds, "filters", ctx.last_list_refines, flush=True)
# several row rounds and column chunks on one CTA (more than 3072 points on both sides), a sweep and a wide-list filter
ctx.close()
ctx = capi.Context(0, max_points=4096, max_slots=1)
pr = synth.make_pair(9, 3300, 3200, "acvo", motion_scale=1.5)
ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
ctx.set_cluster_size(1); ctx.set_group_clusters(1)
gp = capi.default_params("acvo"); gp.max_iter = 3
r = ctx.align([0], gp)
print("acvo 3300 x 3200 on one CTA", r["transform"][0][:3, 3], "sweeps", ctx.last_list_builds, "filters", ctx.last_list_refines, flush=True)
