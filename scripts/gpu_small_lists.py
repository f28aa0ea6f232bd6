"""Scratch: small align runs through the list paths (for compute-sanitizer racecheck / memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
for kind, n, m, G in (("cvo", 700, 650, 1), ("acvo", 600, 640, 2)):
    pr = synth.make_pair(7, n, m, kind)
    ctx = capi.Context(0, max_points=1024, max_slots=2)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    ctx.set_pair(1, pr["y_pos"], pr["y_feat"], pr["x_pos"], pr["x_feat"])
    ctx.set_cluster_size(G)
    gp = capi.default_params(kind)
    gp.fixed_iters = 6 if kind == "cvo" else 40  # (acvo: long enough for ell to come down again)
    r = ctx.align([0, 1], gp)
    print(kind, G, r["transform"][0][:3, 3], ctx.last_list_builds, "refines", ctx.last_list_refines)
    ctx.close()
