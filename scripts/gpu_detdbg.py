"""Scratch: determinism of align across repeats / contexts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
pr = synth.make_pair(5, 1000, 1000, "cvo")
res = []
for mp in (2048, 3008, 2048, 10240):
    ctx = capi.Context(0, max_points=mp, max_slots=2)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    gp = capi.default_params("cvo")
    for rep in range(3):
        r = ctx.align_trace(0, gp, trace_cap=200)
        res.append((mp, rep, r))
        print(mp, rep, r["iters"], r["transform"][:3, 3], ctx.last_list_builds)
    ctx.close()
base = res[0][2]
for mp, rep, r in res[1:]:
    n = min(len(r["trace"]), len(base["trace"]))
    first = None
    for k in range(n):
        a, b = r["trace"][k], base["trace"][k]
        if a["nnz"] != b["nnz"] or a["B"] != b["B"] or not np.array_equal(a["omega"], b["omega"]) or a["sum_a"] != b["sum_a"]:
            first = k; break
    print(mp, rep, "first differing iteration:", first, "" if first is None else (r["trace"][first]["nnz"], base["trace"][first]["nnz"], r["trace"][first]["sum_a"], base["trace"][first]["sum_a"]))
