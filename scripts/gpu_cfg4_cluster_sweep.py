"""Device time of a GPU's share of BASELINE config 4 (ragged pairs, stock cvo schedule, stop tests on) by cluster size:
P = 500 / 250 / 125 / 63 pairs (the share at 1 / 2 / 4 / 8 GPUs), G = automatic and forced.  The pairs' iteration counts
differ by a factor of three, so a single wave lasts as long as its slowest pair: the table is what choose_cluster's
cost model (cvo_api.cu) is fitted to."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
from cvo_rgbd_b200 import capi, synth  # noqa: E402

Gs = [int(g) for g in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 3, 4, 6, 8]
prs = [synth.config_pair(4, i) for i in range(500)]
ctx = capi.Context(0, max_points=3328, max_slots=500)
for s, pr in enumerate(prs):
    ctx.set_pair(s, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
gp = capi.default_params("cvo")
shares = [("W=%d" % W, np.arange(0, 500, W)) for W in (1, 2, 4, 8)]  # rank 0's share: pair p -> rank p mod W
if len(sys.argv) > 2:  # other batch sizes: every (500 // P)-th pair
    shares = [("every %d" % (500 // P), np.arange(0, 500, 500 // P)[:P]) for P in (int(a) for a in sys.argv[2].split(","))]
for W, slots in shares:
    P = len(slots)
    row = []
    for G in Gs:
        ctx.set_cluster_size(G)
        best = 1e9
        for rep in range(3):
            res = ctx.align(slots, gp)
            best = min(best, ctx.last_kernel_ms)
        row.append("G=%d%s: %.3f ms" % (ctx.last_cluster_size, " (auto)" if G == 0 else "", best))
    print("%s P=%3d iters mean %.1f max %d | %s" % (W, P, res["iters"].mean(), res["iters"].max(), " | ".join(row)), flush=True)
ctx.close()
