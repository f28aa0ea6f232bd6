"""Scratch: eval with neighbour lists on / off vs the oracle on a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
from oracle import cvo_oracle as O
R0 = synth._rotvec_to_R(np.array([0.01, -0.02, 0.015])).astype(np.float32)
T0 = np.array([0.01, 0.005, -0.02], np.float32)
ctx = capi.Context(0, max_points=10240, max_slots=4)
for kind, seed, n, m in [("cvo", 1000, 500, 500), ("cvo", 52, 33, 2100), ("acvo", 54, 600, 1000), ("acvo", 53, 900, 650), ("acvo", 3000, 3000, 3000)]:
    pr = synth.make_pair(seed, n, m, kind)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    gp, op = capi.default_params(kind), O.default_params(kind)
    o = O.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], R0, T0, 0.1, op)
    for G in (1, 16):
        ctx.set_cluster_size(G)
        for lists in (False, True):
            ctx.set_neighbor_lists(lists)
            for rep in range(2):
                g = ctx.eval(0, R0, T0, 0.1, gp)
                print(kind, n, m, "G", G, "lists", lists, "nnz", g["nnz"], o["nnz"], "xx", g["nnz_xx"], o["nnz_xx"], "yy", g["nnz_yy"], o["nnz_yy"],
                      "sum_a %.9g %.9g" % (g["sum_a"], o["sum_a"]), "B %.9g %.9g" % (g["B"], o["B"]), "dl %.6g %.6g" % (g["dl"], o["dl"]))
