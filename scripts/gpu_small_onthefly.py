"""Scratch: one small on-the-fly eval (for compute-sanitizer)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
from oracle import cvo_oracle as O
pr = synth.make_pair(1000, 500, 500, "cvo")
ctx = capi.Context(0, max_points=1024, max_slots=1)
ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
ctx.set_neighbor_lists(int(os.environ.get("LISTS", "0")))
ctx.set_cluster_size(int(os.environ.get("G", "1")))
gp = capi.default_params("cvo")
g = ctx.eval(0, np.eye(3), np.zeros(3), 0.1, gp)
o = O.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], np.eye(3), np.zeros(3), 0.1, O.default_params("cvo"))
print("nnz", g["nnz"], o["nnz"], "B", g["B"], o["B"], "omega", g["omega"], o["omega"])
