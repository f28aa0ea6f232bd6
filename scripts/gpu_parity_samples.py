"""Checks, on one GPU, the parity samples bench.py draws at N = 2 / 4 / 8 (rank 0's pairs s * N of the 592-pair batch):
the same pairs through one CTA each (the benchmark's launch geometry) against the oracle, with bench.py's own criterion.
usage: gpu_parity_samples.py [pairs per sample]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import bench  # noqa: E402
from cvo_rgbd_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
P = 592
for world in (2, 4, 8):
    sample = np.unique(np.linspace(0, P - 1, n).astype(int))
    ids = [int(i) * world for i in sample]
    ctx = capi.Context(0, max_points=bench.N_POINTS + 72, max_slots=len(ids))
    ctx.set_cluster_size(1)
    for s, i in enumerate(ids):
        pr = synth.config_pair(2, i)
        ctx.set_pair(s, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    res = ctx.align(np.arange(len(ids)), bench.make_params(capi))
    r = bench.cpu_reference_run(ids)
    rep = bench.parity_report(res["transform"], r["poses"], bench.POSE_TOL)
    print("N=%d: %d pairs, within tol %.3f, max %.2e rad %.2e m, ok %s" % (world, rep["pairs"], rep["frac_within_tol"], rep["max_rot"], rep["max_trans"], rep["ok"]), flush=True)
    ctx.close()
