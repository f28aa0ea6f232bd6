"""Scratch: where the end-to-end step time goes beyond the align kernel (bench.py's e2e arm, one GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cvo_rgbd_b200 import capi, synth
P, N = 296, 3000
ctx = capi.Context(0, max_points=N + 72, max_slots=2 * P)
pin = lambda shape: torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
hx, hfx, hy, hfy = pin((P, N, 3)), pin((P, N, 5)), pin((P, N, 3)), pin((P, N, 5))
for s in range(P):
    pr = synth.config_pair(2, s)
    hx[s], hfx[s], hy[s], hfy[s] = pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"]
counts = np.full(P, N, np.int32)
sets = (np.arange(P, dtype=np.int32), np.arange(P, dtype=np.int32) + P)
gp = capi.default_params("cvo"); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
def steps(n, log=None):
    ctx.set_pairs(sets[0], hx, hfx, counts, hy, hfy, counts)
    for k in range(n):
        t0 = time.perf_counter()
        if k + 1 < n:
            ctx.set_pairs(sets[(k + 1) % 2], hx, hfx, counts, hy, hfy, counts)
        t1 = time.perf_counter()
        ctx.align(sets[k % 2], gp)
        t2 = time.perf_counter()
        if log is not None:
            log.append((t1 - t0, t2 - t1, ctx.last_kernel_ms))
steps(3)
log = []
t0 = time.perf_counter(); steps(8, log); dt = time.perf_counter() - t0
print("e2e ms/step %.3f" % (dt / 8 * 1e3))
for a, b, k in log:
    print("set_pairs call %.3f ms   align call %.3f ms   (kernel %.3f ms, align call - kernel = %.3f ms)" % (a * 1e3, b * 1e3, k, b * 1e3 - k))
