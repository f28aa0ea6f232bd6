"""Tuning aid: per-phase cycle split, mean over the CTAs; `stock` = the stock cvo schedule instead of cfg2 (needs the CVO_PHASE_CLOCKS variant: build_variants.py clk:CVO_PHASE_CLOCKS)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
lib = capi.load()
ctx = capi.Context(0, max_points=3072, max_slots=296)
for s in range(296):
    pr = synth.config_pair(2, s)
    ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
gp = capi.default_params('cvo')
if "stock" not in sys.argv:
    gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
P = 1 if "single" in sys.argv else 296  # `single`: one pair on one 16-CTA cluster (latency mode)
if P == 1:
    ctx.set_cluster_size(16); ctx.set_group_clusters(1)
ctx.align(list(range(P)), gp)
out = (C.c_ulonglong * 24)()
lib.cvo_b200_phase_clocks(out, 1)
ctx.align(list(range(P)), gp)
lib.cvo_b200_phase_clocks(out, 1)
v = np.array(out[:23], float)
names = ["(loop top)", "list build: rest", "FLOW pass (trips of warp 0)", "allreduce + finalize_flow", "STEP pass (trips of warp 0)", "barrier after the serial section", "build: stage", "build: evaluate: wait for the slowest warp", "list passes: tail (slowest warp + reduction)", "build: scatter: wait for the slowest warp", "serial: all-reduce of B..E", "serial: update_state after the step", "serial: prepare_iter", "serial: list_policy", "serial: step_from_coeffs", "list passes: staging barrier + tags", "FLOW: entry (setup + barrier)", "FLOW: column staging (thread 0)", "STEP: entry (setup + barrier)", "STEP: row terms (thread 0)", "build: evaluate (warp 0's units)", "build: count + place", "build: scatter (warp 0's tiles)"]
v /= ctx.last_num_clusters * ctx.last_cluster_size  # summed over the CTAs by the kernel
print("kernel_ms", ctx.last_kernel_ms, "list builds", ctx.last_list_builds, "fill", ctx.last_list_fill, "mean Mcycles per CTA:", v.sum() / 1e6)
for n, x in zip(names, v):
    print("%-45s %8.2f Mcycles  %5.1f %%" % (n, x / 1e6, 100 * x / v.sum()))
