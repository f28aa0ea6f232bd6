"""Scratch GPU probe: parity vs oracle + timing sweep. Not part of the product."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
from oracle import cvo_oracle as O

def show(tag, g, o, keys):
    for k in keys:
        a = np.asarray(g[k], float); b = np.asarray(o[k], float)
        rel = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
        print(f"  {tag} {k}: gpu={a} oracle={b} rel={rel:.2e}")

def main():
    ctx = capi.Context(0, max_points=10240, max_slots=300)
    print("SMs", ctx.num_sms)
    # ---- level 1: eval parity
    for cfg, kind in ((1, 'cvo'), (2, 'cvo'), (3, 'acvo')):
        pr = synth.config_pair(cfg)
        ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
        gp = capi.default_params(kind); op = O.default_params(kind)
        R = synth._rotvec_to_R(np.array([0.01, -0.02, 0.015])).astype(np.float32); T = np.array([0.01, 0.005, -0.02], np.float32)
        for ell in (0.15, 0.1, 0.05):
            g = ctx.eval(0, R, T, ell, gp)
            o = O.evaluate(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], R, T, ell, op)
            print(f"cfg{cfg} {kind} ell={ell} kernel_ms={ctx.last_kernel_ms:.3f} G={ctx.last_cluster_size}")
            keys = ['nnz', 'sum_a', 'omega', 'v', 'B', 'C', 'D', 'E', 'step']
            if kind == 'acvo': keys += ['nnz_xx', 'nnz_yy', 'dl']
            show('', g, o, keys)
    # ---- level 3: align parity
    for cfg, kind in ((1, 'cvo'), (2, 'cvo'), (3, 'acvo')):
        pr = synth.config_pair(cfg)
        ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
        gp = capi.default_params(kind); op = O.default_params(kind)
        t0 = time.time(); g = ctx.align_trace(0, gp); tg = time.time() - t0
        t0 = time.time(); o = O.align(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], op, trace_cap=2048); to = time.time() - t0
        print(f"align cfg{cfg} {kind}: gpu iters={g['iters']} status={g['status']} kernel_ms={ctx.last_kernel_ms:.3f} wall={tg*1e3:.1f}ms | oracle iters={o['iters']} status={o['status']} wall={to*1e3:.1f}ms")
        dT = np.linalg.inv(o['transform'].astype(float)) @ g['transform'].astype(float)
        ang = np.arccos(np.clip((np.trace(dT[:3,:3]) - 1) / 2, -1, 1)); print(f"   pose diff: rot={ang:.2e} rad trans={np.linalg.norm(dT[:3,3]):.2e} m  ell gpu={g['ell']} oracle={o['ell']}")
        for k in (0, 1, 5, 20):
            if k < len(g['trace']) and k < len(o['trace']):
                show(f'k={k}', g['trace'][k], o['trace'][k], ['nnz', 'omega', 'v', 'step'])
    # ---- timing sweep: cfg2 fixed ell, 100 iters
    gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 100
    NP = 296
    pairs = [synth.config_pair(2, i % 8) for i in range(8)]
    for s in range(NP):
        pr = pairs[s % 8]
        ctx.set_pair(s, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    ctx.sync()
    for P, G in ((1, 16), (1, 8), (1, 4), (1, 1), (8, 16), (18, 8), (37, 4), (74, 2), (148, 1), (296, 1), (296, 2)):
        ctx.set_cluster_size(G)
        for rep in range(2):
            r = ctx.align(list(range(P)), gp)
        ms = ctx.last_kernel_ms
        print(f"P={P} G={G} ncl={ctx.last_num_clusters} kernel_ms={ms:.3f} pairs/s={P/ms*1e3:.1f} iters={ctx.last_total_iterations}")
    ctx.set_cluster_size(0)
    # stock cvo schedule batch
    gp = capi.default_params('cvo')
    for rep in range(2):
        r = ctx.align(list(range(148)), gp)
    print(f"cvo stock P=148 kernel_ms={ctx.last_kernel_ms:.3f} pairs/s={148/ctx.last_kernel_ms*1e3:.1f} iters mean={r['iters'].mean():.1f}")
    # 10k stress
    pr = synth.config_pair(5)
    ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    gp = capi.default_params('cvo'); gp.ell_policy = capi.ELL_FIXED; gp.ell_init = 0.10; gp.fixed_iters = 20
    for G in (16, 8):
        ctx.set_cluster_size(G)
        for rep in range(2):
            r = ctx.align([0], gp)
        print(f"cfg5 10k G={G} kernel_ms={ctx.last_kernel_ms:.3f} per-iter={ctx.last_kernel_ms/20*1e3:.1f} us")
    op = O.default_params('cvo'); 
    g = ctx.eval(0, np.eye(3), np.zeros(3), 0.1, capi.default_params('cvo'))
    o = O.evaluate(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], np.eye(3), np.zeros(3), 0.1, op)
    show('10k', g, o, ['nnz', 'sum_a', 'omega', 'v', 'B', 'E', 'step'])
    # inner product
    pr = synth.config_pair(3)
    ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    g = ctx.inner_product(0, 0.1, capi.default_params('acvo'))
    o = O.inner_product(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], 0.1, O.default_params('acvo'))
    print('inner product', g, o)

main()
