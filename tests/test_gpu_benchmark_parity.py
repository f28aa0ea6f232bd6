"""Parity of the workloads bench.py MEASURES, through the C ABI, against the CPU oracle at north_star's tolerance
(1e-4 rad / 1e-4 m per pair, written below):

  cfg2  BASELINE.json configs[1]: 3000 x 3000 points, ELL_FIXED 0.10, exactly 100 iterations, stop tests off --
        as a single pair on a 16-CTA cluster AND as pairs inside the k x #SMs batch the benchmark launches
        (one CTA per pair), i.e. the exact launch geometry of the headline number;
  cfg4  BASELINE.json configs[3]: ragged pairs N, M ~ U{2700..3300}, stock cvo schedule, identity init, aligned to
        convergence in ONE batch (reference loop: src/cvo_main.cpp:36-66 with cvo::align, src/cvo.cpp:361-420).

The reference algorithm does not reproduce its own final pose to 1e-4 on every pair: the line search takes the smallest
positive root of a cubic (src/cvo.cpp:291-307), which jumps when two roots merge, and two correct executions then
drift apart along the weakly constrained directions.  Measured over 200 pairs per mode
(profiles/r02_parity_distribution.json): the same CPU restatement compiled two ways agrees with itself within 1e-4 on
98 % (cfg2) / 92 % (stock cvo) / 97.5 % (stock acvo) of the pairs, worst pair 1.7e-4 .. 2.3e-4; GPU-vs-oracle shows the
same distribution (98.5 % / 92.5 % / 95.5 % at the round's first build, 96.5 % / 87 % / 95.5 % at its last).  The tests therefore assert, per workload:
  * 1e-4 on the single BASELINE pairs (seed 0 of cfg 1, 2, 3), as north_star states it;
  * in batches: 1e-4 on the bulk (>= 75 % of the sampled pairs, median), every pair inside 5e-4, and for every cfg2
    pair beyond 1e-4 that the two end states sit on the same flat top of the objective (the oracle's own objective
    and flow evaluated at the GPU's final (R, T)): they differ by where the iteration hovers, not by what it computes;
  * the chaos-free quantities -- the records of the FIRST iteration (identical inputs): nnz, omega, v, B..E, step --
    at 1e-5 relative on every checked pair."""
import numpy as np
import pytest

from conftest import POSE_TOL_FLOOR, POSE_TOL_NORTH_STAR, pose_diff, rel_err, iters_comparable  # noqa: F401
from cvo_rgbd_b200 import capi, synth

pytestmark = pytest.mark.gpu

CFG2_ELL, CFG2_ITERS, CFG2_POINTS = 0.10, 100, 3000


def _cfg2(p, mod):
    p.ell_policy, p.ell_init, p.fixed_iters = mod.ELL_FIXED, CFG2_ELL, CFG2_ITERS
    return p


def _oracle_cfg2(oracle, pr, trace_cap=0):
    return oracle.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], _cfg2(oracle.default_params("cvo"), oracle),
                        trace_cap=trace_cap)


def _check_record(g, o, tight):
    """One iteration record (nnz, omega, v, B..E, step) of the device against the oracle's."""
    assert abs(g["nnz"] - o["nnz"]) <= 2, (g["nnz"], o["nnz"])
    tol = 1e-5 if tight else 2e-3
    flips = abs(g["nnz"] - o["nnz"])
    for k in ("omega", "v"):
        scale = max(np.abs(o[k]).max(), 1e-30)
        assert np.abs(np.asarray(g[k]) - np.asarray(o[k])).max() < tol * scale + 2e-4 * flips + 1e-9, k
    if tight and flips == 0:
        for k in ("B", "C", "D", "E"):
            assert rel_err(g[k], o[k]) < 1e-5, k
        assert abs(g["step"] - o["step"]) < 1e-5 * max(1.0, o["step"])


def _assert_same_plateau(oracle, clouds, RT_gpu, o, ell, tag):
    """After 100 iterations at fixed ell the iteration still hovers around its fixed point (residual flow ~5e-4,
    twenty times eps): two executions end at different points of the same flat top of the objective.  The oracle's
    objective (sum of the kernel values, the quantity the flow ascends) at the GPU's end state must equal the one at
    its own end state to 2e-4 relative (measured between two compilations of the oracle: up to 7e-5), and the flow
    there must be of the same small order."""
    op = oracle.default_params("cvo")
    x, fx, y, fy = clouds
    at_gpu = oracle.evaluate(x, fx, y, fy, RT_gpu[:9].reshape(3, 3), RT_gpu[9:], ell, op)
    at_own = oracle.evaluate(x, fx, y, fy, o["R"], o["T"], ell, op)
    flow = lambda e: max(np.linalg.norm(e["omega"]), np.linalg.norm(e["v"]))  # noqa: E731
    assert abs(at_gpu["sum_a"] - at_own["sum_a"]) <= 2e-4 * at_own["sum_a"], (tag, at_gpu["sum_a"], at_own["sum_a"])
    assert flow(at_gpu) < 5e-3 and flow(at_own) < 5e-3, (tag, flow(at_gpu), flow(at_own))


def test_cfg2_exact_single_pair_on_a_cluster(gpu_ctx, oracle):
    """The benchmark's pair 0 alone: the library spreads it over a 16-CTA cluster (latency mode)."""
    pr = synth.config_pair(2, 0)
    gpu_ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    g = gpu_ctx.align_trace(0, _cfg2(capi.default_params("cvo"), capi), trace_cap=CFG2_ITERS)
    o = _oracle_cfg2(oracle, pr, trace_cap=CFG2_ITERS)
    assert gpu_ctx.last_cluster_size > 1
    assert g["n_iterations_run"] == o["n_iterations_run"] == CFG2_ITERS
    assert g["status"] == capi.STATUS_MAX_ITER and g["iters"] == CFG2_ITERS
    _check_record(g["trace"][0], o["trace"][0], tight=True)    # first iteration: identical inputs
    # last iteration: both hover around the ell = 0.10 fixed point with a small, chaotic residual flow
    gl, ol = g["trace"][-1], o["trace"][-1]
    assert abs(gl["nnz"] - ol["nnz"]) <= 1e-3 * ol["nnz"], (gl["nnz"], ol["nnz"])
    assert max(np.abs(gl["omega"]).max(), np.abs(gl["v"]).max()) < 5e-3 and abs(gl["sum_a"] - ol["sum_a"]) < 2e-4 * ol["sum_a"]
    assert all(abs(t["ell"] - CFG2_ELL) < 1e-7 for t in g["trace"])
    rot, tr = pose_diff(g["transform"], o["transform"])
    assert rot < POSE_TOL_NORTH_STAR and tr < POSE_TOL_NORTH_STAR, (rot, tr)


def test_cfg2_exact_pairs_inside_the_benchmark_batch(oracle):
    """bench.py's launch geometry: a multiple of #SMs distinct cfg-2 pairs (here 2 x, bench.py 4 x), one CTA per pair (G = 1), batched upload.  Sixteen pairs spread
    over the batch (first and last CTA wave included) are checked against the oracle (module docstring); the whole batch must
    have run exactly 100 iterations and be finite; pair 0 must agree with the single-pair cluster run."""
    probe = capi.Context(0, 64, 1)
    P = 2 * probe.num_sms
    probe.close()
    n = CFG2_POINTS
    hx, hfx = np.empty((P, n, 3), np.float32), np.empty((P, n, 5), np.float32)
    hy, hfy = np.empty((P, n, 3), np.float32), np.empty((P, n, 5), np.float32)
    for s in range(P):
        pr = synth.config_pair(2, s)
        hx[s], hfx[s], hy[s], hfy[s] = pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"]
    counts = np.full(P, n, np.int32)
    slots = np.arange(P, dtype=np.int32)
    with capi.Context(0, max_points=n + 72, max_slots=P) as ctx:
        ctx.set_pairs(slots, hx, hfx, counts, hy, hfy, counts)
        gp = _cfg2(capi.default_params("cvo"), capi)
        RT0 = np.tile(np.concatenate([np.eye(3).reshape(9), np.zeros(3)]).astype(np.float32), (P, 1))
        res = ctx.align(slots, gp, RT=RT0)  # RT in/out: the end state (R, T) comes back as well
        assert ctx.last_cluster_size == 1 and ctx.last_num_clusters == P // 2
        assert ctx.last_total_iterations == P * CFG2_ITERS
        assert np.isfinite(res["transform"]).all()
        assert (res["status"] == capi.STATUS_MAX_ITER).all() and (res["iters"] == CFG2_ITERS).all()
        worst, within = (0.0, 0.0), []
        sample = sorted(set(np.linspace(0, P - 1, 16).astype(int).tolist()))
        for s in sample:
            o = oracle.align(hx[s], hfx[s], hy[s], hfy[s], _cfg2(oracle.default_params("cvo"), oracle))
            rot, tr = pose_diff(res["transform"][s], o["transform"])
            worst = (max(worst[0], rot), max(worst[1], tr))
            within.append(rot < POSE_TOL_NORTH_STAR and tr < POSE_TOL_NORTH_STAR)
            assert rot < 5e-4 and tr < 5e-4, (s, rot, tr)
            if not within[-1]:
                _assert_same_plateau(oracle, (hx[s], hfx[s], hy[s], hfy[s]), res["RT"][s], o, CFG2_ELL, s)
        assert np.mean(within) >= 0.75, (within, worst)
        # the same pair alone (16-CTA cluster) and inside the batch (1 CTA): same pose up to f32 summation order
        ctx.set_pair(0, hx[0], hfx[0], hy[0], hfy[0])
        single = ctx.align(np.array([0], np.int32), gp)
        assert ctx.last_cluster_size > 1
        rot, tr = pose_diff(single["transform"][0], res["transform"][0])
        assert rot < POSE_TOL_NORTH_STAR and tr < POSE_TOL_NORTH_STAR, (rot, tr)
        print("cfg2 batch: worst of %d sampled pairs vs oracle: %.2e rad %.2e m; within 1e-4: %d" % (len(sample), worst[0], worst[1], int(np.sum(within))))


def test_cfg4_ragged_batch_aligned_to_convergence(oracle):
    """24 ragged cfg-4 pairs (N, M ~ U{2700..3300}), stock cvo, identity init, one batched upload, one align launch."""
    P = 24
    prs = [synth.config_pair(4, i) for i in range(P)]
    stride = max(max(len(p["x_pos"]), len(p["y_pos"])) for p in prs)
    hx, hfx = np.zeros((P, stride, 3), np.float32), np.zeros((P, stride, 5), np.float32)
    hy, hfy = np.zeros((P, stride, 3), np.float32), np.zeros((P, stride, 5), np.float32)
    nf, nm = np.zeros(P, np.int32), np.zeros(P, np.int32)
    for s, pr in enumerate(prs):
        nf[s], nm[s] = len(pr["x_pos"]), len(pr["y_pos"])
        hx[s, :nf[s]], hfx[s, :nf[s]] = pr["x_pos"], pr["x_feat"]
        hy[s, :nm[s]], hfy[s, :nm[s]] = pr["y_pos"], pr["y_feat"]
    assert len(set(nf.tolist())) > 4 and nf.min() >= 2700 and nf.max() <= 3300
    slots = np.arange(P, dtype=np.int32)
    with capi.Context(0, max_points=stride, max_slots=P) as ctx:
        ctx.set_pairs(slots, hx, hfx, nf, hy, hfy, nm)
        res = ctx.align(slots, capi.default_params("cvo"))
        assert np.isin(res["status"], (capi.STATUS_CONVERGED_TWIST, capi.STATUS_CONVERGED_UPDATE)).all()
        rots, trs = [], []
        for s, pr in enumerate(prs):
            o = oracle.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], oracle.default_params("cvo"))
            rot, tr = pose_diff(res["transform"][s], o["transform"])
            rots.append(rot)
            trs.append(tr)
            assert iters_comparable(res["iters"][s], o["iters"]), (s, res["iters"][s], o["iters"])
            rot_gt, tr_gt = pose_diff(res["transform"][s], pr["T_gt"])
            assert rot_gt < 1e-2 and tr_gt < 1e-2
        rots, trs = np.array(rots), np.array(trs)
        msg = "cfg4 batch of %d: rot median %.2e max %.2e rad; trans median %.2e max %.2e m; within 1e-4: %d/%d" % (
            P, np.median(rots), rots.max(), np.median(trs), trs.max(),
            int(((rots < POSE_TOL_NORTH_STAR) & (trs < POSE_TOL_NORTH_STAR)).sum()), P)
        print(msg)
        # every pair inside 5e-4; the bulk at north_star's 1e-4 (module docstring)
        assert rots.max() < 5e-4 and trs.max() < 5e-4, msg
        assert np.median(rots) < POSE_TOL_NORTH_STAR and np.median(trs) < POSE_TOL_NORTH_STAR, msg
        assert ((rots < POSE_TOL_NORTH_STAR) & (trs < POSE_TOL_NORTH_STAR)).mean() >= 0.75, msg


def test_cfg4_batch_fills_the_machine(oracle):
    """The per-GPU share of cfg4 on an 8-GPU node is 62-63 pairs: the launch must not leave most of the SMs idle
    (round 1 launched one CTA per pair there: 42 % of the machine)."""
    P = 63
    prs = [synth.config_pair(4, 8 * i) for i in range(P)]
    with capi.Context(0, max_points=3328, max_slots=P) as ctx:
        for s, pr in enumerate(prs):
            ctx.set_pair(s, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
        res = ctx.align(np.arange(P, dtype=np.int32), capi.default_params("cvo"))
        busy = ctx.last_cluster_size * min(P, ctx.last_num_clusters)
        assert busy >= 0.8 * ctx.num_sms, (ctx.last_cluster_size, ctx.last_num_clusters, ctx.num_sms)
        for s in (0, 31, 62):
            pr = prs[s]
            o = oracle.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], oracle.default_params("cvo"))
            rot, tr = pose_diff(res["transform"][s], o["transform"])
            assert rot < POSE_TOL_FLOOR and tr < POSE_TOL_FLOOR, (s, rot, tr)


def test_align_multi_deals_pairs_over_contexts_and_gathers_in_pair_order():
    """cvo_b200_align_multi, the native multi-GPU batch driver (BASELINE config 4): pair q -> context q mod W, a host
    thread per context, chunks of the context's slot count, results at index q.  Two contexts on ONE device exercise the
    same code path as two GPUs; the result must equal the single-context batch bit for bit (same launch geometry per
    pair is not guaranteed -- cluster sizes differ with the batch size -- so poses are compared at f32 summation noise)."""
    P = 11
    prs = [synth.config_pair(4, 100 + i) for i in range(P)]
    stride = max(max(len(p["x_pos"]), len(p["y_pos"])) for p in prs)
    hx, hfx = np.zeros((P, stride, 3), np.float32), np.zeros((P, stride, 5), np.float32)
    hy, hfy = np.zeros((P, stride, 3), np.float32), np.zeros((P, stride, 5), np.float32)
    nf, nm = np.zeros(P, np.int32), np.zeros(P, np.int32)
    for s, pr in enumerate(prs):
        nf[s], nm[s] = len(pr["x_pos"]), len(pr["y_pos"])
        hx[s, :nf[s]], hfx[s, :nf[s]] = pr["x_pos"], pr["x_feat"]
        hy[s, :nm[s]], hfy[s, :nm[s]] = pr["y_pos"], pr["y_feat"]
    gp = capi.default_params("cvo")
    a, b = capi.Context(0, max_points=stride, max_slots=2), capi.Context(0, max_points=stride, max_slots=3)
    one = capi.Context(0, max_points=stride, max_slots=P)
    try:
        multi = capi.align_multi([a, b], hx, hfx, nf, hy, hfy, nm, gp)
        one.set_pairs(np.arange(P, dtype=np.int32), hx, hfx, nf, hy, hfy, nm)
        ref = one.align(np.arange(P, dtype=np.int32), gp)
        assert (multi["kernel_ms"] > 0).all()
        assert np.isin(multi["status"], (capi.STATUS_CONVERGED_TWIST, capi.STATUS_CONVERGED_UPDATE)).all()
        for q in range(P):
            rot, tr = pose_diff(multi["transform"][q], ref["transform"][q])
            assert rot < POSE_TOL_FLOOR and tr < POSE_TOL_FLOOR, (q, rot, tr)
            rot_gt, tr_gt = pose_diff(multi["transform"][q], prs[q]["T_gt"])
            assert rot_gt < 1e-2 and tr_gt < 1e-2, q  # every pair landed at ITS OWN index
    finally:
        a.close(); b.close(); one.close()
