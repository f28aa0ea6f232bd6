"""Whole-GPU mode (cvo_b200_set_group_clusters; group_allreduce in cvo_kernels.cuh): every cluster of the launch works on
the same pair, the clusters' partial sums meet in global memory.  What is computed does not change -- only which CTA
sums which rows -- so the first iteration's records agree with the one-cluster run to f32 summation order, the pose with
the oracle within the tolerance policy of tests/conftest.py, and the run is bit-deterministic.  Reference loop: one pair at a time,
src/cvo_main.cpp:36-52; BASELINE.json configs[4] is the large pair this mode exists for."""
import numpy as np
import pytest

from conftest import POSE_TOL_FLOOR, pose_diff, rel_err
from cvo_rgbd_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _fixed(p, mod, iters, ell=0.10):
    p.ell_policy, p.ell_init, p.fixed_iters = mod.ELL_FIXED, ell, iters
    return p


def _close(a, b, tol, nnz_slack=0):
    assert abs(a["nnz"] - b["nnz"]) <= nnz_slack
    for k in ("omega", "v"):
        assert rel_err(a[k], b[k]) < tol, (k, a[k], b[k])
    for k in ("B", "C", "D", "E"):
        assert abs(a[k] - b[k]) <= tol * max(abs(b[k]), 1e-12) * 10, (k, a[k], b[k])


@pytest.fixture()
def ctx():
    with capi.Context(0, max_points=10240, max_slots=2) as c:
        yield c


@pytest.mark.parametrize("kind,cfg", [("cvo", 2), ("acvo", 3)])
def test_group_mode_matches_one_cluster_and_the_oracle(ctx, oracle, kind, cfg):
    pr = synth.config_pair(cfg)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    iters = 100  # (converged: mid-trajectory poses amplify f32 summation-order differences, see test_gpu_parity level 2)
    runs = {}
    for clusters in (1, 2, 5):
        ctx.set_group_clusters(clusters)
        runs[clusters] = ctx.align_trace(0, _fixed(capi.default_params(kind), capi, iters), trace_cap=iters)
        assert ctx.last_group_clusters == clusters and ctx.last_num_clusters == clusters
        again = ctx.align_trace(0, _fixed(capi.default_params(kind), capi, iters), trace_cap=iters)
        assert np.array_equal(again["transform"], runs[clusters]["transform"])  # bit-deterministic
    o = oracle.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], _fixed(oracle.default_params(kind), oracle, iters),
                     trace_cap=iters)
    for clusters in (2, 5):
        _close(runs[clusters]["trace"][0], runs[1]["trace"][0], 2e-6)
        _close(runs[clusters]["trace"][0], o["trace"][0], 1e-5, nnz_slack=2)
        rot, tr = pose_diff(runs[clusters]["transform"], o["transform"])
        assert rot < POSE_TOL_FLOOR and tr < POSE_TOL_FLOOR, (clusters, rot, tr)  # (tolerance policy: tests/conftest.py)


def test_group_mode_is_automatic_for_a_large_pair_and_takes_pairs_in_order(ctx, oracle):
    """BASELINE config 5 (10 000 x 10 000 points): a single-pair call spreads over every cluster the device holds; a
    forced group call with two pairs takes them one after the other and equals the two single calls bit for bit."""
    big = synth.config_pair(5)
    ctx.set_pair(0, big["x_pos"], big["x_feat"], big["y_pos"], big["y_feat"])
    ctx.set_group_clusters(0)
    gp = _fixed(capi.default_params("cvo"), capi, 6)
    g = ctx.align_trace(0, gp, trace_cap=6)
    assert ctx.last_group_clusters > 1 and ctx.last_group_clusters * ctx.last_cluster_size > ctx.num_sms // 2
    o = oracle.evaluate(big["x_pos"], big["x_feat"], big["y_pos"], big["y_feat"], np.eye(3), np.zeros(3), 0.10,
                        oracle.default_params("cvo"))
    assert abs(g["trace"][0]["nnz"] - o["nnz"]) <= 2
    for k in ("omega", "v"):
        assert rel_err(g["trace"][0][k], o[k]) < 1e-5
    small = synth.config_pair(2)  # (a small pair stays on one cluster in automatic mode)
    ctx.set_pair(1, small["x_pos"], small["x_feat"], small["y_pos"], small["y_feat"])
    ctx.align(np.array([1]), gp)
    assert ctx.last_group_clusters == 1
    ctx.set_group_clusters(4)
    both = ctx.align(np.array([0, 1]), gp)
    for s in (0, 1):
        one = ctx.align(np.array([s]), gp)
        assert np.array_equal(one["transform"][0], both["transform"][s])


def test_every_cluster_rank_runs_on_its_lists(ctx):
    """Guard against a silent fallback: a CTA whose neighbour list fails falls back to on-the-fly passes with identical
    results, five times slower (this happened to every cluster rank > 0 when the quads' shared-memory addresses ignored
    the cluster's shared-memory window).  cfg2 on one 16-CTA cluster takes 2.05 ms with lists, 10.5 ms without; the
    budget sits between the two with a factor of two either side."""
    pr = synth.config_pair(2)
    ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
    ctx.set_cluster_size(16)
    ctx.set_group_clusters(1)
    gp = _fixed(capi.default_params("cvo"), capi, 100)
    best = min(ctx.align(np.array([0]), gp) and ctx.last_kernel_ms for _ in range(3))
    ctx.set_cluster_size(0)
    assert best < 4.5, best
