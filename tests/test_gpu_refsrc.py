"""The CUDA path (through the C ABI) held DIRECTLY to outputs of the reference's own first-party sources
(tests/golden/refsrc_golden.json: src/cvo.cpp, src/adaptive_cvo.cpp, src/LieGroup.cpp compiled unmodified from
/root/reference against oracle/shim -- generator tests/golden/make_refsrc_golden.py), not only to the restatement.

Tolerances: single evaluation -- counts within 2 boundary flips, sum A / omega / v / dl 1e-5 relative, step 2e-3
relative (the reference's root comes out of an f32 eigenvalue iteration); align() -- the three BASELINE configs at
north_star's 1e-4 rad / 1e-4 m, further pairs inside the noise floor of the algorithm (conftest.POSE_TOL_FLOOR);
function_inner_product 1e-5 relative (north_star)."""
import json
import os

import numpy as np
import pytest

from conftest import POSE_TOL_FLOOR, POSE_TOL_NORTH_STAR, pose_diff, iters_comparable
from cvo_rgbd_b200 import capi, frontend, synth
from test_refsrc_golden import GOLD, R0, T0, _clouds, _pair, check_eval_against_reference

pytestmark = pytest.mark.gpu


def _set(ctx, slot, pr):
    ctx.set_pair(slot, *_clouds(pr))


@pytest.mark.parametrize("name", sorted(GOLD["level1"]))
def test_device_single_evaluation_equals_reference_sources(gpu_ctx, name):
    case = GOLD["level1"][name]
    pr = _pair(case)
    _set(gpu_ctx, 0, pr)
    R, T = (np.eye(3, dtype=np.float32), np.zeros(3, np.float32)) if case.get("real") else (R0, T0)
    for ell, want in case["eval"].items():
        got = gpu_ctx.eval(0, R, T, float(ell), capi.default_params(case["kind"]))
        check_eval_against_reference(got, want, case["kind"] == "acvo", exact_counts=False)


@pytest.mark.parametrize("name", sorted(GOLD["align"]))
def test_device_align_equals_reference_align(gpu_ctx, name):
    case = GOLD["align"][name]
    pr = _pair(case)
    _set(gpu_ctx, 0, pr)
    g = gpu_ctx.align_trace(0, capi.default_params(case["kind"]), trace_cap=4)
    rot, tr = pose_diff(g["transform"], np.array(case["transform"]))
    tol = POSE_TOL_NORTH_STAR if name in ("cfg1", "cfg2_stock", "cfg3") else POSE_TOL_FLOOR
    assert rot < tol and tr < tol, (name, rot, tr, g["iters"], case["iters"])
    assert g["status"] in (capi.STATUS_CONVERGED_TWIST, capi.STATUS_CONVERGED_UPDATE)
    assert iters_comparable(g["iters"], case["iters"])
    want0 = case["first_iterations"][0]  # iteration 0: identical inputs on both sides
    got0 = g["trace"][0]
    assert abs(got0["nnz"] - want0["nnz"]) <= 2 and abs(got0["ell"] - want0["ell"]) < 1e-7
    assert np.abs(got0["omega"] - np.array(want0["omega"])).max() < 1e-5 * np.abs(want0["omega"]).max() + 2e-4 * abs(got0["nnz"] - want0["nnz"])
    assert abs(got0["step"] - want0["step"]) <= 2e-3 * want0["step"]


@pytest.mark.parametrize("name", sorted(GOLD["fixed"]))
def test_device_benchmark_schedule_equals_reference_functions(gpu_ctx, name):
    """BASELINE config 2 (fixed ell 0.10, exactly 100 iterations) against the reference's own se_kernel / compute_flow /
    compute_step_size / Exp_SEK3 driven through that schedule."""
    case = GOLD["fixed"][name]
    pr = synth.config_pair(2, case["pair_index"])
    _set(gpu_ctx, 0, pr)
    gp = capi.default_params("cvo")
    gp.ell_policy, gp.ell_init, gp.fixed_iters = capi.ELL_FIXED, 0.10, 100
    g = gpu_ctx.align_trace(0, gp, trace_cap=100)
    assert g["n_iterations_run"] == 100
    assert abs(g["trace"][0]["nnz"] - case["first"]["nnz"]) <= 2
    assert np.abs(g["trace"][0]["omega"] - np.array(case["first"]["omega"])).max() < 1e-5 * np.abs(case["first"]["omega"]).max() + 1e-9
    rot, tr = pose_diff(g["transform"], np.array(case["transform"]))
    # tolerance policy (tests/conftest.py): the BASELINE pair (seed 0) at north_star's 1e-4, extra seeds at the measured
    # noise floor of the algorithm -- two correct executions differ by up to 1.7e-4 m on 2 % of the cfg-2 pairs
    # (profiles/r02_parity_distribution.json: oracle_port_vs_oracle_ref)
    tol = POSE_TOL_NORTH_STAR if case["pair_index"] == 0 else POSE_TOL_FLOOR
    assert rot < tol and tr < tol, (name, rot, tr)


@pytest.mark.parametrize("kind", ["cvo", "acvo"])
def test_frontend_sequence_equals_reference_driver_loop(kind):
    """The Python mirror of the frontends fed like src/cvo_main.cpp:36-66 against ONE reference object fed the same
    frames through its own run_cvo(): transform and accum_transform after every frame (quirks Q3, Q4), and `iter`."""
    case = GOLD["sequence"][kind]
    n = case["n"]
    first = synth.make_pair(case["seed"], n, n, kind)
    frames = [(first["x_pos"], first["x_feat"])]
    for k in range(1, case["n_frames"]):
        pr = synth.make_pair(case["seed"], n, n, kind, motion_scale=0.6 * k)
        frames.append((pr["y_pos"], pr["y_feat"]))
    reg = (frontend.cvo if kind == "cvo" else frontend.acvo)(max_points=2048)
    try:
        for k, (xyz, feat) in enumerate(frames):
            reg.run_cvo(xyz, feat)
            rot, tr = pose_diff(reg.accum_transform, np.array(case["accum_transform"][k]))
            assert rot < 3e-4 * max(k, 1) and tr < 3e-4 * max(k, 1), (k, rot, tr)
            if k:
                rot, tr = pose_diff(reg.transform, np.array(case["transform"][k]))
                assert rot < 2 * POSE_TOL_NORTH_STAR * k and tr < 2 * POSE_TOL_NORTH_STAR * k, (k, rot, tr)
                assert iters_comparable(reg.iter, case["iter"][k])
    finally:
        reg.close()


def test_device_inner_product_equals_reference(gpu_ctx):
    case = GOLD["inner_product"]
    pr = synth.make_pair(case["seed"], case["n"], case["m"], "acvo")
    _set(gpu_ctx, 0, pr)
    for ell, want in case["values"].items():
        got = gpu_ctx.inner_product(0, float(ell), capi.default_params("acvo"))["value"]
        assert abs(got - want) < 1e-5 * abs(want), (ell, got, want)


def test_device_exp_sek3_and_root_selection_equal_reference(gpu_ctx):
    rows = np.array([c["omega"] + c["v"] + [c["dt"]] for c in GOLD["exp_sek3"]], np.float32)
    dR, dT = gpu_ctx.selftest_exp_sek3(rows)
    for i, c in enumerate(GOLD["exp_sek3"]):
        assert np.abs(dR[i] - np.array(c["dR"])).max() < 1e-6 and np.abs(dT[i] - np.array(c["dT"])).max() < 1e-6
    bcde = np.array([[c["B"], c["C"], c["D"], c["E"]] for c in GOLD["step"]])
    got = gpu_ctx.selftest_step_size(bcde)
    for g, c in zip(got, GOLD["step"]):
        assert abs(g - c["step"]) <= 2e-3 * abs(c["step"]), (c, g)


def test_device_equals_refsrc_library_on_fresh_seeds(gpu_ctx):
    """Live: the reference-source library itself (a built .so that travels to the GPU box) on seeds no fixture holds."""
    from oracle import refsrc as Rs
    if not Rs.available():
        pytest.skip("oracle/_ref/libcvo_refsrc.so absent")
    within = []
    for kind, seed, n, m in (("cvo", 91, 1800, 1700), ("acvo", 92, 1200, 1300), ("cvo", 93, 3000, 3000), ("acvo", 94, 3000, 2900)):
        pr = synth.make_pair(seed, n, m, kind)
        _set(gpu_ctx, 0, pr)
        for ell in (0.13, 0.08):
            want = Rs.evaluate(kind, *_clouds(pr), R0, T0, ell)
            got = gpu_ctx.eval(0, R0, T0, ell, capi.default_params(kind))
            check_eval_against_reference(got, want, kind == "acvo", exact_counts=False)
        a = Rs.align(kind, *_clouds(pr))  # the reference's own align()
        g = gpu_ctx.align(np.array([0]), capi.default_params(kind))
        rot, tr = pose_diff(g["transform"][0], a["transform"])
        assert rot < POSE_TOL_FLOOR and tr < POSE_TOL_FLOOR, (kind, seed, rot, tr)
        within.append(rot < POSE_TOL_NORTH_STAR and tr < POSE_TOL_NORTH_STAR)
    assert sum(within) >= 2, within
