"""SURVEY section 8f row 4 (single-pass step size) as a checked CPU prototype (oracle/single_pass_step.py): the flow pass could
accumulate, per fixed point, the A-weighted moment tensors of d = y_j - x_i up to order four (35 distinct values per row);
B, C, D, E of compute_step_size (src/cvo.cpp:275-279) then follow from the twist without a second traversal of A.

Held here: (1) in f64 the moment form IS the two-pass evaluation (the algebra is right); (2) what f32 moments would cost in
accuracy, next to the f32 per-nonzero terms the reference itself uses -- the numbers DESIGN.md section 8 row 4 quotes;
(3) the step size the cubic yields from either.  The cost model (why no CUDA kernel was built from it) is in DESIGN.md."""
import os
import sys
import types

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cvo_rgbd_b200 import synth  # noqa: E402
from oracle import numpy_ref, single_pass_step as sp  # noqa: E402

P = types.SimpleNamespace(sigma=0.1, sp_thres=8e-3, c=7.0, d=7.0, c_ell=200.0, c_sigma=1.0, c_sp_thres=8e-3, min_step=0.2, max_step=0.8)


def _case(seed, n, ell):
    pr = synth.make_pair(seed, n, n, "cvo")
    ell = float(np.float32(ell))  # the length-scale is an f32 member of the reference class
    x, y = pr["x_pos"].astype(np.float64), pr["y_pos"].astype(np.float64)  # identity pose: y is the moving cloud as given
    s2 = float(np.float32(P.sigma) ** 2)
    A, keep, _, _ = numpy_ref.gram(x, pr["x_feat"], y, pr["y_feat"], ell, s2, P.sp_thres, P.c_ell, P.c_sigma, P.sp_thres)
    ii, jj = np.nonzero(keep)
    ev = numpy_ref.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], np.eye(3), np.zeros(3), ell, P)
    return x, y, ii, jj, A[ii, jj], ev, ell


def _step(B, C, D, E):
    roots = np.roots([4 * E, 3 * D, 2 * C, B])
    good = [z.real for z in roots if abs(z.imag) < 1e-12 * max(1.0, abs(z.real)) and z.real > 0]
    return min(min(good) if good else P.min_step, P.max_step)


@pytest.mark.parametrize("seed,n,ell", [(11, 400, 0.10), (12, 600, 0.15), (13, 500, 0.06)])
def test_f64_moments_reproduce_the_two_pass_coefficients(seed, n, ell):
    x, y, ii, jj, a, ev, ell = _case(seed, n, ell)
    assert len(ii) > 500
    mom = sp.moments(x, y, ii, jj, a, np.float64)       # needs A only: could ride in the flow pass
    got = sp.coefficients_from_moments(x, mom, ev["omega"], ev["v"], ell)   # needs the twist: after the flow reduction
    want = (ev["B"], ev["C"], ev["D"], ev["E"])
    for g, w, name in zip(got, want, "BCDE"):
        assert abs(g - w) <= 1e-9 * abs(w), (name, g, w)
    # and the same nonzeros through the reference's own second traversal
    two = sp.two_pass_coefficients(x, y, ii, jj, a, ev["omega"], ev["v"], ell, np.float64)
    for g, w in zip(two, want):
        assert abs(g - w) <= 1e-10 * abs(w)


def test_f32_moments_against_f32_terms():
    """The reference computes every per-nonzero term in f32 and sums in f64; a single pass would hold the MOMENTS in f32
    (35 per row, on chip).  Both are compared with the f64 coefficients: the moment form is no less accurate in B..E on
    these inputs (moments of centimetre-sized d are well conditioned; the twist enters in f64 afterwards), so numerics
    are not what rules the row out -- its cost is."""
    worst_mom, worst_two, worst_step = 0.0, 0.0, 0.0
    for seed, n, ell in [(21, 500, 0.10), (22, 500, 0.15), (23, 700, 0.06), (24, 500, 0.03)]:
        x, y, ii, jj, a, ev, ell = _case(seed, n, ell)
        want = np.array([ev["B"], ev["C"], ev["D"], ev["E"]])
        mom32 = sp.moments(x, y, ii, jj, a, np.float32)
        got_mom = np.array(sp.coefficients_from_moments(x, mom32, ev["omega"], ev["v"], ell))
        got_two = np.array(sp.two_pass_coefficients(x, y, ii, jj, a, ev["omega"], ev["v"], ell, np.float32))
        worst_mom = max(worst_mom, float(np.max(np.abs(got_mom - want) / np.abs(want))))
        worst_two = max(worst_two, float(np.max(np.abs(got_two - want) / np.abs(want))))
        worst_step = max(worst_step, abs(_step(*got_mom) - _step(*want)) / _step(*want))
    print("single pass, f32 moments: max rel. error of B..E %.2e (two passes, f32 terms: %.2e); step size %.2e" % (worst_mom, worst_two, worst_step))
    assert worst_mom < 1e-4 and worst_two < 1e-4
    assert worst_step < 1e-4
