"""Host-side check of the algebra behind the STEP pass (cvo_rgbd_b200/csrc/cvo_onthefly.cuh: step_col, step_terms; the row-factored form of cvo_quads.cuh is checked against the same expressions).

The reference forms xi^k z as matrix powers of Omega = skew(omega) (src/cvo.cpp:229-234) and takes four dot products
with diff_xy per nonzero (:262-271).  The kernel takes three (z1.r, z2.r, omega.r) and uses
    Omega^2 z1 = omega (omega . z1) - |omega|^2 z1,  omega . z1 = omega . v   (z1 = omega x y + v)
    Omega^3 z1 = -|omega|^2 Omega z1
and folds the coefficients of gamma / delta / epsilon into per-column terms.  Here both forms are evaluated in f64 on
random inputs and must agree to rounding; the f32 behaviour of the kernel itself is covered by the GPU parity tests."""
import numpy as np


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def reference_terms(omega, v, y, x, ell):
    """beta, gamma, delta, epsilon of one nonzero as src/cvo.cpp:226-271 writes them."""
    W = skew(omega)
    xiz = np.cross(omega, y) + v
    xi2z = W @ W @ y + W @ v
    xi3z = W @ W @ W @ y + W @ W @ v
    xi4z = W @ W @ W @ W @ y + W @ W @ W @ v
    normxiz2 = xiz @ xiz
    xiz_dot_xi2z = -(xiz @ xi2z)
    epsil_const = xi2z @ xi2z + 2 * (xiz @ xi3z)
    t = 1.0 / (2 * ell * ell)
    d = x - y
    beta = -2 * t * (xiz @ d)
    gamma = -t * (normxiz2 + 2 * (xi2z @ d))
    delta = 2 * t * (xiz_dot_xi2z + (-xi3z @ d))
    epsil = -t * (epsil_const + 2 * (xi4z @ d))
    return beta, gamma, delta, epsil


def kernel_terms(omega, v, y, x, ell):
    """The same four numbers the way step_col + step_terms compute them."""
    t = 1.0 / (2 * ell * ell)
    z1 = np.cross(omega, y) + v
    z2 = np.cross(omega, z1)
    z3 = np.cross(omega, z2)
    nrm = -t * (z1 @ z1)                      # per column, pre-scaled (step_col)
    pdt = 2 * t * (-(z1 @ z2))
    ecn = -t * (z2 @ z2 + 2 * (z1 @ z3))
    ww, wv = omega @ omega, omega @ v         # per iteration
    r = x - y
    p1, p2, pw = z1 @ r, z2 @ r, omega @ r    # per nonzero
    beta = -2 * t * p1
    gamma = -2 * t * p2 + nrm
    delta = (-2 * t * wv) * pw + (2 * t * ww) * p1 + pdt
    epsil = (2 * t * ww) * p2 + ecn
    return beta, gamma, delta, epsil


def test_three_dot_products_reproduce_the_reference_terms():
    rng = np.random.default_rng(7)
    for _ in range(200):
        omega = rng.normal(size=3) * 10 ** rng.uniform(-4, -0.5)
        v = rng.normal(size=3) * 10 ** rng.uniform(-4, -0.5)
        y = rng.normal(size=3) * 2
        x = y + rng.normal(size=3) * 0.1
        ell = rng.uniform(0.03, 0.15)
        want = reference_terms(omega, v, y, x, ell)
        got = kernel_terms(omega, v, y, x, ell)
        scale = max(abs(w) for w in want) + 1e-300
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-12 * max(abs(w), 1e-6 * scale), (g, w)


def test_xiz_is_orthogonal_to_xi2z_up_to_rounding():
    """-xiz . xi2z (src/cvo.cpp:236) is zero in exact arithmetic (xi2z = omega x xiz); the kernel keeps computing it per
    column like the reference, so that its rounding residue enters delta the same way."""
    rng = np.random.default_rng(8)
    for _ in range(50):
        omega, v, y = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
        z1 = np.cross(omega, y) + v
        z2 = np.cross(omega, z1)
        assert abs(z1 @ z2) <= 1e-13 * (np.linalg.norm(z1) * np.linalg.norm(z2) + 1e-300)
