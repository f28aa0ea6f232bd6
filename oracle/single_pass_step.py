"""Prototype of SURVEY section 8f row 4, the single-pass step size (TEST INFRASTRUCTURE ONLY, NumPy; not on any product path).

compute_step_size (src/cvo.cpp:213-289) walks A a second time because its per-nonzero terms depend on the twist
(omega, v) that compute_flow has just reduced from the same A.  With d = y_j - x_i and the ROW vectors
u_i = omega x x_i + v, w_i = omega x u_i (the factorisation of cvo_rgbd_b200/csrc/cvo_quads.cuh) every term is a
polynomial of degree <= 4 in d whose coefficients depend on the row and the twist only:
    beta = b.d (b = 2t u_i)       Q = 3 |omega|^2 |d|^2 - 3 (omega.d)^2 - 4 w_i.d
    gamma = -t |u_i|^2 - t Q      delta = 2t (omega.v)(omega.d) - |omega|^2 beta      epsil = t |w_i|^2 + t |omega|^2 Q
so sum_j a_ij (term) is a contraction of the row's A-weighted MOMENT TENSORS of d,
    m0 = sum a, m1 = sum a d, m2 = sum a d(x)d, m3 = sum a d(x)d(x)d, m4 = sum a d(x)d(x)d(x)d      (1 + 3 + 6 + 10 + 15 = 35 distinct values),
which do not depend on the twist: the flow pass could accumulate them and B..E would follow without a second pass.
`moments` / `coefficients_from_moments` do exactly that; tests/test_single_pass_step.py holds the result to the two-pass
evaluation (numpy_ref.evaluate, f64) and measures what f32 moments cost.  DESIGN.md section 8 row 4 has the cost model that
keeps this off the GPU: 35 multiply-adds plus 31 monomial products per candidate and 35 floats of state per fixed point
against the 40 issue slots per candidate of the second pass it would replace."""
import numpy as np


def moments(x, y, ii, jj, a, dtype=np.float64):
    """Per fixed point i: the moment tensors of d = y_j - x_i over the nonzeros (i, j) with weights a, accumulated in `dtype`
    (products and sums both), returned as f64 arrays m0 [N], m1 [N,3], m2 [N,3,3], m3 [N,3,3,3], m4 [N,3,3,3,3]."""
    N = x.shape[0]
    d = (y[jj].astype(dtype) - x[ii].astype(dtype)).astype(dtype)
    a = a.astype(dtype)
    m0 = np.zeros(N, dtype)
    m1 = np.zeros((N, 3), dtype)
    m2 = np.zeros((N, 3, 3), dtype)
    m3 = np.zeros((N, 3, 3, 3), dtype)
    m4 = np.zeros((N, 3, 3, 3, 3), dtype)
    d2 = (d[:, :, None] * d[:, None, :]).astype(dtype)
    d3 = (d2[:, :, :, None] * d[:, None, None, :]).astype(dtype)
    d4 = (d3[:, :, :, :, None] * d[:, None, None, None, :]).astype(dtype)
    np.add.at(m0, ii, a)
    np.add.at(m1, ii, (a[:, None] * d).astype(dtype))
    np.add.at(m2, ii, (a[:, None, None] * d2).astype(dtype))
    np.add.at(m3, ii, (a[:, None, None, None] * d3).astype(dtype))
    np.add.at(m4, ii, (a[:, None, None, None, None] * d4).astype(dtype))
    return tuple(np.asarray(m, np.float64) for m in (m0, m1, m2, m3, m4))


def coefficients_from_moments(x, mom, omega, v, ell):
    """B, C, D, E of src/cvo.cpp:275-279 from the per-row moments, once the twist is known (f64)."""
    m0, m1, m2, m3, m4 = mom
    x = np.asarray(x, np.float64)
    omega = np.asarray(omega, np.float64)
    v = np.asarray(v, np.float64)
    t = 1.0 / (2.0 * ell * ell)
    u = np.cross(omega, x) + v          # [N,3]
    w = np.cross(omega, u)
    b = 2.0 * t * u
    ww = float(omega @ omega)
    k1 = 2.0 * t * float(omega @ v)
    g0 = -t * (u * u).sum(1)
    e0 = t * (w * w).sum(1)
    I = np.eye(3)
    o = omega
    # contractions (per row)
    b_m1 = np.einsum("ni,ni->n", b, m1)
    o_m1 = m1 @ o
    w_m1 = np.einsum("ni,ni->n", w, m1)
    tr2 = np.einsum("nii->n", m2)
    oo2 = np.einsum("nij,i,j->n", m2, o, o)
    bb2 = np.einsum("nij,ni,nj->n", m2, b, b)
    bo2 = np.einsum("nij,ni,j->n", m2, b, o)
    bw2 = np.einsum("nij,ni,nj->n", m2, b, w)
    ww2 = np.einsum("nij,ni,nj->n", m2, w, w)
    bI3 = np.einsum("nijj,ni->n", m3, b)
    boo3 = np.einsum("nijk,ni,j,k->n", m3, b, o, o)
    bbb3 = np.einsum("nijk,ni,nj,nk->n", m3, b, b, b)
    bbw3 = np.einsum("nijk,ni,nj,nk->n", m3, b, b, w)
    Iw3 = np.einsum("niij,nj->n", m3, w)
    oow3 = np.einsum("nijk,i,j,nk->n", m3, o, o, w)
    bbI4 = np.einsum("nijkk,ni,nj->n", m4, b, b)
    bboo4 = np.einsum("nijkl,ni,nj,k,l->n", m4, b, b, o, o)
    II4 = np.einsum("niijj->n", m4)
    oooo4 = np.einsum("nijkl,i,j,k,l->n", m4, o, o, o, o)
    Ioo4 = np.einsum("niijk,j,k->n", m4, o, o)
    bbbb4 = np.einsum("nijkl,ni,nj,nk,nl->n", m4, b, b, b, b)
    del I
    sQ = 3 * ww * tr2 - 3 * oo2 - 4 * w_m1                       # sum a Q
    sbQ = 3 * ww * bI3 - 3 * boo3 - 4 * bw2                      # sum a beta Q
    sbbQ = 3 * ww * bbI4 - 3 * bboo4 - 4 * bbw3                  # sum a beta^2 Q
    sQQ = 9 * ww * ww * II4 + 9 * oooo4 + 16 * ww2 - 18 * ww * Ioo4 - 24 * ww * Iw3 + 24 * oow3
    s_beta = b_m1
    s_gamma = g0 * m0 - t * sQ
    s_delta = k1 * o_m1 - ww * b_m1
    s_epsil = e0 * m0 + t * ww * sQ
    s_bg = g0 * b_m1 - t * sbQ
    s_bd = k1 * bo2 - ww * bb2
    s_bbg = g0 * bb2 - t * sbbQ
    s_gg = g0 * g0 * m0 - 2 * g0 * t * sQ + t * t * sQQ
    B = s_beta.sum()
    C = (s_gamma + bb2 / 2).sum()
    D = (s_delta + s_bg + bbb3 / 6).sum()
    E = (s_epsil + s_bd + s_bbg / 2 + s_gg / 2 + bbbb4 / 24).sum()
    return float(B), float(C), float(D), float(E)


def two_pass_coefficients(x, y, ii, jj, a, omega, v, ell, dtype=np.float64):
    """The reference's second traversal (src/cvo.cpp:249-289) on the same nonzeros: per-nonzero terms in `dtype`,
    sums in f64 -- the arithmetic the single pass is compared with."""
    omega = np.asarray(omega, np.float64)
    v = np.asarray(v, np.float64)
    W = np.array([[0, -omega[2], omega[1]], [omega[2], 0, -omega[0]], [-omega[1], omega[0], 0]])
    yy = np.asarray(y, np.float64)
    xiz = yy @ W.T + v
    xi2z = xiz @ W.T
    xi3z = xi2z @ W.T
    xi4z = xi3z @ W.T
    c = lambda m: m.astype(dtype)  # noqa: E731
    normxiz2 = c((c(xiz) ** 2).sum(1))
    xiz_dot_xi2z = c(-(c(xiz) * c(xi2z)).sum(1))
    epsil_const = c((c(xi2z) ** 2).sum(1) + 2 * (c(xiz) * c(xi3z)).sum(1))
    t = dtype(1.0 / (2.0 * ell * ell))
    r = c(c(x)[ii] - c(y)[jj])
    a = c(a)
    beta = c(-2 * t * (c(xiz)[jj] * r).sum(1))
    gamma = c(-t * (normxiz2[jj] + 2 * (c(xi2z)[jj] * r).sum(1)))
    delta = c(2 * t * (xiz_dot_xi2z[jj] - (c(xi3z)[jj] * r).sum(1)))
    epsil = c(-t * (epsil_const[jj] + 2 * (c(xi4z)[jj] * r).sum(1)))
    f = lambda z: np.asarray(z, np.float64)  # noqa: E731
    B = (f(a * beta)).sum()
    C = (f(a) * (f(gamma) + f(beta * beta) / 2)).sum()
    D = (f(a) * (f(delta) + f(beta * gamma) + f(beta * beta * beta) / 6)).sum()
    E = (f(a) * (f(epsil) + f(beta * delta) + 0.5 * f(beta * beta * gamma) + 0.5 * f(gamma * gamma) + f(beta * beta * beta * beta) / 24)).sum()
    return float(B), float(C), float(D), float(E)
