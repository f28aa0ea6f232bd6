"""Independent NumPy float64 restatement of ONE evaluation of the hot path (TEST INFRASTRUCTURE ONLY).

Purpose: pin oracle/cvo_oracle.cpp.  It is written from the maths of the reference
(src/cvo.cpp:99-308, src/adaptive_cvo.cpp:92-272, and the MATLAB statement
matlab/@rkhs_se3_registration/rkhs_se3_registration.m:120-197 for compute_flow /
compute_step_size), dense N x M, no kd-tree, no CSR, all f64 -- so a bug shared with the C++
restatement would have to be made twice, in two different formulations.  It is NOT bit-faithful
(the reference is f32): comparisons use relative tolerances and a small slack on nnz.
"""
import numpy as np


def _skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def thresholds(ell, s2, sp_thres, c_ell, c_sigma, c_gate):
    # f32 log as in the reference (std::log(float)), rest in f64, result stored as f32
    d2_thres = np.float32(-2.0 * float(ell) * float(ell) * float(np.log(np.float32(sp_thres) / np.float32(s2))))
    r = np.float32(c_gate) / np.float32(c_sigma) / np.float32(c_sigma)
    d2_c_thres = np.float32(-2.0 * float(c_ell) * float(c_ell) * float(np.log(np.float32(r))))
    return float(d2_thres), float(d2_c_thres)


def gram(a_pos, a_feat, b_pos, b_feat, ell, s2, sp_thres, c_ell, c_sigma, c_gate):
    """Dense affinity matrix with the three strict gates of se_kernel (src/cvo.cpp:143-153)."""
    a_pos = np.asarray(a_pos, np.float64); b_pos = np.asarray(b_pos, np.float64)
    a_feat = np.asarray(a_feat, np.float64); b_feat = np.asarray(b_feat, np.float64)
    d2_thres, d2_c_thres = thresholds(ell, s2, sp_thres, c_ell, c_sigma, c_gate)
    d2 = ((a_pos[:, None, :] - b_pos[None, :, :]) ** 2).sum(-1)
    in_ball = d2 < d2_thres
    d2c = np.zeros_like(d2)
    ii, jj = np.nonzero(in_ball)
    d2c[ii, jj] = ((a_feat[ii] - b_feat[jj]) ** 2).sum(-1)
    k = s2 * np.exp(-d2 / (2.0 * ell * ell))
    ck = c_sigma * c_sigma * np.exp(-d2c / (2.0 * c_ell * c_ell))
    a = ck * k
    keep = in_ball & (d2c < d2_c_thres) & (a > sp_thres)
    return np.where(keep, a, 0.0), keep, d2, int(in_ball.sum())


def evaluate(x_pos, x_feat, y_pos, y_feat, R, T, ell, p, acvo=False):
    """p: any object with sigma, sp_thres, c, d, c_ell, c_sigma, c_sp_thres, min_step, max_step."""
    x = np.asarray(x_pos, np.float64); y0 = np.asarray(y_pos, np.float64)
    R = np.asarray(R, np.float64); T = np.asarray(T, np.float64)
    ell = float(np.float32(ell))
    s2 = float(np.float32(p.sigma) * np.float32(p.sigma))
    sp = float(np.float32(p.sp_thres))
    c_gate = float(np.float32(p.c_sp_thres if acvo else p.sp_thres))
    c_ell, c_sigma = float(np.float32(p.c_ell)), float(np.float32(p.c_sigma))
    # update_tf + transform_pcd: y = R^T (y0 - T)
    y = (y0 - T) @ R
    A, keep, d2, n_in_ball = gram(x, x_feat, y, y_feat, ell, s2, sp, c_ell, c_sigma, c_gate)
    # compute_flow: omega = 1/c sum A_ij x_i x y_j ; v = 1/d sum A_ij (y_j - x_i)
    Ay = A @ y                      # sum_j A_ij y_j
    rowsum = A.sum(1)
    omega = np.cross(x, Ay).sum(0) / float(p.c)
    v = (Ay - rowsum[:, None] * x).sum(0) / float(p.d)
    out = dict(nnz=int(keep.sum()), sum_a=float(A.sum()), omega=omega, v=v, n_in_ball=n_in_ball)

    # compute_step_size in the MATLAB formulation (rkhs_se3_registration.m:149-197)
    W = _skew(omega)
    xiz = y @ W.T + v
    xi2z = xiz @ W.T            # W(Wy+v)
    xi3z = xi2z @ W.T
    xi4z = xi3z @ W.T
    normxiz2 = (xiz ** 2).sum(1)
    xiz_dot_xi2z = -(xiz * xi2z).sum(1)
    epsil_const = (xi2z ** 2).sum(1) + 2 * (xiz * xi3z).sum(1)
    t = 1.0 / (2.0 * ell * ell)
    ii, jj = np.nonzero(keep)
    r = x[ii] - y[jj]
    a = A[ii, jj]
    beta = -2 * t * (xiz[jj] * r).sum(1)
    gamma = -t * (normxiz2[jj] + 2 * (xi2z[jj] * r).sum(1))
    delta = 2 * t * (xiz_dot_xi2z[jj] - (xi3z[jj] * r).sum(1))
    epsil = -t * (epsil_const[jj] + 2 * (xi4z[jj] * r).sum(1))
    B = float((a * beta).sum())
    Cc = float((a * (gamma + beta ** 2 / 2)).sum())
    D = float((a * (delta + beta * gamma + beta ** 3 / 6)).sum())
    E = float((a * (epsil + beta * delta + 0.5 * beta ** 2 * gamma + 0.5 * gamma ** 2 + beta ** 4 / 24)).sum())
    out.update(B=B, C=Cc, D=D, E=E)
    roots = np.roots([4 * E, 3 * D, 2 * Cc, B]) if E != 0 else np.array([])
    good = [z.real for z in roots if abs(z.imag) < 1e-12 * max(1.0, abs(z.real)) and z.real > 0]
    step = min(good) if good else float(p.min_step)
    out["step"] = min(step, float(p.max_step))

    if acvo:
        Axx, kxx, d2xx, _ = gram(x, x_feat, x, x_feat, ell, s2, sp, c_ell, c_sigma, c_gate)
        Ayy, kyy, d2yy, _ = gram(y, y_feat, y, y_feat, ell, s2, sp, c_ell, c_sigma, c_gate)
        N, M = x.shape[0], y.shape[0]
        s_xy = (A * d2).sum()
        s_xx = (Axx * d2xx).sum()
        s_yy = (Ayy[N:] * d2yy[N:]).sum() if M > N else 0.0   # quirk Q1 (src/adaptive_cvo.cpp:213-223,243-265)
        dl_num = (s_xx - 2 * s_xy + s_yy) / ell ** 3
        den = int(kxx.sum()) + int(kyy.sum()) - 2 * int(keep.sum())
        out.update(nnz_xx=int(kxx.sum()), nnz_yy=int(kyy.sum()), dl_num=float(dl_num),
                   dl=float(dl_num / den) if den != 0 else float("nan"))
    return out


def exp_se3(omega, v, s):
    """exp(s * [w^ v; 0 0]) by scipy-free series (f64)."""
    M = np.zeros((4, 4))
    M[:3, :3] = _skew(np.asarray(omega, np.float64))
    M[:3, 3] = np.asarray(v, np.float64)
    M *= s
    out = np.eye(4); term = np.eye(4)
    for k in range(1, 30):
        term = term @ M / k
        out = out + term
    return out
