"""ctypes loader for the CPU parity oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product package (cvo_rgbd_b200) never does.

Two builds of the same restatement (oracle/cvo_oracle.cpp) can be loaded:
  variant="port"  oracle/libcvo_oracle.so          brute-force ball query
  variant="ref"   oracle/_ref/libcvo_oracle_ref.so ball query = the reference's own
                  nanoflann kd-tree (thirdparty/nanoflann.hpp) compiled from /root/reference
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(_HERE, "libcvo_oracle.so"),
    "ref": os.path.join(_HERE, "_ref", "libcvo_oracle_ref.so"),
}

MODE_CVO, MODE_ACVO = 0, 1
ELL_SCHEDULE, ELL_ADAPTIVE, ELL_FIXED = 0, 1, 2


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("ell_policy", C.c_int),
        ("ell_init", C.c_float), ("ell_min", C.c_float), ("ell_max", C.c_float),
        ("dl_step", C.c_double),
        ("sigma", C.c_float), ("sp_thres", C.c_float), ("c", C.c_float), ("d", C.c_float),
        ("c_ell", C.c_float), ("c_sigma", C.c_float), ("c_sp_thres", C.c_float),
        ("max_iter", C.c_int), ("min_step", C.c_float), ("max_step", C.c_float),
        ("eps", C.c_float), ("eps_2", C.c_float), ("fixed_iters", C.c_int),
    ]


class EvalOut(C.Structure):
    _fields_ = [
        ("nnz", C.c_longlong), ("sum_a", C.c_double),
        ("omega", C.c_float * 3), ("v", C.c_float * 3),
        ("B", C.c_double), ("C", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("step", C.c_float),
        ("nnz_xx", C.c_longlong), ("nnz_yy", C.c_longlong),
        ("dl", C.c_double), ("dl_num", C.c_double),
        ("n_in_ball", C.c_longlong),
    ]


class TraceRec(C.Structure):
    _fields_ = [
        ("ell", C.c_float), ("step", C.c_float),
        ("omega", C.c_float * 3), ("v", C.c_float * 3),
        ("B", C.c_double), ("C", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("sum_a", C.c_double), ("dl", C.c_double),
        ("nnz", C.c_longlong), ("nnz_xx", C.c_longlong), ("nnz_yy", C.c_longlong),
        ("R", C.c_float * 9), ("T", C.c_float * 3),
    ]


def build(variant="port", quiet=True):
    """Compile the oracle with oracle/Makefile (building the checker is not using it)."""
    target = [] if variant == "port" else ["ref"]
    subprocess.run(["make", "-C", _HERE] + target, check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_libs = {}


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        assert a.shape == shape, (a.shape, shape)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def load(variant="port"):
    if variant in _libs:
        return _libs[variant]
    path = _PATHS[variant]
    if not os.path.exists(path):
        if variant == "ref" and not os.path.isdir("/root/reference"):
            raise FileNotFoundError(path)
        build(variant)
    lib = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    lib.oracle_default_params_cvo.argtypes = [C.POINTER(Params)]
    lib.oracle_default_params_acvo.argtypes = [C.POINTER(Params)]
    lib.oracle_eval.argtypes = [fp, fp, C.c_int, fp, fp, C.c_int, fp, fp, C.c_float,
                                C.POINTER(Params), C.POINTER(EvalOut)]
    lib.oracle_eval.restype = C.c_int
    lib.oracle_align.argtypes = [fp, fp, C.c_int, fp, fp, C.c_int, C.POINTER(Params),
                                 fp, fp, fp, fp, fp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(TraceRec), C.c_int, C.POINTER(C.c_int)]
    lib.oracle_align.restype = C.c_int
    lib.oracle_inner_product.argtypes = [fp, fp, C.c_int, fp, fp, C.c_int, C.c_float, C.POINTER(Params),
                                         C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.oracle_inner_product.restype = C.c_float
    lib.oracle_exp_sek3.argtypes = [fp, fp, C.c_float, fp, fp]
    lib.oracle_step_from_coeffs.argtypes = [C.c_double] * 4 + [C.c_float] * 2
    lib.oracle_step_from_coeffs.restype = C.c_float
    lib.oracle_ball_query.argtypes = [fp, C.c_int, fp, C.c_float, C.POINTER(C.c_int), fp, C.c_int]
    lib.oracle_ball_query.restype = C.c_int
    lib.oracle_backend.restype = C.c_char_p
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_set_num_threads.argtypes = [C.c_int]
    _libs[variant] = lib
    return lib


def default_params(kind="cvo", variant="port"):
    lib = load(variant)
    p = Params()
    (lib.oracle_default_params_cvo if kind == "cvo" else lib.oracle_default_params_acvo)(C.byref(p))
    return p


def _eval_dict(o):
    return dict(nnz=int(o.nnz), sum_a=float(o.sum_a), omega=np.array(o.omega[:], np.float32),
                v=np.array(o.v[:], np.float32), B=o.B, C=o.C, D=o.D, E=o.E, step=float(o.step),
                nnz_xx=int(o.nnz_xx), nnz_yy=int(o.nnz_yy), dl=float(o.dl), dl_num=float(o.dl_num),
                n_in_ball=int(o.n_in_ball))


def evaluate(x_pos, x_feat, y_pos, y_feat, R, T, ell, params, variant="port"):
    lib = load(variant)
    x_pos, x_feat, y_pos, y_feat = _f32(x_pos), _f32(x_feat), _f32(y_pos), _f32(y_feat)
    R, T = _f32(R, (3, 3)), _f32(T, (3,))
    out = EvalOut()
    rc = lib.oracle_eval(_ptr(x_pos), _ptr(x_feat), x_pos.shape[0], _ptr(y_pos), _ptr(y_feat), y_pos.shape[0],
                         _ptr(R), _ptr(T), C.c_float(ell), C.byref(params), C.byref(out))
    if rc != 0:
        raise RuntimeError("oracle_eval failed: %d" % rc)
    return _eval_dict(out)


def align(x_pos, x_feat, y_pos, y_feat, params, R=None, T=None, ell=None, trace_cap=0, variant="port"):
    """Returns dict(R, T, ell, transform, prev_transform, iters, status, trace[list of dict])."""
    lib = load(variant)
    x_pos, x_feat, y_pos, y_feat = _f32(x_pos), _f32(x_feat), _f32(y_pos), _f32(y_feat)
    R = np.eye(3, dtype=np.float32) if R is None else _f32(R, (3, 3)).copy()
    T = np.zeros(3, np.float32) if T is None else _f32(T, (3,)).copy()
    ell_io = C.c_float(params.ell_init if ell is None else ell)
    tf = np.zeros((4, 4), np.float32)
    ptf = np.zeros((4, 4), np.float32)
    iters, status, tlen = C.c_int(0), C.c_int(0), C.c_int(0)
    tr = (TraceRec * max(trace_cap, 1))()
    rc = lib.oracle_align(_ptr(x_pos), _ptr(x_feat), x_pos.shape[0], _ptr(y_pos), _ptr(y_feat), y_pos.shape[0],
                          C.byref(params), _ptr(R), _ptr(T), C.byref(ell_io), _ptr(tf), _ptr(ptf),
                          C.byref(iters), C.byref(status), tr if trace_cap > 0 else None, trace_cap, C.byref(tlen))
    if rc != 0:
        raise RuntimeError("oracle_align failed: %d" % rc)
    trace = []
    for i in range(min(trace_cap, tlen.value)):
        r = tr[i]
        trace.append(dict(ell=float(r.ell), step=float(r.step), omega=np.array(r.omega[:], np.float32),
                          v=np.array(r.v[:], np.float32), B=r.B, C=r.C, D=r.D, E=r.E, sum_a=r.sum_a, dl=r.dl,
                          nnz=int(r.nnz), nnz_xx=int(r.nnz_xx), nnz_yy=int(r.nnz_yy),
                          R=np.array(r.R[:], np.float32).reshape(3, 3), T=np.array(r.T[:], np.float32)))
    return dict(R=R, T=T, ell=float(ell_io.value), transform=tf, prev_transform=ptf, iters=iters.value,
                status=status.value, n_iterations_run=tlen.value, trace=trace)


def inner_product(a_pos, a_feat, b_pos, b_feat, ell, params, variant="port"):
    lib = load(variant)
    a_pos, a_feat, b_pos, b_feat = _f32(a_pos), _f32(a_feat), _f32(b_pos), _f32(b_feat)
    s, n = C.c_double(0), C.c_longlong(0)
    val = lib.oracle_inner_product(_ptr(a_pos), _ptr(a_feat), a_pos.shape[0], _ptr(b_pos), _ptr(b_feat),
                                   b_pos.shape[0], C.c_float(ell), C.byref(params), C.byref(s), C.byref(n))
    return dict(value=float(val), sum_a=s.value, nnz=n.value)


def exp_sek3(omega, v, dt, variant="port"):
    lib = load(variant)
    omega, v = _f32(omega, (3,)), _f32(v, (3,))
    dR = np.zeros((3, 3), np.float32)
    dT = np.zeros(3, np.float32)
    lib.oracle_exp_sek3(_ptr(omega), _ptr(v), C.c_float(dt), _ptr(dR), _ptr(dT))
    return dR, dT


def step_from_coeffs(B, Cc, D, E, min_step=0.2, max_step=0.8, variant="port"):
    return float(load(variant).oracle_step_from_coeffs(B, Cc, D, E, min_step, max_step))


def ball_query(pts, q, r2, variant="port"):
    lib = load(variant)
    pts, q = _f32(pts), _f32(q, (3,))
    n = pts.shape[0]
    idx = np.zeros(n, np.int32)
    d2 = np.zeros(n, np.float32)
    cnt = lib.oracle_ball_query(_ptr(pts), n, _ptr(q), C.c_float(r2), idx.ctypes.data_as(C.POINTER(C.c_int)),
                                _ptr(d2), n)
    return idx[:cnt].copy(), d2[:cnt].copy()


def backend(variant="port"):
    return load(variant).oracle_backend().decode()


def num_threads(variant="port"):
    return load(variant).oracle_num_threads()


def set_num_threads(n, variant="port"):
    load(variant).oracle_set_num_threads(int(n))
