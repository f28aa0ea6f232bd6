"""ctypes loader for oracle/_ref/libcvo_refsrc.so (TEST INFRASTRUCTURE ONLY): the reference's own first-party sources
(src/cvo.cpp, src/adaptive_cvo.cpp, src/LieGroup.cpp + its vendored nanoflann), compiled unmodified from /root/reference
against the header stand-ins of oracle/shim/ -- see oracle/refsrc_driver.cpp for what is reference code and what is not.

It exists to PIN the restatement (oracle/cvo_oracle.cpp): tests/golden/make_refsrc_golden.py runs it in THIS container
(the only place /root/reference exists) and commits its outputs as tests/golden/refsrc_golden.json; the -m "not gpu"
tests hold the restatement to those fixtures, and -- when the library is present (it travels to the GPU box as a built
.so) -- to the library itself on fresh seeds.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from .cvo_oracle import EvalOut, TraceRec, _eval_dict, _f32, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libcvo_refsrc.so")
KIND = {"cvo": 0, "acvo": 1}
MODE_REFERENCE_ALIGN, MODE_DRIVEN_TRACE, MODE_FIXED = 0, 1, 2
_lib = None


def available():
    return os.path.exists(PATH) or os.path.isdir("/root/reference")


def build(quiet=True):
    subprocess.run(["make", "-C", _HERE, "refsrc"], check=True, stdout=subprocess.DEVNULL if quiet else None)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(PATH):
        if not os.path.isdir("/root/reference"):
            raise FileNotFoundError(PATH)
        build()
    lib = C.CDLL(PATH)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.refsrc_align.argtypes = [C.c_int, fp, fp, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp,
                                 ip, ip, C.POINTER(TraceRec), C.c_int, ip]
    lib.refsrc_align.restype = C.c_int
    lib.refsrc_eval.argtypes = [C.c_int, fp, fp, C.c_int, fp, fp, C.c_int, fp, fp, C.c_float, C.POINTER(EvalOut)]
    lib.refsrc_eval.restype = C.c_int
    lib.refsrc_inner_product.argtypes = [fp, fp, C.c_int, fp, fp, C.c_int, C.c_float]
    lib.refsrc_inner_product.restype = C.c_float
    lib.refsrc_exp_sek3.argtypes = [fp, fp, C.c_float, fp, fp]
    lib.refsrc_step_from_coeffs.argtypes = [C.c_double] * 4 + [C.c_float]
    lib.refsrc_step_from_coeffs.restype = C.c_float
    lib.refsrc_run_sequence.argtypes = [C.c_int, C.c_int, fp, fp, ip, fp, fp, ip, fp]
    lib.refsrc_run_sequence.restype = C.c_int
    lib.refsrc_backend.restype = C.c_char_p
    _lib = lib
    return lib


def align(kind, x_pos, x_feat, y_pos, y_feat, mode=MODE_REFERENCE_ALIGN, fixed_iters=0, max_iter=0, R=None, T=None,
          ell=None, trace_cap=0):
    """mode 0: the reference's own align(); 1: the same loop driven with a per-iteration trace; 2: fixed ell `ell`,
    exactly `fixed_iters` iterations, no stop tests (benchmark config 2).  Returns the oracle's align() dict layout."""
    lib = load()
    x_pos, x_feat, y_pos, y_feat = _f32(x_pos), _f32(x_feat), _f32(y_pos), _f32(y_feat)
    R = np.eye(3, dtype=np.float32) if R is None else _f32(R, (3, 3)).copy()
    T = np.zeros(3, np.float32) if T is None else _f32(T, (3,)).copy()
    ell_io = C.c_float(-1.0 if ell is None else ell)
    tf, ptf = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
    iters, status, tlen = C.c_int(0), C.c_int(0), C.c_int(0)
    tr = (TraceRec * max(trace_cap, 1))()
    rc = lib.refsrc_align(KIND[kind], _ptr(x_pos), _ptr(x_feat), x_pos.shape[0], _ptr(y_pos), _ptr(y_feat), y_pos.shape[0],
                          mode, fixed_iters, max_iter, _ptr(R), _ptr(T), C.byref(ell_io), _ptr(tf), _ptr(ptf),
                          C.byref(iters), C.byref(status), tr if trace_cap > 0 else None, trace_cap, C.byref(tlen))
    if rc != 0:
        raise RuntimeError("refsrc_align failed: %d" % rc)
    trace = []
    for i in range(max(0, min(trace_cap, tlen.value))):
        r = tr[i]
        trace.append(dict(ell=float(r.ell), step=float(r.step), omega=np.array(r.omega[:], np.float32),
                          v=np.array(r.v[:], np.float32), sum_a=r.sum_a, dl=r.dl, nnz=int(r.nnz), nnz_xx=int(r.nnz_xx),
                          nnz_yy=int(r.nnz_yy), R=np.array(r.R[:], np.float32).reshape(3, 3), T=np.array(r.T[:], np.float32)))
    return dict(R=R, T=T, ell=float(ell_io.value), transform=tf, prev_transform=ptf, iters=iters.value, status=status.value,
                n_iterations_run=tlen.value, trace=trace)


def evaluate(kind, x_pos, x_feat, y_pos, y_feat, R, T, ell):
    lib = load()
    x_pos, x_feat, y_pos, y_feat = _f32(x_pos), _f32(x_feat), _f32(y_pos), _f32(y_feat)
    R, T = _f32(R, (3, 3)), _f32(T, (3,))
    out = EvalOut()
    rc = lib.refsrc_eval(KIND[kind], _ptr(x_pos), _ptr(x_feat), x_pos.shape[0], _ptr(y_pos), _ptr(y_feat), y_pos.shape[0],
                         _ptr(R), _ptr(T), C.c_float(ell), C.byref(out))
    if rc != 0:
        raise RuntimeError("refsrc_eval failed: %d" % rc)
    return _eval_dict(out)


def run_sequence(kind, frames):
    """The reference's driver loop (src/cvo_main.cpp:36-66) on one object: frames = [(xyz, feat), ...].
    Returns dict(transform[F,4,4], accum_transform[F,4,4], iter[F], ell[F]) after every run_cvo()."""
    lib = load()
    xyz = np.ascontiguousarray(np.concatenate([_f32(f[0]) for f in frames]))
    feat = np.ascontiguousarray(np.concatenate([_f32(f[1]) for f in frames]))
    counts = np.array([len(f[0]) for f in frames], np.int32)
    F = len(frames)
    tf, acc = np.zeros((F, 4, 4), np.float32), np.zeros((F, 4, 4), np.float32)
    it, ell = np.zeros(F, np.int32), np.zeros(F, np.float32)
    rc = lib.refsrc_run_sequence(KIND[kind], F, _ptr(xyz), _ptr(feat), counts.ctypes.data_as(C.POINTER(C.c_int)), _ptr(tf),
                                 _ptr(acc), it.ctypes.data_as(C.POINTER(C.c_int)), _ptr(ell))
    if rc != 0:
        raise RuntimeError("refsrc_run_sequence failed: %d" % rc)
    return dict(transform=tf, accum_transform=acc, iter=it, ell=ell)


def inner_product(a_pos, a_feat, b_pos, b_feat, ell):
    lib = load()
    a_pos, a_feat, b_pos, b_feat = _f32(a_pos), _f32(a_feat), _f32(b_pos), _f32(b_feat)
    return float(lib.refsrc_inner_product(_ptr(a_pos), _ptr(a_feat), a_pos.shape[0], _ptr(b_pos), _ptr(b_feat),
                                          b_pos.shape[0], C.c_float(ell)))


def exp_sek3(omega, v, dt):
    lib = load()
    omega, v = _f32(omega, (3,)), _f32(v, (3,))
    dR, dT = np.zeros((3, 3), np.float32), np.zeros(3, np.float32)
    lib.refsrc_exp_sek3(_ptr(omega), _ptr(v), C.c_float(dt), _ptr(dR), _ptr(dT))
    return dR, dT


def step_from_coeffs(B, Cc, D, E, min_step=0.2):
    return float(load().refsrc_step_from_coeffs(B, Cc, D, E, min_step))


def backend():
    return load().refsrc_backend().decode()
