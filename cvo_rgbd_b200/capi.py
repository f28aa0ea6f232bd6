"""ctypes binding of libcvo_b200.so (include/cvo_b200.h).

This is the only way the Python host side reaches the GPU: there is no CPU fallback and no
PyTorch path.  If the shared library is missing or a CUDA device is unusable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVO_B200_LIB selects a tuning variant built by scripts/build_variants.py (still this repo's CUDA library)
LIB_PATH = os.environ.get("CVO_B200_LIB") or os.path.join(_HERE, "libcvo_b200.so")

MODE_CVO, MODE_ACVO = 0, 1
ELL_SCHEDULE, ELL_ADAPTIVE, ELL_FIXED = 0, 1, 2
STATUS_MAX_ITER, STATUS_CONVERGED_TWIST, STATUS_CONVERGED_UPDATE, STATUS_NAN = 0, 1, 2, 3
OK, ERR_ARG, ERR_CUDA, ERR_EMPTY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4

# every symbol include/cvo_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "cvo_b200_default_params_cvo", "cvo_b200_default_params_acvo", "cvo_b200_create", "cvo_b200_destroy",
    "cvo_b200_last_error", "cvo_b200_set_pair", "cvo_b200_set_pairs", "cvo_b200_push_frame", "cvo_b200_replace_moving", "cvo_b200_replace_moving_images", "cvo_b200_eval", "cvo_b200_align",
    "cvo_b200_align_trace", "cvo_b200_inner_product", "cvo_b200_sync", "cvo_b200_last_kernel_ms",
    "cvo_b200_kernel_launches", "cvo_b200_last_cluster_size", "cvo_b200_last_num_clusters",
    "cvo_b200_set_cluster_size", "cvo_b200_last_total_iterations", "cvo_b200_num_sms",
    "cvo_b200_set_neighbor_lists", "cvo_b200_last_list_builds", "cvo_b200_last_list_refines",
    "cvo_b200_push_frame_images", "cvo_b200_last_generated_cloud", "cvo_b200_reset_slot", "cvo_b200_selftest_rand_bytes",
    "cvo_b200_last_frame_used_canny", "cvo_b200_selftest_step_size", "cvo_b200_selftest_exp_sek3",
    "cvo_b200_neighbor_lists_active", "cvo_b200_list_scratch_bytes", "cvo_b200_align_multi", "cvo_b200_last_list_fill",
    "cvo_b200_set_group_clusters", "cvo_b200_last_group_clusters", "cvo_b200_prefetch_frame_images",
    "cvo_b200_push_prefetched_frame", "cvo_b200_align_begin", "cvo_b200_align_finish",
]


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


class IterRec(C.Structure):
    _fields_ = [
        ("ell", C.c_float), ("step", C.c_float),
        ("omega", C.c_float * 3), ("v", C.c_float * 3),
        ("B", C.c_double), ("C", C.c_double), ("D", C.c_double), ("E", C.c_double),
        ("sum_a", C.c_double), ("dl", C.c_double),
        ("nnz", C.c_longlong), ("nnz_xx", C.c_longlong), ("nnz_yy", C.c_longlong),
        ("R", C.c_float * 9), ("T", C.c_float * 3),
    ]

    def as_dict(self):
        return dict(ell=float(self.ell), step=float(self.step), omega=np.array(self.omega[:], np.float32),
                    v=np.array(self.v[:], np.float32), B=self.B, C=self.C, D=self.D, E=self.E,
                    sum_a=self.sum_a, dl=self.dl, nnz=int(self.nnz), nnz_xx=int(self.nnz_xx),
                    nnz_yy=int(self.nnz_yy), R=np.array(self.R[:], np.float32).reshape(3, 3),
                    T=np.array(self.T[:], np.float32))


class CvoB200Error(RuntimeError):
    def __init__(self, msg, code=None):
        super().__init__(msg)
        self.code = code


_lib = None


def load():
    """Loads libcvo_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvoB200Error("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
    lib.cvo_b200_default_params_cvo.argtypes = [C.POINTER(Params)]
    lib.cvo_b200_default_params_acvo.argtypes = [C.POINTER(Params)]
    lib.cvo_b200_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
    lib.cvo_b200_destroy.argtypes = [vp]
    lib.cvo_b200_destroy.restype = None
    lib.cvo_b200_last_error.argtypes = [vp]
    lib.cvo_b200_last_error.restype = C.c_char_p
    lib.cvo_b200_set_pair.argtypes = [vp, C.c_int, fp, fp, C.c_int, fp, fp, C.c_int]
    lib.cvo_b200_set_pairs.argtypes = [vp, ip, C.c_int, fp, fp, ip, fp, fp, ip, C.c_int]
    lib.cvo_b200_push_frame.argtypes = [vp, C.c_int, fp, fp, C.c_int]
    lib.cvo_b200_replace_moving.argtypes = [vp, C.c_int, fp, fp, C.c_int]
    lib.cvo_b200_eval.argtypes = [vp, C.c_int, fp, fp, C.c_float, C.POINTER(Params), C.POINTER(IterRec)]
    lib.cvo_b200_align.argtypes = [vp, ip, C.c_int, C.POINTER(Params), fp, fp, fp, fp, ip, ip]
    lib.cvo_b200_align_begin.argtypes = [vp, ip, C.c_int, C.POINTER(Params), fp, fp]
    lib.cvo_b200_align_finish.argtypes = [vp, fp, fp, fp, fp, ip, ip]
    lib.cvo_b200_align_trace.argtypes = [vp, C.c_int, C.POINTER(Params), fp, fp, fp, fp, ip, ip,
                                         C.POINTER(IterRec), C.c_int, ip]
    lib.cvo_b200_inner_product.argtypes = [vp, C.c_int, C.c_float, C.POINTER(Params), fp,
                                           C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    lib.cvo_b200_sync.argtypes = [vp]
    lib.cvo_b200_last_kernel_ms.argtypes = [vp]
    lib.cvo_b200_last_kernel_ms.restype = C.c_float
    lib.cvo_b200_kernel_launches.argtypes = [vp]
    lib.cvo_b200_kernel_launches.restype = C.c_longlong
    lib.cvo_b200_last_cluster_size.argtypes = [vp]
    lib.cvo_b200_last_num_clusters.argtypes = [vp]
    lib.cvo_b200_set_cluster_size.argtypes = [vp, C.c_int]
    lib.cvo_b200_set_group_clusters.argtypes = [vp, C.c_int]
    lib.cvo_b200_last_group_clusters.argtypes = [vp]
    lib.cvo_b200_last_total_iterations.argtypes = [vp]
    lib.cvo_b200_last_total_iterations.restype = C.c_longlong
    lib.cvo_b200_num_sms.argtypes = [vp]
    lib.cvo_b200_push_frame_images.argtypes = [vp, C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_ushort), C.c_int, C.c_int,
                                               C.c_int, C.c_int, ip]
    lib.cvo_b200_replace_moving_images.argtypes = lib.cvo_b200_push_frame_images.argtypes
    lib.cvo_b200_prefetch_frame_images.argtypes = [vp, C.POINTER(C.c_ubyte), C.POINTER(C.c_ushort), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.cvo_b200_push_prefetched_frame.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.cvo_b200_align_multi.argtypes = [C.POINTER(vp), C.c_int, C.c_int, fp, fp, ip, fp, fp, ip, C.c_int, C.POINTER(Params), fp, ip,
                                         ip, fp]
    lib.cvo_b200_last_list_fill.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.cvo_b200_neighbor_lists_active.argtypes = [vp]
    lib.cvo_b200_list_scratch_bytes.argtypes = [vp]
    lib.cvo_b200_list_scratch_bytes.restype = C.c_longlong
    lib.cvo_b200_last_generated_cloud.argtypes = [vp, fp, fp, C.c_int, ip]
    lib.cvo_b200_reset_slot.argtypes = [vp, C.c_int]
    lib.cvo_b200_last_frame_used_canny.argtypes = [vp]
    lib.cvo_b200_selftest_rand_bytes.argtypes = [C.c_uint, C.c_int, C.POINTER(C.c_ubyte)]
    lib.cvo_b200_selftest_step_size.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_float, C.c_float, fp]
    lib.cvo_b200_selftest_exp_sek3.argtypes = [vp, fp, C.c_int, fp]
    lib.cvo_b200_set_neighbor_lists.argtypes = [vp, C.c_int, C.c_float]
    lib.cvo_b200_last_list_builds.argtypes = [vp]
    lib.cvo_b200_last_list_builds.restype = C.c_longlong
    lib.cvo_b200_last_list_refines.argtypes = [vp]
    lib.cvo_b200_last_list_refines.restype = C.c_longlong
    _lib = lib
    return lib


def align_multi(contexts, fx, ff, n_fixed, mx, mf, n_moving, params):
    """cvo_b200_align_multi: pairs q -> contexts[q mod len(contexts)], one host thread per context (= per GPU).
    fx/mx: [P, stride, 3], ff/mf: [P, stride, 5] C-contiguous float32.  Returns dict(transform, iters, status, kernel_ms)."""
    lib = load()
    n_fixed = np.ascontiguousarray(n_fixed, dtype=np.int32)
    n_moving = np.ascontiguousarray(n_moving, dtype=np.int32)
    P, stride = fx.shape[0], fx.shape[1]
    for a, w in ((fx, 3), (ff, 5), (mx, 3), (mf, 5)):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.shape == (P, stride, w)
    hs = (C.c_void_p * len(contexts))(*[c._h.value for c in contexts])
    tf = np.zeros((P, 4, 4), np.float32)
    iters, status = np.zeros(P, np.int32), np.zeros(P, np.int32)
    ms = np.zeros(len(contexts), np.float32)
    rc = lib.cvo_b200_align_multi(hs, len(contexts), P, _fp(fx), _fp(ff), _ipt(n_fixed), _fp(mx), _fp(mf), _ipt(n_moving), stride,
                                  C.byref(params), _fp(tf), _ipt(iters), _ipt(status), _fp(ms))
    if rc != OK:
        raise CvoB200Error("cvo_b200_align_multi failed: %d (%s)" % (rc, "; ".join(lib.cvo_b200_last_error(c._h).decode() for c in contexts)), rc)
    return dict(transform=tf, iters=iters, status=status, kernel_ms=ms)


def selftest_rand_bytes(seed, n):
    """`rand() & 0xFF` of glibc after srand(seed), from the library's own restatement (host code, no GPU needed)."""
    out = np.zeros(n, np.uint8)
    rc = load().cvo_b200_selftest_rand_bytes(seed, n, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    if rc != OK:
        raise CvoB200Error("selftest_rand_bytes failed", rc)
    return out


def default_params(kind="cvo"):
    p = Params()
    lib = load()
    (lib.cvo_b200_default_params_cvo if kind == "cvo" else lib.cvo_b200_default_params_acvo)(C.byref(p))
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ipt(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Context:
    """One GPU context (cvo_b200_ctx): owns the device buffers of `max_slots` frame pairs."""

    def __init__(self, device=0, max_points=4096, max_slots=1):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.cvo_b200_create(C.byref(self._h), device, max_points, max_slots)
        if rc != OK:
            self._h = None
            raise CvoB200Error("cvo_b200_create failed (%d): no usable CUDA device / out of memory" % rc)
        self.max_points, self.max_slots, self.device = max_points, max_slots, device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cvo_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise CvoB200Error("libcvo_b200 error %d: %s" % (rc, self._lib.cvo_b200_last_error(self._h).decode()), rc)

    def set_pair(self, slot, fixed_xyz, fixed_feat, moving_xyz, moving_feat):
        fx, ff, mx, mf = _f32(fixed_xyz), _f32(fixed_feat), _f32(moving_xyz), _f32(moving_feat)
        assert fx.ndim == 2 and fx.shape[1] == 3 and ff.shape == (fx.shape[0], 5)
        assert mx.ndim == 2 and mx.shape[1] == 3 and mf.shape == (mx.shape[0], 5)
        self._check(self._lib.cvo_b200_set_pair(self._h, slot, _fp(fx), _fp(ff), fx.shape[0], _fp(mx), _fp(mf),
                                                mx.shape[0]))

    def set_pair_raw(self, slot, fx, ff, mx, mf):
        """No conversions: arrays must already be C-contiguous float32 (benchmark path)."""
        return self._lib.cvo_b200_set_pair(self._h, slot, _fp(fx), _fp(ff), fx.shape[0], _fp(mx), _fp(mf),
                                           mx.shape[0])

    def set_pairs(self, slots, fx, ff, n_fixed, mx, mf, n_moving):
        """Batched upload: fx/mx are [P, stride, 3], ff/mf are [P, stride, 5] C-contiguous float32 (no conversion
        is done: pass pinned arrays for the benchmark path); n_fixed/n_moving are the per-pair point counts."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        n_fixed = np.ascontiguousarray(n_fixed, dtype=np.int32)
        n_moving = np.ascontiguousarray(n_moving, dtype=np.int32)
        P, stride = slots.shape[0], fx.shape[1]
        for a, w in ((fx, 3), (ff, 5), (mx, 3), (mf, 5)):
            assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.shape == (P, stride, w)
        assert n_fixed.shape == (P,) and n_moving.shape == (P,)
        self._check(self._lib.cvo_b200_set_pairs(self._h, _ipt(slots), P, _fp(fx), _fp(ff), _ipt(n_fixed), _fp(mx),
                                                 _fp(mf), _ipt(n_moving), stride))

    def push_frame(self, slot, xyz, feat, promote=True):
        """promote=True: moving becomes fixed, then the new moving cloud (src/cvo.cpp:417 + the next set_pcd);
        promote=False: the moving cloud is replaced (a second set_pcd without an align, cvo_b200_replace_moving)."""
        x, f = _f32(xyz), _f32(feat)
        fn = self._lib.cvo_b200_push_frame if promote else self._lib.cvo_b200_replace_moving
        self._check(fn(self._h, slot, _fp(x), _fp(f), x.shape[0]))

    def push_frame_images(self, slot, img3, depth, dataset_seq=1, feature_type=1, promote=True):
        """Image front door: h x w x 3 uint8 (as cv::imread returns it) + h x w uint16 depth -> the slot's next cloud,
        generated on the device.  Returns the number of points.  promote: see push_frame."""
        img3 = np.ascontiguousarray(img3, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        assert img3.ndim == 3 and img3.shape[2] == 3 and depth.shape == img3.shape[:2]
        n = C.c_int(0)
        fn = self._lib.cvo_b200_push_frame_images if promote else self._lib.cvo_b200_replace_moving_images
        self._check(fn(
            self._h, slot, img3.ctypes.data_as(C.POINTER(C.c_ubyte)), depth.ctypes.data_as(C.POINTER(C.c_ushort)),
            img3.shape[1], img3.shape[0], dataset_seq, feature_type, C.byref(n)))
        return n.value

    def prefetch_frame_images(self, img3, depth, dataset_seq=1, feature_type=1):
        """Starts the front end for the NEXT frame on the copy stream (overlaps a running align()); the arrays are kept
        alive until push_prefetched_frame."""
        img3 = np.ascontiguousarray(img3, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        assert img3.ndim == 3 and img3.shape[2] == 3 and depth.shape == img3.shape[:2]
        self._prefetched = (img3, depth)
        self._check(self._lib.cvo_b200_prefetch_frame_images(
            self._h, img3.ctypes.data_as(C.POINTER(C.c_ubyte)), depth.ctypes.data_as(C.POINTER(C.c_ushort)),
            img3.shape[1], img3.shape[0], dataset_seq, feature_type))

    def push_prefetched_frame(self, slot, promote=True):
        n = C.c_int(0)
        self._check(self._lib.cvo_b200_push_prefetched_frame(self._h, slot, int(bool(promote)), C.byref(n)))
        self._prefetched = None
        return n.value

    def last_generated_cloud(self):
        """(xyz[n,3], feat[n,5]) of the last push_frame_images, in the reference's raster order."""
        n = C.c_int(0)
        self._check(self._lib.cvo_b200_last_generated_cloud(self._h, None, None, 0, C.byref(n)))
        xyz, feat = np.zeros((n.value, 3), np.float32), np.zeros((n.value, 5), np.float32)
        self._check(self._lib.cvo_b200_last_generated_cloud(self._h, _fp(xyz), _fp(feat), n.value, C.byref(n)))
        return xyz, feat

    @property
    def last_frame_used_canny(self):
        return bool(self._lib.cvo_b200_last_frame_used_canny(self._h))

    def reset_slot(self, slot):
        self._check(self._lib.cvo_b200_reset_slot(self._h, slot))

    def eval(self, slot, R, T, ell, params):
        R, T = _f32(R).reshape(3, 3), _f32(T).reshape(3)
        rec = IterRec()
        self._check(self._lib.cvo_b200_eval(self._h, slot, _fp(R), _fp(T), C.c_float(ell), C.byref(params),
                                            C.byref(rec)))
        return rec.as_dict()

    def align(self, slots, params, RT=None, ell=None):
        """Returns dict(RT, ell, transform[P,4,4], prev_transform, iters, status)."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        P = slots.shape[0]
        RT_io = None if RT is None else _f32(RT).reshape(P, 12).copy()
        ell_io = None if ell is None else _f32(ell).reshape(P).copy()
        tf = np.zeros((P, 4, 4), np.float32)
        ptf = np.zeros((P, 4, 4), np.float32)
        iters = np.zeros(P, np.int32)
        status = np.zeros(P, np.int32)
        self._check(self._lib.cvo_b200_align(self._h, _ipt(slots), P, C.byref(params),
                                             None if RT_io is None else _fp(RT_io),
                                             None if ell_io is None else _fp(ell_io),
                                             _fp(tf), _fp(ptf), _ipt(iters), _ipt(status)))
        return dict(RT=RT_io, ell=ell_io, transform=tf, prev_transform=ptf, iters=iters, status=status)

    def align_begin(self, slots, params, RT=None, ell=None):
        """Launches the align and returns; align_finish() collects it.  In between: prefetch_frame_images."""
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        P = slots.shape[0]
        RT_io = None if RT is None else _f32(RT).reshape(P, 12).copy()
        ell_io = None if ell is None else _f32(ell).reshape(P).copy()
        self._check(self._lib.cvo_b200_align_begin(self._h, _ipt(slots), P, C.byref(params),
                                                   None if RT_io is None else _fp(RT_io),
                                                   None if ell_io is None else _fp(ell_io)))
        self._in_flight = (P, RT_io, ell_io)

    def align_finish(self):
        P, RT_io, ell_io = self._in_flight
        self._in_flight = None
        tf = np.zeros((P, 4, 4), np.float32)
        ptf = np.zeros((P, 4, 4), np.float32)
        iters = np.zeros(P, np.int32)
        status = np.zeros(P, np.int32)
        if RT_io is None:
            RT_io = np.zeros((P, 12), np.float32)
        if ell_io is None:
            ell_io = np.zeros(P, np.float32)
        self._check(self._lib.cvo_b200_align_finish(self._h, _fp(RT_io), _fp(ell_io), _fp(tf), _fp(ptf), _ipt(iters), _ipt(status)))
        return dict(RT=RT_io, ell=ell_io, transform=tf, prev_transform=ptf, iters=iters, status=status)

    def align_trace(self, slot, params, R=None, T=None, ell=None, trace_cap=2048):
        RT = np.zeros(12, np.float32)
        RT[:9] = np.eye(3, dtype=np.float32).reshape(9) if R is None else _f32(R).reshape(9)
        if T is not None:
            RT[9:] = _f32(T).reshape(3)
        ell_io = np.array([params.ell_init if ell is None else ell], np.float32)
        tf = np.zeros((4, 4), np.float32)
        ptf = np.zeros((4, 4), np.float32)
        iters, status, tlen = C.c_int(0), C.c_int(0), C.c_int(0)
        tr = (IterRec * max(1, trace_cap))()
        self._check(self._lib.cvo_b200_align_trace(self._h, slot, C.byref(params), _fp(RT), _fp(ell_io), _fp(tf),
                                                   _fp(ptf), C.byref(iters), C.byref(status), tr, trace_cap,
                                                   C.byref(tlen)))
        n = min(trace_cap, tlen.value)
        return dict(R=RT[:9].reshape(3, 3).copy(), T=RT[9:].copy(), ell=float(ell_io[0]), transform=tf,
                    prev_transform=ptf, iters=iters.value, status=status.value, n_iterations_run=tlen.value,
                    trace=[tr[i].as_dict() for i in range(n)])

    def inner_product(self, slot, ell, params):
        val, s, n = C.c_float(0), C.c_double(0), C.c_longlong(0)
        self._check(self._lib.cvo_b200_inner_product(self._h, slot, C.c_float(ell), C.byref(params), C.byref(val),
                                                     C.byref(s), C.byref(n)))
        return dict(value=float(val.value), sum_a=s.value, nnz=n.value)

    def sync(self):
        self._check(self._lib.cvo_b200_sync(self._h))

    def set_cluster_size(self, g):
        self._check(self._lib.cvo_b200_set_cluster_size(self._h, g))

    def set_group_clusters(self, n):
        """Whole-GPU mode: clusters per pair (0 automatic, 1 off, n > 1 forced); see include/cvo_b200.h."""
        self._check(self._lib.cvo_b200_set_group_clusters(self._h, n))

    def set_neighbor_lists(self, enable=True, skin=0.10):
        self._check(self._lib.cvo_b200_set_neighbor_lists(self._h, int(bool(enable)), C.c_float(skin)))

    @property
    def last_list_builds(self):
        return int(self._lib.cvo_b200_last_list_builds(self._h))

    def selftest_step_size(self, bcde, min_step=0.2, max_step=0.8):
        """The device's line search (src/cvo.cpp:53-69,291-307) on rows {B, C, D, E} of `bcde` (test hook)."""
        bcde = np.ascontiguousarray(bcde, dtype=np.float64).reshape(-1, 4)
        out = np.zeros(len(bcde), np.float32)
        self._check(self._lib.cvo_b200_selftest_step_size(self._h, bcde.ctypes.data_as(C.POINTER(C.c_double)), len(bcde),
                                                          C.c_float(min_step), C.c_float(max_step),
                                                          out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def selftest_exp_sek3(self, omega_v_dt):
        """The device's Exp_SEK3 (src/LieGroup.cpp:159-186) on rows {omega, v, dt}; returns (dR [n,3,3], dT [n,3])."""
        rows = np.ascontiguousarray(omega_v_dt, dtype=np.float32).reshape(-1, 7)
        out = np.zeros((len(rows), 12), np.float32)
        self._check(self._lib.cvo_b200_selftest_exp_sek3(self._h, rows.ctypes.data_as(C.POINTER(C.c_float)), len(rows),
                                                         out.ctypes.data_as(C.POINTER(C.c_float))))
        return out[:, :9].reshape(-1, 3, 3), out[:, 9:]

    @property
    def last_list_fill(self):
        """(candidates kept, quad slots holding them) of the (x, y) lists built by the last align call."""
        e, sl = C.c_longlong(0), C.c_longlong(0)
        self._check(self._lib.cvo_b200_last_list_fill(self._h, C.byref(e), C.byref(sl)))
        return e.value, sl.value

    @property
    def neighbor_lists_active(self):
        return bool(self._lib.cvo_b200_neighbor_lists_active(self._h))

    @property
    def list_capacity(self):
        """Bytes of HBM scratch the neighbour lists occupy right now."""
        return int(self._lib.cvo_b200_list_scratch_bytes(self._h))

    @property
    def last_list_refines(self):
        return int(self._lib.cvo_b200_last_list_refines(self._h))

    @property
    def last_kernel_ms(self):
        return float(self._lib.cvo_b200_last_kernel_ms(self._h))

    @property
    def kernel_launches(self):
        return int(self._lib.cvo_b200_kernel_launches(self._h))

    @property
    def last_cluster_size(self):
        return int(self._lib.cvo_b200_last_cluster_size(self._h))

    @property
    def last_group_clusters(self):
        return int(self._lib.cvo_b200_last_group_clusters(self._h))

    @property
    def last_num_clusters(self):
        return int(self._lib.cvo_b200_last_num_clusters(self._h))

    @property
    def last_total_iterations(self):
        return int(self._lib.cvo_b200_last_total_iterations(self._h))

    @property
    def num_sms(self):
        return int(self._lib.cvo_b200_num_sms(self._h))
