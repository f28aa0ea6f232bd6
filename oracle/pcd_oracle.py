"""ctypes wrapper of oracle/libpcd_oracle.so: the CPU restatement of the reference's image front end
(pcd_generator + DSO PixelSelector2; see the header of pcd_oracle.cpp for the file:line map).
TEST INFRASTRUCTURE ONLY -- the product path never imports this.

What the reference delegates to OpenCV -- the two colour conversions, cv::blur and cv::Canny of the low-texture
top-up (src/pcd_generator.cpp:135-163) -- is restated in C (pcd_oracle.cpp) and pinned against cv2 by
tests/test_pcd_oracle.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "libpcd_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", _HERE, "libpcd_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    u8, u16, fp, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_float), C.POINTER(C.c_int)
    lib.pcd_oracle_rgb2gray.argtypes = [u8, C.c_int, u8]
    lib.pcd_oracle_rgb2hsv.argtypes = [u8, C.c_int, u8]
    lib.pcd_oracle_select.argtypes = [u8, C.c_int, C.c_int, C.c_int, fp, fp, fp, ip, ip]
    lib.pcd_oracle_blur_canny.argtypes = [u8, C.c_int, C.c_int, u8, u8]
    lib.pcd_oracle_canny_topup.argtypes = [u8, C.c_int, C.c_int, C.c_int, C.c_int, fp]
    lib.pcd_oracle_points.argtypes = [fp, u16, u8, u8, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def rgb2gray(img3):
    img3 = np.ascontiguousarray(img3, np.uint8)
    out = np.empty(img3.shape[:2], np.uint8)
    load().pcd_oracle_rgb2gray(_p(img3, C.c_uint8), out.size, _p(out, C.c_uint8))
    return out


def rgb2hsv(img3):
    img3 = np.ascontiguousarray(img3, np.uint8)
    out = np.empty_like(img3)
    load().pcd_oracle_rgb2hsv(_p(img3, C.c_uint8), img3.shape[0] * img3.shape[1], _p(out, C.c_uint8))
    return out


def blur_canny(gray):
    """(cv::blur(gray, 3x3), cv::Canny(blurred, 0, 25, 3)) restated."""
    gray = np.ascontiguousarray(gray, np.uint8)
    blurred, edge = np.empty_like(gray), np.empty_like(gray)
    load().pcd_oracle_blur_canny(_p(gray, C.c_uint8), gray.shape[1], gray.shape[0], _p(blurred, C.c_uint8), _p(edge, C.c_uint8))
    return blurred, edge


def canny_top_up(gray, sel_map, num_want, num_selected):
    """select_point's fallback for low-texture frames (src/pcd_generator.cpp:135-163). Returns True if it ran."""
    gray = np.ascontiguousarray(gray, np.uint8)
    assert sel_map.dtype == np.float32 and sel_map.flags["C_CONTIGUOUS"]
    return bool(load().pcd_oracle_canny_topup(_p(gray, C.c_uint8), gray.shape[1], gray.shape[0], num_want, num_selected,
                                              _p(sel_map, C.c_float)))


def create_pointcloud(img3, depth, dataset_seq=1, feature_type=1, num_want=3000, allow_canny=True):
    """load_image + create_pointcloud (src/pcd_generator.cpp:384-420).  img3: h x w x 3 uint8 as cv::imread returns it
    (the reference calls the channels R, G, B although they are B, G, R -- kept as is); depth: h x w uint16.
    Returns dict(xyz[n,3], feat[n,5], map[h,w], num_selected, pots, canny)."""
    lib = load()
    img3 = np.ascontiguousarray(img3, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    h, w = depth.shape
    gray, hsv = rgb2gray(img3), rgb2hsv(img3)
    sel = np.zeros((h, w), np.float32)
    dx0, dy0 = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
    pots, npots = (C.c_int * 4)(), C.c_int(0)
    n_sel = lib.pcd_oracle_select(_p(gray, C.c_uint8), w, h, num_want, _p(sel, C.c_float), _p(dx0, C.c_float),
                                  _p(dy0, C.c_float), pots, C.byref(npots))
    if n_sel < 0:
        raise ValueError("image size must be a multiple of 32 (see pcd_oracle.cpp)")
    canny = False
    if n_sel < num_want // 3:
        if not allow_canny:
            raise RuntimeError("low-texture frame: Canny top-up needed")
        canny = canny_top_up(gray, sel, num_want, n_sel)
    xyz, feat = np.zeros((w * h, 3), np.float32), np.zeros((w * h, 5), np.float32)
    n = lib.pcd_oracle_points(_p(sel, C.c_float), _p(depth, C.c_uint16), _p(img3, C.c_uint8), _p(hsv, C.c_uint8),
                              _p(dx0, C.c_float), _p(dy0, C.c_float), w, h, dataset_seq, feature_type,
                              _p(xyz, C.c_float), _p(feat, C.c_float))
    return dict(xyz=xyz[:n].copy(), feat=feat[:n].copy(), map=sel, num_selected=n_sel, pots=list(pots[:npots.value]),
                canny=canny, gray=gray, hsv=hsv)
