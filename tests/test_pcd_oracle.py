"""CPU tests of the image front end's oracle (oracle/pcd_oracle.cpp: pcd_generator + DSO PixelSelector2 restated)
and of the host pieces of the product path that need no GPU.

What pins the oracle: (1) its two colour conversions against cv2 -- the reference's own third-party dependency
(cv::cvtColor, src/pcd_generator.cpp:390-391), version 4.13 in this image, unpinned (>= 3) in the reference;
(2) the committed fingerprints tests/golden/pcd_golden.json; (3) structural properties the reference's code implies."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from cvo_rgbd_b200 import capi, synth
from oracle import pcd_oracle as P

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcd_golden.json")


def test_colour_conversions_equal_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    img[:60] = rng.integers(0, 4, (60, 640, 3))            # near black: v, diff small
    img[60:120] = 255 - rng.integers(0, 4, (60, 640, 3))   # near white
    img[120:180, :, 1] = img[120:180, :, 0]                # ties between channels (hue branches)
    grid = np.array([[r, g, b] for r in range(0, 256, 5) for g in range(0, 256, 5) for b in range(0, 256, 5)], np.uint8)
    grid = np.concatenate([grid, np.zeros(((-len(grid)) % 640, 3), np.uint8)]).reshape(-1, 640, 3)
    for im in (img, grid):
        assert np.array_equal(P.rgb2gray(im), cv2.cvtColor(im, cv2.COLOR_RGB2GRAY))
        assert np.array_equal(P.rgb2hsv(im), cv2.cvtColor(im, cv2.COLOR_RGB2HSV))


def test_random_pattern_restatement_equals_libc_rand():
    """thirdparty/PixelSelector2.cpp:36-38: srand(3141592); randomPattern[i] = rand() & 0xFF.  The library restates
    glibc's generator so as not to touch the process-wide state; here it is held against the real rand()."""
    libc = ctypes.CDLL("libc.so.6")
    for seed, n in ((3141592, 640 * 480), (1, 1000), (0, 100), (2 ** 31 + 5, 500)):
        libc.srand(ctypes.c_uint(seed))
        want = np.array([libc.rand() & 0xFF for _ in range(min(n, 20000))], np.uint8)
        got = capi.selftest_rand_bytes(seed, n)
        assert np.array_equal(got[:len(want)], want), seed


def test_oracle_matches_committed_fingerprints():
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for case in json.load(open(GOLD)):
        img, dep = synth.make_frame(case["seed"], texture=case["texture"])
        assert sha(img) == case["img_sha"] and sha(dep) == case["depth_sha"]  # the generator itself is pinned
        r = P.create_pointcloud(img, dep, case["dataset_seq"], case["feature_type"])
        assert (r["num_selected"], len(r["xyz"]), r["pots"], r["canny"]) == (case["num_selected"], case["n"], case["pots"], case["canny"])
        assert sha(r["map"].astype(np.uint8)) == case["map_sha"]
        assert sha(r["xyz"]) == case["xyz_sha"] and sha(r["feat"]) == case["feat_sha"]


def test_oracle_structure():
    img, dep = synth.make_frame(11)
    r = P.create_pointcloud(img, dep, 1, 1)
    ys, xs = np.nonzero(r["map"])
    h, w = dep.shape
    # select() only considers 4 <= x < w-5, 4 <= y <= h-4 (thirdparty/PixelSelector2.cpp:364)
    assert xs.min() >= 4 and xs.max() < w - 5 and ys.min() >= 4 and ys.max() <= h - 4
    assert set(np.unique(r["map"])) <= {0.0, 1.0, 2.0, 4.0}
    # about num_want pixels survive the sub-sampling (:226-243); zero-depth pixels are dropped afterwards (:307)
    assert 2400 < r["num_selected"] < 3800 and len(r["xyz"]) == int(((r["map"] != 0) & (dep != 0)).sum())
    # points come in raster order with the fr1 pinhole model (src/pcd_generator.cpp:304-321)
    keep = (r["map"] != 0) & (dep != 0)
    yy, xx = np.nonzero(keep)
    z = dep[keep].astype(np.float32) / np.float32(5000.0)
    assert np.array_equal(r["xyz"][:, 2], z)
    assert np.allclose(r["xyz"][:, 0], (xx - 318.6) * z / 517.3, rtol=1e-5, atol=1e-6) and np.allclose(r["xyz"][:, 1], (yy - 255.3) * z / 516.5, rtol=1e-5, atol=1e-6)
    # feature type 1 = raw channels + raw central-difference gradient of the gray image (:359-381)
    assert np.array_equal(r["feat"][:, :3], img[keep].astype(np.float32))
    g = r["gray"].astype(np.float32)
    assert np.array_equal(r["feat"][:, 3], 0.5 * (g[yy, xx + 1] - g[yy, xx - 1]))
    assert np.array_equal(r["feat"][:, 4], 0.5 * (g[yy + 1, xx] - g[yy - 1, xx]))
    # feature type 0 = HSV / [180, 255, 255] + gradient * 2 / 255 (:336-358)
    r0 = P.create_pointcloud(img, dep, 1, 0)
    assert np.array_equal(r0["xyz"], r["xyz"])
    assert np.allclose(r0["feat"][:, :3], r["hsv"][keep] / np.array([180.0, 255.0, 255.0]), atol=1e-7)
    assert np.allclose(r0["feat"][:, 3:], r["feat"][:, 3:] * 2 / 255.0, rtol=1e-6)


def test_blur_and_canny_restatements_equal_opencv():
    """cv::blur(3x3) and cv::Canny(0, 25, 3) of the low-texture top-up (src/pcd_generator.cpp:141-142)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for trial in range(5):
        g = rng.integers(0, 256, (480, 640), dtype=np.uint8)
        if trial >= 1:
            g = cv2.GaussianBlur(g, (0, 0), 1 + trial)
        if trial == 4:
            g = (g // 4 * 4).astype(np.uint8)
        b, e = P.blur_canny(g)
        wb = cv2.blur(g, (3, 3))
        assert np.array_equal(b, wb) and np.array_equal(e, cv2.Canny(wb, 0, 25, apertureSize=3))
    img, _ = synth.make_frame(7)
    gray = P.rgb2gray(img)
    b, e = P.blur_canny(gray)
    assert np.array_equal(e, cv2.Canny(cv2.blur(gray, (3, 3)), 0, 25, apertureSize=3)) and (e != 0).sum() > 1000


def test_low_texture_frame_takes_the_canny_branch():
    img, dep = synth.make_frame(12, texture=0.0)
    img[:] = (img.astype(np.int32) // 8 * 8).astype(np.uint8)  # flatten the noise: almost no gradient anywhere
    with pytest.raises(RuntimeError):
        P.create_pointcloud(img, dep, 1, 1, allow_canny=False)
    r = P.create_pointcloud(img, dep, 1, 1)
    assert r["canny"] and r["num_selected"] < 1000 and len(r["xyz"]) >= r["num_selected"] - 50


def test_image_size_must_be_a_multiple_of_32():
    with pytest.raises(ValueError):
        P.create_pointcloud(np.zeros((100, 100, 3), np.uint8), np.ones((100, 100), np.uint16))
