"""GPU parity tests of the image front end: cvo_b200_push_frame_images against the CPU
restatement of pcd_generator + DSO PixelSelector2 (oracle/pcd_oracle.cpp).  The bar is BIT-EXACT: the path is
integer / byte / index work plus a handful of individually rounded float operations."""
import json
import os

import numpy as np
import pytest

from conftest import pose_diff
from cvo_rgbd_b200 import capi, frontend, synth
from oracle import pcd_oracle as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcd_golden.json")


@pytest.fixture(scope="module")
def img_ctx():
    ctx = capi.Context(0, max_points=4096, max_slots=4)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("seed,texture,dataset_seq,feature_type", [(1, 1.0, 1, 1), (2, 1.0, 1, 0), (3, 0.3, 2, 1),
                                                                   (4, 0.05, 3, 0), (5, 3.0, 0, 1), (21, 1.0, 4, 1),
                                                                   (22, 0.6, 5, 0), (23, 2.0, 1, 1)])
def test_generated_cloud_equals_the_oracle_bit_for_bit(img_ctx, seed, texture, dataset_seq, feature_type):
    """Covers no re-selection (potential 3 only), re-selection with a larger potential (too many picks) and with a
    smaller one (too few), every camera model of src/pcd_generator.cpp:241-302 and both feature types."""
    img, dep = synth.make_frame(seed, texture=texture)
    want = P.create_pointcloud(img, dep, dataset_seq, feature_type)
    img_ctx.reset_slot(0)
    n = img_ctx.push_frame_images(0, img, dep, dataset_seq, feature_type)
    xyz, feat = img_ctx.last_generated_cloud()
    assert n == len(want["xyz"]) == len(xyz)
    assert np.array_equal(xyz, want["xyz"])
    assert np.array_equal(feat, want["feat"])


def test_golden_fingerprints(img_ctx):
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for case in json.load(open(GOLD)):
        img, dep = synth.make_frame(case["seed"], texture=case["texture"])
        img_ctx.reset_slot(0)
        assert img_ctx.push_frame_images(0, img, dep, case["dataset_seq"], case["feature_type"]) == case["n"]
        xyz, feat = img_ctx.last_generated_cloud()
        assert sha(xyz) == case["xyz_sha"] and sha(feat) == case["feat_sha"]


def test_other_image_sizes_and_refused_sizes(img_ctx):
    img, dep = synth.make_frame(31, w=320, h=256)
    want = P.create_pointcloud(img, dep, 1, 1)
    img_ctx.reset_slot(1)
    assert img_ctx.push_frame_images(1, img, dep, 1, 1) == len(want["xyz"])
    xyz, feat = img_ctx.last_generated_cloud()
    assert np.array_equal(xyz, want["xyz"]) and np.array_equal(feat, want["feat"])
    with pytest.raises(capi.CvoB200Error) as e:
        img_ctx.push_frame_images(1, np.zeros((100, 100, 3), np.uint8), np.ones((100, 100), np.uint16))
    assert e.value.code == capi.ERR_ARG


def test_low_texture_frames_take_the_canny_top_up_on_the_device(img_ctx):
    """The reference adds Canny edges when the selector keeps fewer than num_want/3 pixels
    (src/pcd_generator.cpp:135-163): cv::blur + cv::Canny + one extra pixel per 8 x 8 block, bit-exact here too."""
    for seed, quant in ((12, 8), (13, 4), (14, 16)):
        img, dep = synth.make_frame(seed, texture=0.0)
        img[:] = (img.astype(np.int32) // quant * quant).astype(np.uint8)
        want = P.create_pointcloud(img, dep, 1, 1)
        img_ctx.reset_slot(2)
        n = img_ctx.push_frame_images(2, img, dep, 1, 1)
        assert img_ctx.last_frame_used_canny == want["canny"]
        xyz, feat = img_ctx.last_generated_cloud()
        assert n == len(want["xyz"]) and np.array_equal(xyz, want["xyz"]) and np.array_equal(feat, want["feat"])
    assert want["canny"] or seed != 12
    n1 = img_ctx.push_frame_images(2, *synth.make_frame(2), 1, 1)  # a textured second frame: no top-up
    assert n1 > 2000 and not img_ctx.last_frame_used_canny
    r = img_ctx.align(np.array([2]), _few_iters(capi.default_params("cvo")))
    assert np.isfinite(r["transform"]).all()


def test_frame_with_too_many_points_leaves_the_bound_slot_unchanged():
    """an overflowing frame must not touch the slot's planes (for a bound slot the target buffer is the
    current FIXED cloud) nor its bookkeeping."""
    with capi.Context(0, max_points=2048, max_slots=1) as ctx:
        frames = []
        for seed, half in ((51, slice(320, None)), (52, slice(0, 320))):
            img, dep = synth.make_frame(seed)
            dep[:, half] = 0  # half of the image has no depth reading: about half of the ~3000 selected pixels survive
            frames.append((img, dep))
        n0 = ctx.push_frame_images(0, *frames[0], 1, 1)
        n1 = ctx.push_frame_images(0, *frames[1], 1, 1)
        assert 0 < n0 <= 2048 and 0 < n1 <= 2048
        before = ctx.align(np.array([0]), _few_iters(capi.default_params("cvo")))
        with pytest.raises(capi.CvoB200Error) as e:
            ctx.push_frame_images(0, *synth.make_frame(53), 1, 1)  # ~3000 points > max_points
        assert e.value.code == capi.ERR_ARG
        after = ctx.align(np.array([0]), _few_iters(capi.default_params("cvo")))
        assert np.array_equal(before["transform"], after["transform"])  # same clouds, deterministic kernel
        # and the slot still takes a frame that fits (promotion: pair becomes (frame 1, frame 0 again))
        assert ctx.push_frame_images(0, *frames[0], 1, 1) == n0
        assert np.isfinite(ctx.align(np.array([0]), _few_iters(capi.default_params("cvo")))["transform"]).all()


def _few_iters(p):
    p.fixed_iters = 3
    return p


@pytest.mark.parametrize("kind", ["cvo", "acvo"])
def test_image_sequence_through_the_frontend_equals_the_array_path(kind):
    """run_cvo on images (device front end) == run_cvo on the clouds the oracle extracts from the same images."""
    frames = []
    base_img, base_dep = synth.make_frame(41)
    for k in range(3):  # a camera panning over a static scene: shift the image and the depth by a few pixels
        frames.append((np.roll(base_img, 3 * k, axis=1), np.roll(base_dep, 3 * k, axis=1)))
    cls = frontend.cvo if kind == "cvo" else frontend.acvo
    a, b = cls(max_points=4096), cls(max_points=4096)
    try:
        for img, dep in frames:
            a.run_cvo_images(1, img, dep)
            o = P.create_pointcloud(img, dep, 1, 1 if kind == "cvo" else 0)
            b.run_cvo(o["xyz"], o["feat"])
            assert np.array_equal(a.transform, b.transform) and np.array_equal(a.accum_transform, b.accum_transform)
            assert a.iter == b.iter
        assert np.linalg.norm(a.accum_transform[:3, 3]) > 1e-4  # it registered a motion
        rot, tr = pose_diff(a.accum_transform, b.accum_transform)
        assert rot == 0 and tr == 0
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("kind", ["cvo", "acvo"])
def test_prefetched_frames_equal_the_synchronous_front_end(kind):
    """cvo_b200_prefetch_frame_images + cvo_b200_push_prefetched_frame (the front end of frame k + 1 overlapping the align of
    frame k, the look-ahead of src/cvo_main.cpp:36-66) give bit-identical clouds, hence bit-identical trajectories."""
    base_img, base_dep = synth.make_frame(43)
    frames = [(np.roll(base_img, 3 * k, axis=1).copy(), np.roll(base_dep, 3 * k, axis=1).copy()) for k in range(5)]
    cls = frontend.cvo if kind == "cvo" else frontend.acvo
    a, b = cls(max_points=4096), cls(max_points=4096)
    try:
        for k, (img, dep) in enumerate(frames):
            a.run_cvo_images(1, img, dep)
            if k == 0:
                b.set_pcd_images(1, img, dep)
            else:
                b.set_pcd_images(1, img, dep)  # (prefetched during the previous align, see below)
                if k + 1 < len(frames) and k % 2:  # both look-ahead styles: prefetch, then a blocking align ...
                    b.prefetch_images(1, *frames[k + 1])
                    b.align()
                else:  # ... and align_begin / prefetch / align_finish
                    b.align(next_frame=(1,) + frames[k + 1] if k + 1 < len(frames) else None)
            if k == 0 and len(frames) > 1:
                b.prefetch_images(1, *frames[1])
            assert np.array_equal(a.transform, b.transform) and np.array_equal(a.accum_transform, b.accum_transform), k
        assert np.linalg.norm(a.accum_transform[:3, 3]) > 1e-4
        # a synchronous frame in between discards a pending prefetch; pushing without one is an error
        b.prefetch_images(1, *frames[0])
        b.set_pcd_images(1, *frames[1])
        with pytest.raises(capi.CvoB200Error):
            b._ctx.push_prefetched_frame(0)
    finally:
        a.close()
        b.close()


def test_align_begin_finish_protocol_errors():
    """One align in flight per context: a second begin, a blocking align, or a finish without a begin are errors (and
    leave the context usable); a refused image size does not start a prefetch."""
    pr = synth.config_pair(1)
    with capi.Context(0, max_points=1024, max_slots=1) as ctx:
        ctx.set_pair(0, pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"])
        gp = capi.default_params("cvo")
        with pytest.raises(capi.CvoB200Error):
            ctx._lib.cvo_b200_align_finish  # noqa: B018  (exists)
            ctx._in_flight = (1, None, None)
            ctx.align_finish()
        ctx.align_begin([0], gp)
        with pytest.raises(capi.CvoB200Error):
            ctx.align_begin([0], gp)
        with pytest.raises(capi.CvoB200Error):
            ctx.align([0], gp)
        with pytest.raises(capi.CvoB200Error):
            ctx.prefetch_frame_images(np.zeros((100, 100, 3), np.uint8), np.zeros((100, 100), np.uint16))  # not a multiple of 32
        a = ctx.align_finish()
        b = ctx.align([0], gp)
        assert np.array_equal(a["transform"], b["transform"]) and a["iters"][0] == b["iters"][0]
        with pytest.raises(capi.CvoB200Error):
            ctx.push_prefetched_frame(0)
