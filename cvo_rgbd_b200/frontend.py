"""Python mirror of the reference's two frontends above the C ABI (same names, members and quirks as
include/cvo_b200_frontend.hpp; reference: cpp/rkhs_registration/include/cvo.hpp:101-107,171-192 and
include/adaptive_cvo.hpp:108-114,169-195).  Used by the tests and the benchmark driver; PyTorch-free.

Inputs are either the OUTPUT contract of the reference's image front end (pcd_generator):
xyz N x 3 and features N x 5 (row-major), or the images themselves (set_pcd_images: the front end runs on the device)."""
import numpy as np

from . import capi


class _Registration:
    _KIND = "cvo"

    def __init__(self, device=0, max_points=16384, ctx=None, slot=0, scratch_slot=1):
        self._own = ctx is None
        self._ctx = capi.Context(device, max_points, 2) if ctx is None else ctx
        self._slot, self._scratch = slot, scratch_slot
        self.params = capi.default_params(self._KIND)
        # public members of the reference classes (inc/cvo.hpp:103-107)
        self.init = False
        self.iter = 0
        self.transform = np.eye(4, dtype=np.float32)
        self.prev_transform = np.eye(4, dtype=np.float32)
        self.accum_transform = np.eye(4, dtype=np.float32)
        # private state that carries across pairs (quirk Q4): R, T, and ell for cvo
        self._RT = np.concatenate([np.eye(3).reshape(9), np.zeros(3)]).astype(np.float32)
        self._ell = float(self.params.ell_init)
        self._first = None
        self._images = False
        self._bound = False
        self._have_moving = False
        self._aligned = False  # an align() has run since the last set_pcd: the next set_pcd promotes moving -> fixed
        self._prefetched_key = None
        self.status = 0

    def close(self):
        if self._own:
            self._ctx.close()

    @property
    def ell(self):
        return self._ell

    def set_pcd(self, xyz, feat):
        """set_pcd (src/cvo.cpp:319-357): the first call stores the fixed cloud, later calls bind a moving cloud."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        feat = np.ascontiguousarray(feat, np.float32)
        if not self.init:
            self._first = (xyz, feat)
            self.init = True
            return
        if not self._bound:
            self._ctx.set_pair(self._slot, self._first[0], self._first[1], xyz, feat)
            self._bound = True
        else:
            # fixed <- moving happened at the end of align() (src/cvo.cpp:417); without an align() since the last
            # set_pcd only the moving cloud is replaced (:336-351)
            self._ctx.push_frame(self._slot, xyz, feat, promote=self._aligned)
        self._aligned = False
        if self._KIND == "acvo":  # src/adaptive_cvo.cpp:476-478
            self._ell = float(self.params.ell_init)
        self._have_moving = True

    def set_pcd_images(self, dataset_seq, img3, depth):
        """set_pcd(dataset_seq, RGB, depth, ...) (src/cvo.cpp:319-357) with the image front end on the device:
        pcd_generator::load_image + create_pointcloud(feature_type 1 for cvo, 0 for acvo; src/cvo.cpp:329,
        src/adaptive_cvo.cpp:451)."""
        if self._first is not None or (self.init and not self._images):
            raise RuntimeError("one frontend object takes either arrays or images, not both")
        self._images = True
        promote = (not self._bound) or self._aligned
        if self._prefetched_key is not None and self._prefetched_key == (id(img3), id(depth)):
            n = self._ctx.push_prefetched_frame(self._slot, promote=promote)  # the look-ahead of prefetch_images
        else:
            n = self._ctx.push_frame_images(self._slot, img3, depth, dataset_seq, 0 if self._KIND == "acvo" else 1, promote=promote)
        self._prefetched_key = None
        self._aligned = False
        if not self.init:
            self.init = True
            return n
        self._bound = True
        if self._KIND == "acvo":  # src/adaptive_cvo.cpp:476-478
            self._ell = float(self.params.ell_init)
        self._have_moving = True
        return n

    def prefetch_images(self, dataset_seq, img3, depth):
        """Look-ahead for a sequence loop (the reference's driver has frame k + 1 on disk while it aligns frame k,
        src/cvo_main.cpp:36-66): starts the front end for the frame that the NEXT set_pcd_images / run_cvo_images will be
        given (the same array objects) so that it overlaps the align() in between."""
        self._ctx.prefetch_frame_images(img3, depth, dataset_seq, 0 if self._KIND == "acvo" else 1)
        self._prefetched_key = (id(img3), id(depth))

    def run_cvo_images(self, dataset_seq, img3, depth):
        """run_cvo(dataset_seq, RGB, depth, ...) (src/cvo.cpp:422-435)."""
        first = not self.init
        self.set_pcd_images(dataset_seq, img3, depth)
        if not first:
            self.align()

    def align(self, next_frame=None):
        """align (src/cvo.cpp:361-420).  next_frame = (dataset_seq, img3, depth): the look-ahead of a sequence loop -- the
        kernel is launched, the front end of the next frame is enqueued while it runs (prefetch_images), then the
        result is collected."""
        if not self._have_moving:
            raise RuntimeError("align() called before a moving cloud was set")
        if next_frame is None:
            r = self._ctx.align([self._slot], self.params, RT=self._RT[None], ell=np.array([self._ell], np.float32))
        else:
            self._ctx.align_begin([self._slot], self.params, RT=self._RT[None], ell=np.array([self._ell], np.float32))
            self.prefetch_images(*next_frame)
            r = self._ctx.align_finish()
        self._RT, self._ell = r["RT"][0], float(r["ell"][0])
        self.status = int(r["status"][0])
        if self.status != capi.STATUS_MAX_ITER:  # Q5: iter only assigned on early exit
            self.iter = int(r["iters"][0])
        self.prev_transform = r["prev_transform"][0]                       # Q3 (src/cvo.cpp:413)
        self.accum_transform = self.accum_transform @ self.prev_transform  # :414
        self.transform = r["transform"][0]                                 # :415
        self._have_moving = False
        self._aligned = True                                               # :417

    def run_cvo(self, xyz, feat):
        """run_cvo (src/cvo.cpp:422-435)."""
        if not self.init:
            self.set_pcd(xyz, feat)
        else:
            self.set_pcd(xyz, feat)
            self.align()


class cvo(_Registration):
    """cvo::cvo"""
    _KIND = "cvo"


class acvo(_Registration):
    """acvo::acvo"""
    _KIND = "acvo"

    def function_inner_product(self, cloud_a, cloud_b):
        """acvo::function_inner_product (src/adaptive_cvo.cpp:385-439); clouds are (xyz, feat) tuples."""
        self._ctx.set_pair(self._scratch, cloud_a[0], cloud_a[1], cloud_b[0], cloud_b[1])
        return self._ctx.inner_product(self._scratch, self._ell, self.params)["value"]
