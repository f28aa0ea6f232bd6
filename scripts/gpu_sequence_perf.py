"""Evidence: a synthetic image sequence through the whole pipeline (device image front end + align), frames per second,
against the same sequence through the CPU restatements (oracle front end + oracle align, all host threads)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import frontend, synth
from oracle import pcd_oracle as P, cvo_oracle as O
base_img, base_dep = synth.make_frame(41)
frames = [(np.roll(base_img, 2 * k, axis=1), np.roll(base_dep, 2 * k, axis=1)) for k in range(41)]
WARM = 3  # frames before the clock starts: the context's first align allocates the list scratch (about 17 ms)
for kind in ("cvo", "acvo"):
    cls = frontend.cvo if kind == "cvo" else frontend.acvo
    reg = cls(max_points=4096)
    iters = []
    for k, (img, dep) in enumerate(frames):
        if k == WARM:
            t0 = time.perf_counter()
        reg.run_cvo_images(1, img, dep)
        iters.append(reg.iter)
    dt = time.perf_counter() - t0
    reg.close()
    # the same loop with the reference driver's look-ahead: the front end of frame k + 1 overlaps the align of frame k
    reg = cls(max_points=4096)
    for k, (img, dep) in enumerate(frames):
        if k == WARM:
            t0 = time.perf_counter()
        reg.set_pcd_images(1, img, dep)
        nxt = (1,) + frames[k + 1] if k + 1 < len(frames) else None
        if k > 0:
            reg.align(next_frame=nxt)  # align_begin, front end of frame k + 1 enqueued while the kernel runs, align_finish
        elif nxt is not None:
            reg.prefetch_images(*nxt)
    dt_ahead = time.perf_counter() - t0
    reg.close()
    nt = len(frames) - WARM
    print("%s: with look-ahead (cvo_b200_align_begin / prefetch_frame_images / align_finish): %d frames in %.1f ms -> %.1f frames/s" % (kind, nt, 1e3 * dt_ahead, nt / dt_ahead), flush=True)
    # CPU: 6 frames are enough for a rate
    n_cpu = 6
    op = O.default_params(kind)
    t0 = time.perf_counter()
    prev = None
    R, T, ell = None, None, None
    for img, dep in frames[:n_cpu]:
        c = P.create_pointcloud(img, dep, 1, 1 if kind == "cvo" else 0)
        if prev is not None:
            r = O.align(prev["xyz"], prev["feat"], c["xyz"], c["feat"], op, R=R, T=T, ell=ell if kind == "cvo" else None)
            R, T, ell = r["R"], r["T"], r["ell"]
        prev = c
    dc = time.perf_counter() - t0
    print("%s: %d frames in %.1f ms -> %.1f frames/s on the device (mean %.1f iterations per pair); CPU restatement %.2f frames/s (%d threads)"
          % (kind, len(frames) - WARM, dt * 1e3, (len(frames) - WARM) / dt, float(np.mean(iters[1:])), (n_cpu - 1) / dc, O.num_threads()))
