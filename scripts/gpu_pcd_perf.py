"""Scratch: time of the device image front end per frame vs the CPU restatement."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
from oracle import pcd_oracle as P
frames = [synth.make_frame(100 + i) for i in range(8)]
ctx = capi.Context(0, max_points=4096, max_slots=1)
for img, dep in frames[:3]:
    ctx.push_frame_images(0, img, dep, 1, 1)
t0 = time.perf_counter()
R = 10
for r in range(R):
    for img, dep in frames:
        ctx.push_frame_images(0, img, dep, 1, 1)
dt = (time.perf_counter() - t0) / (R * len(frames))
t0 = time.perf_counter()
for img, dep in frames:
    P.create_pointcloud(img, dep, 1, 1)
dc = (time.perf_counter() - t0) / len(frames)
print("device front end: %.3f ms/frame (H2D of 1.5 MB + 24 launches + point count back); CPU restatement: %.2f ms/frame (1 thread)" % (dt * 1e3, dc * 1e3))
