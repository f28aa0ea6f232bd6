"""Generates tests/golden/pcd_golden.json: fingerprints of the image front end's outputs (oracle/pcd_oracle.cpp, with
the colour conversions pinned against cv2 by tests/test_pcd_oracle.py) on seeded synthetic frames.
Run here: python tests/golden/make_pcd_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_rgbd_b200 import synth  # noqa: E402
from oracle import pcd_oracle as P  # noqa: E402

CASES = [dict(seed=1, texture=1.0, dataset_seq=1, feature_type=1), dict(seed=2, texture=1.0, dataset_seq=1, feature_type=0),
         dict(seed=3, texture=0.3, dataset_seq=2, feature_type=1), dict(seed=4, texture=0.05, dataset_seq=3, feature_type=0),
         dict(seed=5, texture=3.0, dataset_seq=0, feature_type=1)]


def fingerprint(case):
    img, dep = synth.make_frame(case["seed"], texture=case["texture"])
    r = P.create_pointcloud(img, dep, case["dataset_seq"], case["feature_type"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    return dict(case, num_selected=int(r["num_selected"]), n=int(len(r["xyz"])), pots=r["pots"], canny=bool(r["canny"]),
                levels=[int((r["map"] == v).sum()) for v in (1, 2, 4)], map_sha=sha(r["map"].astype(np.uint8)),
                xyz_sha=sha(r["xyz"]), feat_sha=sha(r["feat"]), img_sha=sha(img), depth_sha=sha(dep),
                xyz_head=r["xyz"][:3].tolist(), feat_head=r["feat"][:3].tolist())


if __name__ == "__main__":
    out = [fingerprint(c) for c in CASES]
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pcd_golden.json"), "w"), indent=1)
    for o in out:
        print(o["seed"], o["num_selected"], o["n"], o["pots"], o["levels"])
