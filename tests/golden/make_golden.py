"""Generates tests/golden/*.json|npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference ships NO golden vectors for this path, and it cannot be built here, so
these fixtures are produced by the CPU oracle (oracle/cvo_oracle.cpp, brute-force variant) on
  (a) seeded synthetic pairs (cvo_rgbd_b200/synth.py), regenerated from the seed at test time, and
  (b) a REAL pair: two consecutive clouds of the reference's TUM fr1_desk sample
      (/root/reference/data/rgbd_dataset/freiburg1_desk/pcd_ds/1305031453.359684.pcd / .391690.pcd),
      sub-sampled to 1500 points with seed 0, colour unpacked from the packed-float rgb field, gradient
      features 0.  The sub-sampled arrays are stored in real_pair.npz because
      /root/reference does not exist on the GPU box.
They pin the oracle against accidental change and give the GPU tests a reference-data case.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cvo_rgbd_b200 import synth  # noqa: E402
from oracle import cvo_oracle as O  # noqa: E402

PCD_DIR = "/root/reference/data/rgbd_dataset/freiburg1_desk/pcd_ds"


def read_pcd_ascii(path):
    with open(path) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("DATA")) + 1
    rows = [l.split() for l in lines[start:] if l.strip()]
    xyz = np.array([[float(r[0]), float(r[1]), float(r[2])] for r in rows])
    rgbf = np.array([float(r[3]) for r in rows], dtype=np.float32)
    packed = rgbf.view(np.uint32)
    r, g, b = (packed >> 16) & 255, (packed >> 8) & 255, packed & 255
    return xyz, np.stack([b, g, r], axis=1).astype(np.float64)  # BGR like the reference's cv::Mat


def real_pair(n=1500):
    out = []
    for k, name in enumerate(["1305031453.359684.pcd", "1305031453.391690.pcd"]):
        xyz, bgr = read_pcd_ascii(os.path.join(PCD_DIR, name))
        ok = np.isfinite(xyz).all(1)
        xyz, bgr = xyz[ok], bgr[ok]
        idx = np.random.default_rng(k).choice(len(xyz), n, replace=False)
        feat = np.concatenate([bgr[idx], np.zeros((n, 2))], axis=1)
        out.append((xyz[idx].astype(np.float32), feat.astype(np.float32)))
    return out


def rec(e):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in e.items()}


def main():
    cases = {}
    R0 = synth._rotvec_to_R(np.array([0.01, -0.02, 0.015])).astype(np.float32)
    T0 = np.array([0.01, 0.005, -0.02], np.float32)
    for name, seed, n, m, kind in [("syn_cvo_500", 1000, 500, 500, "cvo"), ("syn_cvo_ragged", 77, 777, 1234, "cvo"),
                                   ("syn_acvo_ragged", 78, 900, 650, "acvo"), ("syn_acvo_m_gt_n", 79, 600, 1000, "acvo")]:
        pr = synth.make_pair(seed, n, m, kind)
        p = O.default_params(kind)
        ev = {str(ell): rec(O.evaluate(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], R0, T0, ell, p))
              for ell in (0.15, 0.1, 0.05)}
        al = O.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], p)
        cases[name] = dict(seed=seed, n=n, m=m, kind=kind, eval=ev,
                           align=dict(transform=al["transform"].tolist(), iters=al["iters"], status=al["status"],
                                      ell=al["ell"]))
    if os.path.isdir(PCD_DIR):
        (x, fx), (y, fy) = real_pair()
        np.savez_compressed(os.path.join(HERE, "real_pair.npz"), x_pos=x, x_feat=fx, y_pos=y, y_feat=fy)
    d = np.load(os.path.join(HERE, "real_pair.npz"))
    p = O.default_params("cvo")
    ev = {str(ell): rec(O.evaluate(d["x_pos"], d["x_feat"], d["y_pos"], d["y_feat"], np.eye(3), np.zeros(3), ell, p))
          for ell in (0.15, 0.1, 0.06, 0.03)}
    al = O.align(d["x_pos"], d["x_feat"], d["y_pos"], d["y_feat"], p)
    cases["real_fr1_desk_1500"] = dict(kind="cvo", eval=ev, align=dict(transform=al["transform"].tolist(),
                                       iters=al["iters"], status=al["status"], ell=al["ell"]))
    ip = O.inner_product(d["x_pos"], d["x_feat"], d["y_pos"], d["y_feat"], 0.1, p)
    cases["real_fr1_desk_1500"]["inner_product"] = ip
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote", list(cases))


if __name__ == "__main__":
    main()
