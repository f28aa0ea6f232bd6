"""Generates tests/golden/refsrc_golden.json: outputs of the REFERENCE'S OWN first-party sources (src/cvo.cpp,
src/adaptive_cvo.cpp, src/LieGroup.cpp, thirdparty/nanoflann.hpp -- compiled unmodified from /root/reference against
oracle/shim, see oracle/refsrc_driver.cpp) on seeded inputs.  These are the vectors that PIN the oracle restatement
(tests/test_refsrc_golden.py, CPU) and that the CUDA path is held to directly (tests/test_gpu_refsrc.py).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_refsrc_golden.py
Serial execution (CVO_SHIM_THREADS unset): the outputs are bit-reproducible."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cvo_rgbd_b200 import synth  # noqa: E402
from oracle import cvo_oracle as O  # noqa: E402
from oracle import refsrc as Rs  # noqa: E402

GOLD = os.path.dirname(os.path.abspath(__file__))
R0 = synth._rotvec_to_R(np.array([0.01, -0.02, 0.015])).astype(np.float32)
T0 = np.array([0.01, 0.005, -0.02], np.float32)


def lst(a):
    return np.asarray(a, np.float64).tolist()


def clouds(pr):
    return pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"]


def eval_case(kind, pr, R, T, ells):
    out = {}
    for ell in ells:
        e = Rs.evaluate(kind, *clouds(pr), R, T, ell)
        out[repr(ell)] = dict(nnz=e["nnz"], nnz_xx=e["nnz_xx"], nnz_yy=e["nnz_yy"], sum_a=e["sum_a"], omega=lst(e["omega"]),
                              v=lst(e["v"]), step=e["step"], dl=e["dl"])
    return out


def align_case(kind, pr, trace=4):
    a = Rs.align(kind, *clouds(pr), mode=Rs.MODE_REFERENCE_ALIGN)
    t = Rs.align(kind, *clouds(pr), mode=Rs.MODE_DRIVEN_TRACE, trace_cap=4096)
    assert np.array_equal(a["transform"], t["transform"]) and np.array_equal(a["prev_transform"], t["prev_transform"])
    assert a["iters"] == t["iters"]
    recs = [dict(ell=r["ell"], step=r["step"], omega=lst(r["omega"]), v=lst(r["v"]), nnz=r["nnz"], nnz_xx=r["nnz_xx"],
                 nnz_yy=r["nnz_yy"], sum_a=r["sum_a"], dl=r["dl"], R=lst(r["R"]), T=lst(r["T"])) for r in t["trace"][:trace]]
    return dict(transform=lst(a["transform"]), prev_transform=lst(a["prev_transform"]), iters=a["iters"], status=t["status"],
                R=lst(a["R"]), T=lst(a["T"]), ell=a["ell"], n_iterations_run=t["n_iterations_run"], first_iterations=recs)


def main():
    gold = {"generator": "tests/golden/make_refsrc_golden.py", "backend": Rs.backend(), "R0": lst(R0), "T0": lst(T0),
            "level1": {}, "align": {}, "fixed": {}, "sequence": {}, "inner_product": {}, "exp_sek3": [], "step": []}
    # ---- level 1: one evaluation at (R0, T0)
    for name, kind, seed, n, m in (("cvo_500", "cvo", 1000, 500, 500), ("cvo_777x1234", "cvo", 51, 777, 1234),
                                   ("cvo_33x2100", "cvo", 52, 33, 2100), ("acvo_900x650", "acvo", 53, 900, 650),
                                   ("acvo_600x1000", "acvo", 54, 600, 1000), ("cvo_3000", "cvo", 2000, 3000, 3000),
                                   ("acvo_3000", "acvo", 3000, 3000, 3000), ("acvo_2049x2050", "acvo", 57, 2049, 2050)):
        pr = synth.make_pair(seed, n, m, kind)
        gold["level1"][name] = dict(kind=kind, seed=seed, n=n, m=m, eval=eval_case(kind, pr, R0, T0, (0.15, 0.1, 0.05)))
        print("level1", name)
    real = dict(np.load(os.path.join(GOLD, "real_pair.npz")))
    gold["level1"]["real_pair"] = dict(kind="cvo", real=True, eval=eval_case("cvo", real, np.eye(3, dtype=np.float32),
                                                                               np.zeros(3, np.float32), (0.15, 0.1)))
    # ---- level 3: the reference's align() to convergence
    for name, kind, cfg, idx in (("cfg1", "cvo", 1, 0), ("cfg2_stock", "cvo", 2, 0), ("cfg3", "acvo", 3, 0), ("cfg4_0", "cvo", 4, 0),
                                 ("cfg4_1", "cvo", 4, 1), ("cfg3_1", "acvo", 3, 1)):
        gold["align"][name] = dict(kind=kind, cfg=cfg, pair_index=idx, **align_case(kind, synth.config_pair(cfg, idx)))
        print("align", name, gold["align"][name]["iters"])
    gold["align"]["real_pair"] = dict(kind="cvo", real=True, **align_case("cvo", real))
    # ---- benchmark config 2: fixed ell 0.10, exactly 100 iterations (driven loop over the reference's private functions)
    for idx in (0, 1, 2):
        pr = synth.config_pair(2, idx)
        f = Rs.align("cvo", *clouds(pr), mode=Rs.MODE_FIXED, fixed_iters=100, ell=0.10, trace_cap=100)
        gold["fixed"]["cfg2_%d" % idx] = dict(pair_index=idx, transform=lst(f["transform"]), R=lst(f["R"]), T=lst(f["T"]),
                                              first=dict(nnz=f["trace"][0]["nnz"], omega=lst(f["trace"][0]["omega"]),
                                                         v=lst(f["trace"][0]["v"]), step=f["trace"][0]["step"]),
                                              last=dict(nnz=f["trace"][-1]["nnz"], omega=lst(f["trace"][-1]["omega"]),
                                                        v=lst(f["trace"][-1]["v"]), step=f["trace"][-1]["step"]))
        print("fixed", idx)
    # ---- the reference's driver loop on one object (run_cvo per frame): quirks Q3, Q4, Q5
    for kind in ("cvo", "acvo"):
        first = synth.make_pair(900, 1200, 1200, kind)
        frames = [(first["x_pos"], first["x_feat"])]
        for k in range(1, 4):
            pr = synth.make_pair(900, 1200, 1200, kind, motion_scale=0.6 * k)
            frames.append((pr["y_pos"], pr["y_feat"]))
        s = Rs.run_sequence(kind, frames)
        gold["sequence"][kind] = dict(seed=900, n=1200, n_frames=4, transform=lst(s["transform"]), accum_transform=lst(s["accum_transform"]),
                                      iter=[int(i) for i in s["iter"]], ell=lst(s["ell"]))
        print("sequence", kind, s["iter"])
    # ---- acvo::function_inner_product
    pr = synth.make_pair(950, 1500, 1400, "acvo")
    gold["inner_product"] = dict(seed=950, n=1500, m=1400, values={repr(ell): Rs.inner_product(*clouds(pr), ell) for ell in (0.15, 0.1, 0.05)})
    # ---- Exp_SEK3 incl. the small-angle branch (quirk Q2)
    rng = np.random.default_rng(7)
    for i in range(12):
        w = rng.normal(0, 0.05, 3).astype(np.float32)
        if i >= 9:
            w *= np.float32(1e-8)
        v = rng.normal(0, 0.05, 3).astype(np.float32)
        dt = float(np.float32(rng.uniform(0.05, 0.8)))
        dR, dT = Rs.exp_sek3(w, v, dt)
        gold["exp_sek3"].append(dict(omega=lst(w), v=lst(v), dt=dt, dR=lst(dR), dT=lst(dT)))
    # ---- poly_solver + root selection on real coefficient sets (B..E from the restatement: they are locals of the
    #      reference's compute_step_size) and on synthetic ones (three real roots, E = 0 -> NaN -> min_step, clamps)
    sets = []
    for kind, seed, n, m in (("cvo", 1000, 500, 500), ("cvo", 2000, 3000, 3000), ("acvo", 53, 900, 650)):
        pr = synth.make_pair(seed, n, m, kind)
        for ell in (0.15, 0.1, 0.05):
            e = O.evaluate(*clouds(pr), R0, T0, ell, O.default_params(kind))
            sets.append((e["B"], e["C"], e["D"], e["E"]))
    sets += [(-6.0, 11.0 / 2, -6.0 / 3, 1.0 / 4), (1.0, 1.0, 1.0, 0.0), (-1e-3, 1.0, 0.0, 0.25), (1.0, 2.0, 3.0, 4.0),
             (-50.0, 1.0, 0.0, 0.25)]
    for B, Cc, D, E in sets:
        gold["step"].append(dict(B=B, C=Cc, D=D, E=E, step=Rs.step_from_coeffs(B, Cc, D, E)))
    json.dump(gold, open(os.path.join(GOLD, "refsrc_golden.json"), "w"), indent=0)
    print("wrote refsrc_golden.json")


if __name__ == "__main__":
    main()
