"""Generates tests/golden/eval_golden.json: outputs of the REFERENCE's own (vendored TUM) evaluation functions on
seeded synthetic trajectories, used to pin cvo_rgbd_b200/evaluate.py.  Run here (needs /root/reference):
    python tests/golden/make_eval_golden.py
The reference tools are Python-2 scripts (print statements in their __main__ blocks), so only their function
definitions are compiled, straight from the files where they lie -- nothing is copied into this repository."""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
TOOLS = "/root/reference/data/rgbd_dataset/rgbd_benchmark_tools"


def load_functions(fname, extra=None):
    """exec()s only the FunctionDef nodes of a reference tool into a fresh module namespace."""
    src = open(os.path.join(TOOLS, fname)).read()
    # keep the text of every top-level `def` block (the Python-2-only statements live in `if __name__ == ...`)
    lines, keep, out = src.split("\n"), False, []
    for ln in lines:
        if ln.startswith("def "):
            keep = True
        elif ln and not ln[0].isspace() and not ln.startswith("def "):
            keep = False
        if keep:
            out.append(ln)
    mod = types.ModuleType(fname[:-3])
    mod.__dict__.update(dict(numpy=np, random=__import__("random")))
    if not hasattr(np.linalg, "linalg"):  # numpy >= 2 removed the alias the 2012 script uses
        np.linalg.linalg = np.linalg
    if extra:
        mod.__dict__.update(extra)
    exec(compile("\n".join(out), fname, "exec"), mod.__dict__)
    return mod


def synthetic_trajectories(seed, n=120):
    from cvo_rgbd_b200.synth import _rotvec_to_R
    rng = np.random.default_rng(seed)
    stamps = 1000.0 + np.cumsum(rng.uniform(0.025, 0.04, n))
    T = np.eye(4)
    gt, est = {}, {}
    D = np.eye(4)
    D[:3, :3] = _rotvec_to_R(np.array([0.3, -0.2, 0.5]))
    D[:3, 3] = [1.0, -2.0, 0.5]
    E = np.eye(4)
    for k, s in enumerate(stamps):
        d = np.eye(4)
        d[:3, :3] = _rotvec_to_R(rng.normal(0, 0.012, 3))
        d[:3, 3] = rng.normal(0, 0.01, 3)
        T = T @ d
        gt[float(round(s, 4))] = D @ T
        n_ = np.eye(4)
        n_[:3, :3] = _rotvec_to_R(rng.normal(0, 0.002, 3))
        n_[:3, 3] = rng.normal(0, 0.002, 3)
        E = E @ d @ n_
        est[float(round(s + rng.uniform(-0.004, 0.004), 6))] = E
    return gt, est


def main():
    assoc = load_functions("associate.py")
    ate = load_functions("evaluate_ate.py")
    rpe = load_functions("evaluate_rpe.py")
    out = {}
    for seed in (1, 2, 3):
        gt, est = synthetic_trajectories(seed)
        # ATE exactly as evaluate_ate.py's __main__ does it (:121-146)
        matches = assoc.associate({k: 0 for k in gt}, {k: 0 for k in est}, 0.0, 0.02)
        first = np.matrix([gt[a][:3, 3] for a, b in matches]).transpose()
        second = np.matrix([est[b][:3, 3] for a, b in matches]).transpose()
        rot, trans, err = ate.align(second, first)
        # RPE with --fixed_delta --delta 1 --delta_unit s / f (evaluate_rpe.py:204-297)
        res = {}
        for unit, delta in (("s", 1.0), ("f", 5.0)):
            r = np.array(rpe.evaluate_trajectory(gt, est, 0, True, delta, unit, 0.0, 1.0))
            res[unit] = dict(pairs=int(len(r)), trans_rmse=float(np.sqrt(np.dot(r[:, 4], r[:, 4]) / len(r))),
                             rot_rmse=float(np.sqrt(np.dot(r[:, 5], r[:, 5]) / len(r))),
                             trans_mean=float(np.mean(r[:, 4])), rot_mean=float(np.mean(r[:, 5])))
        out[str(seed)] = dict(ate=dict(pairs=int(len(err)), rmse=float(np.sqrt(np.dot(err, err) / len(err))),
                                       mean=float(np.mean(err)), median=float(np.median(err)), max=float(np.max(err))),
                              n_matches=len(matches), rpe=res)
    json.dump(out, open(os.path.join(HERE, "eval_golden.json"), "w"), indent=1)
    print(json.dumps(out, indent=1)[:600])


if __name__ == "__main__":
    main()
