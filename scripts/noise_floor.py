"""Measures the reference algorithm's own numerical noise floor: the same restatement built two ways
(explicit no-contraction + brute force ball  vs  GCC fp-contract=fast + reference nanoflann) on the same inputs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import synth
from oracle import cvo_oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import pose_diff
rows = []
for kind, n in (("cvo", 3000), ("acvo", 3000), ("cvo", 1200)):
    for seed in range(5000, 5008):
        pr = synth.make_pair(seed, n, n, kind)
        p = O.default_params(kind)
        a = O.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], p, variant="port")
        b = O.align(pr["x_pos"], pr["x_feat"], pr["y_pos"], pr["y_feat"], p, variant="ref")
        rot, tr = pose_diff(a["transform"], b["transform"])
        rows.append(dict(kind=kind, n=n, seed=seed, rot=rot, trans=tr, iters_port=a["iters"], iters_ref=b["iters"]))
        print(rows[-1], flush=True)
r = np.array([[x["rot"], x["trans"]] for x in rows])
print("max rot %.2e max trans %.2e median rot %.2e median trans %.2e" % (r[:,0].max(), r[:,1].max(), np.median(r[:,0]), np.median(r[:,1])))
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "oracle_noise_floor_r01.json"), "w"), indent=1)
