import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import synth
from scripts.tile_stats import morton_order
def stats(x, y, ell, rt=32, ct=32):
    r2 = -2*ell*ell*np.log(0.008/0.01)
    x = x[morton_order(x)]; y = y[morton_order(y)]
    nxt, nyt = -(-len(x)//rt), -(-len(y)//ct)
    xlo = np.array([x[a*rt:(a+1)*rt].min(0) for a in range(nxt)]); xhi = np.array([x[a*rt:(a+1)*rt].max(0) for a in range(nxt)])
    ylo = np.array([y[b*ct:(b+1)*ct].min(0) for b in range(nyt)]); yhi = np.array([y[b*ct:(b+1)*ct].max(0) for b in range(nyt)])
    gap = np.maximum(0, np.maximum(xlo[:,None,:]-yhi[None,:,:], ylo[None,:,:]-xhi[:,None,:]))
    live = ((gap**2).sum(-1) <= r2)
    return dict(rt=rt, ct=ct, live=int(live.sum()), warp_steps=int(live.sum())*ct*(rt//32 if rt>=32 else 1))
pr = synth.config_pair(2)
for ell in (0.1, 0.03):
    for rt, ct in ((32,32),(32,16),(32,8),(32,4),(64,32)):
        print(ell, stats(pr['x_pos'], pr['y_pos'], ell, rt, ct))
