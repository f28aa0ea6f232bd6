import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import synth
from scripts.tile_stats import morton_order
def stats(x, y, ell, rt=32, ct=32):
    r2 = -2*ell*ell*np.log(0.008/0.01)
    x = x[morton_order(x)]; y = y[morton_order(y)]
    nxt, nyt = -(-len(x)//rt), -(-len(y)//ct)
    live=0; act_rows=0; act_cols=0; inball=0; both=0
    for a in range(nxt):
        xa = x[a*rt:(a+1)*rt]; alo, ahi = xa.min(0), xa.max(0)
        for b in range(nyt):
            yb = y[b*ct:(b+1)*ct]; blo, bhi = yb.min(0), yb.max(0)
            gap = np.maximum(0, np.maximum(alo-bhi, blo-ahi))
            if (gap**2).sum() <= r2:
                live += 1
                gr = np.maximum(0, np.maximum(blo - xa, xa - bhi)); ar = ((gr**2).sum(1) <= r2)
                gc = np.maximum(0, np.maximum(alo - yb, yb - ahi)); ac = ((gc**2).sum(1) <= r2)
                act_rows += ar.sum(); act_cols += ac.sum(); both += ar.sum()*ac.sum()
                d2 = ((xa[:,None,:]-yb[None,:,:])**2).sum(-1); inball += (d2<r2).sum()
    return dict(live=live, act_rows_per_pair=act_rows/live, act_cols_per_pair=act_cols/live, inball_per_pair=inball/live, cand_both=both/live)
pr = synth.config_pair(2)
for ell in (0.15, 0.1, 0.06, 0.03):
    print(ell, stats(pr['x_pos'], pr['y_pos'], ell))
