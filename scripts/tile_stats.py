"""Scratch: emulate Morton sort + tile-box culling on CPU to count live tile pairs and survivor loop trips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import synth

def spread10(v):
    v = v & 0x3ff
    v = (v | (v << 16)) & 0x030000ff
    v = (v | (v << 8)) & 0x0300f00f
    v = (v | (v << 4)) & 0x030c30c3
    v = (v | (v << 2)) & 0x09249249
    return v
def morton_order(p):
    lo, hi = p.min(0), p.max(0)
    q = ((p - lo) * (1023.999 / (hi - lo))).astype(np.int64)
    code = spread10(q[:,0]) | (spread10(q[:,1]) << 1) | (spread10(q[:,2]) << 2)
    return np.lexsort((np.arange(len(p)), code))
def stats(x, y, ell, tile=32, rows_per_lane=1):
    r2 = -2*ell*ell*np.log(0.008/0.01)
    x = x[morton_order(x)]; y = y[morton_order(y)]
    rt = tile*rows_per_lane
    nxt, nyt = -(-len(x)//rt), -(-len(y)//tile)
    live = 0; trips = 0; inball = 0; trips_mean=0
    for a in range(nxt):
        xa = x[a*rt:(a+1)*rt]; alo, ahi = xa.min(0), xa.max(0)
        for b in range(nyt):
            yb = y[b*tile:(b+1)*tile]; blo, bhi = yb.min(0), yb.max(0)
            gap = np.maximum(0, np.maximum(alo-bhi, blo-ahi))
            if (gap**2).sum() <= r2:
                live += 1
                d2 = ((xa[:,None,:]-yb[None,:,:])**2).sum(-1)
                m = (d2 < r2).sum(1)
                # rows_per_lane rows per lane: trip count = max over lanes of sum of popcounts of its rows
                pad = np.zeros(rt, int); pad[:len(m)] = m
                per_lane = pad.reshape(rows_per_lane, tile).sum(0)
                trips += per_lane.max(); inball += m.sum(); trips_mean += per_lane.mean()
    return dict(tilepairs=nxt*nyt, live=live, live_frac=live/(nxt*nyt), inball=int(inball), trips=int(trips), util=trips_mean/max(trips,1))
if __name__ == "__main__":
  pr = synth.config_pair(2)
  for ell in (0.15, 0.1, 0.06, 0.03):
      print('ell', ell, stats(pr['x_pos'], pr['y_pos'], ell))
  print('2 rows/lane', stats(pr['x_pos'], pr['y_pos'], 0.1, rows_per_lane=2))
  print('tile16', stats(pr['x_pos'], pr['y_pos'], 0.1, tile=16))
