"""Scratch: dump GPU and oracle align traces for offline comparison."""
import sys, os, pickle
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cvo_rgbd_b200 import capi, synth
from oracle import cvo_oracle as O
ctx = capi.Context(0, max_points=4096, max_slots=2)
out = {}
for cfg, kind in ((1, 'cvo'), (2, 'cvo'), (3, 'acvo')):
    pr = synth.config_pair(cfg)
    ctx.set_pair(0, pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'])
    gp = capi.default_params(kind); op = O.default_params(kind)
    for G in (1, 8):
        ctx.set_cluster_size(G)
        out[(cfg, 'gpu', G)] = ctx.align_trace(0, gp)
    out[(cfg, 'port')] = O.align(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], op, trace_cap=2048)
    try:
        out[(cfg, 'ref')] = O.align(pr['x_pos'], pr['x_feat'], pr['y_pos'], pr['y_feat'], op, trace_cap=2048, variant='ref')
    except Exception as e:
        print('ref unavailable', e)
pickle.dump(out, open('gpurun_out/traces.pkl', 'wb'))
print('ok')
