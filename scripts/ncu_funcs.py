"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump by FUNCTION of cvo_kernels.cuh
(line ranges found by scanning the source for function heads). usage: ncu_funcs.py dump.csv [source]"""
import csv, collections, re, sys, os
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cvo_rgbd_b200", "csrc", "cvo_kernels.cuh")
heads = []
for i, l in enumerate(open(src), 1):
    m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", l)
    if m and not l.startswith(" "):
        heads.append((i, m.group(1)))
def func_of(ln):
    name = "?"
    for h, n in heads:
        if h <= ln: name = n
        else: break
    return name
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No'][0]
hdr = rows[hi]
idx = {}
for i, h in enumerate(hdr): idx.setdefault(h, i)
def num(v):
    try: return int(float(v))
    except: return 0
S = collections.Counter(); I = collections.Counter()
for r in rows[hi + 1:]:
    try: ln = int(r[0])
    except: continue
    f = func_of(ln)
    S[f] += num(r[idx['# Samples']]); I[f] += num(r[idx['Instructions Executed']])
ts, ti = sum(S.values()), sum(I.values())
print("function                 samples%%   inst%%   (total samples %d, warp insts %d)" % (ts, ti))
for f, s in S.most_common():
    print("%-24s %7.1f %7.1f" % (f, 100 * s / ts, 100 * I[f] / ti))
