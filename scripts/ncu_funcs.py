"""Aggregates an `ncu -i rep --page source --csv --print-source cuda,sass` dump by (file, function): samples, warp
instructions, and the leading stall reasons.  usage: ncu_funcs2.py dump.csv [--lines N]"""
import csv, collections, re, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
heads_cache = {}
def heads_of(path):
    if path in heads_cache: return heads_cache[path]
    heads = []
    p = path if os.path.exists(path) else os.path.join(root, "cvo_rgbd_b200", "csrc", os.path.basename(path))
    if os.path.exists(p):
        for i, l in enumerate(open(p, errors="replace"), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__)[^;(]*?\b(\w+)\s*\(", l)
            if m and not l.startswith(" "): heads.append((i, m.group(1)))
    heads_cache[path] = heads
    return heads
def func_of(path, ln):
    name = "?"
    for h, n in heads_of(path):
        if h <= ln: name = n
        else: break
    return name
S = collections.Counter(); I = collections.Counter(); ST = collections.defaultdict(collections.Counter)
L = collections.Counter(); LI = collections.Counter()
cur_file = None; idx = None; cur_line = None
csv.field_size_limit(1 << 30)
for r in csv.reader(open(sys.argv[1], errors="replace")):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur_file = r[1]; continue
    if r[0] == "Line No":
        idx = {}
        for i, h in enumerate(r): idx.setdefault(h, i)
        continue
    if idx is None or cur_file is None: continue
    if not r[0].strip(): continue  # SASS rows (elided in the dump): the source-line rows carry the aggregates
    try: cur_line = int(r[0])
    except ValueError: continue
    if len(r) <= idx.get("# Samples", 10**9): continue
    def num(k):
        try: return int(float(r[idx[k]]))
        except Exception: return 0
    key = (os.path.basename(cur_file), func_of(cur_file, cur_line or 0))
    s, i = num("# Samples"), num("Instructions Executed")
    S[key] += s; I[key] += i
    L[(os.path.basename(cur_file), cur_line)] += s; LI[(os.path.basename(cur_file), cur_line)] += i
    for k in idx:
        if k.startswith("stall_") and "Not Issued" not in k: ST[key][k[6:]] += num(k)
ts, ti = sum(S.values()), sum(I.values())
print("file:function                              samples%%  inst%%  top stalls   (samples %d, warp insts %d)" % (ts, ti))
for k, s in S.most_common(28):
    top = ", ".join("%s %.0f%%" % (n, 100 * c / max(1, sum(ST[k].values()))) for n, c in ST[k].most_common(4))
    print("%-42s %7.1f %6.1f  %s" % (k[0][:14] + ":" + k[1], 100 * s / ts, 100 * I[k] / ti, top))
if "--lines" in sys.argv:
    n = int(sys.argv[sys.argv.index("--lines") + 1])
    print("hot lines")
    for k, s in L.most_common(n): print("  %-18s %5d  samples %5.2f%%  inst %5.2f%%" % (k[0], k[1] or 0, 100 * s / ts, 100 * LI[k] / ti))
