"""Tuning aid: finds the packed-arithmetic hot loops of align_kernel in the SASS of the built library and prints, per loop,
the opcode histogram and an estimate of the hot-path instructions per trip (the cold re-decision blocks excluded).
usage: sass_loops.py [lib.so] [--dump N]  (N: print the N-th loop's hot path)"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(root, "cvo_rgbd_b200", "libcvo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
ins = []
for l in txt:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
packed = [i for i, (a, t) in enumerate(ins) if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", t)]
groups, cur = [], [packed[0]]
for i in packed[1:]:
    if i - cur[-1] > 150: groups.append(cur); cur = [i]
    else: cur.append(i)
groups.append(cur)
for gi, g in enumerate(groups):
    lo, hi = g[0], g[-1]
    # extend to the loop: back to the first LDS burst before lo, forward to the backward branch
    while lo > 0 and g[0] - lo < 40 and not re.search(r"\bBRA\b", ins[lo - 1][1]): lo -= 1
    while hi < len(ins) - 1 and hi - g[-1] < 60 and not re.match(r"(@\S+\s+)?BRA\b", ins[hi][1]): hi += 1
    body = [t for a, t in ins[lo:hi + 1]]
    ntrip = max(1, sum(1 for t in body if "VOTE.ANY" in t))
    # cold blocks: from a "@P0 BRA" following FSETP.GEU |x| to its BSYNC
    hot, cold, skip = [], 0, False
    for t in body:
        if "BSSY" in t: skip = True
        if skip: cold += 1
        else: hot.append(t)
        if "BSYNC" in t: skip = False
    ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in hot)
    print("loop %d: %d instr, %d trips unrolled, cold %d -> hot path %.0f per trip" % (gi, len(body), ntrip, cold, (len(hot) - 4 * ntrip) / ntrip))
    print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    if "--dump" in sys.argv and int(sys.argv[sys.argv.index("--dump") + 1]) == gi:
        for t in hot: print("      " + t)
