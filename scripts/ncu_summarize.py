"""Turns the ncu outputs a gpurun call brought back into the small summaries committed under profiles/.
usage: ncu_summarize.py <launches.csv> <full.ncu-rep> <tag> [pairs per launch]   (run here; needs `ncu` on PATH, no GPU;
run it at the commit the capture was taken at: the source fingerprint bench.py compares is computed from the tree)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 592
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_fingerprint, algorithmic bytes)

# 1. launch list -> per-kernel totals and shares
rows = [r for r in csv.reader(open(launches_csv, errors="replace")) if r]
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    ms = v / 1e6 if r[mu] in ("ns", "nsecond") else (v / 1e3 if r[mu] in ("us", "usecond") else v)
    name = r[kn].split("(")[0][:80]
    tot[name][0] += 1
    tot[name][1] += ms
allms = sum(v[1] for v in tot.values())
summary = {k: {"launches": v[0], "total_ms": v[1], "share": v[1] / allms} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}
json.dump(summary, open(os.path.join(ROOT, "profiles", "%s_launch_summary.json" % tag), "w"), indent=1)
print(json.dumps(summary, indent=1))

# 2. full capture -> DRAM traffic + headline pipe metrics
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, vals = r[0], r[2]
m = dict(zip(h, vals))
def f(k):
    try:
        return float(m[k].replace(",", ""))
    except Exception:
        return None
keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
out = {k: f(k) for k in keys}
units = dict(zip(h, r[1]))
rd, wr = out["dram__bytes_read.sum"], out["dram__bytes_write.sum"]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rd *= scale.get(units["dram__bytes_read.sum"], 1.0)
wr *= scale.get(units["dram__bytes_write.sum"], 1.0)
alg = pairs * 100 * bench.algorithmic_bytes_per_iteration(3000, 3000)
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
traffic = {"kernel": "cvo_b200::align_kernel", "source": "ncu --set full --clock-control none -k regex:align_kernel -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cfg4 (%s)" % os.path.basename(rep),
           "workload": "%d cfg-2 pairs (3000x3000, fixed ell 0.10, 100 iterations) in one launch" % pairs,
           "pairs_per_launch": pairs, "source_fingerprint": bench.source_fingerprint(), "git_commit": commit,
           "kernel_ms": out["gpu__time_duration.sum"],
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": alg, "traffic_over_algorithmic": (rd + wr) / alg, "metrics": out,
           "note": "algorithmic bytes by SURVEY 8d (64(N+M)+96 per iteration); the measured DRAM traffic is dominated by the neighbour candidate lists (quads: 6.5 B per candidate slot, about 0.61 MB per pair, read by both passes of every iteration; 148 of them plus the clouds exceed the 126 MB L2, so part of the stream comes from HBM)"}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "align_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))

# 3. per-source-line stall summary
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp = "/tmp/_src_%s.csv" % tag
open(tmp, "w").write(src)
txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_funcs.py"), tmp, "--lines", "40"], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", "%s_align_kernel_by_function.txt" % tag), "w").write(txt)
print(txt[:2500])
