#!/bin/bash
# The round's evidence set in one gpurun call (outputs under gpurun_out/ev/; summarised into profiles/ afterwards by
# scripts/ncu_summarize.py and by hand).  Needs the clk variant: python scripts/build_variants.py clk:CVO_PHASE_CLOCKS
mkdir -p gpurun_out/ev
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/ev/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/ev/bench_line.json 2> gpurun_out/ev/bench_err.txt
timeout 900 python bench.py --impl reference > gpurun_out/ev/bench_ref.json 2>> gpurun_out/ev/bench_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cfg4 > gpurun_out/ev/ncu_bench.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 1 -c 1 -f -o gpurun_out/ev/f_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cfg4 > gpurun_out/ev/ncu_full.log 2>&1
# BASELINE config 5 (the "HBM roofline report"): 10 000 x 10 000 points, fixed ell, whole-GPU mode; and the stock schedules
# (gpurun copies back at most 64 MiB: the four reports together exceed it -- SKIP_OTHER_NCU=1 leaves these three to a second call)
if [ -z "$SKIP_OTHER_NCU" ]; then
timeout 600 ncu --set full --clock-control none -k regex:align_kernel -s 2 -c 1 -f -o gpurun_out/ev/cfg5_prof python scripts/gpu_other_configs.py cfg5 > gpurun_out/ev/ncu_cfg5.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:align_kernel -s 2 -c 1 -f -o gpurun_out/ev/stock_cvo_prof python scripts/gpu_other_configs.py cvo > gpurun_out/ev/ncu_stock_cvo.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:align_kernel -s 2 -c 1 -f -o gpurun_out/ev/stock_acvo_prof python scripts/gpu_other_configs.py acvo > gpurun_out/ev/ncu_stock_acvo.log 2>&1
fi
timeout 300 python scripts/gpu_other_configs.py all > gpurun_out/ev/other_configs.txt 2>&1
for m in cfg2 stock single; do echo "== $m"; CVO_B200_LIB=build/variants/libcvo_b200_clk.so timeout 200 python scripts/gpu_phase_clocks.py $m; done > gpurun_out/ev/phase.txt 2>&1
timeout 300 python scripts/gpu_sequence_perf.py > gpurun_out/ev/sequence.txt 2>&1
timeout 300 python scripts/gpu_pcd_perf.py > gpurun_out/ev/pcd.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev/smoke.txt 2>&1
tail -3 gpurun_out/ev/pytest_gpu.txt; cat gpurun_out/ev/other_configs.txt; tail -c 600 gpurun_out/ev/bench_ref.json
