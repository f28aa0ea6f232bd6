#!/bin/bash
# Scratch: the round's evidence set in one gpurun call (outputs under gpurun_out/).
mkdir -p gpurun_out/ev
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/ev/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/ev/bench_line.json 2> gpurun_out/ev/bench_err.txt
timeout 900 python bench.py --impl reference > gpurun_out/ev/bench_ref.json 2>> gpurun_out/ev/bench_err.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ev/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_kernel -s 1 -c 1 -f -o gpurun_out/ev/f_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ev/ncu_full.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 > gpurun_out/ev/memcheck_smoke.txt
timeout 200 python scripts/gpu_perf2.py > gpurun_out/ev/perf2.txt 2>&1
CVO_B200_LIB=build/variants/libcvo_b200_clk.so timeout 200 python scripts/gpu_phase_clocks.py > gpurun_out/ev/phase.txt 2>&1
tail -3 gpurun_out/ev/pytest_gpu.txt; cat gpurun_out/ev/bench_line.json gpurun_out/ev/bench_ref.json; tail -6 gpurun_out/ev/perf2.txt; cat gpurun_out/ev/memcheck_smoke.txt
