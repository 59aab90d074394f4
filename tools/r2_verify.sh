#!/bin/bash
# First GPU call of round 2, on main, after `bash tools/build_ab.sh` (builds one library per kernel variant):
#   gpurun --timeout 2700 -- 'bash tools/r2_verify.sh'
# 1. device parity of everything that was written without a GPU (main's build is the default library);
# 2. A/B numbers of the kernel variants (one library per branch in serf_b200/ab/), bench + per-tick profile.
# Everything lands in gpurun_out/r2_*.  Exit code = the test suite's.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1

timeout 2000 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_gpu_multi.py > gpurun_out/r2_tests.log 2>&1
rc=$?
tail -5 gpurun_out/r2_tests.log
if [ $rc -ne 0 ]; then     # localise: compaction off
  SERFSIM_COMPACT=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_tests_nocompact.log 2>&1
  tail -3 gpurun_out/r2_tests_nocompact.log
fi
for v in main ab-no-queue-word; do
  lib=$PWD/serf_b200/ab/libserfsim_$v.so
  [ -f "$lib" ] || continue
  SERFSIM_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$v.json 2>> gpurun_out/r2_bench.err
  SERFSIM_LIB=$lib timeout 300 python tools/tick_profile.py --out gpurun_out/r2_ticks_$v.json > gpurun_out/r2_ticks_$v.log 2>&1
  echo "$v: $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_$v.json'));print(d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])" 2>/dev/null)"
done
SERFSIM_COMPACT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_main_nocompact.json 2>> gpurun_out/r2_bench.err
SERFSIM_COMPACT=0 timeout 300 python tools/tick_profile.py --out gpurun_out/r2_ticks_main_nocompact.json > gpurun_out/r2_ticks_main_nocompact.log 2>&1
# issue-rate micro-benchmarks behind the LSU-bound reading of the plateau (DESIGN "What comes next" item 4)
[ -x tools/ubench/lsu_red ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/lsu_red tools/ubench/lsu_red.cu
timeout 120 tools/ubench/lsu_red > gpurun_out/r2_ubench.txt 2>&1; cat gpurun_out/r2_ubench.txt
# the new features at BASELINE scale (only if their parity passed)
if [ $rc -eq 0 ]; then
  timeout 600 python tools/feature_profile.py --what events --out gpurun_out/r2_events.json > gpurun_out/r2_events.log 2>&1; tail -1 gpurun_out/r2_events.log
  timeout 600 python tools/feature_profile.py --what byzantine --out gpurun_out/r2_byzantine.json > gpurun_out/r2_byzantine.log 2>&1; tail -1 gpurun_out/r2_byzantine.log
fi
exit $rc
