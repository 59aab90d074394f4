#!/bin/bash
# First GPU call of round 2, on branch r2-integration (= main + queue word + compaction), after `bash tools/build_ab.sh`:
#   gpurun --timeout 2700 -- 'bash tools/r2_verify.sh'
# 1. device parity of everything that was written without a GPU (the integration build is the default library);
# 2. A/B numbers of the kernel variants (one library per branch in serf_b200/ab/), bench + per-tick profile.
# Everything lands in gpurun_out/r2_*.  Exit code = the test suite's.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1

timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2_tests.log 2>&1
rc=$?
tail -5 gpurun_out/r2_tests.log
if [ $rc -ne 0 ]; then     # localise: the same suites against main's kernels (features only), then compaction off
  SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_main.so timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/r2_tests_mainlib.log 2>&1
  tail -3 gpurun_out/r2_tests_mainlib.log
  SERFSIM_COMPACT=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2_tests_nocompact.log 2>&1
  tail -3 gpurun_out/r2_tests_nocompact.log
fi
for v in main queue-word compaction integration; do
  lib=$PWD/serf_b200/ab/libserfsim_$v.so
  [ -f "$lib" ] || continue
  SERFSIM_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$v.json 2>> gpurun_out/r2_bench.err
  SERFSIM_LIB=$lib timeout 300 python tools/tick_profile.py --out gpurun_out/r2_ticks_$v.json > gpurun_out/r2_ticks_$v.log 2>&1
  echo "$v: $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_$v.json'));print(d['value'], d['kernel_ms_per_step'], d['roofline']['frac'])" 2>/dev/null)"
done
SERFSIM_COMPACT=0 SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_integration.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_integration_nocompact.json 2>> gpurun_out/r2_bench.err
exit $rc
