#!/bin/bash
# GPU call O of round 2 (one GPU): full ncu captures (at most three .ncu-rep per call: gpurun_out is capped at 64 MiB) of two multi-slot
# ticks of the default workload and of the sharded single-slot plateau tick (loopback aid).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2o_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2o_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2o_bench_lf.json 2> $O/r2o_bench_lf.err; echo "bench rc=$?"; summ $O/r2o_bench_lf.json; tail -2 $O/r2o_bench_lf.err
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2o_ticks_storm_fail.json > $O/r2o_ticks_storm_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2o_ticks_storm_fail.json'));print('storm_fail', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][145:175]))"
for t in 20 30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2o_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2o_ncu_tick$t.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2o_loop8_tick13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2o_ncu_loop_tick.log 2>&1
ls -la $O/r2o*.ncu-rep
