#!/bin/bash
# GPU call N of round 2 (one GPU): multi-slot tick kernel with the one-tile-ahead requests (SERFSIM_AHEAD) — device suite, the
# multi-slot parity files again with the path forced in every tick, A/B bench of the default workload, per-tick profile, full ncu
# captures of two multi-slot ticks and of the sharded single-slot plateau tick (loopback aid).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2n_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2n_tests.log
SERFSIM_AHEAD=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_fullsize.py tests/test_gpu_z_multiphase.py -m gpu -q -x > $O/r2n_tests_ahead2.log 2>&1
echo "tests ahead=2 rc=$?"; tail -3 $O/r2n_tests_ahead2.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2n_bench_lf_ahead1.json 2> $O/r2n_bench_lf_ahead1.err; echo "bench ahead=1 rc=$?"; summ $O/r2n_bench_lf_ahead1.json; tail -2 $O/r2n_bench_lf_ahead1.err
SERFSIM_AHEAD=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-check > $O/r2n_bench_lf_ahead0.json 2> $O/r2n_bench_lf_ahead0.err; echo "bench ahead=0 rc=$?"; summ $O/r2n_bench_lf_ahead0.json
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2n_ticks_storm_fail.json > $O/r2n_ticks_storm_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2n_ticks_storm_fail.json'));print('storm_fail', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]))"
for t in 20 30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2n_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2n_ncu_tick$t.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2n_loop8_tick13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2n_ncu_loop_tick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 20 --launch-count 1 -f -o $O/r2n_loop8f_tick20 \
    python tools/loopback_profile.py --world 8 --fail --runs 1 > $O/r2n_ncu_loopf_tick.log 2>&1
ls -la $O/r2n*.ncu-rep
