#!/bin/bash
# GPU call F of round 2 (one GPU): sleeping views (timer wheel, per-view visiting, idle-tick skipping / host jump) on the device:
# the whole single-GPU suite incl. the full-size configs[2] / configs[4] cases, both bench workloads, per-tick profiles, A/B switches.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2f_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/r2f_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  eu %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['edge_updates_per_step'], d['gpu_launches'])); print(d['self_check'])"; }
run() { name=$1; shift; env "$@" timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2f_bench_$name.json 2> $O/r2f_bench_$name.err; echo "bench $name rc=$?"; summ $O/r2f_bench_$name.json; tail -2 $O/r2f_bench_$name.err; }
run leave_fail
run leave_fail_nojump SERFSIM_NO_JUMP=1
run leave_fail_noskip SERFSIM_NO_SKIP=1
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline > $O/r2f_bench_leave.json 2> $O/r2f_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2f_bench_leave.json
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2f_ticks_leave_fail.json > $O/r2f_ticks_leave_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2f_ticks_leave_fail.json'));print(d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:200]))"
timeout 300 python tools/tick_profile.py --out $O/r2f_ticks_leave.json > $O/r2f_ticks_leave.log 2>&1
python -c "import json;d=json.load(open('$O/r2f_ticks_leave.json'));print(d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))"
for t in 20 30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2f_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2f_ncu_tick$t.log 2>&1
done
