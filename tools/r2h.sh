#!/bin/bash
# GPU call H of round 2 (one GPU): node tick with the cold paths called on copies (record stays in registers), asynchronous block
# reservation in the sharded kernel; A/B of the register budgets; loopback profile (W = 8, 2); wire-codec device tests.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_z_fullsize.py > $O/r2h_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2h_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches']))"; }
run() { name=$1; wl=$2; shift 2; env "$@" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-check --workload $wl > $O/r2h_bench_${name}_$wl.json 2> $O/r2h_bench_${name}_$wl.err; echo "bench $name $wl rc=$?"; summ $O/r2h_bench_${name}_$wl.json; tail -2 $O/r2h_bench_${name}_$wl.err; }
run main leave
run main leave_fail
run inl leave SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_inl.so
run inl leave_fail SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_inl.so
run r1mb3 leave SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_r1mb3.so
run rn2nopf leave_fail SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_rn2nopf.so
for wl in storm storm_fail; do
  timeout 300 python tools/tick_profile.py --scenario $wl --out $O/r2h_ticks_$wl.json > $O/r2h_ticks_$wl.log 2>&1
  python -c "import json;d=json.load(open('$O/r2h_ticks_$wl.json'));print('$wl', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]))"
done
for w in 8 2; do
  SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py --world $w --out $O/r2h_loop_w$w.json > $O/r2h_loop_w$w.log 2>&1
  tail -1 $O/r2h_loop_w$w.log; grep -E "^rank" $O/r2h_loop_w$w.log | tail -5
  python -c "import json;d=json.load(open('$O/r2h_loop_w$w.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:40]))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2h_leave_tick13 \
    python tools/tick_profile.py --runs 1 > $O/r2h_ncu_tick13.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2h_loop8_tick13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2h_ncu_loop_tick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:drain_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2h_loop8_drain13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2h_ncu_loop_drain.log 2>&1
ls -la $O/r2h*.ncu-rep
