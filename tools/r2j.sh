#!/bin/bash
# GPU call J of round 2 (one GPU, the last one): the device suite on the shipped binary, both bench workloads with the per-step
# oracle check, the loopback profile of the sharded path, and the ncu evidence of the shipped binary for the default workload
# (launch list of the bench command, DRAM bytes of every tick launch → profiles/r2_traffic_leave_fail.json, one full capture).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2j_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2j_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d['self_check']); print(d.get('cpu_baseline'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2j_bench_leave_fail.json 2> $O/r2j_bench_leave_fail.err; echo "bench leave_fail rc=$?"; summ $O/r2j_bench_leave_fail.json; tail -2 $O/r2j_bench_leave_fail.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline > $O/r2j_bench_leave.json 2> $O/r2j_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2j_bench_leave.json
for wl in storm_fail storm; do
  timeout 300 python tools/tick_profile.py --scenario $wl --out $O/r2j_ticks_$wl.json > $O/r2j_ticks_$wl.log 2>&1
  python -c "import json;d=json.load(open('$O/r2j_ticks_$wl.json'));print('$wl', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]))"
done
for w in 8 2; do
  SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py --world $w --out $O/r2j_loop_w$w.json > $O/r2j_loop_w$w.log 2>&1
  tail -1 $O/r2j_loop_w$w.log; grep -E "^rank" $O/r2j_loop_w$w.log | tail -5
done
SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py --world 8 --fail --out $O/r2j_loop_w8_fail.json > $O/r2j_loop_w8_fail.log 2>&1; tail -1 $O/r2j_loop_w8_fail.log; grep -E "^rank 0:" $O/r2j_loop_w8_fail.log | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2j_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > $O/r2j_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tick_kernel --clock-control none --csv \
    --log-file $O/r2j_traffic_ncu.csv python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2j_traffic.log 2>&1
for t in 20 30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2j_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2j_ncu_tick$t.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2j_loop8_tick13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2j_ncu_loop_tick.log 2>&1
ls -la $O/r2j*.ncu-rep
