#!/bin/bash
# GPU call U of round 2 (one GPU): the ncu evidence of the shipped binary for the default workload — DRAM bytes of every tick launch of one
# run (→ profiles/r2_traffic_leave_fail.json, read by bench.py), launch list of the bench command, full captures of a two-view tick (general
# kernel) and of a single-view tick (single-slot kernel) — and the bench itself with the CPU baseline.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tick_kernel --clock-control none --csv \
    --log-file $O/r2u_traffic_ncu.csv python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2u_traffic.log 2>&1
echo "traffic rc=$?"; tail -1 $O/r2u_traffic.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2u_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > $O/r2u_launches.log 2>&1
echo "launch list rc=$? lines $(wc -l < $O/r2u_launches.csv)"
# tick_kernel launches of one run: two per tick (general, single-view); tick t = launches 2t and 2t+1
for t in 20 40; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $((2 * t)) --launch-count 2 -f -o $O/r2u_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2u_ncu_tick$t.log 2>&1
done
ls -la $O/r2u*.ncu-rep
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check')); print(d.get('cpu_baseline')); print(d.get('host'))"; }
timeout 900 python bench.py > $O/r2u_bench.json 2> $O/r2u_bench.err; echo "bench rc=$?"; summ $O/r2u_bench.json; tail -2 $O/r2u_bench.err
