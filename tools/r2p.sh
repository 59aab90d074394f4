#!/bin/bash
# GPU call P of round 2 (one GPU): the ncu evidence of the shipped binary for the default workload — launch list of the bench command
# (share of the step per kernel), DRAM bytes of every tick launch of one run (→ profiles/r2_traffic_leave_fail.json, read by bench.py),
# then the bench itself (with the CPU baseline) and the reference arm.  TAG names the output files.
TAG=${1:-r2p}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > $O/${TAG}_launches.log 2>&1
echo "launch list rc=$? lines $(wc -l < $O/${TAG}_launches.csv)"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tick_kernel --clock-control none --csv \
    --log-file $O/${TAG}_traffic_ncu.csv python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/${TAG}_traffic.log 2>&1
echo "traffic rc=$?"; tail -1 $O/${TAG}_traffic.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check')); print(d.get('cpu_baseline')); print(d.get('host'))"; }
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; summ $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 900 python bench.py --no-numa-bind --no-cpu-baseline --no-check > $O/${TAG}_bench_unbound.json 2> $O/${TAG}_bench_unbound.err; echo "bench unbound rc=$?"; summ $O/${TAG}_bench_unbound.json
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference rc=$?"; cut -c1-400 $O/${TAG}_bench_reference.json
