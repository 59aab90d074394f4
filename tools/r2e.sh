#!/bin/bash
# GPU call E of round 2 (one GPU): the whole single-GPU device suite on the binary at HEAD, the default bench workload
# (leave + fail, 2 tracked subjects, per-step oracle check) and the round-1 workload, per-tick profiles, the reference arm,
# the RED-path micro-benchmark, and the ncu evidence (launch list, per-launch DRAM bytes, full captures) of the default workload.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2e_gpu.txt 2>&1
nproc > $O/r2e_host.txt; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" >> $O/r2e_host.txt
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2e_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/r2e_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  eu %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['edge_updates_per_step'], d['gpu_launches'])); print(d['self_check']); print(d.get('cpu_baseline'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2e_bench_leave_fail.json 2> $O/r2e_bench_leave_fail.err; echo "bench leave_fail rc=$?"
summ $O/r2e_bench_leave_fail.json; tail -3 $O/r2e_bench_leave_fail.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline > $O/r2e_bench_leave.json 2> $O/r2e_bench_leave.err; echo "bench leave rc=$?"
summ $O/r2e_bench_leave.json
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2e_ticks_leave_fail.json > $O/r2e_ticks_leave_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2e_ticks_leave_fail.json'));print(d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))"
timeout 300 python tools/tick_profile.py --out $O/r2e_ticks_leave.json > $O/r2e_ticks_leave.log 2>&1
python -c "import json;d=json.load(open('$O/r2e_ticks_leave.json'));print(d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/r2e_bench_reference.json 2> $O/r2e_bench_reference.err; cut -c1-700 $O/r2e_bench_reference.json
timeout 200 tools/ubench/red_paths > $O/r2e_ubench_red_paths.txt 2>&1; cat $O/r2e_ubench_red_paths.txt
# ncu: launch list of the bench command, DRAM bytes of every tick launch, full captures of three ticks of the default workload
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2e_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-check > $O/r2e_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tick_kernel --clock-control none --csv \
    --log-file $O/r2e_traffic_ncu.csv python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2e_traffic.log 2>&1
for t in 12 20 30; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2e_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2e_ncu_tick$t.log 2>&1
done
ls -la $O/*.ncu-rep
