#!/bin/bash
# GPU call D (one GPU): the new default bench workload (leave + fail, 2 tracked subjects) with the per-step oracle check,
# the round-1 workload for continuity, per-tick profiles and ncu evidence of the multi-slot kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
nproc > $O/r2d_host.txt; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" >> $O/r2d_host.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2d_bench_leave_fail.json 2> $O/r2d_bench_leave_fail.err; echo "bench leave_fail rc=$?"
python -c "import json;d=json.load(open('$O/r2d_bench_leave_fail.json'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g  ticks %d  eu %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['ticks_to_convergence'], d['edge_updates_per_step'])); print(d['self_check']); print(d['cpu_baseline'])"
tail -3 $O/r2d_bench_leave_fail.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave > $O/r2d_bench_leave.json 2> $O/r2d_bench_leave.err; echo "bench leave rc=$?"
python -c "import json;d=json.load(open('$O/r2d_bench_leave.json'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])); print(d['self_check'])"
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2d_ticks_leave_fail.json > $O/r2d_ticks_leave_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2d_ticks_leave_fail.json'));print(d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/r2d_bench_reference.json 2> $O/r2d_bench_reference.err; cat $O/r2d_bench_reference.json | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_multiphase.py -m gpu -q -x > $O/r2d_tests.log 2>&1; tail -2 $O/r2d_tests.log
for t in 12 20; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2d_lf_tick$t \
      python tools/tick_profile.py --runs 1 --scenario storm_fail > $O/r2d_ncu_tick$t.log 2>&1
done
