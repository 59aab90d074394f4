#!/bin/bash
# GPU call R of round 2 (one GPU): packed counters in the single-view kernels (R1: 12 B of spill instead of 68), lazy node_due load, warmed
# read-back path in bench.py — parity files, both bench workloads, per-tick profiles.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_z_fullsize.py > $O/r2r_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2r_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2r_bench_lf.json 2> $O/r2r_bench_lf.err; echo "bench lf rc=$?"; summ $O/r2r_bench_lf.json; tail -2 $O/r2r_bench_lf.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-check > $O/r2r_bench_lf2.json 2> $O/r2r_bench_lf2.err; echo "bench lf (2nd process) rc=$?"; summ $O/r2r_bench_lf2.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline > $O/r2r_bench_leave.json 2> $O/r2r_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2r_bench_leave.json
for wl in storm_fail storm; do
  timeout 300 python tools/tick_profile.py --scenario $wl --out $O/r2r_ticks_$wl.json > $O/r2r_ticks_$wl.log 2>&1
  python -c "import json;d=json.load(open('$O/r2r_ticks_$wl.json'));print('$wl', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]))"
done
