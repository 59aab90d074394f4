#!/bin/bash
# GPU call V of round 2 (one GPU): launch chunks 8 / 16 / 32 restarting after every jump; the new device test file of the multi-slot paths;
# whole device suite; both bench workloads.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2v_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2v_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2v_bench_lf.json 2> $O/r2v_bench_lf.err; echo "bench lf rc=$?"; summ $O/r2v_bench_lf.json; tail -2 $O/r2v_bench_lf.err
SERFSIM_CHUNK=16 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-check > $O/r2v_bench_lf_c16.json 2> $O/r2v_bench_lf_c16.err; echo "bench lf chunk 16 rc=$?"; summ $O/r2v_bench_lf_c16.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline --no-check > $O/r2v_bench_leave.json 2> $O/r2v_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2v_bench_leave.json
