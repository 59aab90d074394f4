#!/bin/bash
# GPU call W of round 2 (one GPU): the device suite and the default bench command on the final binary.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multi.py > $O/r2w_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/r2w_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2w_smoke.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check')); print(d.get('cpu_baseline')); print(d['roofline']['traffic'], d['roofline']['algorithmic_bytes_per_launch'])"; }
timeout 900 python bench.py > $O/r2w_bench.json 2> $O/r2w_bench.err; echo "bench rc=$?"; summ $O/r2w_bench.json; tail -2 $O/r2w_bench.err
