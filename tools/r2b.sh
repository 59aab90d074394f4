#!/bin/bash
# GPU call B of round 2 (one GPU): the whole single-GPU device suite on the new binary (device-side convergence gate, per-warp
# CSR span staging through cp.async.bulk), then A/B of the staging and one full ncu capture of the plateau tick.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2b_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/r2b_tests.log
run_variant() {   # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2b_bench_$name.json 2>> $O/r2b_bench.err
  env "$@" timeout 300 python tools/tick_profile.py --out $O/r2b_ticks_$name.json > $O/r2b_ticks_$name.log 2>&1
  echo "$name: $(python -c "import json;d=json.load(open('$O/r2b_bench_$name.json'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))" 2>/dev/null)"
  python -c "import json;d=json.load(open('$O/r2b_ticks_$name.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))" 2>/dev/null
}
run_variant main
run_variant nowstage SERFSIM_WSTAGE=0
run_variant chunk8 SERFSIM_CHUNK=8
run_variant chunk32 SERFSIM_CHUNK=32
for t in 13 18; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2b_tick$t \
      python tools/tick_profile.py --runs 1 > $O/r2b_ncu_tick$t.log 2>&1
done
tail -3 $O/r2b_bench.err
