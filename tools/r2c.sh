#!/bin/bash
# GPU call C (one GPU): A/B of kernel variants built by tools/build_variants.sh (serf_b200/ab/libserfsim_<name>.so):
# bench (10 steps) + per-tick profile each; GRIDMUL sweep on the default library.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
run_variant() {   # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2c_bench_$name.json 2>> $O/r2c_bench.err
  env "$@" timeout 300 python tools/tick_profile.py --out $O/r2c_ticks_$name.json > $O/r2c_ticks_$name.log 2>&1
  echo "$name: $(python -c "import json;d=json.load(open('$O/r2c_bench_$name.json'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))" 2>/dev/null)"
  python -c "import json;d=json.load(open('$O/r2c_ticks_$name.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows']))" 2>/dev/null
}
for lib in serf_b200/ab/libserfsim_*.so; do
  n=$(basename $lib .so); n=${n#libserfsim_}
  run_variant $n SERFSIM_LIB=$PWD/$lib
done
run_variant main_grid1 SERFSIM_GRIDMUL=1
run_variant main_grid4 SERFSIM_GRIDMUL=4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config1 or fuzz or multi_slot or failure" > $O/r2c_tests.log 2>&1; tail -2 $O/r2c_tests.log
tail -3 $O/r2c_bench.err
