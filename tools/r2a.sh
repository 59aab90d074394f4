#!/bin/bash
# GPU call A of round 2 (one GPU): device parity of the paths that had none, A/B numbers of the kernel variants, the
# issue-rate micro-benchmark, and the ncu evidence for the binary at HEAD.  Everything lands in gpurun_out/r2a_*.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_tma.py tests/test_gpu_z_multiphase.py -m gpu -q > $O/r2a_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2a_tests.log
run_variant() {   # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2a_bench_$name.json 2>> $O/r2a_bench.err
  env "$@" timeout 300 python tools/tick_profile.py --out $O/r2a_ticks_$name.json > $O/r2a_ticks_$name.log 2>&1
  echo "$name: $(python -c "import json;d=json.load(open('$O/r2a_bench_$name.json'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))" 2>/dev/null)"
}
run_variant main
run_variant nocompact SERFSIM_COMPACT=0
run_variant tma SERFSIM_TMA=1
run_variant chunk8 SERFSIM_CHUNK=8
run_variant spec8 SERFSIM_CHUNK=8 SERFSIM_SPECULATE=1
[ -f serf_b200/ab/libserfsim_ab-no-queue-word.so ] && run_variant noqw SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_ab-no-queue-word.so
timeout 120 tools/ubench/lsu_red > $O/r2a_ubench.txt 2>&1; cat $O/r2a_ubench.txt
# ncu: launch list of the bench command, DRAM bytes of every tick launch, full captures of four ticks
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2a_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:tick_kernel --clock-control none --csv \
    --log-file $O/r2a_traffic_ncu.csv python tools/tick_profile.py --runs 1 > $O/r2a_traffic.log 2>&1
for t in 13 9 18 28; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip $t --launch-count 1 -f -o $O/r2a_tick$t \
      python tools/tick_profile.py --runs 1 > $O/r2a_ncu_tick$t.log 2>&1
done
ls -la $O/*.ncu-rep
