#!/bin/bash
# GPU call K of round 2 (one GPU): the sharded send path after its instruction diet (no divisions, ready-peer mask, reservation one
# flush ahead, publish fused into the tick kernel's last CTA) through the loopback aid; register budget A/B of the sharded kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_z_fullsize.py > $O/r2k_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2k_tests.log
loop() { name=$1; shift; for a in "--world 8" "--world 2" "--world 8 --fail"; do
    tag=$(echo $a | tr -d ' -'); env "$@" SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py $a --out $O/r2k_loop_${name}_$tag.json > $O/r2k_loop_${name}_$tag.log 2>&1
    echo "$name $a: $(tail -1 $O/r2k_loop_${name}_$tag.log)"; grep -E "^rank 0 tick 13|^rank 0:" $O/r2k_loop_${name}_$tag.log | tail -2
    python -c "import json;d=json.load(open('$O/r2k_loop_${name}_$tag.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:40]))"
  done; }
loop main
loop nofuse SERFSIM_NO_FUSE=1
loop r1s3 SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_r1s3.so
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step']))"; }
for wl in leave leave_fail; do timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-check --workload $wl > $O/r2k_bench_$wl.json 2> $O/r2k_bench_$wl.err; echo "bench $wl rc=$?"; summ $O/r2k_bench_$wl.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tick_kernel --launch-skip 13 --launch-count 1 -f -o $O/r2k_loop8_tick13 \
    python tools/loopback_profile.py --world 8 --runs 1 > $O/r2k_ncu_loop_tick.log 2>&1
ls -la $O/r2k*.ncu-rep
