#!/bin/bash
# GPU call L of round 2 (one GPU): HEAD after the sharded send-path diet — device suite, both bench workloads (default one with the
# per-step oracle check and the CPU baseline), per-tick profile of the default workload, loopback profile of the sharded path
# (fused publish on / off).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2l_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2l_tests.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check')); print(d.get('cpu_baseline'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2l_bench_leave_fail.json 2> $O/r2l_bench_leave_fail.err; echo "bench leave_fail rc=$?"; summ $O/r2l_bench_leave_fail.json; tail -2 $O/r2l_bench_leave_fail.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline > $O/r2l_bench_leave.json 2> $O/r2l_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2l_bench_leave.json
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2l_ticks_storm_fail.json > $O/r2l_ticks_storm_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2l_ticks_storm_fail.json'));print('storm_fail', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]))"
loop() { name=$1; shift; for a in "--world 8" "--world 2" "--world 8 --fail"; do
    tag=$(echo $a | tr -d ' -'); env "$@" SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py $a --out $O/r2l_loop_${name}_$tag.json > $O/r2l_loop_${name}_$tag.log 2>&1
    echo "$name $a: $(tail -1 $O/r2l_loop_${name}_$tag.log)"; grep -E "^rank 0 tick 13|^rank 0:" $O/r2l_loop_${name}_$tag.log | tail -2
    python -c "import json;d=json.load(open('$O/r2l_loop_${name}_$tag.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:40]))"
  done; }
loop main
loop nofuse SERFSIM_NO_FUSE=1
