#!/bin/bash
# GPU call Q of round 2 (one GPU): single-view ticks (SERFSIM_SV) — device suite with the dual launch, the multi-slot parity files again in
# check mode (SERFSIM_SV=2), A/B bench of the default workload, the single-slot workload (spill check of the R1 kernel), per-tick profile,
# loopback profile of the sharded leave + fail run.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > $O/r2q_tests.log 2>&1
echo "tests rc=$?"; tail -3 $O/r2q_tests.log
SERFSIM_SV=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_z_fullsize.py tests/test_gpu_z_multiphase.py -m gpu -q -x > $O/r2q_tests_sv2.log 2>&1
echo "tests sv=2 rc=$?"; tail -3 $O/r2q_tests_sv2.log
summ() { python -c "import json;d=json.load(open('$1'));print('%.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2q_bench_lf_sv1.json 2> $O/r2q_bench_lf_sv1.err; echo "bench sv=1 rc=$?"; summ $O/r2q_bench_lf_sv1.json; tail -2 $O/r2q_bench_lf_sv1.err
SERFSIM_SV=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-check > $O/r2q_bench_lf_sv0.json 2> $O/r2q_bench_lf_sv0.err; echo "bench sv=0 rc=$?"; summ $O/r2q_bench_lf_sv0.json
timeout 600 python bench.py --steps 10 --warmup 3 --workload leave --no-cpu-baseline --no-check > $O/r2q_bench_leave.json 2> $O/r2q_bench_leave.err; echo "bench leave rc=$?"; summ $O/r2q_bench_leave.json
timeout 300 python tools/tick_profile.py --scenario storm_fail --out $O/r2q_ticks_storm_fail.json > $O/r2q_ticks_storm_fail.log 2>&1
python -c "import json;d=json.load(open('$O/r2q_ticks_storm_fail.json'));print('storm_fail', d['kernel_ms']);print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:60]));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][145:175]))"
for sv in 1 0; do
  SERFSIM_SV=$sv SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py --world 8 --fail --out $O/r2q_loop_w8fail_sv$sv.json > $O/r2q_loop_w8fail_sv$sv.log 2>&1
  echo "loop w8 fail sv=$sv: $(tail -1 $O/r2q_loop_w8fail_sv$sv.log)"; grep -E "^rank 0:" $O/r2q_loop_w8fail_sv$sv.log | tail -1
  python -c "import json;d=json.load(open('$O/r2q_loop_w8fail_sv$sv.json'));print(' '.join('%d'%(1e3*r['ms']) for r in d['rows'][:50]))"
done
