#!/bin/bash
# GPU call S of round 2 (N GPUs, default 2): sharded parity on real GPUs with the single-view dispatch (dual launch, then check mode), bench of both
# workloads at N.
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29511 tools/multi_parity.py > $O/r2s_parity_w$N.log 2> $O/r2s_parity_w$N.err
echo "parity world $N rc=$?"; grep -E "^world" $O/r2s_parity_w$N.log | cut -c1-150
SERFSIM_SV=2 timeout 600 $TR --nproc-per-node $N --master-port 29512 tools/multi_parity.py > $O/r2s_parity_w${N}_sv2.log 2> $O/r2s_parity_w${N}_sv2.err
echo "parity sv=2 world $N rc=$?"; grep -E "^world" $O/r2s_parity_w${N}_sv2.log | grep -v " ok" | cut -c1-150
summ() { python -c "import json;d=json.load(open('$1'));print('N=%d %.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['n_gpus'], d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches'])); print(d.get('self_check'))"; }
for wl in leave_fail leave; do
  timeout 600 $TR --nproc-per-node $N --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > $O/r2s_bench_n${N}_$wl.json 2> $O/r2s_bench_n${N}_$wl.err
  echo "bench n=$N $wl rc=$?"; summ $O/r2s_bench_n${N}_$wl.json; tail -2 $O/r2s_bench_n${N}_$wl.err | cut -c1-300
done
