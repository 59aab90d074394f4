#!/bin/bash
# GPU call I of round 2 (EIGHT GPUs, charged 8x): sharded parity at HEAD on real GPUs (worlds 8, 4, 2; all scenarios incl. user
# events / injectors / push-pull / prune across shards; configs[4] and configs[2] at full size on 8 GPUs), the scaling bench of both
# workloads, per-tick sharded profile.  Logs are kept under profiles/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r2i_gpus.txt; nvidia-smi topo -m > $O/r2i_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for w in 8 2 4; do
  timeout 600 $TR --nproc-per-node $w --master-port $((29500 + w)) tools/multi_parity.py > $O/r2i_parity_w$w.log 2> $O/r2i_parity_w$w.err
  echo "parity world $w rc=$?"; grep -E "^world" $O/r2i_parity_w$w.log | cut -c1-160
done
timeout 900 $TR --nproc-per-node 8 --master-port 29520 tools/multi_parity.py --full > $O/r2i_parity_full_w8.log 2> $O/r2i_parity_full_w8.err
echo "parity full rc=$?"; grep -E "^world" $O/r2i_parity_full_w8.log | cut -c1-200
summ() { python -c "import json;d=json.load(open('$1'));print('N=%d %.4g eu/s  kernel %.3f ms  step %.3f ms  frac %.3f  e2e %.4g (%.3f ms)  ticks %d  launches %d' % (d['n_gpus'], d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ticks_to_convergence'], d['gpu_launches']))"; }
for n in 8 4 2; do
  for wl in leave_fail leave; do
    timeout 600 $TR --nproc-per-node $n --master-port $((29530 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-check --workload $wl > $O/r2i_bench_n${n}_$wl.json 2> $O/r2i_bench_n${n}_$wl.err
    echo "bench n=$n $wl rc=$?"; summ $O/r2i_bench_n${n}_$wl.json; tail -2 $O/r2i_bench_n${n}_$wl.err | cut -c1-300
  done
done
SERFSIM_XTIMING=1 timeout 300 $TR --nproc-per-node 8 --master-port 29550 tools/multi_profile.py > $O/r2i_multi_profile_n8.log 2>&1; grep -E "^rank 0|^world|^tick (1[0-9]|2[0-9]) " $O/r2i_multi_profile_n8.log | head -40
