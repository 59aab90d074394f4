#!/bin/bash
# GPU call X of round 2 (two GPUs): sharded parity and the default bench at N = 2 with the final host logic (launch chunks, jump rule, window padding).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/multi_parity.py > $O/r2x_parity_w2.log 2> $O/r2x_parity_w2.err
echo "parity world 2 rc=$?"; grep -E "^world" $O/r2x_parity_w2.log | grep -vc " ok"
timeout 600 $TR --nproc-per-node 2 --master-port 29530 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err
echo "bench rc=$?"; grep '^{' $O/r2x_bench_n2.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('N=%d %.4g eu/s kernel %.3f step %.3f e2e %.4g (%.3f ms) launches %d' % (d['n_gpus'], d['value'], d['kernel_ms_per_step'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']), d['self_check'] if isinstance(d['self_check'],str) else 'checked')"
