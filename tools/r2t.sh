#!/bin/bash
# GPU call T of round 2 (one GPU): cross-shard staging with one shared atomic per lane (main) against match_any + leader atomic (ab/libserfsim_xmatch.so),
# loopback aid.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
loop() { name=$1; shift; for a in "--world 8" "--world 2" "--world 8 --fail"; do
    tag=$(echo $a | tr -d ' -'); env "$@" SERFSIM_XTIMING=1 timeout 300 python tools/loopback_profile.py $a --out $O/r2t_loop_${name}_$tag.json > $O/r2t_loop_${name}_$tag.log 2>&1
    echo "$name $a: $(tail -1 $O/r2t_loop_${name}_$tag.log)"; grep -E "^rank 0 tick 13|^rank 0:" $O/r2t_loop_${name}_$tag.log | tail -2
  done; }
loop main
loop xmatch SERFSIM_LIB=$PWD/serf_b200/ab/libserfsim_xmatch.so
