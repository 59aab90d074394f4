#!/bin/bash
# GPU call Z of round 2 (one GPU): compute-sanitizer over the SHARDED kernels (cross-shard staging, flush, fused publish, drain; single-view dispatch)
# through the loopback aid on a small shard.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
for tool in memcheck racecheck; do
  for a in "--world 8 --fail" "--world 4"; do
    tag=$(echo $a | tr -d ' -')
    timeout 60 compute-sanitizer --tool $tool --error-exitcode 9 python tools/loopback_profile.py --nodes 120000 $a --runs 1 > $O/r2z_sanitizer_${tool}_$tag.log 2>&1
    echo "$tool $a rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|loopback, shard" $O/r2z_sanitizer_${tool}_$tag.log | tail -2
  done
done
