#!/bin/bash
# GPU call Y of round 2 (one GPU, the last one): compute-sanitizer memcheck + racecheck over small parity scenarios of the final binary, incl. the
# multi-slot kernel's new paths (requests one tile ahead forced, single-view dispatch, check mode) — logs kept under profiles/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
O=gpurun_out
SEL="tests/test_gpu_parity.py::test_config0_full_mesh_256[1] tests/test_gpu_parity.py::test_fuzz[3] tests/test_gpu_parity.py::test_fuzz[5] tests/test_gpu_parity.py::test_fuzz[11] tests/test_gpu_parity.py::test_fuzz[16] tests/test_gpu_parity.py::test_fuzz_prune[5] tests/test_gpu_z_multislot_paths.py::test_fuzz"
for tool in memcheck racecheck; do
  timeout 125 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -x -q > $O/r2y_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 $O/r2y_sanitizer_$tool.log
done
