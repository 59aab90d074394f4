#!/bin/bash
# Line coverage of the KERNEL SOURCES under the emulated parity suites: builds the host-compiled library with
# --coverage, runs tests/test_emu_*.py against it, and prints per-file coverage plus every kernel line no test reached.
# Usage: bash tools/emu_coverage.sh [summary-file]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$ROOT/.scratch/cov
rm -rf "$W" && mkdir -p "$W" && cd "$W"
SRC=""
for f in "$ROOT"/serf_b200/csrc/*.cu "$ROOT"/tests/emu/emu_engine.cpp; do SRC="$SRC -x c++ $f"; done
g++ -std=c++17 -O0 -g --coverage -fPIC -shared -DSERFSIM_EMU -I"$ROOT/tests/emu" -Wno-unknown-pragmas -o libserfsim_emu_cov.so $SRC
(cd "$ROOT" && SERFSIM_EMU_LIB=$W/libserfsim_emu_cov.so python -m pytest tests/test_emu_parity.py tests/test_emu_multi.py tests/test_emu_uevent.py tests/test_emu_byzantine.py tests/test_emu_host.py -q 2>&1 | tail -1)
OUT=${1:-/dev/stdout}; case "$OUT" in /*) ;; *) OUT="$ROOT/$OUT";; esac
{
  echo "# kernel-source line coverage under tests/test_emu_*.py (host-compiled kernels, tests/emu) — $(date -u +%F)"
  for g in *.gcno; do gcov -o . "$g" >/dev/null 2>&1; done
  echo "# a template / inline line counts as reached if ANY instantiation or translation unit ran it"
  python3 - <<'PY'
import collections, glob, re
for path in sorted(glob.glob("*.cu.gcov")):
    cnt, src = {}, {}
    for l in open(path, errors="replace"):
        m = re.match(r"\s*([0-9#=\-]+\*?):\s*(\d+):(.*)", l)
        if not m or m.group(2) == "0":
            continue
        c, ln = m.group(1), int(m.group(2))
        src[ln] = m.group(3)
        if c.startswith("-"):
            continue
        v = 0 if c[0] in "#=" else int(c.rstrip("*"))
        cnt[ln] = max(cnt.get(ln, 0), v)          # template instances: a line is missed only if NO instance ran it
    missed = [ln for ln in sorted(cnt) if cnt[ln] == 0]
    print(f"{path[:-5]}: {len(cnt)} executable lines, {len(missed)} never executed")
    for ln in missed:
        print(f"    {ln}: {src[ln].strip()[:140]}")
PY
} > "$OUT"
