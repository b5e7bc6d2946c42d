#!/bin/bash
# Usage (on the GPU box, via gpurun):  bash profiles/ncu_capture.sh <kernel-regex> <out-name> [skip] [count]
# One `ncu --set full` capture of the matching kernel from a short bench run (3 scenes, resident inputs).
set -u
REGEX="$1"; OUT="$2"; SKIP="${3:-3}"; COUNT="${4:-2}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:${REGEX}" -s "${SKIP}" -c "${COUNT}" -f -o "gpurun_out/${OUT}" \
    python bench.py --scenes 3 --steps 1 --warmup 1 --no-e2e --no-cpu > "gpurun_out/${OUT}.log" 2>&1
tail -2 "gpurun_out/${OUT}.log" | cut -c1-300
