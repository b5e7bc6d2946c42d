#!/bin/bash
# Build library variants for profiles/variant_sweep.sh:  bash profiles/build_variants.sh name1:"-DFLAG=1" name2:"-DA=1 -DB=0" ...
# Each becomes build/variants/<name>.so (build/ is git-ignored but travels to the GPU box with gpurun).  `base` = the shipped flags.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NV="${NVCC:-/usr/local/cuda/bin/nvcc} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared -I$ROOT/include"
mkdir -p "$ROOT/build/variants"
cd "$ROOT/multi_spatialmllm_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"; [ "$name" = "$spec" ] && flags=""
  ( $NV $flags -o "$ROOT/build/variants/$name.so" msp_kernels.cu -ldl && echo "built $name [$flags]" ) &
done
wait
