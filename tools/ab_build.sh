#!/bin/bash
# Builds A/B variants of the C-ABI library into tools/bin/ab_<name>.so (git-ignored; they travel to the GPU box with gpurun):
#   tools/ab_build.sh "ctas5:-DVC_VERIFY_CTAS=5" "w2c4:-DVC_VERIFY_CTAS_W2=4" ...
# Build-time knobs: DESIGN.md section 9.  Register / spill counts of the verify kernels are printed per variant.
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$HERE/tools/bin"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( cd "$HERE/verticut_b200/csrc" && nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
      -Xptxas -v -shared $flags -o "$HERE/tools/bin/ab_$name.so" api.cu 2> "/tmp/ab_$name.log" ) &
done
wait
for spec in "$@"; do
  name=${spec%%:*}
  echo "== $name (${spec#*:})"
  grep -A2 "Compiling entry function '_ZN2vc18bmih_verify_kernelILi[12]ELb1ELi[248]E" "/tmp/ab_$name.log" | grep -E "spill|registers" | sed 's/ptxas info    : //' | paste - - | cut -c1-200
done
