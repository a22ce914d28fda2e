#!/bin/bash
# Build tuning variants of the CUDA library next to the default one: tools/build_variants.sh name:"-DFLAG ..." ...
set -e
cd "$(dirname "$0")/../topotrafficrl_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared -Wno-deprecated-gpu-targets"
pids=""
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc $F $flags -o "var_${name}.so" ttrl_sim.cu ttrl_qnet.cu &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
ls -la var_*.so
