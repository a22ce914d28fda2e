#!/bin/bash
# Build tuning variants of the CUDA library next to the default one:
#   tools/build_variants.sh name:"-DFLAG ..." ...      (TT_VS limits the compiled capacities, default "50")
set -e
cd "$(dirname "$0")/../topotrafficrl_b200/csrc"
VS="${TT_VS:-50}"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  make --no-print-directory OUT="var_${name}.so" B="build_${name}" TT_VS="$VS" EXTRA="$flags" > /dev/null
done
ls -la var_*.so
