#!/bin/bash
# Run bench.py once per library variant (on the GPU box): tools/bench_variants.sh default t64 b16 ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [ "$v" = default ]; then unset TTRL_B200_LIB; else export TTRL_B200_LIB=$PWD/topotrafficrl_b200/csrc/var_$v.so; fi
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'ms/step', round(d['ms_per_step'],3), 'veh-steps/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
done
