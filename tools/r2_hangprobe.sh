#!/bin/bash
# which configuration of the sub-warp-team kernels hangs: tiny runs, each under its own short timeout
mkdir -p gpurun_out
probe() {  # tag lib workload envs [env assignments]
  tag=$1; lib=$2; w=$3; E=$4; shift 4
  start=$(date +%s.%N)
  env TTRL_B200_LIB=$PWD/topotrafficrl_b200/csrc/$lib "$@" timeout 45 python bench.py --workload $w --envs $E --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  rc=$?
  echo "$tag lib=$lib w=$w E=$E $* rc=$rc $(python -c "import json;d=json.load(open('gpurun_out/$tag.json'));print('ms/step',round(d['ms_per_step'],4))" 2>/dev/null)"
}
