#!/bin/bash
# A/B of the step scheduling knobs on one box: python bench.py per (workload, env knobs) -> one line each
mkdir -p gpurun_out
run() {  # tag workload [env assignments...]
  tag=$1; w=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $w --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/{tag}.json"))
    print(f"{tag:44s} ms/step {d['ms_per_step']:.4f}  veh-steps/s {d['value']:.4g}  e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.3f} ms)  launches {d['gpu_launches']}")
except Exception as e:
    print(tag, "FAILED", e, open(f"gpurun_out/{tag}.err").read()[-400:])
PY
}
"$@"
runx() {  # like run, also prints the episode statistics that tell how often the synchronous reset fallback fired
  tag=$1; run "$@"
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/{sys.argv[1]}.json")); s=d["episode_stats"]
    print(f"    episodes {s['episodes']:.0f}  sync_resets {s.get('sync_resets')}  per step {s.get('sync_resets',0)/d['steps']:.1f}")
except Exception as e:
    print("   ", e)
PY
}
