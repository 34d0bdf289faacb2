#!/bin/bash
# A/B of experimental builds on one box: tools/gpu_ab.sh <workload> <variant> [<variant> ...]   ("base" = the product library)
#   built beforehand with: python -m rl4mm_b200.build --variant NAME -DFLAG=...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
wl=$1; shift
for v in "$@"; do
  lib=""; [ "$v" != base ] && lib="$PWD/rl4mm_b200/_native/variants/liblobsim_$v.so"
  LOBSIM_NATIVE_LIB=$lib timeout 600 python bench.py --workload $wl --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/ab_${wl}_$v.log 2>&1
  python - "$v" gpurun_out/ab_${wl}_$v.log <<'P'
import json, sys
try:
    l = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    extra = {k: l[k] for k in ("env_step_kernel_only_steps_per_sec", "lob_messages_per_sec") if k in l}
    if "env" in l: extra["env"] = l["env"]["value"]
    print(f"{sys.argv[1]:>16s}: value {l['value']:.4e}  ms/step {l['ms_per_step']:.3f}  {extra}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-600:])
P
done
