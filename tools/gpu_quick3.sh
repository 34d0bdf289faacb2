#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "rollout 128" "rollout 64" "collect 128" "collect 64"; do
  set -- $cfg
  LOBSIM_BENCH_NL=$2 timeout 400 python bench.py --workload $1 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_$1_$2.log 2>&1
  python - $1 $2 <<'P'
import json, sys
try:
    l = json.loads(open(f"gpurun_out/q_{sys.argv[1]}_{sys.argv[2]}.log").read().strip().splitlines()[-1])
    extra = {k: l[k] for k in ("env_step_kernel_only_steps_per_sec", "book_forms") if k in l}
    print(f"{sys.argv[1]} NL={sys.argv[2]}: value {l['value']:.4e} ms/step {l['ms_per_step']:.2f} {extra}")
except Exception as e:
    print(sys.argv, "FAILED", e, open(f"gpurun_out/q_{sys.argv[1]}_{sys.argv[2]}.log").read()[-600:])
P
done
