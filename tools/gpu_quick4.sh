#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in 8 5 4 3; do
  LOBSIM_ENV_WPC=$w timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_mt_$w.log 2>&1
  python - $w <<'P'
import json, sys
l = json.loads(open(f"gpurun_out/q_mt_{sys.argv[1]}.log").read().strip().splitlines()[-1])
print(f"env warps per CTA {sys.argv[1]}: replay {l['value']:.4e} env {l['env']['value']:.4e}")
P
done
