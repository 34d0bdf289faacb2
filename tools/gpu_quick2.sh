#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "config5" 2>&1 | tail -3
for wl in multiticker rollout; do
  timeout 400 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/q_$wl.log 2>&1
  python - $wl <<'P'
import json, sys
l = json.loads(open(f"gpurun_out/q_{sys.argv[1]}.log").read().strip().splitlines()[-1])
extra = {k: l[k] for k in ("env_step_kernel_only_steps_per_sec",) if k in l}
if "env" in l: extra["env"] = l["env"]["value"]; extra["overflow"] = l["env"]["agent_overflow_envs"]
print(f"{sys.argv[1]}: value {l['value']:.4e} ms/step {l['ms_per_step']:.2f} {extra}")
P
done
