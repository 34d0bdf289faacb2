#!/bin/bash
# the small-batch PPO collection loop with and without the CUDA graph: ms per env step at 4 096 envs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for g in on off; do
  timeout 600 python examples/train_ppo.py --envs 4096 --iterations 4 --rollout-steps 128 --episode-seconds 30 --n-msgs 1000000 --duration-s 2340 --cuda-graph $g > gpurun_out/ppo_graph_$g.log 2>&1
  python - "$g" <<'P'
import json, sys
g = sys.argv[1]
rows = [json.loads(l) for l in open(f"gpurun_out/ppo_graph_{g}.log") if l.startswith("{")]
for r in rows[1:]:
    print(f"cuda-graph {g}: collect {1e3 * r['collect_s'] / 128:.3f} ms per env step of 4096 envs ({r['env_steps_per_sec']:.3g} env steps/s), update {r['update_s']:.3f} s")
P
done
