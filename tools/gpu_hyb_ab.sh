#!/bin/bash
# A/B of the deep-book replay: hybrid (default) vs flat/sorted (LOBSIM_REPLAY_HYBRID=0), alternating on one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for hyb in 1 0 1 0; do
  LOBSIM_REPLAY_HYBRID=$hyb timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hyb$hyb.log 2>&1
  echo "hyb=$hyb: $(tail -1 gpurun_out/bench_hyb$hyb.log | cut -c1-120)"
done
