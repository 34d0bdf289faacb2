#!/bin/bash
# hybrid deep-book replay kernel: the replay tests of every family + the stress cases, then the A/B of the deep-book bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 -k "replay or config2 or config5 or message_free or stepping_past or hybrid" > gpurun_out/pytest_hyb.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hyb.log
tail -15 gpurun_out/pytest_hyb.log
for hyb in 1 0 1 0; do
  LOBSIM_REPLAY_HYBRID=$hyb timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hyb$hyb.log 2>&1
  echo "hyb=$hyb: $(tail -1 gpurun_out/bench_hyb$hyb.log | cut -c1-120)"
done
