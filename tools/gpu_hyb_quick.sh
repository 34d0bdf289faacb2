#!/bin/bash
# the deep-book replay bench (hybrid book) twice + its hybrid stress / config-5 tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -n 4 --timeout 600 -x -k "hybrid or config5" > gpurun_out/pytest_hyb.log 2>&1; tail -2 gpurun_out/pytest_hyb.log
for k in 1 2; do
  timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hyb1.log 2>&1
  echo "hyb: $(tail -1 gpurun_out/bench_hyb1.log | cut -c1-120)"
done
