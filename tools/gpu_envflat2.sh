#!/bin/bash
# quick correctness (env + replay parity tests on all families) then env kernel flat blobs on / off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
bash tools/gpu_envflat.sh
LOBSIM_REPLAY_FLAT=1 timeout 400 python bench.py --workload replay --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat1.log 2>&1
echo "replay flat: $(tail -1 gpurun_out/bench_flat1.log | cut -c1-100)"
