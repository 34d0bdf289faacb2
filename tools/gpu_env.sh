#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --tb=short --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
LOBSIM_FORCE_GENERAL=1 timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --tb=short --timeout 600 > gpurun_out/pytest_gpu_general.log 2>&1; echo "pytest(general) rc=$?" >> gpurun_out/pytest_gpu_general.log
timeout 900 python bench.py --workload rollout --envs-per-gpu 65536 --steps 3 --warmup 3 > gpurun_out/bench_rollout.log 2>&1; echo "rollout rc=$?" >> gpurun_out/bench_rollout.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-200; tail -4 gpurun_out/pytest_gpu_general.log; tail -3 gpurun_out/bench_rollout.log | cut -c1-330
