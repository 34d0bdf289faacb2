#!/bin/bash
# GPU suite (both kernel families) + A/B of the given variants on the given workload: tools/gpu_check_ab.sh <workload> <variants...>
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
bash tools/gpu_ab.sh "$@"
