#!/bin/bash
# the GPU suite on the experiment build with the HOT / DEFERRED env kernels compiled in (python -m rl4mm_b200.build --variant hot
# -DLOBSIM_ENV_HOT_LAYOUTS=1): keeps the compiled-out path honest
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LOBSIM_NATIVE_LIB=$PWD/rl4mm_b200/_native/variants/liblobsim_hot.so timeout 1500 python -m pytest tests -q -m gpu -n 4 --timeout 900 > gpurun_out/pytest_hot.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hot.log
tail -5 gpurun_out/pytest_hot.log
