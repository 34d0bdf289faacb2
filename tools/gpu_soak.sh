#!/bin/bash
# soak: the randomised differential tests on seeds outside the committed range (replay on all four implementations, env cases)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LOBSIM_RANDOM_SEED0=${SOAK_SEED0:-5000} LOBSIM_RANDOM_CASES=${SOAK_CASES:-96} LOBSIM_RANDOM_REPLAY_CASES=${SOAK_REPLAY_CASES:-160} timeout 2400 python -m pytest tests/test_gpu_random_diff.py -m gpu -q -n 6 --timeout 900 > gpurun_out/pytest_soak.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_soak.log
tail -6 gpurun_out/pytest_soak.log
