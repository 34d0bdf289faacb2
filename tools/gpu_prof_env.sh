#!/bin/bash
# ncu captures of the env step kernel in the rollout bench (configs[2]): launch list + one full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS="--workload rollout --no-cpu-baseline --sub-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file gpurun_out/launches_env.csv python bench.py $ARGS > gpurun_out/bench_env_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_fast -s 300 -c 1 -f -o gpurun_out/prof_env python bench.py $ARGS > gpurun_out/ncu_env.log 2>&1
tail -3 gpurun_out/ncu_env.log; ls -la gpurun_out | tail -5
