#!/bin/bash
# ncu captures of BASELINE configs[4] (8 tickers, 50 levels, deep queues, 8 192 books): the replay kernel and the fused env kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
ARGS="--workload multiticker --no-cpu-baseline --sub-steps 1"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c5.csv python bench.py $ARGS > gpurun_out/bench_c5_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 3 -c 1 -f -o gpurun_out/prof_c5_replay python bench.py $ARGS > gpurun_out/ncu_c5_replay.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_fast -s 4 -c 1 -f -o gpurun_out/prof_c5_env python bench.py $ARGS > gpurun_out/ncu_c5_env.log 2>&1
ls -la gpurun_out/prof_c5*; grep -v "^==" gpurun_out/launches_c5.csv | cut -d, -f5,14- | tail -12
