#!/bin/bash
# launch lists + one full capture per hot kernel (replay fast, env fast); summaries are made on the build box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 4 -c 1 -f -o gpurun_out/prof_replay python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/launches_env.csv python bench.py --workload rollout --envs-per-gpu 65536 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_env_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_fast -s 40 -c 1 -f -o gpurun_out/prof_env python bench.py --workload rollout --envs-per-gpu 65536 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_env.log 2>&1
ls -la gpurun_out | tail -8
