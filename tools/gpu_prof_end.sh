#!/bin/bash
# end-of-round captures: ncu --set full of the flat replay kernel (headline) and of the hybrid kernel (configs[4]), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_replay_end.csv python bench.py --workload replay --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_flat -s 4 -c 1 -f -o gpurun_out/prof_replay_end python bench.py --workload replay --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 3 -c 1 -f -o gpurun_out/prof_c5_hyb python bench.py --workload multiticker --no-cpu-baseline --sub-steps 1 > gpurun_out/ncu_c5_hyb.log 2>&1
ls -la gpurun_out/*.ncu-rep
