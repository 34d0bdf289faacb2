#!/bin/bash
# ncu capture of the hybrid deep-book replay kernel on BASELINE configs[4]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export LOBSIM_REPLAY_HYBRID=1
ARGS="--workload multiticker --no-cpu-baseline --sub-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 3 -c 1 -f -o gpurun_out/prof_c5_hyb python bench.py $ARGS > gpurun_out/ncu_c5_hyb.log 2>&1
ls -la gpurun_out/prof_c5_hyb*
