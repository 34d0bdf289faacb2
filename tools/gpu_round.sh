#!/bin/bash
# one GPU call for a development round: GPU suite, smoke, replay A/B (flat pools vs sorted arrays) with instruction counters,
# launch lists + one full ncu capture per hot kernel, the PPO CUDA-graph timing, the full bench line.
# usage: gpu_round.sh [skip-tests] [skip-prof] [skip-bench]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
if [[ "$*" != *skip-tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -n 4 --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
fi
METRICS=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for flat in 1 0; do
  LOBSIM_REPLAY_FLAT=$flat timeout 400 python bench.py --workload replay --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat$flat.log 2>&1
  echo "flat=$flat: $(tail -1 gpurun_out/bench_flat$flat.log | cut -c1-120)"
  if [[ "$*" != *skip-prof* ]]; then
    LOBSIM_REPLAY_FLAT=$flat timeout 600 ncu --metrics $METRICS --clock-control none -k regex:k_replay -s 3 -c 1 --csv --log-file gpurun_out/ncu_flat$flat.csv python bench.py --workload replay --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    grep -v "^==" gpurun_out/ncu_flat$flat.csv | tail -8 | cut -d, -f5,13-
  fi
done
if [[ "$*" != *skip-prof* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --workload replay --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 4 -c 1 -f -o gpurun_out/prof_replay python bench.py --workload replay --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 60 --csv --log-file gpurun_out/launches_env.csv python bench.py --workload rollout --steps 1 --warmup 3 --sub-steps 1 --no-cpu-baseline > gpurun_out/bench_env_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_fast -s 40 -c 1 -f -o gpurun_out/prof_env python bench.py --workload rollout --steps 1 --warmup 3 --sub-steps 1 --no-cpu-baseline > gpurun_out/ncu_env.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
bash tools/gpu_ppo_graph.sh 2>&1 | tail -8
if [[ "$*" != *skip-bench* ]]; then
  ( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
  tail -4 gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-200
fi
