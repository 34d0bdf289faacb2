#!/bin/bash
# flat-pool replay kernel: the replay parity tests (flat / sorted / general), then the A/B bench lines and instruction counts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 -x -k "replay or config2 or message_free or stepping_past" > gpurun_out/pytest_flat.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_flat.log
tail -5 gpurun_out/pytest_flat.log
for flat in 1 0; do
  LOBSIM_REPLAY_FLAT=$flat timeout 400 python bench.py --workload replay --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat$flat.log 2>&1
  echo "flat=$flat: $(tail -1 gpurun_out/bench_flat$flat.log | cut -c1-160)"
  LOBSIM_REPLAY_FLAT=$flat timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_replay -s 3 -c 1 --csv --log-file gpurun_out/ncu_flat$flat.csv python bench.py --workload replay --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  grep -v "^==" gpurun_out/ncu_flat$flat.csv | tail -6 | cut -d, -f5,13-
done
