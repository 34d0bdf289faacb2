#!/bin/bash
# replay kernels after a change: the replay-path tests of every family, then the headline replay bench line + instruction counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 -x -k "replay or config2 or config5 or message_free or stepping_past or hybrid or volume" > gpurun_out/pytest_replay.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_replay.log
tail -5 gpurun_out/pytest_replay.log
timeout 400 python bench.py --workload replay --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_replay.log 2>&1
echo "replay: $(tail -1 gpurun_out/bench_replay.log | cut -c1-160)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_replay -s 3 -c 1 --csv --log-file gpurun_out/ncu_replay_counters.csv python bench.py --workload replay --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_replay_counters.csv | tail -4 | cut -d, -f5,13-
for hyb in 1 0; do
  LOBSIM_REPLAY_HYBRID=$hyb timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hyb$hyb.log 2>&1
  echo "hyb=$hyb: $(tail -1 gpurun_out/bench_hyb$hyb.log | cut -c1-120)"
done
