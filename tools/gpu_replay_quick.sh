#!/bin/bash
# the headline replay bench line + instruction counters only (no tests)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python bench.py --workload replay --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_replay.log 2>&1
echo "replay: $(tail -1 gpurun_out/bench_replay.log | cut -c1-160)"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_replay -s 3 -c 1 --csv --log-file gpurun_out/ncu_replay_counters.csv python bench.py --workload replay --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_replay_counters.csv | tail -3 | cut -d, -f13-
