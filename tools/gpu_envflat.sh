#!/bin/bash
# env kernel: flat blobs on / off -- env steps/s, share of flat books, instruction counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
METRICS=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum
for flat in 1 0; do
  LOBSIM_FLAT_BLOBS=$flat timeout 400 python bench.py --workload rollout --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/envflat$flat.log 2>&1
  python - $flat <<'P'
import json, sys
l = json.loads(open(f"gpurun_out/envflat{sys.argv[1]}.log").read().strip().splitlines()[-1])
print(f"FLAT_BLOBS={sys.argv[1]}: {l['value']:.4e} env steps/s, kernel only {l['env_step_kernel_only_steps_per_sec']:.4e}, book forms {l.get('book_forms')}")
P
  LOBSIM_FLAT_BLOBS=$flat timeout 600 ncu --metrics $METRICS --clock-control none -k regex:k_env_fast -s 300 -c 1 --csv --log-file gpurun_out/ncu_envflat$flat.csv python bench.py --workload rollout --steps 1 --warmup 3 --sub-steps 1 --no-cpu-baseline > /dev/null 2>&1
  grep -v "^==" gpurun_out/ncu_envflat$flat.csv | tail -7 | cut -d, -f13-
done
LOBSIM_FLAT_BLOBS=1 timeout 400 python bench.py --workload multiticker --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/mt.log 2>&1
python - <<'P'
import json
l = json.loads(open("gpurun_out/mt.log").read().strip().splitlines()[-1])
print(f"multiticker replay {l['value']:.4e} msgs/s; env {l['env']['value']:.4e} env steps/s; overflow envs {l['env']['agent_overflow_envs']}")
P
