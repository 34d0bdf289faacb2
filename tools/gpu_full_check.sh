#!/bin/bash
# after a kernel change: the whole GPU suite (4 workers) + the rollout / collect / multiticker bench records
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -n 4 --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for wl in rollout collect; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1
  python - $wl <<'P'
import json, sys
wl = sys.argv[1]
try:
    l = json.loads(open(f"gpurun_out/bench_{wl}.log").read().strip().splitlines()[-1])
    print(wl, "value %.4e" % l["value"], {k: ("%.4e" % l[k]) for k in ("env_step_kernel_only_steps_per_sec",) if k in l})
except Exception as e:
    print(wl, "FAILED", e)
P
done
