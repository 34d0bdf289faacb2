#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # name lib NL
  LOBSIM_NATIVE_LIB=$2 LOBSIM_BENCH_NL=$3 timeout 400 python bench.py --workload rollout --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ef_$1.log 2>&1
  python - $1 <<'P'
import json, sys
try:
    l = json.loads(open(f"gpurun_out/ef_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:>22s}: {l['value']:.4e} env steps/s, kernel only {l['env_step_kernel_only_steps_per_sec']:.4e}, flat {l['book_forms']['flat_fraction']:.2f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(f"gpurun_out/ef_{sys.argv[1]}.log").read()[-500:])
P
}
run base_NL64 "" 64
run flatonly_NL128 $PWD/rl4mm_b200/_native/variants/liblobsim_flatonly.so 128
