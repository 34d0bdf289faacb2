#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {  # name lib NL flat
  LOBSIM_NATIVE_LIB=$2 LOBSIM_BENCH_NL=$3 LOBSIM_FLAT_BLOBS=$4 timeout 400 python bench.py --workload rollout --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ef_$1.log 2>&1
  python - $1 <<'P'
import json, sys
try:
    l = json.loads(open(f"gpurun_out/ef_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:>22s}: {l['value']:.4e} env steps/s, kernel only {l['env_step_kernel_only_steps_per_sec']:.4e}, flat {l['book_forms']['flat_fraction']:.2f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(f"gpurun_out/ef_{sys.argv[1]}.log").read()[-500:])
P
}
V=$PWD/rl4mm_b200/_native/variants/liblobsim_noflat.so
run noflat_NL64 $V 64 0
run noflat_NL128 $V 128 0
run flatcode_NL64_off "" 64 0
run flatcode_NL64_on "" 64 1
run flatcode_NL128_off "" 128 0
run flatcode_NL128_on "" 128 1
