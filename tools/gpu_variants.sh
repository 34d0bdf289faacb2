#!/bin/bash
# A/B of library variants built next to liblobsim.so (rl4mm_b200/_native/liblobsim_<name>.so): rollout bench per variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp rl4mm_b200/_native/liblobsim.so /tmp/orig.so
for v in "$@"; do
  cp rl4mm_b200/_native/liblobsim_$v.so rl4mm_b200/_native/liblobsim.so
  timeout 600 python bench.py --workload rollout --envs-per-gpu 65536 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rollout_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/bench_rollout_$v.log | cut -c1-110)"
done
cp /tmp/orig.so rl4mm_b200/_native/liblobsim.so
