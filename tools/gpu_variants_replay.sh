#!/bin/bash
# A/B of library variants (rl4mm_b200/_native/liblobsim_<name>.so) on the replay bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp rl4mm_b200/_native/liblobsim.so /tmp/orig.so
for v in "$@"; do
  [ "$v" = "cur" ] && cp /tmp/orig.so rl4mm_b200/_native/liblobsim.so || cp rl4mm_b200/_native/liblobsim_$v.so rl4mm_b200/_native/liblobsim.so
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_replay_$v.log 2>&1
  echo "$v: $(tail -1 gpurun_out/bench_replay_$v.log | cut -c1-110)"
done
cp /tmp/orig.so rl4mm_b200/_native/liblobsim.so
