#!/bin/bash
# compute-sanitizer on a small end-to-end invocation (smoke: k_env_fast, k_advance, k_replay_flat, k_replay_fast, k_to_sorted,
# k_get_state): memcheck + racecheck (shared-memory hazards in the warp-synchronous book code) + synccheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
