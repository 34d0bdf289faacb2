#!/bin/bash
# compute-sanitizer on a small end-to-end invocation (smoke): memcheck + racecheck (shared-memory hazards in the
# warp-synchronous book code) + synccheck
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edges.py -m gpu -q --tb=short --timeout 300 > gpurun_out/pytest_edges.log 2>&1; echo "edges rc=$?" >> gpurun_out/pytest_edges.log
for tool in racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"
done
tail -15 gpurun_out/pytest_edges.log
