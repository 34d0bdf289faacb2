#!/bin/bash
# compute-sanitizer memcheck + racecheck + synccheck on the three small end-to-end drivers (smoke, tiny_flat4, tiny_hybrid)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for drv in "__graft_entry__.py smoke" "tools/tiny_flat4.py" "tools/tiny_hybrid.py"; do
  tag=$(echo $drv | tr -c 'a-zA-Z0-9' '_' | cut -c1-24)
  for tool in memcheck racecheck synccheck; do
    timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python $drv > gpurun_out/san_${tag}_$tool.log 2>&1
    echo "== $drv / $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|tiny_flat4 ok|tiny_hybrid ok' gpurun_out/san_${tag}_$tool.log | cut -c1-90 | tr '\n' ' ')"
  done
done
