#!/bin/bash
# hybrid deep-book replay kernel, final checks: GPU suite + smoke, compute-sanitizer on tools/tiny_hybrid.py, ncu capture, bench A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -n 4 --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 300 python tools/tiny_hybrid.py > gpurun_out/tiny_hybrid.log 2>&1; tail -2 gpurun_out/tiny_hybrid.log
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/tiny_hybrid.py > gpurun_out/sanitizer_hyb_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|tiny_hybrid ok' gpurun_out/sanitizer_hyb_$tool.log | tr '\n' ' ')"
done
ARGS="--workload multiticker --no-cpu-baseline --sub-steps 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_ -s 3 -c 1 -f -o gpurun_out/prof_c5_hyb python bench.py $ARGS > gpurun_out/ncu_c5_hyb.log 2>&1
for hyb in 1 0; do
  LOBSIM_REPLAY_HYBRID=$hyb timeout 400 python bench.py --workload multiticker --no-cpu-baseline > gpurun_out/bench_hyb$hyb.log 2>&1
  echo "hyb=$hyb: $(tail -1 gpurun_out/bench_hyb$hyb.log | cut -c1-120)"
done
