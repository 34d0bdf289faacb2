#!/bin/bash
# end-of-round pass: GPU tests on both kernel families, the two bench lines, sanitizers, launch lists + full captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -n 4 --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
LOBSIM_FORCE_GENERAL=1 timeout 600 python -m pytest tests -m gpu -q -n 4 --timeout 200 > gpurun_out/pytest_gpu_general.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_general.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
timeout 600 python bench.py --workload rollout --envs-per-gpu 65536 --steps 3 --warmup 3 > gpurun_out/bench_rollout.log 2>&1; echo "rollout rc=$?" >> gpurun_out/bench_rollout.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
done
timeout 1500 bash tools/gpu_prof_all.sh > gpurun_out/prof_all.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu_general.log; tail -2 gpurun_out/bench.log | cut -c1-200; tail -1 gpurun_out/bench_reference.log | cut -c1-200; tail -2 gpurun_out/bench_rollout.log | cut -c1-200
for tool in memcheck racecheck synccheck; do echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitizer_$tool.log | tr '\n' ' ')"; done
