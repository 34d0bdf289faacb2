#!/bin/bash
# what the driver runs at round end: the GPU suite (sequential), smoke, the bench line and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; tail -4 gpurun_out/bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/bench_ref.err; tail -4 gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.log | cut -c1-300
