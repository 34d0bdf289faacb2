#!/bin/bash
# (N GPUs) the PPO loop under DDP: a few iterations of examples/train_ppo.py on every rank + the replica checksum; then, on rank 0's
# GPU, the learning-check test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 examples/train_ppo.py --envs 4096 --iterations 4 --rollout-steps 64 --n-msgs 600000 --duration-s 1500 > gpurun_out/ppo_ddp.log 2> gpurun_out/ppo_ddp.err; echo "rc=$?" >> gpurun_out/ppo_ddp.log
tail -6 gpurun_out/ppo_ddp.log | cut -c1-400; tail -3 gpurun_out/ppo_ddp.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "ppo" --timeout 500 > gpurun_out/pytest_ppo.log 2>&1; tail -3 gpurun_out/pytest_ppo.log
