#!/usr/bin/env python
"""Train a Beta-ladder market-making policy with PPO on the device env (the reference's main.py, without RLlib).

    python examples/train_ppo.py --envs 4096 --iterations 20
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_ppo.py --envs 65536
"""
from __future__ import annotations

import argparse
import json
import sys
from datetime import datetime, timedelta
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl4mm_b200 import parallel, synthetic  # noqa: E402
from rl4mm_b200.features import Portfolio  # noqa: E402
from rl4mm_b200.gym import HistoricalOrderbookEnvironment  # noqa: E402
from rl4mm_b200.ppo import PPOConfig, PPOTrainer  # noqa: E402
from rl4mm_b200.rewards import InventoryAdjustedPnL  # noqa: E402
from rl4mm_b200.simulation import DeviceDatabase  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--rollout-steps", type=int, default=128)
    ap.add_argument("--episode-seconds", type=float, default=60.0)
    ap.add_argument("--n-msgs", type=int, default=2_000_000)
    ap.add_argument("--duration-s", type=int, default=4680)
    ap.add_argument("--cuda-graph", choices=["auto", "on", "off"], default="auto",
                    help="capture one collection step (policy forward + lobsim_step) in a CUDA graph; auto: n_envs <= 16384")
    args = ap.parse_args(argv)

    rank, world, local_rank = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    day = datetime(2019, 1, 2)
    db = DeviceDatabase()
    db.add_stream("SPY", day, synthetic.generate(synthetic.spy_day(seed=0, n_msgs=args.n_msgs, duration_s=args.duration_s)))
    step, ep = timedelta(seconds=0.1), timedelta(seconds=args.episode_seconds)
    feats = HistoricalOrderbookEnvironment.get_default_features(step, ep)
    warm = max(f.window_size for f in feats)
    env = HistoricalOrderbookEnvironment(
        features=feats, ticker="SPY", step_size=step, episode_length=ep, min_date=day, max_date=day,
        initial_portfolio=Portfolio(inventory=0, cash=1_000_000), n_levels=10, database=db, n_envs=args.envs, device=local_rank,
        min_start_timedelta=timedelta(hours=9, minutes=30) + warm + timedelta(seconds=10),
        max_end_timedelta=timedelta(hours=9, minutes=30, seconds=args.duration_s - 10),
        per_step_reward_function=InventoryAdjustedPnL(inventory_aversion=1e-4), terminal_reward_function=InventoryAdjustedPnL(inventory_aversion=0.1),
        max_levels_per_side=64, max_orders_per_side=256, max_agent_orders=64, portfolio_carryover=False, seed=1234 + rank)
    trainer = PPOTrainer(env, PPOConfig(rollout_steps=args.rollout_steps, reward_scale=1e-3), seed=rank,
                         use_cuda_graph={"auto": None, "on": True, "off": False}[args.cuda_graph])
    hist = trainer.train(args.iterations, log=(lambda s: print(json.dumps(s))) if rank == 0 else None)
    if world > 1:
        # DDP keeps the replicas identical: the parameter checksum must be the same number on every rank
        chk = torch.stack([p.detach().double().sum() for p in trainer.policy.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"world_size": world, "param_checksum_min": float(lo), "param_checksum_max": float(hi),
                              "replicas_in_sync": bool(lo == hi)}))
        assert bool(lo == hi), "DDP replicas diverged"
        torch.distributed.destroy_process_group()
    return hist


if __name__ == "__main__":
    main()
