#!/usr/bin/env python
"""BASELINE.json configs[3]: PPO-style rollout collection -- envs sharded over the GPUs of one box (env i lives on
GPU i mod world), a fused T-step rollout per shard, then ONE NCCL all-gather of the per-episode statistics.

    python examples/collect_rollouts.py --envs 65536                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/collect_rollouts.py --envs 1048576

There is no data-path collective: the books never talk to each other.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl4mm_b200 import abi, parallel, synthetic  # noqa: E402
from rl4mm_b200.agents import Teradactyl  # noqa: E402
from rl4mm_b200.device import LobSim  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536, help="total number of envs over all GPUs")
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--rollouts", type=int, default=2)
    ap.add_argument("--n-msgs", type=int, default=2_000_000)
    ap.add_argument("--duration-s", type=int, default=4680)
    args = ap.parse_args(argv)

    rank, world, local_rank = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    ids = parallel.shard_env_ids(args.envs, rank, world)
    n_local = -(-args.envs // world)                      # equal shard sizes (the last ranks pad with replicas)
    ids = np.resize(ids, n_local)

    stream = synthetic.generate(synthetic.spy_day(seed=0, n_msgs=args.n_msgs, duration_s=args.duration_s))
    feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_BOOK_IMBALANCE, 0, 100000, -1, 1),
             abi.feature(abi.FEAT_PRICE_MOVE, 10, 100000, -1e4, 1e4), abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6),
             abi.feature(abi.FEAT_VOLATILITY, 100, 100000, 0, 1)]
    cfg = abi.default_cfg(n_envs=n_local, n_levels=stream.n_levels, episode_steps=args.T, warmup_steps=100, features=feats,
                          step_reward=abi.Reward(abi.REWARD_INV_ADJ_PNL, 0, 1e-4), terminal_reward=abi.Reward(abi.REWARD_INV_ADJ_PNL, 0, 0.1),
                          max_levels_per_side=64, max_orders_per_side=256, max_agent_orders=64, portfolio_carryover=0)
    sim = LobSim(cfg, local_rank)
    sim.load_stream(0, stream)
    agent = Teradactyl(max_inventory=500, default_kappa=8.0, default_omega=0.45, max_kappa=12.0, inventory_index=3).to_abi()
    sps, last_start = stream.steps_per_second, stream.n_seconds - (args.T + 100) // stream.steps_per_second - 2
    out = None
    for r in range(args.rollouts):
        # episode start of GLOBAL env i: a hash of (i, rollout) on whole seconds, so the result is independent of `world`
        starts = ((11 + (ids * 2654435761 + r * 40503) % (last_start - 11)) * sps).astype(np.int32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.reset(0, starts)
        obs, act, rew, done = sim.rollout(args.T, agent)
        stats = parallel.episode_stats(rew, done, sim.state(), obs[:, :, 0])
        full = parallel.gather_episode_stats(stats, args.envs)        # the only collective
        torch.cuda.synchronize()
        dt = parallel.max_over_ranks(time.perf_counter() - t0, torch.device("cuda", local_rank))
        if rank == 0:
            col = {k: i for i, k in enumerate(parallel.STAT_FIELDS)}
            out = dict(rollout=r, envs=args.envs, world=world, T=args.T, seconds=dt, env_steps_per_sec=args.envs * args.T / dt,
                       mean_return=float(full[:, col["return"]].mean()), mean_abs_inventory=float(full[:, col["final_inventory"]].abs().mean()),
                       errors=int((full[:, col["err"]] != 0).sum()), gathered_shape=list(full.shape))
            print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
