"""Multi-GPU plumbing: environments shard by rank, streams are replicated, and the only collective is the all-gather of
per-episode statistics.  One process per GPU (torchrun); backend NCCL on GPUs, gloo in the CPU
tests."""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

STAT_FIELDS = ("return", "length", "final_inventory", "final_cash", "aum", "mean_spread", "n_done", "err")


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank


def shard_env_ids(n_envs_total: int, rank: int, world: int) -> np.ndarray:
    """Global env ids owned by `rank`: env i lives on GPU i mod world."""
    return np.arange(rank, n_envs_total, world, dtype=np.int64)


def episode_stats(rew: torch.Tensor, done: torch.Tensor, state: np.ndarray, obs_spread: torch.Tensor = None) -> torch.Tensor:
    """[N_local, 8] f32 statistics of a rollout: rew/done are [T, N]; state is LobSim.state()."""
    T, N = rew.shape
    dev = rew.device
    out = torch.zeros((N, len(STAT_FIELDS)), dtype=torch.float32, device=dev)
    out[:, 0] = rew.sum(dim=0).to(torch.float32)
    out[:, 1] = float(T)
    out[:, 2] = torch.as_tensor(state["inventory"].astype(np.float32), device=dev)
    out[:, 3] = torch.as_tensor(state["cash"].astype(np.float32), device=dev)
    out[:, 4] = torch.as_tensor((state["cash"] + state["price"] * state["inventory"]).astype(np.float32), device=dev)
    if obs_spread is not None:
        out[:, 5] = obs_spread.to(torch.float32).mean(dim=0)
    out[:, 6] = done.to(torch.float32).sum(dim=0)
    out[:, 7] = torch.as_tensor(state["err"].astype(np.float32), device=dev)
    return out


def episode_stats_dev(rew: torch.Tensor, done: torch.Tensor, state_dev: torch.Tensor, obs_spread: torch.Tensor = None) -> torch.Tensor:
    """`episode_stats` without leaving the device: ``state_dev`` is ``LobSim.state_dev()`` ([N, 80] uint8 =
    lobsim_env_state_t records); no host synchronisation, so it can sit between a rollout and the all-gather."""
    T, N = rew.shape
    i64, f64, i32 = state_dev.view(torch.int64), state_dev.view(torch.float64), state_dev.view(torch.int32)
    inv, cash, price, err = i64[:, 0].to(torch.float64), f64[:, 1], f64[:, 2], i32[:, 14]
    out = torch.zeros((N, len(STAT_FIELDS)), dtype=torch.float32, device=rew.device)
    out[:, 0] = rew.sum(dim=0).to(torch.float32)
    out[:, 1] = float(T)
    out[:, 2] = inv.to(torch.float32)
    out[:, 3] = cash.to(torch.float32)
    out[:, 4] = (cash + price * inv).to(torch.float32)
    if obs_spread is not None:
        out[:, 5] = obs_spread.to(torch.float32).mean(dim=0)
    out[:, 6] = done.to(torch.float32).sum(dim=0)
    out[:, 7] = err.to(torch.float32)
    return out


def gather_episode_stats(stats_local: torch.Tensor, n_envs_total: int = None) -> torch.Tensor:
    """All-gather [N_local, K] -> [N_total, K] in GLOBAL env order (env i is row i).  Requires equal N_local on all
    ranks (pad the last shard) -- one `all_gather_into_tensor` per rollout, latency-bound (32 B per env)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return stats_local
    world = dist.get_world_size()
    n_local, k = stats_local.shape
    flat = torch.empty((world * n_local, k), dtype=stats_local.dtype, device=stats_local.device)
    dist.all_gather_into_tensor(flat, stats_local.contiguous())
    # rank r holds global envs r, r + world, ... => interleave back
    out = flat.view(world, n_local, k).transpose(0, 1).reshape(world * n_local, k)
    return out if n_envs_total is None else out[:n_envs_total]


def max_over_ranks(seconds: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
