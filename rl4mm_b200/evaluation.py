"""Evaluation / episode-summary path of the reference (rl4mm/gym/utils.py:100-257, rl4mm/rewards/RewardFunctions.py:10-22),
batched: ONE fused device rollout produces every trajectory (one env = one "iteration" of the reference's loop), the
per-step info series come straight from the kernel (``lobsim_rollout_info``), and the summary dict is reduced from those
tensors.  Keys, per-episode ordering and the reference's quirks (mean action drops its LAST component, utils.py:161;
Sharpe with ddof=1 and +float_min, RewardFunctions.py:20) are kept so that the JSON written by
``save_episode_summary_json`` is interchangeable with the reference's (utils.py:417-420).
"""
from __future__ import annotations

import json
import sys
from typing import Dict

import numpy as np

from . import abi

SUMMARY_KEYS = ("equity_curves", "reward_series", "rewards", "actions", "spread", "inventory", "inventories", "asset_prices",
                "agent_midprice_offsets")


def get_sharpe(aum_array) -> np.ndarray:
    """rl4mm/rewards/RewardFunctions.py:10-22 over the LAST axis (any leading batch axes)."""
    aum = np.asarray(aum_array, dtype=np.float64)
    if np.min(aum) <= 0:
        raise Exception("AUM has gone non_positive")
    simple_returns = np.exp(np.diff(np.log(aum), axis=-1)) - 1
    return np.mean(simple_returns, axis=-1) / (np.std(simple_returns, axis=-1, ddof=1) + sys.float_info.min)


def init_episode_summary_dict() -> Dict:
    return {k: [] for k in SUMMARY_KEYS}      # utils.py:131-143


def weighted_midprice_offsets(actions: np.ndarray, order_distributor) -> np.ndarray:
    """info["weighted_midprice_offset"] (InfoCalculators.py:44-50) for a batch of actions [..., A]: half the difference
    of the volume-weighted mean quote level of the sell and the buy ladder."""
    a = np.asarray(actions, dtype=np.float64)
    flat = a.reshape(-1, a.shape[-1])
    orders = order_distributor.convert_action(flat)
    dist = np.arange(order_distributor.quote_levels)
    total = order_distributor.active_volume
    return ((orders["sell"] @ dist / total - orders["buy"] @ dist / total) / 2).reshape(a.shape[:-1])


def episode_summary_from_rollout(act, rew, info, order_distributor, esd: Dict = None) -> Dict:
    """append_to_episode_summary_dict (utils.py:146-190) for every env of a rollout: act [T, N, A], rew [T, N],
    info [T, N, abi.INFO_DIM] (numpy or torch).  Appends N episodes, env 0 first."""
    to_np = lambda x: x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)  # noqa: E731
    act, rew, info = to_np(act), to_np(rew), to_np(info)
    esd = init_episode_summary_dict() if esd is None else esd
    col = {k: i for i, k in enumerate(abi.INFO_FIELDS)}
    offs = weighted_midprice_offsets(act, order_distributor)                       # [T, N]
    for e in range(rew.shape[1]):
        esd["equity_curves"].append(info[:, e, col["aum"]].copy())
        esd["reward_series"].append([float(r) for r in rew[:, e]])
        esd["rewards"].append(np.mean(rew[:, e]))
        esd["actions"].append(np.mean(act[:, e, :], axis=0)[:-1])
        esd["spread"].append(np.mean(info[:, e, col["market_spread"]]))
        inv = info[:, e, col["inventory"]].astype(np.int64)
        esd["inventory"].append(np.mean(inv))
        esd["inventories"].append(inv)
        esd["asset_prices"].append(info[:, e, col["asset_price"]].copy())
        esd["agent_midprice_offsets"].append(offs[:, e].copy())
    return esd


def get_episode_summary_dict(agent, env, n_iterations: int = None) -> Dict:
    """get_episode_summary_dict (utils.py:120-128): `n_iterations` trajectories of `agent`.  `env` is the batched
    HistoricalOrderbookEnvironment; each pass resets all its envs (random starts, as the reference's env.reset does) and
    runs one full episode of every env in a single launch.  The agent must have a device implementation (to_abi)."""
    n_iterations = env.n_envs if n_iterations is None else n_iterations
    desc = agent.to_abi()
    if desc is None:
        raise NotImplementedError("agent has no device implementation")
    esd = init_episode_summary_dict()
    while len(esd["rewards"]) < n_iterations:
        env.reset()
        _, act, rew, done, info = env.sim.rollout(env.n_steps, desc, want_obs=False, want_info=True)
        env._raise_on_errors()
        assert bool(done[-1].all()) and not bool(done[:-1].any()), "episodes end exactly at episode_length"
        episode_summary_from_rollout(act, rew, info, env.order_distributor, esd)
    return {k: v[:n_iterations] for k, v in esd.items()}


class _NumpyEncoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, np.generic):
            return o.item()
        return super().default(o)


def save_episode_summary_json(esd: Dict, path) -> None:
    """utils.py:417-420 (json.dump(episode_summary_dict, cls=NumpyEncoder))."""
    with open(path, "w") as f:
        json.dump(esd, f, cls=_NumpyEncoder)
