"""Rule-based-agent tuning as a batched parameter sweep (the reference runs one Ray Tune trial per
Teradactyl parameter set, tune_rule_based_agents.py:24-72, each trial stepping its own env and reporting
``episode_reward_mean``).  Here K parameter sets x M episodes are the envs of ONE fused rollout
(``lobsim_rollout_agents``: one built-in agent per env); the search strategy on top (grid, random, Bayesian) only sees
the returned table."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import abi, evaluation
from .agents import Teradactyl


def sweep_agents(env, agents: Sequence, episodes_per_agent: int = None) -> List[Dict]:
    """Evaluate each agent (objects with ``to_abi()``: Teradactyl, FixedActionAgent) on ``episodes_per_agent`` episodes in one
    launch.  ``env.n_envs`` must equal ``len(agents) * episodes_per_agent``; env ``k * M + j`` runs agent k, episode j.
    Episode j of every agent starts at the same (random) time, so the agents are compared on identical market data."""
    K = len(agents)
    M = env.n_envs // K if episodes_per_agent is None else episodes_per_agent
    assert K * M == env.n_envs, f"n_envs ({env.n_envs}) must be len(agents) * episodes_per_agent ({K} x {M})"
    env.reset()
    # identical episode starts across agents: re-reset env k*M + j to the start of env j
    sids, starts = env.stream_ids.reshape(K, M).copy(), env.episode_start_steps.reshape(K, M).copy()
    sids[:], starts[:] = sids[0], starts[0]
    env.stream_ids[:], env.episode_start_steps[:] = sids.ravel(), starts.ravel()
    env.sim.reset(env.stream_ids, env.episode_start_steps)
    descs = [a.to_abi() for a in agents for _ in range(M)]
    _, act, rew, done, info = env.sim.rollout(env.n_steps, descs, want_obs=False, want_info=True)
    env._raise_on_errors()
    rew, info = rew.cpu().numpy(), info.cpu().numpy()
    col = {k: i for i, k in enumerate(abi.INFO_FIELDS)}
    ret = rew.sum(axis=0).reshape(K, M)
    aum = np.moveaxis(info[:, :, col["aum"]], 0, -1).reshape(K, M, -1)
    inv = np.abs(info[:, :, col["inventory"]]).mean(axis=0).reshape(K, M)
    out = []
    for k, a in enumerate(agents):
        positive = aum[k].min() > 0
        out.append(dict(agent=a.get_name(), episode_reward_mean=float(ret[k].mean()), episode_reward_std=float(ret[k].std()),
                        mean_abs_inventory=float(inv[k].mean()),
                        sharpe_mean=float(np.mean(evaluation.get_sharpe(aum[k]))) if positive else float("nan"),
                        episodes=M))
    return out


def teradactyl_grid(max_inventory, default_kappas, default_omegas, max_kappas, exponents=(1.0,), inventory_index: int = 3) -> List[Teradactyl]:
    return [Teradactyl(max_inventory=max_inventory, default_kappa=k, default_omega=w, max_kappa=mk, exponent=e, inventory_index=inventory_index)
            for k in default_kappas for w in default_omegas for mk in max_kappas for e in exponents]
