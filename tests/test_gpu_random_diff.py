"""Randomised differential test: the CUDA path vs the CPU oracle on seeded random streams x random env configurations
(tests/random_cases.py): quote-level ranges, concentration mode, enter_spread, clearing market orders, every feature kind
with random windows and z-score normalisation, every reward kind, resync on/off, thin books that run empty, non-zero
initial inventory, portfolio carry-over across a mid-run reset.  Integer state bit-exact, floats within 1e-6 relative,
error bits identical."""
import numpy as np
import pytest

import parity_helpers as H
import random_cases as R
from rl4mm_b200 import abi

pytestmark = pytest.mark.gpu

N_CASES = 64


@pytest.mark.parametrize("seed", range(N_CASES))
def test_random_env_case(seed):
    import torch

    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic
    from test_gpu_parity import compare_books, make_sim

    c = R.random_case(seed)
    s = synthetic.generate(c["synth"])
    n, ep = c["n_envs"], c["cfg_kw"]["episode_steps"]
    sim = make_sim(abi.default_cfg(n_envs=n, **c["cfg_kw"]), [s])
    oracles = [Oracle(abi.default_cfg(n_envs=1, **c["cfg_kw"]), s) for _ in range(n)]
    agent = abi.Agent(kind=abi.AGENT_EXTERNAL)
    for part in range(2):
        starts = (c["starts"] + part * 10).astype(np.int32)
        obs0 = sim.reset(0, starts).cpu().numpy()
        acts = c["actions"][part * ep:(part + 1) * ep]
        # first half of the episode as single steps (lobsim_step), the rest as one fused rollout
        k = ep // 2
        outs = [sim.step(torch.tensor(acts[t], device="cuda")) for t in range(k)]
        obs_a = np.stack([o[0].cpu().numpy() for o in outs]); rew_a = np.stack([o[1].cpu().numpy() for o in outs])
        done_a = np.stack([o[2].cpu().numpy() for o in outs])
        obs_b, _, rew_b, done_b, info_b = (x.cpu().numpy() for x in sim.rollout(ep - k, agent, torch.tensor(acts[k:], device="cuda"), want_info=True))
        obs, rew, done = np.concatenate([obs_a, obs_b]), np.concatenate([rew_a, rew_b]), np.concatenate([done_a, done_b])
        st = sim.state()
        for env, o in enumerate(oracles):
            what = f"seed {seed} part {part} env {env}"
            H.assert_close_vec(obs0[env], o.reset(int(starts[env])), what + " reset obs")
            oo, _, orw, od, oi = o.rollout(ep, agent, acts[:, env], want_info=True)
            os_ = o.state()
            assert int(st["err"][env]) == int(os_["err"]), (what, int(st["err"][env]), int(os_["err"]))
            for t in range(ep):
                H.assert_close_vec(obs[t, env], oo[t], f"{what} t {t} obs")
                assert H.close(rew[t, env], orw[t]), (what, t, rew[t, env], orw[t])
                assert done[t, env] == od[t], (what, t)
            for t in range(ep - k):
                H.assert_close_vec(info_b[t, env], oi[k + t], f"{what} t {k + t} info")
            compare_books(sim, env, o, what)
            assert st["inventory"][env] == os_["inventory"] and H.close(st["cash"][env], os_["cash"]), what
            assert H.close(st["price"][env], os_["price"]), what
            for f in ("now_step", "min_buy_price", "max_sell_price", "best_buy", "best_sell", "best_buy_volume", "best_sell_volume"):
                assert st[f][env] == os_[f], (what, f, st[f][env], os_[f])
    sim.close()
