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
    agent, external = c["agent"], c["agent_kind"] == "external"
    for part in range(2):
        starts = (c["starts"] + part * 10).astype(np.int32)
        obs0 = sim.reset(0, starts).cpu().numpy()
        acts = c["actions"][part * ep:(part + 1) * ep]
        # first half of the episode as single steps (lobsim_step), the rest as one fused rollout
        # (a built-in agent: two fused rollouts, its actions are compared as well)
        k = ep // 2
        if external:
            outs = [sim.step(torch.tensor(acts[t], device="cuda")) for t in range(k)]
            obs_a = np.stack([o[0].cpu().numpy() for o in outs]); rew_a = np.stack([o[1].cpu().numpy() for o in outs])
            done_a = np.stack([o[2].cpu().numpy() for o in outs])
            act_a = acts[:k]
        else:
            obs_a, act_a, rew_a, done_a = (x.cpu().numpy() for x in sim.rollout(k, agent))
        obs_b, act_b, rew_b, done_b, info_b = (x.cpu().numpy() for x in sim.rollout(
            ep - k, agent, torch.tensor(acts[k:], device="cuda") if external else None, want_info=True))
        obs, rew, done = np.concatenate([obs_a, obs_b]), np.concatenate([rew_a, rew_b]), np.concatenate([done_a, done_b])
        act = np.concatenate([act_a, act_b])
        st = sim.state()
        for env, o in enumerate(oracles):
            what = f"seed {seed} part {part} env {env}"
            H.assert_close_vec(obs0[env], o.reset(int(starts[env])), what + " reset obs")
            oo, oa, orw, od, oi = o.rollout(ep, agent, acts[:, env] if external else None, want_info=True)
            os_ = o.state()
            assert int(st["err"][env]) == int(os_["err"]), (what, int(st["err"][env]), int(os_["err"]))
            for t in range(ep):
                H.assert_close_vec(act[t, env], oa[t], f"{what} t {t} action")
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


@pytest.mark.parametrize("seed", range(24))
def test_random_replay_case(seed, monkeypatch):
    """Pure replay (both kernel families: the straight-line k_replay_fast and, forced, the general k_advance) vs the oracle."""
    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic
    from test_gpu_parity import compare_books, make_sim

    c = R.random_replay_case(seed)
    s = synthetic.generate(c["synth"])
    n = c["n_envs"]
    okw = {k: v for k, v in c["cfg_kw"].items() if not k.startswith("max_")}
    for force_general in ("0", "1"):
        monkeypatch.setenv("LOBSIM_FORCE_GENERAL", force_general)
        sim = make_sim(abi.default_cfg(n_envs=n, **c["cfg_kw"]), [s])
        oracles = [Oracle(abi.default_cfg(n_envs=1, **okw), s) for _ in range(n)]
        sim.reset_book(0, c["starts"])
        for o, st in zip(oracles, c["starts"]):
            o.reset_book(int(st))
        for chunk in c["chunks"]:
            sim.replay(chunk)
            st = sim.state()
            for env, o in enumerate(oracles):
                what = f"replay seed {seed} general={force_general} env {env} chunk {chunk}"
                o.replay(chunk)
                os_ = o.state()
                overflow = int(st["err"][env]) & (abi.ERR_LEVEL_OVERFLOW | abi.ERR_ORDER_OVERFLOW)
                if overflow:          # fixed capacities are a device-side limit (the oracle is unbounded): flagged, not compared
                    continue
                assert int(st["err"][env]) == int(os_["err"]), (what, int(st["err"][env]), int(os_["err"]))
                for f in ("now_step", "min_buy_price", "max_sell_price", "best_buy", "best_sell", "best_buy_volume", "best_sell_volume"):
                    assert st[f][env] == os_[f], (what, f, st[f][env], os_[f])
                compare_books(sim, env, o, what)
        sim.close()
