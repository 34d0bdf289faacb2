"""Seeded random (stream, env config, action) cases for the differential tests: every knob of lobsim_cfg_t that changes
behaviour is drawn at random, so that the CUDA path and the oracle are compared off the beaten track of the goldens."""
import numpy as np

from rl4mm_b200 import abi, synthetic


def random_features(rng, step_us=100_000):
    kinds = [
        lambda: abi.feature(abi.FEAT_SPREAD, 0, step_us, 0, 5000),
        lambda: abi.feature(abi.FEAT_BOOK_IMBALANCE, 0, step_us, -1, 1),
        lambda: abi.feature(abi.FEAT_PRICE_MOVE, int(rng.integers(1, 12)), step_us * int(rng.choice([1, 2, 10])), -1e4, 1e4),
        lambda: abi.feature(abi.FEAT_PRICE_RANGE, int(rng.integers(1, 8)), step_us * int(rng.choice([1, 5])), 0, 1e4),
        lambda: abi.feature(abi.FEAT_VOLATILITY, int(rng.integers(2, 40)), step_us, 0, float(rng.choice([1.0, 1e-9]))),
        lambda: abi.feature(abi.FEAT_PRICE, 0, step_us * int(rng.choice([1, 10])), 0, 1e8),
        lambda: abi.feature(abi.FEAT_TRADE_DIR_IMBALANCE, int(rng.integers(1, 30)), step_us, -1, 1, iparam=int(rng.integers(0, 2))),
        lambda: abi.feature(abi.FEAT_TRADE_VOL_IMBALANCE, int(rng.integers(1, 30)), step_us, -1, 1, iparam=int(rng.integers(0, 2))),
        lambda: abi.feature(abi.FEAT_INVENTORY, 0, step_us, -float(rng.choice([50, 1e6])), float(rng.choice([50, 1e6]))),
        lambda: abi.feature(abi.FEAT_EPISODE_PROPORTION, 0, step_us, 0, 1, dparam=1.0 / 77),
        lambda: abi.feature(abi.FEAT_TIME_OF_DAY, 0, step_us * 10, 0, 9, iparam=10),
        lambda: abi.amihud(int(rng.integers(2, 6)), int(rng.integers(1, 4)), step_us, 0, 1e-3),
    ]
    n = int(rng.integers(3, 13))
    feats = [kinds[i]() for i in rng.permutation(len(kinds))[:n]]
    for f in feats:                                   # rolling z-score on a few of them, short histories => eviction
        if rng.random() < 0.25 and f.kind not in (abi.FEAT_EPISODE_PROPORTION,):
            f.norm_len = int(rng.integers(3, 40))
    return feats


def random_reward(rng):
    k = int(rng.integers(0, 4))
    if k == 0:
        return abi.Reward(abi.REWARD_PNL, 0, 0.0)
    if k == 3:
        mx = int(rng.integers(4, 30))
        return abi.rolling_sharpe(mx, int(rng.integers(2, mx + 1)))
    return abi.Reward(abi.REWARD_INV_ADJ_PNL, int(k == 2), float(rng.choice([1e-4, 0.01, 0.5])))


def random_case(seed: int):
    rng = np.random.default_rng(1000 + seed)
    n_levels = int(rng.choice([5, 10, 50]))
    thin = rng.random() < 0.3
    sc = synthetic.SynthConfig(
        seed=seed, n_msgs=int(rng.integers(40_000, 90_000)), duration_s=int(rng.integers(120, 260)), n_levels=n_levels,
        mid0=int(rng.choice([300_000, 1_000_000, 4_000_000])), p_limit=0.44, p_cancel=float(rng.choice([0.02, 0.2])),
        p_delete=0.0, p_exec=float(rng.choice([0.05, 0.12, 0.2])), geom_p=float(rng.choice([0.12, 0.35, 0.6])),
        init_levels=int(rng.integers(n_levels + 2, 60)), mean_queue=int(rng.choice([1, 4, 10])),
        target_orders=int(rng.choice([12, 25]) if thin else rng.choice([50, 200, 500])), max_offset_ticks=int(rng.choice([8, 40, 70])),
        size_sigma=float(rng.choice([0.3, 0.8, 1.4])), p_sweep=float(rng.choice([0.0005, 0.01, 0.03])))
    sc.p_delete = 1.0 - sc.p_limit - sc.p_cancel - sc.p_exec
    feats = random_features(rng)
    warm = max(int(f.lookback * (f.update_us // 100_000)) for f in feats)
    warm = (warm + 9) // 10 * 10 + 10 * int(rng.integers(0, 2))       # the book reset (start - warm-up) must fall on a whole second
    conc = float(rng.choice([-1.0, 12.0]))
    minq = int(rng.choice([0, 0, 2]))
    maxq = minq + int(rng.choice([3, 5, 10]))
    clearing = int(rng.random() < 0.4)
    cfg_kw = dict(
        n_levels=n_levels, episode_steps=int(rng.integers(20, 120)), warmup_steps=warm, min_quote_level=minq, max_quote_level=maxq,
        outer_levels=int(rng.choice([2, 20])) * n_levels // 50 if n_levels == 50 else int(rng.integers(1, n_levels)),
        resync=int(rng.random() < 0.8), active_volume=int(rng.choice([10, 100, 1000])), market_order_clearing=clearing,
        market_order_fraction_of_inventory=float(rng.choice([0.25, 1.0])) if clearing else 0.0,
        enter_spread=int(rng.random() < 0.4), inc_prev_action_in_obs=int(rng.random() < 0.4),
        portfolio_carryover=int(rng.random() < 0.6), concentration=conc, initial_cash=float(rng.choice([1e9, 1e12])),
        initial_inventory=int(rng.choice([0, 0, 40, -300])), features=feats, step_reward=random_reward(rng),
        terminal_reward=random_reward(rng), max_levels_per_side=128, max_orders_per_side=1024, max_agent_orders=64,
        fill_log_capacity=int(rng.choice([0, 4096])))
    n_envs = int(rng.integers(2, 6))
    last = sc.duration_s * 10 - 2 * cfg_kw["episode_steps"] - 20
    starts = np.sort(((warm + 10) // 10 + 1 + rng.integers(0, max(last // 10 - (warm + 10) // 10 - 1, 1), size=n_envs)) * 10).astype(np.int32)
    ad = (2 if conc >= 0 else 4) + clearing
    hi = np.array(([12.0] * 2 if conc >= 0 else [10.0] * 4) + ([200.0] if clearing else []))
    T = 2 * cfg_kw["episode_steps"]                   # two episodes: reset in between (portfolio carry-over or not)
    acts = rng.uniform(0.0, 1.0, size=(T, n_envs, ad)) * hi
    acts[rng.random(acts.shape) < 0.03] = 0.0        # the reference's a + 1e-6 corner
    # agent: separate generator so that the (stream, config, action) draws above do not depend on it
    rng2 = np.random.default_rng(5000 + seed)
    kind = str(rng2.choice(["external", "external", "fixed", "teradactyl"]))
    if kind == "teradactyl" and conc >= 0:
        kind = "fixed"                                # Teradactyl emits (alpha, beta) x 2: needs the 4(+1)-dimensional action
    agent = abi.Agent(kind=abi.AGENT_EXTERNAL)
    if kind == "fixed":
        import ctypes

        fa = list(rng2.uniform(0.0, 1.0, size=ad) * hi) + [0.0] * (5 - ad)
        agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(*fa))
    elif kind == "teradactyl":
        idx = [i for i, f in enumerate(feats) if f.kind == abi.FEAT_INVENTORY]
        if not idx:
            feats.append(abi.feature(abi.FEAT_INVENTORY, 0, 100_000, -1e6, 1e6))
            idx = [len(feats) - 1]
        agent = abi.Agent(kind=abi.AGENT_TERADACTYL, inventory_index=idx[0], max_inventory=float(rng2.choice([50.0, 300.0, 5000.0])),
                          default_kappa=float(rng2.uniform(3.0, 10.0)), default_omega=float(rng2.uniform(0.2, 0.8)),
                          max_kappa=float(rng2.uniform(10.0, 14.0)), exponent=float(rng2.choice([1.0, 1.5, 2.0])), market_clearing=clearing)
    # capacities: two of the three choices have a compiled StaticLayout (the straight-line k_env_fast), the third runs the
    # general runtime-layout kernel; small capacities that overflow are flagged by the device and skipped by the test
    rng3 = np.random.default_rng(7000 + seed)
    caps = [(64, 256), (128, 512), (128, 1024)] if sc.target_orders <= 100 else [(128, 512), (128, 1024)]
    cfg_kw["max_levels_per_side"], cfg_kw["max_orders_per_side"] = caps[int(rng3.integers(0, len(caps)))]
    return dict(synth=sc, cfg_kw=cfg_kw, n_envs=n_envs, starts=starts, actions=acts, T=T, agent=agent, agent_kind=kind)


def random_replay_case(seed: int):
    """Replay-only case: a random stream shape (thin / deep books, many sweeps) and random resync settings."""
    rng = np.random.default_rng(9000 + seed)
    n_levels = int(rng.choice([5, 10, 50]))
    sc = synthetic.SynthConfig(
        seed=100 + seed, n_msgs=int(rng.integers(60_000, 150_000)), duration_s=int(rng.integers(100, 300)), n_levels=n_levels,
        mid0=int(rng.choice([300_000, 2_000_000, 5_000_000])), p_limit=float(rng.choice([0.35, 0.48])), p_cancel=float(rng.choice([0.02, 0.25])),
        p_delete=0.0, p_exec=float(rng.choice([0.05, 0.1, 0.2])), geom_p=float(rng.choice([0.1, 0.35, 0.6])),
        init_levels=int(rng.integers(n_levels + 2, 60)), mean_queue=int(rng.choice([1, 4, 12])),
        target_orders=int(rng.choice([10, 30, 100, 600])), max_offset_ticks=int(rng.choice([8, 40, 70])),
        size_sigma=float(rng.choice([0.3, 0.8, 1.4])), p_sweep=float(rng.choice([0.0005, 0.01, 0.05])))
    sc.p_delete = 1.0 - sc.p_limit - sc.p_cancel - sc.p_exec
    cfg_kw = dict(n_levels=n_levels, outer_levels=int(rng.integers(1, n_levels)), resync=int(rng.random() < 0.8),
                  max_levels_per_side=int(rng.choice([64, 128])), max_orders_per_side=int(rng.choice([256, 512, 1536])))
    n_envs = int(rng.integers(3, 9))
    starts = (np.sort(rng.integers(0, sc.duration_s - 60, size=n_envs)) * 10).astype(np.int32)
    chunks = [int(x) for x in rng.choice([1, 3, 10, 57, 200], size=5)]
    return dict(synth=sc, cfg_kw=cfg_kw, n_envs=n_envs, starts=starts, chunks=chunks)
