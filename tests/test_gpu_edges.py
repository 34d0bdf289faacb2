"""Edge cases the reference's tests touch only implicitly: empty / ragged inputs, capacity limits, error flags,
call-order errors, env subsets.  All through the C ABI."""
import ctypes

import numpy as np
import pytest

import parity_helpers as H
from rl4mm_b200 import abi

pytestmark = pytest.mark.gpu


def _sim(cfg, streams=()):
    import torch

    assert torch.cuda.is_available()
    from rl4mm_b200.device import LobSim

    sim = LobSim(cfg, 0)
    for i, s in enumerate(streams):
        sim.load_stream(i, s)
    return sim


def _order(env, type, direction, price, volume, is_external, ref):
    o = np.zeros(1, abi.ORDER_DTYPE)
    o[0] = (env, type, direction, price, volume, int(is_external), ref, 0)
    return o


@pytest.mark.replay_path
def test_message_free_steps_and_zero_step_calls():
    """A stream with no messages at all: replay / rollout advance the clock and leave the snapshot book untouched."""
    s = H.snapshot_stream([[0, 1000, 5], [0, 900, 7], [1, 1100, 3]], n_levels=50)
    sim = _sim(abi.default_cfg(n_envs=5, features=[abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000)], episode_steps=5,
                               portfolio_carryover=0), [s])
    sim.reset_book(0, 0)
    before = [sim.dump_book(3, side).copy() for side in (0, 1)]
    sim.replay(0)
    sim.replay(7)
    st = sim.state()
    assert np.all(st["now_step"] == 7) and np.all(st["err"] == 0)
    for side in (0, 1):
        assert np.array_equal(sim.dump_book(3, side), before[side])
    obs = sim.reset(0, 0)
    assert obs.shape == (5, 1) and float(obs[0, 0]) == 100.0
    o, a, r, d = sim.rollout(0, abi.Agent(kind=abi.AGENT_NONE))
    assert o.shape == (0, 5, 1)


@pytest.mark.replay_path
def test_stepping_past_the_grid_sets_end_of_stream():
    s = H.snapshot_stream([[0, 1000, 5], [1, 1100, 3]], n_levels=50)
    sim = _sim(abi.default_cfg(n_envs=2), [s])
    sim.reset_book(0, 0)
    sim.replay(10)
    assert np.all(sim.state()["err"] == 0)
    sim.replay(1)
    assert np.all(sim.state()["err"] & abi.ERR_END_OF_STREAM)


def test_reset_without_snapshot_and_off_second():
    s = H.load_fixture_stream("reference")   # second 0 (35999 s) has no data before it
    sim = _sim(abi.default_cfg(n_envs=3), [s])
    sim.reset_book(0, np.array([0, 10, 13], np.int32))   # no snapshot / ok / not on a whole second
    err = sim.state()["err"]
    assert err[0] & abi.ERR_NO_SNAPSHOT and err[1] == 0 and err[2] & abi.ERR_NO_SNAPSHOT


def test_market_order_on_empty_side_raises_empty_book_flag():
    s = H.snapshot_stream([[0, 1000, 5]], n_levels=50)     # no sell side at all
    sim = _sim(abi.default_cfg(n_envs=2), [s])
    sim.reset_book(0, 0)
    fills, _ = sim.process_orders(_order(1, abi.MSG_MARKET, abi.BUY, 0, 10, True, 7))
    st = sim.state()
    assert st["err"][1] & abi.ERR_EMPTY_BOOK and st["err"][0] == 0 and len(fills) == 0
    # a market sell consumes the whole bid side and then hits the empty book too; the partial fill is reported
    sim.reset_book(0, 0)
    fills, _ = sim.process_orders(_order(0, abi.MSG_MARKET, abi.SELL, 0, 8, True, 9))
    assert [(int(f["volume"]), int(f["price"])) for f in fills] == [(5, 1000)]
    assert sim.state()["err"][0] & abi.ERR_EMPTY_BOOK


def test_capacity_overflow_flags_not_ub():
    s = H.snapshot_stream([[0, 1000, 5], [1, 2000, 3]], n_levels=4)
    cfg = abi.default_cfg(n_envs=1, n_levels=4, max_levels_per_side=4, max_orders_per_side=6, max_agent_orders=2)
    sim = _sim(cfg, [s])
    sim.reset_book(0, 0)
    for k in range(3):   # levels 1000 (snapshot) + 3 new = 4 = capacity
        sim.process_orders(_order(0, abi.MSG_LIMIT, abi.BUY, 990 - 10 * k, 1, True, 10 + k))
    assert sim.state()["err"][0] == 0
    sim.process_orders(_order(0, abi.MSG_LIMIT, abi.BUY, 900, 1, True, 20))
    assert sim.state()["err"][0] & abi.ERR_LEVEL_OVERFLOW and len(sim.dump_book(0, 0)) == 4
    sim.reset_book(0, 0)
    for k in range(5):   # 1 aggregate + 5 orders = 6 = capacity
        sim.process_orders(_order(0, abi.MSG_LIMIT, abi.BUY, 1000, 1, True, 30 + k))
    assert sim.state()["err"][0] == 0
    sim.process_orders(_order(0, abi.MSG_LIMIT, abi.BUY, 1000, 1, True, 40))
    assert sim.state()["err"][0] & abi.ERR_ORDER_OVERFLOW and len(sim.dump_book(0, 0)) == 6
    sim.reset_book(0, 0)
    for k in range(3):
        sim.process_orders(_order(0, abi.MSG_LIMIT, abi.SELL, 2100 + k, 1, False, 0))
    st = sim.state()
    assert st["err"][0] & abi.ERR_AGENT_OVERFLOW and st["n_agent_orders"][0][1] == 2


def test_call_order_and_argument_errors():
    from rl4mm_b200._lib import LobsimError
    import torch

    s = H.snapshot_stream([[0, 1000, 5], [1, 1100, 3]], n_levels=50)
    from rl4mm_b200.device import LobSim

    sim = LobSim(abi.default_cfg(n_envs=2), 0)
    with pytest.raises(LobsimError):
        sim.replay(1)                      # no stream loaded
    sim.load_stream(0, s)
    with pytest.raises(LobsimError):
        sim.step(torch.zeros((2, 4), dtype=torch.float64, device="cuda"))   # step before reset
    with pytest.raises(LobsimError):
        LobSim(abi.default_cfg(n_envs=2, max_quote_level=40), 0)
    deep = LobSim(abi.default_cfg(n_envs=2, max_levels_per_side=4096, max_orders_per_side=60000), 0)   # > smem of an SM: deep-book mode
    assert deep.kernel_path == "deep"
    deep.close()
    with pytest.raises(LobsimError):
        LobSim(abi.default_cfg(n_envs=2, max_levels_per_side=4096, max_orders_per_side=70000), 0)   # level ends are 16-bit


def test_env_subset_reset_and_odd_env_counts():
    """n_envs not a multiple of the CTA size; resetting a subset leaves the other envs untouched."""
    from rl4mm_b200 import synthetic

    s = synthetic.generate(synthetic.spy_day(seed=2, n_msgs=20_000, duration_s=50))
    feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6)]
    n = 37
    cfg = abi.default_cfg(n_envs=n, n_levels=10, features=feats, episode_steps=30, warmup_steps=10, portfolio_carryover=0,
                          max_levels_per_side=64, max_orders_per_side=256, max_agent_orders=64)
    sim = _sim(cfg, [s])
    sim.reset(0, 100)
    agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(1, 1, 1, 1, 0))
    sim.rollout(20, agent)
    before = sim.state().copy()
    ids = [3, 17, 36]
    obs = sim.reset(0, 200, env_ids=ids)
    assert obs.shape == (3, 2)
    after = sim.state()
    others = np.setdiff1d(np.arange(n), ids)
    assert np.array_equal(before[others], after[others])
    assert np.all(after["now_step"][ids] == 200) and np.all(after["inventory"][ids] == 0)
    assert np.all(before["now_step"] == 120)
    # identical envs stay identical, the reset ones are identical among themselves
    o, a, r, d = sim.rollout(5, agent)
    assert bool((o[:, others] == o[:, others[:1]]).all()) and bool((o[:, ids] == o[:, ids[:1]]).all())
    assert np.all(sim.state()["err"] == 0)


def test_deep_books_in_hbm_match_oracle():
    """books deeper than the shared memory of an SM allows (the reference's SortedDict / deque are unbounded,
    rl4mm/orderbook/models.py:66-67).  Capacities of 256 levels x 16 384 orders per side = 265 KB per book put the handle in
    deep-book mode (blob worked on in place in HBM); a synthetic stream with a mean queue of 80 orders per level and ~6 000
    resting orders: replay and a fused agent rollout against the (unbounded) oracle, no ORDER_OVERFLOW."""
    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic
    from test_gpu_parity import compare_books

    sc = synthetic.SynthConfig(seed=11, n_msgs=300_000, duration_s=600, n_levels=50, mid0=2_000_000, p_limit=0.40, p_cancel=0.15,
                               p_delete=0.37, p_exec=0.08, geom_p=0.10, init_levels=70, mean_queue=80, target_orders=6000,
                               max_offset_ticks=80)
    s = synthetic.generate(sc)
    feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_BOOK_IMBALANCE, 0, 100000, -1, 1),
             abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6)]
    kw = dict(n_levels=50, outer_levels=20, features=feats, episode_steps=600, warmup_steps=0)
    starts = [0, 500, 1500]
    sim = _sim(abi.default_cfg(n_envs=len(starts), max_levels_per_side=256, max_orders_per_side=16384, max_agent_orders=64, **kw), [s])
    assert sim.kernel_path == "deep"
    sim.reset_book(0, np.array(starts, np.int32))
    oracles = [Oracle(abi.default_cfg(**kw), s) for _ in starts]
    for o, st in zip(oracles, starts):
        o.reset_book(st)
    for chunk in (1, 99, 1900):
        sim.replay(chunk)
        st = sim.state()
        assert np.all(st["err"] == 0), st["err"]
        for env, o in enumerate(oracles):
            o.replay(chunk)
            compare_books(sim, env, o, ("deep replay", env, chunk))
    n_side = [len(sim.dump_book(0, side)) for side in (0, 1)]
    assert max(n_side) > 1600, n_side                       # deeper than the largest compiled shared-memory layout (1 536)
    agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(1, 2, 1, 2, 0))
    obs0 = sim.reset(0, np.array([1000, 2000, 3000], np.int32)).cpu().numpy()
    obs, act, rew, done = (x.cpu().numpy() for x in sim.rollout(50, agent))
    assert np.all(sim.state()["err"] == 0)
    for env, start in enumerate((1000, 2000, 3000)):
        o = Oracle(abi.default_cfg(**kw), s)
        o.reset(start)
        oo, oa, orw, od = o.rollout(50, agent)
        assert np.allclose(obs[:, env], oo, rtol=1e-6, atol=1e-9) and np.allclose(rew[:, env], orw, rtol=1e-6, atol=1e-9)
        compare_books(sim, env, o, ("deep env", env))


@pytest.mark.replay_path
@pytest.mark.parametrize("deep", [False, True])
def test_messages_with_nonpositive_volume_set_bad_volume_and_are_skipped(deep):
    """`assert order.volume > 0` (Exchange.py:59-60): the reference raises; the device flags the env (LOBSIM_ERR_BAD_VOLUME), skips
    the message and goes on -- on every replay implementation (the flat / hybrid loops look at the volumes of a whole message
    segment at once), with the books still equal to the oracle's."""
    from oracle.oracle import Oracle
    from rl4mm_b200 import synthetic
    from test_gpu_parity import compare_books

    sc = synthetic.SynthConfig(seed=21, n_msgs=30_000, duration_s=60, n_levels=50 if deep else 10, target_orders=500 if deep else 50,
                               mean_queue=8 if deep else 4, init_levels=55 if deep else 30)
    s = synthetic.generate(sc)
    rng = np.random.default_rng(3)
    bad = rng.choice(np.arange(200, s.n_msgs), size=40, replace=False)
    s.msgs["volume"][bad[:20]] = 0
    s.msgs["volume"][bad[20:]] = -7
    kw = dict(n_levels=sc.n_levels, outer_levels=5, resync=1)
    cap = dict(max_levels_per_side=128, max_orders_per_side=1024) if deep else {}
    n = 3
    sim = _sim(abi.default_cfg(n_envs=n, **kw, **cap), [s])
    starts = np.array([0, 100, 200], np.int32)
    sim.reset_book(0, starts)
    oracles = [Oracle(abi.default_cfg(n_envs=1, **kw), s) for _ in range(n)]
    for o, st in zip(oracles, starts):
        o.reset_book(int(st))
    for chunk in (3, 150, 200):
        sim.replay(chunk)
        st = sim.state()
        for env, o in enumerate(oracles):
            o.replay(chunk)
            os_ = o.state()
            assert int(st["err"][env]) == int(os_["err"]), (env, chunk, int(st["err"][env]), int(os_["err"]))
            compare_books(sim, env, o, f"bad volume env {env} chunk {chunk}")
    assert all(int(e) & abi.ERR_BAD_VOLUME for e in sim.state()["err"])
    sim.close()
