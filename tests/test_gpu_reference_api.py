"""The reference's own unit / integration tests, re-expressed against the drop-in façade (rl4mm_b200.orderbook /
.simulation / .gym), which drives the CUDA path through the C ABI.  Each test names the reference test it mirrors."""
from collections import deque
from copy import copy, deepcopy
from datetime import datetime, timedelta

import numpy as np
import pytest
from sortedcontainers import SortedDict

import parity_helpers as H
from rl4mm_b200 import abi

pytestmark = pytest.mark.gpu

TICKER, TICK_SIZE = "MSFT", 100
T0 = datetime(2012, 6, 21, 12, 0)


def _orders():
    from rl4mm_b200.orderbook import Cancellation, Deletion, LimitOrder, MarketOrder

    # rl4mm/orderbook/tests/mock_orders.py:9-145
    m = {}
    m["LIMIT_1"] = LimitOrder(T0, "buy", TICKER, None, 50, True, int(30.1 * 10000), 1000)
    m["LIMIT_2"] = LimitOrder(T0 + timedelta(minutes=1), "buy", TICKER, None, 100, True, int(30.1 * 10000), 200)
    m["LIMIT_3"] = LimitOrder(T0 + timedelta(minutes=1), "buy", TICKER, None, 110, True, int(30.2 * 10000), 200)
    m["LIMIT_4"] = LimitOrder(T0 + timedelta(minutes=1), "sell", TICKER, None, None, False, int(30.3 * 10000), 200)
    m["LIMIT_5"] = LimitOrder(T0 + timedelta(minutes=1), "sell", TICKER, None, None, False, int(30.2 * 10000), 300)
    m["CANCELLATION_1"] = Cancellation(T0 + timedelta(minutes=1), "buy", TICKER, None, 50, True, int(30.1 * 10000), 200)
    m["DELETION_1"] = Deletion(T0 + timedelta(minutes=1), "buy", TICKER, None, 50, True, int(30.1 * 10000), 1000)
    m["DELETION_2"] = Deletion(T0 + timedelta(minutes=2), "buy", TICKER, None, 100, True, int(30.1 * 10000), None)
    m["DELETION_3"] = Deletion(T0 + timedelta(minutes=3), "sell", TICKER, 2, None, False, int(30.3 * 10000), None)
    m["MARKET_1"] = MarketOrder(T0 + timedelta(minutes=1), "sell", TICKER, None, 50, False, 1000)
    m["MARKET_2"] = MarketOrder(T0 + timedelta(minutes=2), "sell", TICKER, None, 100, True, 200)
    m["MARKET_3"] = MarketOrder(T0 + timedelta(minutes=3), "buy", TICKER, None, 110, True, 200)
    for i, k in enumerate(("LIMIT_1", "LIMIT_2", "LIMIT_3", "LIMIT_4"), 1):
        sub = deepcopy(m[k])
        sub.internal_id = i
        m[f"submission_{i}"] = sub
    return m


def get_mock_orderbook(m):
    from rl4mm_b200.orderbook import Orderbook

    return deepcopy(Orderbook(
        buy=SortedDict({int(30.1 * 10000): deque([m["submission_1"], m["submission_2"]]),
                        int(30.2 * 10000): deque([m["submission_3"]])}),
        sell=SortedDict({int(30.3 * 10000): deque([m["submission_4"]])}), ticker=TICKER, tick_size=TICK_SIZE))


@pytest.fixture()
def m():
    import torch

    assert torch.cuda.is_available()
    return _orders()


def new_exchange(*a, **k):
    from rl4mm_b200.orderbook import Exchange

    return Exchange(*a, **k)


# ---- rl4mm/orderbook/tests/testExchange.py ---------------------------------------------------------------------------
def test_post_init(m):
    from rl4mm_b200.orderbook import Orderbook

    exchange = new_exchange(TICKER)
    assert exchange.name == "NASDAQ"
    assert exchange.central_orderbook == Orderbook(buy=SortedDict(), sell=SortedDict(), ticker=TICKER, tick_size=TICK_SIZE)


def test_get_initial_orderbook_from_orders(m):
    exchange = new_exchange()
    initial_orders = [deepcopy(m["LIMIT_1"]), deepcopy(m["LIMIT_3"])]
    with pytest.raises(AssertionError):
        exchange.get_initial_orderbook_from_orders(initial_orders)
    for order in initial_orders:
        order.internal_id = -1
        order.external_id = None  # snapshot aggregates carry no external id (OrderbookSimulator.py:164-173)
    book = exchange.get_initial_orderbook_from_orders(initial_orders)
    assert list(book.buy.keys()) == [301000, 302000] and book.buy[301000][0] is initial_orders[0]
    exchange.central_orderbook = book  # and it round-trips through the device
    assert exchange.central_orderbook == book


def test_order_tracking(m):
    exchange = new_exchange(TICKER)
    for k in ("LIMIT_1", "LIMIT_2", "LIMIT_3", "LIMIT_4"):
        exchange.submit_order(m[k])
    count = 1
    for direction in ["buy", "sell"]:
        side = getattr(exchange.central_orderbook, direction)
        for level in side.keys():
            for order in side[level]:
                assert count == order.internal_id
                count += 1
    for count, k in enumerate(["LIMIT_1", "LIMIT_2", "LIMIT_3", "submission_4"]):
        assert count + 1 == exchange.order_id_convertor.get_internal_order_id(m[k])


def test_submit_order_basic(m):
    from rl4mm_b200.orderbook import Orderbook

    exchange = new_exchange(TICKER)
    for k in ("LIMIT_1", "LIMIT_2", "LIMIT_3", "LIMIT_4"):
        exchange.submit_order(m[k])
    assert get_mock_orderbook(m) == exchange.central_orderbook
    expected_internal = Orderbook(buy=SortedDict(), sell=SortedDict({303000: deque([m["submission_4"]])}),
                                  ticker=TICKER, tick_size=TICK_SIZE)
    assert expected_internal == exchange.internal_orderbook


def test_execute_market_order(m):
    orderbook = get_mock_orderbook(m)
    exchange = new_exchange(TICKER, deepcopy(orderbook))
    exchange.execute_order(m["MARKET_1"])
    orderbook.buy.pop(302000)
    orderbook.buy[301000][0].volume -= m["MARKET_1"].volume - m["LIMIT_3"].volume
    assert orderbook == exchange.central_orderbook
    exchange.execute_order(m["MARKET_2"])
    orderbook.buy[301000].popleft()
    assert orderbook == exchange.central_orderbook
    exchange.execute_order(m["MARKET_3"])
    orderbook.sell.pop(303000)
    assert orderbook == exchange.central_orderbook


def test_cancel_order_basic(m):
    from rl4mm_b200.orderbook import Orderbook

    exchange = new_exchange(TICKER)
    modified = deepcopy(m["submission_1"])
    modified.volume = 1000 - 200
    expected = Orderbook(buy=SortedDict({301000: deque([modified])}), sell=SortedDict({}), ticker=TICKER, tick_size=TICK_SIZE)
    exchange.submit_order(m["LIMIT_1"])
    exchange.remove_order(m["CANCELLATION_1"])
    assert expected == exchange.central_orderbook


def test_large_cancellation_is_clamped(m):
    """testExchange.py:119-123 asserts nothing; the behaviour (Exchange.py:142-146) is: remove the resting volume."""
    from rl4mm_b200.orderbook import Cancellation

    exchange = new_exchange(TICKER)
    exchange.submit_order(m["LIMIT_1"])
    exchange.remove_order(Cancellation(T0, "buy", TICKER, None, 50, True, 301000, 1100))
    assert exchange.central_orderbook == exchange.get_empty_orderbook()


def test_delete_order_basic(m):
    exchange = new_exchange(TICKER)
    exchange.submit_order(m["LIMIT_1"])
    exchange.remove_order(m["DELETION_1"])
    assert exchange.get_empty_orderbook() == exchange.central_orderbook


def test_delete_order_internal(m):
    from rl4mm_b200.orderbook import Orderbook

    exchange = new_exchange(TICKER)
    exchange.submit_order(m["LIMIT_1"])
    exchange.submit_order(m["LIMIT_4"])
    exchange.remove_order(m["DELETION_3"])
    assert exchange.get_empty_orderbook() == exchange.internal_orderbook
    expected = Orderbook(buy=SortedDict({301000: deque([m["submission_1"]])}), sell=SortedDict(), ticker=TICKER, tick_size=TICK_SIZE)
    assert expected == exchange.central_orderbook


def test_delete_order_with_no_volume_given(m):
    exchange = new_exchange(TICKER)
    exchange.submit_order(m["LIMIT_2"])
    exchange.remove_order(m["DELETION_2"])
    assert exchange.get_empty_orderbook() == exchange.central_orderbook


def test_submit_limit_order_crossing_spread(m):
    from rl4mm_b200.orderbook import Orderbook

    exchange = new_exchange(TICKER, get_mock_orderbook(m))
    submission_5 = copy(m["LIMIT_5"])
    submission_5.internal_id = 1
    filled = exchange.submit_order(submission_5)
    rest = copy(submission_5)
    rest.volume = m["LIMIT_5"].volume - m["LIMIT_3"].volume
    got = exchange.central_orderbook
    assert list(got.buy.keys()) == [301000] and list(got.sell.keys()) == [302000, 303000]
    assert got.buy[301000] == get_mock_orderbook(m).buy[301000]
    assert [(o.volume, o.is_external) for o in got.sell[302000]] == [(100, False)]
    assert got.sell[303000] == deque([m["submission_4"]])
    assert [(o.volume, o.price, o.direction) for o in filled.external] == [(200, 302000, "buy")]
    assert [(type(o).__name__, o.volume, o.price, o.direction) for o in filled.internal] == [("MarketOrder", 200, 302000, "sell")]


def test_orderbook_price_range(m):
    exchange = new_exchange(TICKER)
    exchange.central_orderbook = get_mock_orderbook(m)
    assert exchange.orderbook_price_range == (301000, 303000)


def test_empty_orderbook_error(m):
    from rl4mm_b200.orderbook import EmptyOrderbookError

    exchange = new_exchange(TICKER, get_mock_orderbook(m))
    with pytest.raises(EmptyOrderbookError):
        exchange.execute_order(m["MARKET_3"].__class__(T0, "buy", TICKER, None, 7, True, 100000))


def test_self_match_deletes_own_resting_order(m):
    """mock_orders.py:171 TODO in the reference; behaviour from Exchange.py:91-94."""
    from rl4mm_b200.orderbook import LimitOrder

    exchange = new_exchange(TICKER, get_mock_orderbook(m))
    buy = LimitOrder(T0, "buy", TICKER, None, None, False, 303000, 50)  # crosses the agent's own sell at 30.3
    filled = exchange.submit_order(buy)
    assert filled.internal == [] and filled.external == []
    assert list(exchange.central_orderbook.sell.keys()) == []            # own order deleted, nothing filled
    assert [o.volume for o in exchange.central_orderbook.buy[303000]] == [50]
    assert [o.volume for o in exchange.internal_orderbook.buy[303000]] == [50]


# ---- rl4mm/simulation/tests/testOrderbookSimulator.py ------------------------------------------------------------------
def make_database():
    from rl4mm_b200.simulation import DeviceDatabase

    db = DeviceDatabase()
    db.add_stream(TICKER, datetime(2012, 6, 21), H.load_fixture_stream("reference"))
    return db


def lobster_dict(orderbook, n_levels):
    """rl4mm/extras/orderbook_comparison.py:6-19"""
    out = {}
    for direction in ["buy", "sell"]:
        half = getattr(orderbook, direction)
        prices = reversed(half) if direction == "buy" else half
        for level, price in enumerate(prices):
            if level < n_levels:
                out[f"{direction}_price_{level}"] = float(price)
                out[f"{direction}_volume_{level}"] = float(sum(o.volume for o in half[price]))
    return out


def snapshot_dict(stream, second_index):
    out = {}
    for side, direction in enumerate(("buy", "sell")):
        for level in range(stream.n_levels):
            p, v = stream.snapshots[second_index, side, level]
            out[f"{direction}_price_{level}"], out[f"{direction}_volume_{level}"] = float(p), float(v)
    return out


def test_agreement_with_lobster(m):
    """testOrderbookSimulator.py:46-75: the simulated L2 book equals the LOBSTER orderbook file at t0, +1 s, +2 s."""
    from rl4mm_b200.simulation import OrderbookSimulator

    db = make_database()
    stream = db.streams[0]
    sim = OrderbookSimulator(TICKER, None, None, 50, db, preload_orders=False, outer_levels=48)
    start = datetime(2012, 6, 21, 10, 0, 0)
    sim.reset_episode(start)
    assert snapshot_dict(stream, 1) == lobster_dict(sim.exchange.central_orderbook, 50)
    sim.forward_step(start + timedelta(seconds=1))
    expected, actual = snapshot_dict(stream, 2), lobster_dict(sim.exchange.central_orderbook, 50)
    for key in expected.keys() - actual.keys():
        expected.pop(key)
    assert expected == actual
    sim.forward_step(start + timedelta(seconds=2))
    expected, actual = snapshot_dict(stream, 3), lobster_dict(sim.exchange.central_orderbook, 50)
    for key in (expected.keys() - actual.keys()) | {"buy_price_45", "buy_volume_45"}:
        expected.pop(key, None)
        actual.pop(key, None)
    assert expected == actual


def test_update_outer_levels_with_internal_orders(m):
    """testOrderbookSimulator.py:77-108: agent orders on resynchronised levels are re-queued behind the aggregate."""
    from rl4mm_b200.orderbook import LimitOrder
    from rl4mm_b200.simulation import OrderbookSimulator

    db = make_database()
    sim = OrderbookSimulator(TICKER, None, None, 50, db, preload_orders=False, outer_levels=48)
    start = datetime(2012, 6, 21, 10, 0, 0)
    sim.reset_episode(start)
    min_buy_price_0 = sim.min_buy_price
    book = sim.exchange.central_orderbook
    worst_buy, worst_sell = min(book.buy.keys()), max(book.sell.keys())
    internal_buy = LimitOrder(start, "buy", TICKER, None, None, False, worst_buy - 100, 100)
    internal_sell = LimitOrder(start, "sell", TICKER, None, None, False, worst_sell, 200)
    sim.forward_step(start + timedelta(seconds=1), internal_orders=[internal_buy, internal_sell])
    assert sim.min_buy_price < min_buy_price_0
    internal_level = sim.exchange.internal_orderbook.buy[worst_buy - 100]
    external_level = sim.exchange.central_orderbook.buy[worst_buy - 100]
    assert internal_level[0].volume == 100
    assert external_level[0].is_external and not external_level[1].is_external
    assert external_level[0].volume == 200 and external_level[1].volume == 100


def test_hidden_executions_are_dropped(m):
    """testHistoricalOrderGenerator.py:34-40: 1000 rows - hidden executions (5 on the fixture); one more row precedes the
    grid origin + 1 step."""
    s = H.load_fixture_stream("reference")
    assert s.n_msgs == 995


# ---- rl4mm/gym/tests/testHistoricalOrderbookEnvironment.py --------------------------------------------------------------
def make_env(**kw):
    from rl4mm_b200.features import Inventory, PriceMove, PriceRange, Spread
    from rl4mm_b200.gym import HistoricalOrderbookEnvironment

    args = dict(
        step_size=timedelta(milliseconds=100), episode_length=timedelta(seconds=1), min_date=datetime(2012, 6, 21),
        max_date=datetime(2012, 6, 21), min_start_timedelta=timedelta(hours=10, seconds=1),
        max_end_timedelta=timedelta(hours=10, seconds=2), max_quote_level=10, database=make_database(),
        features=[Inventory(), Spread(), PriceMove(lookback_periods=1), PriceRange(lookback_periods=1)],
        fill_log_capacity=256)
    args.update(kw)
    return HistoricalOrderbookEnvironment(**args)


def test_env_reset(m):
    """testHistoricalOrderbookEnvironment.py:59-64"""
    env = make_env()
    env.reset()
    actual = env.reset()
    for a, e in zip(actual, [0, 100, -343.1, 343.1]):
        assert round(abs(a - e), 1) == 0


def test_env_ladders(m):
    """testHistoricalOrderbookEnvironment.py:73-103: Beta(1,1) -> 10 lots of 10 per side; then Beta(1,2) on top of it
    -> the resting ladder becomes [19, 17, ..., 1] through limits (diff > 0) and cancels (diff < 0)."""
    env = make_env(portfolio_carryover=False)
    env.reset()
    st0 = env.sim.state()[0]
    env.step(np.array([1, 1, 1, 1]))
    fills = env.sim.fills(0)
    for side, base, sgn in ((0, int(st0["best_buy"]), -1), (1, int(st0["best_sell"]), 1)):
        got = {int(e["price"]): int(e["volume"]) for e in env.sim.dump_agent_orders(0, side)}
        if not len(fills):
            assert got == {base + sgn * 100 * k: 10 for k in range(10)}
        assert sum(got.values()) <= 100
    env2 = make_env(portfolio_carryover=False)
    env2.reset()
    obs, reward, done, info = env2.step(np.array([1, 2, 1, 2]))
    assert obs.shape == (4,) and isinstance(reward, float) and done is False and info == {}
    if not len(env2.sim.fills(0)):
        for side in (0, 1):
            vols = sorted((int(e["volume"]) for e in env2.sim.dump_agent_orders(0, side)), reverse=True)
            assert vols == [19, 17, 15, 13, 11, 9, 7, 5, 3, 1]


def test_env_episode_runs_to_done_and_info(m):
    from rl4mm_b200.agents import FixedActionAgent
    from rl4mm_b200.gym import SimpleInfoCalculator, generate_trajectory

    env = make_env(info_calculator=SimpleInfoCalculator())
    traj = generate_trajectory(FixedActionAgent(np.array([1, 2, 1, 2])), env)
    assert len(traj["rewards"]) == 10 and len(traj["observations"]) == 11
    info = traj["infos"][-1]
    assert {"asset_price", "inventory", "cash", "aum", "market_spread", "agent_spread", "midprice_offset"} <= set(info)
    # n_envs == 1: the reference's shapes (InfoCalculators.py:31-59) -- scalars, and 1-tuples of 1-D action slices
    assert info["market_spread"] == 100.0 and np.isscalar(info["aum"]) and isinstance(info["inventory"], int)
    assert isinstance(info["bid_action"], tuple) and info["bid_action"][0].shape == (2,) and info["ask_action"][0].shape == (2,)
    aum = np.array([i["aum"] for i in traj["infos"]])
    assert aum.shape == (10,) and np.diff(aum).shape == (9,)          # a [T] series: get_sharpe(aum_array) works as in the reference


def test_trade_imbalance_goldens(m):
    """testFeatures.py:152-196: TradeDirectionImbalance / TradeVolumeImbalance on the real MSFT flow."""
    import ctypes

    from rl4mm_b200.device import LobSim

    s = H.load_fixture_stream("reference")
    feats = [abi.feature(abi.FEAT_TRADE_DIR_IMBALANCE, 10, 100000, -1, 1),
             abi.feature(abi.FEAT_TRADE_VOL_IMBALANCE, 10, 100000, -1, 1, iparam=1)]
    sim = LobSim(abi.default_cfg(n_envs=1, warmup_steps=0, features=feats, episode_steps=1000), 0)
    sim.load_stream(0, s)
    sim.reset(0, s.step_of_time(36000.0))
    obs, _, _, _ = sim.rollout(18, abi.Agent(kind=abi.AGENT_NONE))
    obs = obs.cpu().numpy()[:, 0, :]
    assert np.all(obs[:9] == 0)
    exp_dir = [1.0] * 7 + [6 / 36, 13 / 43]
    exp_vol = [1.0] * 7 + [-988 / 14642, -188 / 15442]
    for i in range(9):
        assert abs(obs[9 + i, 0] - exp_dir[i]) < 1e-7 and abs(obs[9 + i, 1] - exp_vol[i]) < 1e-7, (i, obs[9 + i])


def test_teradactyl_matches_reference_formula(m):
    from rl4mm_b200.agents import Teradactyl

    ag = Teradactyl(max_inventory=100, default_kappa=8.0, default_omega=0.4, max_kappa=12.0, exponent=2.0, inventory_index=0)
    a = ag.get_action(np.array([[30.0], [-250.0], [0.0]]))
    assert a.shape == (3, 4)
    # inventory 0 => omega_bid = omega_ask = default_omega, kappa = default_kappa
    assert np.allclose(a[2], [0.4 * 6 + 1, 0.6 * 6 + 1, 0.4 * 6 + 1, 0.6 * 6 + 1])
    assert a[0, 0] > a[0, 2] and a[1, 0] < a[1, 2]  # long => skew bids away; short => the opposite


# ---- rl4mm/gym/utils.py evaluation path ---------------------------------------------------------------------------------
@pytest.mark.parametrize("case_idx", range(2))
def test_episode_summary_dict_matches_reference(m, case_idx, tmp_path):
    """get_episode_summary_dict (utils.py:120-201) through the façade: fused device rollout + info series vs the summary
    dict the unmodified reference produced (tests/golden/episode_summary.json.gz), then the JSON round trip (:417-420)."""
    import json

    from rl4mm_b200 import evaluation
    from rl4mm_b200.agents import FixedActionAgent, Teradactyl
    from rl4mm_b200.features import Inventory, Portfolio, Spread
    from rl4mm_b200.rewards import PnL

    case = H.load_golden("episode_summary.json.gz")[case_idx]
    a = case["agent"]
    agent = (FixedActionAgent(np.array(a["action"], dtype=float)) if a["kind"] == "fixed" else
             Teradactyl(max_inventory=a["max_inventory"], default_kappa=a["default_kappa"], default_omega=a["default_omega"],
                        max_kappa=a["max_kappa"], exponent=a["exponent"], inventory_index=a["inventory_index"]))
    kw = dict(features=[Spread(), Inventory(max_value=100000)], episode_length=timedelta(seconds=1.5),
              min_start_timedelta=timedelta(hours=10, seconds=1), max_end_timedelta=timedelta(hours=10, seconds=2.5),
              initial_portfolio=Portfolio(inventory=0, cash=case["initial_cash"]), per_step_reward_function=PnL(),
              terminal_reward_function=PnL(), n_levels=50, fill_log_capacity=0)
    env = make_env(**kw)
    esd = evaluation.get_episode_summary_dict(agent, env, case["n_iterations"])          # n_envs = 1: carry-over chain
    H.assert_summary_matches(esd, case, [float(evaluation.get_sharpe(c)) for c in esd["equity_curves"]])
    evaluation.save_episode_summary_json(esd, tmp_path / "esd.json")
    back = json.loads((tmp_path / "esd.json").read_text())
    assert set(back) == set(case["esd"]) and back["inventories"] == case["esd"]["inventories"]
    env3 = make_env(n_envs=3, **kw)                                                      # batch: every env = iteration 0
    esd3 = evaluation.get_episode_summary_dict(agent, env3)
    for key in esd3:
        for e in range(3):
            H.assert_close_vec(np.atleast_1d(np.asarray(esd3[key][e], float)), np.atleast_1d(np.asarray(case["esd"][key][0], float)), key)
    sh = evaluation.get_sharpe(np.stack(esd3["equity_curves"]))
    assert sh.shape == (3,) and H.close(sh[0], case["sharpe"][0])
