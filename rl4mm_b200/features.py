"""Mirror of ``rl4mm/features/Features.py``: the same feature classes and constructor arguments, as *descriptors* of
what the kernel computes per lane (csrc/env.cuh ``feature_update_raw``).  ``Portfolio`` and ``State`` are kept for API
compatibility."""
from __future__ import annotations

from dataclasses import dataclass
from datetime import datetime, timedelta
from typing import Any

from . import abi


@dataclass
class Portfolio:
    inventory: int
    cash: float


@dataclass
class State:
    filled_orders: Any
    orderbook: Any
    price: float
    portfolio: Portfolio
    now_is: datetime


def _us(td: timedelta) -> int:
    return td // timedelta(microseconds=1)


class Feature:
    kind: int = -1

    def __init__(self, name: str, min_value: float, max_value: float, update_frequency: timedelta,
                 lookback_periods: int, normalisation_on: bool = False, max_norm_len: int = 10000):
        assert update_frequency <= timedelta(minutes=1), "HFT update frequency must be less than 1 minute."
        self.name, self.min_value, self.max_value = name, min_value, max_value
        self.update_frequency, self.lookback_periods = update_frequency, lookback_periods
        self.normalisation_on, self.max_norm_len = normalisation_on, max_norm_len
        self.current_value = 0.0

    @property
    def window_size(self) -> timedelta:
        return self.lookback_periods * self.update_frequency

    def _params(self):
        return dict()

    def to_abi(self) -> abi.Feature:
        return abi.feature(self.kind, self.lookback_periods, _us(self.update_frequency), self.min_value, self.max_value,
                           norm_len=self.max_norm_len if self.normalisation_on else 0, **self._params())


class Spread(Feature):
    kind = abi.FEAT_SPREAD

    def __init__(self, name="Spread", min_value=0, max_value=(50 * 100), update_frequency=timedelta(seconds=0.1),
                 normalisation_on=False, max_norm_len=10000):
        super().__init__(name, min_value, max_value, update_frequency, 0, normalisation_on, max_norm_len)


class BookImbalance(Feature):
    kind = abi.FEAT_BOOK_IMBALANCE

    def __init__(self, update_frequency=timedelta(seconds=0.1)):
        super().__init__("BookImbalance", -1, 1, update_frequency, 0, False, 0)


class PriceMove(Feature):
    kind = abi.FEAT_PRICE_MOVE

    def __init__(self, name="MidpriceMove", min_value=-100 * 100, max_value=100 * 100,
                 update_frequency=timedelta(seconds=1), lookback_periods=10, normalisation_on=False,
                 max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, lookback_periods, normalisation_on, max_norm_len)


class PriceRange(Feature):
    kind = abi.FEAT_PRICE_RANGE

    def __init__(self, name="PriceRange", min_value=0, max_value=(100 * 100), update_frequency=timedelta(seconds=1),
                 lookback_periods=10, normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, lookback_periods, normalisation_on, max_norm_len)


class Volatility(Feature):
    kind = abi.FEAT_VOLATILITY

    def __init__(self, name="Volatility", min_value=0, max_value=1.0, update_frequency=timedelta(seconds=1),
                 lookback_periods=10, normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, lookback_periods, normalisation_on, max_norm_len)


class Price(Feature):
    kind = abi.FEAT_PRICE

    def __init__(self, name="Price", min_value=0, max_value=(10_000 * 10_000), update_frequency=timedelta(seconds=1),
                 normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, 0, normalisation_on, max_norm_len)


class TradeDirectionImbalance(Feature):
    kind = abi.FEAT_TRADE_DIR_IMBALANCE

    def __init__(self, name="TradeImbalance", update_frequency=timedelta(seconds=0.1), lookback_periods=600,
                 track_internal=False, normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, -1.0, 1.0, update_frequency, lookback_periods, normalisation_on, max_norm_len)
        self.track_internal = track_internal

    def _params(self):
        return dict(iparam=int(self.track_internal))


class TradeVolumeImbalance(TradeDirectionImbalance):
    kind = abi.FEAT_TRADE_VOL_IMBALANCE

    def __init__(self, name="TradeVolumeImbalance", update_frequency=timedelta(seconds=0.1), lookback_periods=600,
                 track_internal=False, normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, update_frequency, lookback_periods, track_internal, normalisation_on, max_norm_len)


class Inventory(Feature):
    kind = abi.FEAT_INVENTORY

    def __init__(self, name="Inventory", min_value=-1000000, max_value=1000000,
                 update_frequency=timedelta(seconds=0.1), normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, 0, normalisation_on, max_norm_len)


class EpisodeProportion(Feature):
    kind = abi.FEAT_EPISODE_PROPORTION

    def __init__(self, name="EpisodeProportion", update_frequency=timedelta(seconds=0.1),
                 episode_length=timedelta(minutes=60), normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, 0.0, 1.0, update_frequency, 0, normalisation_on, max_norm_len)
        self.episode_length = episode_length
        self.step_size = update_frequency / episode_length

    def _params(self):
        return dict(dparam=self.step_size)


class TimeOfDay(Feature):
    kind = abi.FEAT_TIME_OF_DAY

    def __init__(self, name="TimeOfDay", update_frequency=timedelta(minutes=1), normalisation_on=False,
                 max_norm_len=100_000, n_buckets=10):
        super().__init__(name, 0, n_buckets - 1, update_frequency, 0, normalisation_on, max_norm_len)
        self.n_buckets = n_buckets

    def _params(self):
        return dict(iparam=self.n_buckets)


class AmihudLambda(Feature):
    kind = abi.FEAT_AMIHUD_LAMBDA

    def __init__(self, name="AmihudLambda", min_value=0, max_value=1.0, update_frequency=timedelta(seconds=0.1),
                 lookback_periods=10, slowing_factor=10, normalisation_on=False, max_norm_len=100_000):
        super().__init__(name, min_value, max_value, update_frequency, (lookback_periods + 1) * slowing_factor,
                         normalisation_on, max_norm_len)
        self.true_lookback_periods = lookback_periods
        self.slowing_factor = slowing_factor

    def _params(self):
        return dict(iparam=self.slowing_factor)
