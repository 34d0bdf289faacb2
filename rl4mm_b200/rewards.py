"""Mirror of ``rl4mm/rewards/RewardFunctions.py``: descriptors of the rewards the kernel computes (csrc/env.cuh
``reward_calc``) plus the same host-side ``calculate`` for API compatibility."""
from __future__ import annotations

import abc

from . import abi


class RewardFunction(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def calculate(self, current_state, next_state) -> float:
        pass

    def reset(self):
        pass

    @abc.abstractmethod
    def to_abi(self) -> abi.Reward:
        pass


class PnL(RewardFunction):
    def calculate(self, current_state, next_state) -> float:
        current_value = current_state.portfolio.cash + current_state.portfolio.inventory * current_state.price
        next_value = next_state.portfolio.cash + next_state.portfolio.inventory * next_state.price
        return next_value - current_value

    def to_abi(self):
        return abi.Reward(abi.REWARD_PNL, 0, 0.0)


class InventoryAdjustedPnL(RewardFunction):
    def __init__(self, inventory_aversion: float, asymmetrically_dampened: bool = False):
        self.inventory_aversion = inventory_aversion
        self.pnl = PnL()
        self.asymmetrically_dampened = asymmetrically_dampened

    def calculate(self, current_state, next_state) -> float:
        delta_midprice = next_state.price - current_state.price
        dampened_inventory_term = self.inventory_aversion * next_state.portfolio.inventory * delta_midprice
        if self.asymmetrically_dampened:
            dampened_inventory_term = max(0, dampened_inventory_term)
        return self.pnl.calculate(current_state, next_state) - dampened_inventory_term

    def to_abi(self):
        return abi.Reward(abi.REWARD_INV_ADJ_PNL, int(self.asymmetrically_dampened), float(self.inventory_aversion))


class RollingSharpe(RewardFunction):
    """RewardFunctions.py:38-94.  On the device the AUM window lives in HBM per env and, as in the reference, is never
    reset by the environment."""

    def __init__(self, max_window_size: int = 120, min_window_size: int = 60):
        assert max_window_size >= min_window_size, "Error with window sizes"
        self.max_window_size, self.min_window_size = max_window_size, min_window_size

    def calculate(self, current_state, next_state):
        raise NotImplementedError("RollingSharpe is stateful: it is evaluated by the kernel (csrc/env.cuh rolling_sharpe_step)")

    def to_abi(self):
        return abi.rolling_sharpe(self.max_window_size, self.min_window_size)
