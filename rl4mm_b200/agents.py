"""Mirror of ``rl4mm/agents`` with batched ``get_action`` (obs [N, F] -> actions [N, A]) and the descriptor the fused
rollout kernel needs (``to_abi``)."""
from __future__ import annotations

import abc
import ctypes

import numpy as np

from . import abi


class Agent(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def get_action(self, state: np.ndarray) -> np.ndarray:
        pass

    def to_abi(self):
        return None  # no fused-rollout implementation: the env steps with externally supplied actions


class RandomAgent(Agent):
    """baseline_agents.py:9-18.  ``get_action`` samples the env's action box on the host (one vectorised draw for a batch of
    states); ``env.rollout(agent, T)`` runs the same kind of agent fused on the device (``LOBSIM_AGENT_RANDOM``: a
    counter-based Philox stream keyed by (seed, env, grid step) -- a different stream than the host generator's)."""

    def __init__(self, env, seed: int = None):
        self.action_space = env.action_space
        self.action_space.seed(seed)
        self.n_envs = getattr(env, "n_envs", 1)
        self.seed = 0 if seed is None else int(seed)
        self._rng = np.random.default_rng(seed)

    def get_action(self, state: np.ndarray) -> np.ndarray:
        state = np.asarray(state)
        if state.ndim == 1:
            return self.action_space.sample()
        low, high = np.asarray(self.action_space.low, np.float64), np.asarray(self.action_space.high, np.float64)
        return low + (high - low) * self._rng.random((state.shape[0], len(low)))

    def to_abi(self):
        high = [float(x) for x in np.asarray(self.action_space.high, np.float64).reshape(-1)]
        assert np.all(np.asarray(self.action_space.low) == 0), "the device RandomAgent samples Box(0, high)"
        return abi.Agent(kind=abi.AGENT_RANDOM, fixed_action=(ctypes.c_double * 5)(*(high + [0.0] * (5 - len(high)))),
                         reserved=self.seed & 0x7FFFFFFF)

    def get_name(self):
        return "RandomAgent"


class FixedActionAgent(Agent):
    """baseline_agents.py:21-30"""

    def __init__(self, fixed_action: np.ndarray):
        self.fixed_action = np.asarray(fixed_action, dtype=np.float64)

    def get_action(self, state: np.ndarray) -> np.ndarray:
        state = np.asarray(state)
        if state.ndim == 1:
            return self.fixed_action
        return np.broadcast_to(self.fixed_action, (state.shape[0], len(self.fixed_action))).copy()

    def get_name(self):
        return "FixedAction_" + "_".join(map(str, self.fixed_action))

    def to_abi(self):
        a = list(self.fixed_action) + [0.0] * (5 - len(self.fixed_action))
        return abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(*a))


class Teradactyl(Agent):
    """baseline_agents.py:33-108 (vectorised over envs)."""

    def __init__(self, max_inventory=None, default_kappa: float = 10.0, default_omega: float = 0.5,
                 max_kappa: float = 10.0, exponent: float = 1.0, market_clearing: bool = False,
                 inventory_index: int = 3):
        self.max_inventory, self.default_kappa, self.default_omega = max_inventory, default_kappa, default_omega
        self.max_kappa, self.exponent, self.market_clearing = max_kappa, exponent, market_clearing
        self.inventory_index = inventory_index
        self.eps = 0.00001
        self.denom = 100 if max_inventory is None else max_inventory

    def clamp_to_unit(self, x, strict_containment: bool = True):
        if strict_containment:
            return np.maximum(np.minimum(x, 1 - self.eps), -1 + self.eps)
        return np.maximum(np.minimum(x, 1), -1)

    def get_omega_bid_and_ask(self, inventory):
        inventory = np.asarray(inventory, dtype=np.float64)
        c = self.clamp_to_unit(inventory / self.denom)
        w = self.default_omega
        pos_bid = w * (1 + (1 / w - 1) * np.abs(c) ** self.exponent)
        pos_ask = w * (1 - np.abs(c) ** self.exponent)
        omega_bid = np.where(inventory >= 0, pos_bid, pos_ask)
        omega_ask = np.where(inventory >= 0, pos_ask, pos_bid)
        return omega_bid, omega_ask

    def get_kappa(self, inventory):
        return (self.max_kappa - self.default_kappa) * np.abs(np.asarray(inventory) / self.max_inventory) ** self.exponent \
            + self.default_kappa

    @staticmethod
    def calculate_alpha(omega, kappa):
        return (omega * (kappa - 2)) + 1

    @staticmethod
    def calculate_beta(omega, kappa):
        return (1 - omega) * (kappa - 2) + 1

    def get_action(self, state: np.ndarray) -> np.ndarray:
        state = np.asarray(state, dtype=np.float64)
        inventory = state[..., self.inventory_index]
        omega_bid, omega_ask = self.get_omega_bid_and_ask(inventory)
        kappa = self.get_kappa(inventory)
        cols = [self.calculate_alpha(omega_bid, kappa), self.calculate_beta(omega_bid, kappa),
                self.calculate_alpha(omega_ask, kappa), self.calculate_beta(omega_ask, kappa)]
        if self.market_clearing is True:
            cols.append(np.full_like(cols[0], self.max_inventory * 2))
        return np.stack(cols, axis=-1)

    def get_name(self):
        return (f"Teradactyl_def_omega_{self.default_omega}_def_kappa_{self.default_kappa}_"
                f"max_inv_{self.max_inventory}_max_kappa_{self.max_kappa}_exponent_{self.exponent}")

    def to_abi(self):
        return abi.Agent(kind=abi.AGENT_TERADACTYL, inventory_index=self.inventory_index,
                         max_inventory=float(self.max_inventory or 0.0), default_kappa=self.default_kappa,
                         default_omega=self.default_omega, max_kappa=self.max_kappa, exponent=self.exponent,
                         market_clearing=int(bool(self.market_clearing)))
