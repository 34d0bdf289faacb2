"""Policy-update loop around the device env (replaces the RLlib PPO trainer of the reference's
main.py:48-58 / rl4mm/utils/utils.py get_ray_config).  Plumbing only: torch modules + torch.distributed (DDP over NCCL),
the env step is the CUDA path (``LobSim.step`` on device tensors, no host round trip).

* policy: MLP (2 x 64 tanh, RLlib's default fcnet) -> per-action-dimension Beta(a, b) on (0, 1), scaled to the env's
  action box (0, max_distribution_param]; value head on a separate MLP;
* rollouts: all envs of a rank step in lock-step for T steps (fixed episode length => every env finishes together and
  the batch is reset at once, i.e. the reference's env.reset() per worker);
* update: GAE(lambda) + clipped surrogate + value loss + entropy bonus, minibatch Adam; with world_size > 1 the
  modules are wrapped in DistributedDataParallel so gradients are all-reduced over NVLink;
* observation normalisation: the batch stores the observations AS THE POLICY SAW THEM (normalised with the running
  statistics in force during collection); the statistics are updated only after the PPO epochs, so ratio == 1 at the
  first minibatch and returns / values share one input scaling;
* small batches are launch-bound (about 40 small torch kernels + one env kernel per step, 2.8 ms per step of Python and launch
  overhead at 4 096 envs): with ``use_cuda_graph`` (default for n_envs <= 16 384) ONE collection step -- normalise, policy
  forward, Beta sample, ``lobsim_step`` on the capturing stream, and the writes into slot t of the rollout tensors through a
  device-side step counter -- is captured once in a CUDA graph and replayed T times (the reference's per-step caller is
  rl4mm/gym/utils.py:100-117);
* device errors: ``step_torch`` does not poll, so the per-env error column is read once per rollout and before every
  batch reset (the reset kernel clears the flags): EmptyOrderbookError / overflow / bad-action envs raise, exactly as
  ``env.step`` would have (``on_error="raise"``), or are masked out of the update (``on_error="mask"``).
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn


@dataclass
class PPOConfig:
    rollout_steps: int = 128
    epochs: int = 4
    minibatches: int = 4
    gamma: float = 0.99
    lam: float = 0.95
    clip: float = 0.2
    vf_coef: float = 0.5
    ent_coef: float = 0.0
    lr: float = 3e-4
    max_grad_norm: float = 0.5
    hidden: int = 64
    reward_scale: float = 1.0


class BetaPolicy(nn.Module):
    def __init__(self, obs_dim: int, action_dim: int, action_low: torch.Tensor, action_high: torch.Tensor, hidden: int = 64):
        super().__init__()
        mlp = lambda out: nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, out))  # noqa: E731
        self.pi, self.vf = mlp(2 * action_dim), mlp(1)
        self.register_buffer("low", action_low.float())
        self.register_buffer("span", (action_high - action_low).float())

    def dist(self, obs: torch.Tensor) -> torch.distributions.Beta:
        a, b = (nn.functional.softplus(self.pi(obs)) + 1.0).chunk(2, dim=-1)      # unimodal: a, b > 1
        return torch.distributions.Beta(a, b, validate_args=False)   # (validation synchronises: not allowed while a CUDA graph is captured)

    def forward(self, obs: torch.Tensor):
        return self.dist(obs), self.vf(obs).squeeze(-1)

    def to_env(self, x: torch.Tensor) -> torch.Tensor:
        """(0,1)^A sample -> env action (Beta-ladder parameters in the env's action box, open at 0)."""
        return self.low + self.span * x.clamp(1e-6, 1.0)


class RunningNorm:
    """Running mean / variance of the observations (Welford, batched); synchronised over ranks when distributed."""

    def __init__(self, dim: int, device):
        self.n = torch.zeros((), dtype=torch.float64, device=device)
        self.mean = torch.zeros(dim, dtype=torch.float64, device=device)
        self.m2 = torch.zeros(dim, dtype=torch.float64, device=device)

    def update(self, x: torch.Tensor) -> None:
        x = x.reshape(-1, x.shape[-1]).double()
        x = torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)
        stats = torch.cat([torch.tensor([x.shape[0]], dtype=torch.float64, device=x.device), x.sum(0), (x * x).sum(0)])
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(stats)
        d = x.shape[1]
        nb, sb, qb = stats[0], stats[1:1 + d], stats[1 + d:]
        mb = sb / nb
        m2b = qb - nb * mb * mb
        delta = mb - self.mean
        tot = self.n + nb
        # in place: a captured CUDA graph of the collection step keeps reading these tensors
        self.m2.add_(m2b + delta * delta * self.n * nb / tot)
        self.mean.add_(delta * nb / tot)
        self.n.copy_(tot)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        std = torch.sqrt(torch.clamp(self.m2, min=0.0) / torch.clamp(self.n, min=1.0)).clamp(min=1e-8)   # (m2 of a constant feature can round below 0)
        z = (torch.nan_to_num(x.double(), nan=0.0, posinf=0.0, neginf=0.0) - self.mean) / std
        return z.clamp(-10.0, 10.0).float()


def gae(rew: torch.Tensor, val: torch.Tensor, last_val: torch.Tensor, done: torch.Tensor, gamma: float, lam: float):
    """Generalised advantage estimation over [T, N] tensors; `done[t]` ends the episode AFTER step t."""
    T = rew.shape[0]
    adv = torch.zeros_like(rew)
    nxt, run = last_val, torch.zeros_like(last_val)
    for t in range(T - 1, -1, -1):
        nd = 1.0 - done[t].float()
        delta = rew[t] + gamma * nxt * nd - val[t]
        run = delta + gamma * lam * nd * run
        adv[t], nxt = run, val[t]
    return adv, adv + val


class PPOTrainer:
    def __init__(self, env, cfg: PPOConfig = PPOConfig(), seed: int = 0, on_error: str = "raise", use_cuda_graph: Optional[bool] = None):
        """`env`: rl4mm_b200.gym.HistoricalOrderbookEnvironment (batched).  One trainer per rank / GPU."""
        assert on_error in ("raise", "mask")
        self.env, self.cfg, self.on_error = env, cfg, on_error
        self.use_cuda_graph = env.n_envs <= 16_384 if use_cuda_graph is None else bool(use_cuda_graph)
        self._graph = None
        self.device = env.sim.device
        torch.manual_seed(seed)
        low = torch.as_tensor(env.action_space.low, device=self.device)
        high = torch.as_tensor(env.action_space.high, device=self.device)
        self.policy = BetaPolicy(env.sim.obs_dim, env.sim.action_dim, low, high, cfg.hidden).to(self.device)
        self.module = self.policy
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            self.module = nn.parallel.DistributedDataParallel(self.policy, device_ids=[self.device.index])
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=cfg.lr, eps=1e-5)
        self.norm = RunningNorm(env.sim.obs_dim, self.device)
        self.obs: Optional[torch.Tensor] = None
        self.steps_left = 0
        self.episode_return = torch.zeros(env.n_envs, dtype=torch.float64, device=self.device)
        self.finished_returns = []
        self.bad_envs = torch.zeros(env.n_envs, dtype=torch.bool, device=self.device)   # envs that died in the current episode

    def _poll_errors(self) -> None:
        """Read the per-env error flags (one D2H copy).  Must run BEFORE a reset: the reset kernel clears them."""
        from . import abi

        err = torch.as_tensor(self.env.sim.errors().astype("int64"), device=self.device)
        fatal = abi.ERR_EMPTY_BOOK | abi.ERR_LEVEL_OVERFLOW | abi.ERR_ORDER_OVERFLOW | abi.ERR_AGENT_OVERFLOW | abi.ERR_BAD_ACTION \
            | abi.ERR_END_OF_STREAM | abi.ERR_NO_SNAPSHOT
        bad = (err & fatal) != 0
        if bool(bad.any()):
            if self.on_error == "raise":
                self.env._raise_on_errors()
            self.bad_envs |= bad

    def _reset(self):
        obs = torch.as_tensor(self.env.reset(), device=self.device).reshape(self.env.n_envs, -1)
        if self.obs is None:
            self.obs = obs.clone()
        else:
            self.obs.copy_(obs)                         # in place: the captured step reads this tensor
        self.steps_left = self.env.n_steps
        self.episode_return.zero_()
        self.bad_envs.zero_()

    def _one_step(self, b: Dict[str, torch.Tensor], t_idx: torch.Tensor) -> None:
        """One collection step into slot ``t_idx`` (a 1-element device tensor) of the rollout tensors; reads and rewrites
        ``self.obs`` in place.  Pure device work on the current stream: this is the body of the captured CUDA graph."""
        cfg = self.cfg
        obsn = self.norm(self.obs)
        dist, val = self.policy(obsn)
        x = dist.sample()
        obs, rew, done = self.env.step_torch(self.policy.to_env(x).double())
        logp = dist.log_prob(x.clamp(1e-6, 1 - 1e-6)).sum(-1)
        finite = torch.isfinite(rew)                    # an empty book side has no price: the reward of a dead env is NaN
        rew = torch.where(finite, rew, torch.zeros_like(rew))
        for key, v in (("obs", self.obs), ("obs_n", obsn), ("x", x), ("logp", logp), ("val", val), ("valid", finite),
                       ("rew", (rew * cfg.reward_scale).float()), ("done", done.bool())):
            b[key].index_copy_(0, t_idx, v.unsqueeze(0))
        self.episode_return += rew
        self.obs.copy_(obs)
        t_idx += 1

    def _capture(self, b: Dict[str, torch.Tensor]) -> None:
        """Capture ``_one_step`` once (after the side-stream warm-up torch asks for); the warm-up steps really step the envs,
        so it runs at construction-time state and the envs are reset afterwards."""
        self._t_idx = torch.zeros(1, dtype=torch.long, device=self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self._t_idx.zero_()
                self._one_step(b, self._t_idx)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        self._t_idx.zero_()
        with torch.cuda.graph(self._graph):
            self._one_step(b, self._t_idx)
        self._graph_bufs = b
        self._poll_errors()
        self._reset()                                   # the warm-up / capture steps moved the envs: start the episode afresh

    @torch.no_grad()
    def collect(self) -> Dict[str, torch.Tensor]:
        T, N, cfg = self.cfg.rollout_steps, self.env.n_envs, self.cfg
        if self.obs is None:
            self._reset()
        if float(self.norm.n) == 0:                     # first rollout: seed the statistics with the reset observations
            self.norm.update(self.obs)
        if self.use_cuda_graph and self._graph is not None:
            b = self._graph_bufs                        # the captured graph writes into these tensors
        else:
            b = dict(obs=torch.empty((T, N, self.env.sim.obs_dim), dtype=torch.float64, device=self.device),
                     obs_n=torch.empty((T, N, self.env.sim.obs_dim), device=self.device),
                     valid=torch.ones((T, N), dtype=torch.bool, device=self.device),
                     x=torch.empty((T, N, self.env.sim.action_dim), device=self.device),
                     logp=torch.empty((T, N), device=self.device), val=torch.empty((T, N), device=self.device),
                     rew=torch.empty((T, N), device=self.device), done=torch.zeros((T, N), dtype=torch.bool, device=self.device))
            if self.use_cuda_graph:
                self._capture(b)
        obs_b, obsn_b, valid_b, x_b, logp_b, val_b, rew_b, done_b = (b[k] for k in ("obs", "obs_n", "valid", "x", "logp", "val", "rew", "done"))
        t_idx = self._t_idx if self.use_cuda_graph else torch.zeros(1, dtype=torch.long, device=self.device)
        t_idx.zero_()
        for t in range(T):
            if self.use_cuda_graph:
                self._graph.replay()
            else:
                self._one_step(b, t_idx)
            self.steps_left -= 1
            if self.steps_left == 0:                    # every env ends together: batch reset (env.reset per worker)
                self._poll_errors()                     # before the reset kernel clears the flags
                valid_b[: t + 1] &= ~self.bad_envs
                good = ~self.bad_envs
                self.finished_returns.append(float(self.episode_return[good].mean()) if bool(good.any()) else float("nan"))
                self._reset()
        self._poll_errors()
        valid_b &= ~self.bad_envs                       # (conservative: the whole rollout of an env that died in it)
        last_val = self.policy(self.norm(self.obs))[1]
        adv, ret = gae(rew_b, val_b, last_val, done_b, cfg.gamma, cfg.lam)
        return dict(obs=obs_b, obs_n=obsn_b, x=x_b, logp=logp_b, adv=adv, ret=ret, rew=rew_b, valid=valid_b)

    def update(self, batch: Dict[str, torch.Tensor]) -> Dict[str, float]:
        cfg = self.cfg
        obs = batch["obs_n"].flatten(0, 1)              # as seen by the policy during collection
        x, logp0 = batch["x"].flatten(0, 1).clamp(1e-6, 1 - 1e-6), batch["logp"].flatten()
        adv, ret = batch["adv"].flatten(), batch["ret"].flatten()
        keep = batch["valid"].flatten()
        if not bool(keep.all()):                        # dead envs (on_error="mask") / non-finite rewards stay out of the update
            obs, x, logp0, adv, ret = obs[keep], x[keep], logp0[keep], adv[keep], ret[keep]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        n = obs.shape[0]
        stats = {}
        for _ in range(cfg.epochs):
            perm = torch.randperm(n, device=self.device)
            for idx in perm.chunk(cfg.minibatches):
                dist, val = self.module(obs[idx])
                logp = dist.log_prob(x[idx]).sum(-1)
                ratio = torch.exp(logp - logp0[idx])
                pg = -torch.min(ratio * adv[idx], ratio.clamp(1 - cfg.clip, 1 + cfg.clip) * adv[idx]).mean()
                vf = 0.5 * (val - ret[idx]).pow(2).mean()
                ent = dist.entropy().sum(-1).mean()
                loss = pg + cfg.vf_coef * vf - cfg.ent_coef * ent
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(self.policy.parameters(), cfg.max_grad_norm)
                self.opt.step()
                stats = dict(loss=loss.item(), pg=pg.item(), vf=vf.item(), entropy=ent.item(),
                             kl=(logp0[idx] - logp).mean().item())
        self.norm.update(batch["obs"])                  # running statistics move only after the PPO epochs
        stats["masked_fraction"] = 1.0 - float(keep.float().mean())
        return stats

    def train(self, iterations: int, log=None):
        history = []
        for it in range(iterations):
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            batch = self.collect()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            stats = self.update(batch)
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            stats.update(iteration=it, mean_step_reward=float(batch["rew"].mean()),
                         env_steps_per_sec=self.cfg.rollout_steps * self.env.n_envs / (t1 - t0), collect_s=t1 - t0, update_s=t2 - t1,
                         episode_reward_mean=self.finished_returns[-1] if self.finished_returns else float("nan"))
            history.append(stats)
            if log:
                log(stats)
        return history
