#!/usr/bin/env python
"""Experiment: configs[2] rollout (torch MLP policy between one-step env launches) with the envs split into G groups, each with its
own handle and CUDA stream, so that the policy kernels of one group overlap the env-step kernel of another.
    python tools/pipe_rollout.py [n_envs] [groups ...]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from rl4mm_b200 import synthetic  # noqa: E402
from rl4mm_b200.device import LobSim  # noqa: E402

n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
groups = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 0]      # 0: one group, policy + step captured in a CUDA graph
dev = torch.device("cuda:0")
stream = synthetic.generate(synthetic.spy_day(seed=0, n_msgs=10_000_000))
T, reps, warm = 128, 3, 2
for G_ in groups:
    G = max(G_, 1)
    n = n_envs // G
    sims, obs, strs = [], [], []
    rng = np.random.default_rng(1234)
    for g in range(G):
        sim = LobSim(bench.rollout_cfg(n, stream), 0)
        sim.load_stream(0, stream)
        starts = ((1800 + rng.integers(0, 5 * 3600, size=n)) * stream.steps_per_second).astype(np.int32)
        obs.append(sim.reset(0, starts))
        sims.append(sim)
        strs.append(torch.cuda.Stream(dev) if G > 1 else torch.cuda.current_stream(dev))
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(obs[0].shape[1], 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 4)).to(dev)
    scale = torch.tensor([1e-2, 1e-2, 1e-2, 1e3, 1e3, 1e-2, 1.0, 0.1, 1.0, 1.0], device=dev, dtype=torch.float64)

    def act(o):
        with torch.no_grad():
            return (torch.sigmoid(policy((o * scale).float())) * 10.0).double()

    torch.cuda.synchronize(dev)

    def run():
        for _ in range(T):
            for g in range(G):
                with torch.cuda.stream(strs[g]):
                    obs[g], r, d = sims[g].step(act(obs[g]), stream=strs[g])

    if G_ == 0:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                obs[0].copy_(sims[0].step(act(obs[0]))[0])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            o, r, d = sims[0].step(act(obs[0]))
            obs[0].copy_(o)

        def run():  # noqa: F811
            for _ in range(T):
                graph.replay()

    for _ in range(warm):
        run()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    errs = sum(int((s.errors() != 0).sum()) for s in sims)
    print(f"groups {G_}: {reps * T * n * G / dt:.4e} env steps/s  ({1e3 * dt / (reps * T):.3f} ms per step of all groups)  errs {errs}", flush=True)
    for s in sims:
        s.close()
