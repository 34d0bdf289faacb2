#!/usr/bin/env python
"""Throughput of deep-book mode (books worked on in place in HBM): 1 024 books of a synthetic stream with a
mean queue of 80 orders per level (~4 400 resting orders per side), capacities 256 levels x 16 384 orders per side."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl4mm_b200 import abi, synthetic  # noqa: E402
from rl4mm_b200.device import LobSim  # noqa: E402

sc = synthetic.SynthConfig(seed=11, n_msgs=300_000, duration_s=600, n_levels=50, mid0=2_000_000, p_limit=0.40, p_cancel=0.15, p_delete=0.37,
                           p_exec=0.08, geom_p=0.10, init_levels=70, mean_queue=80, target_orders=6000, max_offset_ticks=80)
s = synthetic.generate(sc)
n = 1024
sim = LobSim(abi.default_cfg(n_envs=n, n_levels=50, outer_levels=20, max_levels_per_side=256, max_orders_per_side=16384, max_agent_orders=64), 0)
assert sim.kernel_path == "deep"
sim.load_stream(0, s)
sim.reset_book(0, 0)
sim.replay(2000)                       # grow the books
torch.cuda.synchronize()
t0 = time.perf_counter()
sim.replay(2000)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
msgs = int(s.step_off[4000]) - int(s.step_off[2000])
st = sim.state()
assert np.all(st["err"] == 0)
print(f"deep-book mode: {n} books x {msgs} messages in {dt * 1e3:.1f} ms = {n * msgs / dt:.3e} msgs/s; orders per side "
      f"{int(((st['reserved'] >> 8) & 0xfff).max())}+ (clamped), blob {sim.state_bytes} bytes per book")
