#!/usr/bin/env python
"""A tiny run on the 128/256/64 layout (flat pools of 4 chunks = 128 orders per side) for compute-sanitizer: replay on k_replay_flat,
env steps on the classic kernel (reads the flat blobs), the L3 dump (k_to_sorted), all against the oracle."""
import ctypes
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.oracle import Oracle  # noqa: E402
from rl4mm_b200 import abi, synthetic  # noqa: E402
from rl4mm_b200.device import LobSim  # noqa: E402

s = synthetic.generate(synthetic.SynthConfig(seed=5, n_msgs=60_000, duration_s=120, target_orders=150, mean_queue=6))
feats = [abi.feature(abi.FEAT_SPREAD, 0, 100000, 0, 5000), abi.feature(abi.FEAT_INVENTORY, 0, 100000, -1e6, 1e6)]
kw = dict(n_levels=10, outer_levels=20, features=feats, episode_steps=100, warmup_steps=0)
starts = np.array([0, 100, 300, 500], np.int32)
sim = LobSim(abi.default_cfg(n_envs=4, max_levels_per_side=128, max_orders_per_side=256, max_agent_orders=64, **kw), 0)
sim.load_stream(0, s)
sim.reset_book(0, starts)
sim.replay(300)
forms = sim.state()["reserved"]
oracles = []
for env, st in enumerate(starts):
    o = Oracle(abi.default_cfg(**kw), s)
    o.reset_book(int(st))
    o.replay(300)
    oracles.append(o)
    for side in (0, 1):
        d, e = sim.dump_book(env, side), o.dump_book(side)
        assert np.array_equal(d[["price", "volume", "ref"]], e[["price", "volume", "ref"]]), (env, side)
agent = abi.Agent(kind=abi.AGENT_FIXED, fixed_action=(ctypes.c_double * 5)(1, 2, 1, 2, 0))
sim.replay(50)                                  # flat blobs again
obs0 = sim.reset(0, starts + 400).cpu().numpy()
obs, act, rew, done = (x.cpu().numpy() for x in sim.rollout(40, agent))
sim.replay(0)
for env, st in enumerate(starts):
    o = Oracle(abi.default_cfg(**kw), s)
    o.reset(int(st) + 400)
    oo, oa, orw, od = o.rollout(40, agent)
    assert np.allclose(obs[:, env], oo, rtol=1e-6, atol=1e-9) and np.allclose(rew[:, env], orw, rtol=1e-6, atol=1e-9)
print("tiny_flat4 ok: book forms after the replay", [int(f & 1) for f in forms], "orders per side", [int((f >> 8) & 0xfff) for f in forms])
