#!/usr/bin/env python
"""A tiny run of the hybrid deep-book replay kernel (k_replay_hyb, book_hybrid.cuh) for compute-sanitizer: 50-level books with
deep queues on the 128/1024/64 layout -- level moves between the hot pool and the cold arrays, spills, bails to the sorted form --
with every book compared with the oracle."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.oracle import Oracle  # noqa: E402
from rl4mm_b200 import abi, synthetic  # noqa: E402
from rl4mm_b200.device import LobSim  # noqa: E402

for name, sk in (("deep", dict(geom_p=0.12, max_offset_ticks=70, target_orders=600, mean_queue=12, p_sweep=0.01)),
                 ("long_queues", dict(geom_p=0.6, max_offset_ticks=2, target_orders=500, mean_queue=12, p_sweep=0.01))):
    s = synthetic.generate(synthetic.SynthConfig(seed=11, n_msgs=40_000, duration_s=60, n_levels=50, mid0=2_000_000, p_limit=0.4,
                                                 p_cancel=0.2, p_delete=0.3, p_exec=0.1, init_levels=55, **sk))
    kw = dict(n_levels=50, outer_levels=20, resync=1)
    starts = np.array([0, 50, 150, 250], np.int32)
    sim = LobSim(abi.default_cfg(n_envs=4, max_levels_per_side=128, max_orders_per_side=1024, max_agent_orders=64, **kw), 0)
    sim.load_stream(0, s)
    sim.reset_book(0, starts)
    oracles = [Oracle(abi.default_cfg(**kw), s) for _ in starts]
    for o, st in zip(oracles, starts):
        o.reset_book(int(st))
    for chunk in (7, 120, 150):
        sim.replay(chunk)
        for env, o in enumerate(oracles):
            o.replay(chunk)
            for side in (0, 1):
                d, e = sim.dump_book(env, side), o.dump_book(side)
                assert np.array_equal(d[["price", "volume", "ref"]], e[["price", "volume", "ref"]]), (name, env, side, chunk)
    assert not sim.errors().any()
    sim.close()
print("tiny_hybrid ok")
