import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from rl4mm_b200 import abi, synthetic
from rl4mm_b200.device import LobSim
s = synthetic.generate(synthetic.spy_day(seed=1, n_msgs=20000, duration_s=60))
cfg = abi.default_cfg(n_envs=4, n_levels=10, outer_levels=20, max_levels_per_side=64, max_orders_per_side=256, max_agent_orders=32)
sim = LobSim(cfg, 0); sim.load_stream(0, s)
sim.reset_book(0, np.array([0, 10, 20, 30], np.int32)); torch.cuda.synchronize(); print("reset ok", flush=True)
for n in (0, 1, 5, 50, 200):
    sim.replay(n); torch.cuda.synchronize(); print("replay", n, "ok", sim.state()["now_step"], sim.state()["err"], flush=True)
# parity of the restructured kernel against the oracle on the same tiny stream (includes resync bail-outs)
from oracle.oracle import Oracle
o = [Oracle(abi.default_cfg(n_levels=10, outer_levels=20), s) for _ in range(4)]
for k, st in enumerate([0, 10, 20, 30]):
    o[k].reset_book(st); o[k].replay(256)
    a, b_ = sim.dump_book(k, 0), o[k].dump_book(0)
    assert len(a) == len(b_) and np.array_equal(a["price"], b_["price"]) and np.array_equal(a["volume"], b_["volume"]), k
print("parity ok", flush=True)
