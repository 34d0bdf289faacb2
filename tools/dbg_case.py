"""Debug helper: run one random env case on the device and the oracle, print the state around the second reset."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import random_cases as R
from oracle.oracle import Oracle
from rl4mm_b200 import abi, synthetic
from rl4mm_b200.device import LobSim
seed, env = int(sys.argv[1]), int(sys.argv[2])
c = R.random_case(seed); s = synthetic.generate(c["synth"]); n = c["n_envs"]; ep = c["cfg_kw"]["episode_steps"]
sim = LobSim(abi.default_cfg(n_envs=n, **c["cfg_kw"]), 0); sim.load_stream(0, s)
o = Oracle(abi.default_cfg(n_envs=1, **c["cfg_kw"]), s)
ext = c["agent_kind"] == "external"
for part in range(2):
    starts = (c["starts"] + part * 10).astype(np.int32)
    obs0 = sim.reset(0, starts).cpu().numpy(); oo0 = o.reset(int(starts[env]))
    st = sim.state(); os_ = o.state()
    print("part", part, "reset: dev inv", int(st["inventory"][env]), "oracle inv", int(os_["inventory"]), "dev nag", st["n_agent_orders"][env], "oracle nag", os_["n_agent_orders"], "err", int(st["err"][env]), int(os_["err"]))
    print("  dev obs", obs0[env]); print("  orc obs", oo0)
    for side in (0, 1):
        d = sim.dump_book(env, side); print("  side", side, "dev agent-tagged entries in book:", int(((d["ref"] & abi.REF_AGENT) != 0).sum()), "orc:", int(((o.dump_book(side)["ref"] & abi.REF_AGENT) != 0).sum()))
    acts = c["actions"][part * ep:(part + 1) * ep]
    k = ep // 2
    if ext:
        for t in range(k): sim.step(torch.tensor(acts[t], device="cuda"))
        sim.rollout(ep - k, c["agent"], torch.tensor(acts[k:], device="cuda"))
    else:
        sim.rollout(k, c["agent"]); sim.rollout(ep - k, c["agent"])
    o.rollout(ep, c["agent"], acts[:, env] if ext else None)
    st = sim.state(); os_ = o.state()
    print("  end: dev inv", int(st["inventory"][env]), "oracle inv", int(os_["inventory"]), "dev nag", st["n_agent_orders"][env], "oracle nag", os_["n_agent_orders"], "dev next_agent_id", int(st["next_agent_id"][env]))
