#!/usr/bin/env python
"""Packing throughput for a 1e7-row LOBSTER day: writes a synthetic LOBSTER message / orderbook file pair (the SPY-shaped
day of BASELINE configs[1], 10 levels) to a scratch directory, then packs it with the native packer (csrc/lobster_ingest.cpp), with
the C++ reader + numpy packer, and checks that both give the same stream.

    python tools/ingest_rate.py [n_rows] [scratch_dir]
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rl4mm_b200 import abi, synthetic  # noqa: E402
from rl4mm_b200.packing import pack_lobster  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
out = Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp/lobster_day")
out.mkdir(parents=True, exist_ok=True)
L = 10
s = synthetic.generate(synthetic.spy_day(seed=0, n_msgs=n, duration_s=23_400 if n >= 5_000_000 else 2_340))
step = np.repeat(np.arange(s.n_grid_steps), np.diff(s.step_off.astype(np.int64)))
us = s.t0_us + step * s.step_us + 1 + (np.arange(n) - s.step_off.astype(np.int64)[step])
ty = (s.msgs["meta"] & 7).astype(np.int64)
side = ((s.msgs["meta"] >> 3) & 1).astype(np.int64)
di = np.where(np.where(ty == 4, 1 - side, side) == 0, 1, -1)
msg, book = out / f"SYN_message_{L}.csv", out / f"SYN_orderbook_{L}.csv"
t0 = time.perf_counter()
with open(msg, "w") as f:
    for a in range(0, n, 1_000_000):
        b = min(n, a + 1_000_000)
        f.write("".join(f"{u // 1_000_000}.{u % 1_000_000:06d}000,{t},{r},{v},{p},{d}\n" for u, t, r, v, p, d in
                        zip(us[a:b].tolist(), ty[a:b].tolist(), s.msgs["ref"][a:b].tolist(), s.msgs["volume"][a:b].tolist(),
                            s.msgs["price"][a:b].tolist(), di[a:b].tolist())))
snap = s.snapshots.astype(np.int64)
rows = np.zeros((snap.shape[0], 4 * L), np.int64)
for lv in range(L):
    ap, av, bp, bv = snap[:, 1, lv, 0], snap[:, 1, lv, 1], snap[:, 0, lv, 0], snap[:, 0, lv, 1]
    rows[:, 4 * lv + 0] = np.where(ap == abi.NO_PRICE, 9999999999, ap)
    rows[:, 4 * lv + 1] = av
    rows[:, 4 * lv + 2] = np.where(bp == abi.NO_PRICE, -9999999999, bp)
    rows[:, 4 * lv + 3] = bv
lines = [",".join(map(str, r)) + "\n" for r in rows.tolist()]
sec_of = ((us - s.t0_us) // 1_000_000).tolist()
with open(book, "w") as f:
    for a in range(0, n, 1_000_000):
        f.write("".join(lines[k] for k in sec_of[a:a + 1_000_000]))
print(f"wrote {msg.stat().st_size / 1e6:.0f} MB + {book.stat().st_size / 1e6:.0f} MB in {time.perf_counter() - t0:.0f} s")
res = {}
for name, fast in (("native packer (lobingest_pack_open)", True), ("C++ reader + numpy packer", "reader")):
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        res[name] = pack_lobster(msg, book, L, fast=fast, t0_us=s.t0_us)
        best = min(best, time.perf_counter() - t0)
    print(f"{name}: {best:.2f} s = {n / best:.3g} rows/s")
a, b = res.values()
for f_ in ("msgs", "step_off", "snapshots", "snap_valid", "ext_ids"):
    assert np.array_equal(getattr(a, f_), getattr(b, f_)), f_
assert np.array_equal(a.step_off, s.step_off) and np.array_equal(a.msgs["price"], s.msgs["price"])
print("identical streams; round trip ok")
