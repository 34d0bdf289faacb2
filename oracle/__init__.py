"""TEST INFRASTRUCTURE: the CPU parity oracle (C restatement of the reference + shims to run the reference itself).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``rl4mm_b200``) never does.
"""
