"""LOBSTER message / orderbook files -> packed, device-ready stream buffers.

This replaces ``rl4mm/database`` for the hot path: instead of loading the
CSVs into Postgres and querying a range per episode / per step, a ticker-day is packed ONCE into

* ``msgs``      16-byte records in replay order (hidden executions dropped, execution direction flipped),
* ``step_off``  CSR offsets of the messages of every ``step_us`` grid interval ``(t0 + k*step, t0 + (k+1)*step]``,
* ``snapshots`` the L-level book at every whole second (``get_last_snapshot`` with ``book_snapshot_freq="S"``),

and kept resident in HBM.  The semantics restated here, with the reference lines they follow:

* type map 1 limit / 2 cancellation / 3 deletion / 4 market / 5 market_hidden / 6 cross_trade / 7 trading_halt --
  rl4mm/database/database_population_helpers.py:151-160;
* direction: +1 buy / -1 sell, flipped for executions -- :132-136;
* timestamps are SQL ``DateTime`` => truncated to microseconds -- rl4mm/database/models.py:14;
* range query ``start < ts <= end`` ordered by ``(timestamp, id)`` where ``id`` is the STRING
  ``"{freq}_L{levels}_NASDAQ_{ticker}_{date}_{row}"`` => same-microsecond ties are ordered lexicographically by the
  decimal row number -- rl4mm/database/HistoricalDatabase.py:103-119, database_population_helpers.py:163-181;
* hidden executions dropped, cross trades rejected -- rl4mm/simulation/HistoricalOrderGenerator.py:49-57;
* snapshot at second T = orderbook row of the last message with time <= T --
  database_population_helpers.py:45-62,139-148, HistoricalDatabase.py:46-62.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi

LOBSTER_DUMMY = 9_999_999_999


@dataclass
class PackedStream:
    msgs: np.ndarray        # abi.MSG_DTYPE [n_msgs]
    step_off: np.ndarray    # uint32 [n_grid_steps + 1]
    snapshots: np.ndarray   # int32 [n_seconds + 1, 2, n_levels, 2]  (side, level, (price, volume))
    snap_valid: np.ndarray  # uint8 [n_seconds + 1]
    t0_us: int              # grid origin, microseconds after midnight (whole second)
    step_us: int
    n_levels: int
    ext_ids: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int64))  # ref -> original external id
    ticker: str = ""
    date: str = ""
    generators: tuple = ("historical",)   # names of the OrderGenerators merged into this stream (pack_merged)

    @property
    def n_msgs(self) -> int:
        return len(self.msgs)

    @property
    def n_grid_steps(self) -> int:
        return len(self.step_off) - 1

    @property
    def n_seconds(self) -> int:
        return len(self.snap_valid) - 1

    @property
    def steps_per_second(self) -> int:
        return 1_000_000 // self.step_us

    def step_of_time(self, seconds_after_midnight: float) -> int:
        """Grid step index whose boundary is the given clock time (must lie on the grid)."""
        us = int(round(seconds_after_midnight * 1_000_000)) - self.t0_us
        if us % self.step_us:
            raise ValueError("time is not on the step grid")
        return us // self.step_us

    def validate(self) -> None:
        assert self.msgs.dtype == abi.MSG_DTYPE and self.msgs.flags.c_contiguous
        assert self.step_off.dtype == np.uint32 and self.step_off[0] == 0 and self.step_off[-1] == len(self.msgs)
        assert np.all(np.diff(self.step_off.astype(np.int64)) >= 0)
        assert self.snapshots.dtype == np.int32 and self.snapshots.shape == (self.n_seconds + 1, 2, self.n_levels, 2)
        assert self.t0_us % 1_000_000 == 0 and 1_000_000 % self.step_us == 0
        assert self.n_seconds * self.steps_per_second >= self.n_grid_steps


def parse_time_ns(text: str) -> int:
    """Exact decimal parse of a LOBSTER ``seconds.after.midnight`` field into integer nanoseconds."""
    if "." in text:
        sec, frac = text.split(".")
    else:
        sec, frac = text, ""
    return int(sec) * 1_000_000_000 + int((frac + "000000000")[:9])


def _lexicographic_tie_order(ts_us: np.ndarray, row_ids: np.ndarray) -> np.ndarray:
    """Permutation that sorts by (ts_us, str(row_id)) -- the reference's ``ORDER BY timestamp, id`` on a string id."""
    order = np.argsort(ts_us, kind="stable")
    t = ts_us[order]
    run_start = np.flatnonzero(np.r_[True, t[1:] != t[:-1]])
    run_end = np.r_[run_start[1:], len(t)]
    for a, b in zip(run_start[run_end - run_start > 1], run_end[run_end - run_start > 1]):
        seg = order[a:b]
        order[a:b] = seg[np.argsort(np.array([str(int(r)) for r in row_ids[seg]]), kind="stable")]
    return order


@dataclass
class RawMessages:
    """The messages of ONE OrderGenerator (rl4mm/simulation/OrderGenerator.py:9-21) as LOBSTER columns."""
    time_ns: np.ndarray
    msg_type: np.ndarray
    ext_id: np.ndarray
    size: np.ndarray
    price: np.ndarray
    direction: np.ndarray

    def take(self, idx) -> "RawMessages":
        return RawMessages(*(np.asarray(getattr(self, f))[idx] for f in ("time_ns", "msg_type", "ext_id", "size", "price", "direction")))

    def __len__(self):
        return len(self.time_ns)


class _MergeOrder:
    """What ``min(order_dict, key=order_dict.get)`` sees of an Order (rl4mm/orderbook/models.py:17-52): ``==`` is the dataclass
    equality over every field (same class, i.e. same message type; a market order has no price), ``<`` is
    ``Order.__lt__`` = the timestamp alone (:27-28)."""
    __slots__ = ("ts", "fields", "idx")

    def __init__(self, ts, fields, idx):
        self.ts, self.fields, self.idx = ts, fields, idx

    def __eq__(self, other):
        return self.fields == other.fields

    def __lt__(self, other):
        return self.ts < other.ts


def compress_order_dict(order_dict):
    """``OrderbookSimulator._compress_order_dict`` (rl4mm/simulation/OrderbookSimulator.py:137-148): the orders of several
    generators for one step, merged by repeatedly taking the head of the SMALLEST deque -- deques compare
    lexicographically (first unequal pair decides by ``<`` = timestamp; equal prefixes => the shorter one), and ``min``
    keeps the first generator of the dict on a tie.  Works on deques of anything with that ``==`` / ``<`` (the façade's
    Order dataclasses, or :class:`_MergeOrder`).  Consumes ``order_dict``."""
    if len(order_dict) == 1:
        return list(next(iter(order_dict.values())))
    merged = []
    while order_dict:
        key = min(order_dict, key=order_dict.get)
        merged.append(order_dict[key].popleft())
        if not order_dict[key]:
            del order_dict[key]
    return merged


def generator_order(raw: RawMessages, t0_us: int, tie_order: str = "reference", db_batch_size: int = 1_000_000) -> RawMessages:
    """One generator's messages in the order ``generate_orders`` hands them out (HistoricalOrderGenerator.py:32-57):
    ``ORDER BY (timestamp, id-string)``, hidden executions dropped, nothing at or before the grid origin."""
    time_ns, mt = np.asarray(raw.time_ns, np.int64), np.asarray(raw.msg_type, np.int64)
    ts_us = time_ns // 1000
    row_ids = np.arange(len(raw), dtype=np.int64)
    row_ids = row_ids + (row_ids // db_batch_size) * db_batch_size
    order = _lexicographic_tie_order(ts_us, row_ids) if tie_order == "reference" else np.argsort(ts_us, kind="stable")
    keep = order[(ts_us[order] > t0_us) & (mt[order] != 5)]
    return raw.take(keep)


def compress_order_sources(sources, t0_us: int, step_us: int) -> RawMessages:
    """Pack-time form of ``forward_step``'s merge (OrderbookSimulator.py:74-75,137-148): ``sources`` maps generator name ->
    :class:`RawMessages` already in generator order (:func:`generator_order`); every step window
    ``(t0 + k*step, t0 + (k+1)*step]`` is merged with :func:`compress_order_dict`, exactly what the reference does when it
    pulls that window from each generator.  Returns the merged messages (time truncated to microseconds)."""
    from collections import deque

    names = list(sources)
    srcs = [sources[k] for k in names]
    ts = [np.asarray(s.time_ns, np.int64) // 1000 for s in srcs]
    if len(srcs) == 1:
        out = srcs[0].take(slice(None))
        out.time_ns = ts[0] * 1000
        return out
    steps = [-(-(t - t0_us) // step_us) - 1 for t in ts]
    for st in steps:
        assert np.all(np.diff(st) >= 0), "a generator's messages must be time ordered"
    n_steps = max(int(st[-1]) + 1 if len(st) else 0 for st in steps)
    bounds = [np.searchsorted(st, np.arange(n_steps + 1)) for st in steps]
    cols = [[np.asarray(getattr(s, f), np.int64) for f in ("msg_type", "ext_id", "size", "price", "direction")] for s in srcs]
    picks = []                                       # (source, index) in merged order, as array chunks
    for k in range(n_steps):
        live = [g for g in range(len(srcs)) if bounds[g][k + 1] > bounds[g][k]]
        if not live:
            continue
        if len(live) == 1:
            g = live[0]
            idx = np.arange(bounds[g][k], bounds[g][k + 1])
            picks.append(np.stack([np.full(len(idx), g), idx], 1))
            continue
        od = {}
        for g in live:
            mt, eid, sz, pr, di = cols[g]
            od[names[g]] = deque(
                _MergeOrder(int(ts[g][i]), (int(ts[g][i]), int(mt[i]), int(di[i]), int(eid[i]), int(sz[i]), None if mt[i] == 4 else int(pr[i])), (g, i))
                for i in range(bounds[g][k], bounds[g][k + 1]))
        picks.append(np.array([o.idx for o in compress_order_dict(od)], np.int64).reshape(-1, 2))
    pick = np.concatenate(picks) if picks else np.zeros((0, 2), np.int64)
    out = {}
    for f in ("time_ns", "msg_type", "ext_id", "size", "price", "direction"):
        col = np.zeros(len(pick), np.int64)
        for g, s in enumerate(srcs):
            m = pick[:, 0] == g
            src = ts[g] * 1000 if f == "time_ns" else np.asarray(getattr(s, f), np.int64)
            col[m] = src[pick[m, 1]]
        out[f] = col
    return RawMessages(**out)


def pack_merged(sources, book_rows, n_levels: int, step_us: int = 100_000, t0_us: Optional[int] = None, tie_order: str = "reference",
                db_batch_size: int = 1_000_000, snapshot_source: Optional[str] = None, **kw) -> PackedStream:
    """Several OrderGenerators -> ONE packed stream (the reference's ``order_generators`` list, OrderbookSimulator.py:24-53):
    each source is put in generator order, the sources are merged step by step with the reference's comparison, and the
    result is packed as is.  ``book_rows`` (and the per-second snapshots) belong to ``snapshot_source`` (default: the first
    source, the historical one -- snapshots come from the database, not from the generators)."""
    names = list(sources)
    snapshot_source = snapshot_source or names[0]
    first = min(int(np.asarray(s.time_ns)[0]) for s in sources.values() if len(s))
    if t0_us is None:
        t0_us = first // 1000 // 1_000_000 * 1_000_000
    ordered = {k: generator_order(v, t0_us, tie_order, db_batch_size) for k, v in sources.items()}
    merged = compress_order_sources(ordered, t0_us, step_us)
    out = pack_arrays(merged.time_ns, merged.msg_type, merged.ext_id, merged.size, merged.price, merged.direction, book_rows, n_levels,
                      step_us=step_us, t0_us=t0_us, tie_order="file", snapshot_time_ns=np.asarray(sources[snapshot_source].time_ns, np.int64), **kw)
    out.generators = tuple(names)
    return out


def pack_arrays(
    time_ns: np.ndarray,
    msg_type: np.ndarray,
    ext_id: np.ndarray,
    size: np.ndarray,
    price: np.ndarray,
    direction: np.ndarray,
    book_rows,
    n_levels: int,
    step_us: int = 100_000,
    t0_us: Optional[int] = None,
    tie_order: str = "reference",
    db_batch_size: int = 1_000_000,
    ticker: str = "",
    date: str = "",
    snapshot_time_ns: Optional[np.ndarray] = None,
) -> PackedStream:
    """Pack raw LOBSTER columns.  ``book_rows`` is the orderbook file as int64 [n_rows, 4*n_levels] (one row per
    message row, columns ask price, ask size, bid price, bid size per level -- rl4mm/orderbook/helpers.py:52-55), or a
    callable ``rows(idx) -> int64 [len(idx), 4*n_levels]`` that loads just the requested (ascending) row indices.
    ``snapshot_time_ns``: the time column the orderbook rows are aligned with when it is not ``time_ns`` (merged streams:
    the snapshots follow the historical source, :func:`pack_merged`)."""
    time_ns = np.asarray(time_ns, np.int64)
    msg_type = np.asarray(msg_type, np.int64)
    n = len(time_ns)
    assert np.all(np.diff(time_ns) >= 0), "LOBSTER messages must be time ordered"
    snap_time = time_ns if snapshot_time_ns is None else np.asarray(snapshot_time_ns, np.int64)
    assert callable(book_rows) or book_rows.shape == (len(snap_time), 4 * n_levels)
    if 1_000_000 % step_us:
        raise ValueError("step_us must divide one second")
    ts_us = time_ns // 1000
    if t0_us is None:
        t0_us = int(ts_us[0] // 1_000_000) * 1_000_000
    if t0_us % 1_000_000:
        raise ValueError("t0_us must be a whole second")
    last_us = int(ts_us[-1])
    n_grid = max(1, -(-(last_us - t0_us) // step_us))
    n_seconds = -(-(n_grid * step_us) // 1_000_000)
    n_grid = n_seconds * (1_000_000 // step_us)

    # --- replay order ------------------------------------------------------------------------------------------
    row_ids = np.arange(n, dtype=np.int64)
    row_ids = row_ids + (row_ids // db_batch_size) * db_batch_size  # message.name + start_index (:171)
    if tie_order == "reference":
        order = _lexicographic_tie_order(ts_us, row_ids)
    elif tie_order == "file":
        order = np.arange(n)
    else:
        raise ValueError(tie_order)
    keep = order[(ts_us[order] > t0_us) & (msg_type[order] != 5)]
    bad = np.isin(msg_type[keep], (6, 7))
    if bad.any():
        raise ValueError("cross_trade / trading_halt messages inside the replay range (the reference asserts / crashes: "
                         "HistoricalOrderGenerator.py:52-56, create_order.py:9-24)")
    if not np.isin(msg_type[keep], (1, 2, 3, 4)).all():
        raise ValueError("unknown LOBSTER message type")
    mt = msg_type[keep]
    d = np.asarray(direction, np.int64)[keep]
    is_exec = mt == 4
    side = np.where(d == 1, abi.BUY, abi.SELL)
    side = np.where(is_exec, 1 - side, side)  # executions carry the aggressor's direction
    ids = np.asarray(ext_id, np.int64)[keep]
    uniq, inv = np.unique(ids, return_inverse=True)
    if len(uniq) + 1 >= 2**31:
        raise ValueError("too many distinct order ids")
    msgs = np.zeros(len(keep), abi.MSG_DTYPE)
    pr = np.asarray(price, np.int64)[keep]
    sz = np.asarray(size, np.int64)[keep]
    if pr.max(initial=0) >= 2**31 or sz.max(initial=0) >= 2**31:
        raise ValueError("price / size does not fit int32")
    msgs["price"] = pr
    msgs["volume"] = sz
    msgs["ref"] = inv + 1
    msgs["meta"] = mt.astype(np.uint32) | (side.astype(np.uint32) << 3)
    step_idx = -(-(ts_us[keep] - t0_us) // step_us) - 1  # ts in (t0 + k*step, t0 + (k+1)*step]  =>  k
    assert step_idx.min(initial=0) >= 0 and step_idx.max(initial=0) < n_grid
    step_off = np.zeros(n_grid + 1, np.uint32)
    step_off[1:] = np.cumsum(np.bincount(step_idx, minlength=n_grid))

    # --- per-second snapshots ----------------------------------------------------------------------------------
    sec_ns = (t0_us + np.arange(n_seconds + 1, dtype=np.int64) * 1_000_000) * 1000
    idx = np.searchsorted(snap_time, sec_ns, side="right") - 1
    snap_valid = (idx >= 0).astype(np.uint8)
    need = np.maximum(idx, 0)
    rows = book_rows(need) if callable(book_rows) else np.asarray(book_rows, np.int64)[need]
    rows = np.asarray(rows, np.int64).reshape(n_seconds + 1, n_levels, 4)
    snapshots = np.zeros((n_seconds + 1, 2, n_levels, 2), np.int32)
    for s, (pc, vc) in ((abi.SELL, (0, 1)), (abi.BUY, (2, 3))):
        p, v = rows[:, :, pc], rows[:, :, vc]
        dummy = (np.abs(p) >= LOBSTER_DUMMY)
        snapshots[:, s, :, 0] = np.where(dummy, abi.NO_PRICE, p)
        snapshots[:, s, :, 1] = np.where(dummy, 0, v)
    out = PackedStream(msgs, step_off, snapshots, snap_valid, int(t0_us), int(step_us), n_levels,
                       np.r_[np.int64(0), uniq], ticker, date)
    out.validate()
    return out


_INGEST = None


def _ingest_lib():
    global _INGEST
    if _INGEST is None:
        import ctypes as C

        from .build import build_ingest

        L = C.CDLL(str(build_ingest()))
        L.lobingest_count_lines.restype = C.c_int64
        L.lobingest_count_lines.argtypes = [C.c_char_p]
        L.lobingest_parse_messages.argtypes = [C.c_char_p, C.c_int64] + [C.c_void_p] * 6 + [C.POINTER(C.c_int64)]
        L.lobingest_parse_book_rows.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.lobingest_count_rows.restype = C.c_int64
        L.lobingest_count_rows.argtypes = [C.c_char_p]
        L.lobingest_pack_open.restype = C.c_void_p
        L.lobingest_pack_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                          C.POINTER(C.c_int32), C.c_char_p, C.c_int32]
        L.lobingest_pack_sizes.argtypes = [C.c_void_p, C.c_void_p]
        L.lobingest_pack_copy.argtypes = [C.c_void_p] * 6
        L.lobingest_pack_close.argtypes = [C.c_void_p]
        _INGEST = L
    return _INGEST


def read_lobster_messages(message_csv, max_rows: Optional[int] = None):
    """Columns of a LOBSTER message file through csrc/lobster_ingest.cpp: (time_ns, type, order_id, size, price,
    direction) as numpy arrays; the time is parsed exactly from the decimal text."""
    import ctypes as C

    L = _ingest_lib()
    path = str(message_csv).encode()
    n = L.lobingest_count_rows(path)
    if n < 0:
        raise FileNotFoundError(message_csv)
    if max_rows is not None:
        n = min(n, max_rows)
    t, oid, sz, pr = (np.zeros(n, np.int64) for _ in range(4))
    ty, di = np.zeros(n, np.int32), np.zeros(n, np.int32)
    got = C.c_int64(0)
    rc = L.lobingest_parse_messages(path, n, t.ctypes.data, ty.ctypes.data, oid.ctypes.data, sz.ctypes.data,
                                    pr.ctypes.data, di.ctypes.data, C.byref(got))
    if rc != 0:
        raise ValueError(f"malformed LOBSTER message file {message_csv} (rc {rc})")
    k = got.value
    return t[:k], ty[:k], oid[:k], sz[:k], pr[:k], di[:k]


def read_lobster_book_rows(orderbook_csv, row_idx: np.ndarray, n_levels: int) -> np.ndarray:
    """Only the requested rows (ascending indices) of a LOBSTER orderbook file."""
    L = _ingest_lib()
    idx = np.ascontiguousarray(row_idx, np.int64)
    assert np.all(np.diff(idx) >= 0)
    out = np.zeros((len(idx), 4 * n_levels), np.int64)
    rc = L.lobingest_parse_book_rows(str(orderbook_csv).encode(), idx.ctypes.data, len(idx), 4 * n_levels, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"could not read the requested rows of {orderbook_csv} (rc {rc})")
    return out


def pack_lobster_native(message_csv, orderbook_csv, n_levels: int, max_rows: Optional[int] = None, step_us: int = 100_000,
                        t0_us: Optional[int] = None, tie_order: str = "reference", db_batch_size: int = 1_000_000,
                        ticker: str = "", date: str = "") -> PackedStream:
    """The whole packer in C++ (csrc/lobster_ingest.cpp::lobingest_pack_open): CSV parse, type map, direction flip,
    microsecond truncation, the reference's lexicographic tie order, hidden-execution drop, dense order references,
    CSR step offsets and the per-second snapshot alignment -- bit-identical to :func:`pack_arrays`."""
    import ctypes as C

    if tie_order not in ("reference", "file"):
        raise ValueError(tie_order)
    L = _ingest_lib()
    rc, err = C.c_int32(0), C.create_string_buffer(256)
    h = L.lobingest_pack_open(str(message_csv).encode(), str(orderbook_csv).encode(), n_levels, step_us,
                              -1 if t0_us is None else int(t0_us), int(tie_order == "reference"), db_batch_size,
                              -1 if max_rows is None else int(max_rows), C.byref(rc), err, 256)
    if not h:
        msg = err.value.decode() or f"rc {rc.value}"
        raise (FileNotFoundError if rc.value == -1 else ValueError)(f"{message_csv}: {msg}")
    try:
        sizes = np.zeros(6, np.int64)
        L.lobingest_pack_sizes(h, sizes.ctypes.data)
        n_msgs, n_grid, n_sec, n_ids, t0, _ = (int(x) for x in sizes)
        msgs = np.zeros(n_msgs, abi.MSG_DTYPE)
        step_off = np.zeros(n_grid + 1, np.uint32)
        snapshots = np.zeros((n_sec + 1, 2, n_levels, 2), np.int32)
        snap_valid = np.zeros(n_sec + 1, np.uint8)
        ext_ids = np.zeros(n_ids, np.int64)
        L.lobingest_pack_copy(h, msgs.ctypes.data, step_off.ctypes.data, snapshots.ctypes.data, snap_valid.ctypes.data, ext_ids.ctypes.data)
    finally:
        L.lobingest_pack_close(h)
    out = PackedStream(msgs, step_off, snapshots, snap_valid, t0, int(step_us), n_levels, ext_ids, ticker, date)
    out.validate()
    return out


def pack_lobster(message_csv, orderbook_csv, n_levels: int, max_rows: Optional[int] = None, fast=True, **kw) -> PackedStream:
    """Pack a LOBSTER ``*_message_L.csv`` / ``*_orderbook_L.csv`` pair (populate_database.py:71-78 column layout).
    ``fast=True``: the native packer (:func:`pack_lobster_native`); ``fast="reader"``: the C++ CSV reader (only the
    orderbook rows the snapshots need) + the numpy packer :func:`pack_arrays`; ``fast=False``: the pure-Python path.
    All three give bit-identical streams (tests/test_abi_cpu.py)."""
    if fast is True:
        return pack_lobster_native(message_csv, orderbook_csv, n_levels, max_rows, **kw)
    if fast:
        t, ty, oid, sz, pr, di = read_lobster_messages(message_csv, max_rows)
        if max_rows is None and _ingest_lib().lobingest_count_rows(str(orderbook_csv).encode()) != len(t):
            raise ValueError("message file and orderbook file have different row counts")
        return pack_arrays(t, ty, oid, sz, pr, di, lambda idx: read_lobster_book_rows(orderbook_csv, idx, n_levels),
                           n_levels, **kw)
    t, ty, oid, sz, pr, di = [], [], [], [], [], []
    with open(message_csv) as f:
        for i, line in enumerate(f):
            if max_rows is not None and i >= max_rows:
                break
            p = line.rstrip("\n").split(",")
            t.append(parse_time_ns(p[0]))
            ty.append(int(p[1]))
            oid.append(int(p[2]))
            sz.append(int(p[3]))
            pr.append(int(p[4]))
            di.append(int(p[5]))
    books = np.loadtxt(orderbook_csv, delimiter=",", dtype=np.int64, max_rows=len(t), ndmin=2)
    return pack_arrays(np.array(t), np.array(ty), np.array(oid), np.array(sz), np.array(pr), np.array(di), books,
                       n_levels, **kw)
