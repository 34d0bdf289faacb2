"""Drop-in mirror of ``rl4mm/orderbook`` (models.py, Exchange.py, OrderIDConvertor.py, create_order.py) on top of
the device book.

Same class names, constructor arguments, return values and exceptions as the reference, so code and tests written
against ``rl4mm.orderbook.Exchange`` run against this one.  The book itself lives on the GPU (one env of a
:class:`rl4mm_b200.device.LobSim`); ``central_orderbook`` / ``internal_orderbook`` are materialised from an L3 dump on
access.  This single-book view is the compatibility layer -- the hot path is the batched kernel.
"""
from __future__ import annotations

import warnings
from collections import deque
from copy import copy
from dataclasses import dataclass, field
from datetime import datetime
from typing import Dict, List, Literal, Optional, Union

import numpy as np
from sortedcontainers import SortedDict

from . import abi

# ---- rl4mm/orderbook/models.py ---------------------------------------------------------------------------------------


@dataclass
class Order:
    timestamp: datetime
    direction: Literal["buy", "sell"]
    ticker: str
    internal_id: Optional[int]
    external_id: Optional[int]
    is_external: bool

    def __lt__(self, other):
        return self.timestamp < other.timestamp


@dataclass
class MarketOrder(Order):
    volume: int
    price: Optional[int] = None


@dataclass
class LimitOrder(Order):
    price: int
    volume: int


@dataclass
class Deletion(Order):
    price: int
    volume: Optional[int]


@dataclass
class Cancellation(Deletion):
    volume: int


FillableOrder = Union[MarketOrder, LimitOrder]


@dataclass
class FilledOrders:
    internal: List[FillableOrder] = field(default_factory=list)
    external: List[FillableOrder] = field(default_factory=list)


@dataclass
class Orderbook:
    buy: SortedDict
    sell: SortedDict
    ticker: str
    tick_size: int

    @property
    def best_buy_price(self):
        return next(reversed(self.buy), 0)

    @property
    def best_sell_price(self):
        return next(iter(self.sell.keys()), np.inf)

    @property
    def best_buy_volume(self):
        return sum(order.volume for order in self.buy[self.best_buy_price])

    @property
    def best_sell_volume(self):
        return sum(order.volume for order in self.sell[self.best_sell_price])

    @property
    def midprice(self):
        return (self.best_sell_price + self.best_buy_price) / 2

    @property
    def imbalance(self):
        return (self.best_buy_volume - self.best_sell_volume) / (self.best_buy_volume + self.best_sell_volume)

    @property
    def microprice(self):
        return (1 + self.imbalance) / 2 * self.best_sell_price + (1 - self.imbalance) / 2 * self.best_buy_price

    @property
    def spread(self):
        return self.best_sell_price - self.best_buy_price


OrderDict = dict


def create_order(order_type: str, order_dict: dict):
    """rl4mm/orderbook/create_order.py:4-42."""
    d = dict(order_dict)
    if order_type == "market":
        d.pop("price", None)
        return MarketOrder(**d)
    if order_type == "limit":
        return LimitOrder(**d)
    if order_type == "cancellation":
        return Cancellation(**d)
    if order_type == "deletion":
        return Deletion(**d)
    raise NotImplementedError(order_type)


class EmptyOrderbookError(Exception):
    pass


class CancellationVolumeExceededError(Exception):
    pass


class OrderIdConvertor:
    """rl4mm/orderbook/OrderIDConvertor.py:7-37 -- host-side shadow of the ids the reference hands out."""

    def __init__(self):
        self.external_to_internal_lookup: Dict[int, int] = dict()
        self.counter = 0

    def get_internal_order_id(self, order: Order) -> Optional[int]:
        if not order.is_external:
            return order.internal_id
        return self.external_to_internal_lookup.get(order.external_id)

    def remove_external_order_id(self, external_id: int) -> None:
        self.external_to_internal_lookup.pop(external_id, None)

    def reset(self):
        self.external_to_internal_lookup = dict()
        self.counter = 0


_SIDE = {"buy": abi.BUY, "sell": abi.SELL}
_DIR = ("buy", "sell")


class Exchange:
    """rl4mm/orderbook/Exchange.py:40-247 backed by one device book."""

    def __init__(self, ticker: str = "MSFT", central_orderbook: Orderbook = None, internal_orderbook: Orderbook = None,
                 tick_size: int = 100, *, sim=None, env: int = 0, device: int = 0, **capacity):
        self.ticker = ticker
        self.tick_size = tick_size
        self.name = "NASDAQ"
        self.order_id_convertor = OrderIdConvertor()
        if sim is None:
            from .device import LobSim

            cfg = abi.default_cfg(n_envs=1, tick_size=tick_size, fill_log_capacity=0, **capacity)
            sim = LobSim(cfg, device)
        self.sim, self.env = sim, env
        self._ext_ref: Dict[int, int] = {}        # external_id -> dense device ref
        self._ref_ext: Dict[int, int] = {}
        self._iid_of_ref: Dict[int, int] = {}     # device ref (agent refs carry LOBSIM_REF_AGENT) -> internal_id
        self._agent_of_iid: Dict[int, int] = {}   # internal_id -> device agent id
        self._meta: Dict[int, tuple] = {}         # device ref -> (timestamp,)
        self._stream_ext_ids = np.zeros(1, np.int64)  # dense refs 1..n-1 of the loaded stream (packing.ext_ids)
        for book in (central_orderbook, internal_orderbook):
            if book is not None:
                assert book.ticker == self.ticker, "Orderbook ticker must agree with the exchange ticker."
        if central_orderbook is not None:
            self.central_orderbook = central_orderbook

    # ---- id plumbing -------------------------------------------------------------------------------------------------
    def use_stream_ids(self, ext_ids: np.ndarray) -> None:
        """Share the dense reference numbering of a packed stream (PackedStream.ext_ids, sorted ascending)."""
        self._stream_ext_ids = np.asarray(ext_ids, np.int64)

    def _lookup_ref(self, external_id) -> Optional[int]:
        ids = self._stream_ext_ids
        if len(ids) > 1:
            i = int(np.searchsorted(ids[1:], external_id)) + 1
            if i < len(ids) and ids[i] == external_id:
                return i
        return self._ext_ref.get(external_id)

    def _ref_for_external(self, external_id) -> int:
        if external_id is None:
            return 0
        r = self._lookup_ref(external_id)
        if r is None:
            r = len(self._stream_ext_ids) + len(self._ext_ref)
            self._ext_ref[external_id] = r
            self._ref_ext[r] = external_id
        return r

    def _ext_of_ref(self, ref: int):
        if 0 < ref < len(self._stream_ext_ids):
            return int(self._stream_ext_ids[ref])
        return self._ref_ext.get(ref, ref)

    # ---- books -------------------------------------------------------------------------------------------------------
    def get_empty_orderbook(self) -> Orderbook:
        return Orderbook(buy=SortedDict(), sell=SortedDict(), ticker=self.ticker, tick_size=self.tick_size)

    def _view(self, entries_by_side) -> Orderbook:
        book = self.get_empty_orderbook()
        for side, entries in enumerate(entries_by_side):
            half = book.sell if side else book.buy
            for e in entries:
                ref = int(e["ref"])
                is_agent = bool(ref & abi.REF_AGENT)
                if ref == abi.REF_AGGREGATE:
                    iid, ext = -1, None
                elif is_agent:
                    iid, ext = self._iid_of_ref.get(ref, ref & 0x7FFFFFFF), None
                else:
                    iid, ext = self._iid_of_ref.get(ref), self._ext_of_ref(ref)
                ts = self._meta.get(ref, (None,))[0] if ref else self._meta.get(("agg", side, int(e["price"])), (None,))[0]
                order = LimitOrder(timestamp=ts, direction=_DIR[side], ticker=self.ticker, internal_id=iid,
                                   external_id=ext, is_external=not is_agent, price=int(e["price"]), volume=int(e["volume"]))
                half.setdefault(order.price, deque()).append(order)
        return book

    @property
    def central_orderbook(self) -> Orderbook:
        return self._view([self.sim.dump_book(self.env, s) for s in (0, 1)])

    @central_orderbook.setter
    def central_orderbook(self, book: Orderbook) -> None:
        sides = []
        for side, half in ((0, book.buy), (1, book.sell)):
            prices = list(reversed(half)) if side == 0 else list(half)
            rows = []
            for p in prices:
                for o in half[p]:
                    if not o.is_external:
                        aid = o.internal_id if o.internal_id and o.internal_id > 0 else len(self._agent_of_iid) + 1
                        ref = abi.REF_AGENT | aid
                        self._agent_of_iid[aid] = aid
                        self._iid_of_ref[ref] = aid
                        self._meta[ref] = (o.timestamp,)
                    elif o.internal_id == -1:
                        ref = abi.REF_AGGREGATE
                        self._meta[("agg", side, int(p))] = (o.timestamp,)
                    else:
                        ref = self._ref_for_external(o.external_id)
                        if o.internal_id is not None:
                            self._iid_of_ref[ref] = o.internal_id
                            self.order_id_convertor.external_to_internal_lookup[o.external_id] = o.internal_id
                        self._meta[ref] = (o.timestamp,)
                    if o.internal_id is not None and o.internal_id > self.order_id_convertor.counter:
                        self.order_id_convertor.counter = o.internal_id
                    rows.append((int(p), int(o.volume), ref, 0))
            sides.append(np.array(rows, dtype=abi.BOOK_ENTRY_DTYPE))
        self.sim.set_book(self.env, sides[0], sides[1])

    @property
    def internal_orderbook(self) -> Orderbook:
        entries = []
        for s in (0, 1):
            e = self.sim.dump_agent_orders(self.env, s)
            # the internal book is price-sorted with FIFO (= ascending id) inside a level
            order = np.lexsort((e["ref"], e["price"]))
            entries.append(e[order])
        return self._view(entries)

    @internal_orderbook.setter
    def internal_orderbook(self, book: Orderbook) -> None:
        if len(book.buy) or len(book.sell):
            raise NotImplementedError("set the central_orderbook with is_external=False orders instead")

    def reset_internal_orderbook(self):
        if len(self.sim.dump_agent_orders(self.env, 0)) or len(self.sim.dump_agent_orders(self.env, 1)):
            raise NotImplementedError("reset_internal_orderbook with agent orders still resting in the central book")

    def get_initial_orderbook_from_orders(self, orders: List[LimitOrder]) -> Orderbook:
        assert all(order.internal_id == -1 for order in orders), "internal_ids of orders in the initial book must be -1"
        orderbook = self.get_empty_orderbook()
        for order in orders:
            assert order.is_external, "Initial orders must all be external."
            getattr(orderbook, order.direction)[order.price] = deque([order])
        return orderbook

    # ---- prices ------------------------------------------------------------------------------------------------------
    def _state(self):
        return self.sim.state(self.env, 1)[0]

    @property
    def best_sell_price(self):
        st = self._state()
        return np.inf if st["best_sell"] == 2**31 - 1 else int(st["best_sell"])

    @property
    def best_buy_price(self):
        return int(self._state()["best_buy"])

    @property
    def orderbook_price_range(self):
        buy, sell = self.sim.dump_book(self.env, 0), self.sim.dump_book(self.env, 1)
        if not len(buy) or not len(sell):
            raise StopIteration
        return int(buy["price"].min()), int(sell["price"].max())

    # ---- order processing --------------------------------------------------------------------------------------------
    def process_order(self, order: Order) -> Optional[FilledOrders]:
        if hasattr(order, "volume") and order.volume is not None:
            assert order.volume > 0, f"Order volume must be positive. Instead, order.volume = {order.volume}."
        if isinstance(order, LimitOrder):
            return self.submit_order(order)
        elif isinstance(order, MarketOrder):
            return self.execute_order(order)
        elif isinstance(order, (Cancellation, Deletion)):
            self.remove_order(order)
            return None
        raise NotImplementedError(f"Cannot process order of type {type(order)}.")

    def _run(self, mtype: int, order, volume: int, ref: int) -> FilledOrders:
        rec = np.zeros(1, abi.ORDER_DTYPE)
        price = int(order.price) if getattr(order, "price", None) is not None else 0
        rec[0] = (self.env, mtype, _SIDE[order.direction], price, int(volume), int(order.is_external), ref, 0)
        err_before = int(self._state()["err"])
        fills, refs = self.sim.process_orders(rec)
        out = FilledOrders()
        for f in fills:
            fref = int(f["ref"])
            if f["is_market"]:
                o = MarketOrder(order.timestamp, order.direction, order.ticker, order.internal_id, order.external_id,
                                order.is_external, volume=int(f["volume"]), price=int(f["price"]))
            else:
                is_agent = bool(fref & abi.REF_AGENT)
                iid = -1 if fref == 0 else self._iid_of_ref.get(fref)
                o = LimitOrder(self._meta.get(fref, (None,))[0], _DIR[int(f["direction"])], self.ticker, iid,
                               None if is_agent or fref == 0 else self._ext_of_ref(fref), not is_agent,
                               price=int(f["price"]), volume=int(f["volume"]))
            (out.internal if f["list"] == 0 else out.external).append(o)
        rested = int(refs[0])
        if rested:  # OrderIdConvertor.add_internal_id_to_order_and_track
            conv = self.order_id_convertor
            conv.counter += 1
            if order.is_external:
                conv.external_to_internal_lookup[order.external_id] = conv.counter
                self._iid_of_ref[ref] = conv.counter
                self._meta[ref] = (order.timestamp,)
            else:
                aref = abi.REF_AGENT | rested
                self._iid_of_ref[aref] = conv.counter
                self._agent_of_iid[conv.counter] = rested
                self._meta[aref] = (order.timestamp,)
        err = int(self._state()["err"])
        if err & abi.ERR_EMPTY_BOOK and not err_before & abi.ERR_EMPTY_BOOK:
            opposite = "sell" if order.direction == "buy" else "buy"
            raise EmptyOrderbookError(f"Trying take liquidity from empty {opposite} side of the book.")
        if err & ~err_before & (abi.ERR_LEVEL_OVERFLOW | abi.ERR_ORDER_OVERFLOW | abi.ERR_AGENT_OVERFLOW):
            raise MemoryError("device book capacity exceeded: " + ", ".join(n for b, n in abi.ERR_NAMES.items() if err & b))
        return out

    def submit_order(self, order: LimitOrder) -> Optional[FilledOrders]:
        ref = self._ref_for_external(order.external_id) if order.is_external else 0
        crosses = self._does_order_cross_spread(order)
        filled = self._run(abi.MSG_LIMIT, order, order.volume, ref)
        return filled if crosses else None

    def execute_order(self, order: FillableOrder) -> FilledOrders:
        ref = self._ref_for_external(order.external_id) if order.is_external else 0
        return self._run(abi.MSG_LIMIT if isinstance(order, LimitOrder) else abi.MSG_MARKET, order, order.volume, ref)

    def remove_order(self, order: Union[Cancellation, Deletion]) -> None:
        if order.is_external:
            ref = self._lookup_ref(order.external_id) if order.external_id is not None else None
            ref = 0x7FFFFFFF if ref is None else ref
        else:
            ref = self._agent_of_iid.get(order.internal_id, 0x7FFFFFF0)
        level = [e for e in self.sim.dump_book(self.env, _SIDE[order.direction]) if e["price"] == order.price]
        if not level:
            warnings.warn(f"No {order.direction} orders found at level {order.price}")
        full_ref = ref if order.is_external else (abi.REF_AGENT | ref)
        hit = [e for e in level if int(e["ref"]) == full_ref]
        if level and not hit:
            warnings.warn(f"No order found with internal_id = {order.internal_id}")
            if int(level[0]["ref"]) == abi.REF_AGGREGATE:
                assert order.volume is not None, "When deleting an initial order, a volume must be provided."
                order.internal_id = -1  # the reference mutates the caller's order, Exchange.py:136
        elif hit and order.volume is None:
            order.volume = int(hit[0]["volume"])  # Exchange.py:140-141
        self._run(abi.MSG_CANCEL if isinstance(order, Cancellation) else abi.MSG_DELETE, order,
                  0 if order.volume is None else order.volume, ref)
        if hit and int(hit[0]["volume"]) <= (order.volume or 0) and order.is_external:
            self.order_id_convertor.remove_external_order_id(order.external_id)

    def _does_order_cross_spread(self, order: FillableOrder):
        if isinstance(order, MarketOrder):
            return True
        if order.direction == "buy":
            return order.price >= self.best_sell_price
        return order.price <= self.best_buy_price
