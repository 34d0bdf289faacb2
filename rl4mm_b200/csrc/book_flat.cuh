// book_flat.cuh -- the FLAT book of the replay kernel: a side is an unordered pool of resting orders
//
//     pool[s][i] = { price, volume, ref, seq }      i < n[s] <= FLAT_CAP (64), 16 bytes per order, no gaps
//
// with NO level structure at all.  It is the reference's `SortedDict[price -> deque[LimitOrder]]` (rl4mm/orderbook/models.py:64-69)
// for books that are small enough that one or two warp-wide compares see every resting order of a side (BASELINE config 2: 8-12
// levels and 30-55 orders per side):
//   * price-time priority is carried by `seq`, a per-book counter stamped when an order starts resting (Exchange.py:78-83 appends
//     to the level's deque): the head of a level = the order with the smallest seq at that price (one REDUX.MIN);
//   * a new order is ONE 16-byte store at pool[n] (no level search, no level insert, no queue shift, no prefix-end update);
//   * a cancellation / deletion finds (price, ref) with two 16-byte loads per lane + ballots and fills the hole with the last
//     order (one load + one store) -- Exchange.remove_order / _find_queue_position, Exchange.py:122-147,196-217;
//   * the best price of a side lives in a register and is recomputed (REDUX.MAX / MIN over the prices already in registers)
//     only when the last order at the best price leaves.
// About 45 warp instructions per message against 110 for the sorted level arrays of book_fast.cuh (profiles/r02_replay_ab.md).
//
// The HBM blob stays in the canonical sorted layout (book.cuh): the kernel converts on entry (flat_enter: trivial, seq = position
// in the sorted order array) and on exit (flat_leave: rank-by-counting sort on (price, seq)), so every other kernel, the L3 dump
// and the env path are untouched.  A book that does not fit (more than FLAT_CAP orders on a side) runs on the sorted path of
// book_fast.cuh in the same kernel and moves back when it has shrunk; update_outer_levels (OrderbookSimulator.py:105-135) is
// run on the sorted form (leave -> fast_resync -> enter) on the seconds where the snapshot really has levels beyond the range.
#pragma once
#include "book_fast.cuh"

#define FLAT_CAP 64

struct FlatState {
  int n0, n1;      // resting orders per side
  uint32_t seq;    // next time-priority stamp
};

template <class LT>
__device__ __forceinline__ uint4* flat_pool(unsigned char* blob, int s) {
  static_assert(LT::NO * 8 >= FLAT_CAP * 16, "the flat pool (FLAT_CAP x 16 B) lives in the order array of the side");
  static_assert(LT::NL >= FLAT_CAP, "a flat side may hold FLAT_CAP distinct prices");
  static_assert(LT::ord_off % 16 == 0 && LT::side_stride % 16 == 0 && LT::side_off % 16 == 0, "16-byte aligned pool");
  return reinterpret_cast<uint4*>(blob + LT::side_off + s * LT::side_stride + LT::ord_off);
}

#define FLAT_EMPTY_ENTRY make_uint4(0u, 0u, 0xffffffffu, 0xffffffffu)   // ref / seq that no resting order has

// both chunks of a side's pool, one order per lane and chunk
__device__ __forceinline__ void flat_load(const uint4* pool, int n, int lane, uint4& e0, uint4& e1) {
  e0 = FLAT_EMPTY_ENTRY; e1 = FLAT_EMPTY_ENTRY;
  if (lane < n) e0 = pool[lane];
  if (lane + 32 < n) e1 = pool[lane + 32];
}

// best price of side S over the orders in registers, leaving out entry `skip` (-1: none)
template <int S>
__device__ __forceinline__ int flat_best_of(const uint4& e0, const uint4& e1, int n, int lane, int skip) {
  const int sent = S ? INT32_MAX : INT32_MIN;
  const int k0 = (lane < n && lane != skip) ? (int)e0.x : sent;
  const int k1 = (lane + 32 < n && lane + 32 != skip) ? (int)e1.x : sent;
  return S ? __reduce_min_sync(FULL_MASK, k0 < k1 ? k0 : k1) : __reduce_max_sync(FULL_MASK, k0 > k1 ? k0 : k1);
}

// One order of side S (0 buy, 1 sell) through the flat book.  Same results as fast_order_full<LT,TR> on the sorted book; TR: fills /
// flows / the agent's order tables are tracked (env kernels), exactly as in book_fast.cuh.
// Returns false -- with the book untouched -- when the order may have to rest and the pool of its side is full: the caller
// converts the book to the sorted form and runs the order there.
template <class LT, int S, bool TR>
__device__ __forceinline__ bool flat_order(unsigned char* blob, int lane, FastState& f, FlatState& st, int type, int price, int vol, uint32_t ref, bool is_agent) {
  constexpr int OPP = S ^ 1;
  if (!TR) is_agent = false;
  int& n_own = S ? st.n1 : st.n0;
  int& n_opp = S ? st.n0 : st.n1;
  int& best_own = S ? f.best1 : f.best0;
  int& best_opp = S ? f.best0 : f.best1;
  uint4* own = flat_pool<LT>(blob, S);
  uint4* opp = flat_pool<LT>(blob, OPP);
  FastBook<LT> fb; fb.blob = blob; fb.lane = lane;
  if (vol <= 0) { f.err |= LOBSIM_ERR_BAD_VOLUME; return true; }          // assert order.volume > 0, Exchange.py:59-60
  __syncwarp();                                                             // the previous order's stores are visible
  if (type == LOBSIM_MSG_LIMIT || type == LOBSIM_MSG_MARKET) {
    int rem = vol;
    if (type == LOBSIM_MSG_LIMIT && n_own >= FLAT_CAP) return false;
    const bool crosses = S ? price <= best_opp : price >= best_opp;        // empty opposite side: INT32_MIN / INT32_MAX
    if (type == LOBSIM_MSG_MARKET || crosses) {
      // ---- execution against the opposite side, best price first, oldest order first (Exchange.py:85-120) -------------
#pragma unroll 1
      while (rem > 0) {
        if (n_opp == 0) {
          if (type == LOBSIM_MSG_MARKET) { f.err |= LOBSIM_ERR_EMPTY_BOOK; f.dead = 1; }   // EmptyOrderbookError :183-186
          break;
        }
        const int bp = best_opp;
        if (type == LOBSIM_MSG_LIMIT && !(S ? price <= bp : price >= bp)) break;
        uint4 e0, e1;
        flat_load(opp, n_opp, lane, e0, e1);
        const unsigned q0 = (int)e0.x == bp ? e0.w : 0xffffffffu, q1 = (int)e1.x == bp ? e1.w : 0xffffffffu;   // empty lanes: seq = ~0
        const unsigned head_seq = __reduce_min_sync(FULL_MASK, q0 < q1 ? q0 : q1);
        const unsigned b0 = __ballot_sync(FULL_MASK, q0 == head_seq), b1 = __ballot_sync(FULL_MASK, q1 == head_seq);
        const int i = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;              // the head of the best queue
        const int hv = (int)__shfl_sync(FULL_MASK, b0 ? e0.y : e1.y, i & 31);
        const uint32_t href = TR ? __shfl_sync(FULL_MASK, b0 ? e0.z : e1.z, i & 31) : 0u;
        const bool hagent = TR && (href & LOBSIM_REF_AGENT) != 0;
        const bool self_match = TR && is_agent && hagent;                  // cannot fill our own order => delete it, :91-94
        if (!self_match) {
          const int v = rem < hv ? rem : hv;
          if (TR) {
            if (hagent) fast_record(fb, f, 0, OPP, bp, v, 0, href);
            else fast_record(fb, f, 1, OPP, bp, v, 0, href);
            if (is_agent) fast_record(fb, f, 0, S, bp, v, 1, href);         // the synthetic MarketOrder fill, :111-115
          }
          if (rem < hv) {                                                  // partial fill of the head
            if (lane == 0) opp[i].y = (unsigned)(hv - rem);
            if (TR && hagent) { __syncwarp(); fast_agent_reduce(fb, OPP, href & 0x7fffffffu, rem, false); }
            rem = 0;
            break;
          }
          rem -= hv;                                                       // the head is consumed
        }
        const int at_best = __popc(__ballot_sync(FULL_MASK, q0 != 0xffffffffu)) + __popc(__ballot_sync(FULL_MASK, q1 != 0xffffffffu));
        if (at_best == 1) best_opp = flat_best_of<OPP>(e0, e1, n_opp, lane, i);   // the level emptied
        if (lane == 0) opp[i] = opp[n_opp - 1];
        n_opp -= 1;
        __syncwarp();
        if (TR && hagent) fast_agent_reduce(fb, OPP, href & 0x7fffffffu, 0, true);   // the resting agent order is gone
      }
      if (!(rem > 0 && type == LOBSIM_MSG_LIMIT && !f.dead)) return true;
      // the remainder of a crossing limit order rests (Exchange.py:116-119)
    }
    // ---- the order rests at the back of its price's queue (Exchange.py:74-83) ------------------------------------------
    if (TR && is_agent) {   // OrderIdConvertor.add_internal_id_to_order_and_track + internal book append
      BookHdr* h = reinterpret_cast<BookHdr*>(blob);
      const int nag = h->nag[S];
      if (nag >= LT::NA) { f.err |= LOBSIM_ERR_AGENT_OVERFLOW; return true; }
      const uint32_t id = h->next_agent_id;
      ref = LOBSIM_REF_AGENT | id;
      __syncwarp();
      if (lane == 0) {
        int32_t* ap = reinterpret_cast<int32_t*>(blob + LT::agent_off + S * LT::NA * 12);
        ap[nag] = price; ap[LT::NA + nag] = rem; reinterpret_cast<uint32_t*>(ap)[2 * LT::NA + nag] = id;
        h->nag[S] = nag + 1; h->next_agent_id = id + 1;
      }
    }
    if (lane == 0) own[n_own] = make_uint4((unsigned)price, (unsigned)rem, ref, st.seq);
    n_own += 1; st.seq += 1;
    if (S ? price < best_own : price > best_own) best_own = price;
    return true;
  }
  // ---- cancellation / deletion (Exchange.py:122-147) ----------------------------------------------------------------------
  uint4 e0, e1;
  flat_load(own, n_own, lane, e0, e1);
  unsigned m0 = __ballot_sync(FULL_MASK, (int)e0.x == price && e0.z == ref), m1 = __ballot_sync(FULL_MASK, (int)e1.x == price && e1.z == ref);
  bool aggregate = false;
  if (!(m0 | m1)) {
    // unknown id: the level's snapshot aggregate (internal_id -1, always the head of its level) takes the hit (:133-137);
    // no such level, or no aggregate left at it: nothing happens (:129-132,138-139)
    m0 = __ballot_sync(FULL_MASK, lane < n_own && (int)e0.x == price && e0.z == LOBSIM_REF_AGGREGATE);
    m1 = __ballot_sync(FULL_MASK, lane + 32 < n_own && (int)e1.x == price && e1.z == LOBSIM_REF_AGGREGATE);
    if (!(m0 | m1)) return true;
    aggregate = true;
  }
  const int i = m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1;
  const int cur = (int)__shfl_sync(FULL_MASK, m0 ? e0.y : e1.y, i & 31);
  if (vol < cur) {                                                         // partial: reduce in place
    if (lane == 0) own[i].y = (unsigned)(cur - vol);
    if (TR && is_agent && !aggregate) { __syncwarp(); fast_agent_reduce(fb, S, ref & 0x7fffffffu, vol, false); }
    return true;
  }
  // full removal (over-size requests remove the resting volume, :142-146): the last order of the pool fills the hole
  if (price == best_own) best_own = flat_best_of<S>(e0, e1, n_own, lane, i);
  if (lane == 0) own[i] = own[n_own - 1];
  n_own -= 1;
  if (TR && is_agent && !aggregate) { __syncwarp(); fast_agent_reduce(fb, S, ref & 0x7fffffffu, cur, true); }
  return true;
}

template <class LT>
__device__ __forceinline__ bool flat_message(unsigned char* blob, int lane, FastState& f, FlatState& st, int price, int vol, uint32_t ref, uint32_t meta) {
  const int type = (int)(meta & 7u);
  if (meta & 8u) return flat_order<LT, 1, false>(blob, lane, f, st, type, price, vol, ref, false);
  return flat_order<LT, 0, false>(blob, lane, f, st, type, price, vol, ref, false);
}
// the tracked form (env kernels): historical messages and the agent's own orders
template <class LT>
__device__ __forceinline__ bool flat_order_tracked(unsigned char* blob, int lane, FastState& f, FlatState& st, int type, int side, int price, int vol, uint32_t ref, bool is_agent) {
  if (side) return flat_order<LT, 1, true>(blob, lane, f, st, type, price, vol, ref, is_agent);
  return flat_order<LT, 0, true>(blob, lane, f, st, type, price, vol, ref, is_agent);
}

// volume resting at the best price of side S (Orderbook.best_buy_volume / best_sell_volume, models.py:80-85)
template <class LT, int S>
__device__ __forceinline__ int flat_best_volume(unsigned char* blob, int lane, const FastState& f, const FlatState& st) {
  const int n = S ? st.n1 : st.n0, best = S ? f.best1 : f.best0;
  uint4 e0, e1;
  flat_load(flat_pool<LT>(blob, S), n, lane, e0, e1);
  const int v = ((lane < n && (int)e0.x == best) ? (int)e0.y : 0) + ((lane + 32 < n && (int)e1.x == best) ? (int)e1.y : 0);
  return __reduce_add_sync(FULL_MASK, v);
}

// ---- sorted -> flat (kernel entry, and after a resync / when a book has shrunk) ---------------------------------------------
template <class LT>
__device__ __forceinline__ bool flat_fits(const FastBook<LT>& fb, int slack) {
  return fb.cnt(0)->y <= FLAT_CAP - slack && fb.cnt(1)->y <= FLAT_CAP - slack;
}
// precondition: flat_fits(fb, 0); f.best0 / f.best1 are current
template <class LT>
__device__ __forceinline__ void flat_enter(const FastBook<LT>& fb, FlatState& st) {
  const int lane = fb.lane;
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 2; s++) {
    unsigned char* sb = fb.side(s);
    const int2 c = *fb.cnt(s);
    const int nlv = c.x, n = c.y;
    uint2 o0 = make_uint2(0u, 0u), o1 = make_uint2(0u, 0u);
    int p0 = 0, p1 = 0;
    // the level of order i = the first level whose end offset is beyond i (binary search over the level ends)
    auto price_of = [&](int i) {
      int lo = 0, hi = nlv - 1;
#pragma unroll 1
      while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)fb.LE(sb)[mid] > i) hi = mid; else lo = mid + 1; }
      return fb.P(sb)[lo];
    };
    if (lane < n) { o0 = fb.O(sb)[lane]; p0 = price_of(lane); }
    if (lane + 32 < n) { o1 = fb.O(sb)[lane + 32]; p1 = price_of(lane + 32); }
    __syncwarp();                                                           // the pool overlays the order array
    uint4* pool = flat_pool<LT>(fb.blob, s);
    if (lane < n) pool[lane] = make_uint4((unsigned)p0, o0.x, o0.y, (unsigned)lane);
    if (lane + 32 < n) pool[lane + 32] = make_uint4((unsigned)p1, o1.x, o1.y, (unsigned)(lane + 32));
    if (s) st.n1 = n; else st.n0 = n;
  }
  st.seq = FLAT_CAP;
  __syncwarp();
}

// ---- flat -> sorted (kernel exit, pool overflow, resync): rank every order by (price worst -> best, seq), scatter, rebuild the
//      level prices / ends ----------------------------------------------------------------------------------------------------
template <class LT>
__device__ __forceinline__ void flat_leave(const FastBook<LT>& fb, const FlatState& st) {
  const int lane = fb.lane;
  __syncwarp();
#pragma unroll
  for (int s = 0; s < 2; s++) {
    unsigned char* sb = fb.side(s);
    const uint4* pool = flat_pool<LT>(fb.blob, s);
    const int n = s ? st.n1 : st.n0;
    uint4 e0, e1;
    flat_load(pool, n, lane, e0, e1);
    auto key = [&](const uint4& e) {                                        // ascending = worst -> best price, then oldest first
      uint32_t pk = e.x ^ 0x80000000u;
      if (s) pk = ~pk;
      return ((unsigned long long)pk << 32) | e.w;
    };
    const unsigned long long k0 = key(e0), k1 = key(e1);
    int r0 = 0, r1 = 0;
#pragma unroll 1
    for (int m = 0; m < n; m++) {
      const unsigned long long km = key(pool[m]);
      r0 += km < k0 ? 1 : 0; r1 += km < k1 ? 1 : 0;
    }
    __syncwarp();                                                           // every lane holds its orders: the arrays can be rewritten
    if (lane < n) { fb.O(sb)[r0] = make_uint2(e0.y, e0.z); fb.P(sb)[r0] = (int)e0.x; }          // P: sorted price per ORDER for now
    if (lane + 32 < n) { fb.O(sb)[r1] = make_uint2(e1.y, e1.z); fb.P(sb)[r1] = (int)e1.x; }
    __syncwarp();
    int t0 = 0, t1 = 0, pv0 = 0, pv1 = 0;
    if (lane < n) t0 = fb.P(sb)[lane];
    if (lane + 32 < n) t1 = fb.P(sb)[lane + 32];
    if (lane > 0 && lane < n) pv0 = fb.P(sb)[lane - 1];
    if (lane + 32 < n) pv1 = fb.P(sb)[lane + 31];
    const bool new0 = lane < n && (lane == 0 || t0 != pv0), new1 = lane + 32 < n && t1 != pv1;   // first order of a level
    const unsigned nb0 = __ballot_sync(FULL_MASK, new0), nb1 = __ballot_sync(FULL_MASK, new1);
    const unsigned le_mask = 0xffffffffu >> (31 - lane);                    // lanes <= this lane
    const int lvl0 = __popc(nb0 & le_mask) - 1, lvl1 = __popc(nb0) + __popc(nb1 & le_mask) - 1;
    // last order of a level: the next order starts a new level, or there is no next order
    const bool next_new0 = lane == 31 ? (nb1 & 1u) != 0 : ((nb0 >> (lane + 1)) & 1u) != 0;
    const bool next_new1 = lane == 31 ? false : ((nb1 >> (lane + 1)) & 1u) != 0;
    const bool last0 = lane < n && (lane == n - 1 || next_new0), last1 = lane + 32 < n && (lane + 32 == n - 1 || next_new1);
    __syncwarp();                                                           // the per-order prices have been read
    if (new0) fb.P(sb)[lvl0] = t0;
    if (new1) fb.P(sb)[lvl1] = t1;
    if (last0) fb.LE(sb)[lvl0] = (uint16_t)(lane + 1);
    if (last1) fb.LE(sb)[lvl1] = (uint16_t)(lane + 33);
    if (lane == 0) *fb.cnt(s) = make_int2(__popc(nb0) + __popc(nb1), n);
  }
  __syncwarp();
}

// does the snapshot row of this second hold a level beyond the tracked price range (OrderbookSimulator.py:99-103,113-115)?
__device__ __forceinline__ bool flat_resync_needed(const BookHdr* h, const int32_t* __restrict__ row, int L, int lane) {
  const int min_buy = h->min_buy, max_sell = h->max_sell;
  unsigned any = 0;
  for (int base = 0; base < 2 * L; base += 32) {
    const int idx = base + lane;
    int price = LOBSIM_NO_PRICE;
    if (idx < 2 * L) price = __ldg(&row[idx * 2]);
    any |= __ballot_sync(FULL_MASK, price != LOBSIM_NO_PRICE && (idx < L ? price < min_buy : price > max_sell));
  }
  return any != 0;
}
// the price-range trackers when no level is overwritten (OrderbookSimulator.py:134-135): worst resting price of each side
template <class LT>
__device__ __forceinline__ void flat_update_trackers(unsigned char* blob, int lane, const FlatState& st) {
  BookHdr* h = reinterpret_cast<BookHdr*>(blob);
  __syncwarp();
  uint4 e0, e1;
  flat_load(flat_pool<LT>(blob, 0), st.n0, lane, e0, e1);
  const int w0 = flat_best_of<1>(e0, e1, st.n0, lane, -1);                // lowest bid
  flat_load(flat_pool<LT>(blob, 1), st.n1, lane, e0, e1);
  const int w1 = flat_best_of<0>(e0, e1, st.n1, lane, -1);                // highest ask
  if (lane == 0) {
    if (st.n0 && w0 < h->min_buy) h->min_buy = w0;
    if (st.n1 && w1 > h->max_sell) h->max_sell = w1;
  }
  __syncwarp();
}
