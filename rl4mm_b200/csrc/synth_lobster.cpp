// synth_lobster.cpp -- seeded generator of self-consistent LOBSTER-shaped message streams, written straight into
// the packed device format (include/lobsim.h: lobsim_msg_t records, per-step CSR offsets, per-second snapshots).
//
// Host-side input generator for the benchmarks and the large-scale parity tests (BASELINE.json configs 2-5: "synthetic
// SPY-shaped day", "multi-ticker heavy cancel/modify flow").  It keeps its own small price-time-priority book so
// that every cancellation / deletion / execution refers to an order that is really resting with at least that size,
// exactly like a LOBSTER file (SURVEY.md App. A.6 describes the file semantics):
//   type 1  new limit order (never crossing)                      -> LOBSIM_MSG_LIMIT
//   type 2  partial cancellation of a live order                  -> LOBSIM_MSG_CANCEL
//   type 3  deletion of a live order, size = its remaining size   -> LOBSIM_MSG_DELETE
//   type 4  execution of the order at the head of the best level  -> LOBSIM_MSG_MARKET with the AGGRESSOR direction
//           (the packer's direction flip, rl4mm/database/database_population_helpers.py:132-136, already applied)
// Orders of the initial book get references too; they appear in the snapshot as aggregates, so messages that touch
// them exercise the simulator's "unknown id hits the aggregate" path (rl4mm/orderbook/Exchange.py:128-137).
//
// This file is independent of oracle/ (the product never links the oracle) and of the CUDA path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <vector>

#include "../../include/lobsim.h"

extern "C" {

typedef struct {
  uint64_t seed;
  uint64_t n_msgs;
  int32_t duration_s;       // messages are spread over (0, duration_s] seconds after the grid origin
  int32_t n_levels;         // snapshot depth L
  int32_t tick;
  int32_t mid0;             // initial best ask (multiple of tick); best bid = mid0 - tick
  int64_t step_us;
  double p_limit, p_cancel, p_delete, p_exec;
  double geom_p;            // limit price distance from the opposite best ~ 1 + Geometric(geom_p) ticks
  int32_t init_levels;      // populated levels per side in the initial book
  int32_t mean_queue;       // mean orders per level in the initial book
  int32_t target_orders;    // soft cap on live orders per side (keeps the book stationary)
  int32_t max_offset_ticks; // hard cap on the limit price distance
  double size_sigma;        // log-normal sigma of order sizes (median 100 shares)
  double p_sweep;           // per-message probability of starting a sweep that executes a whole best level
} lobsynth_cfg_t;

int lobsynth_generate(const lobsynth_cfg_t* cfg, lobsim_msg_t* msgs, uint32_t* step_off, uint32_t n_grid_steps,
                      int32_t* snapshots, uint8_t* snap_valid, uint32_t n_seconds);
}

namespace {

struct Rng {  // xoshiro256**
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0, 1)
  uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n); }
  double normal() {
    double u1 = 1.0 - uniform(), u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
  double exponential() { return -std::log(1.0 - uniform()); }
};

struct Ord {
  int32_t price, vol;
  uint8_t side, alive;
  uint32_t slot;  // index in live[side] while alive
};

struct Level {
  std::deque<uint32_t> q;  // order ids in arrival order (dead ones are skipped lazily)
  int64_t vol = 0;
  int32_t live = 0;
};

struct Book {
  std::map<int32_t, Level> side[2];  // 0 = buy (best = rbegin), 1 = sell (best = begin)
  std::vector<Ord> ords;             // id = index + 1
  std::vector<uint32_t> live[2];

  int32_t best(int s) const { return s == 0 ? side[0].rbegin()->first : side[1].begin()->first; }

  uint32_t add(int s, int32_t price, int32_t vol) {
    ords.push_back(Ord{price, vol, (uint8_t)s, 1, (uint32_t)live[s].size()});
    uint32_t id = (uint32_t)ords.size();
    live[s].push_back(id);
    Level& l = side[s][price];
    l.q.push_back(id); l.vol += vol; l.live++;
    return id;
  }
  void reduce(uint32_t id, int32_t v) {  // v <= remaining volume
    Ord& o = ords[id - 1];
    auto it = side[o.side].find(o.price);
    o.vol -= v; it->second.vol -= v;
    if (o.vol == 0) {
      o.alive = 0;
      uint32_t last = live[o.side].back();
      live[o.side][o.slot] = last; ords[last - 1].slot = o.slot; live[o.side].pop_back();
      if (--it->second.live == 0) side[o.side].erase(it);
    }
  }
  uint32_t head(int s) {  // live order at the head of the best level
    Level& l = s == 0 ? side[0].rbegin()->second : side[1].begin()->second;
    while (!ords[l.q.front() - 1].alive) l.q.pop_front();
    return l.q.front();
  }
  void snapshot(int L, int32_t* out) const {  // [2][L][2]
    for (int s = 0; s < 2; s++) {
      int k = 0;
      auto emit = [&](int32_t p, const Level& l) { out[(s * L + k) * 2] = p; out[(s * L + k) * 2 + 1] = (int32_t)l.vol; k++; };
      if (s == 0) for (auto it = side[0].rbegin(); it != side[0].rend() && k < L; ++it) emit(it->first, it->second);
      else for (auto it = side[1].begin(); it != side[1].end() && k < L; ++it) emit(it->first, it->second);
      for (; k < L; k++) { out[(s * L + k) * 2] = LOBSIM_NO_PRICE; out[(s * L + k) * 2 + 1] = 0; }
    }
  }
};

int32_t draw_size(Rng& r, double sigma) {
  if (r.uniform() < 0.15) return 1 + (int32_t)r.below(99);  // odd lot
  double lots = std::exp(sigma * r.normal());
  int32_t n = (int32_t)std::llround(lots);
  return 100 * std::max(1, std::min(n, 200));
}

}  // namespace

int lobsynth_generate(const lobsynth_cfg_t* c, lobsim_msg_t* msgs, uint32_t* step_off, uint32_t n_grid_steps,
                      int32_t* snapshots, uint8_t* snap_valid, uint32_t n_seconds) {
  if (!c || !msgs || !step_off || !snapshots || !snap_valid) return LOBSIM_E_INVALID;
  if (c->n_levels <= 0 || c->tick <= 0 || c->duration_s <= 0 || c->step_us <= 0 || 1000000 % c->step_us) return LOBSIM_E_INVALID;
  if ((uint64_t)n_seconds != (uint64_t)c->duration_s || (uint64_t)n_grid_steps * (uint64_t)c->step_us != (uint64_t)c->duration_s * 1000000ULL)
    return LOBSIM_E_INVALID;
  const int L = c->n_levels;
  Rng rng(c->seed), trng(c->seed ^ 0x5851f42d4c957f2dULL);
  Book b;
  b.ords.reserve((size_t)(c->n_msgs * (c->p_limit + 0.05)) + 4096);
  // ---- initial book ------------------------------------------------------------------------------------------
  for (int k = 0; k < c->init_levels; k++)
    for (int s = 0; s < 2; s++) {
      int32_t price = s == 0 ? c->mid0 - c->tick * (k + 1) : c->mid0 + c->tick * k;
      int n = 1 + (int)rng.below((uint64_t)std::max(1, 2 * c->mean_queue - 1));
      for (int j = 0; j < n; j++) b.add(s, price, draw_size(rng, c->size_sigma));
    }
  const size_t snap_stride = (size_t)2 * L * 2;
  b.snapshot(L, snapshots); snap_valid[0] = 1;
  // ---- arrival times: exponential gaps rescaled to fill the duration exactly (two passes over the same RNG) -----
  double total = 0.0;
  { Rng t2 = trng; for (uint64_t i = 0; i <= c->n_msgs; i++) total += t2.exponential(); }
  const double scale = (double)c->duration_s * 1e6 / total;
  double tacc = 0.0;
  uint32_t next_sec = 1;
  int sweep_left = 0, sweep_side = 0;
  int32_t sweep_price = 0;
  std::vector<uint32_t> counts(n_grid_steps, 0);
  const double pl = c->p_limit, pc = pl + c->p_cancel, pd = pc + c->p_delete;
  for (uint64_t i = 0; i < c->n_msgs; i++) {
    tacc += trng.exponential();
    int64_t ts = (int64_t)(tacc * scale);
    if (ts < 1) ts = 1;
    if (ts > (int64_t)c->duration_s * 1000000) ts = (int64_t)c->duration_s * 1000000;
    while (next_sec <= n_seconds && (int64_t)next_sec * 1000000 < ts) {  // book after the last message <= boundary
      b.snapshot(L, snapshots + snap_stride * next_sec); snap_valid[next_sec] = 1; next_sec++;
    }
    counts[(size_t)((ts + c->step_us - 1) / c->step_us - 1)]++;
    // ---- choose the event ------------------------------------------------------------------------------------
    double u = rng.uniform();
    int type = u < pl ? 1 : u < pc ? 2 : u < pd ? 3 : 4;
    int s = (int)(rng.next() & 1);
    if (sweep_left == 0 && rng.uniform() < c->p_sweep) {  // a burst of executions that clears one best level
      sweep_side = (int)(rng.next() & 1);
      sweep_price = b.best(sweep_side);
      sweep_left = 64;
    }
    bool sweeping = false;
    if (sweep_left > 0) {
      if (b.side[sweep_side].size() >= 4 && b.best(sweep_side) == sweep_price) { type = 4; s = sweep_side; sweeping = true; sweep_left--; }
      else sweep_left = 0;
    }
    for (int t = 0; t < 2; t++)  // keep both sides populated and the book stationary
      if (b.side[t].size() < 4 || b.live[t].size() < 8) { type = 1; s = t; sweeping = false; }
    if (type == 1 && (int)b.live[s].size() > c->target_orders && rng.uniform() < 0.5) type = 3;
    if ((type == 2 || type == 3) && (int)b.live[s].size() < c->target_orders / 4) type = 1;
    lobsim_msg_t m;
    if (type == 2) {  // partial cancellation: needs an order with volume >= 2
      uint32_t id = 0;
      for (int tries = 0; tries < 4 && !id; tries++) { uint32_t cand = b.live[s][rng.below(b.live[s].size())]; if (b.ords[cand - 1].vol >= 2) id = cand; }
      if (!id) type = 3;
      else {
        Ord& o = b.ords[id - 1];
        int32_t v = o.vol >= 200 && rng.uniform() < 0.7 ? 100 * (1 + (int32_t)rng.below((uint64_t)((o.vol - 1) / 100)))
                                                        : 1 + (int32_t)rng.below((uint64_t)o.vol - 1);
        m.price = o.price; m.volume = v; m.ref = id; m.meta = LOBSIM_META(LOBSIM_MSG_CANCEL, s);
        b.reduce(id, v);
      }
    }
    if (type == 3) {
      uint32_t id = b.live[s][rng.below(b.live[s].size())];
      Ord& o = b.ords[id - 1];
      m.price = o.price; m.volume = o.vol; m.ref = id; m.meta = LOBSIM_META(LOBSIM_MSG_DELETE, s);
      b.reduce(id, o.vol);
    } else if (type == 4) {  // execution against the head of side s; the aggressor has the opposite direction
      uint32_t id = b.head(s);
      Ord& o = b.ords[id - 1];
      int32_t want = draw_size(rng, c->size_sigma), v = sweeping ? o.vol : std::min(want, o.vol);
      m.price = o.price; m.volume = v; m.ref = id; m.meta = LOBSIM_META(LOBSIM_MSG_MARKET, s ^ 1);
      b.reduce(id, v);
    } else if (type == 1) {
      int32_t off = 1;
      while (off < c->max_offset_ticks && rng.uniform() >= c->geom_p) off++;
      int32_t price = s == 0 ? b.best(1) - c->tick * off : b.best(0) + c->tick * off;
      if (price <= 0) price = c->tick;
      int32_t v = draw_size(rng, c->size_sigma);
      uint32_t id = b.add(s, price, v);
      m.price = price; m.volume = v; m.ref = id; m.meta = LOBSIM_META(LOBSIM_MSG_LIMIT, s);
    }
    msgs[i] = m;
  }
  for (; next_sec <= n_seconds; next_sec++) { b.snapshot(L, snapshots + snap_stride * next_sec); snap_valid[next_sec] = 1; }
  step_off[0] = 0;
  for (uint32_t k = 0; k < n_grid_steps; k++) step_off[k + 1] = step_off[k] + counts[k];
  return LOBSIM_OK;
}
