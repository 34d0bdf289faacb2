// kernels.cuh -- the kernels of the B200-native LOB simulation step.
//
// Kernel design:
//   * one warp per book; a CTA is `warps_per_cta` independent warps, no block-level synchronisation at all;
//   * the book blob lives in shared memory for the whole launch: HBM -> smem by one TMA bulk copy
//     (cp.async.bulk + mbarrier) at the start, smem -> HBM by one bulk store at the end;
//   * historical messages are streamed by TMA bulk copies of 512-byte tiles (32 x 16 B records) into a per-warp
//     double buffer, prefetched one tile ahead; all 32 lanes read a record with one broadcast LDS.128;
//   * book operations are warp-collective (ballot / popc / ffs level and order search, <=32-entry shifts);
//   * features run one per lane, the reward and the rollout tensors are written straight from the kernel.
// fp64 everywhere the reference uses Python floats, compiled with --fmad=false (no FMA contraction).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include "book_flat.cuh"
#include "book_hybrid.cuh"
#include "env.cuh"
#include "lobsim.h"

// ====================================================================================================================
//  PTX helpers: mbarrier + TMA bulk copies (sm_90+; SASS: UBLKCP / SYNCS)
// ====================================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_part(void* dst_gmem, const void* src_smem, uint32_t bytes) { // no commit: several parts, one group
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ====================================================================================================================
//  kernel parameters
// ====================================================================================================================
#ifndef LOBSIM_WS_IN_SMEM
#define LOBSIM_WS_IN_SMEM 0
#endif
#ifndef LOBSIM_ENV_MIN_BLOCKS
#define LOBSIM_ENV_MIN_BLOCKS 3
#endif
#define MSG_TILE 32                      // records per TMA tile
#define MSG_TILE_BYTES (MSG_TILE * 16)
// per-warp scratch behind the message tiles: int2[2 * NA] for update_outer_levels (the agent's re-queued orders), at least
// int[64] for the agent's resting volume per ladder level (agent_prepare) and the 5 action doubles
__host__ __device__ constexpr int scratch_bytes(int NA) { return 2 * NA * 8 > 256 ? 2 * NA * 8 : 256; }

struct AdvParams {
  unsigned char* blobs;             // [n_envs][blob_bytes]
  FeatState* fstate;                // [n_envs][LOBSIM_MAX_FEATURES]
  NormState* nstate;                // [n_envs][LOBSIM_MAX_FEATURES] rolling z-score state, or null (no normalised feature configured)
  const double* beta_tab;           // [2][32]: ln x_k, ln(1 - x_k) at the quote-level midpoints x_k = (k + 0.5) / Q
  double* rings;                    // [n_envs][ring_stride]
  double* rs_ring;                  // RollingSharpe [n_envs][3][LOBSIM_MAX_SHARPE_WINDOW]: two AUM windows + a scratch row of returns, or null
  int32_t* rs_state;                // [n_envs][2][2] = {n_filled, head}
  const lobsim_stream_t* streams;   // device array [n_streams]
  int32_t n_streams;
  lobsim_fill_t* fill_log;          // [n_envs][fill_cap] or null
  int32_t* fill_count;              // [n_envs]
  int32_t fill_cap;
  int32_t n_envs;                   // total envs of the handle
  int32_t n_sel;                    // number of envs this launch works on
  int32_t sel_offset;               // first selection index handled by this launch (tail launches)
  const int32_t* env_ids;           // [n_sel] or null (identity)
  int32_t T;                        // simulation steps to run
  int32_t reset_mode;               // 0 none, 1 book reset only, 2 full env reset (+ warm-up of T steps)
  const int32_t* reset_stream_ids;  // [n_sel]
  const int32_t* reset_steps;       // [n_sel]: book reset: start step; env reset: episode start step
  int32_t agent_kind;               // LOBSIM_AGENT_*
  int32_t out_final_obs_only;       // reset: write obs once, after the warm-up
  int32_t resync_last_only;         // forward_step over several grid steps: resync check only at the end
  int32_t blob_in_global;           // deep books: the blob does not fit in the shared memory of an SM and is worked on in place in HBM / L2
  int32_t* defer_count;             // env HOT kernel: number of (env, step) items handed to the deferred kernel
  int2* defer_list;                 // [n_sel] {selection index, env step to resume at}
  int32_t hybrid;                   // deep layouts: the replay launch runs k_replay_hyb (book_hybrid.cuh); blobs stay sorted in HBM
  int32_t allow_flat;               // fast kernels: books that fit run on -- and are stored in -- the flat order pools (book_flat.cuh)
  const double* actions_in;         // EXTERNAL: [T][n_sel][action_dim]
  double* obs; double* act; double* rew; uint8_t* done; // [T][n_sel][...] (any may be null)
  double* info;                     // [T][n_sel][LOBSIM_INFO_DIM] per-step info series (SimpleInfoCalculator source) or null
  lobsim_agent_t agent;
  const lobsim_agent_t* agents;     // per-env built-in agents [n_sel] (parameter sweeps) or null: every env runs `agent`
  Layout L;
  int32_t warp_smem;                // bytes of shared memory per warp
};

// one row of the per-step info series (InfoCalculators.py:31-59: asset_price, inventory, cash, aum, market_spread) --
// the state the reference's info_calculator sees at the end of HistoricalOrderbookEnvironment.step (HOE.py:175-177)
__device__ __forceinline__ void write_info(double* row, int lane, const StepView& v, double cash, long long inv, uint32_t err) {
  double x = v.price;
  if (lane == LOBSIM_INFO_INVENTORY) x = (double)inv;
  else if (lane == LOBSIM_INFO_CASH) x = cash;
  else if (lane == LOBSIM_INFO_AUM) x = cash + v.price * (double)inv;
  else if (lane == LOBSIM_INFO_MARKET_SPREAD) x = v.have_tops ? (double)(v.bs - v.bb) : NAN;
  else if (lane == LOBSIM_INFO_BEST_BUY) x = v.have_tops ? (double)v.bb : NAN;
  else if (lane == LOBSIM_INFO_BEST_SELL) x = v.have_tops ? (double)v.bs : NAN;
  else if (lane == LOBSIM_INFO_ERR) x = (double)err;
  if (lane < LOBSIM_INFO_DIM) row[lane] = x;
}

// per_step / terminal reward of one env step (HOE.py:170-174); RollingSharpe keeps one AUM window per reward function
template <bool RARE>
__device__ __forceinline__ double step_reward(const AdvParams& p, const lobsim_cfg_t& c, int env, int lane, bool done, double cash0, long long inv0, double p0,
                                              double cash1, long long inv1, double p1, uint32_t& err) {
  double r;
  if (RARE && c.step_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE) {
    SharpeOut o = rolling_sharpe_step(p.rs_ring + ((size_t)env * 3 + 0) * LOBSIM_MAX_SHARPE_WINDOW, p.rs_state + ((size_t)env * 2 + 0) * 2, c.step_reward.asymmetric, cash1 + p1 * (double)inv1, lane, p.rs_ring + ((size_t)env * 3 + 2) * LOBSIM_MAX_SHARPE_WINDOW);
    r = o.reward; err |= o.err;
  } else r = reward_calc(c.step_reward, cash0, inv0, p0, cash1, inv1, p1);
  if (done) {
    if (RARE && c.terminal_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE) {
      SharpeOut o = rolling_sharpe_step(p.rs_ring + ((size_t)env * 3 + 1) * LOBSIM_MAX_SHARPE_WINDOW, p.rs_state + ((size_t)env * 2 + 1) * 2, c.terminal_reward.asymmetric, cash1 + p1 * (double)inv1, lane, p.rs_ring + ((size_t)env * 3 + 2) * LOBSIM_MAX_SHARPE_WINDOW);
      r = o.reward; err |= o.err;
    } else r = reward_calc(c.terminal_reward, cash0, inv0, p0, cash1, inv1, p1);
  }
  return r;
}

// (second * 1e6 + microsecond) of now = t0 + now_step * step_us, in 32-bit arithmetic (t0 % 60 s comes with the stream)
__device__ __forceinline__ int us_in_minute(int t0_mod_min, int now_step, const EnvConst& ec) {
  int u = t0_mod_min + (now_step % ec.steps_per_min) * (int)ec.cfg.step_us;
  return u >= 60000000 ? u - 60000000 : u;
}

__device__ __forceinline__ unsigned char* warp_smem_base(unsigned char* smem, int warp, int warp_smem) { return smem + (size_t)warp * warp_smem; }

// ====================================================================================================================
//  the advance kernel: [reset] + T x ([agent orders] + messages of the step + [resync] + [features, reward])
// ====================================================================================================================
// kEnv: agent + features + rewards (HistoricalOrderbookEnvironment.step); kTrack: fills / flows / agent orders are
// tracked (always with kEnv; the pure replay fast path <false,false> is used when no agent order can be resting).
template <bool kEnv, bool kTrack>
__global__ void __launch_bounds__(128, kEnv ? LOBSIM_ENV_MIN_BLOCKS : 4) k_advance(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sel = blockIdx.x * (blockDim.x >> 5) + warp;
  if (sel >= p.n_sel) return;
  const int env = p.env_ids ? p.env_ids[sel] : sel;
  const lobsim_cfg_t& c = ec.cfg;

  // Deep-book mode (p.blob_in_global): capacities whose blob exceeds the shared memory of an SM (thousands of resting orders per side,
  // rl4mm/orderbook/models.py:66-67 is unbounded) -- the book is worked on IN PLACE in HBM / L2 through the same routines (generic
  // addressing), only the message tiles / scratch / barriers live in shared memory.  Slower per order, but no capacity cliff.
  unsigned char* wbase = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* gblob = p.blobs + (size_t)env * p.L.blob_bytes;
  unsigned char* base = p.blob_in_global ? gblob : wbase;
  Book b; b.blob = base; b.L = p.L; b.lane = lane;
  unsigned char* msgbuf = p.blob_in_global ? wbase : wbase + p.L.blob_bytes;       // 2 x 512 B
  int2* scratch = reinterpret_cast<int2*>(msgbuf + 2 * MSG_TILE_BYTES);             // [2*NA]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + scratch_bytes(p.L.NA)); // 3 barriers

  // ---- book blob: HBM -> shared memory ---------------------------------------------------------------------------
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    if (!p.blob_in_global) {
      mbar_expect_tx(&bars[2], (uint32_t)p.L.blob_bytes);
      tma_load(base, gblob, (uint32_t)p.L.blob_bytes, &bars[2]);
    }
  }
  __syncwarp();
  if (!p.blob_in_global) mbar_wait(&bars[2], 0);

#if LOBSIM_WS_IN_SMEM
  // the uniform per-warp state lives in shared memory (not registers): every lane stores identical values, so plain
  // accesses are race-free; it trades a few broadcast LDS for ~35 registers per thread => more resident books per SM
  WarpState& w = *reinterpret_cast<WarpState*>(reinterpret_cast<unsigned char*>(bars) + 32);
#else
  WarpState w;
#endif
  load_state<kTrack>(b, w);
  w.fill_log = p.fill_log ? p.fill_log + (size_t)env * p.fill_cap : nullptr;
  w.fill_cap = p.fill_cap; w.n_fills = 0;
  const long long inventory_in = w.inventory; const double cash_in = w.cash;
  BookHdr* h = b.hdr();
  const int F = c.n_features;

  // ---- reset prologue ----------------------------------------------------------------------------------------------
  int stream_id = h->stream_id;
  if (kTrack && p.reset_mode) {
    stream_id = p.reset_stream_ids[sel];
    int start = p.reset_steps[sel] - (p.reset_mode == 2 ? c.warmup_steps : 0);
    if (stream_id < 0 || stream_id >= p.n_streams) { stream_id = 0; start = -1; }
    w = init_book_cold(b, w, &ec.cfg, &p.streams[stream_id], stream_id, start);
    if (p.reset_mode == 2 && lane == 0) {
      h->episode_start_step = p.reset_steps[sel];
      if (!c.portfolio_carryover || !h->has_reset) { h->inventory = c.initial_inventory; h->cash = c.initial_cash; }
      h->has_reset = 1;
    }
    __syncwarp();
    w.inventory = h->inventory; w.cash = h->cash;
  }
  const lobsim_stream_t* stp = &p.streams[stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  const long long st_t0_us = kEnv ? stp->t0_us : 0;
  int now_step = h->now_step;
  const long long episode_start_us = st_t0_us + (long long)h->episode_start_step * c.step_us;

  FeatState* fstate_env = p.fstate + (size_t)env * LOBSIM_MAX_FEATURES;
  NormState* nstate_env = p.nstate ? p.nstate + (size_t)env * LOBSIM_MAX_FEATURES : nullptr;
  double* rings_env = p.rings + (size_t)env * ec.ring_stride;
  double feat_cur = 0.0; // Feature.current_value of this lane's feature
  if (kEnv && lane < F) feat_cur = fstate_env[lane].cur;
  const int t0_mod_min = (int)stp->reserved;   // t0_us % 60 s, filled in by lobsim_load_stream

  // price / tops of the current book
  auto tops = [&](StepView& v) {
    v.have_tops = w.nlv0 > 0 && w.nlv1 > 0;
    if (v.have_tops) {
      v.bb = b.lvp(0)[w.nlv0 - 1]; v.bs = b.lvp(1)[w.nlv1 - 1];
      v.bv = best_level_volume(b, 0, w.nlv0); v.sv = best_level_volume(b, 1, w.nlv1);
      double imb; v.price = microprice(v.bb, v.bs, v.bv, v.sv, imb);
    } else { v.bb = v.bs = v.bv = v.sv = 0; v.price = NAN; }
  };

  double price = h->price;
  if (kEnv && p.reset_mode == 2) { // State(...) + _reset_features, HOE.py:152-154,218-221
    StepView v; tops(v);
    v.inventory = w.inventory; v.now_us = st_t0_us + (long long)now_step * c.step_us;
    v.us_in_min = us_in_minute(t0_mod_min, now_step, ec);
    v.n_ext0 = v.n_ext1 = v.vol_ext0 = v.vol_ext1 = v.n_int0 = v.n_int1 = v.vol_int0 = v.vol_int1 = 0;
    price = v.price;
    feat_cur = features_step<true>(&ec, fstate_env, nstate_env, rings_env, lane, v, episode_start_us, 1);
  }

  // ---- message pipeline ----------------------------------------------------------------------------------------------
  const int T = p.T;
  // steps beyond the end of the grid: the steps that exist are run, the first one past the end sets END_OF_STREAM
  const int n_grid = (int)stp->n_grid_steps;
  if ((now_step < 0 || now_step > n_grid) && !w.dead && T > 0) { w.err |= LOBSIM_ERR_END_OF_STREAM; w.dead = 1; }
  unsigned g = 0, g_end_all = 0;
  if (!w.dead && T > 0) { g = __ldg(&st_step_off[now_step]); g_end_all = __ldg(&st_step_off[min((long long)now_step + T, (long long)n_grid)]); }
  // tiles are MSG_TILE-aligned in the global message index space; `rel` tile r lives in buffer r & 1 and completes
  // phase (r >> 1) & 1 of that buffer's mbarrier.  Tiles are issued and consumed strictly in order.
  const unsigned tile0 = g / MSG_TILE;
  unsigned next_issue = 0, next_wait = 0;
  auto issue_tile = [&]() { // uniform; lane 0 talks to the TMA unit
    const unsigned first = (tile0 + next_issue) * MSG_TILE;
    if (first >= g_end_all) return;
    if (lane == 0) {
      const unsigned n_total = (unsigned)stp->n_msgs;
      unsigned cnt = n_total - first < MSG_TILE ? n_total - first : MSG_TILE;
      uint64_t* bar = &bars[next_issue & 1];
      mbar_expect_tx(bar, cnt * 16);
      tma_load(msgbuf + (next_issue & 1) * MSG_TILE_BYTES, st_msgs + first, cnt * 16, bar);
    }
    next_issue++;
  };
  auto wait_tile = [&]() { // wait for relative tile `next_wait`
    mbar_wait(&bars[next_wait & 1], (next_wait >> 1) & 1);
    next_wait++;
  };
  if (g < g_end_all) { issue_tile(); issue_tile(); }
  __syncwarp();

  AgentGen gen; gen.side = 3;
  bool agent_phase = false;
#pragma unroll 1
  for (int t = 0; t < T; t++) {
    // ---- agent: obs -> action -> orders (processed before the step's history, OrderbookSimulator.py:76-77) -------
    double cash0 = w.cash, p0 = price; long long inv0 = w.inventory;
    if (kEnv) {
      reset_flow(w);
      if (p.agent_kind != LOBSIM_AGENT_NONE) {
        double* act_sm = reinterpret_cast<double*>(scratch); // 5 doubles of per-warp scratch
        if (p.agent_kind == LOBSIM_AGENT_EXTERNAL) {
          const double* a = p.actions_in + ((size_t)t * p.n_sel + sel) * ec.action_dim;
          if (lane < 5) act_sm[lane] = lane < ec.action_dim ? __ldg(&a[lane]) : 0.0;
        } else {
          const lobsim_agent_t* agp = p.agents ? p.agents + sel : &p.agent;
          const double inv_obs = __shfl_sync(FULL_MASK, feat_cur, agp->inventory_index & 31);
          if (lane == 0) agent_action_cold(agp, inv_obs, act_sm, env, now_step);
        }
        __syncwarp();
        const double a0 = act_sm[0], a1 = act_sm[1], a2 = act_sm[2], a3 = act_sm[3], a4 = act_sm[4];
        const double mine = lane < 5 ? act_sm[lane] : 0.0;
        __syncwarp();
        if (p.act && p.agent_kind != LOBSIM_AGENT_EXTERNAL && lane < ec.action_dim) p.act[((size_t)t * p.n_sel + sel) * ec.action_dim + lane] = mine;
        if (p.obs && !p.out_final_obs_only && c.inc_prev_action_in_obs && lane < ec.action_dim)
          p.obs[((size_t)t * p.n_sel + sel) * ec.obs_dim + F + lane] = mine; // get_observation(action), HOE.py:171
        gen = agent_prepare(b, w.nlv0 ? b.lvp(0)[w.nlv0 - 1] : INT32_MIN, w.nlv1 ? b.lvp(1)[w.nlv1 - 1] : INT32_MAX, w.nag0, w.nag1, w.inventory, &ec, a0, a1, a2, a3, a4, p.beta_tab, reinterpret_cast<int*>(scratch));
        w.err |= gen.err_out; if (gen.dead_out) w.dead = 1;
        agent_phase = true;
      }
    }
    // ---- the step's orders: the agent's first, then the historical messages of (now, now + step] ------------------
    {
      if (!w.dead && now_step >= n_grid) { w.err |= LOBSIM_ERR_END_OF_STREAM; w.dead = 1; }
      const unsigned g_step_end = w.dead ? g : __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
      for (;;) {
        int type, side, price, vol; uint32_t ref; bool is_agent;
        if (kEnv && agent_phase) {
          if (!agent_next(b, w, gen, type, side, price, vol, ref)) { agent_phase = false; continue; }
          is_agent = true;
        } else {
          if (w.dead || g >= g_step_end) break;
          const unsigned tile = g / MSG_TILE - tile0;
          if (tile == next_wait) wait_tile();
          const uint4 m = *reinterpret_cast<const uint4*>(msgbuf + (tile & 1) * MSG_TILE_BYTES + (g % MSG_TILE) * 16);
          price = (int)m.x; vol = (int)m.y; ref = m.z; type = (int)LOBSIM_META_TYPE(m.w); side = (int)LOBSIM_META_DIR(m.w);
          is_agent = false;
          g++;
          if (g % MSG_TILE == 0) { __syncwarp(); issue_tile(); } // tile consumed: refill its buffer
        }
        process_order<kTrack>(b, w, type, side, price, vol, ref, is_agent);
      }
    }
    if (!w.dead) {
      now_step++;
      // ---- resync, OrderbookSimulator.py:86-87 ----------------------------------------------------------------------
      if (c.resync && (!p.resync_last_only || t == T - 1)) {
        long long rel = (long long)now_step * c.step_us;
        if (rel % 1000000 == 0 && near_exiting(b, w, c)) {
          long long sec = rel / 1000000;
          if (sec <= (long long)stp->n_seconds && stp->snap_valid[sec]) {
            w = update_outer_levels<kTrack>(b, w, &ec.cfg, stp->snapshots + (size_t)sec * 2 * c.n_levels * 2, scratch);
                    }
        }
      }
    }
    if (kEnv) {
      // ---- update_internal_state + _update_features + reward, HOE.py:163-178,199-204 -------------------------------
      StepView v; tops(v);
      if (!v.have_tops) w.err |= LOBSIM_ERR_EMPTY_BOOK;
      price = v.price;
      v.inventory = w.inventory; v.now_us = st_t0_us + (long long)now_step * c.step_us;
      v.us_in_min = us_in_minute(t0_mod_min, now_step, ec);
      v.n_ext0 = w.n_ext0; v.n_ext1 = w.n_ext1; v.vol_ext0 = w.vol_ext0; v.vol_ext1 = w.vol_ext1;
      v.n_int0 = w.n_int0; v.n_int1 = w.n_int1; v.vol_int0 = w.vol_int0; v.vol_int1 = w.vol_int1;
      feat_cur = features_step<true>(&ec, fstate_env, nstate_env, rings_env, lane, v, episode_start_us, 0);
      const bool write_now = !p.out_final_obs_only || t == T - 1;
      if (p.obs && write_now) {
        double* o = p.obs + ((size_t)(p.out_final_obs_only ? 0 : t) * p.n_sel + sel) * ec.obs_dim;
        if (lane < F) o[lane] = feat_cur;
        if (c.inc_prev_action_in_obs && lane < ec.action_dim && (p.agent_kind == LOBSIM_AGENT_NONE || p.out_final_obs_only)) o[F + lane] = 0.0;
      }
      if (p.agent_kind != LOBSIM_AGENT_NONE) {
        const bool d = now_step >= h->episode_start_step + c.episode_steps; // terminal_time - now < step/2, HOE.py:172
        const double r = step_reward<true>(p, c, env, lane, d, cash0, inv0, p0, w.cash, w.inventory, price, w.err);
        if (lane == 0) {
          if (p.rew) p.rew[(size_t)t * p.n_sel + sel] = r;
          if (p.done) p.done[(size_t)t * p.n_sel + sel] = d ? 1 : 0;
        }
        if (p.info) write_info(p.info + ((size_t)t * p.n_sel + sel) * LOBSIM_INFO_DIM, lane, v, w.cash, w.inventory, w.err);
      }
    }
  }
  if (kEnv && T == 0 && p.obs && p.reset_mode == 2) { // reset with no warm-up: obs straight after _reset_features
    double* o = p.obs + (size_t)sel * ec.obs_dim;
    if (lane < F) o[lane] = feat_cur;
    if (c.inc_prev_action_in_obs && lane < ec.action_dim) o[F + lane] = 0.0;
  }

  while (next_wait < next_issue) wait_tile(); // drain TMA loads still in flight (only after an aborted episode)

  // ---- write back ----------------------------------------------------------------------------------------------------
  // OrderbookSimulator.forward_step only returns the fills: the portfolio belongs to the env (HOE.py:280-289)
  if (!kEnv && !p.reset_mode) { w.inventory = inventory_in; w.cash = cash_in; }
  if (lane == 0) {
    h->now_step = now_step;
    h->price = price;
    if (p.fill_count) p.fill_count[env] = w.n_fills;
  }
  store_state<kTrack>(b, w);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0 && !p.blob_in_global) { tma_store(gblob, base, (uint32_t)p.L.blob_bytes); tma_store_wait(); }
  __syncwarp();
}

// ====================================================================================================================
//  the replay fast kernel: T x (messages of the step + [resync]) with the straight-line message path of book_fast.cuh.
//  Used by lobsim_replay when no fill log is requested, no agent order can be resting and the book capacities match
//  one of the compiled StaticLayouts.  72 registers => 7 CTAs x 4 warps = 28 books resident per SM.
// ====================================================================================================================
template <class LT>
__global__ void __launch_bounds__(128, 7) k_replay_fast(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  if (env >= p.n_sel) return;
  const lobsim_cfg_t& c = ec.cfg;
  // a replay book holds no agent orders: the agent tables at the end of the blob stay in HBM (deep books are shared-memory bound:
  // 128/1536/64 capacities = 8 instead of 7 resident books per SM)
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::agent_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(msgbuf + 2 * MSG_TILE_BYTES);
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)LT::agent_off);
    tma_load(base, gblob, (uint32_t)LT::agent_off, &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);

  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  BookHdr* h = reinterpret_cast<BookHdr*>(base);
  FastState f; f.err = h->err; f.dead = h->dead; f.bail = 0; f.bail_vol = 0;
  fast_refresh_best(fb, f);
  const lobsim_stream_t* stp = &p.streams[h->stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  int now_step = h->now_step;
  const int T = p.T;
  const int n_grid = (int)stp->n_grid_steps;   // steps beyond the end: the existing ones are run, the first one past the end sets END_OF_STREAM
  if ((now_step < 0 || now_step > n_grid) && !f.dead && T > 0) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; }
  unsigned g = 0, g_end_all = 0;
  if (!f.dead && T > 0) { g = __ldg(&st_step_off[now_step]); g_end_all = __ldg(&st_step_off[min((long long)now_step + T, (long long)n_grid)]); }
  const unsigned tile0 = g / MSG_TILE;
  unsigned next_issue = 0, next_wait = 0;
  auto issue_tile = [&]() {
    const unsigned first = (tile0 + next_issue) * MSG_TILE;
    if (first >= g_end_all) return;
    if (lane == 0) {
      const unsigned n_total = (unsigned)stp->n_msgs;
      const unsigned cnt = n_total - first < MSG_TILE ? n_total - first : MSG_TILE;
      uint64_t* bar = &bars[next_issue & 1];
      mbar_expect_tx(bar, cnt * 16);
      tma_load(msgbuf + (next_issue & 1) * MSG_TILE_BYTES, st_msgs + first, cnt * 16, bar);
    }
    next_issue++;
  };
  auto wait_tile = [&]() { mbar_wait(&bars[next_wait & 1], (next_wait >> 1) & 1); next_wait++; };
  if (g < g_end_all) { issue_tile(); issue_tile(); }
  __syncwarp();
  const int steps_per_sec = ec.steps_per_sec;
  int sub = now_step >= 0 ? now_step % steps_per_sec : 0;   // position inside the current second

#pragma unroll 1
  for (int t = 0; t < T && !f.dead; t++) {
    if (now_step >= n_grid) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; break; }
    const unsigned g_step_end = __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
    while (g < g_step_end) {
      const unsigned tile = g / MSG_TILE - tile0;
      if (tile == next_wait) wait_tile();
      const unsigned tile_end = (g / MSG_TILE + 1) * MSG_TILE;
      const unsigned lim = g_step_end < tile_end ? g_step_end : tile_end;
      const uint4* mp = reinterpret_cast<const uint4*>(msgbuf + (tile & 1) * MSG_TILE_BYTES) + (g % MSG_TILE);
      const unsigned cnt = lim - g;
#pragma unroll 1
      for (unsigned i = 0; i < cnt; i++) {
        const uint4 m = mp[i];
        fast_message(fb, f, (int)m.x, (int)m.y, m.z, m.w);
        if (f.dead) break;
      }
      if (f.dead) break;
      g = lim;
      if (g == tile_end) { __syncwarp(); issue_tile(); }
    }
    if (f.dead) break;
    now_step++;
    if (++sub == steps_per_sec) {                            // whole second: outer-level resync, OrderbookSimulator.py:86-87
      sub = 0;
      if (c.resync && (!p.resync_last_only || t == T - 1)) {
        const double prop = ec.outer_prop;
        const double bb = f.best0 == INT32_MIN ? 0.0 : (double)f.best0;
        const double bs = f.best1 == INT32_MAX ? (double)INFINITY : (double)f.best1;
        if (bb < (double)h->min_buy + prop * (double)h->init_buy_range || bs > (double)h->max_sell - prop * (double)h->init_sell_range) {
          const int sec = now_step / steps_per_sec;
          if (sec <= (int)stp->n_seconds && stp->snap_valid[sec]) {
            const int32_t* row = stp->snapshots + (size_t)sec * 2 * c.n_levels * 2;
            fast_resync(fb, f, row, c.n_levels);   // straight-line: a replay book holds no agent orders
          }
        }
      }
    }
  }
  while (next_wait < next_issue) wait_tile();                // drain TMA loads still in flight (aborted episode)
  __syncwarp();
  if (lane == 0) { h->now_step = now_step; h->err = f.err; h->dead = f.dead; }
  __syncwarp();
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) { tma_store(gblob, base, (uint32_t)LT::agent_off); tma_store_wait(); }
  __syncwarp();
}

__device__ __forceinline__ bool hdr_is_flat(const BookHdr* h) { return h->cnt[0][0] < 0; }
// ====================================================================================================================
//  the replay HYBRID kernel (deep layouts): k_replay_fast with the hybrid book of book_hybrid.cuh -- the levels near the touch in
//  a flat order pool, the rest in the sorted arrays -- for every book that can take that form, the sorted straight-line path for
//  the others.  The blob in HBM is the canonical sorted layout on entry and on exit.
// ====================================================================================================================
template <class LT>
__global__ void __launch_bounds__(128, 3) k_replay_hyb(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  if (env >= p.n_sel) return;
  const lobsim_cfg_t& c = ec.cfg;
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::agent_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(msgbuf + 2 * MSG_TILE_BYTES);
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)LT::agent_off);
    tma_load(base, gblob, (uint32_t)LT::agent_off, &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);

  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  BookHdr* h = reinterpret_cast<BookHdr*>(base);
  FastState f; f.err = h->err; f.dead = h->dead; f.bail = 0; f.bail_vol = 0;
  if (hdr_is_flat(h)) flat_leave_fn<LT>(base, lane, h->cnt[0][1], h->cnt[1][1]);   // (a blob stored by the flat kernels)
  fast_refresh_best(fb, f);
  HybState hs; hs.n0 = hs.n1 = 0; hs.floor0 = INT32_MIN; hs.floor1 = INT32_MAX; hs.seq = 0;
  bool hyb = false;
  const lobsim_stream_t* stp = &p.streams[h->stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  int now_step = h->now_step;
  const int T = p.T;
  const int n_grid = (int)stp->n_grid_steps;
  if ((now_step < 0 || now_step > n_grid) && !f.dead && T > 0) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; }
  unsigned g = 0, g_end_all = 0;
  if (!f.dead && T > 0) { g = __ldg(&st_step_off[now_step]); g_end_all = __ldg(&st_step_off[min((long long)now_step + T, (long long)n_grid)]); }
  auto try_enter = [&]() {
    if (hyb_enter(fb, f, hs)) hyb = true;
    else if (hs.n0 | hs.n1) { hyb_leave(fb, f, hs); hs.n0 = hs.n1 = 0; }      // one side was moved, the other could not be: put it back
  };
  if (g < g_end_all) try_enter();
  const unsigned tile0 = g / MSG_TILE;
  unsigned next_issue = 0, next_wait = 0;
  auto issue_tile = [&]() {
    const unsigned first = (tile0 + next_issue) * MSG_TILE;
    if (first >= g_end_all) return;
    if (lane == 0) {
      const unsigned n_total = (unsigned)stp->n_msgs;
      const unsigned cnt = n_total - first < MSG_TILE ? n_total - first : MSG_TILE;
      uint64_t* bar = &bars[next_issue & 1];
      mbar_expect_tx(bar, cnt * 16);
      tma_load(msgbuf + (next_issue & 1) * MSG_TILE_BYTES, st_msgs + first, cnt * 16, bar);
    }
    next_issue++;
  };
  auto wait_tile = [&]() { mbar_wait(&bars[next_wait & 1], (next_wait >> 1) & 1); next_wait++; };
  if (g < g_end_all) { issue_tile(); issue_tile(); }
  __syncwarp();
  const int steps_per_sec = ec.steps_per_sec;
  int sub = now_step >= 0 ? now_step % steps_per_sec : 0;

#pragma unroll 1
  for (int t = 0; t < T && !f.dead; t++) {
    if (now_step >= n_grid) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; break; }
    const unsigned g_step_end = __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
    while (g < g_step_end) {
      const unsigned tile = g / MSG_TILE - tile0;
      if (tile == next_wait) wait_tile();
      const unsigned tile_end = (g / MSG_TILE + 1) * MSG_TILE;
      const unsigned lim = g_step_end < tile_end ? g_step_end : tile_end;
      const uint4* mp = reinterpret_cast<const uint4*>(msgbuf + (tile & 1) * MSG_TILE_BYTES) + (g % MSG_TILE);
      const unsigned cnt = lim - g;
      unsigned i = 0;
      int redo_vol = 0;
      // the volumes of the whole segment (<= 32 messages, one per lane) at once: a segment with a bad one (never in valid data) runs on
      // the sorted book, whose routines test every order (assert order.volume > 0, Exchange.py:59-60)
      if (hyb && __any_sync(FULL_MASK, (unsigned)lane < cnt && (int)mp[(unsigned)lane < cnt ? lane : 0].y <= 0)) {
        hyb_leave(fb, f, hs); hyb = false; hs.n0 = hs.n1 = 0;
      }
      if (hyb) {
#pragma unroll 1
        for (; i < cnt; i++) {
          const uint4 m = mp[i];
          if (hyb_message<LT>(fb, f, hs, (int)m.x, (int)m.y, m.z, m.w)) break;
        }
        if (f.bail) {                                        // this book cannot stay hybrid: back to the sorted form
          const int why = f.bail, rest = f.bail_vol;
          f.bail = 0; f.bail_vol = 0;
          hyb_leave(fb, f, hs); hyb = false; hs.n0 = hs.n1 = 0;
          if (why == FLAT_BAIL_FULL) redo_vol = rest;        // the unfinished part of message i runs on the sorted book
          else i++;
        }
      }
      if (!hyb && !f.dead) {
#pragma unroll 1
        for (; i < cnt; i++) {
          const uint4 m = mp[i];
          const int v = redo_vol ? redo_vol : (int)m.y;
          redo_vol = 0;
          fast_message(fb, f, (int)m.x, v, m.z, m.w);
          if (f.dead) break;
        }
      }
      if (f.dead) break;
      g = lim;
      if (g == tile_end) { __syncwarp(); issue_tile(); }
    }
    if (f.dead) break;
    now_step++;
    if (hyb) hyb_top_up(fb, f, hs);
    if (++sub == steps_per_sec) {                            // whole second: outer-level resync, OrderbookSimulator.py:86-87
      sub = 0;
      if (c.resync && (!p.resync_last_only || t == T - 1)) {
        const double prop = ec.outer_prop;
        const double bb = f.best0 == INT32_MIN ? 0.0 : (double)f.best0;
        const double bs = f.best1 == INT32_MAX ? (double)INFINITY : (double)f.best1;
        if (bb < (double)h->min_buy + prop * (double)h->init_buy_range || bs > (double)h->max_sell - prop * (double)h->init_sell_range) {
          const int sec = now_step / steps_per_sec;
          if (sec <= (int)stp->n_seconds && stp->snap_valid[sec]) {
            const int32_t* row = stp->snapshots + (size_t)sec * 2 * c.n_levels * 2;
            if (hyb && !flat_resync_needed(h, row, c.n_levels, lane)) hyb_update_trackers(fb, hs);
            else {
              if (hyb) { hyb_leave(fb, f, hs); hyb = false; hs.n0 = hs.n1 = 0; }
              fast_resync(fb, f, row, c.n_levels);
            }
          }
        }
      }
      if (!hyb) try_enter();
    }
  }
  if (hyb) hyb_leave(fb, f, hs);
  while (next_wait < next_issue) wait_tile();
  __syncwarp();
  if (lane == 0) { h->now_step = now_step; h->err = f.err; h->dead = f.dead; }
  __syncwarp();
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) { tma_store(gblob, base, (uint32_t)LT::agent_off); tma_store_wait(); }
  __syncwarp();
}

// ---- the flat form of a blob in HBM (book_flat.cuh): header with cnt[0] = {-1, n0}, cnt[1] = {seq, n1}; per side the order pool
//      (n x 16 B at the start of the side's order array); the agent tables as always.  Parts: lanes 0-1 the pools, 2-7 the agent tables.
template <class LT, bool LOAD>
__device__ __forceinline__ uint32_t flat_body_copy(unsigned char* sm, unsigned char* gm, uint64_t* bar, int lane, int n0, int n1) {
  const BookHdr* h = reinterpret_cast<const BookHdr*>(sm);
  uint32_t off = 0, len = 0;
  if (lane < 2) { off = LT::side_off + lane * LT::side_stride + LT::ord_off; len = (uint32_t)(lane ? n1 : n0) * 16u; }
  else if (lane < 8) {
    const int s = (lane - 2) / 3, part = (lane - 2) % 3;
    off = LT::agent_off + s * LT::NA * 12 + part * LT::NA * 4;
    len = (min((uint32_t)h->nag[s], (uint32_t)LT::NA) * 4 + 15) & ~15u;
  }
  const uint32_t total = __reduce_add_sync(FULL_MASK, len);
  if (total == 0) return 0;
  if (LOAD) { if (lane == 0) mbar_expect_tx(bar, total); __syncwarp(); }
  if (len) { if (LOAD) tma_load(sm + off, gm + off, len, bar); else tma_store_part(gm + off, sm + off, len); }
  return total;
}
// FlatState of a flat blob whose header and pools are in shared memory; the best prices are recomputed from the pools
template <class LT>
__device__ __forceinline__ void flat_adopt(unsigned char* sm, int lane, FastState& f, FlatState& fs) {
  const BookHdr* h = reinterpret_cast<const BookHdr*>(sm);
  fs.n0 = h->cnt[0][1]; fs.n1 = h->cnt[1][1]; fs.seq = (uint32_t)h->cnt[1][0];
  f.best0 = flat_best_scan<LT, 0>(sm, lane, fs.n0);
  f.best1 = flat_best_scan<LT, 1>(sm, lane, fs.n1);
}
__device__ __forceinline__ void flat_mark(BookHdr* h, const FlatState& fs) {   // lane 0, before the blob goes back to HBM
  h->cnt[0][0] = -1; h->cnt[0][1] = fs.n0; h->cnt[1][0] = (int32_t)fs.seq; h->cnt[1][1] = fs.n1;
}

// Every flat blob back to the canonical sorted layout, in place (run by the host before anything that reads the level arrays:
// the general kernels, k_process_orders, the sorted-only replay kernel, the L3 dump).  One warp per book; sorted books: nothing.
template <class LT>
__global__ void __launch_bounds__(128) k_to_sorted(unsigned char* blobs, int n_envs) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  if (env >= n_envs) return;
  unsigned char* gblob = blobs + (size_t)env * LT::blob_bytes;
  if (!hdr_is_flat(reinterpret_cast<const BookHdr*>(gblob))) return;
  unsigned char* base = smem + (size_t)warp * LT::blob_bytes;
  if (lane < 8) reinterpret_cast<uint4*>(base)[lane] = reinterpret_cast<const uint4*>(gblob)[lane];   // the 128-byte header
  __syncwarp();
  const BookHdr* h = reinterpret_cast<const BookHdr*>(base);
  FlatState fs; fs.n0 = h->cnt[0][1]; fs.n1 = h->cnt[1][1]; fs.seq = (uint32_t)h->cnt[1][0];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int n = s ? fs.n1 : fs.n0;
    const uint4* gp = flat_pool<LT>(gblob, s); uint4* sp = flat_pool<LT>(base, s);
    for (int i = lane; i < n; i += 32) sp[i] = gp[i];
  }
  __syncwarp();
  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  flat_leave(fb, fs);
  __syncwarp();
  if (lane == 0) reinterpret_cast<uint4*>(gblob)[0] = reinterpret_cast<const uint4*>(base)[0];          // cnt[2][2]
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int2 c = *fb.cnt(s);
    unsigned char* ss = fb.side(s); unsigned char* gs = gblob + LT::side_off + s * LT::side_stride;
    for (int i = lane; i < c.x; i += 32) { fb.P(gs)[i] = fb.P(ss)[i]; fb.LE(gs)[i] = fb.LE(ss)[i]; }
    for (int i = lane; i < c.y; i += 32) fb.O(gs)[i] = fb.O(ss)[i];
  }
}

// ====================================================================================================================
//  the replay FLAT kernel: k_replay_fast with the flat order pools of book_flat.cuh for every book that fits them (at most
//  FLAT_CAP resting orders per side), the sorted straight-line path for the others -- per book, re-decided every second.
//  The blob in HBM is the canonical sorted layout on entry and on exit.
// ====================================================================================================================
#ifndef LOBSIM_REPLAY_FLAT_MIN_BLOCKS
#define LOBSIM_REPLAY_FLAT_MIN_BLOCKS 7      // 72 registers: 28 resident warps per SM with the 64/256/32 blob (8 = 64 registers: A/B in profiles/)
#endif
template <class LT>
__global__ void __launch_bounds__(128, LOBSIM_REPLAY_FLAT_MIN_BLOCKS) k_replay_flat(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + warp;
  if (env >= p.n_sel) return;
  const lobsim_cfg_t& c = ec.cfg;
  // a replay book holds no agent orders: the agent tables at the end of the blob stay in HBM (deep books are shared-memory bound:
  // 128/1536/64 capacities = 8 instead of 7 resident books per SM)
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::agent_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(msgbuf + 2 * MSG_TILE_BYTES);
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)LT::agent_off);
    tma_load(base, gblob, (uint32_t)LT::agent_off, &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);

  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  BookHdr* h = reinterpret_cast<BookHdr*>(base);
  FastState f; f.err = h->err; f.dead = h->dead; f.bail = 0; f.bail_vol = 0;
  if (!hdr_is_flat(h)) fast_refresh_best(fb, f);
  FlatState fs; fs.n0 = fs.n1 = 0; fs.seq = 0;
  bool flat = false;
  if (hdr_is_flat(h)) { flat_adopt<LT>(base, lane, f, fs); flat = true; }   // the blob was stored in the flat form
  const lobsim_stream_t* stp = &p.streams[h->stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  int now_step = h->now_step;
  const int T = p.T;
  const int n_grid = (int)stp->n_grid_steps;
  if ((now_step < 0 || now_step > n_grid) && !f.dead && T > 0) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; }
  unsigned g = 0, g_end_all = 0;
  if (!f.dead && T > 0) { g = __ldg(&st_step_off[now_step]); g_end_all = __ldg(&st_step_off[min((long long)now_step + T, (long long)n_grid)]); }
  if (!flat && g < g_end_all && flat_fits(fb, 0)) { flat_enter(fb, fs); flat = true; }
  const unsigned tile0 = g / MSG_TILE;
  unsigned next_issue = 0, next_wait = 0;
  auto issue_tile = [&]() {
    const unsigned first = (tile0 + next_issue) * MSG_TILE;
    if (first >= g_end_all) return;
    if (lane == 0) {
      const unsigned n_total = (unsigned)stp->n_msgs;
      const unsigned cnt = n_total - first < MSG_TILE ? n_total - first : MSG_TILE;
      uint64_t* bar = &bars[next_issue & 1];
      mbar_expect_tx(bar, cnt * 16);
      tma_load(msgbuf + (next_issue & 1) * MSG_TILE_BYTES, st_msgs + first, cnt * 16, bar);
    }
    next_issue++;
  };
  auto wait_tile = [&]() { mbar_wait(&bars[next_wait & 1], (next_wait >> 1) & 1); next_wait++; };
  if (g < g_end_all) { issue_tile(); issue_tile(); }
  __syncwarp();
  const int steps_per_sec = ec.steps_per_sec;
  int sub = now_step >= 0 ? now_step % steps_per_sec : 0;

#pragma unroll 1
  for (int t = 0; t < T && !f.dead; t++) {
    if (now_step >= n_grid) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; break; }
    const unsigned g_step_end = __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
    while (g < g_step_end) {
      const unsigned tile = g / MSG_TILE - tile0;
      if (tile == next_wait) wait_tile();
      const unsigned tile_end = (g / MSG_TILE + 1) * MSG_TILE;
      const unsigned lim = g_step_end < tile_end ? g_step_end : tile_end;
      const uint4* mp = reinterpret_cast<const uint4*>(msgbuf + (tile & 1) * MSG_TILE_BYTES) + (g % MSG_TILE);
      const unsigned cnt = lim - g;
      unsigned i = 0;
      if (flat) {
        // the volumes of the whole segment (<= 32 messages, one per lane) are checked at once; the per-message test is one predicate
        const bool bad_vol = __any_sync(FULL_MASK, (unsigned)lane < cnt && (int)mp[(unsigned)lane < cnt ? lane : 0].y <= 0);
        unsigned off = 0;                                    // (the loop walks a byte offset: no index arithmetic per message)
        const unsigned off_end = cnt * 16u;
        auto run = [&](auto vchk) {                          // (two copies of the loop: the one that runs has no volume test)
#pragma unroll 1
          for (; off != off_end; off += 16u) {
            const uint4 m = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(mp) + off);
            if (flat_message<LT, decltype(vchk)::value>(base, lane, f, fs, (int)m.x, (int)m.y, m.z, m.w)) return true;   // pool full
                                                             // (this message runs on the sorted book) or f.dead: EmptyOrderbookError
          }
          return false;
        };
        const bool stopped = __builtin_expect(bad_vol, 0) ? run(std::true_type{}) : run(std::false_type{});
        i = off >> 4;
        f.bail = 0;                                          // (not read: `stopped` without f.dead is the full pool)
        if (stopped && !f.dead) { flat_leave(fb, fs); flat = false; }
      }
      if (!flat && !f.dead) {
#pragma unroll 1
        for (; i < cnt; i++) {
          const uint4 m = mp[i];
          fast_message(fb, f, (int)m.x, (int)m.y, m.z, m.w);
          if (f.dead) break;
        }
      }
      if (f.dead) break;
      g = lim;
      if (g == tile_end) { __syncwarp(); issue_tile(); }
    }
    if (f.dead) break;
    now_step++;
    if (++sub == steps_per_sec) {                            // whole second: outer-level resync, OrderbookSimulator.py:86-87
      sub = 0;
      if (c.resync && (!p.resync_last_only || t == T - 1)) {
        const double prop = ec.outer_prop;
        const double bb = f.best0 == INT32_MIN ? 0.0 : (double)f.best0;
        const double bs = f.best1 == INT32_MAX ? (double)INFINITY : (double)f.best1;
        if (bb < (double)h->min_buy + prop * (double)h->init_buy_range || bs > (double)h->max_sell - prop * (double)h->init_sell_range) {
          const int sec = now_step / steps_per_sec;
          if (sec <= (int)stp->n_seconds && stp->snap_valid[sec]) {
            const int32_t* row = stp->snapshots + (size_t)sec * 2 * c.n_levels * 2;
            if (flat && !flat_resync_needed(h, row, c.n_levels, lane)) flat_update_trackers<LT>(base, lane, fs);
            else {
              if (flat) { flat_leave(fb, fs); flat = false; }
              fast_resync(fb, f, row, c.n_levels);
            }
          }
        }
      }
      if (flat && fs.seq > 0xE0000000u) { flat_leave(fb, fs); flat = false; }   // renumber the time-priority stamps long before they wrap
      if (!flat && flat_fits(fb, 8)) { flat_enter(fb, fs); flat = true; }   // back to the flat pools once the book has shrunk
    }
  }
  if (flat && !p.allow_flat) { flat_leave(fb, fs); flat = false; }   // the caller wants the canonical sorted layout in HBM
  while (next_wait < next_issue) wait_tile();                // drain TMA loads still in flight (aborted episode)
  __syncwarp();
  if (lane == 0) { h->now_step = now_step; h->err = f.err; h->dead = f.dead; if (flat) flat_mark(h, fs); }
  __syncwarp();
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) { tma_store(gblob, base, (uint32_t)LT::agent_off); tma_store_wait(); }
  __syncwarp();
}

// ====================================================================================================================
//  the env fast kernel: k_advance<true,true> with the straight-line tracked order path (book_fast.cuh fast_order<LT,true>)
//  and all per-env scalars (counters, portfolio, per-step flow) in the shared-memory header instead of registers.
// ====================================================================================================================
static __device__ __noinline__ uint32_t reset_book_cold(unsigned char* blob, const Layout* L, int lane, const lobsim_cfg_t* c, const lobsim_stream_t* st, int stream_id, int start_step) {
  Book b; b.blob = blob; b.L = *L; b.lane = lane;
  WarpState w;
  __syncwarp();
  load_state<true>(b, w);
  w.fill_log = nullptr; w.fill_cap = 0; w.n_fills = 0;
  init_book_from_snapshot(b, w, *c, *st, stream_id, start_step);
  store_state<true>(b, w);
  return pack_errdead(w.err, w.dead);
}

// ---- the OCCUPIED part of a book blob (env kernels: one blob round trip per env step when a policy runs between steps) ----
// Only what the counters in the header say is in use moves between HBM and shared memory: per side the level prices, the
// level ends and the order entries up to {nlv, nord}, and the agent tables up to nag -- about 1.5 KB of a 6.5 KB blob for a
// 10-level book.  Sizes are rounded up to the 16-byte granularity of cp.async.bulk; every array of a StaticLayout starts
// 16-byte aligned and its capacity is a multiple of 16 bytes, so a rounded-up part never leaves its array.
// Warp-collective: lane k < 12 owns part k (its offset and length are a few integer operations on the header counters) and
// issues its own bulk copy; lane 0 posts the byte total on the mbarrier first.  LOAD: the header has already arrived in shared
// memory; returns the bytes expected on `bar` (0: nothing was issued).
template <class LT, bool LOAD>
__device__ __forceinline__ uint32_t blob_body_copy(unsigned char* sm, unsigned char* gm, uint64_t* bar, int lane) {
  static_assert(LT::NL % 8 == 0 && LT::NO % 2 == 0 && LT::NA % 4 == 0 && LT::ord_off % 16 == 0 && LT::side_stride % 16 == 0 && LT::agent_off % 16 == 0,
                "StaticLayout arrays must be 16-byte aligned for the partial blob copies");
  const BookHdr* h = reinterpret_cast<const BookHdr*>(sm);
  const int s = lane >= 6 ? 1 : 0, part = lane - s * 6;     // parts 0-2: level prices, level ends, orders; 3-5: agent price / volume / id
  uint32_t off = 0, len = 0;
  if (lane < 12) {
    const uint32_t so = LT::side_off + s * LT::side_stride, ao = LT::agent_off + s * LT::NA * 12;
    if (part == 0)      { off = so;                 len = min((uint32_t)h->cnt[s][0], (uint32_t)LT::NL) * 4; }
    else if (part == 1) { off = so + LT::lvend_off; len = min((uint32_t)h->cnt[s][0], (uint32_t)LT::NL) * 2; }
    else if (part == 2) { off = so + LT::ord_off;   len = min((uint32_t)h->cnt[s][1], (uint32_t)LT::NO) * 8; }
    else                { off = ao + (part - 3) * LT::NA * 4; len = min((uint32_t)h->nag[s], (uint32_t)LT::NA) * 4; }
    len = (len + 15) & ~15u;
  }
  const uint32_t total = __reduce_add_sync(FULL_MASK, len);
  if (total == 0) return 0;
  if (LOAD) { if (lane == 0) mbar_expect_tx(bar, total); __syncwarp(); }
  if (len) { if (LOAD) tma_load(sm + off, gm + off, len, bar); else tma_store_part(gm + off, sm + off, len); }
  return total;
}

#ifndef LOBSIM_ENVFAST_WARPS
#define LOBSIM_ENVFAST_WARPS 8    // warps per CTA of the env fast kernel: the warps of a CTA move through the phases of a step together, so
                                  // fewer, larger CTAs = fewer phases in flight per SM = a warmer instruction cache ("no instruction" is the
                                  // top stall).  Measured at 80 registers / 24 resident warps per SM (profiles/r02_env_ab.txt): 4 warps per CTA
                                  // 4.36e7 env steps/s, 8: 4.77e7, 12: 4.53e7, 24: 4.21e7; no phase sync at all: 3.41e7
#endif
#ifndef LOBSIM_ENVFAST_MINB
#define LOBSIM_ENVFAST_MINB (24 / LOBSIM_ENVFAST_WARPS)   // resident CTAs per SM the register allocation aims at (80 registers: measured
#endif                                                     // +9 % over 128 registers / 16 warps per SM, profiles/r02_env_ab.txt)
// ... but never more CTAs than the shared memory of an SM can hold for this layout (deep 50-level books: 2 CTAs, no register cap)
template <class LT>
constexpr int env_min_blocks() {
  const int warp_bytes = (LT::blob_bytes + 2 * MSG_TILE_BYTES + scratch_bytes(LT::NA) + 32 + 128 + 127) & ~127;
  const int fit = (227 * 1024) / (LOBSIM_ENVFAST_WARPS * warp_bytes);
  return fit < 1 ? 1 : (fit < LOBSIM_ENVFAST_MINB ? fit : LOBSIM_ENVFAST_MINB);
}
#ifndef LOBSIM_PHASE_SYNC
#define LOBSIM_PHASE_SYNC 1
#endif
// MODE of k_env_fast:
//   ENV_CLASSIC  -- the sorted level arrays of book_fast.cuh with every rare path in the kernel; flat blobs are converted on load.
//   ENV_HOT      -- flat order pools ONLY (book_flat.cuh): no sorted order path, no any-depth routines, no flat_leave, no tracked
//                   resync in the kernel.  A book that does not fit the pools, a pool that fills up, or an update_outer_levels with a
//                   level to overwrite makes the warp ABORT the env step: nothing of it has reached HBM (the blob is stored only
//                   at step ends, the feature state is written in phase C), so the env is handed -- {selection index, step} in
//                   p.defer_list -- to the ENV_DEFERRED launch that follows, which redoes that step and the rest of the launch.
//   ENV_DEFERRED -- ENV_CLASSIC over the items of p.defer_list (usually none: the launch exits at once).
// Why: the env kernel is bound by instruction fetch, not issue slots (profiles/r02_env_ab.txt): the flat path next to the sorted path
// in one kernel is 16 % slower than the sorted path alone, the flat path ALONE is 6 % faster (and 11 % fewer instructions).
#define ENV_CLASSIC 0
#define ENV_HOT 1
#define ENV_DEFERRED 2
#ifndef LOBSIM_SYNC_B
#define LOBSIM_SYNC_B 1           // barrier between phase A (ladders) and phase B (orders)
#endif
#ifndef LOBSIM_SYNC_C
#define LOBSIM_SYNC_C 1           // barrier between phase B (orders) and phase C (features, reward)
#endif
#if LOBSIM_PHASE_SYNC
#define PHASE_SYNC() __syncthreads()
#else
#define PHASE_SYNC() ((void)0)
#endif
// SYNC: the launch consists of full CTAs only, whose warps move through the phases of a step together
// (launch_env puts the n_sel % warps-per-CTA tail into a second, free-running launch).
// Per-warp values that are only needed outside the order-processing phase (B) live in the spare shared memory behind the
// mbarriers instead of in registers: the kernel is register-bound (128 registers at 16 resident warps per SM) and phase B is where it spills.
struct StepSave { double cash0, p0, price; long long inv0, episode_start_us, st_t0_us; };
// RARE: the configuration uses z-score normalisation or a RollingSharpe reward (their code is compiled out otherwise).
template <class LT, bool SYNC, bool RARE, int MODE>
__global__ void __launch_bounds__(32 * LOBSIM_ENVFAST_WARPS, env_min_blocks<LT>()) k_env_fast(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The warps of a CTA move through the phases of a step together (PHASE_SYNC = __syncthreads): the instruction
  // working set at any moment is a single phase, which is what keeps the 32 KB L1.5 I-cache warm
  // (profiles/r01_envstep_*: "no instruction" was the top stall with free-running warps).
  int sel = p.sel_offset + blockIdx.x * (blockDim.x >> 5) + warp, t_start = 0;
  if (MODE == ENV_DEFERRED) {
    static_assert(MODE != ENV_DEFERRED || !SYNC, "the deferred launch has no block-level barriers");
    if (sel >= *p.defer_count) return;
    const int2 item = p.defer_list[sel];
    sel = item.x; t_start = item.y;
  }
  if (sel >= p.n_sel) return; // only in SYNC == false launches
  const int env = p.env_ids ? p.env_ids[sel] : sel;
  const lobsim_cfg_t& c = ec.cfg;
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::blob_bytes;
  int2* scratch = reinterpret_cast<int2*>(msgbuf + 2 * MSG_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + scratch_bytes(LT::NA));
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  // ---- book blob, HBM -> shared memory: the 128-byte header first, then only the occupied part of the arrays (a reset rebuilds
  //      the book from the snapshot and needs the header alone)
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)sizeof(BookHdr));
    tma_load(base, gblob, (uint32_t)sizeof(BookHdr), &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  Book b; b.blob = base; b.L = p.L; b.lane = lane;
  BookHdr* h = reinterpret_cast<BookHdr*>(base);
  // flat: the book is in the flat order pools of book_flat.cuh (every book that fits them, when the handle allows it) -- in shared
  // memory during the launch AND in HBM between launches (one launch = one env step when a policy runs in between, so a
  // conversion per launch would cost more than the flat order path saves)
  bool flat = false;
  bool aborted = false;                                  // ENV_HOT: this env step goes to the deferred launch
  FlatState fs; fs.n0 = fs.n1 = 0; fs.seq = 0;
  auto defer = [&](int t) { aborted = true; if (lane == 0) { const int k = atomicAdd(p.defer_count, 1); p.defer_list[k] = make_int2(sel, t); } };
  if (!p.reset_mode) {
    if (hdr_is_flat(h)) {
      if (flat_body_copy<LT, true>(base, gblob, &bars[2], lane, h->cnt[0][1], h->cnt[1][1])) mbar_wait(&bars[2], 1);
      if (MODE == ENV_HOT) flat = true;
      else flat_leave_fn<LT>(base, lane, h->cnt[0][1], h->cnt[1][1]);     // the sorted kernels convert a flat blob on load
    } else if (blob_body_copy<LT, true>(base, gblob, &bars[2], lane)) mbar_wait(&bars[2], 1);
  }
  FastState f; f.err = h->err; f.dead = h->dead; f.bail = 0; f.bail_vol = 0;
  f.fill_log = p.fill_log ? p.fill_log + (size_t)env * p.fill_cap : nullptr; f.fill_cap = p.fill_cap;
  if (lane == 0) h->n_fills = 0;
  const int F = c.n_features;

  // ---- reset prologue ----------------------------------------------------------------------------------------------
  int stream_id = h->stream_id;
  if (p.reset_mode) {
    stream_id = p.reset_stream_ids[sel];
    int start = p.reset_steps[sel] - (p.reset_mode == 2 ? c.warmup_steps : 0);
    if (stream_id < 0 || stream_id >= p.n_streams) { stream_id = 0; start = -1; }
    const uint32_t ed = reset_book_cold(base, &p.L, lane, &ec.cfg, &p.streams[stream_id], stream_id, start);
    f.err = ed & 0x7fffffffu; f.dead = (int)(ed >> 31);
    if (p.reset_mode == 2 && lane == 0) {
      h->episode_start_step = p.reset_steps[sel];
      if (!c.portfolio_carryover || !h->has_reset) { h->inventory = c.initial_inventory; h->cash = c.initial_cash; }
      h->has_reset = 1;
    }
    __syncwarp();
  }
  if (MODE == ENV_HOT && flat) flat_adopt<LT>(base, lane, f, fs);
  else {
    fast_refresh_best(fb, f);
    if (MODE == ENV_HOT) {
      if (flat_fits(fb, 8)) { flat_enter(fb, fs); flat = true; }
      else defer(0);                                       // too many resting orders for the pools: the sorted kernel takes this env
    }
  }
  const lobsim_stream_t* stp = &p.streams[stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  StepSave* sv = reinterpret_cast<StepSave*>(bars + 4);
  int now_step = h->now_step;
  if (lane == 0) { sv->st_t0_us = stp->t0_us; sv->episode_start_us = stp->t0_us + (long long)h->episode_start_step * c.step_us; sv->price = h->price; }
  __syncwarp();
  FeatState* fstate_env = p.fstate + (size_t)env * LOBSIM_MAX_FEATURES;
  NormState* nstate_env = RARE && p.nstate ? p.nstate + (size_t)env * LOBSIM_MAX_FEATURES : nullptr;
  double* rings_env = p.rings + (size_t)env * ec.ring_stride;
  double feat_cur = 0.0;
  if (lane < F) feat_cur = fstate_env[lane].cur;
  const int t0_mod_min = (int)stp->reserved;   // t0_us % 60 s, filled in by lobsim_load_stream

  auto tops = [&](StepView& v) { // Orderbook.best_* / microprice, models.py:72-101
    v.have_tops = f.best0 != INT32_MIN && f.best1 != INT32_MAX;
    if (v.have_tops) {
      v.bb = f.best0; v.bs = f.best1;
      __syncwarp();
      if (MODE == ENV_HOT) { v.bv = flat_best_volume<LT, 0>(base, lane, f, fs); v.sv = flat_best_volume<LT, 1>(base, lane, f, fs); }
      else { v.bv = best_level_volume(b, 0, h->cnt[0][0]); v.sv = best_level_volume(b, 1, h->cnt[1][0]); }
      double imb; v.price = microprice(v.bb, v.bs, v.bv, v.sv, imb);
    } else { v.bb = v.bs = v.bv = v.sv = 0; v.price = NAN; }
  };
  if (p.reset_mode == 2) { // State(...) + _reset_features, HOE.py:152-154,218-221
    StepView v; tops(v);
    v.inventory = h->inventory; v.now_us = sv->st_t0_us + (long long)now_step * c.step_us;
    v.us_in_min = us_in_minute(t0_mod_min, now_step, ec);
    v.n_ext0 = v.n_ext1 = v.vol_ext0 = v.vol_ext1 = v.n_int0 = v.n_int1 = v.vol_int0 = v.vol_int1 = 0;
    __syncwarp();
    if (lane == 0) sv->price = v.price;
    feat_cur = features_step<RARE>(&ec, fstate_env, nstate_env, rings_env, lane, v, sv->episode_start_us, 1);
    __syncwarp();
  }

  // ---- message pipeline ----------------------------------------------------------------------------------------------
  const int T = p.T;
  const int n_grid = (int)stp->n_grid_steps;   // steps beyond the end: the existing ones are run, the first one past the end sets END_OF_STREAM
  if ((now_step < 0 || now_step > n_grid) && !f.dead && T > t_start) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; }
  unsigned g = 0, g_end_all = 0;
  if (!f.dead && !aborted && T > t_start) { g = __ldg(&st_step_off[now_step]); g_end_all = __ldg(&st_step_off[min((long long)now_step + (T - t_start), (long long)n_grid)]); }
  const unsigned tile0 = g / MSG_TILE;
  unsigned next_issue = 0, next_wait = 0;
  auto issue_tile = [&]() {
    const unsigned first = (tile0 + next_issue) * MSG_TILE;
    if (first >= g_end_all) return;
    if (lane == 0) {
      const unsigned n_total = (unsigned)stp->n_msgs;
      const unsigned cnt = n_total - first < MSG_TILE ? n_total - first : MSG_TILE;
      uint64_t* bar = &bars[next_issue & 1];
      mbar_expect_tx(bar, cnt * 16);
      tma_load(msgbuf + (next_issue & 1) * MSG_TILE_BYTES, st_msgs + first, cnt * 16, bar);
    }
    next_issue++;
  };
  auto wait_tile = [&]() { mbar_wait(&bars[next_wait & 1], (next_wait >> 1) & 1); next_wait++; };
  if (g < g_end_all) { issue_tile(); issue_tile(); }
  __syncwarp();
  if (SYNC) PHASE_SYNC();
  const int steps_per_sec = ec.steps_per_sec;
  int sub = now_step >= 0 ? now_step % steps_per_sec : 0;

  // shared memory -> HBM: the header and the occupied part of the arrays (sorted form) or of the pools (flat form); every issuing lane
  // commits and waits for its own copy.  ENV_HOT stores after EVERY step of a multi-step launch, so that an aborted step finds HBM
  // consistent (book, feature state, outputs) at the end of the step before.
  auto store_blob = [&]() {
    __syncwarp();
    if (lane == 0) {
      if (f.fill_log && h->n_fills > f.fill_cap) f.err |= LOBSIM_ERR_FILL_LOG_FULL;
      h->now_step = now_step; h->price = sv->price; h->err = f.err; h->dead = f.dead;
      if (p.fill_count) p.fill_count[env] = h->n_fills;
      if (MODE == ENV_HOT) flat_mark(h, fs);
    }
    __syncwarp();
    fence_proxy_async();
    __syncwarp();
    if (lane == 12) tma_store_part(gblob, base, (uint32_t)sizeof(BookHdr));
    if (MODE == ENV_HOT) flat_body_copy<LT, false>(base, gblob, nullptr, lane, fs.n0, fs.n1);
    else blob_body_copy<LT, false>(base, gblob, nullptr, lane);
    if (lane <= 12) { tma_store_commit(); tma_store_wait(); }
    __syncwarp();
  };
  AgentGenFast gen; gen.side = 3; gen.stage = 0; gen.need = 0; gen.todo0 = gen.todo1 = gen.wm0 = gen.wm1 = 0; gen.wide_removed = gen.wide_pos = 0;
  gen.pending_id = 0; gen.base0 = gen.base1 = 0; gen.clear_vol = -1; gen.clear_side = 0;
  const int Q = c.max_quote_level - c.min_quote_level;
  int* diff_scratch = reinterpret_cast<int*>(scratch);
  bool agent_phase = false;
#pragma unroll 1
  for (int t = t_start; t < T; t++) {
    if (SYNC) PHASE_SYNC(); // ---- phase A: action -> ladders (fp64) --------------------------------------------------------
    __syncwarp();
    if (lane == 0) { sv->cash0 = h->cash; sv->p0 = sv->price; sv->inv0 = h->inventory; } // deepcopy(self.state), HOE.py:166
    if (lane < 8) h->flow[lane] = 0;
    __syncwarp();
    if (p.agent_kind != LOBSIM_AGENT_NONE && !(MODE == ENV_HOT && aborted)) {
      double* act_sm = reinterpret_cast<double*>(scratch);
      if (p.agent_kind == LOBSIM_AGENT_EXTERNAL) {
        const double* a = p.actions_in + ((size_t)t * p.n_sel + sel) * ec.action_dim;
        if (lane < 5) act_sm[lane] = lane < ec.action_dim ? __ldg(&a[lane]) : 0.0;
      } else {
        const lobsim_agent_t* agp = p.agents ? p.agents + sel : &p.agent;
        const double inv_obs = __shfl_sync(FULL_MASK, feat_cur, agp->inventory_index & 31);
        if (lane == 0) agent_action_cold(agp, inv_obs, act_sm, env, now_step);
      }
      __syncwarp();
      const double a0 = act_sm[0], a1 = act_sm[1], a2 = act_sm[2], a3 = act_sm[3], a4 = act_sm[4];
      const double mine = lane < 5 ? act_sm[lane] : 0.0;
      __syncwarp();
      if (p.act && p.agent_kind != LOBSIM_AGENT_EXTERNAL && lane < ec.action_dim) p.act[((size_t)t * p.n_sel + sel) * ec.action_dim + lane] = mine;
      if (p.obs && !p.out_final_obs_only && c.inc_prev_action_in_obs && lane < ec.action_dim)
        p.obs[((size_t)t * p.n_sel + sel) * ec.obs_dim + F + lane] = mine;
      if (!f.dead) {
        const AgentGen g0 = agent_prepare(b, f.best0, f.best1, h->nag[0], h->nag[1], h->inventory, &ec, a0, a1, a2, a3, a4, p.beta_tab, diff_scratch);
        f.err |= g0.err_out; if (g0.dead_out) f.dead = 1;
        gen = agent_gen_fast_init(g0, lane, diff_scratch);
        agent_phase = true;
      }
    }
    if (SYNC && LOBSIM_SYNC_B) PHASE_SYNC(); // ---- phase B: the step's orders: the agent's first, then the historical messages of (now, now + step]
    {
      if (!f.dead && now_step >= (int)stp->n_grid_steps) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; } // re-read: keeps a register free
      const unsigned g_step_end = f.dead ? g : __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
      for (;;) {
        int type, side, oprice, vol; uint32_t ref; bool is_agent;
        if (MODE == ENV_HOT && aborted) break;
        if (agent_phase) {
          if (!agent_next_fast(fb, f, gen, diff_scratch, Q, c.tick_size, type, side, oprice, vol, ref)) { agent_phase = false; continue; }
          is_agent = true;
        } else {
          if (f.dead || g >= g_step_end) break;
          const unsigned tile = g / MSG_TILE - tile0;
          if (tile == next_wait) wait_tile();
          const uint4 m = *reinterpret_cast<const uint4*>(msgbuf + (tile & 1) * MSG_TILE_BYTES + (g % MSG_TILE) * 16);
          oprice = (int)m.x; vol = (int)m.y; ref = m.z; type = (int)(m.w & 7u); side = (int)((m.w >> 3) & 1u);
          is_agent = false;
          g++;
          if (g % MSG_TILE == 0) { __syncwarp(); issue_tile(); }
        }
        if (!f.dead) {
          if (MODE == ENV_HOT) { if (!flat_order_tracked<LT>(base, lane, f, fs, type, side, oprice, vol, ref, is_agent)) defer(t); }   // pool full
          else fast_order_full<LT, true>(fb, f, type, side, oprice, vol, ref, is_agent);
        }
      }
    }
    if (!f.dead && !(MODE == ENV_HOT && aborted)) {
      now_step++;
      if (++sub == steps_per_sec) {                          // whole second: outer-level resync, OrderbookSimulator.py:86-87
        sub = 0;
        if (c.resync && (!p.resync_last_only || t == T - 1)) {
          const double prop = ec.outer_prop;
          const double bbd = f.best0 == INT32_MIN ? 0.0 : (double)f.best0;
          const double bsd = f.best1 == INT32_MAX ? (double)INFINITY : (double)f.best1;
          if (bbd < (double)h->min_buy + prop * (double)h->init_buy_range || bsd > (double)h->max_sell - prop * (double)h->init_sell_range) {
            const int sec = now_step / steps_per_sec;
            if (sec <= (int)stp->n_seconds && stp->snap_valid[sec]) {
              const int32_t* row = stp->snapshots + (size_t)sec * 2 * c.n_levels * 2;
              if (MODE == ENV_HOT) {
                if (flat_resync_needed(h, row, c.n_levels, lane)) defer(t);   // a level to overwrite: the sorted kernel redoes this step
                else flat_update_trackers<LT>(base, lane, fs);
              } else fast_resync_tracked(fb, f, row, c.n_levels, scratch);
            }
          }
        }
      }
    }
    if (SYNC && LOBSIM_SYNC_C) PHASE_SYNC(); // ---- phase C: update_internal_state + _update_features + reward, HOE.py:163-178,199-204 -------------
    if (MODE == ENV_HOT && aborted) continue;              // (keeps taking part in the phase barriers of its CTA)
    StepView v; tops(v);
    if (!v.have_tops) f.err |= LOBSIM_ERR_EMPTY_BOOK;
    const double price = v.price;
    if (lane == 0) sv->price = price;
    v.inventory = h->inventory; v.now_us = sv->st_t0_us + (long long)now_step * c.step_us;
    v.us_in_min = us_in_minute(t0_mod_min, now_step, ec);
    v.n_ext0 = h->flow[0]; v.n_ext1 = h->flow[1]; v.vol_ext0 = h->flow[2]; v.vol_ext1 = h->flow[3];
    v.n_int0 = h->flow[4]; v.n_int1 = h->flow[5]; v.vol_int0 = h->flow[6]; v.vol_int1 = h->flow[7];
    __syncwarp(); // every lane has read this step's flow counters before lanes 0-7 zero them for the next step
    feat_cur = features_step<RARE>(&ec, fstate_env, nstate_env, rings_env, lane, v, sv->episode_start_us, 0);
    const bool write_now = !p.out_final_obs_only || t == T - 1;
    if (p.obs && write_now) {
      double* o = p.obs + ((size_t)(p.out_final_obs_only ? 0 : t) * p.n_sel + sel) * ec.obs_dim;
      if (lane < F) o[lane] = feat_cur;
      if (c.inc_prev_action_in_obs && lane < ec.action_dim && (p.agent_kind == LOBSIM_AGENT_NONE || p.out_final_obs_only)) o[F + lane] = 0.0;
    }
    if (p.agent_kind != LOBSIM_AGENT_NONE) {
      const double cash1 = h->cash; const long long inv1 = h->inventory;
      const bool d = now_step >= h->episode_start_step + c.episode_steps; // terminal_time - now < step/2, HOE.py:172
      const double r = step_reward<RARE>(p, c, env, lane, d, sv->cash0, sv->inv0, sv->p0, cash1, inv1, price, f.err);
      if (lane == 0) {
        if (p.rew) p.rew[(size_t)t * p.n_sel + sel] = r;
        if (p.done) p.done[(size_t)t * p.n_sel + sel] = d ? 1 : 0;
      }
      if (p.info) write_info(p.info + ((size_t)t * p.n_sel + sel) * LOBSIM_INFO_DIM, lane, v, cash1, inv1, f.err);
    }
    if (MODE == ENV_HOT && t + 1 < T) store_blob();       // checkpoint (see store_blob)
  }
  if (T == 0 && p.obs && p.reset_mode == 2) { // reset with no warm-up: obs straight after _reset_features
    double* o = p.obs + (size_t)sel * ec.obs_dim;
    if (lane < F) o[lane] = feat_cur;
    if (c.inc_prev_action_in_obs && lane < ec.action_dim) o[F + lane] = 0.0;
  }
  while (next_wait < next_issue) wait_tile();
  if (!(MODE == ENV_HOT && aborted)) store_blob();        // (an aborted env step leaves HBM as it was at the end of the step before)
}
