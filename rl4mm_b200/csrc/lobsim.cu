// lobsim.cu -- kernels and the C ABI (include/lobsim.h) of the B200-native LOB simulation step.
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
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <utility>
#include <new>
#include <string>
#include <vector>

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

struct AdvParams {
  unsigned char* blobs;             // [n_envs][blob_bytes]
  FeatState* fstate;                // [n_envs][LOBSIM_MAX_FEATURES]
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

  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  Book b; b.blob = base; b.L = p.L; b.lane = lane;
  unsigned char* msgbuf = base + p.L.blob_bytes;                                   // 2 x 512 B
  int2* scratch = reinterpret_cast<int2*>(msgbuf + 2 * MSG_TILE_BYTES);             // [2*NA]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + 2 * p.L.NA * 8); // 3 barriers
  unsigned char* gblob = p.blobs + (size_t)env * p.L.blob_bytes;

  // ---- book blob: HBM -> shared memory ---------------------------------------------------------------------------
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)p.L.blob_bytes);
    tma_load(base, gblob, (uint32_t)p.L.blob_bytes, &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);

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
  double* rings_env = p.rings + (size_t)env * ec.ring_stride;
  double feat_cur = 0.0; // Feature.current_value of this lane's feature
  if (kEnv && lane < F) feat_cur = fstate_env[lane].cur;

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
    v.n_ext0 = v.n_ext1 = v.vol_ext0 = v.vol_ext1 = v.n_int0 = v.n_int1 = v.vol_int0 = v.vol_int1 = 0;
    price = v.price;
    feat_cur = features_step<true>(&ec, fstate_env, rings_env, lane, v, episode_start_us, 1);
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
          if (lane == 0) agent_action_cold(agp, inv_obs, act_sm);
        }
        __syncwarp();
        const double a0 = act_sm[0], a1 = act_sm[1], a2 = act_sm[2], a3 = act_sm[3], a4 = act_sm[4];
        const double mine = lane < 5 ? act_sm[lane] : 0.0;
        __syncwarp();
        if (p.act && p.agent_kind != LOBSIM_AGENT_EXTERNAL && lane < ec.action_dim) p.act[((size_t)t * p.n_sel + sel) * ec.action_dim + lane] = mine;
        if (p.obs && !p.out_final_obs_only && c.inc_prev_action_in_obs && lane < ec.action_dim)
          p.obs[((size_t)t * p.n_sel + sel) * ec.obs_dim + F + lane] = mine; // get_observation(action), HOE.py:171
        gen = agent_prepare(b, w.nlv0, w.nlv1, w.nag0, w.nag1, w.inventory, &ec, a0, a1, a2, a3, a4);
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
      v.n_ext0 = w.n_ext0; v.n_ext1 = w.n_ext1; v.vol_ext0 = w.vol_ext0; v.vol_ext1 = w.vol_ext1;
      v.n_int0 = w.n_int0; v.n_int1 = w.n_int1; v.vol_int0 = w.vol_int0; v.vol_int1 = w.vol_int1;
      feat_cur = features_step<true>(&ec, fstate_env, rings_env, lane, v, episode_start_us, 0);
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
  if (lane == 0) { tma_store(gblob, base, (uint32_t)p.L.blob_bytes); tma_store_wait(); }
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
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::blob_bytes;
  int2* scratch = reinterpret_cast<int2*>(msgbuf + 2 * MSG_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + 2 * LT::NA * 8);
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)LT::blob_bytes);
    tma_load(base, gblob, (uint32_t)LT::blob_bytes, &bars[2]);
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
  if (lane == 0) { tma_store(gblob, base, (uint32_t)LT::blob_bytes); tma_store_wait(); }
  __syncwarp();
}

// ====================================================================================================================
//  the env fast kernel: k_advance<true,true> with the straight-line tracked order path (book_fast.cuh fast_order<LT,true>)
//  and all per-env scalars (counters, portfolio, per-step flow) in the shared-memory header instead of registers.
// ====================================================================================================================
__device__ __noinline__ uint32_t reset_book_cold(unsigned char* blob, const Layout* L, int lane, const lobsim_cfg_t* c, const lobsim_stream_t* st, int stream_id, int start_step) {
  Book b; b.blob = blob; b.L = *L; b.lane = lane;
  WarpState w;
  __syncwarp();
  load_state<true>(b, w);
  w.fill_log = nullptr; w.fill_cap = 0; w.n_fills = 0;
  init_book_from_snapshot(b, w, *c, *st, stream_id, start_step);
  store_state<true>(b, w);
  return pack_errdead(w.err, w.dead);
}

#ifndef LOBSIM_ENVFAST_WARPS
#define LOBSIM_ENVFAST_WARPS 4    // warps per CTA of the env fast kernel (four CTAs per SM at 128 registers; 8 measured 1.3 % slower)
#endif
#ifndef LOBSIM_PHASE_SYNC
#define LOBSIM_PHASE_SYNC 1
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
template <class LT, bool SYNC, bool RARE>
__global__ void __launch_bounds__(32 * LOBSIM_ENVFAST_WARPS, 16 / LOBSIM_ENVFAST_WARPS) k_env_fast(const __grid_constant__ AdvParams p, const __grid_constant__ EnvConst ec) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sel = p.sel_offset + blockIdx.x * (blockDim.x >> 5) + warp;
  // The warps of a CTA move through the phases of a step together (PHASE_SYNC = __syncthreads): the instruction
  // working set at any moment is a single phase, which is what keeps the 32 KB L1.5 I-cache warm
  // (profiles/r01_envstep_*: "no instruction" was the top stall with free-running warps).
  if (sel >= p.n_sel) return; // only in SYNC == false launches
  const int env = p.env_ids ? p.env_ids[sel] : sel;
  const lobsim_cfg_t& c = ec.cfg;
  unsigned char* base = warp_smem_base(smem, warp, p.warp_smem);
  unsigned char* msgbuf = base + LT::blob_bytes;
  int2* scratch = reinterpret_cast<int2*>(msgbuf + 2 * MSG_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(scratch) + 2 * LT::NA * 8);
  unsigned char* gblob = p.blobs + (size_t)env * LT::blob_bytes;
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[2], (uint32_t)LT::blob_bytes);
    tma_load(base, gblob, (uint32_t)LT::blob_bytes, &bars[2]);
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);

  FastBook<LT> fb; fb.blob = base; fb.lane = lane;
  Book b; b.blob = base; b.L = p.L; b.lane = lane;
  BookHdr* h = reinterpret_cast<BookHdr*>(base);
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
  fast_refresh_best(fb, f);
  const lobsim_stream_t* stp = &p.streams[stream_id];
  const lobsim_msg_t* __restrict__ st_msgs = stp->msgs;
  const uint32_t* __restrict__ st_step_off = stp->step_off;
  StepSave* sv = reinterpret_cast<StepSave*>(bars + 4);
  int now_step = h->now_step;
  if (lane == 0) { sv->st_t0_us = stp->t0_us; sv->episode_start_us = stp->t0_us + (long long)h->episode_start_step * c.step_us; sv->price = h->price; }
  __syncwarp();
  FeatState* fstate_env = p.fstate + (size_t)env * LOBSIM_MAX_FEATURES;
  double* rings_env = p.rings + (size_t)env * ec.ring_stride;
  double feat_cur = 0.0;
  if (lane < F) feat_cur = fstate_env[lane].cur;

  auto tops = [&](StepView& v) { // Orderbook.best_* / microprice, models.py:72-101
    const int n0 = h->cnt[0][0], n1 = h->cnt[1][0];
    v.have_tops = n0 > 0 && n1 > 0;
    if (v.have_tops) {
      v.bb = f.best0; v.bs = f.best1;
      v.bv = best_level_volume(b, 0, n0); v.sv = best_level_volume(b, 1, n1);
      double imb; v.price = microprice(v.bb, v.bs, v.bv, v.sv, imb);
    } else { v.bb = v.bs = v.bv = v.sv = 0; v.price = NAN; }
  };
  if (p.reset_mode == 2) { // State(...) + _reset_features, HOE.py:152-154,218-221
    StepView v; tops(v);
    v.inventory = h->inventory; v.now_us = sv->st_t0_us + (long long)now_step * c.step_us;
    v.n_ext0 = v.n_ext1 = v.vol_ext0 = v.vol_ext1 = v.n_int0 = v.n_int1 = v.vol_int0 = v.vol_int1 = 0;
    __syncwarp();
    if (lane == 0) sv->price = v.price;
    feat_cur = features_step<RARE>(&ec, fstate_env, rings_env, lane, v, sv->episode_start_us, 1);
    __syncwarp();
  }

  // ---- message pipeline ----------------------------------------------------------------------------------------------
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
  if (SYNC) PHASE_SYNC();
  const int steps_per_sec = ec.steps_per_sec;
  int sub = now_step >= 0 ? now_step % steps_per_sec : 0;

  AgentGen gen; gen.side = 3;
  bool agent_phase = false;
#pragma unroll 1
  for (int t = 0; t < T; t++) {
    if (SYNC) PHASE_SYNC(); // ---- phase A: action -> ladders (fp64) --------------------------------------------------------
    __syncwarp();
    if (lane == 0) { sv->cash0 = h->cash; sv->p0 = sv->price; sv->inv0 = h->inventory; } // deepcopy(self.state), HOE.py:166
    if (lane < 8) h->flow[lane] = 0;
    __syncwarp();
    if (p.agent_kind != LOBSIM_AGENT_NONE) {
      double* act_sm = reinterpret_cast<double*>(scratch);
      if (p.agent_kind == LOBSIM_AGENT_EXTERNAL) {
        const double* a = p.actions_in + ((size_t)t * p.n_sel + sel) * ec.action_dim;
        if (lane < 5) act_sm[lane] = lane < ec.action_dim ? __ldg(&a[lane]) : 0.0;
      } else {
        const lobsim_agent_t* agp = p.agents ? p.agents + sel : &p.agent;
        const double inv_obs = __shfl_sync(FULL_MASK, feat_cur, agp->inventory_index & 31);
        if (lane == 0) agent_action_cold(agp, inv_obs, act_sm);
      }
      __syncwarp();
      const double a0 = act_sm[0], a1 = act_sm[1], a2 = act_sm[2], a3 = act_sm[3], a4 = act_sm[4];
      const double mine = lane < 5 ? act_sm[lane] : 0.0;
      __syncwarp();
      if (p.act && p.agent_kind != LOBSIM_AGENT_EXTERNAL && lane < ec.action_dim) p.act[((size_t)t * p.n_sel + sel) * ec.action_dim + lane] = mine;
      if (p.obs && !p.out_final_obs_only && c.inc_prev_action_in_obs && lane < ec.action_dim)
        p.obs[((size_t)t * p.n_sel + sel) * ec.obs_dim + F + lane] = mine;
      if (!f.dead) {
        gen = agent_prepare(b, h->cnt[0][0], h->cnt[1][0], h->nag[0], h->nag[1], h->inventory, &ec, a0, a1, a2, a3, a4);
        f.err |= gen.err_out; if (gen.dead_out) f.dead = 1;
        agent_phase = true;
      }
    }
    if (SYNC) PHASE_SYNC(); // ---- phase B: the step's orders: the agent's first, then the historical messages of (now, now + step]
    {
      if (!f.dead && now_step >= (int)stp->n_grid_steps) { f.err |= LOBSIM_ERR_END_OF_STREAM; f.dead = 1; } // re-read: keeps a register free
      const unsigned g_step_end = f.dead ? g : __ldg(&st_step_off[now_step + 1]);
#pragma unroll 1
      for (;;) {
        int type, side, oprice, vol; uint32_t ref; bool is_agent;
        if (agent_phase) {
          if (!agent_next_fast(fb, f, gen, type, side, oprice, vol, ref)) { agent_phase = false; continue; }
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
        if (!f.dead) fast_order_full<LT, true>(fb, f, type, side, oprice, vol, ref, is_agent);
      }
    }
    if (!f.dead) {
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
              fast_resync_tracked(fb, f, row, c.n_levels, scratch);
            }
          }
        }
      }
    }
    if (SYNC) PHASE_SYNC(); // ---- phase C: update_internal_state + _update_features + reward, HOE.py:163-178,199-204 -------------
    StepView v; tops(v);
    if (!v.have_tops) f.err |= LOBSIM_ERR_EMPTY_BOOK;
    const double price = v.price;
    if (lane == 0) sv->price = price;
    v.inventory = h->inventory; v.now_us = sv->st_t0_us + (long long)now_step * c.step_us;
    v.n_ext0 = h->flow[0]; v.n_ext1 = h->flow[1]; v.vol_ext0 = h->flow[2]; v.vol_ext1 = h->flow[3];
    v.n_int0 = h->flow[4]; v.n_int1 = h->flow[5]; v.vol_int0 = h->flow[6]; v.vol_int1 = h->flow[7];
    __syncwarp(); // every lane has read this step's flow counters before lanes 0-7 zero them for the next step
    feat_cur = features_step<RARE>(&ec, fstate_env, rings_env, lane, v, sv->episode_start_us, 0);
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
  }
  if (T == 0 && p.obs && p.reset_mode == 2) { // reset with no warm-up: obs straight after _reset_features
    double* o = p.obs + (size_t)sel * ec.obs_dim;
    if (lane < F) o[lane] = feat_cur;
    if (c.inc_prev_action_in_obs && lane < ec.action_dim) o[F + lane] = 0.0;
  }
  while (next_wait < next_issue) wait_tile();
  __syncwarp();
  if (lane == 0) {
    if (f.fill_log && h->n_fills > f.fill_cap) f.err |= LOBSIM_ERR_FILL_LOG_FULL;
    h->now_step = now_step; h->price = sv->price; h->err = f.err; h->dead = f.dead;
    if (p.fill_count) p.fill_count[env] = h->n_fills;
  }
  __syncwarp();
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) { tma_store(gblob, base, (uint32_t)LT::blob_bytes); tma_store_wait(); }
  __syncwarp();
}

typedef StaticLayout<64, 256, 32> FastLayoutA;   // BASELINE config 2 (10-level books)
typedef StaticLayout<128, 512, 64> FastLayoutB;  // the default capacities (50-level books)
typedef StaticLayout<64, 256, 64> FastLayoutC;   // 10-level books with a 64-order agent table

// ====================================================================================================================
//  Exchange.process_order for a list of orders (drop-in / test entry point; one warp, sequential)
// ====================================================================================================================
struct OrdParams {
  unsigned char* blobs; Layout L; int32_t n_envs;
  const lobsim_order_t* orders; int32_t n;
  lobsim_fill_t* fills; int32_t max_fills; int32_t* n_fills; uint32_t* refs_out;
};

__global__ void __launch_bounds__(32) k_process_orders(const __grid_constant__ OrdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x;
  Book b; b.blob = smem; b.L = p.L; b.lane = lane;
  WarpState w;
  int cur_env = -1, total_fills = 0;
  auto load_env = [&](int env) {
    const uint4* src = reinterpret_cast<const uint4*>(p.blobs + (size_t)env * p.L.blob_bytes);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = lane; i < p.L.blob_bytes / 16; i += 32) dst[i] = src[i];
    __syncwarp();
    load_state<true>(b, w);
    w.fill_log = p.fills ? p.fills + total_fills : nullptr;
    w.fill_cap = p.max_fills - total_fills; w.n_fills = 0;
  };
  auto store_env = [&](int env) {
    store_state<true>(b, w);
    uint4* dst = reinterpret_cast<uint4*>(p.blobs + (size_t)env * p.L.blob_bytes);
    const uint4* src = reinterpret_cast<const uint4*>(smem);
    for (int i = lane; i < p.L.blob_bytes / 16; i += 32) dst[i] = src[i];
    __syncwarp();
    total_fills += w.n_fills < w.fill_cap ? w.n_fills : w.fill_cap;
  };
  for (int k = 0; k < p.n; k++) {
    const lobsim_order_t o = p.orders[k];
    if (o.env < 0 || o.env >= p.n_envs) continue;
    if (o.env != cur_env) { if (cur_env >= 0) store_env(cur_env); load_env(o.env); cur_env = o.env; }
    uint32_t id = 0;
    const bool is_agent = !o.is_external;
    const bool has_vol = !((o.type == LOBSIM_MSG_DELETE || o.type == LOBSIM_MSG_CANCEL) && o.volume <= 0);
    if (!w.dead) {
      if (has_vol && o.volume <= 0) w.err |= LOBSIM_ERR_BAD_VOLUME;
      else if (o.type == LOBSIM_MSG_LIMIT || o.type == LOBSIM_MSG_MARKET)
        id = submit_or_execute<true>(b, w, o.direction, o.price, o.volume, is_agent ? 0u : o.ref, o.type == LOBSIM_MSG_LIMIT, is_agent);
      else remove_order<true>(b, w, o.direction, o.price, o.volume, has_vol, is_agent ? (LOBSIM_REF_AGENT | (o.ref & 0x7fffffffu)) : o.ref, is_agent);
    }
    if (p.refs_out && lane == 0) p.refs_out[k] = id;
  }
  if (cur_env >= 0) store_env(cur_env);
  if (lane == 0 && p.n_fills) *p.n_fills = total_fills;
}

// ====================================================================================================================
//  per-env state summary (one thread per env, reads the blob in HBM)
// ====================================================================================================================
__global__ void k_get_state(const unsigned char* blobs, Layout L, int first, int n, lobsim_env_state_t* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* blob = blobs + (size_t)(first + i) * L.blob_bytes;
  const BookHdr* h = reinterpret_cast<const BookHdr*>(blob);
  lobsim_env_state_t s;
  s.inventory = h->inventory; s.cash = h->cash; s.price = h->price; s.now_step = h->now_step;
  s.episode_start_step = h->episode_start_step; s.min_buy_price = h->min_buy; s.max_sell_price = h->max_sell;
  s.err = h->err; s.stream_id = h->stream_id; s.n_agent_orders[0] = h->nag[0]; s.n_agent_orders[1] = h->nag[1];
  s.next_agent_id = h->next_agent_id; s.reserved = 0;
  int best[2] = {0, INT32_MAX}, bvol[2] = {0, 0};
  for (int side = 0; side < 2; side++) {
    int nlv = h->cnt[side][0];
    if (!nlv) continue;
    const unsigned char* sb = blob + L.side_off + side * L.side_stride;
    const int32_t* lvp = reinterpret_cast<const int32_t*>(sb);
    const uint16_t* lvend = reinterpret_cast<const uint16_t*>(sb + L.lvend_off);
    const uint2* ord = reinterpret_cast<const uint2*>(sb + L.ord_off);
    best[side] = lvp[nlv - 1];
    int start = nlv > 1 ? lvend[nlv - 2] : 0, end = lvend[nlv - 1], v = 0;
    for (int k = start; k < end; k++) v += (int)ord[k].x;
    bvol[side] = v;
  }
  s.best_buy = best[0]; s.best_sell = best[1]; s.best_buy_volume = bvol[0]; s.best_sell_volume = bvol[1];
  out[i] = s;
}

__global__ void k_init_blobs(unsigned char* blobs, Layout L, int n, long long inventory, double cash) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BookHdr* h = reinterpret_cast<BookHdr*>(blobs + (size_t)i * L.blob_bytes);
  memset(h, 0, sizeof(BookHdr));
  h->next_agent_id = 1; h->inventory = inventory; h->cash = cash; h->stream_id = 0;
}

// ====================================================================================================================
//  host side: handle + C ABI
// ====================================================================================================================
static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CUDA_TRY(expr)                                                                                             \
  do {                                                                                                             \
    cudaError_t e__ = (expr);                                                                                      \
    if (e__ != cudaSuccess) return fail(LOBSIM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));       \
  } while (0)

struct lobsim {
  lobsim_cfg_t cfg;
  EnvConst ec;
  Layout L;
  int device;
  int warps_per_cta;
  int env_warps_per_cta;
  int warp_smem;
  unsigned char* blobs = nullptr;
  FeatState* fstate = nullptr;
  double* rings = nullptr;
  double* rs_ring = nullptr;
  int32_t* rs_state = nullptr;
  lobsim_fill_t* fill_log = nullptr;
  int32_t* fill_count = nullptr;
  lobsim_agent_t* agents_dev = nullptr; // per-env agents of lobsim_rollout_agents (allocated on first use)
  bool rare_paths = false;              // z-score normalisation or a RollingSharpe reward configured: kernels with that code
  std::vector<lobsim_stream_t> streams;
  lobsim_stream_t* streams_dev = nullptr;
  int streams_cap = 0;
  // staging for the *_host entry points
  double* st_actions = nullptr; double* st_obs = nullptr; double* st_rew = nullptr; uint8_t* st_done = nullptr;
  lobsim_env_state_t* st_state = nullptr;
  lobsim_msg_t* st_msgs = nullptr; uint64_t st_msgs_cap = 0;
  bool has_reset = false;
  bool force_general = false;         // LOBSIM_FORCE_GENERAL=1: always use the runtime-layout kernels (testing)
  bool agent_orders_possible = false; // an agent order may rest in some book (disables the replay fast path)
  int64_t launches = 0;
};

extern "C" {

const char* lobsim_last_error(void) { return g_last_error.c_str(); }
int lobsim_abi_version(void) { return LOBSIM_ABI_VERSION; }
int lobsim_action_dim(const lobsim_cfg_t* c) { return (c->concentration >= 0 ? 2 : 4) + (c->market_order_clearing ? 1 : 0); }
int lobsim_obs_dim(const lobsim_cfg_t* c) { return c->n_features + (c->inc_prev_action_in_obs ? lobsim_action_dim(c) : 0); }

static int validate_cfg(const lobsim_cfg_t* c) {
  if (!c) return fail(LOBSIM_E_INVALID, "cfg is null");
  if (c->abi_version != LOBSIM_ABI_VERSION) return fail(LOBSIM_E_INVALID, "abi_version mismatch");
  if (c->n_envs <= 0 || c->n_levels <= 0 || c->tick_size <= 0 || c->step_us <= 0) return fail(LOBSIM_E_INVALID, "n_envs, n_levels, tick_size, step_us must be positive");
  if (1000000 % c->step_us) return fail(LOBSIM_E_INVALID, "step_us must divide one second");
  int Q = c->max_quote_level - c->min_quote_level;
  if (Q <= 0 || Q > 32) return fail(LOBSIM_E_INVALID, "1 <= max_quote_level - min_quote_level <= 32");
  if (c->n_features < 0 || c->n_features > LOBSIM_MAX_FEATURES) return fail(LOBSIM_E_INVALID, "too many features");
  if (c->max_levels_per_side < 4 || c->max_levels_per_side % 4 || c->max_levels_per_side > 4096) return fail(LOBSIM_E_INVALID, "max_levels_per_side must be a multiple of 4 in [4, 4096]");
  if (c->max_orders_per_side < c->max_levels_per_side || c->max_orders_per_side > 65535) return fail(LOBSIM_E_INVALID, "max_orders_per_side must be in [max_levels_per_side, 65535]");
  if (c->max_agent_orders < 1 || c->max_agent_orders > 1024) return fail(LOBSIM_E_INVALID, "max_agent_orders must be in [1, 1024]");
  if (c->n_levels > c->max_levels_per_side) return fail(LOBSIM_E_INVALID, "n_levels exceeds max_levels_per_side");
  if (c->warmup_steps < 0 || c->episode_steps <= 0) return fail(LOBSIM_E_INVALID, "bad episode_steps / warmup_steps");
  const lobsim_reward_t* rw[2] = {&c->step_reward, &c->terminal_reward};
  for (int i = 0; i < 2; i++) {
    if (rw[i]->kind < 0 || rw[i]->kind > LOBSIM_REWARD_ROLLING_SHARPE) return fail(LOBSIM_E_INVALID, "unknown reward kind");
    if (rw[i]->kind == LOBSIM_REWARD_ROLLING_SHARPE) {
      const int maxw = rw[i]->asymmetric & 0xffff, minw = (rw[i]->asymmetric >> 16) & 0xffff;
      if (minw < 2 || minw > maxw || maxw > LOBSIM_MAX_SHARPE_WINDOW) return fail(LOBSIM_E_INVALID, "ROLLING_SHARPE: need 2 <= min_window <= max_window <= 256");
    }
  }
  for (int i = 0; i < c->n_features; i++) {
    const lobsim_feature_t& f = c->features[i];
    if (f.kind < 0 || f.kind > LOBSIM_FEAT_AMIHUD_LAMBDA) return fail(LOBSIM_E_INVALID, "unknown feature kind");
    if (f.kind == LOBSIM_FEAT_AMIHUD_LAMBDA && (f.iparam < 1 || f.iparam > 16383 || f.lookback % f.iparam || f.lookback / f.iparam < 2 || f.lookback / f.iparam > 65535))
      return fail(LOBSIM_E_INVALID, "AMIHUD_LAMBDA: lookback must be (true_lookback + 1) * slowing_factor with true_lookback >= 1");
    if (f.update_us <= 0 || f.update_us > 60000000 || f.lookback < 0) return fail(LOBSIM_E_INVALID, "bad feature update_us / lookback");
    if (f.norm_len < 0 || f.norm_len > 10000000) return fail(LOBSIM_E_INVALID, "bad feature norm_len");
    if (f.kind == LOBSIM_FEAT_TIME_OF_DAY && f.iparam <= 0) return fail(LOBSIM_E_INVALID, "TIME_OF_DAY needs n_buckets > 0");
    if ((f.kind == LOBSIM_FEAT_VOLATILITY || f.kind == LOBSIM_FEAT_TRADE_DIR_IMBALANCE || f.kind == LOBSIM_FEAT_TRADE_VOL_IMBALANCE) && f.lookback < 1)
      return fail(LOBSIM_E_INVALID, "windowed feature needs lookback >= 1");
  }
  return LOBSIM_OK;
}

int64_t lobsim_state_bytes(const lobsim_cfg_t* c) {
  if (validate_cfg(c)) return LOBSIM_E_INVALID;
  return make_layout(c->max_levels_per_side, c->max_orders_per_side, c->max_agent_orders).blob_bytes;
}

static int warp_smem_bytes(const Layout& L) { return (L.blob_bytes + 2 * MSG_TILE_BYTES + 2 * L.NA * 8 + 32 + 128 + 127) & ~127; }

int lobsim_create(const lobsim_cfg_t* cfg, int device, lobsim_t** out) {
  if (!out) return fail(LOBSIM_E_INVALID, "out is null");
  int rc = validate_cfg(cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LOBSIM_E_CUDA, "no CUDA device: the lobsim hot path has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(LOBSIM_E_INVALID, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  lobsim* h = new (std::nothrow) lobsim();
  if (!h) return fail(LOBSIM_E_NOMEM, "out of host memory");
  h->cfg = *cfg; h->device = device;
  { const char* e = getenv("LOBSIM_FORCE_GENERAL"); h->force_general = e && e[0] == '1'; }
  h->rare_paths = cfg->step_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE || cfg->terminal_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE;
  for (int i = 0; i < cfg->n_features; i++) h->rare_paths = h->rare_paths || cfg->features[i].norm_len > 0;
  h->L = make_layout(cfg->max_levels_per_side, cfg->max_orders_per_side, cfg->max_agent_orders);
  h->warp_smem = warp_smem_bytes(h->L);
  int max_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (h->warp_smem > max_smem) { delete h; return fail(LOBSIM_E_INVALID, "book capacities exceed the shared memory of one SM"); }
  h->warps_per_cta = 4;
  while (h->warps_per_cta > 1 && h->warps_per_cta * h->warp_smem > max_smem) h->warps_per_cta >>= 1;
  h->env_warps_per_cta = LOBSIM_ENVFAST_WARPS;
  while (h->env_warps_per_cta > 1 && h->env_warps_per_cta * h->warp_smem > max_smem - 1024) h->env_warps_per_cta >>= 1;
  CUDA_TRY((cudaFuncSetAttribute(k_advance<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->warps_per_cta * h->warp_smem)));
  CUDA_TRY((cudaFuncSetAttribute(k_advance<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->warps_per_cta * h->warp_smem)));
  {
    const void* fast_kernels[] = {(const void*)k_replay_fast<FastLayoutA>, (const void*)k_replay_fast<FastLayoutB>, (const void*)k_replay_fast<FastLayoutC>,
                                  (const void*)k_env_fast<FastLayoutA, true, false>, (const void*)k_env_fast<FastLayoutB, true, false>, (const void*)k_env_fast<FastLayoutC, true, false>,
                                  (const void*)k_env_fast<FastLayoutA, false, false>, (const void*)k_env_fast<FastLayoutB, false, false>, (const void*)k_env_fast<FastLayoutC, false, false>,
                                  (const void*)k_env_fast<FastLayoutA, true, true>, (const void*)k_env_fast<FastLayoutB, true, true>, (const void*)k_env_fast<FastLayoutC, true, true>,
                                  (const void*)k_env_fast<FastLayoutA, false, true>, (const void*)k_env_fast<FastLayoutB, false, true>, (const void*)k_env_fast<FastLayoutC, false, true>};
    for (int i = 0; i < 15; i++) {
      const void* k = fast_kernels[i];
      const int dyn = (i < 3 ? h->warps_per_cta : h->env_warps_per_cta) * h->warp_smem;
      CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
      CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
  }
  CUDA_TRY((cudaFuncSetAttribute(k_advance<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)));
  CUDA_TRY((cudaFuncSetAttribute(k_advance<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)));
  CUDA_TRY(cudaFuncSetAttribute(k_process_orders, cudaFuncAttributeMaxDynamicSharedMemorySize, h->L.blob_bytes));
  // feature rings
  memset(&h->ec, 0, sizeof h->ec);
  h->ec.cfg = *cfg;
  int slots = 0;
  for (int i = 0; i < cfg->n_features; i++) {
    const lobsim_feature_t& ft = cfg->features[i];
    int need = ft.lookback + 2;
    if (ft.kind == LOBSIM_FEAT_AMIHUD_LAMBDA) { const int kk = ft.lookback / ft.iparam - 1; if (2 * kk + 2 > need) need = 2 * kk + 2; }
    h->ec.ring_off[i] = slots; slots += need;
    h->ec.hist_off[i] = slots; slots += ft.norm_len > 0 ? ft.norm_len : 0;
  }
  h->ec.ring_stride = (slots + 1) & ~1;
  h->ec.action_dim = lobsim_action_dim(cfg); h->ec.obs_dim = lobsim_obs_dim(cfg);
  h->ec.steps_per_sec = (int)(1000000 / cfg->step_us); h->ec.outer_prop = (double)cfg->outer_levels / (double)cfg->n_levels;
  const size_t n = (size_t)cfg->n_envs;
  CUDA_TRY(cudaMalloc(&h->blobs, n * h->L.blob_bytes));
  CUDA_TRY(cudaMalloc(&h->fstate, n * LOBSIM_MAX_FEATURES * sizeof(FeatState)));
  CUDA_TRY(cudaMemset(h->fstate, 0, n * LOBSIM_MAX_FEATURES * sizeof(FeatState)));
  {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const size_t ring_bytes = n * (size_t)(h->ec.ring_stride > 0 ? h->ec.ring_stride : 2) * sizeof(double);
    if (ring_bytes > free_b) { lobsim_destroy(h); return fail(LOBSIM_E_NOMEM, "feature windows / normalisation histories (n_envs x sum(lookback + max_norm_len) x 8 B) exceed the free HBM"); }
    CUDA_TRY(cudaMalloc(&h->rings, ring_bytes));
  }
  CUDA_TRY(cudaMemset(h->blobs, 0, n * h->L.blob_bytes));
  if (cfg->step_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE || cfg->terminal_reward.kind == LOBSIM_REWARD_ROLLING_SHARPE) {
    CUDA_TRY(cudaMalloc(&h->rs_ring, n * 3 * LOBSIM_MAX_SHARPE_WINDOW * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->rs_state, n * 4 * sizeof(int32_t)));
    CUDA_TRY(cudaMemset(h->rs_state, 0, n * 4 * sizeof(int32_t)));
  }
  if (cfg->fill_log_capacity > 0) {
    CUDA_TRY(cudaMalloc(&h->fill_log, n * (size_t)cfg->fill_log_capacity * sizeof(lobsim_fill_t)));
    CUDA_TRY(cudaMalloc(&h->fill_count, n * sizeof(int32_t)));
    CUDA_TRY(cudaMemset(h->fill_count, 0, n * sizeof(int32_t)));
  }
  k_init_blobs<<<(cfg->n_envs + 127) / 128, 128>>>(h->blobs, h->L, cfg->n_envs, cfg->initial_inventory, cfg->initial_cash);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  h->launches++;
  *out = h;
  return LOBSIM_OK;
}

int lobsim_destroy(lobsim_t* h) {
  if (!h) return LOBSIM_OK;
  cudaSetDevice(h->device);
  cudaFree(h->blobs); cudaFree(h->fstate); cudaFree(h->rings); cudaFree(h->rs_ring); cudaFree(h->rs_state); cudaFree(h->fill_log); cudaFree(h->fill_count); cudaFree(h->agents_dev);
  cudaFree(h->streams_dev); cudaFree(h->st_actions); cudaFree(h->st_obs); cudaFree(h->st_rew); cudaFree(h->st_done);
  cudaFree(h->st_state); cudaFree(h->st_msgs);
  delete h;
  return LOBSIM_OK;
}

int lobsim_load_stream(lobsim_t* h, int stream_id, const lobsim_stream_t* s) {
  if (!h || !s) return fail(LOBSIM_E_INVALID, "null argument");
  if (stream_id < 0 || stream_id > 65535) return fail(LOBSIM_E_INVALID, "bad stream id");
  if (!s->step_off || !s->snapshots || !s->snap_valid || (!s->msgs && s->n_msgs)) return fail(LOBSIM_E_INVALID, "stream arrays missing");
  if (s->t0_us % 1000000) return fail(LOBSIM_E_INVALID, "t0_us must be a whole second");
  if (s->n_msgs >= 0xffffffffull) return fail(LOBSIM_E_INVALID, "more than 2^32-1 messages in one stream");
  if (((uintptr_t)s->msgs & 15) != 0) return fail(LOBSIM_E_INVALID, "msgs must be 16-byte aligned");
  CUDA_TRY(cudaSetDevice(h->device));
  if ((int)h->streams.size() <= stream_id) {
    lobsim_stream_t empty; memset(&empty, 0, sizeof empty);
    h->streams.resize(stream_id + 1, empty);
  }
  h->streams[stream_id] = *s;
  if (h->streams_cap < (int)h->streams.size()) {
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(h->streams_dev);
    h->streams_cap = (int)h->streams.size() * 2;
    CUDA_TRY(cudaMalloc(&h->streams_dev, h->streams_cap * sizeof(lobsim_stream_t)));
  }
  CUDA_TRY(cudaMemcpy(h->streams_dev, h->streams.data(), h->streams.size() * sizeof(lobsim_stream_t), cudaMemcpyHostToDevice));
  return LOBSIM_OK;
}

static void base_params(lobsim* h, AdvParams& p) {
  memset(&p, 0, sizeof p);
  p.blobs = h->blobs; p.fstate = h->fstate; p.rings = h->rings; p.rs_ring = h->rs_ring; p.rs_state = h->rs_state; p.streams = h->streams_dev; p.n_streams = (int)h->streams.size();
  p.fill_log = h->fill_log; p.fill_count = h->fill_count; p.fill_cap = h->cfg.fill_log_capacity;
  p.n_envs = h->cfg.n_envs; p.n_sel = h->cfg.n_envs; p.L = h->L; p.warp_smem = h->warp_smem;
}

} // extern "C"

template <bool kEnv, bool kTrack>
static int launch_advance(lobsim* h, const AdvParams& p, cudaStream_t stream) {
  if (h->streams.empty()) return fail(LOBSIM_E_STATE, "no stream loaded");
  CUDA_TRY(cudaSetDevice(h->device));
  int wpc = h->warps_per_cta;
  int grid = (p.n_sel + wpc - 1) / wpc;
  if (grid <= 0) return LOBSIM_OK;
  k_advance<kEnv, kTrack><<<grid, wpc * 32, (size_t)wpc * h->warp_smem, stream>>>(p, h->ec);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return LOBSIM_OK;
}

static int launch_replay_fast(lobsim* h, const AdvParams& p, cudaStream_t stream) {
  if (h->streams.empty()) return fail(LOBSIM_E_STATE, "no stream loaded");
  const bool a = FastLayoutA::matches(h->L), b = FastLayoutB::matches(h->L), cc = FastLayoutC::matches(h->L);
  if (!a && !b && !cc) return 1;
  CUDA_TRY(cudaSetDevice(h->device));
  const int wpc = h->warps_per_cta, grid = (p.n_sel + wpc - 1) / wpc;
  if (grid <= 0) return LOBSIM_OK;
  const size_t dyn = (size_t)wpc * h->warp_smem;
  if (a) k_replay_fast<FastLayoutA><<<grid, wpc * 32, dyn, stream>>>(p, h->ec);
  else if (b) k_replay_fast<FastLayoutB><<<grid, wpc * 32, dyn, stream>>>(p, h->ec);
  else k_replay_fast<FastLayoutC><<<grid, wpc * 32, dyn, stream>>>(p, h->ec);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return LOBSIM_OK;
}

template <bool SYNC, bool RARE>
static void launch_env_fast_t(int layout, int grid, int block, size_t dyn, cudaStream_t stream, const AdvParams& p, const EnvConst& ec) {
  if (layout == 0) k_env_fast<FastLayoutA, SYNC, RARE><<<grid, block, dyn, stream>>>(p, ec);
  else if (layout == 1) k_env_fast<FastLayoutB, SYNC, RARE><<<grid, block, dyn, stream>>>(p, ec);
  else k_env_fast<FastLayoutC, SYNC, RARE><<<grid, block, dyn, stream>>>(p, ec);
}
static void launch_env_fast(int layout, bool sync, bool rare, int grid, int block, size_t dyn, cudaStream_t stream, const AdvParams& p, const EnvConst& ec) {
  if (sync) { if (rare) launch_env_fast_t<true, true>(layout, grid, block, dyn, stream, p, ec); else launch_env_fast_t<true, false>(layout, grid, block, dyn, stream, p, ec); }
  else { if (rare) launch_env_fast_t<false, true>(layout, grid, block, dyn, stream, p, ec); else launch_env_fast_t<false, false>(layout, grid, block, dyn, stream, p, ec); }
}

// env launches (reset / step / rollout): the straight-line kernel when a compiled StaticLayout matches the capacities,
// the general runtime-layout kernel otherwise
static int launch_env(lobsim* h, const AdvParams& p, cudaStream_t stream) {
  if (h->streams.empty()) return fail(LOBSIM_E_STATE, "no stream loaded");
  const bool a = FastLayoutA::matches(h->L), b = FastLayoutB::matches(h->L), cc = FastLayoutC::matches(h->L);
  if ((!a && !b && !cc) || h->force_general) return launch_advance<true, true>(h, p, stream);
  CUDA_TRY(cudaSetDevice(h->device));
  const int wpc = h->env_warps_per_cta;
  const size_t dyn = (size_t)wpc * h->warp_smem;
  const int full = p.n_sel / wpc, tail = p.n_sel % wpc;
  const int layout = a ? 0 : (b ? 1 : 2);
  if (full > 0) { // full CTAs: phase-synchronous
    launch_env_fast(layout, true, h->rare_paths, full, wpc * 32, dyn, stream, p, h->ec);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
  }
  if (tail > 0) { // the remaining n_sel % wpc envs: one partially filled CTA without block-level barriers
    AdvParams pt = p;
    pt.sel_offset = full * wpc;
    launch_env_fast(layout, false, h->rare_paths, 1, wpc * 32, dyn, stream, pt, h->ec);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
  }
  return LOBSIM_OK;
}

extern "C" {

int lobsim_reset_book(lobsim_t* h, const int32_t* env_ids, int32_t n, const int32_t* stream_ids, const int32_t* start_steps, void* stream) {
  if (!h || !stream_ids || !start_steps) return fail(LOBSIM_E_INVALID, "null argument");
  AdvParams p; base_params(h, p);
  p.env_ids = env_ids; p.n_sel = env_ids ? n : h->cfg.n_envs;
  p.T = 0; p.reset_mode = 1; p.reset_stream_ids = stream_ids; p.reset_steps = start_steps; p.agent_kind = LOBSIM_AGENT_NONE;
  if (!env_ids) h->agent_orders_possible = false; // every book is a fresh snapshot again
  return launch_advance<false, true>(h, p, (cudaStream_t)stream);
}

int lobsim_reset(lobsim_t* h, const int32_t* env_ids, int32_t n, const int32_t* stream_ids, const int32_t* episode_start_steps, double* obs_out, void* stream) {
  if (!h || !stream_ids || !episode_start_steps) return fail(LOBSIM_E_INVALID, "null argument");
  AdvParams p; base_params(h, p);
  p.env_ids = env_ids; p.n_sel = env_ids ? n : h->cfg.n_envs;
  p.T = h->cfg.warmup_steps; p.reset_mode = 2; p.reset_stream_ids = stream_ids; p.reset_steps = episode_start_steps;
  p.agent_kind = LOBSIM_AGENT_NONE; p.obs = obs_out; p.out_final_obs_only = 1;
  h->has_reset = true;
  return launch_env(h, p, (cudaStream_t)stream);
}

int lobsim_step(lobsim_t* h, const double* actions, double* obs_out, double* reward_out, uint8_t* done_out, void* stream) {
  if (!h || !actions) return fail(LOBSIM_E_INVALID, "null argument");
  if (!h->has_reset) return fail(LOBSIM_E_STATE, "step before reset");
  AdvParams p; base_params(h, p);
  p.T = 1; p.agent_kind = LOBSIM_AGENT_EXTERNAL; p.actions_in = actions; p.obs = obs_out; p.rew = reward_out; p.done = done_out;
  h->agent_orders_possible = true;
  return launch_env(h, p, (cudaStream_t)stream);
}

static int ensure_staging(lobsim* h) {
  if (h->st_actions) return LOBSIM_OK;
  const size_t n = (size_t)h->cfg.n_envs;
  CUDA_TRY(cudaMalloc(&h->st_actions, n * 8 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->st_obs, n * (LOBSIM_MAX_FEATURES + 8) * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->st_rew, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&h->st_done, n));
  CUDA_TRY(cudaMalloc(&h->st_state, n * sizeof(lobsim_env_state_t)));
  return LOBSIM_OK;
}

int lobsim_step_host(lobsim_t* h, const double* actions, double* obs_out, double* reward_out, uint8_t* done_out) {
  if (!h || !actions) return fail(LOBSIM_E_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_staging(h);
  if (rc) return rc;
  const size_t n = (size_t)h->cfg.n_envs;
  CUDA_TRY(cudaMemcpyAsync(h->st_actions, actions, n * h->ec.action_dim * sizeof(double), cudaMemcpyHostToDevice, 0));
  rc = lobsim_step(h, h->st_actions, h->st_obs, h->st_rew, h->st_done, nullptr);
  if (rc) return rc;
  if (obs_out) CUDA_TRY(cudaMemcpyAsync(obs_out, h->st_obs, n * h->ec.obs_dim * sizeof(double), cudaMemcpyDeviceToHost, 0));
  if (reward_out) CUDA_TRY(cudaMemcpyAsync(reward_out, h->st_rew, n * sizeof(double), cudaMemcpyDeviceToHost, 0));
  if (done_out) CUDA_TRY(cudaMemcpyAsync(done_out, h->st_done, n, cudaMemcpyDeviceToHost, 0));
  CUDA_TRY(cudaStreamSynchronize(0));
  return LOBSIM_OK;
}

int lobsim_rollout(lobsim_t* h, int32_t T, const lobsim_agent_t* agent, double* obs, double* act, double* rew, uint8_t* done, void* stream) {
  return lobsim_rollout_info(h, T, agent, obs, act, rew, done, nullptr, stream);
}

int lobsim_rollout_info(lobsim_t* h, int32_t T, const lobsim_agent_t* agent, double* obs, double* act, double* rew, uint8_t* done, double* info, void* stream) {
  if (!h || !agent || T < 0) return fail(LOBSIM_E_INVALID, "bad argument");
  if (!h->has_reset) return fail(LOBSIM_E_STATE, "rollout before reset");
  if (agent->kind == LOBSIM_AGENT_EXTERNAL && !act) return fail(LOBSIM_E_INVALID, "EXTERNAL agent needs the act tensor as input");
  if (agent->kind == LOBSIM_AGENT_TERADACTYL && (agent->inventory_index < 0 || agent->inventory_index >= h->cfg.n_features)) return fail(LOBSIM_E_INVALID, "bad inventory_index");
  AdvParams p; base_params(h, p);
  p.T = T; p.agent_kind = agent->kind; p.agent = *agent; p.obs = obs; p.rew = rew; p.done = done; p.info = info;
  if (agent->kind == LOBSIM_AGENT_EXTERNAL) p.actions_in = act; else p.act = act;
  if (agent->kind != LOBSIM_AGENT_NONE) h->agent_orders_possible = true;
  return launch_env(h, p, (cudaStream_t)stream);
}

int lobsim_rollout_agents(lobsim_t* h, int32_t T, const lobsim_agent_t* agents_host, double* obs, double* act, double* rew, uint8_t* done, double* info, void* stream) {
  if (!h || !agents_host || T < 0) return fail(LOBSIM_E_INVALID, "bad argument");
  if (!h->has_reset) return fail(LOBSIM_E_STATE, "rollout before reset");
  const int n = h->cfg.n_envs;
  for (int i = 0; i < n; i++) {
    const lobsim_agent_t& a = agents_host[i];
    if (a.kind != LOBSIM_AGENT_FIXED && a.kind != LOBSIM_AGENT_TERADACTYL) return fail(LOBSIM_E_INVALID, "per-env agents must be FIXED or TERADACTYL");
    if (a.kind == LOBSIM_AGENT_TERADACTYL && (a.inventory_index < 0 || a.inventory_index >= h->cfg.n_features)) return fail(LOBSIM_E_INVALID, "bad inventory_index");
  }
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->agents_dev) CUDA_TRY(cudaMalloc(&h->agents_dev, (size_t)n * sizeof(lobsim_agent_t)));
  CUDA_TRY(cudaMemcpyAsync(h->agents_dev, agents_host, (size_t)n * sizeof(lobsim_agent_t), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  AdvParams p; base_params(h, p);
  p.T = T; p.agent_kind = LOBSIM_AGENT_TERADACTYL; p.agent = agents_host[0]; p.agents = h->agents_dev;
  p.obs = obs; p.act = act; p.rew = rew; p.done = done; p.info = info;
  h->agent_orders_possible = true;
  return launch_env(h, p, (cudaStream_t)stream);
}

int lobsim_replay(lobsim_t* h, int32_t n_steps, void* stream) {
  if (!h || n_steps < 0) return fail(LOBSIM_E_INVALID, "bad argument");
  AdvParams p; base_params(h, p);
  p.T = n_steps; p.agent_kind = LOBSIM_AGENT_NONE;
  // fast path: no fill log requested and no agent order can be resting in any book
  if (!h->fill_log && !h->agent_orders_possible && !h->force_general) {
    int rc = launch_replay_fast(h, p, (cudaStream_t)stream);
    if (rc != 1) return rc; // 1: no compiled StaticLayout matches these capacities
  }
  return launch_advance<false, true>(h, p, (cudaStream_t)stream);
}

int lobsim_forward_step(lobsim_t* h, int32_t n_steps, void* stream) {
  if (!h || n_steps < 0) return fail(LOBSIM_E_INVALID, "bad argument");
  AdvParams p; base_params(h, p);
  p.T = n_steps; p.agent_kind = LOBSIM_AGENT_NONE; p.resync_last_only = 1;
  return launch_advance<false, true>(h, p, (cudaStream_t)stream);
}

int lobsim_set_book(lobsim_t* h, int32_t env, const lobsim_book_entry_t* buy, int32_t n_buy, const lobsim_book_entry_t* sell, int32_t n_sell) {
  if (!h || env < 0 || env >= h->cfg.n_envs || n_buy < 0 || n_sell < 0 || (!buy && n_buy) || (!sell && n_sell)) return fail(LOBSIM_E_INVALID, "bad argument");
  const Layout& L = h->L;
  if (n_buy > L.NO || n_sell > L.NO) return fail(LOBSIM_E_INVALID, "more orders than max_orders_per_side");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaDeviceSynchronize());
  std::vector<unsigned char> buf(L.blob_bytes);
  unsigned char* gblob = h->blobs + (size_t)env * L.blob_bytes;
  CUDA_TRY(cudaMemcpy(buf.data(), gblob, L.blob_bytes, cudaMemcpyDeviceToHost));
  BookHdr* hd = reinterpret_cast<BookHdr*>(buf.data());
  uint32_t max_agent_id = 0;
  for (int side = 0; side < 2; side++) {
    const lobsim_book_entry_t* e = side ? sell : buy;
    const int n = side ? n_sell : n_buy;
    unsigned char* sb = buf.data() + L.side_off + side * L.side_stride;
    int32_t* lvp = reinterpret_cast<int32_t*>(sb);
    uint16_t* lvend = reinterpret_cast<uint16_t*>(sb + L.lvend_off);
    uint2* ord = reinterpret_cast<uint2*>(sb + L.ord_off);
    int32_t* ap = reinterpret_cast<int32_t*>(buf.data() + L.agent_off + side * L.NA * 12);
    int32_t* av = ap + L.NA;
    uint32_t* ai = reinterpret_cast<uint32_t*>(ap + 2 * L.NA);
    // entries arrive best level first; storage is worst level first => walk the levels backwards
    std::vector<std::pair<int, int>> levels; // [first, last) index ranges of equal price, best first
    for (int i = 0; i < n;) {
      int j = i;
      while (j < n && e[j].price == e[i].price) j++;
      if (!levels.empty()) {
        int prev = e[levels.back().first].price;
        if (side == 0 ? e[i].price >= prev : e[i].price <= prev) return fail(LOBSIM_E_INVALID, "entries must be sorted best level first");
      }
      levels.push_back({i, j});
      i = j;
    }
    if ((int)levels.size() > L.NL) return fail(LOBSIM_E_INVALID, "more levels than max_levels_per_side");
    int nlv = (int)levels.size(), pos = 0, nag = 0;
    std::vector<std::pair<uint32_t, int>> agent; // (id, entry index)
    for (int k = nlv - 1; k >= 0; k--) {
      int j = nlv - 1 - k;
      lvp[j] = e[levels[k].first].price;
      for (int i = levels[k].first; i < levels[k].second; i++) {
        if (e[i].volume < 0) return fail(LOBSIM_E_INVALID, "negative volume");
        ord[pos++] = make_uint2((unsigned)e[i].volume, e[i].ref);
        if (e[i].ref & LOBSIM_REF_AGENT) agent.push_back({e[i].ref & 0x7fffffffu, i});
      }
      lvend[j] = (uint16_t)pos;
    }
    if ((int)agent.size() > L.NA) return fail(LOBSIM_E_INVALID, "more agent orders than max_agent_orders");
    std::sort(agent.begin(), agent.end());
    for (auto& a : agent) { ap[nag] = e[a.second].price; av[nag] = e[a.second].volume; ai[nag] = a.first; nag++; if (a.first > max_agent_id) max_agent_id = a.first; }
    hd->cnt[side][0] = nlv; hd->cnt[side][1] = pos; hd->nag[side] = nag;
  }
  if (hd->next_agent_id <= max_agent_id) hd->next_agent_id = max_agent_id + 1;
  hd->dead = 0; hd->err = 0;
  CUDA_TRY(cudaMemcpy(gblob, buf.data(), L.blob_bytes, cudaMemcpyHostToDevice));
  h->agent_orders_possible = true;
  return LOBSIM_OK;
}

int lobsim_get_state_dev(lobsim_t* h, lobsim_env_state_t* out_dev, void* stream) {
  if (!h || !out_dev) return fail(LOBSIM_E_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  k_get_state<<<(h->cfg.n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->blobs, h->L, 0, h->cfg.n_envs, out_dev);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  return LOBSIM_OK;
}

int lobsim_get_state(lobsim_t* h, int32_t first_env, int32_t n, lobsim_env_state_t* out) {
  if (!h || !out || first_env < 0 || n < 0 || first_env + n > h->cfg.n_envs) return fail(LOBSIM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_staging(h);
  if (rc) return rc;
  if (n == 0) return LOBSIM_OK;
  k_get_state<<<(n + 127) / 128, 128>>>(h->blobs, h->L, first_env, n, h->st_state);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  CUDA_TRY(cudaMemcpy(out, h->st_state, (size_t)n * sizeof(lobsim_env_state_t), cudaMemcpyDeviceToHost));
  return LOBSIM_OK;
}

int lobsim_replay_host(lobsim_t* h, int stream_id, const lobsim_msg_t* msgs_host, uint64_t first_msg, uint64_t n_msgs, int32_t n_steps, lobsim_env_state_t* state_out) {
  if (!h || stream_id < 0 || stream_id >= (int)h->streams.size()) return fail(LOBSIM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc = ensure_staging(h);
  if (rc) return rc;
  lobsim_stream_t& s = h->streams[stream_id];
  if (first_msg + n_msgs > s.n_msgs) return fail(LOBSIM_E_INVALID, "segment outside the stream");
  if (n_msgs) { // H2D of the segment straight into the resident stream buffer (same layout, same offsets)
    if (!msgs_host) return fail(LOBSIM_E_INVALID, "msgs_host is null");
    CUDA_TRY(cudaMemcpyAsync(const_cast<lobsim_msg_t*>(s.msgs) + first_msg, msgs_host, n_msgs * sizeof(lobsim_msg_t), cudaMemcpyHostToDevice, 0));
  }
  rc = lobsim_replay(h, n_steps, nullptr);
  if (rc) return rc;
  if (state_out) {
    rc = lobsim_get_state_dev(h, h->st_state, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(state_out, h->st_state, (size_t)h->cfg.n_envs * sizeof(lobsim_env_state_t), cudaMemcpyDeviceToHost, 0));
  }
  CUDA_TRY(cudaStreamSynchronize(0));
  return LOBSIM_OK;
}

int lobsim_process_orders(lobsim_t* h, const lobsim_order_t* orders, int32_t n, lobsim_fill_t* fills_out, int32_t max_fills, int32_t* n_fills_out, uint32_t* refs_out) {
  if (!h || (!orders && n) || n < 0 || max_fills < 0) return fail(LOBSIM_E_INVALID, "bad argument");
  CUDA_TRY(cudaSetDevice(h->device));
  if (n_fills_out) *n_fills_out = 0;
  if (n == 0) return LOBSIM_OK;
  lobsim_order_t* d_orders = nullptr; lobsim_fill_t* d_fills = nullptr; int32_t* d_nf = nullptr; uint32_t* d_refs = nullptr;
  CUDA_TRY(cudaMalloc(&d_orders, (size_t)n * sizeof(lobsim_order_t)));
  CUDA_TRY(cudaMalloc(&d_fills, (size_t)(max_fills > 0 ? max_fills : 1) * sizeof(lobsim_fill_t)));
  CUDA_TRY(cudaMalloc(&d_nf, sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&d_refs, (size_t)n * sizeof(uint32_t)));
  CUDA_TRY(cudaMemcpy(d_orders, orders, (size_t)n * sizeof(lobsim_order_t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(d_nf, 0, sizeof(int32_t)));
  OrdParams p; p.blobs = h->blobs; p.L = h->L; p.n_envs = h->cfg.n_envs; p.orders = d_orders; p.n = n;
  p.fills = max_fills > 0 ? d_fills : nullptr; p.max_fills = max_fills; p.n_fills = d_nf; p.refs_out = d_refs;
  h->agent_orders_possible = true;
  k_process_orders<<<1, 32, h->L.blob_bytes>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  h->launches++;
  int32_t nf = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&nf, d_nf, sizeof nf, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && fills_out && nf > 0) e = cudaMemcpy(fills_out, d_fills, (size_t)nf * sizeof(lobsim_fill_t), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && refs_out) e = cudaMemcpy(refs_out, d_refs, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaFree(d_orders); cudaFree(d_fills); cudaFree(d_nf); cudaFree(d_refs);
  if (e != cudaSuccess) return fail(LOBSIM_E_CUDA, cudaGetErrorString(e));
  if (n_fills_out) *n_fills_out = nf;
  return LOBSIM_OK;
}

// host-side decode of one env's blob (layout = book.cuh)
static int fetch_blob(lobsim* h, int env, std::vector<unsigned char>& buf) {
  if (env < 0 || env >= h->cfg.n_envs) return fail(LOBSIM_E_INVALID, "bad env index");
  CUDA_TRY(cudaSetDevice(h->device));
  buf.resize(h->L.blob_bytes);
  CUDA_TRY(cudaMemcpy(buf.data(), h->blobs + (size_t)env * h->L.blob_bytes, h->L.blob_bytes, cudaMemcpyDeviceToHost));
  return LOBSIM_OK;
}

int lobsim_dump_book(lobsim_t* h, int32_t env, int32_t side, lobsim_book_entry_t* out, int32_t capacity) {
  if (!h || side < 0 || side > 1) return fail(LOBSIM_E_INVALID, "bad argument");
  std::vector<unsigned char> buf;
  int rc = fetch_blob(h, env, buf);
  if (rc) return rc;
  const Layout& L = h->L;
  const BookHdr* hd = reinterpret_cast<const BookHdr*>(buf.data());
  const unsigned char* sb = buf.data() + L.side_off + side * L.side_stride;
  const int32_t* lvp = reinterpret_cast<const int32_t*>(sb);
  const uint16_t* lvend = reinterpret_cast<const uint16_t*>(sb + L.lvend_off);
  const uint2* ord = reinterpret_cast<const uint2*>(sb + L.ord_off);
  int n = 0;
  for (int k = 0; k < hd->cnt[side][0]; k++) { // best level first
    int j = hd->cnt[side][0] - 1 - k;
    int start = j > 0 ? lvend[j - 1] : 0, end = lvend[j];
    for (int i = start; i < end; i++) {
      if (out && n < capacity) { out[n].price = lvp[j]; out[n].volume = (int32_t)ord[i].x; out[n].ref = ord[i].y; out[n].level = k; }
      n++;
    }
  }
  return n;
}

int lobsim_dump_agent_orders(lobsim_t* h, int32_t env, int32_t side, lobsim_book_entry_t* out, int32_t capacity) {
  if (!h || side < 0 || side > 1) return fail(LOBSIM_E_INVALID, "bad argument");
  std::vector<unsigned char> buf;
  int rc = fetch_blob(h, env, buf);
  if (rc) return rc;
  const Layout& L = h->L;
  const BookHdr* hd = reinterpret_cast<const BookHdr*>(buf.data());
  const int32_t* ap = reinterpret_cast<const int32_t*>(buf.data() + L.agent_off + side * L.NA * 12);
  const int32_t* av = ap + L.NA;
  const uint32_t* ai = reinterpret_cast<const uint32_t*>(ap + 2 * L.NA);
  int n = hd->nag[side];
  for (int i = 0; i < n && out && i < capacity; i++) { out[i].price = ap[i]; out[i].volume = av[i]; out[i].ref = LOBSIM_REF_AGENT | ai[i]; out[i].level = -1; }
  return n;
}

int lobsim_get_fills(lobsim_t* h, int32_t env, lobsim_fill_t* out, int32_t capacity, int32_t* n_out) {
  if (!h || env < 0 || env >= h->cfg.n_envs || !n_out) return fail(LOBSIM_E_INVALID, "bad argument");
  *n_out = 0;
  if (!h->fill_log) return LOBSIM_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  int32_t n = 0;
  CUDA_TRY(cudaMemcpy(&n, h->fill_count + env, sizeof n, cudaMemcpyDeviceToHost));
  *n_out = n;
  int m = n < h->cfg.fill_log_capacity ? n : h->cfg.fill_log_capacity;
  if (m > capacity) m = capacity;
  if (out && m > 0) CUDA_TRY(cudaMemcpy(out, h->fill_log + (size_t)env * h->cfg.fill_log_capacity, (size_t)m * sizeof(lobsim_fill_t), cudaMemcpyDeviceToHost));
  return LOBSIM_OK;
}

int lobsim_errors(lobsim_t* h, uint32_t* err_out) {
  if (!h || !err_out) return fail(LOBSIM_E_INVALID, "null argument");
  std::vector<lobsim_env_state_t> st(h->cfg.n_envs);
  int rc = lobsim_get_state(h, 0, h->cfg.n_envs, st.data());
  if (rc) return rc;
  for (int i = 0; i < h->cfg.n_envs; i++) err_out[i] = st[i].err;
  return LOBSIM_OK;
}

int64_t lobsim_launch_count(lobsim_t* h) { return h ? h->launches : 0; }

} // extern "C"
