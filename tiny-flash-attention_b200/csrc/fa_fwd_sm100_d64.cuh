// fa_fwd_sm100_d64.cuh -- head_dim 64 variant of the persistent attention-forward kernel (sm_100a, B200).
//
// At D=64 the tensor work per KV tile halves while the softmax work per score does not: the one-thread-per-row kernel is
// bound by the LATENCY of one warp walking 128 columns (profiles/r01: 2750 cycles per KV tile pair for 1552 cycles of
// tensor work; no pipe is more than half busy).  D=64 leaves TMEM room that D=128 does not have, and this kernel spends
// it on instruction-level parallelism instead of a deeper S pipeline:
//
//   * every 128-row Q tile is served by TWO softmax warpgroups: (t, h) owns key columns [64h, 64h+64) of every KV tile of
//     tile t -- 16 softmax warps, four per SM sub-partition instead of two;
//   * the two halves are INDEPENDENT online-softmax streams: each keeps its own running max / sum and its own O
//     accumulator in TMEM (S0 | S1 | O0A | O0B | O1A | O1B = 512 columns), so no per-tile exchange or barrier couples them;
//     P_{t,h} aliases the first 32 columns of its half of S_t and feeds its own TS-form PV chain (K = 64 keys);
//   * the halves are merged ONCE per work item in the epilogue, exactly like split-KV partials: m = max(mA, mB),
//     O = (OA 2^((mA-m)c) + OB 2^((mB-m)c)) / (lA 2^((mA-m)c) + lB 2^((mB-m)c)); thread (row, h) produces output columns
//     [32h, 32h+32) from both accumulators (TMEM loads are per lane, any column), so O never travels through smem.
//
// Everything between work items (atomic scheduler, cross-item K/V ring, hoisted first S, TMA-store epilogue, fused
// peer stores) is the persistent design of fa_fwd_sm100_persist.cuh.  Arithmetic per score is unchanged (reference:
// /root/reference/flash_attention_cutlass/csrc/flash_attention.cu:228-316,601).
//
// Warp roles (640 threads): warps 0-15 softmax (warpgroup g = warp/4: tile g/2, half g%2), warp 16 scheduler + TMA
// producer, warp 17 TMEM allocator + UMMA issuer, warps 18-19 idle (register donors).
#pragma once
#include "fa_fwd_sm100_persist.cuh"

namespace tfa {

struct P64Cfg {
  static constexpr int D = 64;
  static constexpr int BM = 128, BN = 128;
  static constexpr int TILE_BYTES = 128 * 128;                      // one 64-column slab of 128 rows
  static constexpr int NSTAGE = 8, NSTAGE_LOG2 = 3;
  static constexpr int STG_WARP_BYTES = 32 * 64;                    // epilogue staging: 32 rows x 32 columns (64 B) per warp
  static constexpr int STG_BYTES = 16 * STG_WARP_BYTES;
  static constexpr int XCH_BYTES = 2 * 128 * 2 * 8;                 // [tile][row][half] {m, l}
  static constexpr uint32_t Q_FULL = 0, Q_EMPTY = 2, KV_FULL = 4, KV_EMPTY = KV_FULL + NSTAGE, S_FULL = KV_EMPTY + NSTAGE,
                            P_HALF = S_FULL + 2, P_FULL = P_HALF + 2, O_FULL = P_FULL + 2, SCHED_FULL = O_FULL + 2,
                            SCHED_EMPTY = SCHED_FULL + 2, NUM_BARS = SCHED_EMPTY + 2;
  static constexpr int SMEM_BYTES = 1024 + 2 * TILE_BYTES + NSTAGE * TILE_BYTES + STG_BYTES + XCH_BYTES + NUM_BARS * 8 + 32;
  static constexpr int TM_S0 = 0, TM_S1 = 128, TM_O = 256;          // O_{t,h} at TM_O + (2t+h)*64
  static constexpr int TM_COLS = 512;
  static constexpr int THREADS = 640;
  // 640 threads launch with 96 registers each (61440); 16 softmax warps x 104 + 4 service warps x 40 = 58368
  static constexpr uint32_t REGS_SOFTMAX = 104, REGS_OTHER = 40;
};
static_assert((512 * P64Cfg::REGS_SOFTMAX + 128 * P64Cfg::REGS_OTHER) <= 640 * 96, "setmaxnreg budget exceeds the CTA pool");

template <bool CAUSAL, bool IS_BF16, bool OUT_F32>
__global__ void __launch_bounds__(640, 1)
fa_fwd_sm100_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ OutMaps tmO, const FwdParams p) {
  using C = P64Cfg;
  constexpr int D = 64;
  constexpr int TILE = C::TILE_BYTES;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr uint32_t SLOT_LO = TILE >> 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + 2 * TILE;
  uint8_t* sStg = sKV + NSTAGE * TILE;
  float2* sXch = reinterpret_cast<float2*>(sStg + C::STG_BYTES);   // [2][128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sXch) + C::XCH_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](uint32_t which, uint32_t i) -> uint32_t { return bar_base + 8u * (which + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);
  volatile int* sched_ring = reinterpret_cast<volatile int*>(tmem_slot + 2);
  const uint32_t sQ_addr = smem_u32(sQ);
  const uint32_t sKV_addr = smem_u32(sKV);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.total_items;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (!OUT_F32) tma_prefetch_desc(&tmO.m[0]);
    for (uint32_t t = 0; t < 2; ++t) {
      mbar_init(bar(C::Q_FULL, t), 1);
      mbar_init(bar(C::Q_EMPTY, t), 1);
      mbar_init(bar(C::S_FULL, t), 1);
      mbar_init(bar(C::O_FULL, t), 1);
      mbar_init(bar(C::SCHED_FULL, t), 1);
      mbar_init(bar(C::SCHED_EMPTY, t), 17);   // UMMA warp + 16 softmax warps
    }
    for (uint32_t t = 0; t < 2; ++t) {
      mbar_init(bar(C::P_HALF, t), 8);         // one arrival per warp of BOTH half-warpgroups of tile t: the issuer pays
      mbar_init(bar(C::P_FULL, t), 8);         // two waits per tile, not four (its fixed costs dominate at D=64)
    }
    for (uint32_t i = 0; i < NSTAGE; ++i) {
      mbar_init(bar(C::KV_FULL, i), 1);
      mbar_init(bar(C::KV_EMPTY, i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc(tmem_slot, C::TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto read_tmem_base = [&]() {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(tmem_slot)));
    return v;
  };
  auto sched_get = [&](int k) -> int {
    mbar_wait(bar(C::SCHED_FULL, k & 1), (k >> 1) & 1, p.dbg, SITE_P_SCHED_FULL);
    // REDUX makes the item number PROVABLY warp-uniform for the compiler.  Without it everything derived from a value
    // loaded from shared memory (tile counts, loop bounds, ring slots, barrier parities, MMA descriptors) is treated as
    // divergent: the issuer's operands go through vector registers + R2UR, every loop gets reconvergence scaffolding --
    // measured on B200 (r02, profiles/r02_persist_bisect.txt): 25 % slower steady state, the whole deficit of the first
    // persistent kernels, which round 1 had attributed to the instruction cache.
    const int item = __reduce_max_sync(0xffffffffu, sched_ring[k & 1]);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(C::SCHED_EMPTY, k & 1));
    return item;
  };

  if (warp == 16) {
    // ==================== scheduler + TMA producer (same protocol as fa_fwd_sm100_persist.cuh) ====================
    setmaxnreg_dec<C::REGS_OTHER>();
    if (lane == 0) {
      bool first = true;                       // CTA c starts with item c: no atomic in front of the first loads
      auto fetch = [&]() -> int {
        for (;;) {
          const int i = first ? static_cast<int>(blockIdx.x) : atomicAdd(p.sched_counter, 1) + static_cast<int>(gridDim.x);
          first = false;
          if (i >= total) {
            if (atomicAdd(p.sched_counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
              p.sched_counter[0] = 0;
              p.sched_counter[1] = 0;
            }
            return total;
          }
          if (decode_pitem<CAUSAL>(i, p).nmax > 0) return i;
        }
      };
      auto publish = [&](int k, int item) {
        mbar_wait(bar(C::SCHED_EMPTY, k & 1), ((k >> 1) & 1) ^ 1, p.dbg, SITE_P_SCHED_EMPTY);
        sched_ring[k & 1] = item;
        mbar_arrive(bar(C::SCHED_FULL, k & 1));     // release: the store above is visible to the waiters (acquire in try_wait)
      };
      uint32_t ent = 0, qpar = 0;
      int k = 0;
      int cur = fetch();
      publish(0, cur);
      while (cur < total) {
        const PItem w = decode_pitem<CAUSAL>(cur, p);
        auto load_q = [&](int t) {
          if (w.nblk[t] > 0) {
            mbar_wait(bar(C::Q_EMPTY, t), ((qpar >> t) & 1u) ^ 1u, p.dbg, SITE_P_QEMPTY);
            qpar ^= (1u << t);
            mbar_arrive_expect_tx(bar(C::Q_FULL, t), TILE);
            tma_load_4d(sQ_addr + t * TILE, &tmQ, bar(C::Q_FULL, t), 0, w.row0[t], w.hidx, w.bidx);
          }
        };
        auto load_kv = [&](int j, int kv) {
          const uint32_t slot = ent & (NSTAGE - 1);
          const uint32_t par = (ent >> C::NSTAGE_LOG2) & 1u;
          mbar_wait(bar(C::KV_EMPTY, slot), par ^ 1u, p.dbg, SITE_LOAD_EMPTY);
          mbar_arrive_expect_tx(bar(C::KV_FULL, slot), TILE);
          tma_load_4d(sKV_addr + slot * TILE, (kv == 0) ? &tmK : &tmV, bar(C::KV_FULL, slot), 0, (w.jb + j) * C::BN, w.hkv,
                      w.bidx);
          ++ent;
        };
        load_q(0);
        load_kv(0, 0);
        load_kv(0, 1);
        load_q(1);
        for (int j = 1; j < w.nmax; ++j) {
          load_kv(j, 0);
          load_kv(j, 1);
        }
        const int nxt = fetch();               // drawn as late as the pipeline allows (see fa_fwd_sm100_persist.cuh)
        publish(k + 1, nxt);
#if TFA_Q_PREFETCH
        if (nxt < total) {
          const PItem wn = decode_pitem<CAUSAL>(nxt, p);
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (wn.nblk[t] > 0) tma_prefetch_l2_4d(&tmQ, 0, wn.row0[t], wn.hidx, wn.bidx);
        }
#endif
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp == 17) {
    // =========================== UMMA issuer ===========================
    setmaxnreg_dec<C::REGS_OTHER>();
    {
      const uint32_t tmem_base = read_tmem_base();
      constexpr uint32_t FMT = IS_BF16 ? 1u : 0u;
      const uint32_t idescS = umma_idesc_f16(FMT, 128, 128, 0, 0);  // A,B K-major
      const uint32_t idescO = umma_idesc_f16(FMT, 128, D, 0, 1);    // B (=V) MN-major, N = 64
      auto opaque = [](uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
      const uint32_t q_lo0 = umma_desc_lo(sQ_addr, 16);
      const uint32_t k_lo_base = umma_desc_lo(sKV_addr, 16);
      const uint32_t v_lo_base = umma_desc_lo(sKV_addr, 16384);     // LBO unused at N = 64 (one slab)

      auto issue_S = [&](int t, uint32_t kslot, bool release_kv, bool release_q) {
        const uint32_t q_lo = opaque(q_lo0) + t * SLOT_LO;
        const uint32_t k_lo = opaque(k_lo_base) + kslot * SLOT_LO;
        const uint32_t d_tmem = opaque(tmem_base) + t * (C::TM_S1 - C::TM_S0);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) umma_ss_lo(d_tmem, q_lo + k * 2, k_lo + k * 2, idescS, k > 0 ? 1u : 0u);
          umma_commit(bar(C::S_FULL, t));
          if (release_kv) umma_commit(bar(C::KV_EMPTY, kslot));
          if (release_q) umma_commit(bar(C::Q_EMPTY, t));
        }
        __syncwarp();
      };
      // O_{t,h} += P_{t,h} V[64h + 16k .. ) for k-steps [k0, k1) of the half's four
      auto issue_PV = [&](int t, int h, uint32_t vslot, bool acc, int k0, int k1, bool release_kv, bool done) {
        const uint32_t v_lo = opaque(v_lo_base) + vslot * SLOT_LO + h * (64 * 128 >> 4);   // 64 key rows = 8192 B
        const uint32_t tb = opaque(tmem_base);
        const uint32_t d_tmem = tb + C::TM_O + (2 * t + h) * 64;
        const uint32_t p_tmem = tb + t * (C::TM_S1 - C::TM_S0) + h * 64;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k >= k0 && k < k1) umma_ts_lo(d_tmem, p_tmem + k * 8, v_lo + k * 128, idescO, (acc || k > 0) ? 1u : 0u);
          }
          if (release_kv) umma_commit(bar(C::KV_EMPTY, vslot));
          if (done) umma_commit(bar(C::O_FULL, t));
        }
        __syncwarp();
      };
      auto ent_slot = [&](uint32_t e) { return e & (NSTAGE - 1); };
      auto ent_par = [&](uint32_t e) { return (e >> C::NSTAGE_LOG2) & 1u; };

      uint32_t ent_base = 0;
      uint32_t st = 0;      // bit t: q_full parity | bit 2+t: P barriers' parity | bit 4+t: first S of the next item hoisted
      int k = 0;
      int cur = sched_get(0);
      while (cur < total) {
        int n0, n1, nn0 = 0, nn1 = 0;
        {
          const PItem x = decode_pitem<CAUSAL>(cur, p);
          n0 = x.nblk[0];
          n1 = x.nblk[1];
        }
        const int nmax = max(n0, n1);
        int nxt = -1;                                        // picked up lazily, see fa_fwd_sm100_persist.cuh
        auto poll_nxt = [&](bool block) {
          if (nxt >= 0) return;
          if (!block && !__all_sync(0xffffffffu, mbar_test_wait(bar(C::SCHED_FULL, (k + 1) & 1), ((k + 1) >> 1) & 1))) return;
          nxt = sched_get(k + 1);
          if (nxt < total) {
            const PItem x = decode_pitem<CAUSAL>(nxt, p);
            nn0 = x.nblk[0];
            nn1 = x.nblk[1];
          }
        };
        const uint32_t ent_next = ent_base + 2u * static_cast<uint32_t>(nmax);
        auto first_S = [&](int t, int x_nt, bool other_done, uint32_t e0) {
          mbar_wait(bar(C::Q_FULL, t), (st >> t) & 1u, p.dbg, SITE_P_FIRST_Q);
          st ^= (1u << t);
          mbar_wait(bar(C::KV_FULL, ent_slot(e0)), ent_par(e0), p.dbg, SITE_P_FIRST_K);
          tc_fence_after();
          issue_S(t, ent_slot(e0), other_done, x_nt == 1);
        };
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int nt = (t == 0) ? n0 : n1;
          const int no = (t == 0) ? n1 : n0;
          if (nt > 0 && !((st >> (4 + t)) & 1u)) {
            const bool other_done = (no == 0) || ((st >> (4 + (t ^ 1))) & 1u) || (t == 1);
            first_S(t, nt, other_done, ent_base);
            st |= (1u << (4 + t));
          }
        }
        st &= ~(3u << 4);

        for (int j = 0; j < nmax; ++j) {
          const uint32_t ev = ent_base + 2u * j + 1u, ek = ent_base + 2u * j + 2u;
          const uint32_t vslot = ent_slot(ev), kslot = ent_slot(ek);
          // the 8-deep ring is far ahead: these waits are satisfied except right behind an item boundary
          mbar_wait(bar(C::KV_FULL, vslot), ent_par(ev), p.dbg, SITE_MMA_V);
          if (j + 1 < nmax) mbar_wait(bar(C::KV_FULL, kslot), ent_par(ek), p.dbg, SITE_MMA_K);
#if TFA_ISSUER_UNROLL_T
#pragma unroll
#else
#pragma unroll 1
#endif
          for (int t = 0; t < 2; ++t) {
            const int nt = (t == 0) ? n0 : n1;
            const int no = (t == 0) ? n1 : n0;
            const bool active = (j < nt);
            const bool last_v_user = (t == 1) || (j >= no);
            const bool last_k_user = (t == 1) || (j + 1 >= no);
            const bool has_next = (j + 1 < nt);
            if (active) {
              const uint32_t ppar = (st >> (2 + t)) & 1u;
              mbar_wait(bar(C::P_HALF, t), ppar, p.dbg, SITE_MMA_PH);
              tc_fence_after();
              issue_PV(t, 0, vslot, j > 0, 0, 2, false, false);
              issue_PV(t, 1, vslot, j > 0, 0, 2, false, false);
              mbar_wait(bar(C::P_FULL, t), ppar, p.dbg, SITE_MMA_P);
              st ^= (1u << (2 + t));
              tc_fence_after();
              issue_PV(t, 0, vslot, true, 2, 4, false, false);
              issue_PV(t, 1, vslot, true, 2, 4, last_v_user, !has_next);
            }
            // ONE S site (instruction-cache footprint, see fa_fwd_sm100_persist.cuh): next KV tile, or the hoisted first
            // S of the next item as soon as its Q_t and K_0 have landed (non-blocking probe, retried while the other tile
            // is still running)
            bool do_S = active && has_next, rel_kv = last_k_user, rel_q = (j + 2 == nt);
            uint32_t s_slot = kslot;
            if (!do_S) {
              bool landed = false;
              if (TFA_HOIST && !((st >> (4 + t)) & 1u)) poll_nxt(false);
              const bool has_nxt = nxt >= 0 && nxt < total;
              const int nnt = (t == 0) ? nn0 : nn1;
              const int nno = (t == 0) ? nn1 : nn0;
              if (TFA_HOIST && has_nxt && nnt > 0 && !((st >> (4 + t)) & 1u)) {
                const bool q_ok = mbar_test_wait(bar(C::Q_FULL, t), (st >> t) & 1u);
                const bool k_ok = mbar_test_wait(bar(C::KV_FULL, ent_slot(ent_next)), ent_par(ent_next));
                landed = __all_sync(0xffffffffu, q_ok && k_ok);
              }
              if (landed) {
                st ^= (1u << t);
                rel_kv = (nno == 0) || ((st >> (4 + (t ^ 1))) & 1u);
                rel_q = (nnt == 1);
                s_slot = ent_slot(ent_next);
                st |= (1u << (4 + t));
                do_S = true;
                tc_fence_after();
              }
            }
            if (do_S) issue_S(t, s_slot, rel_kv, rel_q);
          }
        }
        poll_nxt(true);
        ent_base = ent_next;
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp < 16) {
    // ================= softmax warpgroup (t, h): key columns [64h, 64h+64) of every KV tile of Q tile t =================
    setmaxnreg_inc<C::REGS_SOFTMAX>();
    const int g = warp >> 2;
    const int t = g >> 1, h = g & 1;
    const int r = ((warp & 3) << 5) | lane;                 // row inside the Q tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_base = read_tmem_base();
    const uint32_t tS = tmem_base + lane_base + (t == 0 ? C::TM_S0 : C::TM_S1) + h * 64;   // my half of S_t; P aliases [0,32)
    const uint32_t tOh = tmem_base + lane_base + C::TM_O + (2 * t + h) * 64;               // my accumulator
    const uint32_t tOt = tmem_base + lane_base + C::TM_O + (2 * t) * 64;                   // O_{t,A}; O_{t,B} = +64
    const float c = p.scale_log2;
    const int S = p.S, Sk = p.Sk;
    const int wq = __reduce_max_sync(0xffffffffu, warp & 3);   // warp index inside the warpgroup, PROVABLY uniform
    const uint32_t stg = smem_u32(sStg) + warp * C::STG_WARP_BYTES;
    float2* xch = sXch + (t * 128 + r) * 2;

    uint32_t scnt = 0, ocnt = 0;
    for (int k = 0;; ++k) {
      const int item = sched_get(k);
      if (item >= total) break;
      const PItem w = decode_pitem<CAUSAL>(item, p);
      const int n = (t == 0) ? w.nblk[0] : w.nblk[1];
      if (n == 0) continue;
      const int trow0 = (t == 0) ? w.row0[0] : w.row0[1];
      const int row_g = trow0 + r;
      const int jb = w.jb;

      float m_ref = 0.f, l = 0.f;
      for (int j = 0; j < n; ++j) {
        mbar_wait(bar(C::S_FULL, t), scnt & 1u, p.dbg, SITE_SM_S);
        tc_fence_after();
        // Register budget: 104 per softmax thread (640-thread CTA).  The row max needs all 64 scores, the exponentials
        // are taken 32 at a time: keys 0..31 stay in registers from the first load, keys 32..63 are dropped after the
        // max and loaded AGAIN after P[0..31] has been handed over (P quarter 0 overwrites S columns 0..15 only).
        uint32_t sa[32];
        const int col0 = (jb + j) * C::BN + h * 64;          // first key of my half
        int lim = Sk - col0;
        if (CAUSAL) lim = min(lim, row_g + p.causal_off - col0 + 1);
        float mx;
        uint32_t dead = 0;                                   // bit c: 32-key chunk c of my half is masked for the whole warp
        {
          uint32_t sb[32];
          tmem_ld_x32(tS, sa);
          tmem_ld_x32(tS + 32, sb);
          tmem_wait_ld();
          // masking by 32-key chunk, warp-uniformly (see fa_fwd_sm100_persist.cuh): untouched / dead / mixed
          {
            int lim_lo = Sk - col0, lim_hi = lim_lo;
            if (CAUSAL) {
              const int b0 = trow0 + wq * 32 + p.causal_off - col0 + 1;
              lim_lo = min(lim_lo, b0);
              lim_hi = min(lim_hi, b0 + 31);
            }
            if (lim_lo < 64) {
              if (lim_hi <= 0) dead |= 1u;
              else if (lim_lo < 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i >= lim) sa[i] = 0xff800000u;
              }
              if (lim_hi <= 32) dead |= 2u;
              else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i + 32 >= lim) sb[i] = 0xff800000u;
              }
            }
          }
          float mxa = -INFINITY, mxb = -INFINITY, mxc = -INFINITY, mxd = -INFINITY;
          if (!(dead & 1u)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mxa = fmax3(mxa, __uint_as_float(sa[i]), __uint_as_float(sa[i + 1]));
              mxb = fmax3(mxb, __uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3]));
            }
          }
          if (!(dead & 2u)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              mxc = fmax3(mxc, __uint_as_float(sb[i]), __uint_as_float(sb[i + 1]));
              mxd = fmax3(mxd, __uint_as_float(sb[i + 2]), __uint_as_float(sb[i + 3]));
            }
          }
          mx = fmaxf(fmaxf(mxa, mxc), fmaxf(mxb, mxd));
        }
        if (j == 0) {
          m_ref = fmaxf(mx, -1.0e30f);
        } else {
          const bool need = (mx - m_ref) * c > kRescaleThresholdLog2;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? mx : m_ref;
            const float alpha = ex2_approx((m_ref - m_new) * c);
            m_ref = m_new;
            l *= alpha;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t o[16];
              tmem_ld_x16(tOh + ch * 16, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(tOh + ch * 16, o);
            }
          }
        }
        constexpr int kEmuPairsPer8 = kEmuPairsPer8For<64>;
        const float2 c2 = make_float2(c, c);
        const float2 nm2 = make_float2(-m_ref * c, -m_ref * c);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          const bool dq = (dead >> qt) & 1u;                  // warp-uniform: every key of this quarter is masked
          if (qt == 1 && !dq) {
            tmem_ld_x32(tS + 32, sa);                        // keys 32..63 again (their S columns are untouched so far)
            tmem_wait_ld();
            if (lim < 64) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i + 32 >= lim) sa[i] = 0xff800000u;
            }
          }
          uint32_t pk[16];
          if (dq) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int pi = qt * 16 + i;
              const float2 x = ffma2(make_float2(__uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1])), c2, nm2);
              float2 e;
              if (((pi * kEmuPairsPer8) & 7) < kEmuPairsPer8) {
                e = ex2_poly2(x);
              } else {
                e.x = ex2_approx(x.x);
                e.y = ex2_approx(x.y);
              }
              if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
              pk[i] = pack_16x2<IS_BF16>(e.x, e.y);
            }
          }
          tmem_st_x16(tS + qt * 16, pk);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(qt == 0 ? C::P_HALF : C::P_FULL, t));
        }
        acc0 = fadd2(acc0, acc1);
        l += acc0.x + acc0.y;
        ++scnt;
      }

      // ---------------------------- epilogue: merge the two halves, normalise, store ----------------------------
      mbar_wait(bar(C::O_FULL, t), ocnt & 1u, p.dbg, SITE_EPI_O);
      ++ocnt;
      tc_fence_after();
      xch[h] = make_float2(m_ref, l);
      named_bar_sync(1 + t, 256);
      const float2 other = xch[h ^ 1];
      const float mA = h == 0 ? m_ref : other.x, lA = h == 0 ? l : other.y;
      const float mB = h == 0 ? other.x : m_ref, lB = h == 0 ? other.y : l;
      const float m = fmaxf(mA, mB);
      const float a = ex2_approx((mA - m) * c), b = ex2_approx((mB - m) * c);
      float L = lA * a + lB * b;
      // a row none of whose keys lies in this item's KV range (split-KV partial above the row's causal limit): L = 0
      if ((CAUSAL ? min(Sk, row_g + p.causal_off + 1) : Sk) <= jb * C::BN) L = 0.f;
      const float inv = (L > 0.f) ? 1.0f / L : 0.f;
      const float fa = a * inv, fb = b * inv;
      if (h == 0 && p.lse != nullptr && row_g < S)
        p.lse[w.split * p.lse_part_stride + static_cast<long long>(w.bh) * p.lse_stride_bh + row_g] = m * p.scale + logf(L);

      // thread (row, h) produces output columns [32h, 32h + 32) from BOTH accumulators, 16 columns at a time
      uint32_t pk[16];
      {
        const long long tile_off =
            static_cast<long long>(w.bidx) * p.o_stride_b + static_cast<long long>(w.hidx) * p.o_stride_h;
        float* orow = OUT_F32 ? p.out_f32 + w.split * p.part_stride + tile_off + static_cast<long long>(row_g) * p.o_stride_s + h * 32
                              : nullptr;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t oa[16], ob[16];
          tmem_ld_x16(tOt + h * 32 + ch * 16, oa);
          tmem_ld_x16(tOt + 64 + h * 32 + ch * 16, ob);
          tmem_wait_ld();
          if constexpr (OUT_F32) {
            if (row_g < S) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 v4 = make_float4(__uint_as_float(oa[i]) * fa + __uint_as_float(ob[i]) * fb,
                                        __uint_as_float(oa[i + 1]) * fa + __uint_as_float(ob[i + 1]) * fb,
                                        __uint_as_float(oa[i + 2]) * fa + __uint_as_float(ob[i + 2]) * fb,
                                        __uint_as_float(oa[i + 3]) * fa + __uint_as_float(ob[i + 3]) * fb);
                *reinterpret_cast<float4*>(orow + ch * 16 + i) = v4;
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              pk[ch * 8 + i] = pack_16x2<IS_BF16>(__uint_as_float(oa[2 * i]) * fa + __uint_as_float(ob[2 * i]) * fb,
                                                  __uint_as_float(oa[2 * i + 1]) * fa + __uint_as_float(ob[2 * i + 1]) * fb);
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1 + t, 256);          // both halves have read both accumulators: PV of the next item may overwrite them
      if constexpr (!OUT_F32) {
        if (lane == 0) bulk_wait_group_read0();           // the previous item's store has read this staging
        __syncwarp();
        // 32 rows x 64 B, SWIZZLE_64B: 16-byte chunk q of row i lands at chunk q ^ ((i >> 1) & 3)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t addr = stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[q * 4 + 0]), "r"(pk[q * 4 + 1]),
                       "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int wrow0 = trow0 + (warp & 3) * 32;
          tma_store_4d(&tmO.m[0], stg, h * 32, wrow0, w.hidx, w.bidx);
          for (int d = 0; d < p.n_extra_dst; ++d) tma_store_4d(&tmO.m[1 + d], stg, h * 32, wrow0, w.hidx, w.bidx);
          bulk_commit_group();
        }
      }
    }
    if (!OUT_F32 && lane == 0) bulk_wait_group0();
  } else {
    setmaxnreg_dec<C::REGS_OTHER>();   // warps 18-19: idle, give their registers away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(read_tmem_base(), C::TM_COLS);
  }
}

}  // namespace tfa
