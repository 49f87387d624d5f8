// fa_fwd_sm100_persistent2.cuh -- EXPERIMENTAL (TFA_KERNEL=persistent2), written at the end of round 1 and NOT yet
// run on hardware; never selected by default, no test selects it.
//
// fa_fwd_sm100_persistent.cuh (persistent CTAs, atomic work counter, K/V ring running ahead across work items, first S
// of the next item hoisted into the current item's tail) predates four optimisations of the default kernel.  This file
// is that kernel with them ported, so that round 2 can measure what cross-item overlap is worth against TODAY's
// default:
//   * softmax: four back-to-back TMEM loads + one wait + four max chains (was: per-chunk load/max overlap);
//   * P produced in four quarters with three hand-offs p_half / p_3q / p_full (was: two halves, two hand-offs);
//   * issuer: PV in three stages (k-steps 0-3, 4-5, 6-7) and K/V readiness confirmed one tile ahead inside an item;
//   * work items are handed out in the default kernel's launch order (decode_work), not (b,h)-major;
// Everything else (scheduler ring, producer order, hoisting rule, per-warp epilogue staging, square MHA problems only)
// is unchanged from fa_fwd_sm100_persistent.cuh -- read that file's header for the design.
#pragma once
#include "fa_fwd_sm100_persistent.cuh"

namespace tfa {

// work item -> coordinates, in the launch order of the default kernel (decode_work: chunks of 8 heads, heaviest pair
// first across the chunk); the atomic counter hands items out in exactly this order
template <bool CAUSAL>
__device__ __forceinline__ WorkItem decode_item2(int item, const FwdParams& p) {
  WorkItem w;
  int split, pr;
  decode_work(item, p.npairs, 1, p.head_chunk, p.BH, w.bh, split, pr);
  w.bidx = w.bh / p.H;
  w.hidx = w.bh % p.H;
  const int nkv_total = (p.S + 127) / 128;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    w.row0[t] = pr * 256 + t * 128;
    const bool active = w.row0[t] < p.S;
    w.nblk[t] = active ? (CAUSAL ? min(nkv_total, w.row0[t] / 128 + 1) : nkv_total) : 0;
  }
  w.nmax = max(w.nblk[0], w.nblk[1]);
  return w;
}

template <int D>
struct P2Cfg : PFwdCfg<D> {
  static constexpr int NUM_BARS = PFwdCfg<D>::NUM_BARS + 2;          // + p_3q[2]
  static constexpr int SMEM_BYTES = PFwdCfg<D>::SMEM_BYTES + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

template <int D, bool CAUSAL, bool IS_BF16, bool OUT_F32>
__global__ void __launch_bounds__(384, 1)
fa_fwd_sm100_persistent2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using C = P2Cfg<D>;
  constexpr int TILE = C::TILE_BYTES;
  constexpr int NSTAGE = C::NSTAGE;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment (the swizzle is a function of address bits 7..9)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sKV = smem + 2 * TILE;           // NSTAGE tiles
  uint8_t* sStg = sKV + NSTAGE * TILE;      // 8 x 4 KB epilogue staging (one per softmax warp)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + C::STG_BYTES);
  // barriers are addressed by 32-bit shared addresses: bar_base + 8 * index
  const uint32_t bar_base = smem_u32(bars);
  constexpr uint32_t Q_FULL = 0, Q_EMPTY = 2, KV_FULL = 4, KV_EMPTY = KV_FULL + NSTAGE, S_FULL = KV_EMPTY + NSTAGE,
                     P_HALF = S_FULL + 2, P_FULL = P_HALF + 2, P_3Q = P_FULL + 2, O_FULL = P_3Q + 2, SCHED_FULL = O_FULL + 2,
                     SCHED_EMPTY = SCHED_FULL + 2, NBARS = SCHED_EMPTY + 2;
  static_assert(NBARS == C::NUM_BARS, "barrier table out of sync with FwdCfg::NUM_BARS");
  auto bar = [&](uint32_t which, uint32_t i) -> uint32_t { return bar_base + 8u * (which + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  volatile int* sched_ring = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2]
  const uint32_t sQ_addr = smem_u32(sQ);
  const uint32_t sKV_addr = smem_u32(sKV);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S;
  const int total = p.total_items;

  // ---- one-time setup ----
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(Q_FULL, t), 1);
      mbar_init(bar(Q_EMPTY, t), 1);
      mbar_init(bar(S_FULL, t), 1);
      mbar_init(bar(P_HALF, t), 4);      // one arrival per softmax warp
      mbar_init(bar(P_FULL, t), 4);
      mbar_init(bar(P_3Q, t), 4);
      mbar_init(bar(O_FULL, t), 1);
      mbar_init(bar(SCHED_FULL, t), 1);
      mbar_init(bar(SCHED_EMPTY, t), 9); // UMMA warp + 8 softmax warps
    }
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(bar(KV_FULL, i), 1);
      mbar_init(bar(KV_EMPTY, i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, C::TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // consumer side of the scheduler ring: item number k lives in slot k&1
  auto sched_get = [&](int k) -> int {
    mbar_wait(bar(SCHED_FULL, k & 1), (k >> 1) & 1, p.dbg, SITE_SCHED_FULL);
    const int item = sched_ring[k & 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(SCHED_EMPTY, k & 1));
    return item;
  };

  if (warp == 8) {
    // ==================== scheduler + TMA producer ====================
    setmaxnreg_dec<kRegsOtherPersistent>();
    if (lane == 0) {
      TFA_PTRACE_DECL(3, true)
      TFA_TRACE_EV(1);
      auto publish = [&](int k, int item) {
        mbar_wait(bar(SCHED_EMPTY, k & 1), ((k >> 1) & 1) ^ 1, p.dbg, SITE_SCHED_EMPTY);
        sched_ring[k & 1] = item;
        mbar_arrive(bar(SCHED_FULL, k & 1));       // release: the store above is visible to the waiters
      };
      uint32_t ent = 0;                        // running K/V ring entry (never reset between items)
      uint32_t qpar = 0;                       // bit t: parity of the Q loads issued for tile t
      int k = 0;
      int cur = atomicAdd(p.sched_counter, 1);
      publish(0, cur);
      while (cur < total) {
        const int nxt = atomicAdd(p.sched_counter, 1);
        publish(k + 1, nxt);                   // consumers always know one item ahead
        const WorkItem w = decode_item2<CAUSAL>(cur, p);
        auto load_q = [&](int t) {
          if (w.nblk[t] > 0) {
            mbar_wait(bar(Q_EMPTY, t), ((qpar >> t) & 1u) ^ 1u, p.dbg, SITE_LOAD_QEMPTY);
            qpar ^= (1u << t);
            mbar_arrive_expect_tx(bar(Q_FULL, t), TILE);
#pragma unroll
            for (int sl = 0; sl < C::SLABS; ++sl)
              tma_load_4d(sQ_addr + t * TILE + sl * C::SLAB_BYTES, &tmQ, bar(Q_FULL, t), sl * 64, w.row0[t], w.hidx, w.bidx);
          }
        };
        auto load_kv = [&](int j, int kv) {
          const int slot = ent % NSTAGE;
          const uint32_t par = (ent / NSTAGE) & 1;
          mbar_wait(bar(KV_EMPTY, slot), par ^ 1, p.dbg, SITE_LOAD_EMPTY);
          TFA_TRACE_EV(2);
          mbar_arrive_expect_tx(bar(KV_FULL, slot), TILE);
          const CUtensorMap* tm = (kv == 0) ? &tmK : &tmV;
#pragma unroll
          for (int sl = 0; sl < C::SLABS; ++sl)
            tma_load_4d(sKV_addr + slot * TILE + sl * C::SLAB_BYTES, tm, bar(KV_FULL, slot), sl * 64, j * C::BN, w.hidx, w.bidx);
          ++ent;
        };
        // Order matters for liveness: Q0, K0, V0, then Q1 (whose buffer frees last), then the rest.
        load_q(0);
        load_kv(0, 0);
        load_kv(0, 1);
        load_q(1);
        for (int j = 1; j < w.nmax; ++j) {
          load_kv(j, 0);
          load_kv(j, 1);
        }
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =========================== UMMA issuer ===========================
    // The whole warp stays converged (descriptors live in uniform registers); one elected lane issues the
    // tcgen05.mma / tcgen05.commit instructions.  Barrier waits that can be satisfied early (K/V tiles)
    // are taken BEFORE the P waits, so that once P_t is ready its PV and the next S are issued back to back.
    setmaxnreg_dec<kRegsOtherPersistent>();
    {
      constexpr uint32_t FMT = IS_BF16 ? 1u : 0u;
      const uint32_t idescS = umma_idesc_f16(FMT, 128, 128, 0, 0);  // A,B K-major
      const uint32_t idescO = umma_idesc_f16(FMT, 128, D, 0, 1);    // B (=V) MN-major
      TFA_PTRACE_DECL(2, lane == 0)
      TFA_TRACE_EV(1);

      // descriptor low words: Q tiles (fixed), ring slot 0 as K-major operand (K) and as MN-major operand (V);
      // another slot / k-step is one add (TILE>>4 per slot)
      const uint32_t q_lo0 = umma_desc_lo(sQ_addr, 16), q_lo1 = umma_desc_lo(sQ_addr + TILE, 16);
      const uint32_t k_lo_base = umma_desc_lo(sKV_addr, 16);
      const uint32_t v_lo_base = umma_desc_lo(sKV_addr, C::SLAB_BYTES);   // LBO = next 64-column slab
      constexpr uint32_t SLOT_LO = TILE >> 4;

      // S_t = Q_t K^T, then commit -> s_full[t]; optionally release the K slot and the Q buffer
      // `opaque` stops the compiler from pre-computing (and then spilling) the 16 TMEM operand addresses and
      // descriptor words as loop invariants: each is ONE add at issue time.
      auto opaque = [](uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
      auto issue_S = [&](int t, int kslot, uint32_t release_kv, bool release_q) {
        const uint32_t q_lo = opaque((t == 0) ? q_lo0 : q_lo1);
        const uint32_t k_lo = opaque(k_lo_base) + kslot * SLOT_LO;
        const uint32_t d_tmem = opaque(tmem_base) + (t == 0 ? C::TM_S0 : C::TM_S1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = (k / 4) * (C::SLAB_BYTES >> 4) + (k % 4) * 2;   // 16-byte units
            umma_ss_lo(d_tmem, q_lo + off, k_lo + off, idescS, k > 0 ? 1u : 0u);
          }
          umma_commit(bar(S_FULL, t));  // also covers every earlier MMA (incl. PV_t of the previous KV tile)
          if (release_kv != 0u) umma_commit(release_kv);
          if (release_q) umma_commit(bar(Q_EMPTY, t));
        }
        __syncwarp();
      };
      // O_t += P_t V for k-steps [k0, k1): 16 kv rows per step = 2048 B (128 units); SBO = 8-row group
      auto issue_PV = [&](int t, int vslot, bool acc, int k0, int k1, uint32_t release_kv, uint32_t done_bar) {
        const uint32_t v_lo = opaque(v_lo_base) + vslot * SLOT_LO;
        const uint32_t tb = opaque(tmem_base);
        const uint32_t d_tmem = tb + (t == 0 ? C::TM_O0 : C::TM_O1);
        const uint32_t p_tmem = tb + (t == 0 ? C::TM_S0 : C::TM_S1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C::BN / 16; ++k) {
            if (k >= k0 && k < k1) umma_ts_lo(d_tmem, p_tmem + k * 8, v_lo + k * 128, idescO, (acc || k > 0) ? 1u : 0u);
          }
          if (release_kv != 0u) umma_commit(release_kv);
          if (done_bar != 0u) umma_commit(done_bar);
        }
        __syncwarp();
      };
      auto ent_slot = [&](uint32_t e) { return static_cast<int>(e % NSTAGE); };
      auto ent_par = [&](uint32_t e) { return (e / NSTAGE) & 1u; };

      uint32_t ent_base = 0;                 // ring entry of K_0 of the current item
      // per-tile 1-bit state packed in one register (arrays indexed by t would live in local memory):
      //   bit t: q_full parity, bit 2+t: p_half/p_full parity, bit 4+t: S_t(0) of the current item already issued
      uint32_t st = 0;
      int k = 0;
      int cur = sched_get(0);
      // the issuer only needs the KV-tile counts of an item (never its coordinates)
      auto item_counts = [&](int item, int& n0, int& n1) {
        const WorkItem x = decode_item2<CAUSAL>(item, p);
        n0 = x.nblk[0];
        n1 = x.nblk[1];
      };
      while (cur < total) {
        int n0, n1, nn0 = 0, nn1 = 0;
        item_counts(cur, n0, n1);
        const int nmax = max(n0, n1);
        const int nxt = sched_get(k + 1);
        const bool has_nxt = nxt < total;
        if (has_nxt) item_counts(nxt, nn0, nn1);
        const uint32_t ent_next = ent_base + 2u * static_cast<uint32_t>(nmax);

        // first S of tile t of an item whose K_0 sits at ring entry e0 (x_nt = that item's tile count for t,
        // x_n1 = its tile-1 count: tile 1, when active, is the last user of K_0)
        auto first_S = [&](int t, int x_nt, int x_n1, uint32_t e0, uint32_t site_q, uint32_t site_k) {
          mbar_wait(bar(Q_FULL, t), (st >> t) & 1u, p.dbg, site_q);
          st ^= (1u << t);
          mbar_wait(bar(KV_FULL, ent_slot(e0)), ent_par(e0), p.dbg, site_k);
          tc_fence_after();
          const bool last_k_user = (t == 1) || (x_n1 == 0);
          umma_issue_first_S<D, IS_BF16>(tmem_base + C::TM_S0 + t * (C::TM_S1 - C::TM_S0), q_lo0 + t * SLOT_LO,
                                         k_lo_base + ent_slot(e0) * SLOT_LO, bar(S_FULL, t),
                                         last_k_user ? bar(KV_EMPTY, ent_slot(e0)) : 0u,
                                         x_nt == 1 ? bar(Q_EMPTY, t) : 0u);
        };

        // prologue: whatever was not hoisted out of the previous item
        // (the t loops are deliberately NOT unrolled: one copy of the issue code keeps the issuer in registers)
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int nt = (t == 0) ? n0 : n1;
          if (nt > 0 && !((st >> (4 + t)) & 1u)) {
            first_S(t, nt, n1, ent_base, SITE_MMA_Q, SITE_MMA_K0);
            TFA_TRACE_EV(4);
          }
          st &= ~(1u << (4 + t));
        }

        // K/V readiness is confirmed one KV tile AHEAD (inside the item), in the shadow of the first PV of tile 1
        bool kv_confirmed = false;
        for (int j = 0; j < nmax; ++j) {
          const uint32_t ev = ent_base + 2u * j + 1u, ek = ent_base + 2u * j + 2u;
          const int vslot = ent_slot(ev), kslot = ent_slot(ek);
          if (!kv_confirmed) {
            mbar_wait(bar(KV_FULL, vslot), ent_par(ev), p.dbg, SITE_MMA_V);
            TFA_TRACE_EV(5);
            if (j + 1 < nmax) {
              mbar_wait(bar(KV_FULL, kslot), ent_par(ek), p.dbg, SITE_MMA_K);
              TFA_TRACE_EV(10);
            }
          }
          kv_confirmed = false;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int nt = (t == 0) ? n0 : n1;
            if (j >= nt) continue;
            const bool last_v_user = !(t == 0 && j < n1);
            const bool has_next = (j + 1 < nt);
            const uint32_t ppar = (st >> (2 + t)) & 1u;
            mbar_wait(bar(P_HALF, t), ppar, p.dbg, SITE_MMA_PH);          // keys 0..63
            TFA_TRACE_EV(6 + t);
            tc_fence_after();
            issue_PV(t, vslot, j > 0, 0, 4, 0u, 0u);
            if (t == 1 && j + 1 < nmax) {
              const uint32_t ev2 = ev + 2u, ek2 = ek + 2u;
              mbar_wait(bar(KV_FULL, ent_slot(ev2)), ent_par(ev2), p.dbg, SITE_MMA_V);
              if (j + 2 < nmax) mbar_wait(bar(KV_FULL, ent_slot(ek2)), ent_par(ek2), p.dbg, SITE_MMA_K);
              kv_confirmed = true;
              TFA_TRACE_EV(11);
            }
            mbar_wait(bar(P_3Q, t), ppar, p.dbg, SITE_MMA_P3);            // keys 64..95
            tc_fence_after();
            issue_PV(t, vslot, true, 4, 6, 0u, 0u);
            mbar_wait(bar(P_FULL, t), ppar, p.dbg, SITE_MMA_P);           // keys 96..127
            st ^= (1u << (2 + t));
            TFA_TRACE_EV(8 + t);
            tc_fence_after();
            issue_PV(t, vslot, true, 6, 8, last_v_user ? bar(KV_EMPTY, vslot) : 0u,
                     has_next ? 0u : bar(O_FULL, t));
            if (has_next) {
              const bool last_k_user = !(t == 0 && j + 1 < n1);
              issue_S(t, kslot, last_k_user ? bar(KV_EMPTY, kslot) : 0u, j + 2 == nt);
              TFA_TRACE_EV(12 + t);
            } else {
              const int nnt = (t == 0) ? nn0 : nn1;
              // tile t is done with this item: if the next item's Q_t and K_0 have ALREADY landed, keep the tensor
              // pipe busy with its first S while the other tile finishes and this tile's warpgroup runs its epilogue.
              // Never block here: the other tile's P may be waiting for us.
              if (has_nxt && nnt > 0 && mbar_try_wait(bar(Q_FULL, t), (st >> t) & 1u) &&
                  mbar_try_wait(bar(KV_FULL, ent_slot(ent_next)), ent_par(ent_next))) {
                first_S(t, nnt, nn1, ent_next, SITE_MMA_HOIST_Q, SITE_MMA_HOIST_K);
                st |= (1u << (4 + t));
                TFA_TRACE_EV(14 + t);
              }
            }
          }
        }
        ent_base = ent_next;
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ================= softmax / correction / epilogue warpgroup t =================
    setmaxnreg_inc<kRegsSoftmaxPersistent>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;                       // row inside the Q tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + (t == 0 ? C::TM_S0 : C::TM_S1);
    const uint32_t tO = tmem_base + lane_base + (t == 0 ? C::TM_O0 : C::TM_O1);
    const float c = p.scale_log2;
    uint8_t* stg = sStg + warp * C::STG_WARP_BYTES;          // this warp's private staging
    TFA_PTRACE_DECL(t, r == 0)
    TFA_TRACE_EV(1);

    uint32_t scnt = 0;     // S tiles consumed  -> s_full parity
    uint32_t ocnt = 0;     // items finished    -> o_full parity
    for (int k = 0;; ++k) {
      const int item = sched_get(k);
      if (item >= total) break;
      const WorkItem w = decode_item2<CAUSAL>(item, p);
      const int n = (t == 0) ? w.nblk[0] : w.nblk[1];
      if (n == 0) continue;
      const int trow0 = (t == 0) ? w.row0[0] : w.row0[1];
      const int row_g = trow0 + r;                            // global query row

      float m_ref = 0.f;   // reference max the exponentials are taken against (raw score units)
      float l = 0.f;       // running sum of exp2((s - m_ref) * c)

      for (int j = 0; j < n; ++j) {
        mbar_wait(bar(S_FULL, t), scnt & 1, p.dbg, SITE_SM_S);
        ++scnt;
        TFA_TRACE_EV(2);
        tc_fence_after();

        // ---- S row -> registers: four back-to-back 32-column TMEM loads, ONE wait, mask, four max chains ----
        uint32_t sr[128];
        const int col0 = j * C::BN;
        int lim = S - col0;                                  // valid keys in this tile
        if (CAUSAL) lim = min(lim, row_g - col0 + 1);        // keys after the query (diagonal tile only)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) tmem_ld_x32(tS + q4 * 32, &sr[q4 * 32]);
        tmem_wait_ld();
        if (lim < C::BN) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= lim) sr[i] = 0xff800000u;                 // -inf
        }
        float mxa = -INFINITY, mxb = -INFINITY;
        {
          float mxc = -INFINITY, mxd = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            mxa = fmax3(mxa, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
            mxb = fmax3(mxb, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
            mxc = fmax3(mxc, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
            mxd = fmax3(mxd, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
          }
          mxa = fmaxf(mxa, mxc);
          mxb = fmaxf(mxb, mxd);
        }
        const float mx = fmaxf(mxa, mxb);
        TFA_TRACE_EV(3);

        // ---- lazy rescale of l and O (only when the max moved by more than 2^8) ----
        if (j == 0) {
          m_ref = mx;      // always finite: key 0 is visible to every row
        } else {
          const bool need = (mx - m_ref) * c > kRescaleThresholdLog2;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? mx : m_ref;
            const float alpha = ex2_approx((m_ref - m_new) * c);   // == 1 when !need
            m_ref = m_new;
            l *= alpha;
            // PV_t(j-1) has completed (s_full covers it) and PV_t(j) waits for p_half: O_t is ours.
#pragma unroll
            for (int ch = 0; ch < D / 32; ++ch) {
              uint32_t o[32];
              tmem_ld_x32(tO + ch * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(tO + ch * 32, o);
            }
          }
        }

        TFA_TRACE_EV(4);
        // ---- P = exp2(s*c - m_ref*c); l += rowsum(P) (fp32, before rounding); pack to 16 bit ----
        // Two lanes per instruction (FFMA2/FADD2).  MUFU.EX2 (16/clk/SM) is co-critical with the tensor
        // pipe, so kEmuPairsPer8 of every 8 element pairs take the polynomial exp2 on the FMA/ALU pipes.
        const float2 c2 = make_float2(c, c);
        const float2 nm2 = make_float2(-m_ref * c, -m_ref * c);
        constexpr int kEmuPairsPer8 = kEmuPairsPer8For<D>;
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        // P in four quarters of 32 keys; hand-offs after quarter 1 (p_half), 2 (p_3q) and 3 (p_full)
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int pi = qt * 16 + i;
            const float2 x = ffma2(make_float2(__uint_as_float(sr[2 * pi]), __uint_as_float(sr[2 * pi + 1])), c2, nm2);
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
          tmem_st_x16(tS + qt * 16, pk);    // P aliases columns [0,64) of S
          if (qt >= 1) {
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qt == 1 ? bar(P_HALF, t) : (qt == 2 ? bar(P_3Q, t) : bar(P_FULL, t)));
            if (qt == 1) TFA_TRACE_EV(5);
          }
        }
        acc0 = fadd2(acc0, acc1);
        l += acc0.x + acc0.y;
        TFA_TRACE_EV(6);
      }

      // ---------------------------- epilogue ----------------------------
      mbar_wait(bar(O_FULL, t), ocnt & 1, p.dbg, SITE_EPI_O);
      ++ocnt;
      TFA_TRACE_EV(7);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const long long tile_off =
          static_cast<long long>(w.bidx) * p.o_stride_b + static_cast<long long>(w.hidx) * p.o_stride_h;

      if (p.lse != nullptr && row_g < S)
        p.lse[static_cast<long long>(w.bh) * S + row_g] = m_ref * p.scale + logf(l);

      if constexpr (OUT_F32) {
        float* orow = p.out_f32 + tile_off + static_cast<long long>(row_g) * p.o_stride_s;
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld_x32(tO + ch * 32, o);
          tmem_wait_ld();
          if (row_g < S) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 v4 = make_float4(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l,
                                      __uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
              *reinterpret_cast<float4*>(orow + ch * 32 + i) = v4;
            }
          }
        }
      } else {
        // registers -> this warp's swizzled 32 x 128 B staging -> coalesced 128-bit stores (4 full 128-byte
        // row segments per warp instruction).  64 output columns per pass.
        uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + tile_off * 2;
        const int wrow0 = trow0 + (warp & 3) * 32;          // first global row owned by this warp
#pragma unroll
        for (int half = 0; half < D / 64; ++half) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t o[32];
            tmem_ld_x32(tO + half * 64 + ch * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 v4;
              v4.x = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 0]) * inv_l, __uint_as_float(o[q * 8 + 1]) * inv_l);
              v4.y = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 2]) * inv_l, __uint_as_float(o[q * 8 + 3]) * inv_l);
              v4.z = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 4]) * inv_l, __uint_as_float(o[q * 8 + 5]) * inv_l);
              v4.w = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 6]) * inv_l, __uint_as_float(o[q * 8 + 7]) * inv_l);
              const int chunk = ch * 4 + q;                  // 16-byte chunk inside the 128-byte row
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((chunk ^ (lane & 7)) * 16)) = v4;
            }
          }
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), chunk = lane & 7;
            const uint4 v4 = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((chunk ^ (rr & 7)) * 16));
            const int rg = wrow0 + rr;
            if (rg < S)
              st_global_v4(obase + static_cast<long long>(rg) * p.o_stride_s * 2 + half * 128 + chunk * 16, v4);
          }
          __syncwarp();
        }
      }
      TFA_TRACE_EV(8);
    }
    tc_fence_before();
  } else {
    setmaxnreg_dec<kRegsOtherPersistent>();   // warps 10-11: idle, give their registers away
  }

  // ---- teardown ----
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TM_COLS);
  }
}

}  // namespace tfa
