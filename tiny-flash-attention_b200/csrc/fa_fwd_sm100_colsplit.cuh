// fa_fwd_sm100_colsplit.cuh -- EXPERIMENTAL variant of the forward kernel (TFA_KERNEL=colsplit), written at the end
// of round 1 from that round's measurements and NOT yet run on hardware: it is never selected by default.
//
// Why (DESIGN.md sections 4 and 8): in fa_fwd_sm100_kernel every Q tile has ONE softmax warp per SM sub-partition;
// that warp needs ~1520 cycles for the exponentials of a 128-key row (issue/latency bound on its own), which makes the
// per-tile chain  S (816) -> softmax (2235) -> PV tail -> S  3250 cycles long against 2800 cycles of tensor work per KV
// tile pair.  Two warps sharing a sub-partition get through the same work at 1054 cycles per row
// (scripts/microbench.py softmax).  Here BOTH warpgroups work on the SAME Q tile: thread (wg, r) owns columns
// [64*wg, 64*wg + 64) of row r of S_t; tile 0 and tile 1 are processed one after the other by all eight warps.
//
// Differences from fa_fwd_sm100_kernel (everything else -- TMA producer, K/V ring, TMEM map, UMMA descriptors,
// work decode, split-KV parameters, epilogue staging -- is the same):
//   * row max: each thread reduces its 64 columns; on the first KV tile of an item the two halves are exchanged
//     through shared memory; on later tiles only the lazy-rescale PREDICATE is combined, with one `bar.red.or` over the
//     256 softmax threads (that barrier is also what orders "all of S_t is in registers" before the first P store,
//     because the P columns of warpgroup 1 alias S columns that warpgroup 0 loads); the rare rescale takes the
//     exchange path and each thread rescales its half of O_t's columns;
//   * row sum: per-thread partial sums, merged once per item in the epilogue;
//   * P hand-off: two stages per tile, after each thread's first 32 keys (PV k-steps {0,1} and {4,5}) and after the
//     rest ({2,3} and {6,7}); eight arrivals per barrier;
//   * epilogue: thread (wg, r) scales/packs D/2 columns of O_t; staging and the coalesced stores use all 256 threads.
#pragma once
#include "fa_fwd_sm100.cuh"

namespace tfa {

template <int D>
struct CsCfg : FwdCfg<D> {
  using Base = FwdCfg<D>;
  // q_full[2] kv_full[N] kv_empty[N] s_full[2] p_a[2] p_b[2] o_full[2]
  static constexpr int NUM_BARS = 2 + 2 * Base::NSTAGE + 2 + 2 + 2 + 2;
  static constexpr int XCH_FLOATS = 2 * 2 * 128;        // [tile][warpgroup][row]
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + 2 * Base::TILE_BYTES + Base::NSTAGE * Base::TILE_BYTES +
                                    XCH_FLOATS * 4 + NUM_BARS * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

enum : uint32_t { SITE_CS_PA = 20, SITE_CS_PB = 21 };

// barrier.red.or over `nthreads` threads of named barrier `id`: true if ANY participant passed true
__device__ __forceinline__ bool named_bar_red_or(uint32_t id, uint32_t nthreads, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t"
      ".reg .pred pin, pout;\n\t"
      "setp.ne.b32 pin, %1, 0;\n\t"
      "bar.red.or.pred pout, %2, %3, pin;\n\t"
      "selp.u32 %0, 1, 0, pout;\n\t"
      "}\n"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred)), "r"(id), "r"(nthreads)
      : "memory");
  return out != 0;
}

template <int D, bool CAUSAL, bool IS_BF16, bool OUT_F32>
__global__ void __launch_bounds__(384, 1)
fa_fwd_sm100_colsplit_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using C = CsCfg<D>;
  constexpr int TILE = C::TILE_BYTES;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr uint32_t SM_THREADS = 256;       // the softmax group: warps 0-7
  constexpr uint32_t BAR_SYNC = 1, BAR_RED = 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sKV = smem + 2 * TILE;           // NSTAGE tiles
  float* xch = reinterpret_cast<float*>(sKV + NSTAGE * TILE);    // [2][2][128] row-max / row-sum exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + C::XCH_FLOATS);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* kv_full = bars + 2;             // [NSTAGE]
  uint64_t* kv_empty = kv_full + NSTAGE;    // [NSTAGE]
  uint64_t* s_full = kv_empty + NSTAGE;     // [2]
  uint64_t* p_a = s_full + 2;               // [2]  first 32 keys of both column halves stored
  uint64_t* p_b = p_a + 2;                  // [2]  all of P_t stored
  uint64_t* o_full = p_b + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work decode (identical to fa_fwd_sm100_kernel) ----
  int bh, split, pr;
  decode_work(static_cast<int>(blockIdx.x), p.npairs, p.nsplit, p.head_chunk, p.BH, bh, split, pr);
  const int bidx = bh / p.H, hidx = bh % p.H;
  const int hkv = hidx / p.kv_group;
  const int S = p.S, Sk = p.Sk;
  const int nkv_total = (Sk + C::BN - 1) / C::BN;
  const int jb = split * p.split_tiles;
  int row0[2], nblk[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    row0[t] = pr * 256 + t * 128;
    const bool active = row0[t] < S;
    const int nfull = active ? (CAUSAL ? min(nkv_total, (row0[t] + (C::BM - 1) + p.causal_off) / C::BN + 1) : nkv_total) : 0;
    nblk[t] = max(0, min(nfull - jb, p.split_tiles));
  }
  const int nmax = max(nblk[0], nblk[1]);
  if (nmax == 0) return;

  // ---- one-time setup ----
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_a[t], 8);         // one arrival per softmax warp, all eight work on every tile
      mbar_init(&p_b[t], 8);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (nblk[t] > 0) {
        mbar_arrive_expect_tx(&q_full[t], TILE);
#pragma unroll
        for (int sl = 0; sl < C::SLABS; ++sl)
          tma_load_4d(sQ + t * TILE + sl * C::SLAB_BYTES, &tmQ, &q_full[t], sl * 64, row0[t], hidx, bidx);
      }
    }
    for (int it = 0; it < NSTAGE && it < 2 * nmax; ++it) {
      mbar_arrive_expect_tx(&kv_full[it], TILE);
      const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
#pragma unroll
      for (int sl = 0; sl < C::SLABS; ++sl)
        tma_load_4d(sKV + it * TILE + sl * C::SLAB_BYTES, tm, &kv_full[it], sl * 64, (jb + (it >> 1)) * C::BN, hkv, bidx);
    }
  }
  if (warp == 9) {
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

  if (warp == 8) {
    // =========================== TMA producer (unchanged) ===========================
    setmaxnreg_dec<kRegsOther>();
    if (lane == 0) {
      for (int it = NSTAGE; it < 2 * nmax; ++it) {
        const int slot = it % NSTAGE;
        const uint32_t par = (it / NSTAGE) & 1;
        mbar_wait(&kv_empty[slot], par ^ 1, p.dbg, SITE_LOAD_EMPTY, it);
        mbar_arrive_expect_tx(&kv_full[slot], TILE);
        const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
#pragma unroll
        for (int sl = 0; sl < C::SLABS; ++sl)
          tma_load_4d(sKV + slot * TILE + sl * C::SLAB_BYTES, tm, &kv_full[slot], sl * 64, (jb + (it >> 1)) * C::BN, hkv, bidx);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =========================== UMMA issuer ===========================
    setmaxnreg_dec<kRegsOther>();
    {
      const uint32_t tmem_base = read_tmem_base();
      constexpr uint32_t FMT = IS_BF16 ? 1u : 0u;
      const uint32_t idescS = umma_idesc_f16(FMT, 128, 128, 0, 0);
      const uint32_t idescO = umma_idesc_f16(FMT, 128, D, 0, 1);
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sKV_addr = smem_u32(sKV);
      auto opaque = [](uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
      const uint32_t q_lo0 = umma_desc_lo(sQ_addr, 16), q_lo1 = umma_desc_lo(sQ_addr + TILE, 16);

      auto issue_S = [&](int t, uint32_t k_addr, uint64_t* release_bar) {
        const uint32_t q_lo = opaque((t == 0) ? q_lo0 : q_lo1);
        const uint32_t k_lo = umma_desc_lo(k_addr, 16);
        const uint32_t d_tmem = opaque(tmem_base) + (t == 0 ? C::TM_S0 : C::TM_S1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = (k / 4) * (C::SLAB_BYTES >> 4) + (k % 4) * 2;
            umma_ss_lo(d_tmem, q_lo + off, k_lo + off, idescS, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[t]);      // also covers PV_t of the previous KV tile
          if (release_bar != nullptr) umma_commit(release_bar);
        }
        __syncwarp();
      };
      // O_t += P_t V for the k-steps in `kmask` (bit k = keys [16k, 16k+16)); `fresh`: the first MMA overwrites O_t
      auto issue_PV = [&](int t, uint32_t v_addr, bool fresh, uint32_t kmask, uint64_t* release_bar, uint64_t* done_bar) {
        const uint32_t v_lo = umma_desc_lo(v_addr, C::SLAB_BYTES);
        const uint32_t tb = opaque(tmem_base);
        const uint32_t d_tmem = tb + (t == 0 ? C::TM_O0 : C::TM_O1);
        const uint32_t p_tmem = tb + (t == 0 ? C::TM_S0 : C::TM_S1);
        if (elect_one()) {
          bool first = fresh;
#pragma unroll
          for (int k = 0; k < C::BN / 16; ++k) {
            if ((kmask >> k) & 1u) {
              umma_ts_lo(d_tmem, p_tmem + k * 8, v_lo + k * 128, idescO, first ? 0u : 1u);
              first = false;
            }
          }
          if (release_bar != nullptr) umma_commit(release_bar);
          if (done_bar != nullptr) umma_commit(done_bar);
        }
        __syncwarp();
      };

      // prologue: S_t(0) = Q_t K_0^T
      mbar_wait(&kv_full[0], 0, p.dbg, SITE_MMA_K0, 0);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nblk[t] > 0) {
          mbar_wait(&q_full[t], 0, p.dbg, SITE_MMA_Q, t);
          tc_fence_after();
          const bool last_user = (t == 1) || (nblk[1] == 0);
          issue_S(t, sKV_addr, last_user ? &kv_empty[0] : nullptr);
        }
      }

      bool kv_confirmed = false;
      for (int j = 0; j < nmax; ++j) {
        const int v_it = 2 * j + 1, k_it = 2 * j + 2;
        const int vslot = v_it % NSTAGE, kslot = k_it % NSTAGE;
        const uint32_t vpar = (v_it / NSTAGE) & 1, kpar = (k_it / NSTAGE) & 1;
        if (!kv_confirmed) {
          mbar_wait(&kv_full[vslot], vpar, p.dbg, SITE_MMA_V, j);
          if (j + 1 < nmax) mbar_wait(&kv_full[kslot], kpar, p.dbg, SITE_MMA_K, j);
        }
        kv_confirmed = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j >= nblk[t]) continue;
          const bool last_v_user = !(t == 0 && j < nblk[1]);
          const bool has_next = (j + 1 < nblk[t]);
          // keys 0..31 and 64..95 of the tile (the first half of each thread's columns)
          mbar_wait(&p_a[t], j & 1, p.dbg, SITE_CS_PA, j * 2 + t);
          tc_fence_after();
          issue_PV(t, sKV_addr + vslot * TILE, j == 0, 0x33u, nullptr, nullptr);
          if (t == 1 && j + 1 < nmax) {
            const int v2 = 2 * j + 3, k2 = 2 * j + 4;
            mbar_wait(&kv_full[v2 % NSTAGE], (v2 / NSTAGE) & 1, p.dbg, SITE_MMA_V, j + 1);
            if (j + 2 < nmax) mbar_wait(&kv_full[k2 % NSTAGE], (k2 / NSTAGE) & 1, p.dbg, SITE_MMA_K, j + 1);
            kv_confirmed = true;
          }
          mbar_wait(&p_b[t], j & 1, p.dbg, SITE_CS_PB, j * 2 + t);
          tc_fence_after();
          issue_PV(t, sKV_addr + vslot * TILE, false, 0xCCu, last_v_user ? &kv_empty[vslot] : nullptr,
                   has_next ? nullptr : &o_full[t]);
          if (has_next) {
            const bool last_k_user = !(t == 0 && j + 1 < nblk[1]);
            issue_S(t, sKV_addr + kslot * TILE, last_k_user ? &kv_empty[kslot] : nullptr);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ============ softmax / correction / epilogue: all eight warps on every Q tile, 64 columns per thread ============
    setmaxnreg_inc<kRegsSoftmax>();
    const int wg = warp >> 2;                                  // which half of the columns
    const int r = (warp & 3) * 32 + lane;                      // row inside the Q tile == TMEM lane
    const int tid = wg * 128 + r;                              // 0..255 inside the softmax group
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_base = read_tmem_base();
    const float c = p.scale_log2;
    constexpr int kEmuPairsPer8 = kEmuPairsPer8For<D>;
    constexpr int OH = D / 2;                                  // O columns per thread

    float m0 = 0.f, m1 = 0.f;     // reference max of tile 0 / 1 (identical in the two threads of a row)
    float l0 = 0.f, l1 = 0.f;     // this thread's PARTIAL row sums

    auto hand_off = [&](uint64_t* bar) {
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };

    // ---- one KV tile of Q tile t (t is a runtime value: ONE copy of this code) ----
    auto softmax_tile = [&](int t, int j) {
      const int row_g = (t ? row0[1] : row0[0]) + r;
      const uint32_t tS = tmem_base + lane_base + static_cast<uint32_t>(t * (C::TM_S1 - C::TM_S0) + wg * 64);
      const uint32_t tP = tmem_base + lane_base + static_cast<uint32_t>(t * (C::TM_S1 - C::TM_S0) + wg * 32);
      float m = t ? m1 : m0;
      float lsum = t ? l1 : l0;
      float* my_slot = xch + (t * 2 + wg) * 128 + r;
      const float* other_slot = xch + (t * 2 + (wg ^ 1)) * 128 + r;

      mbar_wait(&s_full[t], j & 1, p.dbg, SITE_SM_S, j * 2 + t);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld_x32(tS, &sr[0]);
      tmem_ld_x32(tS + 32, &sr[32]);
      tmem_wait_ld();
      const int col0 = (jb + j) * C::BN + wg * 64;             // first key of this thread's columns
      int lim = Sk - col0;
      if (CAUSAL) lim = min(lim, row_g + p.causal_off - col0 + 1);
      if (lim < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= lim) sr[i] = 0xff800000u;                   // -inf
      }
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        mxa = fmax3(mxa, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
        mxb = fmax3(mxb, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
      }
      const float mxh = fmaxf(mxa, mxb);

      // Every path below contains exactly one barrier over the 256 softmax threads between "S_t is in registers" and
      // the first P store: warpgroup 1 stores P into S columns [32,64), which warpgroup 0 has just loaded.
      if (j == 0) {
        *my_slot = mxh;
        named_bar_sync(BAR_SYNC, SM_THREADS);
        m = fmaxf(fmaxf(mxh, *other_slot), -1.0e30f);          // a fully masked row (split-KV) must not give -inf
      } else {
        const bool need_h = (mxh - m) * c > kRescaleThresholdLog2;
        if (named_bar_red_or(BAR_RED, SM_THREADS, need_h)) {     // rare: some row of this tile moved by > 2^8
          *my_slot = mxh;
          named_bar_sync(BAR_SYNC, SM_THREADS);
          const float mxr = fmaxf(mxh, *other_slot);
          const bool need = (mxr - m) * c > kRescaleThresholdLog2;
          const float m_new = need ? mxr : m;
          const float alpha = ex2_approx((m - m_new) * c);     // == 1 when !need
          m = m_new;
          lsum *= alpha;
          // PV_t(j-1) has completed (s_full covers it) and PV_t(j) waits for p_a: this half of O_t is ours
          const uint32_t tOh = tmem_base + lane_base + static_cast<uint32_t>(C::TM_O0 + t * D + wg * OH);
#pragma unroll
          for (int ch = 0; ch < OH / 32; ++ch) {
            uint32_t o[32];
            tmem_ld_x32(tOh + ch * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tOh + ch * 32, o);
          }
        }
      }

      const float2 c2 = make_float2(c, c);
      const float2 nm2 = make_float2(-m * c, -m * c);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int pi = h * 16 + i;
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
        tmem_st_x16(tP + h * 16, pk);                           // keys [64 wg + 32 h, +32) -> packed columns
        hand_off(h == 0 ? &p_a[t] : &p_b[t]);
      }
      acc0 = fadd2(acc0, acc1);
      lsum += acc0.x + acc0.y;
      if (t) { m1 = m; l1 = lsum; } else { m0 = m; l0 = lsum; }
    };

    // ---- epilogue of Q tile t ----
    auto epilogue_tile = [&](int t) {
      const int trow0 = t ? row0[1] : row0[0];
      const int row_g = trow0 + r;
      const float m = t ? m1 : m0;
      const float lsum = t ? l1 : l0;
      mbar_wait(&o_full[t], 0, p.dbg, SITE_EPI_O, t);
      tc_fence_after();
      xch[(t * 2 + wg) * 128 + r] = lsum;
      named_bar_sync(BAR_SYNC, SM_THREADS);
      const float ltot = xch[(t * 2) * 128 + r] + xch[(t * 2 + 1) * 128 + r];     // same order in both threads
      const float inv_l = (ltot > 0.f) ? 1.0f / ltot : 0.f;
      const long long tile_off = static_cast<long long>(bidx) * p.o_stride_b + static_cast<long long>(hidx) * p.o_stride_h;
      if (wg == 0 && p.lse != nullptr && row_g < S)
        p.lse[split * p.lse_part_stride + static_cast<long long>(bh) * p.lse_stride_bh + row_g] = m * p.scale + logf(ltot);

      const uint32_t tOh = tmem_base + lane_base + static_cast<uint32_t>(C::TM_O0 + t * D + wg * OH);
      if constexpr (OUT_F32) {
        float* orow = p.out_f32 + split * p.part_stride + tile_off + static_cast<long long>(row_g) * p.o_stride_s + wg * OH;
#pragma unroll
        for (int ch = 0; ch < OH / 32; ++ch) {
          uint32_t o[32];
          tmem_ld_x32(tOh + ch * 32, o);
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
        // registers -> swizzled staging in this tile's (dead) Q buffer -> coalesced 128-bit stores by all 256 threads
        uint8_t* stg = sQ + t * TILE;
        constexpr int ROW_BYTES = D * 2;
        constexpr int CHUNKS = ROW_BYTES / 16;
#pragma unroll
        for (int ch = 0; ch < OH / 32; ++ch) {
          uint32_t o[32];
          tmem_ld_x32(tOh + ch * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v4;
            v4.x = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 0]) * inv_l, __uint_as_float(o[q * 8 + 1]) * inv_l);
            v4.y = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 2]) * inv_l, __uint_as_float(o[q * 8 + 3]) * inv_l);
            v4.z = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 4]) * inv_l, __uint_as_float(o[q * 8 + 5]) * inv_l);
            v4.w = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 6]) * inv_l, __uint_as_float(o[q * 8 + 7]) * inv_l);
            const int chunk = wg * (CHUNKS / 2) + ch * 4 + q;
            const int phys = (chunk & ~7) | ((chunk ^ r) & 7);
            *reinterpret_cast<uint4*>(stg + r * ROW_BYTES + phys * 16) = v4;
          }
        }
        named_bar_sync(BAR_SYNC, SM_THREADS);
        uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + tile_off * 2;
#pragma unroll 4
        for (int idx = tid; idx < 128 * CHUNKS; idx += static_cast<int>(SM_THREADS)) {
          const int rr = idx / CHUNKS, chunk = idx % CHUNKS;
          const int phys = (chunk & ~7) | ((chunk ^ rr) & 7);
          const uint4 v4 = *reinterpret_cast<const uint4*>(stg + rr * ROW_BYTES + phys * 16);
          const int rg = trow0 + rr;
          if (rg < S) {
            const long long off = tile_off * 2 + static_cast<long long>(rg) * p.o_stride_s * 2 + chunk * 16;
            st_global_v4(obase + (off - tile_off * 2), v4);
            for (int d = 0; d < p.n_extra_dst; ++d)
              st_global_v4(reinterpret_cast<uint8_t*>(p.extra_dst[d]) + off, v4);
          }
        }
      }
    };

    // Program order is identical in all 256 threads (it depends on j and the tile counts only), so the named barriers
    // inside the two lambdas always meet.  A tile's epilogue runs in ITS slot of the first iteration it has no S in,
    // i.e. after the other tile's softmax of the previous slot: its last PV has long finished by then.
    const int n0 = nblk[0], n1 = nblk[1];
    for (int j = 0; j < nmax; ++j) {
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int nt = t ? n1 : n0;
        if (j < nt) softmax_tile(t, j);
        else if (j == nt && nt > 0) epilogue_tile(t);
      }
    }
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
      const int nt = t ? n1 : n0;
      if (nt == nmax) epilogue_tile(t);
    }
    tc_fence_before();
  } else {
    setmaxnreg_dec<kRegsOther>();   // warps 10-11: idle, give their registers away
  }

  // ---- teardown ----
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(read_tmem_base(), C::TM_COLS);
  }
}

}  // namespace tfa
