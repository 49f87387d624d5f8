// fa_fwd_sm100_persist.cuh -- the PERSISTENT fused attention-forward kernel for sm_100a (B200).
//
//   O = softmax(scale * Q K^T  [+ causal mask]) V ,  LSE = scale*max + ln(sum)
//
// Replaces the reference's CuTe/sm80 kernel (/root/reference/flash_attention_cutlass/csrc/flash_attention.cu:373-685).
// Same arithmetic and the same per-tile machine mapping as fa_fwd_sm100.cuh (TMA -> smem, tcgen05 SS/TS UMMA with
// accumulators in TMEM, one-thread-per-row softmax with lazy rescale, two ping-ponging 128-row Q tiles per work item);
// what this file adds is everything BETWEEN work items, which the one-CTA-per-item kernel pays ~6100 exposed cycles for
// (11 % of a causal S=4096 item, profiles/r01_trace_cfg3.txt):
//
//   * one CTA per SM pulls work items (same launch order as decode_work: 8-head chunks, heaviest causal pair first) from
//     an atomic counter; barriers, TMEM, tensor-map prefetch and register re-allocation happen once per CTA;
//   * the K/V ring runs ahead ACROSS items (the producer is never drained), and Q tiles of the next item are loaded as
//     soon as the last QK^T of the current item has consumed them;
//   * the first S = Q K^T of the next item is issued right behind the last PV of the current one ("hoist"), so the tensor
//     pipe works through the epilogue of tile t while the other tile is still in its main loop;
//   * the epilogue never touches the Q buffers: O/l -> 16 bit -> per-warp swizzled staging (32 rows x 128 B) -> ONE TMA
//     store (cp.async.bulk.tensor shared -> global) per warp per 64 columns, issued by one lane; the softmax warps move on
//     at once.  The fused multi-GPU exchange (tfa_fwd_multi) is the same TMA store repeated for up to 7 peer tensor maps
//     (NVLink writes leave from the TMA engine, not from st.global in the softmax warps).
//
// Warp roles (384 threads): warps 0-3 / 4-7 softmax + correction + epilogue of Q tile 0 / 1, warp 8 scheduler + TMA
// producer (one lane), warp 9 TMEM allocator + UMMA issuer (warp-uniform, one elected lane issues), warps 10-11 idle
// (register donors: setmaxnreg works on warpgroups).
#pragma once
#include "fa_fwd_sm100.cuh"
#include <type_traits>

namespace tfa {

struct alignas(64) OutMaps {
  CUtensorMap m[8];     // [0] = the caller's out tensor, [1..7] = peer copies (fused exchange); box = 64 x 32 elements
};

template <int D>
struct PCfg {
  static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
  static constexpr int BM = 128, BN = 128;
  static constexpr int SLABS = D / 64;
  static constexpr int SLAB_BYTES = 128 * 128;
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int NSTAGE = (D == 128) ? 4 : 8;                 // K/V ring depth (tiles), power of two
  static constexpr int NSTAGE_LOG2 = (D == 128) ? 2 : 3;
  static constexpr int STG_WARP_BYTES = 32 * 128;                   // epilogue staging: 32 rows x 128 B per softmax warp
  static constexpr int STG_BYTES = 8 * STG_WARP_BYTES;
  // barrier table (index of the first barrier of each kind)
  static constexpr uint32_t Q_FULL = 0, Q_EMPTY = 2, KV_FULL = 4, KV_EMPTY = KV_FULL + NSTAGE, S_FULL = KV_EMPTY + NSTAGE,
                            P_HALF = S_FULL + 2, P_3Q = P_HALF + 2, P_FULL = P_3Q + 2, O_FULL = P_FULL + 2,
                            SCHED_FULL = O_FULL + 2, SCHED_EMPTY = SCHED_FULL + 2, NUM_BARS = SCHED_EMPTY + 2;
  static constexpr int SMEM_BYTES =
      1024 /*align slack*/ + 2 * TILE_BYTES + NSTAGE * TILE_BYTES + STG_BYTES + NUM_BARS * 8 + 32;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static constexpr int TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 256 + D;   // P_t aliases columns [0,64) of S_t
  static constexpr int TM_COLS = 512;
  static constexpr int THREADS = 384;
};

enum : uint32_t {
  SITE_P_QEMPTY = 20, SITE_P_SCHED_EMPTY = 21, SITE_P_SCHED_FULL = 22, SITE_P_FIRST_Q = 23, SITE_P_FIRST_K = 24
};

// One work item: two adjacent 128-row Q tiles of one (batch, head) [and one KV split].
struct PItem {
  int bh, bidx, hidx, hkv, split, jb;
  int row0[2];
  int nblk[2];      // KV tiles [jb, jb + nblk[t]) for Q tile t (0 = tile inactive)
  int nmax;
};

// Tuning switches (A/B builds: build.py --variant NAME -DSWITCH=v).  Measured in r02 and reverted: alternating the long
// causal tile between the two tile resources from item to item (-8 %), two UMMA issuer warps (-28 %), a rolled issuer loop
// over the tile index (-3 %), no L2 prefetch of the next Q tiles (-3 %), deferring the P hand-offs behind the next quarter's
// exponentials (0 %, kept behind TFA_DEFER_HANDOFF).
#ifndef TFA_ISSUER_UNROLL_T
#define TFA_ISSUER_UNROLL_T 1   // the issuer's main loop is unrolled over the two tiles: compile-time tile index on the issue path
#endif
#ifndef TFA_P_HANDOFFS
#define TFA_P_HANDOFFS 2        // P_t is handed to the issuer after 64 and 128 keys.  A third hand-off after 96 keys (what the
                                // one-CTA-per-item kernel does) costs the issuer one more barrier wait per tile step: here it
                                // measured 0.5-1 % slower on cfg3 / cfg4 / non-causal S=4096 (profiles/r02_ab_spin_two_handoffs.txt)
#endif
#ifndef TFA_HOIST
#define TFA_HOIST 1             // 0: never issue the next item's first S early
#endif

template <bool CAUSAL>
__device__ __forceinline__ PItem decode_pitem(int item, const FwdParams& p) {
  PItem w;
  int pr;
  decode_work(item, p.npairs, p.nsplit, p.head_chunk, p.BH, w.bh, w.split, pr);
  w.bidx = w.bh / p.H;
  w.hidx = w.bh - w.bidx * p.H;
  w.hkv = w.hidx / p.kv_group;
  const int nkv_total = (p.Sk + 127) >> 7;
  w.jb = w.split * p.split_tiles;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    w.row0[t] = pr * 256 + t * 128;
    const bool active = w.row0[t] < p.S;
    const int nfull = active ? (CAUSAL ? min(nkv_total, ((w.row0[t] + 127 + p.causal_off) >> 7) + 1) : nkv_total) : 0;
    w.nblk[t] = max(0, min(nfull - w.jb, p.split_tiles));
  }
  w.nmax = max(w.nblk[0], w.nblk[1]);
  return w;
}

// TMA store of one staged (64 columns x 32 rows) box: shared -> global, completion tracked by the bulk async-group
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t smem_src, int x, int y, int z, int w) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}
// L2 prefetch of a (64 x 128) box: the later real load of the same box is an L2 hit
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* m, int x, int y, int z, int w) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}
// Q tiles are read exactly once, i.e. always from HBM (~2600 cycles after the buffer frees up, r01 trace), which is later
// than the hoist point of the next item's first S.  Prefetching the NEXT item's Q tiles into L2 one item ahead turns that
// into an L2 hit.  -DTFA_Q_PREFETCH=0 disables (A/B).
#ifndef TFA_Q_PREFETCH
#define TFA_Q_PREFETCH 1
#endif
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- UMMA issue helpers.  The issuer warp runs its loop body once per ~2000 cycles while eight softmax warps stream
// ~16 KB of unrolled math through the same instruction caches: an issuer unrolled over both tiles with the first-S code
// inlined at four sites (~10 KB per loop trip, measured r02) misses the instruction cache on nearly every fetch and
// stretches every "barrier satisfied -> MMA issued" gap (r02 timeline: issuer period 4080 cycles vs 3250).  So the
// issuer below is ROLLED over the tile index and has ONE S site in its loop.  (Out-of-line functions would be smaller
// still, but ptxas cannot allocate the 216-register softmax role once the kernel contains calls.)
// Called with warp-uniform arguments by the whole (converged) issuer warp. ----
template <int D, bool IS_BF16>
__device__ __forceinline__ void pumma_issue_S(uint32_t d_tmem, uint32_t q_lo, uint32_t k_lo, uint32_t bar_s, uint32_t bar_kv,
                                           uint32_t bar_q) {
  if (elect_one()) {
    const uint32_t idescS = umma_idesc_f16(IS_BF16 ? 1u : 0u, 128, 128, 0, 0);
#pragma unroll
    for (int k = 0; k < D / 16; ++k) {
      const uint32_t off = (k / 4) * ((128 * 128) >> 4) + (k % 4) * 2;   // 16-byte units: next slab every 4 k-steps
      umma_ss_lo(d_tmem, q_lo + off, k_lo + off, idescS, k > 0 ? 1u : 0u);
    }
    umma_commit(bar_s);
    if (bar_kv != 0u) umma_commit(bar_kv);
    if (bar_q != 0u) umma_commit(bar_q);
  }
  __syncwarp();
}
// O_t += P_t V for k-steps [k0, k1) (16 keys each); `acc` = accumulate onto O already at k0
template <int D, bool IS_BF16>
__device__ __forceinline__ void pumma_issue_PV(uint32_t d_tmem, uint32_t p_tmem, uint32_t v_lo, uint32_t acc, int k0, int k1,
                                            uint32_t bar_kv, uint32_t bar_done) {
  if (elect_one()) {
    const uint32_t idescO = umma_idesc_f16(IS_BF16 ? 1u : 0u, 128, D, 0, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k >= k0 && k < k1) {
        umma_ts_lo(d_tmem, p_tmem + k * 8, v_lo + k * 128, idescO, acc);
        acc = 1u;
      }
    }
    if (bar_kv != 0u) umma_commit(bar_kv);
    if (bar_done != 0u) umma_commit(bar_done);
  }
  __syncwarp();
}

template <int D, bool CAUSAL, bool IS_BF16, bool OUT_F32>
__global__ void __launch_bounds__(384, 1)
fa_fwd_sm100_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const __grid_constant__ OutMaps tmO,
                            const FwdParams p) {
  using C = PCfg<D>;
  constexpr int TILE = C::TILE_BYTES;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr uint32_t SLOT_LO = TILE >> 4;          // descriptor low-word step between ring slots / Q tiles

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sKV = smem + 2 * TILE;           // NSTAGE tiles
  uint8_t* sStg = sKV + NSTAGE * TILE;      // 8 x 4 KB epilogue staging (1024-aligned: SWIZZLE_128B boxes)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStg + C::STG_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto bar = [&](uint32_t which, uint32_t i) -> uint32_t { return bar_base + 8u * (which + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);
  volatile int* sched_ring = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [2]
  const uint32_t sQ_addr = smem_u32(sQ);
  const uint32_t sKV_addr = smem_u32(sKV);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = p.total_items;

  // ---- one-time setup ----
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    if (!OUT_F32) tma_prefetch_desc(&tmO.m[0]);
    for (uint32_t t = 0; t < 2; ++t) {
      mbar_init(bar(C::Q_FULL, t), 1);
      mbar_init(bar(C::Q_EMPTY, t), 1);
      mbar_init(bar(C::S_FULL, t), 1);
      mbar_init(bar(C::P_HALF, t), 4);      // one arrival per softmax warp
      mbar_init(bar(C::P_3Q, t), 4);
      mbar_init(bar(C::P_FULL, t), 4);
      mbar_init(bar(C::O_FULL, t), 1);
      mbar_init(bar(C::SCHED_FULL, t), 1);
      mbar_init(bar(C::SCHED_EMPTY, t), 9); // UMMA warp + 8 softmax warps
    }
    for (uint32_t i = 0; i < NSTAGE; ++i) {
      mbar_init(bar(C::KV_FULL, i), 1);
      mbar_init(bar(C::KV_EMPTY, i), 1);
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
  auto read_tmem_base = [&]() {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(tmem_slot)));
    return v;
  };

  // consumer side of the scheduler ring: the k-th item handed to this CTA lives in slot k&1 (>= total: no more work)
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

  if (warp == 8) {
    // ==================== scheduler + TMA producer ====================
    setmaxnreg_dec<kRegsOther>();
    if (lane == 0) {
      // next non-empty work item (a split-KV item wholly above the causal diagonal has no tiles: never handed out)
      // p.sched_counter = {next item, CTAs that ran out of work}: both are 0 at launch, and the last CTA to draw the
      // terminator puts them back to 0 for the next launch that is handed this pair (no memset in front of each launch)
      int fetched = 0;
      auto fetch = [&]() -> int {
        for (;;) {
          ++fetched;
          // CTA c starts with item c (no atomic, no ~1500-cycle round trip in front of the first load); the counter hands
          // out the items from gridDim.x on
          const int i = (fetched == 1) ? static_cast<int>(blockIdx.x) : atomicAdd(p.sched_counter, 1) + static_cast<int>(gridDim.x);
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
      // (compute-sanitizer racecheck reports this store against the consumers' loads: it does not model mbarrier
      // release/acquire for generic shared-memory accesses.  st.async, which would carry data and signal together, is
      // an illegal instruction outside a cluster launch on sm_100a -- measured.)
      auto publish = [&](int k, int item) {
        mbar_wait(bar(C::SCHED_EMPTY, k & 1), ((k >> 1) & 1) ^ 1, p.dbg, SITE_P_SCHED_EMPTY);
        sched_ring[k & 1] = item;
        mbar_arrive(bar(C::SCHED_FULL, k & 1));     // release: the store above is visible to the waiters (acquire in try_wait)
      };
      TFA_TRACE_DECL(3)
      TFA_TRACE_EV(1);
      uint32_t ent = 0;                        // running K/V ring entry (never reset between items)
      uint32_t qpar = 0;                       // bit t: parity of the next Q_EMPTY wait of tile t
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
#pragma unroll
            for (int sl = 0; sl < C::SLABS; ++sl)
              tma_load_4d(sQ_addr + t * TILE + sl * C::SLAB_BYTES, &tmQ, bar(C::Q_FULL, t), sl * 64, w.row0[t], w.hidx, w.bidx);
          }
        };
        auto load_kv = [&](int j, int kv) {
          const uint32_t slot = ent & (NSTAGE - 1);
          const uint32_t par = (ent >> C::NSTAGE_LOG2) & 1u;
          mbar_wait(bar(C::KV_EMPTY, slot), par ^ 1u, p.dbg, SITE_LOAD_EMPTY);
          TFA_TRACE_EV(2);
          mbar_arrive_expect_tx(bar(C::KV_FULL, slot), TILE);
          const CUtensorMap* tm = (kv == 0) ? &tmK : &tmV;
#pragma unroll
          for (int sl = 0; sl < C::SLABS; ++sl)
            tma_load_4d(sKV_addr + slot * TILE + sl * C::SLAB_BYTES, tm, bar(C::KV_FULL, slot), sl * 64,
                        (w.jb + j) * C::BN, w.hkv, w.bidx);
          ++ent;
        };
        // Q0, K0, V0 first (tile 0 can start), then Q1 (whose buffer frees last) -- all BEFORE the next item number is
        // drawn: the atomic's round trip must not sit in front of the loads the tensor pipe is waiting for
        load_q(0);
        load_kv(0, 0);
        load_kv(0, 1);
        load_q(1);
        for (int j = 1; j < w.nmax; ++j) {
          load_kv(j, 0);
          load_kv(j, 1);
        }
        // The next item is drawn as LATE as the pipeline allows: only now, with every load of this item issued (the ring
        // keeps the producer ~2 KV tiles ahead of the tensor pipe, which is enough for the next Q/K_0 to land before the
        // hoist point).  Drawing it an item ahead hands the last ~148 items of the queue to CTAs that are still busy
        // instead of to idle ones: with the heaviest items of the last head chunk among them that cost ~10 % of the
        // makespan on B4 H32 S4096 causal (r02, profiles/r02_persist_steps.md).
        const int nxt = fetch();
        publish(k + 1, nxt);
#if TFA_Q_PREFETCH
        if (nxt < total) {
          const PItem wn = decode_pitem<CAUSAL>(nxt, p);
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (wn.nblk[t] > 0) {
#pragma unroll
              for (int sl = 0; sl < C::SLABS; ++sl) tma_prefetch_l2_4d(&tmQ, sl * 64, wn.row0[t], wn.hidx, wn.bidx);
            }
        }
#endif
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // =========================== UMMA issuer ===========================
    setmaxnreg_dec<kRegsOther>();
    {
      const uint32_t tmem_base = read_tmem_base();
      const uint32_t q_lo0 = umma_desc_lo(sQ_addr, 16);
      const uint32_t k_lo_base = umma_desc_lo(sKV_addr, 16);
      const uint32_t v_lo_base = umma_desc_lo(sKV_addr, C::SLAB_BYTES);   // LBO = next 64-column slab
      auto ent_slot = [&](uint32_t e) { return e & (NSTAGE - 1); };
      auto ent_par = [&](uint32_t e) { return (e >> C::NSTAGE_LOG2) & 1u; };
      // S_t = Q_t K^T from ring slot `kslot`; commit -> s_full[t] (covers every earlier MMA incl. PV_t of the previous KV
      // tile), optionally release the K slot (last user) and the Q buffer (last S of tile t in this item)
      auto issue_S = [&](int t, uint32_t kslot, bool release_kv, bool release_q) {
        pumma_issue_S<D, IS_BF16>(tmem_base + t * (C::TM_S1 - C::TM_S0), q_lo0 + t * SLOT_LO, k_lo_base + kslot * SLOT_LO,
                                  bar(C::S_FULL, t), release_kv ? bar(C::KV_EMPTY, kslot) : 0u,
                                  release_q ? bar(C::Q_EMPTY, t) : 0u);
      };
      auto issue_PV = [&](int t, uint32_t vslot, bool acc, int k0, int k1, bool release_kv, bool done) {
        pumma_issue_PV<D, IS_BF16>(tmem_base + C::TM_O0 + t * (C::TM_O1 - C::TM_O0), tmem_base + t * (C::TM_S1 - C::TM_S0),
                                   v_lo_base + vslot * SLOT_LO, acc ? 1u : 0u, k0, k1,
                                   release_kv ? bar(C::KV_EMPTY, vslot) : 0u, done ? bar(C::O_FULL, t) : 0u);
      };

      TFA_TRACE_DECL(2)
#ifdef TFA_TRACE
      const bool trace_on_mma = trace_on && (lane == 0);
#define TFA_PTRACE_MMA(id) do { if (trace_on_mma) { TFA_TRACE_EV(id); } } while (0)
#else
#define TFA_PTRACE_MMA(id) do { } while (0)
#endif
      TFA_PTRACE_MMA(1);
      uint32_t ent_base = 0;       // ring entry of K_0 of the current item
      // per-tile 1-bit state in one register:  bit t: q_full parity | bit 2+t: P barriers' parity |
      //                                        bit 4+t: S_t(0) of the CURRENT item was already issued (hoisted)
      uint32_t st = 0;
      int k = 0;
      int cur = sched_get(0);
      while (cur < total) {
        int n0, n1;
        {
          const PItem x = decode_pitem<CAUSAL>(cur, p);
          n0 = x.nblk[0];
          n1 = x.nblk[1];
        }
        const int nmax = max(n0, n1);
        const uint32_t ent_next = ent_base + 2u * static_cast<uint32_t>(nmax);

        // First S = Q_t K_0^T of an item whose K_0 sits at ring entry e0.  K_0 is released by whichever tile issues its
        // first S LAST: `other_done` says the other tile's first S of that item was already issued (or it is inactive).
        auto first_S = [&](int t, int x_nt, bool other_done, uint32_t e0) {
          mbar_wait(bar(C::Q_FULL, t), (st >> t) & 1u, p.dbg, SITE_P_FIRST_Q);
          st ^= (1u << t);
          mbar_wait(bar(C::KV_FULL, ent_slot(e0)), ent_par(e0), p.dbg, SITE_P_FIRST_K);
          tc_fence_after();
          issue_S(t, ent_slot(e0), other_done, x_nt == 1);
        };

        // prologue: whatever was not hoisted out of the previous item (rolled: one copy of the code)
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int nt = (t == 0) ? n0 : n1;
          const int no = (t == 0) ? n1 : n0;
          if (nt > 0 && !((st >> (4 + t)) & 1u)) {
            const bool other_done = (no == 0) || ((st >> (4 + (t ^ 1))) & 1u) || (t == 1);
            first_S(t, nt, other_done, ent_base);
            st |= (1u << (4 + t));
            TFA_PTRACE_MMA(4);
          }
        }
        st &= ~(3u << 4);                      // the flags now describe the NEXT item: nothing hoisted yet
        // The next item is only needed for hoisting and is published late (see the producer): it is picked up by a
        // NON-blocking poll at the hoist probes and, at the latest, by a blocking read at the end of this item.
        int nxt = -1, nn0 = 0, nn1 = 0;                      // -1 = not known yet
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

        bool kv_confirmed = false;             // V_j and K_{j+1} of the upcoming iteration already waited for
        for (int j = 0; j < nmax; ++j) {
          const uint32_t ev = ent_base + 2u * j + 1u, ek = ent_base + 2u * j + 2u;
          const uint32_t vslot = ent_slot(ev), kslot = ent_slot(ek);
          if (!kv_confirmed) {
            mbar_wait(bar(C::KV_FULL, vslot), ent_par(ev), p.dbg, SITE_MMA_V);
            if (j + 1 < nmax) mbar_wait(bar(C::KV_FULL, kslot), ent_par(ek), p.dbg, SITE_MMA_K);
          }
          kv_confirmed = false;
#if TFA_ISSUER_UNROLL_T
#pragma unroll
#else
#pragma unroll 1
#endif
          for (int t = 0; t < 2; ++t) {
            const int nt = (t == 0) ? n0 : n1;
            const int no = (t == 0) ? n1 : n0;
            // tile 0 is served first: it is the last user of V_j / K_{j+1} only when tile 1 does not use them
            const bool active = (j < nt);
            const bool last_v_user = (t == 1) || (j >= no);
            const bool last_k_user = (t == 1) || (j + 1 >= no);
            const bool has_next = (j + 1 < nt);
            if (active) {
              const uint32_t ppar = (st >> (2 + t)) & 1u;
              mbar_wait(bar(C::P_HALF, t), ppar, p.dbg, SITE_MMA_PH);
              TFA_PTRACE_MMA(6 + t);
              tc_fence_after();
              issue_PV(t, vslot, j > 0, 0, 4, false, false);
              if (t == 1 && j + 1 < nmax) {
                // look-ahead: V_{j+1} and K_{j+2} were requested a full iteration ago; confirm them in the shadow of PV
                mbar_wait(bar(C::KV_FULL, ent_slot(ev + 2u)), ent_par(ev + 2u), p.dbg, SITE_MMA_V);
                if (j + 2 < nmax) mbar_wait(bar(C::KV_FULL, ent_slot(ek + 2u)), ent_par(ek + 2u), p.dbg, SITE_MMA_K);
                kv_confirmed = true;
              }
#if TFA_P_HANDOFFS == 3
              mbar_wait(bar(C::P_3Q, t), ppar, p.dbg, SITE_MMA_P3);
              tc_fence_after();
              issue_PV(t, vslot, true, 4, 6, false, false);
#endif
              mbar_wait(bar(C::P_FULL, t), ppar, p.dbg, SITE_MMA_P);
              TFA_PTRACE_MMA(8 + t);
              st ^= (1u << (2 + t));
              tc_fence_after();
              issue_PV(t, vslot, true, TFA_P_HANDOFFS == 3 ? 6 : 4, 8, last_v_user, !has_next);
            }
            // ONE S site: the next KV tile of this item, or -- tile t is done with this item (just now, or in an earlier
            // iteration while the other tile is still running) -- the first S of the NEXT item as soon as its Q_t and K_0
            // have landed.  The probe never blocks (test_wait): the other tile's P may be waiting for this warp.  The
            // tensor pipe runs the hoisted S while this tile's warpgroup does its epilogue and the other tile finishes.
            bool do_S = active && has_next, rel_kv = last_k_user, rel_q = (j + 2 == nt);
            uint32_t s_slot = kslot;
            if (!do_S) {
              // the probe result is voted: a per-thread predicate would make `st`, the slot and the release flags
              // divergent in the compiler's eyes and drag the whole issue path onto vector registers (see sched_get)
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
                st ^= (1u << t);                                   // Q_FULL parity consumed
                rel_kv = (nno == 0) || ((st >> (4 + (t ^ 1))) & 1u);   // K_0 is released by the LAST first-S of the item
                rel_q = (nnt == 1);
                s_slot = ent_slot(ent_next);
                st |= (1u << (4 + t));
                do_S = true;
                tc_fence_after();
                TFA_PTRACE_MMA(14 + t);
              }
            }
            if (do_S) {
              issue_S(t, s_slot, rel_kv, rel_q);
              TFA_PTRACE_MMA(12 + t);
            }
          }
        }
        poll_nxt(true);
        ent_base = ent_next;
        cur = nxt;
        ++k;
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ================= softmax / correction / epilogue warpgroup t =================
    setmaxnreg_inc<kRegsSoftmax>();
    const int t = warp >> 2;
    const int r = threadIdx.x & 127;                       // row inside the Q tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tmem_base = read_tmem_base();
    const uint32_t tS = tmem_base + lane_base + (t == 0 ? C::TM_S0 : C::TM_S1);
    const uint32_t tO = tmem_base + lane_base + (t == 0 ? C::TM_O0 : C::TM_O1);
    const uint32_t tP = tS;
    const float c = p.scale_log2;
    const int S = p.S, Sk = p.Sk;
    const uint32_t stg = smem_u32(sStg) + warp * C::STG_WARP_BYTES;   // this warp's private staging (1024-aligned)

    TFA_TRACE_DECL(t)
#ifdef TFA_TRACE
    const bool trace_on_sm = trace_on && (r == 0);
#define TFA_PTRACE_SM(id) do { if (trace_on_sm) { TFA_TRACE_EV(id); } } while (0)
#else
#define TFA_PTRACE_SM(id) do { } while (0)
#endif
    TFA_PTRACE_SM(1);
    uint32_t scnt = 0;     // S tiles consumed  -> s_full / P barriers' parity
    uint32_t ocnt = 0;     // items finished    -> o_full parity
    for (int k = 0;; ++k) {
      const int item = sched_get(k);
      if (item >= total) break;
      const PItem w = decode_pitem<CAUSAL>(item, p);
      const int n = (t == 0) ? w.nblk[0] : w.nblk[1];
      if (n == 0) continue;
      const int trow0 = (t == 0) ? w.row0[0] : w.row0[1];
      const int row_g = trow0 + r;                            // global query row
      const int jb = w.jb;

      float m_ref = 0.f;   // reference max the exponentials are taken against (raw score units)
      float l = 0.f;       // running sum of exp2((s - m_ref) * c)

      for (int j = 0; j < n; ++j) {
        mbar_wait(bar(C::S_FULL, t), scnt & 1u, p.dbg, SITE_SM_S);
        TFA_PTRACE_SM(2);
        tc_fence_after();

        // ---- S row -> registers: four back-to-back 32-column TMEM loads, ONE wait, mask, 4 independent max chains ----
        uint32_t sr[128];
        const int col0 = (jb + j) * C::BN;
        int lim = Sk - col0;                                 // valid keys in this tile
        if (CAUSAL) lim = min(lim, row_g + p.causal_off - col0 + 1);   // keys after the query (diagonal tiles only)
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) tmem_ld_x32(tS + q4 * 32, &sr[q4 * 32]);
        tmem_wait_ld();
        // (Tried in r02 and measured slower on B200: chunk-wise warp-uniform masking -- per-chunk max chains behind branches
        //  serialise (+210 cycles per tile), writing -inf into dead chunks under a branch makes ptxas handle the whole 128-
        //  register row conditionally (+1600).  The per-element select below only runs on diagonal / ragged tiles.)
        if (lim < C::BN) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= lim) sr[i] = 0xff800000u;                 // -inf
        }
        float mx;
        {
          float mxa = -INFINITY, mxb = -INFINITY, mxc = -INFINITY, mxd = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            mxa = fmax3(mxa, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
            mxb = fmax3(mxb, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
            mxc = fmax3(mxc, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
            mxd = fmax3(mxd, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
          }
          mx = fmaxf(fmaxf(mxa, mxc), fmaxf(mxb, mxd));
        }
        TFA_PTRACE_SM(3);
        if (j == 0) {
          m_ref = fmaxf(mx, -1.0e30f);                        // a fully masked row (split-KV) must not give -inf
        } else {
          // lazy rescale of l and O: only when the row max moved by more than 2^8 (warp-uniform branch, rare)
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
        TFA_PTRACE_SM(4);
        // ---- P = exp2(s*c - m_ref*c); l += rowsum(P) (fp32, before rounding); pack to 16 bit; TFA_P_HANDOFFS hand-offs ----
        constexpr int kEmuPairsPer8 = kEmuPairsPer8For<D>;
        const float2 c2 = make_float2(c, c);
        const float2 nm2 = make_float2(-m_ref * c, -m_ref * c);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#ifndef TFA_DEFER_HANDOFF
#define TFA_DEFER_HANDOFF 0   // 1: the drain of quarter q's TMEM stores (tcgen05.wait::st, ~100+ cycles in which the warp
                              // issues nothing) is taken AFTER quarter q+1's exponentials; a fake register dependency on
                              // the wait keeps ptxas from sinking the exponentials back below it (r01 tried without one)
#endif
        auto hand_off = [&](int which) {       // which: 1 = p_half, 2 = p_3q, 3 = p_full
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(which == 1 ? C::P_HALF : (which == 2 ? C::P_3Q : C::P_FULL), t));
        };
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
#if TFA_DEFER_HANDOFF
          static_assert(TFA_P_HANDOFFS == 3, "the deferred hand-off experiment assumes three hand-offs");
          if (qt >= 2) {       // quarters 0..qt-1 are stored; their drain was deferred to here, behind this quarter's math
            asm volatile("tcgen05.wait::st.sync.aligned; // after %0" ::"r"(pk[15]) : "memory");
            hand_off(qt - 1);
            TFA_PTRACE_SM(5);
          }
          tmem_st_x16(tP + qt * 16, pk);
          if (qt == 3) {
            tmem_wait_st();
            hand_off(3);
          }
#else
          tmem_st_x16(tP + qt * 16, pk);
          if (qt == 1 || qt == 3 || (qt == 2 && TFA_P_HANDOFFS == 3)) {   // after keys 0..63 (p_half), [64..95 (p_3q),] 96..127 (p_full)
            tmem_wait_st();
            hand_off(qt);
            if (qt < 3) TFA_PTRACE_SM(5);
          }
#endif
        }
        acc0 = fadd2(acc0, acc1);
        l += acc0.x + acc0.y;
        ++scnt;
        TFA_PTRACE_SM(6);
      }

      // ---------------------------- epilogue ----------------------------
      mbar_wait(bar(C::O_FULL, t), ocnt & 1u, p.dbg, SITE_EPI_O);
      TFA_PTRACE_SM(7);
      ++ocnt;
      tc_fence_after();
      // a row none of whose keys lies in this item's KV range (split-KV partial above the row's causal limit): the
      // polynomial exponentials of -inf are 2^-126, not 0 -- decide l = 0 analytically (O = 0, LSE = -inf, weight 0)
      if ((CAUSAL ? min(Sk, row_g + p.causal_off + 1) : Sk) <= jb * C::BN) l = 0.f;
      const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;

      if (p.lse != nullptr && row_g < S)
        p.lse[w.split * p.lse_part_stride + static_cast<long long>(w.bh) * p.lse_stride_bh + row_g] = m_ref * p.scale + logf(l);

      if constexpr (OUT_F32) {
        const long long tile_off =
            static_cast<long long>(w.bidx) * p.o_stride_b + static_cast<long long>(w.hidx) * p.o_stride_h;
        float* orow = p.out_f32 + w.split * p.part_stride + tile_off + static_cast<long long>(row_g) * p.o_stride_s;
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
        // registers -> this warp's swizzled 32 x 128 B staging -> one TMA store per destination; rows >= S are clipped
        // by the tensor map.  64 output columns per pass; the staging is re-used once the previous pass has been READ.
        const int wrow0 = trow0 + (warp & 3) * 32;          // first global row owned by this warp
#pragma unroll
        for (int half = 0; half < D / 64; ++half) {
          uint32_t pk[32];
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t o[32];
            tmem_ld_x32(tO + half * 64 + ch * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              pk[ch * 16 + i] = pack_16x2<IS_BF16>(__uint_as_float(o[2 * i]) * inv_l, __uint_as_float(o[2 * i + 1]) * inv_l);
          }
          if (lane == 0) bulk_wait_group_read0();           // previous pass / previous item: staging has been read
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q) {                     // 8 x 16-byte chunks of this row's 128 bytes
            const uint32_t addr = stg + lane * 128 + ((q ^ (lane & 7)) * 16);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[q * 4 + 0]), "r"(pk[q * 4 + 1]),
                         "r"(pk[q * 4 + 2]), "r"(pk[q * 4 + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmO.m[0], stg, half * 64, wrow0, w.hidx, w.bidx);
            for (int d = 0; d < p.n_extra_dst; ++d)          // peer copies: NVLink writes issued by the TMA engine
              tma_store_4d(&tmO.m[1 + d], stg, half * 64, wrow0, w.hidx, w.bidx);
            bulk_commit_group();
          }
        }
      }
      tc_fence_before();
      TFA_PTRACE_SM(8);
    }
    if (!OUT_F32 && lane == 0) bulk_wait_group0();          // all stores of this warp have completed before exit
  } else {
    setmaxnreg_dec<kRegsOther>();   // warps 10-11: idle, give their registers away
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(read_tmem_base(), C::TM_COLS);
  }
}

}  // namespace tfa
