// fa_fwd_sm100.cuh -- the fused attention-forward kernel for sm_100a (B200).
//
//   O = softmax(scale * Q K^T  [+ causal mask]) V ,  LSE = scale*max + ln(sum)
//
// One hand-written kernel replaces the reference's CuTe/sm80 kernel
// (/root/reference/flash_attention_cutlass/csrc/flash_attention.cu:373-685).  The algorithm is
// the same FA-2 recurrence; the machine mapping is Blackwell-first:
//
//   * one CTA = one work item = TWO 128-row Q tiles of one (batch, head) that share every K/V tile
//     (reference: one 64-row tile per CTA, flash_attention.cu:695-699); items are launched in chunks of 8 heads,
//     heaviest causal pair first across the chunk (decode_work), so K/V stay L2 resident and the grid's tail is light;
//   * Q/K/V tiles are staged HBM -> shared memory by TMA (SWIZZLE_128B boxes of 64 x 128
//     elements) behind mbarriers, a 4 (D=128) or 8 (D=64) deep K/V ring (reference: cp.async, single
//     buffered, flash_attention.cu:521-525,556-565,581-590);
//   * S = Q K^T and O += P V run on tcgen05 tensor cores with fp32 accumulators in TMEM:
//     S is an SS-form UMMA (both operands K-major in smem), O is a TS-form UMMA whose A operand P
//     is read from TMEM (it aliases S) and whose B operand is the V tile consumed in place as an
//     MN-major operand -- no transpose, no ldmatrix.trans (reference: mma.sync m16n8k16 +
//     ldmatrix(.trans), flash_attention.cu:84-132, kernel_traits.h:26-39);
//   * softmax is one thread == one row (tcgen05.ld 32x32b): row max / row sum are thread-local,
//     zero shuffles (reference: quad shfl.bfly reductions, utils.h:22-91); exp is ex2.approx with
//     scale*log2(e) folded into one FFMA2 per element pair, a measured fraction of the exponentials runs as a
//     polynomial on the FMA pipe to unload MUFU; O is rescaled lazily, only when the running max moved by
//     more than 2^8 (reference: unconditional rescale every tile, flash_attention.cu:264-316);
//   * the two Q tiles ping-pong: while softmax warpgroup 0 works on S0, the tensor core runs
//     P1 V and the next Q1 K^T, and vice versa; P is handed over in three stages (keys 0-63, 64-95, 96-127) so PV
//     starts early and only two of its eight k-steps are left behind the last hand-off;
//   * causal: KV tiles above the diagonal are skipped, only the diagonal tile is masked
//     (reference: flash_attention.cu:536-540,576-578 with 64-wide tiles);
//   * epilogue: O/l -> 16-bit -> swizzled smem staging -> coalesced 128-bit st.global.v4
//     (reference: flash_attention.cu:608-663), optionally repeated for up to 7 peer GPUs' buffers (fused
//     all-gather, tfa_fwd_multi); LSE by the row-owner threads (:666-683).
//
// Warp roles (384 threads): warps 0-3 softmax/correction/epilogue for Q tile 0, warps 4-7 the same
// for Q tile 1, warp 8 TMA producer (one lane), warp 9 TMEM allocator + UMMA issuer (warp-uniform code, one
// elected lane issues), warps 10-11 idle (they exist so register re-allocation is warpgroup aligned).
// Measurements, machine constants and the variants that were tried and rejected: DESIGN.md section 4.
#pragma once
#include "ptx_sm100.cuh"

namespace tfa {

struct FwdParams {
  void* out;          // 16-bit output, element strides below
  // Fused exchange (multi-GPU): when n_extra_dst > 0 every 16-byte chunk of O is ALSO stored to these buffers
  // (same strides) -- peer GPUs' copies of the gathered output, mapped over NVLink.  Replaces the all-gather.
  void* extra_dst[7];
  int n_extra_dst;
  float* out_f32;     // fp32 output (validation build), same strides
  float* lse;         // (BH, S) fp32 or nullptr
  long long o_stride_b, o_stride_h, o_stride_s;  // elements
  int H;             // query heads
  int S;             // query rows
  int npairs;         // ceil(S / 256)
  // Generalised problem (reference: flash_attention_c/csrc/attn.cpp:121-124,182-183 for
  // Sq != Sk with a bottom-right aligned causal mask, csrc/archive_)/attn.cpp:61,375 for grouped K/V heads):
  int Sk;             // keys (== S for the reference's CuTe path)
  int causal_off;     // Sk - S >= 0: query row i sees keys j <= i + causal_off
  int kv_group;       // query heads per K/V head (1 = MHA)
  // split-KV: the grid carries nsplit CTAs per work item, each covering split_tiles KV tiles and writing a
  // normalised fp32 partial O plus its LSE; tfa_splitkv_combine merges them.  nsplit == 1: split_tiles = all.
  int nsplit;
  int split_tiles;
  long long part_stride;      // out_f32 elements between consecutive splits
  long long lse_stride_bh;    // lse elements between consecutive (b,h)
  long long lse_part_stride;  // lse elements between consecutive splits
  // launch order (decode_work below): heads are taken in chunks of `head_chunk`; total (batch*head) count `BH`
  int head_chunk;
  int BH;
  int total_items;    // npairs * B * H                         (persistent variant only)
  int* sched_counter; // zeroed before the launch; work counter  (persistent variant only)
  float scale;        // softmax_scale
  float scale_log2;   // softmax_scale * log2(e)
  DebugRecord* dbg;
  unsigned long long* trace;   // TFA_TRACE builds only: [4 roles][512] (clock64 << 8 | event id)
  int trace_block;
};

// ---- optional in-kernel timeline tracing (variant builds with -DTFA_TRACE; zero cost otherwise) ----
#ifdef TFA_TRACE
#define TFA_TRACE_DECL(role_expr)                                                         \
  const bool trace_on = (p.trace != nullptr) && (static_cast<int>(blockIdx.x) == p.trace_block); \
  unsigned long long* trace_ptr = p.trace + (role_expr) * 512;                            \
  int trace_n = 0;
#define TFA_TRACE_EV(id)                                                                  \
  do {                                                                                    \
    if (trace_on && trace_n < 511) {                                                      \
      trace_ptr[1 + trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(id); \
      trace_ptr[0] = trace_n;                                                             \
    }                                                                                     \
  } while (0)
#else
#define TFA_TRACE_DECL(role_expr)
#define TFA_TRACE_EV(id) do { } while (0)
#endif

template <int D>
struct FwdCfg {
  static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
  static constexpr int BM = 128;                    // rows per Q tile
  static constexpr int BN = 128;                    // keys per KV tile
  static constexpr int SLABS = D / 64;              // 128-byte-wide swizzle slabs per tile row
  static constexpr int SLAB_BYTES = 128 * 128;      // 128 rows x 128 B
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int NSTAGE = (D == 128) ? 4 : 8; // K/V ring depth (tiles)
  static constexpr int NUM_BARS = 2 + 2 * NSTAGE + 2 + 2 + 2 + 2 + 2 + 2;
#ifndef TFA_PAD_SMEM
#define TFA_PAD_SMEM 0        // experiment: extra (unused) dynamic shared memory, to isolate the effect of the carve-out
#endif
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + 2 * TILE_BYTES + NSTAGE * TILE_BYTES + NUM_BARS * 8 + 16 + TFA_PAD_SMEM;
  // TMEM columns (fp32): S0 | S1 | O0 | O1 ; P_t aliases the first 64 columns of S_t
  static constexpr int TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 256 + D;
  // D=64 leaves 128 TMEM columns unused.  -DTFA_D64_SEPARATE_P=1 (experiment, parity-tested) gives P_t its OWN 64
  // columns there instead of aliasing S_t, so S_t(j+1) need not wait for PV_t(j) to finish reading P_t(j): the issuer
  // launches it right after the first half of P_t(j) is handed over and the softmax warpgroup goes from tile j
  // straight into tile j+1.  Measured on B200: 1-2 % SLOWER (0.684 vs 0.668 ms, B4 H32 S4096 D64 non-causal) -- with
  // both warpgroups busy all the time they contend for the same sub-partitions' issue slots and XU, which is what
  // bounds D=64 (2 x ~1350 cycles of softmax work per KV tile per sub-partition), not the S->P->S chain.  Default off.
#ifndef TFA_D64_SEPARATE_P
#define TFA_D64_SEPARATE_P 0
#endif
  static constexpr bool P_SEPARATE = (D == 64) && (TFA_D64_SEPARATE_P != 0);
  static constexpr int TM_P0 = P_SEPARATE ? 384 : TM_S0, TM_P1 = P_SEPARATE ? 448 : TM_S1;
  static constexpr int TM_COLS = 512;
  static constexpr int THREADS = 384;
};

// watchdog call sites
enum : uint32_t {
  SITE_LOAD_EMPTY = 1, SITE_MMA_K0 = 2, SITE_MMA_Q = 3, SITE_MMA_V = 4, SITE_MMA_P = 5, SITE_MMA_K = 6,
  SITE_SM_S = 7, SITE_EPI_O = 8, SITE_MMA_PH = 9, SITE_MMA_P3 = 10, SITE_SM_PV = 11
};

// P hand-off point: the first kPSplitQ of 4 key-quarters go to the issuer early (PV k-steps [0, 2*kPSplitQ)).
#ifndef TFA_P_SPLITQ
#define TFA_P_SPLITQ 2
#endif
constexpr int kPSplitQ = TFA_P_SPLITQ;
static_assert(kPSplitQ >= 1 && kPSplitQ <= 3, "P split point must leave work on both sides");
// Hand-off stages: 3 (default) adds a hand-off after quarter 2, so the PV tail that sits between "P complete" and the
// next S_t on the per-tile dependency chain is 2 k-steps instead of 4 (measured +0.5..1 % on B200); 2 = p_half + p_full.
#ifndef TFA_P_STAGES
#define TFA_P_STAGES 3
#endif
static_assert(TFA_P_STAGES == 2 || (TFA_P_STAGES == 3 && TFA_P_SPLITQ == 2), "3-stage hand-off publishes after quarters 1, 2, 3");
constexpr float kRescaleThresholdLog2 = 8.0f;  // lazy rescale: tolerate P up to 2^8
// Of every 8 element pairs, this many use the polynomial exp2 instead of MUFU.  Measured on B200 (r01):
// with the staged P hand-off, D=128 is best at 2 (+4.6%; re-checked at the end of r01: 0 and 1 are 1-2.5% slower), D=64 at 3 (+20%).  -DTFA_EMU_PAIRS_PER_8=n overrides both (tuning).
#ifdef TFA_EMU_PAIRS_PER_8
template <int D> constexpr int kEmuPairsPer8For = TFA_EMU_PAIRS_PER_8;
#else
template <int D> constexpr int kEmuPairsPer8For = (D == 64) ? 3 : 2;
#endif
// Register re-allocation after the prologue.  setmaxnreg moves registers inside the CTA's OWN pool, which is
// what the launch allocated: 384 threads x 168 = 64512.  2 softmax warpgroups x 216 + 1 service warpgroup x 64
// = 496 x 128 = 63488 <= 64512 (224 would need 65536 and the second .inc could never be satisfied).
#ifndef TFA_REGS_SOFTMAX
#define TFA_REGS_SOFTMAX 216
#define TFA_REGS_OTHER 64
#endif
constexpr uint32_t kRegsSoftmax = TFA_REGS_SOFTMAX;
constexpr uint32_t kRegsOther = TFA_REGS_OTHER;
static_assert((2 * kRegsSoftmax + kRegsOther) * 128 <= 384 * 168, "setmaxnreg budget exceeds the CTA register pool");

// blockIdx -> work item.  CTAs are handed to SMs in blockIdx order, so this order IS the schedule.  (b,h)-major with
// the heaviest causal item first inside each head (the round-1 order) ends the grid with one head's 32-, 30-, ... tile
// items trickling onto idle SMs: +8 % over a balanced schedule on B4 H32 S4096 causal, +7 % on S=16384
// (scripts/tail_model.py reproduces the measured times).  Here heads are taken in chunks of `head_chunk` (8) and
// INSIDE a chunk the order is (split, pair)-major, head-minor: all heads' heaviest item first, then their second
// heaviest, ... -- the tail is made of the lightest items of the last chunk, while the heads running concurrently
// (and with them the K/V working set in L2, ~10 heads) stay what they were.  Host-callable so that a CPU test can
// check the mapping exhaustively (tests/test_abi.py).
__host__ __device__ __forceinline__ void decode_work(int block, int npairs, int nsplit, int head_chunk, int BH,
                                                     int& bh, int& split, int& pr) {
  const int per_bh = npairs * nsplit;                 // items per (batch, head)
  const int per_chunk = head_chunk * per_bh;
  const int chunk = block / per_chunk;
  const int r = block - chunk * per_chunk;
  const int g = min(head_chunk, BH - chunk * head_chunk);   // heads in this chunk (the last one may be short)
  const int sp = r / g;                               // (split, pair) index, heaviest pair first
  bh = chunk * head_chunk + (r - sp * g);
  split = sp / npairs;
  pr = npairs - 1 - (sp - split * npairs);
}

template <int D, bool CAUSAL, bool IS_BF16, bool OUT_F32>
__global__ void __launch_bounds__(384, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const FwdParams p) {
  using C = FwdCfg<D>;
  constexpr int TILE = C::TILE_BYTES;
  constexpr int NSTAGE = C::NSTAGE;

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment (the swizzle is a function of address bits 7..9)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sKV = smem + 2 * TILE;           // NSTAGE tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NSTAGE * TILE);
  uint64_t* q_full = bars;                  // [2]
  uint64_t* kv_full = bars + 2;             // [NSTAGE]
  uint64_t* kv_empty = kv_full + NSTAGE;    // [NSTAGE]
  uint64_t* s_full = kv_empty + NSTAGE;     // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_full = p_full + 2;            // [2]
  uint64_t* p_half = o_full + 2;            // [2]
  uint64_t* p_3q = p_half + 2;              // [2]  (TFA_P_STAGES == 3 only)
  uint64_t* pv_done = p_3q + 2;             // [2]  (P_SEPARATE only: PV_t(j) finished, P_t / O_t may be rewritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work decode (see decode_work) ----
  int bh, split, pr;
  decode_work(static_cast<int>(blockIdx.x), p.npairs, p.nsplit, p.head_chunk, p.BH, bh, split, pr);
  const int bidx = bh / p.H, hidx = bh % p.H;
  const int hkv = hidx / p.kv_group;                   // K/V head feeding this query head
  const int S = p.S, Sk = p.Sk;
  const int nkv_total = (Sk + C::BN - 1) / C::BN;
  const int jb = split * p.split_tiles;                // first KV tile of this CTA
  int row0[2], nblk[2];                                // nblk[t] = KV tiles [jb, jb + nblk[t]) for Q tile t
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    row0[t] = pr * 256 + t * 128;
    const bool active = row0[t] < S;
    // causal: the tile's last row sees keys <= row0 + 127 + causal_off
    const int nfull = active ? (CAUSAL ? min(nkv_total, (row0[t] + (C::BM - 1) + p.causal_off) / C::BN + 1) : nkv_total) : 0;
    nblk[t] = max(0, min(nfull - jb, p.split_tiles));
  }
  const int nmax = max(nblk[0], nblk[1]);
  if (nmax == 0) return;   // split-KV only: this split lies entirely above the causal diagonal (CTA-uniform)

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
      mbar_init(&p_full[t], 4);      // one arrival per softmax warp
      mbar_init(&p_half[t], 4);
      mbar_init(&p_3q[t], 4);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
    // First loads go out BEFORE the CTA-wide sync: they need neither TMEM nor the other warps, and the ~1.3 us of
    // first-touch TMA latency is the largest part of the per-item prologue.  Q tiles, then the first ring fill.
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (nblk[t] > 0) {
        mbar_arrive_expect_tx(&q_full[t], TILE);
#pragma unroll
        for (int sl = 0; sl < C::SLABS; ++sl)
          tma_load_4d(sQ + t * TILE + sl * C::SLAB_BYTES, &tmQ, &q_full[t], sl * 64, row0[t], hidx, bidx);
      }
    }
    for (int it = 0; it < NSTAGE && it < 2 * nmax; ++it) {      // slots are empty on the first pass
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
  // Each role re-reads the TMEM base address from shared memory (volatile) instead of carrying one value across the
  // role dispatch: a single long live range gets spilled to local memory as soon as any role is register-tight.
  auto read_tmem_base = [&]() {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(tmem_slot)));
    return v;
  };

  if (warp == 8) {
    // =========================== TMA producer ===========================
    setmaxnreg_dec<kRegsOther>();
    if (lane == 0) {
      TFA_TRACE_DECL(3)
      TFA_TRACE_EV(1);
      // ring entries [0, NSTAGE) were issued in the prologue, before the CTA-wide sync
      for (int it = NSTAGE; it < 2 * nmax; ++it) {
        const int slot = it % NSTAGE;
        const uint32_t par = (it / NSTAGE) & 1;
        mbar_wait(&kv_empty[slot], par ^ 1, p.dbg, SITE_LOAD_EMPTY, it);
        TFA_TRACE_EV(2);
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
    // The whole warp stays converged (descriptors live in uniform registers); one elected lane issues the
    // tcgen05.mma / tcgen05.commit instructions.  All barrier waits that can be satisfied early (K/V tiles)
    // are taken BEFORE the P waits, so that once P_t is ready its PV and the next S are issued back to back.
    setmaxnreg_dec<kRegsOther>();
    {
      const uint32_t tmem_base = read_tmem_base();
      constexpr uint32_t FMT = IS_BF16 ? 1u : 0u;
      const uint32_t idescS = umma_idesc_f16(FMT, 128, 128, 0, 0);  // A,B K-major
      const uint32_t idescO = umma_idesc_f16(FMT, 128, D, 0, 1);    // B (=V) MN-major
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sKV_addr = smem_u32(sKV);
      TFA_TRACE_DECL(2)
#ifdef TFA_TRACE
      const bool trace_on_mma = trace_on && (lane == 0);
#define TFA_TRACE_MMA(id) do { if (trace_on_mma) { TFA_TRACE_EV(id); } } while (0)
#else
#define TFA_TRACE_MMA(id) do { } while (0)
#endif
      TFA_TRACE_MMA(1);

      // Descriptor low words (start address >> 4 | LBO << 16): stepping along K or to another ring slot is ONE add.
      // `opaque` stops the compiler from pre-computing the 16 TMEM operand addresses / descriptor words as loop
      // invariants (it then spills them): every instruction between "barrier satisfied" and "MMA issued" is
      // exposed tensor-pipe idle time.
      auto opaque = [](uint32_t x) { uint32_t y; asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x)); return y; };
      const uint32_t q_lo0 = umma_desc_lo(sQ_addr, 16), q_lo1 = umma_desc_lo(sQ_addr + TILE, 16);

      // S_t = Q_t K^T, then commit -> s_full[t] (and optionally release the K slot)
      auto issue_S = [&](int t, uint32_t k_addr, uint64_t* release_bar) {
        const uint32_t q_lo = opaque((t == 0) ? q_lo0 : q_lo1);
        const uint32_t k_lo = umma_desc_lo(k_addr, 16);
        const uint32_t d_tmem = opaque(tmem_base) + (t == 0 ? C::TM_S0 : C::TM_S1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t off = (k / 4) * (C::SLAB_BYTES >> 4) + (k % 4) * 2;   // 16-byte units
            umma_ss_lo(d_tmem, q_lo + off, k_lo + off, idescS, k > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[t]);      // also covers every earlier MMA (incl. PV_t of the previous KV tile)
          if (release_bar != nullptr) umma_commit(release_bar);
        }
        __syncwarp();
      };
      // O_t += P_t V for k-steps [k0, k1): 16 kv rows per step = 2048 B (128 units); LBO = next 64-column slab
      auto issue_PV = [&](int t, uint32_t v_addr, bool acc, int k0, int k1, uint64_t* release_bar, uint64_t* done_bar,
                          uint64_t* done_bar2 = nullptr) {
        const uint32_t v_lo = umma_desc_lo(v_addr, C::SLAB_BYTES);
        const uint32_t tb = opaque(tmem_base);
        const uint32_t d_tmem = tb + (t == 0 ? C::TM_O0 : C::TM_O1);
        const uint32_t p_tmem = tb + (t == 0 ? C::TM_P0 : C::TM_P1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < C::BN / 16; ++k) {
            if (k >= k0 && k < k1) umma_ts_lo(d_tmem, p_tmem + k * 8, v_lo + k * 128, idescO, (acc || k > 0) ? 1u : 0u);
          }
          if (release_bar != nullptr) umma_commit(release_bar);
          if (done_bar != nullptr) umma_commit(done_bar);
          if (done_bar2 != nullptr) umma_commit(done_bar2);
        }
        __syncwarp();
      };

      // prologue: S_t(0) = Q_t K_0^T
      mbar_wait(&kv_full[0], 0, p.dbg, SITE_MMA_K0, 0);
      TFA_TRACE_MMA(2);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (nblk[t] > 0) {
          mbar_wait(&q_full[t], 0, p.dbg, SITE_MMA_Q, t);
          TFA_TRACE_MMA(3);
          tc_fence_after();
          const bool last_user = (t == 1) || (nblk[1] == 0);
          issue_S(t, sKV_addr, last_user ? &kv_empty[0] : nullptr);
          TFA_TRACE_MMA(4);
        }
      }

      // K/V readiness is confirmed one KV tile AHEAD, in the shadow of the first-half PV of tile 1 (the tensor pipe
      // is busy with those MMAs and the issuer is about to wait for P1 anyway).  A satisfied mbarrier wait still costs
      // ~200 cycles of issuer time; at the top of the loop that was pure tensor-pipe idle time in front of PV0.
      bool kv_confirmed = false;              // V_j and K_{j+1} of the upcoming iteration already waited for
      for (int j = 0; j < nmax; ++j) {
        const int v_it = 2 * j + 1, k_it = 2 * j + 2;
        const int vslot = v_it % NSTAGE, kslot = k_it % NSTAGE;
        const uint32_t vpar = (v_it / NSTAGE) & 1, kpar = (k_it / NSTAGE) & 1;
        if (!kv_confirmed) {
          mbar_wait(&kv_full[vslot], vpar, p.dbg, SITE_MMA_V, j);
          TFA_TRACE_MMA(5);
          if (j + 1 < nmax) {
            mbar_wait(&kv_full[kslot], kpar, p.dbg, SITE_MMA_K, j);
            TFA_TRACE_MMA(10);
          }
        }
        kv_confirmed = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j >= nblk[t]) continue;
          const bool last_v_user = !(t == 0 && j < nblk[1]);
          const bool has_next = (j + 1 < nblk[t]);
#define TFA_PV(acc_, k0_, k1_, rel_, done_, ...) \
  issue_PV(t, sKV_addr + vslot * TILE, acc_, k0_, k1_, rel_, done_, ##__VA_ARGS__)
#define TFA_S(rel_) issue_S(t, sKV_addr + kslot * TILE, rel_)
          // first half of P (keys 0..63 of the tile) is published early: start PV on it while the softmax
          // warpgroup is still exponentiating the second half
          mbar_wait(&p_half[t], j & 1, p.dbg, SITE_MMA_PH, j * 2 + t);
          TFA_TRACE_MMA(6 + t);
          tc_fence_after();
          TFA_PV(j > 0, 0, 2 * kPSplitQ, nullptr, nullptr);
          if (t == 1 && j + 1 < nmax) {
            // look-ahead: V_{j+1} and K_{j+2} were requested a full iteration ago
            const int v2 = 2 * j + 3, k2 = 2 * j + 4;
            mbar_wait(&kv_full[v2 % NSTAGE], (v2 / NSTAGE) & 1, p.dbg, SITE_MMA_V, j + 1);
            if (j + 2 < nmax) mbar_wait(&kv_full[k2 % NSTAGE], (k2 / NSTAGE) & 1, p.dbg, SITE_MMA_K, j + 1);
            kv_confirmed = true;
            TFA_TRACE_MMA(11);
          }
          const bool last_k_user = !(t == 0 && j + 1 < nblk[1]);
          if (C::P_SEPARATE && has_next) {
            // P_t has its own TMEM columns: S_t(j+1) may overwrite S_t now (the softmax warpgroup holds row j in
            // registers since before it handed over the first half of P)
            TFA_S(last_k_user ? &kv_empty[kslot] : nullptr);
            TFA_TRACE_MMA(12 + t);
          }
#if TFA_P_STAGES == 3
          mbar_wait(&p_3q[t], j & 1, p.dbg, SITE_MMA_P3, j * 2 + t);
          tc_fence_after();
          TFA_PV(true, 4, 6, nullptr, nullptr);
          constexpr int kTailK0 = 6;
#else
          constexpr int kTailK0 = 2 * kPSplitQ;
#endif
          mbar_wait(&p_full[t], j & 1, p.dbg, SITE_MMA_P, j * 2 + t);
          TFA_TRACE_MMA(8 + t);
          tc_fence_after();
          TFA_PV(true, kTailK0, 8, last_v_user ? &kv_empty[vslot] : nullptr, has_next ? nullptr : &o_full[t],
                 (C::P_SEPARATE && has_next) ? &pv_done[t] : nullptr);
          if (!C::P_SEPARATE && has_next) {
            TFA_S(last_k_user ? &kv_empty[kslot] : nullptr);
            TFA_TRACE_MMA(12 + t);
          }
#undef TFA_PV
#undef TFA_S
        }
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ================= softmax / correction / epilogue warpgroup t =================
    setmaxnreg_inc<kRegsSoftmax>();
    const int t = warp >> 2;
    const int n = (t == 0) ? nblk[0] : nblk[1];
    const int trow0 = (t == 0) ? row0[0] : row0[1];
    if (n > 0) {
      const int r = threadIdx.x & 127;                       // row inside the Q tile == TMEM lane
      const int row_g = trow0 + r;                            // global query row
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t tmem_base = read_tmem_base();
      const uint32_t tS = tmem_base + lane_base + (t == 0 ? C::TM_S0 : C::TM_S1);
      const uint32_t tO = tmem_base + lane_base + (t == 0 ? C::TM_O0 : C::TM_O1);
      const uint32_t tP = tmem_base + lane_base + (t == 0 ? C::TM_P0 : C::TM_P1);   // == tS unless P_SEPARATE
      const float c = p.scale_log2;

      TFA_TRACE_DECL(t)
#ifdef TFA_TRACE
      const bool trace_on_sm = trace_on && (r == 0);
#define TFA_TRACE_SM(id) do { if (trace_on_sm) { TFA_TRACE_EV(id); } } while (0)
#else
#define TFA_TRACE_SM(id) do { } while (0)
#endif
      TFA_TRACE_SM(1);
      float m_ref = 0.f;   // reference max the exponentials are taken against (raw score units)
      float l = 0.f;       // running sum of exp2((s - m_ref) * c)

      for (int j = 0; j < n; ++j) {
        mbar_wait(&s_full[t], j & 1, p.dbg, SITE_SM_S, j * 2 + t);
        TFA_TRACE_SM(2);
        tc_fence_after();

        // ---- S row -> registers: four back-to-back 32-column TMEM loads, ONE wait (measured 2-3 % faster than
        //      waiting per chunk to overlap the max with the next load), then mask + 4 independent max chains ----
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
        // lazy rescale of l and O: only when the row max moved by more than 2^8 (warp-uniform branch, rare)
        auto rescale_if_needed = [&](float mx) -> bool {
          const bool need = (mx - m_ref) * c > kRescaleThresholdLog2;
          if (!__any_sync(0xffffffffu, need)) return false;
          const float m_new = need ? mx : m_ref;
          const float alpha = ex2_approx((m_ref - m_new) * c);   // == 1 when !need
          m_ref = m_new;
          l *= alpha;
          // PV_t(j-1) has completed (s_full covers it) and PV_t(j) waits for p_full: O_t is ours.
#pragma unroll
          for (int ch = 0; ch < D / 32; ++ch) {
            uint32_t o[32];
            tmem_ld_x32(tO + ch * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO + ch * 32, o);
          }
          return true;
        };
        // ---- P = exp2(s*c - m_ref*c); l += rowsum(P) (fp32, before rounding); pack to 16 bit ----
        // Two lanes per instruction (FFMA2/FADD2).  MUFU.EX2 (16/clk/SM) would be co-critical with the tensor
        // pipe, so kEmuPairsPer8 of every 8 element pairs take the polynomial exp2 on the FMA/ALU pipes.
        // P is produced in four quarters of 32 keys, stored to TMEM as they finish (P aliases columns [0,64) of S).
        constexpr int kEmuPairsPer8 = kEmuPairsPer8For<D>;
        const float2 c2 = make_float2(c, c);
        auto p_compute = [&](int qt, float2 nm2, float2& acc0, float2& acc1, uint32_t (&pk)[16]) {
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
        };
        // hand-off of everything stored so far: drain the TMEM stores, fence, one arrival per warp
        auto hand_off = [&](uint64_t* bar) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar);
        };

        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        {
          TFA_TRACE_SM(3);
          if (j == 0) {
            m_ref = fmaxf(mx, -1.0e30f);                      // a fully masked row (split-KV) must not give -inf
          } else {
            if (C::P_SEPARATE) {
              // S_t(j) was issued BEFORE the tail of PV_t(j-1), so s_full no longer implies that PV_t(j-1) is done:
              // wait for its own commit before O_t is rescaled or P_t overwritten (normally long satisfied)
              mbar_wait(&pv_done[t], (j - 1) & 1, p.dbg, SITE_SM_PV, j * 2 + t);
              tc_fence_after();
            }
            rescale_if_needed(mx);
          }
          TFA_TRACE_SM(4);
          const float2 nm2 = make_float2(-m_ref * c, -m_ref * c);
#pragma unroll
          for (int qt = 0; qt < 4; ++qt) {
            uint32_t pk[16];
            p_compute(qt, nm2, acc0, acc1, pk);
            tmem_st_x16(tP + qt * 16, pk);
            // the first kPSplitQ quarters are handed over early (p_half) so the issuer can start PV on them while
            // the rest is still being exponentiated, [quarter 2 with p_3q,] the remainder with p_full
            if (qt == kPSplitQ - 1 || qt == 3 || (TFA_P_STAGES == 3 && qt == 2)) {
              hand_off(qt == 3 ? &p_full[t] : (qt == kPSplitQ - 1 ? &p_half[t] : &p_3q[t]));
              if (qt != 3) TFA_TRACE_SM(5);
            }
          }
        }
        acc0 = fadd2(acc0, acc1);
        l += acc0.x + acc0.y;
        TFA_TRACE_SM(6);
      }

      // ---------------------------- epilogue ----------------------------
      mbar_wait(&o_full[t], 0, p.dbg, SITE_EPI_O, t);
      TFA_TRACE_SM(7);
      tc_fence_after();
      // A row none of whose keys lies in this CTA's KV range (a split-KV partial wholly above the row's causal limit):
      // the MUFU exponentials of its -inf scores are exact zeros but the polynomial ones are 2^-126 (ex2_poly2 clamps),
      // so l would be tiny instead of 0.  Decide it analytically: O = 0, LSE = -inf (combine weight 0), also for
      // softmax_scale == 0 where the clamped exponent scale never rescales m_ref = -1e30 away.
      if ((CAUSAL ? min(Sk, row_g + p.causal_off + 1) : Sk) <= jb * C::BN) l = 0.f;
      const float inv_l = (l > 0.f) ? 1.0f / l : 0.f;
      const long long tile_off = static_cast<long long>(bidx) * p.o_stride_b + static_cast<long long>(hidx) * p.o_stride_h;

      if (p.lse != nullptr && row_g < S)
        p.lse[split * p.lse_part_stride + static_cast<long long>(bh) * p.lse_stride_bh + row_g] = m_ref * p.scale + logf(l);

      if constexpr (OUT_F32) {
        float* orow = p.out_f32 + split * p.part_stride + tile_off + static_cast<long long>(row_g) * p.o_stride_s;
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
        // registers -> swizzled staging (re-uses this tile's Q buffer: Q_t is dead once o_full fired)
        uint8_t* stg = sQ + t * TILE;
        constexpr int ROW_BYTES = D * 2;
        constexpr int CHUNKS = ROW_BYTES / 16;               // 16-byte chunks per row
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld_x32(tO + ch * 32, o);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v4;
            v4.x = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 0]) * inv_l, __uint_as_float(o[q * 8 + 1]) * inv_l);
            v4.y = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 2]) * inv_l, __uint_as_float(o[q * 8 + 3]) * inv_l);
            v4.z = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 4]) * inv_l, __uint_as_float(o[q * 8 + 5]) * inv_l);
            v4.w = pack_16x2<IS_BF16>(__uint_as_float(o[q * 8 + 6]) * inv_l, __uint_as_float(o[q * 8 + 7]) * inv_l);
            const int chunk = ch * 4 + q;
            const int phys = (chunk & ~7) | ((chunk ^ r) & 7);
            *reinterpret_cast<uint4*>(stg + r * ROW_BYTES + phys * 16) = v4;
          }
        }
        named_bar_sync(1 + t, 128);
        // coalesced 128-bit stores: consecutive threads write consecutive 16-byte chunks of a row
        uint8_t* obase = reinterpret_cast<uint8_t*>(p.out) + tile_off * 2;
#pragma unroll 4
        for (int idx = r; idx < 128 * CHUNKS; idx += 128) {
          const int rr = idx / CHUNKS, chunk = idx % CHUNKS;
          const int phys = (chunk & ~7) | ((chunk ^ rr) & 7);
          const uint4 v4 = *reinterpret_cast<const uint4*>(stg + rr * ROW_BYTES + phys * 16);
          const int rg = trow0 + rr;
          if (rg < S) {
            const long long off = tile_off * 2 + static_cast<long long>(rg) * p.o_stride_s * 2 + chunk * 16;
            st_global_v4(obase + (off - tile_off * 2), v4);
            for (int d = 0; d < p.n_extra_dst; ++d)          // peer copies: posted stores over NVLink
              st_global_v4(reinterpret_cast<uint8_t*>(p.extra_dst[d]) + off, v4);
          }
        }
      }
      TFA_TRACE_SM(8);
      tc_fence_before();
    }
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
