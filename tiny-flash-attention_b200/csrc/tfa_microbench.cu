// tfa_microbench.cu -- development micro-benchmarks (not part of the public C ABI header).
// They measure, on the actual B200, the constants the kernel design leans on:
//   * per-SMSP issue/throughput of MUFU.EX2, FFMA, FFMA2, FADD2, FMNMX3, F2FP and the polynomial exp2;
//   * SM-clock cycles per tcgen05.mma for the shapes the forward uses (SS N=128, TS N=128, TS N=64),
//     issued back to back from one thread while EVERY SM does the same (power-realistic).
// Results are clock64() deltas written per CTA; the host prints cycles per warp-instruction / per MMA.
#include "../../include/tfa_b200.h"
#include "ptx_sm100.cuh"

extern "C" void tfa_internal_count_launch(void);

namespace {
using namespace tfa;

// mode: 0 MUFU.EX2, 1 FFMA, 2 FFMA2, 3 FADD2, 4 FMNMX3, 5 F2FP(bf16x2), 6 ex2_poly2, 7 MUFU+FFMA2 mix (1:1 pairs)
template <int MODE>
__global__ void pipe_bench_kernel(float* sink, long long* cycles, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) {
        v[i] = ex2_approx(v[i]);
        v[i + 1] = ex2_approx(v[i + 1]);
      } else if (MODE == 1) {
        v[i] = fmaf(v[i], 1.0001f, 0.5f);
        v[i + 1] = fmaf(v[i + 1], 1.0001f, 0.5f);
      } else if (MODE == 2) {
        float2 r = ffma2(make_float2(v[i], v[i + 1]), make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
        v[i] = r.x; v[i + 1] = r.y;
      } else if (MODE == 3) {
        float2 r = fadd2(make_float2(v[i], v[i + 1]), make_float2(0.5f, 0.25f));
        v[i] = r.x; v[i + 1] = r.y;
      } else if (MODE == 4) {
        v[i] = fmax3(v[i], v[i + 1], 0.3f);
        v[i + 1] = fmax3(v[i + 1], v[i], 0.2f);
      } else if (MODE == 5) {
        v[i] = __uint_as_float(pack_16x2<true>(v[i], v[i + 1]));
        v[i + 1] = __uint_as_float(pack_16x2<true>(v[i + 1], v[i]));
      } else if (MODE == 6) {
        float2 r = ex2_poly2(make_float2(v[i], v[i + 1]));
        v[i] = r.x - 1.0f; v[i + 1] = r.y - 1.0f;
      } else {
        float2 r = ffma2(make_float2(v[i], v[i + 1]), make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));
        v[i] = ex2_approx(r.x) - 1.f; v[i + 1] = ex2_approx(r.y) - 1.f;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += v[i];
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// The exponential phase of the forward's softmax in isolation: exactly the per-row instruction mix of
// fa_fwd_sm100.cuh (FFMA2 scale, MUFU.EX2 / polynomial exp2 in the EMU-of-8 pattern, FADD2 row sum, F2FP pack) on a
// 128-element row held in registers, four quarters, the packed quarter stored to shared memory (stand-in for
// tcgen05.st).  `compute_warps` warps per CTA do that; `spin_warps` more warps sit in mbar_wait() on a barrier that
// completes only at the end, like the kernel's TMA/issuer/other-tile warps do.  Answers: how many cycles does ONE
// warp need per 128-element row, alone and with spinning neighbours on its SMSP?
template <int EMU>
__global__ void __launch_bounds__(384, 1) softmax_bench_kernel(const float* in, float* sink, long long* cycles, int iters,
                                                               int compute_warps) {
  __shared__ uint64_t done_bar;
  __shared__ __align__(16) uint32_t stage[12 * 32 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&done_bar, compute_warps); fence_mbar_init(); }
  __syncthreads();
  if (warp >= compute_warps) {                      // spinning neighbours
    mbar_wait(&done_bar, 0, nullptr, 0, 0);
    return;
  }
  uint32_t sr[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) sr[i] = __float_as_uint(in[(threadIdx.x * 128 + i) & 4095]);
  const float c = in[4096];
  const float2 c2 = make_float2(c, c);
  float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
  uint32_t nmb = __float_as_uint(in[4097]);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    asm volatile("mov.u32 %0, %0;" : "+r"(nmb));   // opaque per iteration: nothing below is loop invariant
    const float2 nm2 = make_float2(__uint_as_float(nmb), __uint_as_float(nmb));
#pragma unroll
    for (int qt = 0; qt < 4; ++qt) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pi = qt * 16 + i;
        const float2 x = ffma2(make_float2(__uint_as_float(sr[2 * pi]), __uint_as_float(sr[2 * pi + 1])), c2, nm2);
        float2 e;
        if (((pi * EMU) & 7) < EMU) {
          e = ex2_poly2(x);
        } else {
          e.x = ex2_approx(x.x);
          e.y = ex2_approx(x.y);
        }
        if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
        pk[i] = pack_16x2<true>(e.x, e.y);
      }
      const uint32_t dst = smem_u32(stage + (warp * 32 + lane) * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i)   // volatile: every quarter's store stays (they all hit the same 64 bytes)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i * 16), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                     "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
    }
  }
  const long long t1 = clock64();
  __syncwarp();
  if (lane == 0) mbar_arrive(&done_bar);
  const float a = acc0.x + acc0.y + acc1.x + acc1.y + __uint_as_float(stage[threadIdx.x]);
  if (a == 123.456f) sink[0] = a;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// UMMA throughput: one CTA per SM, one thread issues `n_mma` MMAs back to back (operands = garbage smem/TMEM,
// only timing matters), then commits and waits.  form: 0 = SS (A,B smem K-major), 1 = TS (A tmem, B smem MN-major).
// uniform: 0 = issue inside `if (lane == 0)` (what kernel v1 did), 1 = whole warp converged + elect.sync.
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(long long* cycles, int n_mma, int N, int form, int uniform) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(1, 128, N, 0, form == 0 ? 0 : 1);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    long long t0 = 0;
    if (uniform == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (form == 0) {
              const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
              umma_ss(tb, umma_smem_desc(a_addr + off, 16, 1024), umma_smem_desc(b_addr + off, 16, 1024), idesc, 1);
            } else {
              umma_ts(tb + 256, tb + k * 8, umma_smem_desc(b_addr + k * 2048, 16384, 1024), idesc, 1);
            }
          }
        }
        umma_commit(bar);
        mbar_wait(bar, 0, nullptr, 0, 0);
        cycles[blockIdx.x] = clock64() - t0;
      }
    } else {
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (form == 0) {
              const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
              umma_ss(tb, umma_smem_desc(a_addr + off, 16, 1024), umma_smem_desc(b_addr + off, 16, 1024), idesc, 1);
            } else {
              umma_ts(tb + 256, tb + k * 8, umma_smem_desc(b_addr + k * 2048, 16384, 1024), idesc, 1);
            }
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, 0, nullptr, 0, 0);
      if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// tcgen05.mma.cta_group::2 throughput (design input for a 2-CTA forward): a cluster of two CTAs on one TPC forms an
// M=256 x N MMA; each CTA holds its own 128 rows of A and HALF of B, so per SM the shared-memory operand traffic of
// the SS form drops from 8 KB to 6 KB per K=16 step -- the 1-CTA SS form is paced by exactly that traffic (102
// cycles against 64-73 for the TS form).  form 0 = SS, 1 = TS (A from TMEM).  Operands are garbage; timing only.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_bench_kernel(long long* cycles, int n_mma, int N, int form) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) {   // one warp of EACH CTA of the pair, same slot offset in both
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tb = *slot;
  if (warp == 1) {
    long long t0 = 0;
    if (cta_rank == 0) {           // the leader CTA issues for the pair
      const uint32_t idesc = umma_idesc_f16(1, 256, N, 0, form == 0 ? 0 : 1);
      const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
      t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (form == 0) {
              const uint32_t off = (k / 4) * 16384 + (k % 4) * 32;
              const uint64_t ad = umma_smem_desc(a_addr + off, 16, 1024), bd = umma_smem_desc(b_addr + off, 16, 1024);
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb), "l"(ad), "l"(bd),
                           "r"(idesc), "r"(1u) : "memory");
            } else {
              const uint64_t bd = umma_smem_desc(b_addr + k * 2048, 16384, 1024);
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tb + 256), "r"(tb + k * 8),
                           "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"(mask) : "memory");
      }
      __syncwarp();
    }
    mbar_wait(bar, 0, nullptr, 0, 0);          // both CTAs: the multicast commit arrives on each CTA's barrier
    if (cta_rank == 0 && lane == 0) cycles[blockIdx.x / 2] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

}  // namespace

extern "C" {

// which: 0..7 pipe modes (see above).  Launches `nblocks` CTAs of `nthreads`; writes one cycle count per CTA.
int tfa_microbench_pipe(int which, int nblocks, int nthreads, int iters, float* sink, long long* cycles, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (which) {
#define TFA_CASE(M) case M: pipe_bench_kernel<M><<<nblocks, nthreads, 0, s>>>(sink, cycles, iters); break;
    TFA_CASE(0) TFA_CASE(1) TFA_CASE(2) TFA_CASE(3) TFA_CASE(4) TFA_CASE(5) TFA_CASE(6) TFA_CASE(7)
#undef TFA_CASE
    default: return TFA_EINVAL_SHAPE;
  }
  tfa_internal_count_launch();
  return static_cast<int>(cudaGetLastError());
}

// emu: polynomial pairs of every 8 (0..4); compute_warps + spin_warps <= 12 warps per CTA (one CTA per SM).
int tfa_microbench_softmax(int emu, int nblocks, int compute_warps, int spin_warps, int iters, const float* in,
                           float* sink, long long* cycles, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nthreads = (compute_warps + spin_warps) * 32;
  if (compute_warps < 1 || nthreads > 384) return TFA_EINVAL_SHAPE;
  switch (emu) {
#define TFA_CASE(M) case M: softmax_bench_kernel<M><<<nblocks, nthreads, 0, s>>>(in, sink, cycles, iters, compute_warps); break;
    TFA_CASE(0) TFA_CASE(1) TFA_CASE(2) TFA_CASE(3) TFA_CASE(4)
#undef TFA_CASE
    default: return TFA_EINVAL_SHAPE;
  }
  tfa_internal_count_launch();
  return static_cast<int>(cudaGetLastError());
}

// nblocks must be even (clusters of 2); writes one cycle count per PAIR.
int tfa_microbench_umma2(int nblocks, int n_mma, int N, int form, long long* cycles, void* stream) {
  const int smem = 65536 + 1024 + 64;
  static bool attr = false;
  if (nblocks & 1) return TFA_EINVAL_SHAPE;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma2_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  umma2_bench_kernel<<<nblocks, 128, smem, static_cast<cudaStream_t>(stream)>>>(cycles, n_mma, N, form);
  tfa_internal_count_launch();
  return static_cast<int>(cudaGetLastError());
}

int tfa_microbench_umma(int nblocks, int n_mma, int N, int form, int uniform, long long* cycles, void* stream) {
  const int smem = 65536 + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  umma_bench_kernel<<<nblocks, 128, smem, static_cast<cudaStream_t>(stream)>>>(cycles, n_mma, N, form, uniform);
  tfa_internal_count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
