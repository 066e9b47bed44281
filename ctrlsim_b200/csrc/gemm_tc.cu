// Tensor-core GEMM for the linear layers: C[M,N] = act(A[M,K] W[N,K]^T + bias + table[tidx]) on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), fp32 in / fp32 out with fp32-class accuracy.
//
// Accuracy: tf32 keeps 10 mantissa bits, which would put ~1e-3 relative error on every product and break sampled-bin
// parity with the fp32 reference.  Each operand is therefore split on the fly into hi = x with the 13 low mantissa
// bits cleared (exactly representable in tf32) and lo = x - hi (exact in fp32), and three MMAs are issued per k-step:
//   D += A_lo B_hi ;  D += A_hi B_lo ;  D += A_hi B_hi          (the dropped A_lo B_lo term is ~2^-22 relative)
// ("3xTF32").  Ceiling: 1/3 of the dense tf32 rate (~370 TFLOP/s nominal) instead of the 74 TFLOP/s FP32 FFMA peak.
// The tensor core adds into its fp32 accumulator with truncation, so the error grows linearly with the number of MMAs
// chained on one accumulator; the two small cross terms therefore go to a SECOND TMEM accumulator (their truncation
// is 2^-11 smaller) and are added once, with round-to-nearest, in the epilogue: K/8 chained adds instead of 3K/8.
//
// Structure (one CTA per 128 x 256 output tile, 288 threads):
//   warps 0-7  producers: global fp32 -> registers -> hi/lo split -> st.shared into the canonical K-major SWIZZLE_128B
//              layout (rows of 32 floats = 128 B, 8-row groups of 1024 B, 16-byte chunk index XOR (row & 7));
//              fence.proxy.async + mbarrier arrive on full[stage].  Afterwards the same warps run the epilogue:
//              tcgen05.ld 32x32b.x32 (warp w owns TMEM lanes 32*(w%4).., column half w/4) -> bias/table/ReLU -> global.
//   warp 8     allocates 256 TMEM columns, then one elected lane waits full[stage], issues 12 tcgen05.mma
//              (4 k-steps of 8 x 3 split products, M=128, N=256) and tcgen05.commit's to empty[stage]; the last commit
//              signals the epilogue.
// 2 stages x 96 KB (A_hi, A_lo 16 KB each; B_hi, B_lo 32 KB each) of dynamic shared memory.
#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32, TC_STAGES = 2;
constexpr int TC_PRODUCERS = 256, TC_THREADS = TC_PRODUCERS + 32;
constexpr uint32_t TC_TMEM_COLS = 512;  // columns [0,256): A_hi B_hi accumulator; [256,512): cross-term accumulator

struct alignas(1024) TcStage {
  float a_hi[TC_BM * TC_BK];
  float a_lo[TC_BM * TC_BK];
  float b_hi[TC_BN * TC_BK];
  float b_lo[TC_BN * TC_BK];
};
struct TcSmem {
  TcStage stage[TC_STAGES];
  uint64_t full[TC_STAGES];
  uint64_t empty[TC_STAGES];
  uint64_t accum_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(tc_smem_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // a lost arrival must not hang the GPU box
  }
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm100): start>>4 | LBO=1 |
// SBO = 1024 B (one 8-row group) | version 1 | layout_type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, K-major both,
// n_dim = N>>3 @17, m_dim = M>>4 @24.
__device__ __forceinline__ uint32_t tc_make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tc_split_store(float* hi, float* lo, int off, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

template <bool RELU>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const float* __restrict__ Aa, const float* __restrict__ W, const float* __restrict__ bias,
               const float* __restrict__ table, const int* __restrict__ tidx, const int* __restrict__ agather,
               float* __restrict__ C, int M, int N, int K, int lda, int ldw, int ldc, int ldt) {
  extern __shared__ unsigned char tc_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;
  const int nk = K / TC_BK;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&sm.full[s], TC_PRODUCERS); tc_mbar_init(&sm.empty[s], 1); }
    tc_mbar_init(&sm.accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sm.tmem_base)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  if (warp < 8) {
    // ------------------------------------------------------------------ producers
    const int c = tid & 7;                 // 16-byte chunk of the 128-byte k-slab
    const int rbase = tid >> 3;            // 0..31
    const int sc = (c ^ (rbase & 7)) << 2; // swizzled chunk, in floats (row & 7 == rbase & 7 for rows rbase + 32 i)
    const float* ap[4];
    const float* wp[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + rbase + 32 * i;
      ap[i] = m < M ? Aa + (size_t)(agather ? agather[m] : m) * lda + c * 4 : nullptr;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + rbase + 32 * i;
      wp[i] = n < N ? W + (size_t)n * ldw + c * 4 : nullptr;
    }
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TC_STAGES;
      const int k0 = kc * TC_BK;
      float4 va[4], vb[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) va[i] = ap[i] ? *reinterpret_cast<const float4*>(ap[i] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) vb[i] = wp[i] ? __ldg(reinterpret_cast<const float4*>(wp[i] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (kc >= TC_STAGES) tc_mbar_wait(&sm.empty[s], ((kc / TC_STAGES) - 1) & 1);
      TcStage& st = sm.stage[s];
#pragma unroll
      for (int i = 0; i < 4; ++i) tc_split_store(st.a_hi, st.a_lo, (rbase + 32 * i) * TC_BK + sc, va[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) tc_split_store(st.b_hi, st.b_lo, (rbase + 32 * i) * TC_BK + sc, vb[i]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_mbar_arrive(&sm.full[s]);
    }
    // ------------------------------------------------------------------ epilogue
    tc_mbar_wait(&sm.accum_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lg = warp & 3, half = warp >> 2;
    const int m = m0 + 32 * lg + lane;
    const float* trow = (table && m < M) ? table + (size_t)tidx[m] * ldt : nullptr;
    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int col0 = half * 128 + j * 32;
      uint32_t r[32], rx[32];
      tc_ld32(tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)col0, r);
      tc_ld32(tmem + ((uint32_t)(32 * lg) << 16) + (uint32_t)(TC_BN + col0), rx);
      if (m < M) {
        float* dst = C + (size_t)m * ldc + n0 + col0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = n0 + col0 + q + e;
            float x = __uint_as_float(r[q + e]) + __uint_as_float(rx[q + e]);
            if (n < N) {
              if (bias) x += __ldg(bias + n);
              if (trow) x += __ldg(trow + n);
            }
            v[e] = RELU ? fmaxf(x, 0.f) : x;
          }
          const int n = n0 + col0 + q;
          if (vec_ok && n + 3 < N) {
            *reinterpret_cast<float4*>(dst + q) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < N) dst[q + e] = v[e];
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 8)
    const uint32_t idesc = tc_make_idesc(TC_BM, TC_BN);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % TC_STAGES;
      tc_mbar_wait(&sm.full[s], (kc / TC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        TcStage& st = sm.stage[s];
        const uint64_t dah = tc_make_desc(tc_smem_u32(st.a_hi)), dal = tc_make_desc(tc_smem_u32(st.a_lo));
        const uint64_t dbh = tc_make_desc(tc_smem_u32(st.b_hi)), dbl = tc_make_desc(tc_smem_u32(st.b_lo));
#pragma unroll
        for (int ks = 0; ks < TC_BK / 8; ++ks) {   // UMMA_K = 8 tf32 = 32 bytes -> start address advances by 2 (x16 B)
          const uint64_t o = (uint64_t)(2 * ks);
          const uint32_t acc = (kc > 0 || ks > 0) ? 1u : 0u;
          tc_mma(tmem + TC_BN, dal + o, dbh + o, idesc, acc);
          tc_mma(tmem + TC_BN, dah + o, dbl + o, idesc, 1u);
          tc_mma(tmem, dah + o, dbh + o, idesc, acc);
        }
        tc_commit(&sm.empty[s]);
        if (kc == nk - 1) tc_commit(&sm.accum_full);
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
  }
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0) return 0;
  if (g.K % TC_BK != 0 || (g.lda & 3) || (g.ldw & 3))
    return set_error(-2, "gemm_tc: K=%d must be a multiple of %d and lda/ldw multiples of 4", g.K, TC_BK);
  static bool attr_set = false;
  const int smem = (int)sizeof(TcSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "gemm_tc smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((g.M + TC_BM - 1) / TC_BM, (g.N + TC_BN - 1) / TC_BN);
  if (g.relu)
    gemm_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(g.A, g.W, g.bias, g.table, g.tidx, g.agather, g.C, g.M, g.N, g.K,
                                                         g.lda, g.ldw, g.ldc, g.ldt);
  else
    gemm_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(g.A, g.W, g.bias, g.table, g.tidx, g.agather, g.C, g.M, g.N, g.K,
                                                          g.lda, g.ldw, g.ldc, g.ldt);
  CS_CHECK_LAUNCH("gemm_tc");
  return 0;
}

}  // namespace ctrlsim
