// tcgen05 flash attention (8 heads x d_h = 32, fp32 in / out, fp32-class accuracy via the 3xTF32 split).
//   causal  decoder self-attention with mask rule M1 (utils/train_utils.py:82-130, evaluated arithmetically)
//   padded  attention over the 224 memory tokens with a key-padding mask (encoder self-attn, decoder cross-attn)
//
// One CTA = one (group, head, 128-query tile); key tiles of 64 stream through a 2-stage shared-memory ring.
//   warp 7      TMA: Q tile once, then raw K tiles (cp.async.bulk.tensor.2d, SWIZZLE_128B).  The raw fp32 tiles are
//               used directly as the "hi" operands (kind::tf32 ignores the 13 low mantissa bits).
//   warps 4-5   derive the lo = x - trunc13(x) tiles (Q once, K per tile) and stage V^T hi / lo from global memory
//   warp 6      MMA issuer.  S = Q K^T: A = Q (smem, K-major), B = K tile (smem, K-major), M=128 N=64, 4 k-steps x 3
//               split products into two TMEM accumulators (hi*hi and cross terms).  O_tile = P V: A = P read FROM TMEM
//               (the softmax warps overwrite S in place with P_hi / P_lo, FlashAttention-4 style, so P never touches
//               shared memory), B = V^T tile (K-major; the producer warps transpose V while splitting it - an MN-major
//               tf32 B operand needs the SWIZZLE_128B_BASE32B layout, which this kernel avoids), M=128 N=32,
//               8 k-steps x 3, fresh accumulators per tile (the tensor core adds with truncation; short chains only).
//   warps 0-3   softmax: thread = query row = TMEM lane.  Two sweeps over S (max, then exp2 / row sum / hi-lo split /
//               tcgen05.st of P); after the PV commit the tile's output is added to the register-resident running
//               output with a rounded fp32 FMA together with the online-softmax rescale.
// TMEM (256 columns per CTA, two CTAs per SM): [0,64) S_main -> P_hi, [64,128) S_cross -> P_lo, [128,160) O_main,
// [160,192) O_cross.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int AT_QT = 128, AT_KT = 64, AT_STAGES = 2, AT_THREADS = 256;
constexpr uint32_t AT_TMEM_COLS = 256;
constexpr uint32_t AT_S_MAIN = 0, AT_S_CROSS = 64, AT_O_MAIN = 128, AT_O_CROSS = 160;

struct alignas(1024) AtKV {
  float k_raw[AT_KT * DH];   // [64 keys][32 dims], K-major (TMA, SWIZZLE_128B)
  float k_lo[AT_KT * DH];
  float vt_hi[DH * AT_KT];   // V^T: [32 dims][64 keys], K-major: two 4 KB column blocks of 32 keys (written by warps 4-5)
  float vt_lo[DH * AT_KT];
};
struct AtSmem {
  float q_raw[AT_QT * DH];
  float q_lo[AT_QT * DH];
  AtKV kv[AT_STAGES];
  uint64_t q_full, q_lo_ready, kv_full[AT_STAGES], lo_ready[AT_STAGES], kv_empty[AT_STAGES], s_full, p_ready, o_full;
  uint32_t tmem_base;
  unsigned long long pad_mask[4];  // padded mode: bit c of word j = key 64 j + c is valid and inside Lk
};

__device__ __forceinline__ uint32_t at_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(at_u32(bar)), "r"(count));
}
__device__ __forceinline__ void at_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(at_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(at_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void at_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(at_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t at_desc(uint32_t saddr) {  // SWIZZLE_128B, 8-row groups of 1024 B (see gemm_tc.cu)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t at_idesc(int M, int N, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void at_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void at_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void at_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void at_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void at_lo_tile(float* lo, const float* raw, int n_float4, int tid, int nthreads) {
  for (int i = tid; i < n_float4; i += nthreads) {
    const float4 v = reinterpret_cast<const float4*>(raw)[i];
    float4 l;
    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}
__device__ __forceinline__ bool at_m1_allowed(int tq, int aq, int kq, int key) {
  const int tk = key / TOK_T;
  if (tk < tq) return true;
  if (tk > tq) return false;
  const int rem = key - tk * TOK_T;
  const int ak = rem / KT, kk = rem - ak * KT;
  return kk == 0 || (ak == aq && kk <= kq);
}

__device__ int g_attn_debug = 0;  // bring-up aid: 2 dump raw PV, 3 dump P_hi read back after PV, 4 record a timeline
__device__ long long g_attn_trace[8 * 64];  // mode 4: clock64() at 8 events x up to 64 key tiles of CTA (0, 0, 0)
#define AT_TRACE(slot) do { if (trace_on && j < 64) g_attn_trace[j * 8 + (slot)] = clock64(); } while (0)

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int q_col0, int k_col0,
               const float* __restrict__ Vbase, int ldkv, const uint8_t* __restrict__ key_pad, float* __restrict__ O,
               int ldo, int Lq, int Lk) {
  extern __shared__ unsigned char at_raw[];
  AtSmem& sm = *reinterpret_cast<AtSmem*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.z, h = blockIdx.y;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;  // heavy (late) causal tiles first
  const int r0 = qt * AT_QT;
  int kend = Lk;
  if (CAUSAL) kend = min(Lk, (min(r0 + AT_QT - 1, Lq - 1) / TOK_T + 1) * TOK_T);
  const int n_tiles = (kend + AT_KT - 1) / AT_KT;

  if (tid == 0) {
    at_mbar_init(&sm.q_full, 1); at_mbar_init(&sm.q_lo_ready, 2);
    for (int s = 0; s < AT_STAGES; ++s) { at_mbar_init(&sm.kv_full[s], 1); at_mbar_init(&sm.lo_ready[s], 2); at_mbar_init(&sm.kv_empty[s], 1); }
    at_mbar_init(&sm.s_full, 1); at_mbar_init(&sm.p_ready, 128); at_mbar_init(&sm.o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!CAUSAL && tid < 4) {
    unsigned long long mk = 0ull;
    for (int c = 0; c < 64; ++c) {
      const int key = tid * 64 + c;
      if (key < Lk && !key_pad[(size_t)g * Lk + key]) mk |= 1ull << c;
    }
    sm.pad_mask[tid] = mk;
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_u32(&sm.tmem_base)), "r"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  const bool trace_on = g_attn_debug == 4 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                        (warp == 0 || warp == 4 || warp == 6);

  if (warp < 4) {
    // ------------------------------------------------------------------ softmax + output (thread = query row)
    const int row = r0 + 32 * warp + lane;
    const bool row_ok = row < Lq;
    int tq = 0, aq = 0, kq = 0;
    if (CAUSAL) { tq = row / TOK_T; const int rem = row - tq * TOK_T; aq = rem / KT; kq = rem - aq * KT; }
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    const float scale = 0.17677669529663687f * 1.4426950408889634f;  // d_h^-0.5 * log2(e)
    float m = -INFINITY, l = 0.f, o[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * AT_KT;
      // visibility of the 64 keys of this tile for this row, as a bit mask (bit c <-> key k0 + c)
      unsigned long long okm;
      if (CAUSAL) {
        if (!row_ok) okm = 0ull;
        else if (k0 + AT_KT <= tq * TOK_T) okm = ~0ull;
        else {
          okm = 0ull;
#pragma unroll 4
          for (int c = 0; c < AT_KT; ++c) {
            const int key = k0 + c;
            const int tk = key / TOK_T;
            const int rem = key - tk * TOK_T;
            const int ak = rem / KT;
            const int kk = rem - ak * KT;
            const bool ok = (key < kend) & ((tk < tq) | ((tk == tq) & ((kk == 0) | ((ak == aq) & (kk <= kq)))));
            okm |= (unsigned long long)ok << c;
          }
        }
      } else {
        okm = sm.pad_mask[j];
      }
      const bool fast = __all_sync(0xffffffffu, okm == ~0ull);
      AT_TRACE(7);
      at_wait(&sm.s_full, j & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      AT_TRACE(0);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32], rx[32];
        at_ld32(lane_addr + AT_S_MAIN + 32 * c, r);
        at_ld32(lane_addr + AT_S_CROSS + 32 * c, rx);
        if (fast) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]) + __uint_as_float(rx[i]));
        } else {
          const unsigned bits = (unsigned)(okm >> (32 * c));
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = __uint_as_float(r[i]) + __uint_as_float(rx[i]);
            mx4[i & 3] = fmaxf(mx4[i & 3], ((bits >> i) & 1u) ? s : -INFINITY);
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale;  // scale > 0 commutes with max
      const float mn = fmaxf(m, mx);
      const float ref = mn == -INFINITY ? 0.f : mn;
      const float corr = exp2f(m - ref);
      AT_TRACE(1);
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32], rx[32];
        at_ld32(lane_addr + AT_S_MAIN + 32 * c, r);
        at_ld32(lane_addr + AT_S_CROSS + 32 * c, rx);
        const unsigned bits = fast ? 0xffffffffu : (unsigned)(okm >> (32 * c));
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = __uint_as_float(r[i]) + __uint_as_float(rx[i]);
          float p = exp2f(fmaf(s, scale, -ref));
          p = ((bits >> i) & 1u) ? p : 0.f;
          sum4[i & 3] += p;
          const uint32_t hi = __float_as_uint(p) & 0xFFFFE000u;
          r[i] = hi;
          rx[i] = __float_as_uint(p - __uint_as_float(hi));
        }
        at_st32(lane_addr + AT_S_MAIN + 32 * c, r);   // P_hi over S_main
        at_st32(lane_addr + AT_S_CROSS + 32 * c, rx); // P_lo over S_cross
      }
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      l = l * corr + sum;
      m = mn;
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      at_arrive(&sm.p_ready);
      AT_TRACE(2);
      at_wait(&sm.o_full, j & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      AT_TRACE(3);
      if (g_attn_debug == 3 && j == 0) {
        uint32_t r[32];
        at_ld32(lane_addr + AT_S_MAIN, r);
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = __uint_as_float(r[i]);
        l = 1.f;
        break;
      }
      {
        uint32_t r[32], rx[32];
        at_ld32(lane_addr + AT_O_MAIN, r);
        at_ld32(lane_addr + AT_O_CROSS, rx);
        if (g_attn_debug == 2 && j == 0) {
#pragma unroll
          for (int i = 0; i < DH; ++i) o[i] = __uint_as_float(r[i]);
          l = 1.f;
          break;
        }
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = fmaf(o[i], corr, __uint_as_float(r[i]) + __uint_as_float(rx[i]));
      }
      AT_TRACE(4);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (row_ok) {
      const float inv = 1.0f / l;
      float* dst = O + ((size_t)g * Lq + row) * ldo + h * DH;
#pragma unroll
      for (int i = 0; i < DH; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ lo producers (64 threads)
    const int pt = tid - 128;
    at_wait(&sm.q_full, 0);
    at_lo_tile(sm.q_lo, sm.q_raw, AT_QT * DH / 4, pt, 64);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) at_arrive(&sm.q_lo_ready);
    const float* vsrc = Vbase + (size_t)g * Lk * ldkv + h * DH + (pt & 7) * 4;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % AT_STAGES;
      float4 vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // V rows of this tile: key = j*64 + pt/8 + 8 i, dims 4*(pt&7) .. +3
        const int key = j * AT_KT + (pt >> 3) + 8 * i;
        vv[i] = key < Lk ? *reinterpret_cast<const float4*>(vsrc + (size_t)key * ldkv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      at_wait(&sm.kv_full[s], (j / AT_STAGES) & 1);
      at_lo_tile(sm.kv[s].k_lo, sm.kv[s].k_raw, AT_KT * DH / 4, pt, 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kl = (pt >> 3) + 8 * i;                 // key within the tile
        const int cb = kl >> 5, kk = kl & 31;
        const float x[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int d = (pt & 7) * 4 + e;
          const int off = cb * (DH * 32) + d * 32 + ((((kk >> 2) ^ (d & 7)) << 2) | (kk & 3));
          sm.kv[s].vt_hi[off] = x[e];
          sm.kv[s].vt_lo[off] = x[e] - __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_arrive(&sm.lo_ready[s]);
      AT_TRACE(5);
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t id_qk = at_idesc(AT_QT, AT_KT, false), id_pv = at_idesc(AT_QT, DH, false);
    at_wait(&sm.q_lo_ready, 0);
    const uint64_t dqh = at_desc(at_u32(sm.q_raw)), dql = at_desc(at_u32(sm.q_lo));
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % AT_STAGES;
      at_wait(&sm.lo_ready[s], (j / AT_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint64_t dkh = at_desc(at_u32(sm.kv[s].k_raw)), dkl = at_desc(at_u32(sm.kv[s].k_lo));
#pragma unroll
        for (int ks = 0; ks < DH / 8; ++ks) {
          const uint64_t o2 = (uint64_t)(2 * ks);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          at_mma_ss(tmem + AT_S_CROSS, dql + o2, dkh + o2, id_qk, acc);
          at_mma_ss(tmem + AT_S_CROSS, dqh + o2, dkl + o2, id_qk, 1u);
          at_mma_ss(tmem + AT_S_MAIN, dqh + o2, dkh + o2, id_qk, acc);
        }
        at_commit(&sm.s_full);
      }
      __syncwarp();
      at_wait(&sm.p_ready, j & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      AT_TRACE(6);
      if (lane == 0) {
        const uint64_t dvh = at_desc(at_u32(sm.kv[s].vt_hi)), dvl = at_desc(at_u32(sm.kv[s].vt_lo));
#pragma unroll
        for (int ks = 0; ks < AT_KT / 8; ++ks) {
          // V^T is [32 dims][64 keys] K-major: 8 keys = 32 bytes inside a 128-byte atom row, 32 keys per 4 KB column block
          const uint64_t ob = (uint64_t)((ks >> 2) * ((DH * 128) >> 4) + (ks & 3) * 2);
          const uint32_t acc = ks > 0 ? 1u : 0u;
          at_mma_ts(tmem + AT_O_CROSS, tmem + AT_S_CROSS + 8 * ks, dvh + ob, id_pv, acc);
          at_mma_ts(tmem + AT_O_CROSS, tmem + AT_S_MAIN + 8 * ks, dvl + ob, id_pv, 1u);
          at_mma_ts(tmem + AT_O_MAIN, tmem + AT_S_MAIN + 8 * ks, dvh + ob, id_pv, acc);
        }
        at_commit(&sm.kv_empty[s]);
        at_commit(&sm.o_full);
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------------------------------------------ TMA issuer
    if (lane == 0) {
      at_expect_tx(&sm.q_full, AT_QT * DH * 4);
      at_tma_2d(sm.q_raw, &tmQ, q_col0 + h * DH, g * Lq + r0, &sm.q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % AT_STAGES;
        at_wait(&sm.kv_empty[s], ((j / AT_STAGES) & 1) ^ 1);
        at_expect_tx(&sm.kv_full[s], AT_KT * DH * 4);
        at_tma_2d(sm.kv[s].k_raw, &tmKV, k_col0 + h * DH, g * Lk + j * AT_KT, &sm.kv_full[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 6) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}

void set_attn_debug(int v) { cudaMemcpyToSymbol(g_attn_debug, &v, sizeof(int)); }
void read_attn_trace(long long* out) { cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * 8 * 64); }

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int at_make_map(AtEncodeFn enc, CUtensorMap* map, const float* base, long long rows, int cols, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)DH, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-5, "attn_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// Q rows: [G*Lq, q_cols] at Qbase (ldq), head h at columns q_col0 + 32h; K / V rows: [G*Lk, kv_cols] at KVbase (ldkv).
int launch_attn_tc(bool causal, const float* Qbase, int ldq, int q_cols, int q_col0, const float* KVbase, int ldkv,
                   int kv_cols, int k_col0, int v_col0, const uint8_t* key_pad, float* O, int ldo, int G, int Lq, int Lk,
                   cudaStream_t st) {
  static AtEncodeFn enc = nullptr;
  const int smem = (int)sizeof(AtSmem) + 1024;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return set_error(-5, "attn_tc: cuTensorMapEncodeTiled unavailable");
    e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "attn_tc smem attr: %s", cudaGetErrorString(e));
    enc = reinterpret_cast<AtEncodeFn>(fn);
  }
  if (!causal && Lk > 256) return set_error(-2, "attn_tc: padded mode supports at most 256 keys");
  CUtensorMap tmQ, tmKV;
  int rc;
  if ((rc = at_make_map(enc, &tmQ, Qbase, (long long)G * Lq, q_cols, ldq, AT_QT))) return rc;
  if ((rc = at_make_map(enc, &tmKV, KVbase, (long long)G * Lk, kv_cols, ldkv, AT_KT))) return rc;
  dim3 grid((Lq + AT_QT - 1) / AT_QT, NH, G);
  if (causal)
    attn_tc_kernel<true><<<grid, AT_THREADS, smem, st>>>(tmQ, tmKV, q_col0, k_col0, KVbase + v_col0, ldkv, key_pad, O, ldo, Lq, Lk);
  else
    attn_tc_kernel<false><<<grid, AT_THREADS, smem, st>>>(tmQ, tmKV, q_col0, k_col0, KVbase + v_col0, ldkv, key_pad, O, ldo, Lq, Lk);
  CS_CHECK_LAUNCH("attn_tc");
  return 0;
}

}  // namespace ctrlsim
