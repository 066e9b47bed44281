// attn_step on tcgen05: the 24 query rows of one (group, head) - the state rows of the last decoder layer in pass 1,
// the rtg rows in pass 2 - against every earlier token plus the 24 state tokens of the current step (rule M1 for those
// rows; the row's own new rtg key of pass 2 is folded in at the end).  Replaces the FP32 FFMA kernel of attention.cu.
//
// 24 queries are far too few for the M dimension of an MMA, so the roles are swapped: KEYS are M.
//   S^T [128 keys x 64]  = K_hi x [Q_hi ; Q_lo]^T  (N = 64)   and   S^T[:, 32:64] += K_lo x Q_hi^T  (N = 32)
//   O^T [128    x 64]    = [V_hi^T ; V_lo^T ; *] x [P_hi ; P_lo]^T  (N = 64), one MMA per 8 keys:
//        rows 0..31 = V_hi P_hi | V_hi P_lo,  rows 32..63 = V_lo P_hi | (V_lo P_lo),  rows 64..127 unused
// i.e. all split products of the 3xTF32 scheme come out of ONE instruction stream of N = 64 MMAs (24 per 128-key tile
// against 56 for the query-major kernel of attention_tc.cu).  Softmax runs along TMEM LANES (thread = key): the
// per-query reference exponent is an integer kept in shared memory and moved lazily (a block-wide vote; column maxima
// and column sums use a 31-shuffle transposing butterfly), P is written to shared memory transposed as the K-major
// B operand.  One CTA per (head, group), 256 threads in lock-step phases, two CTAs per SM.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int SK_KT = 128;
constexpr uint32_t SK_TMEM_COLS = 128;  // [0,64) S^T, [64,128) O^T

struct alignas(1024) SkSmem {
  float k_raw[SK_KT * DH];          // [128 keys][32 dims], K-major SWIZZLE_128B
  float k_lo[SK_KT * DH];
  float vt[4 * 2 * DH * 32];        // 4 key blocks x ([V_hi^T 32 dims][32 keys] | [V_lo^T 32 dims][32 keys])
  float p[4 * 2 * 32 * 32];         // 4 key blocks x ([P_hi 32 queries][32 keys] | [P_lo 32 queries][32 keys])
  float q[2 * 32 * DH];             // [Q_hi 32 rows ; Q_lo 32 rows] x 32 dims (rows >= 24 are zero)
  float ref[32], fac[32], wred[4][32];
  uint64_t mma_bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t sk_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sk_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row groups of 1024 B
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t sk_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void sk_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void sk_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sk_u32(bar)) : "memory");
}
__device__ __forceinline__ void sk_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(sk_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void sk_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ float sk_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// element (row r, column c) of a K-major SWIZZLE_128B tile with 32-float rows
__device__ __forceinline__ int sk_off(int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }

// v[i] of every lane -> lane L returns op over all lanes of v[L]  (31 shuffles; v is destroyed)
template <bool IS_MAX>
__device__ __forceinline__ float sk_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool up = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float keep = up ? v[i + w] : v[i], send = up ? v[i] : v[i + w];
      const float got = __shfl_xor_sync(0xffffffffu, send, w);
      v[i] = IS_MAX ? fmaxf(keep, got) : keep + got;
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(256, 2)
attn_step_tc_kernel(const float* __restrict__ KVbuf, int ld, int k_off, int v_off, int group_rows,
                    const float* __restrict__ qkv_rows, float* __restrict__ O, int ti, int own_row) {
  extern __shared__ unsigned char sk_raw[];
  SkSmem& sm = *reinterpret_cast<SkSmem*>((reinterpret_cast<uintptr_t>(sk_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.y, h = blockIdx.x;
  const float* base = KVbuf + (size_t)g * group_rows * ld + h * DH;
  const int n_hist = ti * TOK_T, n_keys = n_hist + A;
  const int n_tiles = (n_keys + SK_KT - 1) / SK_KT;
  const float scale = 0.17677669529663687f * 1.4426950408889634f;  // d_h^-0.5 * log2(e)

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sk_u32(&sm.mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) { sm.ref[tid] = -INFINITY; sm.fac[tid] = 1.f; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sk_u32(&sm.tmem_base)), "r"(SK_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // Q tile: thread -> (row r = tid / 8, 4 dims), scaled, split into hi (raw bits: the MMA truncates) and lo
    const int r = tid >> 3, c4 = (tid & 7) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < A) {
      v = *reinterpret_cast<const float4*>(qkv_rows + ((size_t)g * A + r) * (3 * H) + h * DH + c4);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    float4 lo;
    lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    sts128(sk_u32(sm.q) + 4u * (uint32_t)sk_off(r, c4), v);
    sts128(sk_u32(sm.q) + 4u * (uint32_t)(32 * 32 + sk_off(r, c4)), lo);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;

  // K / V rows of a tile: thread -> (key = tid / 2, 16 dims of K and of V)
  const int lkey = tid >> 1, half = tid & 1;
  float4 kq[4], vq[4];
  auto gload = [&](int tile) {
    const int key = tile * SK_KT + lkey;
    if (key < n_keys) {
      const int tok = key < n_hist ? key : n_hist + (key - n_hist) * KT;  // state token of agent (key - n_hist)
      const float* src = base + (size_t)tok * ld + 16 * half;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        kq[j] = *reinterpret_cast<const float4*>(src + k_off + 4 * j);
        vq[j] = *reinterpret_cast<const float4*>(src + v_off + 4 * j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { kq[j] = make_float4(0.f, 0.f, 0.f, 0.f); vq[j] = kq[j]; }
    }
  };
  gload(0);

  float o[32];      // warps 4, 5: running O^T row (dim = lane) over the 32 queries: V_hi part (warp 4), V_lo part (warp 5)
  float lpart = 0.f;  // warps 0-3: lane q holds this warp's partial row sum of query q
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = 0.f;
  uint32_t ph = 0;
  const uint32_t id64 = sk_idesc(128, 64), id32 = sk_idesc(128, 32);

  for (int tile = 0; tile < n_tiles; ++tile) {
    // ---- stage K (raw + lo) and V^T (hi + lo) of this tile -----------------------------------------------------------
    {
      const uint32_t kr = sk_u32(sm.k_raw), kl = sk_u32(sm.k_lo), vt = sk_u32(sm.vt);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = 4u * (uint32_t)sk_off(lkey, 16 * half + 4 * j);
        float4 lo;
        lo.x = kq[j].x - __uint_as_float(__float_as_uint(kq[j].x) & 0xFFFFE000u); lo.y = kq[j].y - __uint_as_float(__float_as_uint(kq[j].y) & 0xFFFFE000u);
        lo.z = kq[j].z - __uint_as_float(__float_as_uint(kq[j].z) & 0xFFFFE000u); lo.w = kq[j].w - __uint_as_float(__float_as_uint(kq[j].w) & 0xFFFFE000u);
        sts128(kr + off, kq[j]);
        sts128(kl + off, lo);
      }
      const float x[16] = {vq[0].x, vq[0].y, vq[0].z, vq[0].w, vq[1].x, vq[1].y, vq[1].z, vq[1].w,
                           vq[2].x, vq[2].y, vq[2].z, vq[2].w, vq[3].x, vq[3].y, vq[3].z, vq[3].w};
      const int kb = lkey >> 5, kk = lkey & 31;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        // the two halves of a key store dims that differ in bit 2 (rotation by 4): with the swizzle below a warp's 32
        // stores land in 32 distinct banks
        const int ee = half ? ((e + 4) & 15) : e;
        const float val = half ? x[(e + 4) & 15] : x[e];
        const int d = 16 * half + ee;
        const uint32_t off = 4u * (uint32_t)(kb * 2048 + sk_off(d, kk));
        sts32(vt + off, val);
        sts32(vt + off + 4096u, val - __uint_as_float(__float_as_uint(val) & 0xFFFFE000u));
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tile + 1 < n_tiles) gload(tile + 1);  // next tile's rows travel while this one is processed
    // ---- S^T = K Q^T ----------------------------------------------------------------------------------------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dkh = sk_desc(sk_u32(sm.k_raw)), dkl = sk_desc(sk_u32(sm.k_lo)), dq = sk_desc(sk_u32(sm.q));
#pragma unroll
      for (int ks = 0; ks < DH / 8; ++ks) {
        const uint64_t o2 = (uint64_t)(2 * ks);
        sk_mma(tmem, dkh + o2, dq + o2, id64, ks > 0 ? 1u : 0u);   // [K_hi Q_hi | K_hi Q_lo]
        sk_mma(tmem + 32, dkl + o2, dq + o2, id32, 1u);             // + K_lo Q_hi
      }
      sk_commit(&sm.mma_bar);
    }
    sk_wait(&sm.mma_bar, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- softmax along the keys (thread = key = TMEM lane) ------------------------------------------------------------
    float s[32];
    const bool smx = warp < 4;
    const bool valid = smx && (tile * SK_KT + tid) < n_keys;
    bool need = false;
    if (smx) {
      uint32_t r[32], rx[32];
      const uint32_t la = tmem + ((uint32_t)(32 * warp) << 16);
      sk_ld32(la, r);
      sk_ld32(la + 32, rx);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s[i] = valid ? __uint_as_float(r[i]) + __uint_as_float(rx[i]) : -INFINITY;  // already scaled through Q
        need |= s[i] > sm.ref[i] + 16.f;  // also true while ref is still -inf
      }
    }
    if (__syncthreads_or(need ? 1 : 0)) {
      // move the reference exponents to the (integer) running column maxima; exact power-of-two rescale of l and O^T
      if (smx) {
        float t[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = s[i];
        const float cm = sk_transpose_reduce<true>(t, lane);
        sm.wred[warp][lane] = cm;
      }
      __syncthreads();
      if (tid < 32) {
        const float nm = fmaxf(fmaxf(sm.wred[0][tid], sm.wred[1][tid]), fmaxf(sm.wred[2][tid], sm.wred[3][tid]));
        const float old = sm.ref[tid];
        const float nref = nm == -INFINITY ? old : fmaxf(old, ceilf(nm));
        const float dref = old - nref;  // 0, a negative integer, or -inf / nan (old == -inf)
        sm.fac[tid] = (old == -INFINITY) ? (nref == -INFINITY ? 1.f : 0.f)
                                         : (dref < -126.f ? 0.f : __int_as_float((127 + (int)dref) << 23));
        sm.ref[tid] = nref;
      }
      __syncthreads();
      if (smx) lpart *= sm.fac[lane];
      if (warp == 4 || warp == 5) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= sm.fac[i];
      }
    }
    if (smx) {
      const uint32_t pa = sk_u32(sm.p);
      const int kb = tid >> 5, kk = tid & 31;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float rr = sm.ref[i];
        const float pv = valid ? sk_ex2(s[i] - (rr == -INFINITY ? 0.f : rr)) : 0.f;
        const uint32_t hi = __float_as_uint(pv) & 0xFFFFE000u;
        const uint32_t off = 4u * (uint32_t)(kb * 2048 + sk_off(i, kk));
        sts32(pa + off, __uint_as_float(hi));
        sts32(pa + off + 4096u, pv - __uint_as_float(hi));
        s[i] = pv;
      }
      lpart += sk_transpose_reduce<false>(s, lane);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- O^T tile = [V_hi^T ; V_lo^T] [P_hi ; P_lo]^T ----------------------------------------------------------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t dv = sk_desc(sk_u32(sm.vt)), dp = sk_desc(sk_u32(sm.p));
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ob = (uint64_t)(kb * (8192 >> 4) + ks * 2);  // 8 KB per key block, 32 B per 8 keys
          sk_mma(tmem + 64, dv + ob, dp + ob, id64, (kb | ks) ? 1u : 0u);
        }
      sk_commit(&sm.mma_bar);
    }
    sk_wait(&sm.mma_bar, ph); ph ^= 1;  // everybody: the next tile overwrites V^T and P
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 || warp == 5) {
      uint32_t r[32];
      const uint32_t la = tmem + ((uint32_t)(32 * (warp - 4)) << 16) + 64;
      sk_ld32(la, r);  // V_hi P_hi (warp 4) / V_lo P_hi (warp 5)
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(r[i]);
      if (warp == 4) {
        sk_ld32(la + 32, r);  // V_hi P_lo
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(r[i]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  // ---- combine: O[q, d] = (O_hi^T + O_lo^T)[d, q] / l[q], plus the row's own new rtg key of the second pass ------------
  float (*oex)[DH][33] = reinterpret_cast<float (*)[DH][33]>(sm.p);  // the P tile is free now (all MMAs have retired)
  if (warp == 4 || warp == 5) {
#pragma unroll
    for (int i = 0; i < 32; ++i) oex[warp - 4][lane][i] = o[i];
  }
  if (warp < 4) sm.wred[warp][lane] = lpart;
  __syncthreads();
  {
    const int d = tid & 31;
    for (int qi = tid >> 5; qi < A; qi += 8) {
      float acc = oex[0][d][qi] + oex[1][d][qi];
      float l = (sm.wred[0][qi] + sm.wred[1][qi]) + (sm.wred[2][qi] + sm.wred[3][qi]);
      if (own_row) {
        const float* row = qkv_rows + ((size_t)g * A + qi) * (3 * H) + h * DH;
        float so = 0.f;
#pragma unroll
        for (int c = 0; c < DH; ++c) so = fmaf(row[c] * scale, row[H + c], so);
        const float rf = sm.ref[qi];
        const float mn = fmaxf(rf, so);
        const float f = rf == -INFINITY ? 0.f : exp2f(rf - mn), pw = exp2f(so - mn);
        acc = fmaf(acc, f, pw * row[2 * H + d]);
        l = fmaf(l, f, pw);
      }
      O[((size_t)g * A + qi) * H + h * DH + d] = acc / l;
    }
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(SK_TMEM_COLS) : "memory");
  }
}

int launch_attn_step_tc(const KvView& kv, const float* qkv_rows, float* O, int G, int ti, bool own_row, cudaStream_t st) {
  if (G <= 0) return 0;
  static bool attr_set = false;
  const int smem = (int)sizeof(SkSmem) + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "attn_step_tc smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if ((kv.ld & 3) || (kv.k_off & 3) || (kv.v_off & 3) || (reinterpret_cast<uintptr_t>(kv.base) & 15))
    return set_error(-2, "attn_step_tc: K/V rows must be 16-byte aligned");
  dim3 grid(NH, G);
  attn_step_tc_kernel<<<grid, 256, smem, st>>>(kv.base, kv.ld, kv.k_off, kv.v_off, kv.group_rows, qkv_rows, O, ti, own_row ? 1 : 0);
  CS_CHECK_LAUNCH("attn_step_tc");
  return 0;
}

}  // namespace ctrlsim
