// Tensor-core flash attention for the two big attention shapes of the decoder (and the scene encoder):
//   attn_causal  self-attention over the (timestep, agent, type) tokens with mask rule M1 (utils/train_utils.py:82-130)
//   attn_padded  attention over the 224 memory tokens with a key-padding mask (modules/encoder.py:155-168, decoder.py:52)
// 8 heads x d_h = 32, fp32 in / fp32 out.
//
// Arithmetic: warp-level mma.sync.m16n8k8 TF32 with the same hi/lo operand split as gemm_tc.cu ("3xTF32": x = hi + lo,
// three MMAs per product, the lo*lo term dropped), so scores and outputs keep fp32-class accuracy.  The tensor core
// accumulates with truncation, therefore no accumulator chains more than 24 MMAs: S uses 12 (4 k-steps x 3), the PV
// product of a 64-key tile uses 24 and is then added to the running output with an ordinary rounded fp32 FMA together
// with the online-softmax rescale.
// Layout: one warp owns 16 query rows (FlashAttention-2 register dataflow: S, P and O never leave registers; the
// C-fragment of S is reused directly as the A-fragment of P by permuting the key order inside each block of 8, which
// is legal because softmax(QK^T)V is a sum over keys).  A CTA of 4 warps = 64 query rows streams K/V tiles of 64 keys
// through shared memory (row stride 36 floats: both fragment access patterns are bank-conflict free).
// Why mma.sync and not tcgen05 here: P would have to round-trip through shared memory as an MMA A-operand (128 KB per
// 128x128 tile with the hi/lo split) - DESIGN.md section 8 keeps the tcgen05/TMEM version as the next step.
#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int FA_ROWS = 64;      // query rows per CTA (4 warps x 16)
constexpr int FA_KT = 64;        // keys per shared-memory tile
constexpr int FA_LD = 36;        // padded row stride (floats)
constexpr float kFaLog2e = 1.4426950408889634f;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ bool fa_m1_allowed(int tq, int aq, int kq, int key, int si) {
  const int tk = key / TOK_T;
  if (tk < tq) return true;
  if (tk > tq) return false;
  const int rem = key - tk * TOK_T;
  const int ak = rem / KT, kk = rem - ak * KT;
  return kk == si || (ak == aq && kk <= kq);
}

// CAUSAL: Q/K/V are column blocks of QKV [G*L, 768]; PADDED: Q [G*Lq, ldq], K/V [G*Lk, ldkv] + key_pad [G*Lk].
template <bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_mma_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ Kp, const float* __restrict__ Vp, int ldkv,
                const uint8_t* __restrict__ key_pad, float* __restrict__ O, int ldo, int Lq, int Lk, int si) {
  __shared__ __align__(16) float sK[FA_KT][FA_LD];
  __shared__ __align__(16) float sV[FA_KT][FA_LD];
  __shared__ uint8_t sPad[FA_KT];
  const int g = blockIdx.z, h = blockIdx.y;
  const int r0 = blockIdx.x * FA_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq4 = lane & 3;            // fragment coordinates: group id, thread-in-group
  const int row_a = r0 + warp * 16 + gq, row_b = row_a + 8;
  const bool ok_a = row_a < Lq, ok_b = row_b < Lq;

  // Q fragments (scaled by d_h^-0.5 * log2 e), split once
  uint32_t qh[4][4], ql[4][4];
  {
    const float sc = 0.17677669529663687f * kFaLog2e;
    const float* qa = Q + ((size_t)g * Lq + (ok_a ? row_a : 0)) * ldq + h * DH;
    const float* qb = Q + ((size_t)g * Lq + (ok_b ? row_b : 0)) * ldq + h * DH;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float v0 = ok_a ? qa[8 * ks + tq4] * sc : 0.f, v1 = ok_b ? qb[8 * ks + tq4] * sc : 0.f;
      const float v2 = ok_a ? qa[8 * ks + tq4 + 4] * sc : 0.f, v3 = ok_b ? qb[8 * ks + tq4 + 4] * sc : 0.f;
      split_tf32(v0, qh[ks][0], ql[ks][0]); split_tf32(v1, qh[ks][1], ql[ks][1]);
      split_tf32(v2, qh[ks][2], ql[ks][2]); split_tf32(v3, qh[ks][3], ql[ks][3]);
    }
  }
  // causal bookkeeping of this lane's two rows
  int t_a = 0, a_a = 0, k_a = 0, t_b = 0, a_b = 0, k_b = 0;
  if (CAUSAL) {
    t_a = row_a / TOK_T; { const int rem = row_a - t_a * TOK_T; a_a = rem / KT; k_a = rem - a_a * KT; }
    t_b = row_b / TOK_T; { const int rem = row_b - t_b * TOK_T; a_b = rem / KT; k_b = rem - a_b * KT; }
  }
  // keys needed by this CTA / this warp
  int cta_end = Lk, warp_end = Lk;
  if (CAUSAL) {
    const int last_row = min(r0 + FA_ROWS - 1, Lq - 1);
    cta_end = min(Lk, (last_row / TOK_T + 1) * TOK_T);
    const int wlast = min(r0 + warp * 16 + 15, Lq - 1);
    warp_end = (r0 + warp * 16 < Lq) ? min(Lk, (wlast / TOK_T + 1) * TOK_T) : 0;
  }
  float o[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

  for (int k0 = 0; k0 < cta_end; k0 += FA_KT) {
    const int nk = min(FA_KT, cta_end - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < FA_KT * (DH / 4); i += 128) {
      const int r = i >> 3, c = (i & 7) << 2;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (r < nk) {
        const size_t src = ((size_t)g * Lk + k0 + r) * ldkv + h * DH + c;
        kk = *reinterpret_cast<const float4*>(Kp + src);
        vv = *reinterpret_cast<const float4*>(Vp + src);
      }
      *reinterpret_cast<float4*>(&sK[r][c]) = kk;
      *reinterpret_cast<float4*>(&sV[r][c]) = vv;
    }
    if (!CAUSAL) for (int i = threadIdx.x; i < FA_KT; i += 128) sPad[i] = (i < nk) ? key_pad[(size_t)g * Lk + k0 + i] : 1;
    __syncthreads();
    if (k0 >= warp_end) continue;

    // ---- S = Q K^T for the 8 key blocks of this tile (each block: 16 rows x 8 keys) ----
    float s[8][4];
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      s[kb][0] = s[kb][1] = s[kb][2] = s[kb][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(sK[kb * 8 + gq][8 * ks + tq4], bh0, bl0);
        split_tf32(sK[kb * 8 + gq][8 * ks + tq4 + 4], bh1, bl1);
        mma_tf32(s[kb], ql[ks], bh0, bh1);
        mma_tf32(s[kb], qh[ks], bl0, bl1);
        mma_tf32(s[kb], qh[ks], bh0, bh1);
      }
    }
    // ---- mask + online softmax (this lane holds keys kb*8 + 2*tq4 + {0,1} of rows a (c0,c1) and b (c2,c3)) ----
    float mx_a = -INFINITY, mx_b = -INFINITY;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kl = kb * 8 + 2 * tq4 + e;
        const int key = k0 + kl;
        bool va, vb;
        if (CAUSAL) {
          va = ok_a && kl < nk && (key < t_a * TOK_T || fa_m1_allowed(t_a, a_a, k_a, key, si));
          vb = ok_b && kl < nk && (key < t_b * TOK_T || fa_m1_allowed(t_b, a_b, k_b, key, si));
        } else {
          const bool kv = kl < nk && !sPad[kl];
          va = kv; vb = kv;
        }
        if (!va) s[kb][e] = -INFINITY;
        if (!vb) s[kb][2 + e] = -INFINITY;
        mx_a = fmaxf(mx_a, s[kb][e]);
        mx_b = fmaxf(mx_b, s[kb][2 + e]);
      }
    }
    mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 1)); mx_a = fmaxf(mx_a, __shfl_xor_sync(0xffffffffu, mx_a, 2));
    mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 1)); mx_b = fmaxf(mx_b, __shfl_xor_sync(0xffffffffu, mx_b, 2));
    const float mn_a = fmaxf(m_a, mx_a), mn_b = fmaxf(m_b, mx_b);
    // rows with nothing visible yet keep m = -inf; use 0 as the reference to avoid inf - inf
    const float ref_a = mn_a == -INFINITY ? 0.f : mn_a, ref_b = mn_b == -INFINITY ? 0.f : mn_b;
    const float corr_a = exp2f(m_a - ref_a), corr_b = exp2f(m_b - ref_b);
    float sum_a = 0.f, sum_b = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      s[kb][0] = exp2f(s[kb][0] - ref_a); s[kb][1] = exp2f(s[kb][1] - ref_a);
      s[kb][2] = exp2f(s[kb][2] - ref_b); s[kb][3] = exp2f(s[kb][3] - ref_b);
      sum_a += s[kb][0] + s[kb][1];
      sum_b += s[kb][2] + s[kb][3];
    }
    sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 1); sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 2);
    sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 1); sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 2);
    l_a = l_a * corr_a + sum_a; l_b = l_b * corr_b + sum_b;
    m_a = mn_a; m_b = mn_b;
    // ---- tile output: PV with a fresh accumulator (<= 24 chained MMAs), then one rounded fp32 update of O ----
    float pv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pv[j][0] = pv[j][1] = pv[j][2] = pv[j][3] = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      // A fragment of P: k-index tq4 <-> key 2*tq4 (c0 / c2), k-index tq4+4 <-> key 2*tq4+1 (c1 / c3)
      uint32_t ph[4], pl[4];
      split_tf32(s[kb][0], ph[0], pl[0]); split_tf32(s[kb][2], ph[1], pl[1]);
      split_tf32(s[kb][1], ph[2], pl[2]); split_tf32(s[kb][3], ph[3], pl[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t vh0, vl0, vh1, vl1;
        split_tf32(sV[kb * 8 + 2 * tq4][8 * j + gq], vh0, vl0);
        split_tf32(sV[kb * 8 + 2 * tq4 + 1][8 * j + gq], vh1, vl1);
        mma_tf32(pv[j], pl, vh0, vh1);
        mma_tf32(pv[j], ph, vl0, vl1);
        mma_tf32(pv[j], ph, vh0, vh1);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j][0] = fmaf(o[j][0], corr_a, pv[j][0]); o[j][1] = fmaf(o[j][1], corr_a, pv[j][1]);
      o[j][2] = fmaf(o[j][2], corr_b, pv[j][2]); o[j][3] = fmaf(o[j][3], corr_b, pv[j][3]);
    }
  }
  // ---- normalise and store: lane holds dims 8j + 2*tq4 + {0,1} of rows a and b ----
  const float inv_a = 1.0f / l_a, inv_b = 1.0f / l_b;
  if (ok_a) {
    float* dst = O + ((size_t)g * Lq + row_a) * ldo + h * DH + 2 * tq4;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(o[j][0] * inv_a, o[j][1] * inv_a);
  }
  if (ok_b) {
    float* dst = O + ((size_t)g * Lq + row_b) * ldo + h * DH + 2 * tq4;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(o[j][2] * inv_b, o[j][3] * inv_b);
  }
}

int launch_attn_padded_mma(const float* Q, int ldq, const float* Kp, const float* Vp, int ldkv, const uint8_t* key_pad,
                           float* O, int ldo, int G, int Lq, int Lk, cudaStream_t st) {
  if (G <= 0 || Lq <= 0) return 0;
  dim3 grid((Lq + FA_ROWS - 1) / FA_ROWS, NH, G);
  attn_mma_kernel<false><<<grid, 128, 0, st>>>(Q, ldq, Kp, Vp, ldkv, key_pad, O, ldo, Lq, Lk, 0);
  CS_CHECK_LAUNCH("attn_padded_mma");
  return 0;
}

int launch_attn_causal_mma(const float* QKV, float* O, int G, int n_t, cudaStream_t st, int si) {
  if (G <= 0 || n_t <= 0) return 0;
  const int Lcur = n_t * TOK_T;
  dim3 grid((Lcur + FA_ROWS - 1) / FA_ROWS, NH, G);
  attn_mma_kernel<true><<<grid, 128, 0, st>>>(QKV, 3 * H, QKV + H, QKV + 2 * H, 3 * H, nullptr, O, H, Lcur, Lcur, si);
  CS_CHECK_LAUNCH("attn_causal_mma");
  return 0;
}

}  // namespace ctrlsim
