// attn_step: decoder self-attention (mask rule M1, utils/train_utils.py:82-130) for the A rows of the CURRENT window
// step only - the state rows of the last first-pass layer (the rows the RTG head reads, modules/decoder.py:75) and the
// rtg rows of the second pass (policies/autoregressive_policy.py:210) - against the keys / values the first pass left
// in HBM.  Visible keys of row (ti, a, k): every token of window steps < ti, the A state tokens of step ti, and - second
// pass - the row's own freshly computed rtg key.  Decision-transformer order (rtg, state, action): the state tokens sit at
// position 1 of an agent's step and a state row also sees its own rtg token of step ti (own_mode 2, read from the K/V buffer).
//
// HBM-bound by construction: A query rows per (group, head) against up to 2304 K/V rows of 128 B each (4.7 MB per
// group and layer), 2 x 32 flops per (query, key).  A CTA is one (group, head); warp w owns query rows 16 w .. 16 w + 15
// (FlashAttention-2 register dataflow on warp-level mma.sync.m16n8k8 TF32 with the 3xTF32 hi/lo split of
// attention_mma.cu: S, P and O never leave registers) and all warps stream the same 64-key K/V tiles through a
// double-buffered cp.async ring.  A first SIMT version (thread = query, later a register-tiled outer product) was
// bound by shared-memory wavefronts - a broadcast LDS.128 still costs 4 - at 0.21 of the HBM roofline; with the MMA
// fragments a K/V element is read from shared memory once per 16 query rows.
#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int ST_KT = 64;                 // keys per tile
constexpr int ST_LD = DH + 4;             // padded row stride (floats): both fragment access patterns are conflict free
constexpr int ST_WARPS = (A + 15) / 16;   // 2 for the default model (24 rows), 4 for the wide one (64 rows)
constexpr float kStLog2e = 1.4426950408889634f;

struct StSmem {
  float k[2][ST_KT][ST_LD];
  float v[2][ST_KT][ST_LD];
};

__device__ __forceinline__ void st_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// 16-byte asynchronous copy; n_src = 0 writes zeros (rows past the end of a tile must be finite: 0 * garbage = NaN)
__device__ __forceinline__ void st_cp16(uint32_t dst, const void* src, int n_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n_src) : "memory");
}
__device__ __forceinline__ void st_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void st_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(ST_WARPS * 32)
attn_step_kernel(const float* __restrict__ KVbuf, int ld, int k_off, int v_off, int group_rows,
                 const float* __restrict__ qkv_rows, float* __restrict__ O, int ti, int own_row, int si) {
  // own_row: 0 none; 1 the row's own NEW key / value (qkv_rows); 2 the row's own first token of step ti (K/V buffer).
  // si: position of the state token inside an agent's step (0, or 1 for the decision transformer)
  __shared__ __align__(16) StSmem sm;
  const int g = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq4 = lane & 3;  // fragment coordinates: group id, thread-in-group
  const int row_a = warp * 16 + gq, row_b = row_a + 8;
  const bool ok_a = row_a < A, ok_b = row_b < A;
  const float* base = KVbuf + (size_t)g * group_rows * ld + h * DH;
  const float* rows = qkv_rows + (size_t)g * A * (3 * H) + h * DH;
  const int n_hist = ti * TOK_T;                         // every token of the earlier window steps
  const int n_tail = own_row ? 2 * A : A;                // state tokens of step ti (+ the rows' own new keys)
  const int hist_tiles = (n_hist + ST_KT - 1) / ST_KT;
  const int n_tiles = hist_tiles + (n_tail + ST_KT - 1) / ST_KT;

  // tile t -> ring slot s: 64 keys x (K | V) x 8 chunks of 16 B
  auto issue = [&](int t, int s) {
    const bool hist = t < hist_tiles;
    const int e0 = hist ? t * ST_KT : (t - hist_tiles) * ST_KT;
    const int n = hist ? min(ST_KT, n_hist - e0) : min(ST_KT, n_tail - e0);
    for (int idx = tid; idx < ST_KT * 8; idx += ST_WARPS * 32) {
      const int r = idx >> 3, c = idx & 7;
      const float *ks = base + k_off, *vs = base + v_off;
      const bool in = r < n;
      if (in) {
        const int e = e0 + r;
        if (hist) { ks = base + (size_t)e * ld + k_off; vs = base + (size_t)e * ld + v_off; }
        else if (e < A) { const size_t tok = (size_t)n_hist + (size_t)e * KT + si; ks = base + tok * ld + k_off; vs = base + tok * ld + v_off; }
        else if (own_row == 2) { const size_t tok = (size_t)n_hist + (size_t)(e - A) * KT; ks = base + tok * ld + k_off; vs = base + tok * ld + v_off; }
        else { ks = rows + (size_t)(e - A) * (3 * H) + H; vs = ks + H; }
      }
      st_cp16((uint32_t)__cvta_generic_to_shared(&sm.k[s][r][c * 4]), ks + c * 4, in ? 16 : 0);
      st_cp16((uint32_t)__cvta_generic_to_shared(&sm.v[s][r][c * 4]), vs + c * 4, in ? 16 : 0);
    }
  };
  issue(0, 0);
  st_commit();

  // Q fragments (scaled by d_h^-0.5 log2 e), split once
  uint32_t qh[4][4], ql[4][4];
  {
    const float sc = 0.17677669529663687f * kStLog2e;
    const float* qa = rows + (size_t)(ok_a ? row_a : 0) * (3 * H);
    const float* qb = rows + (size_t)(ok_b ? row_b : 0) * (3 * H);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float v0 = ok_a ? qa[8 * ks + tq4] * sc : 0.f, v1 = ok_b ? qb[8 * ks + tq4] * sc : 0.f;
      const float v2 = ok_a ? qa[8 * ks + tq4 + 4] * sc : 0.f, v3 = ok_b ? qb[8 * ks + tq4 + 4] * sc : 0.f;
      st_split(v0, qh[ks][0], ql[ks][0]); st_split(v1, qh[ks][1], ql[ks][1]);
      st_split(v2, qh[ks][2], ql[ks][2]); st_split(v3, qh[ks][3], ql[ks][3]);
    }
  }
  float o[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_a = -INFINITY, m_b = -INFINITY, l_a = 0.f, l_b = 0.f;

  for (int t = 0; t < n_tiles; ++t) {
    const int s = t & 1;
    if (t + 1 < n_tiles) issue(t + 1, s ^ 1);  // slot s^1 was released by the __syncthreads that ended iteration t-1
    st_commit();
    st_wait<1>();
    __syncthreads();
    const bool hist = t < hist_tiles;
    const int e0 = hist ? t * ST_KT : (t - hist_tiles) * ST_KT;
    const int nk = hist ? min(ST_KT, n_hist - e0) : min(ST_KT, n_tail - e0);
    const float(*sK)[ST_LD] = sm.k[s];
    const float(*sV)[ST_LD] = sm.v[s];
    // ---- S = Q K^T for the 8 key blocks of the tile (each block: 16 rows x 8 keys), cross products first -----------
    float sc[8][4];
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      sc[kb][0] = sc[kb][1] = sc[kb][2] = sc[kb][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bh0, bl0, bh1, bl1;
        st_split(sK[kb * 8 + gq][8 * ks + tq4], bh0, bl0);
        st_split(sK[kb * 8 + gq][8 * ks + tq4 + 4], bh1, bl1);
        st_mma(sc[kb], ql[ks], bh0, bh1);
        st_mma(sc[kb], qh[ks], bl0, bl1);
        st_mma(sc[kb], qh[ks], bh0, bh1);
      }
    }
    // ---- mask + online softmax: this lane holds keys kb*8 + 2*tq4 + {0,1} of rows a (c0, c1) and b (c2, c3) --------
    float mx_a = -INFINITY, mx_b = -INFINITY;
    if (hist && nk == ST_KT) {  // a full tile of earlier steps: everything is visible to every real row
#pragma unroll
      for (int kb = 0; kb < 8; ++kb) {
        if (!ok_a) { sc[kb][0] = -INFINITY; sc[kb][1] = -INFINITY; }
        if (!ok_b) { sc[kb][2] = -INFINITY; sc[kb][3] = -INFINITY; }
        mx_a = fmaxf(mx_a, fmaxf(sc[kb][0], sc[kb][1]));
        mx_b = fmaxf(mx_b, fmaxf(sc[kb][2], sc[kb][3]));
      }
    } else {
#pragma unroll
      for (int kb = 0; kb < 8; ++kb)
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int kl = kb * 8 + 2 * tq4 + e2, e = e0 + kl;
          const bool vis = kl < nk && (hist || e < A);   // history and the step's state tokens: visible to all rows
          const bool va = ok_a && (vis || (kl < nk && e - A == row_a));  // own new key (tail entries A .. 2A-1)
          const bool vb = ok_b && (vis || (kl < nk && e - A == row_b));
          if (!va) sc[kb][e2] = -INFINITY;
          if (!vb) sc[kb][2 + e2] = -INFINITY;
          mx_a = fmaxf(mx_a, sc[kb][e2]);
          mx_b = fmaxf(mx_b, sc[kb][2 + e2]);
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
      sc[kb][0] = exp2f(sc[kb][0] - ref_a); sc[kb][1] = exp2f(sc[kb][1] - ref_a);
      sc[kb][2] = exp2f(sc[kb][2] - ref_b); sc[kb][3] = exp2f(sc[kb][3] - ref_b);
      sum_a += sc[kb][0] + sc[kb][1];
      sum_b += sc[kb][2] + sc[kb][3];
    }
    sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 1); sum_a += __shfl_xor_sync(0xffffffffu, sum_a, 2);
    sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 1); sum_b += __shfl_xor_sync(0xffffffffu, sum_b, 2);
    l_a = l_a * corr_a + sum_a; l_b = l_b * corr_b + sum_b;
    m_a = mn_a; m_b = mn_b;
    // ---- tile output: P V with a fresh accumulator (24 chained MMAs), then one rounded fp32 update of O -----------
    float pv[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pv[j][0] = pv[j][1] = pv[j][2] = pv[j][3] = 0.f;
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      // A fragment of P: k-index tq4 <-> key 2*tq4 (c0 / c2), k-index tq4+4 <-> key 2*tq4+1 (c1 / c3): the keys of a
      // block are consumed in a permuted order, which a sum over keys does not notice
      uint32_t ph[4], pl[4];
      st_split(sc[kb][0], ph[0], pl[0]); st_split(sc[kb][2], ph[1], pl[1]);
      st_split(sc[kb][1], ph[2], pl[2]); st_split(sc[kb][3], ph[3], pl[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t vh0, vl0, vh1, vl1;
        st_split(sV[kb * 8 + 2 * tq4][8 * j + gq], vh0, vl0);
        st_split(sV[kb * 8 + 2 * tq4 + 1][8 * j + gq], vh1, vl1);
        st_mma(pv[j], pl, vh0, vh1);
        st_mma(pv[j], ph, vl0, vl1);
        st_mma(pv[j], ph, vh0, vh1);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j][0] = fmaf(o[j][0], corr_a, pv[j][0]); o[j][1] = fmaf(o[j][1], corr_a, pv[j][1]);
      o[j][2] = fmaf(o[j][2], corr_b, pv[j][2]); o[j][3] = fmaf(o[j][3], corr_b, pv[j][3]);
    }
    __syncthreads();  // ring slot s is free again
  }
  // ---- normalise and store: lane holds dims 8j + 2*tq4 + {0,1} of rows a and b ---------------------------------------
  if (ok_a) {
    const float inv = 1.0f / l_a;
    float* dst = O + ((size_t)g * A + row_a) * H + h * DH + 2 * tq4;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(o[j][0] * inv, o[j][1] * inv);
  }
  if (ok_b) {
    const float inv = 1.0f / l_b;
    float* dst = O + ((size_t)g * A + row_b) * H + h * DH + 2 * tq4;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(dst + 8 * j) = make_float2(o[j][2] * inv, o[j][3] * inv);
  }
}

int launch_attn_step(const KvView& kv, const float* qkv_rows, float* O, int G, int ti, int own_mode, cudaStream_t st,
                     int si) {
  if (G <= 0) return 0;
  if (own_mode < 0 || own_mode > 2 || si < 0 || si >= KT) return set_error(-2, "attn_step: own_mode=%d state_index=%d", own_mode, si);
  dim3 grid(NH, G);
  attn_step_kernel<<<grid, ST_WARPS * 32, 0, st>>>(kv.base, kv.ld, kv.k_off, kv.v_off, kv.group_rows, qkv_rows, O, ti,
                                                   own_mode, si);
  CS_CHECK_LAUNCH("attn_step");
  return 0;
}

}  // namespace ctrlsim
