// attn_step: decoder self-attention (mask rule M1, utils/train_utils.py:82-130) for the A rows of the CURRENT window
// step only - the state rows of the last first-pass layer (the rows the RTG head reads, modules/decoder.py:75) and the
// rtg rows of the second pass (policies/autoregressive_policy.py:210) - against the keys / values the first pass left
// in HBM.  Visible keys of row (ti, a, k): every token of window steps < ti, the A state tokens of step ti, and - second
// pass - the row's own freshly computed rtg key.
//
// HBM-bound by construction: A query rows per (group, head) against up to 2304 K/V rows of 128 B each (4.7 MB per
// group and layer), 2 x 32 flops per (query, key).  fp32 SIMT, laid out so that the FP32 pipe - not the load/store
// unit - is the busiest unit:
//   * one CTA = one (group, head), 4 warps; warp w streams the 32-key tiles w, w+4, ... through its own 2-stage
//     cp.async ring (K and V rows, 16 B chunks XOR-swizzled by the row index: conflict-free 128-bit reads);
//   * S = Q K^T and O += P V as register-tiled outer products: lane (qg = lane / 4, kg = lane % 4) owns the queries
//     {qg + 8 i} and the keys {kg + 4 j} (then the dims {8 kg .. 8 kg + 7}): 11 LDS.128 feed 96 FMAs;
//   * online softmax per tile with two shuffles per row (the 4 lanes that share a query), P staged through 3.4 KB of
//     shared memory per warp; the four warps' partial (m, l, O) states are merged once at the end.
#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int ST_TILE = 32;            // keys per tile
constexpr int ST_STAGES = 2;
constexpr int ST_WARPS = 4;
constexpr int ST_QPL = A / 8;          // queries per lane
constexpr int ST_QS = DH + 4;          // padded row stride (floats) of the Q and P tiles: 8 consecutive rows hit 8 bank groups
static_assert(A % 8 == 0, "attn_step: the query rows are dealt to 8 lane groups");

struct StWarp {
  float kv[ST_STAGES][2][ST_TILE * DH];  // [stage][K | V][key][dim], chunk c of row r stored at chunk c ^ (r & 7)
  float p[A * ST_QS];
};
struct StSmem {
  float q[A * ST_QS];
  StWarp w[ST_WARPS];
};

__device__ __forceinline__ void st_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void st_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void st_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(ST_WARPS * 32, 2)
attn_step_kernel(const float* __restrict__ KVbuf, int ld, int k_off, int v_off, int group_rows,
                 const float* __restrict__ qkv_rows, float* __restrict__ O, int ti, int own_row) {
  extern __shared__ __align__(16) unsigned char st_raw[];
  StSmem& sm = *reinterpret_cast<StSmem*>(st_raw);
  const int g = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qg = lane >> 2, kg = lane & 3;
  const float* base = KVbuf + (size_t)g * group_rows * ld + h * DH;
  const float* rows = qkv_rows + (size_t)g * A * (3 * H) + h * DH;
  const int n_hist = ti * TOK_T;                         // every token of the earlier window steps
  const int n_tail = own_row ? 2 * A : A;                // state tokens of step ti (+ the rows' own new keys)
  const int hist_tiles = (n_hist + ST_TILE - 1) / ST_TILE;
  const int n_tiles = hist_tiles + (n_tail + ST_TILE - 1) / ST_TILE;

  // Q, scaled by d_h^-0.5 log2(e) (scores live in the log2 domain)
  {
    const float sc = 0.17677669529663687f * 1.4426950408889634f;
    for (int i = tid; i < A * (DH / 4); i += ST_WARPS * 32) {
      const int r = i >> 3, c = (i & 7) << 2;
      float4 t = *reinterpret_cast<const float4*>(rows + (size_t)r * (3 * H) + c);
      t.x *= sc; t.y *= sc; t.z *= sc; t.w *= sc;
      *reinterpret_cast<float4*>(&sm.q[r * ST_QS + c]) = t;
    }
  }
  StWarp& W = sm.w[warp];
  const uint32_t kv_a = (uint32_t)__cvta_generic_to_shared(&W.kv[0][0][0]);

  // asynchronous copy of tile t into ring slot s: 32 keys x (K | V) x 8 chunks of 16 B = 16 copies per lane
  auto issue = [&](int t, int s) {
    const bool hist = t < hist_tiles;
    const int e0 = hist ? t * ST_TILE : (t - hist_tiles) * ST_TILE;
    const int n = hist ? min(ST_TILE, n_hist - e0) : min(ST_TILE, n_tail - e0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = lane + 32 * i, r = idx >> 3, c = idx & 7;
      if (r < n) {
        const int e = e0 + r;
        const float *ks, *vs;
        if (hist) { ks = base + (size_t)e * ld + k_off; vs = base + (size_t)e * ld + v_off; }
        else if (e < A) { const size_t tok = (size_t)n_hist + (size_t)e * KT; ks = base + tok * ld + k_off; vs = base + tok * ld + v_off; }
        else { ks = rows + (size_t)(e - A) * (3 * H) + H; vs = ks + H; }
        const uint32_t d = kv_a + (uint32_t)(((s * 2) * ST_TILE * DH + r * DH + ((c ^ (r & 7)) << 2)) * 4);
        st_cp16(d, ks + c * 4);
        st_cp16(d + ST_TILE * DH * 4, vs + c * 4);
      }
    }
  };

  float o[ST_QPL][8], m[ST_QPL], l[ST_QPL];
#pragma unroll
  for (int i = 0; i < ST_QPL; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[i][j] = 0.f;
  }

  if (warp < n_tiles) issue(warp, 0);
  st_commit();
  __syncthreads();  // Q tile visible
  int it = 0;
  for (int t = warp; t < n_tiles; t += ST_WARPS, ++it) {
    const int s = it & 1;
    if (t + ST_WARPS < n_tiles) issue(t + ST_WARPS, s ^ 1);
    st_commit();
    st_wait<1>();
    __syncwarp();
    const bool hist = t < hist_tiles;
    const int e0 = hist ? t * ST_TILE : (t - hist_tiles) * ST_TILE;
    const int n = hist ? min(ST_TILE, n_hist - e0) : min(ST_TILE, n_tail - e0);
    const float* Ks = &W.kv[s][0][0];
    const float* Vs = &W.kv[s][1][0];
    // ---- S = Q K^T: rows {qg + 8 i}, keys {kg + 4 j} ----------------------------------------------------------
    float sc[ST_QPL][8];
#pragma unroll
    for (int i = 0; i < ST_QPL; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[i][j] = 0.f;
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      float4 kk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = kg + 4 * j;
        kk[j] = *reinterpret_cast<const float4*>(Ks + r * DH + ((c ^ (r & 7)) << 2));
      }
#pragma unroll
      for (int i = 0; i < ST_QPL; ++i) {
        const float4 qq = *reinterpret_cast<const float4*>(&sm.q[(qg + 8 * i) * ST_QS + 4 * c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sc[i][j] = fmaf(qq.x, kk[j].x, sc[i][j]); sc[i][j] = fmaf(qq.y, kk[j].y, sc[i][j]);
          sc[i][j] = fmaf(qq.z, kk[j].z, sc[i][j]); sc[i][j] = fmaf(qq.w, kk[j].w, sc[i][j]);
        }
      }
    }
    // ---- mask + online softmax (the 4 lanes of a query group share every row statistic) ------------------------
#pragma unroll
    for (int i = 0; i < ST_QPL; ++i) {
      const int q = qg + 8 * i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = kg + 4 * j, e = e0 + r;
        const bool ok = r < n && (hist || e < A || e - A == q);
        sc[i][j] = ok ? sc[i][j] : -INFINITY;  // rows beyond n hold stale shared memory: never used
        mx = fmaxf(mx, sc[i][j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float mn = fmaxf(m[i], mx);
      const float corr = mn == -INFINITY ? 1.f : exp2f(m[i] - mn);  // m = -inf on first use -> 0
      l[i] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[i][j] *= corr;
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = sc[i][j] == -INFINITY ? 0.f : exp2f(sc[i][j] - mn);
        ps += p;
        W.p[q * ST_QS + kg + 4 * j] = p;
      }
      l[i] += ps;  // per-lane partial; the 4 lanes are added once at the end
      m[i] = mn;
    }
    __syncwarp();
    // ---- O += P V: rows {qg + 8 i}, dims {8 kg .. 8 kg + 7}; masked keys carry p = 0, rows >= n are skipped ------
    const int nk4 = (n + 3) >> 2;
    for (int c = 0; c < nk4; ++c) {
      float4 pp[ST_QPL];
#pragma unroll
      for (int i = 0; i < ST_QPL; ++i) pp[i] = *reinterpret_cast<const float4*>(&W.p[(qg + 8 * i) * ST_QS + 4 * c]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int r = 4 * c + kk;
        if (r < n) {  // never touch a V row that was not loaded (0 * garbage could be NaN)
          const float4 v0 = *reinterpret_cast<const float4*>(Vs + r * DH + (((2 * kg) ^ (r & 7)) << 2));
          const float4 v1 = *reinterpret_cast<const float4*>(Vs + r * DH + (((2 * kg + 1) ^ (r & 7)) << 2));
#pragma unroll
          for (int i = 0; i < ST_QPL; ++i) {
            const float p = kk == 0 ? pp[i].x : kk == 1 ? pp[i].y : kk == 2 ? pp[i].z : pp[i].w;
            o[i][0] = fmaf(p, v0.x, o[i][0]); o[i][1] = fmaf(p, v0.y, o[i][1]);
            o[i][2] = fmaf(p, v0.z, o[i][2]); o[i][3] = fmaf(p, v0.w, o[i][3]);
            o[i][4] = fmaf(p, v1.x, o[i][4]); o[i][5] = fmaf(p, v1.y, o[i][5]);
            o[i][6] = fmaf(p, v1.z, o[i][6]); o[i][7] = fmaf(p, v1.w, o[i][7]);
          }
        }
      }
    }
    __syncwarp();  // the ring slot and the P tile are free again
  }
  st_wait<0>();
  // ---- merge the four warps' partial states (through the, now idle, K/V ring of each warp) -----------------------
#pragma unroll
  for (int i = 0; i < ST_QPL; ++i) {
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 1);
    l[i] += __shfl_xor_sync(0xffffffffu, l[i], 2);
  }
  __syncthreads();
  float* part = &W.kv[0][0][0];  // [A][DH + 2]: o, m, l
#pragma unroll
  for (int i = 0; i < ST_QPL; ++i) {
    const int q = qg + 8 * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) part[q * (DH + 2) + 8 * kg + j] = o[i][j];
    if (kg == 0) { part[q * (DH + 2) + DH] = m[i]; part[q * (DH + 2) + DH + 1] = l[i]; }
  }
  __syncthreads();
  for (int i = tid; i < A * DH; i += ST_WARPS * 32) {
    const int q = i >> 5, d = i & 31;
    float ms = -INFINITY;
#pragma unroll
    for (int w = 0; w < ST_WARPS; ++w) ms = fmaxf(ms, sm.w[w].kv[0][0][q * (DH + 2) + DH]);
    float lt = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < ST_WARPS; ++w) {
      const float* pw = &sm.w[w].kv[0][0][q * (DH + 2)];
      const float f = pw[DH] == -INFINITY ? 0.f : exp2f(pw[DH] - ms);
      lt = fmaf(pw[DH + 1], f, lt);
      acc = fmaf(pw[d], f, acc);
    }
    O[((size_t)g * A + q) * H + h * DH + d] = acc / lt;
  }
}

int launch_attn_step(const KvView& kv, const float* qkv_rows, float* O, int G, int ti, bool own_row, cudaStream_t st) {
  if (G <= 0) return 0;
  static bool attr_set = false;
  const int smem = (int)sizeof(StSmem);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "attn_step smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(NH, G);
  attn_step_kernel<<<grid, ST_WARPS * 32, smem, st>>>(kv.base, kv.ld, kv.k_off, kv.v_off, kv.group_rows, qkv_rows, O, ti,
                                                      own_row ? 1 : 0);
  CS_CHECK_LAUNCH("attn_step");
  return 0;
}

}  // namespace ctrlsim
