// Small-head (8 x 32) attention kernels of the policy network, fp32 SIMT with warp/online softmax.
//
//  attn_padded   Q[G,Lq] x K,V[G,Lk] with a per-group key-padding mask: the 2-layer scene encoder over 224
//                [polyline | initial-state] tokens (modules/encoder.py:155-168) and the decoder's cross-attention
//                (modules/decoder.py:52, memory_key_padding_mask).  torch semantics: q scaled by d_h^-0.5, masked
//                keys get -inf, softmax, P.V  (torch.nn.functional.multi_head_attention_forward).
//  attn_causal   decoder self-attention over (timestep, agent, {state,rtg,action}) tokens with the structured mask
//                of utils/train_utils.py:82-130 evaluated arithmetically (rule M1) instead of reading the 21 MB mask.
//  attn_step     the same rule for 24 rows of the current timestep only, streaming the layer's K/V rows written by the
//                first pass: the rtg-token rows of the second pass (after the RTGs were sampled; own new key appended)
//                and the state-token rows of the last first-pass layer (the only rows the RTG head reads).
//  map_pool      the polyline encoder's single learned-query attention over the 100 points of a polyline
//                (modules/map_encoder.py:41-45) in its algebraically reduced form: scores = feats . U_h with
//                U_h = W_k,h^T q_h (the key bias shifts every score of a head equally and cancels in the softmax),
//                pooled_h = sum_p softmax_h(p) feats_p; the value/out projections are applied afterwards to the 8
//                pooled vectors by one GEMM.  This is the HBM-bound "encoder-attn" kernel of BASELINE.json: each
//                polyline's 100x256 fp32 feature tile (102,400 B) is fetched by one TMA bulk copy
//                (cp.async.bulk + mbarrier) into a 2-stage shared-memory ring by a persistent CTA per SM.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr float kLog2e = 1.4426950408889634f;

// -------------------------------------------------------------------------------------------------------------
// one thread = one query row; K/V of one (group, head) staged through shared memory in chunks of CH keys.
template <int CH>
struct KVTile {
  float k[CH][DH];
  float v[CH][DH];
};

__device__ __forceinline__ void online_block8(const float (&q)[DH], const float (*ks)[DH], const float (*vs)[DH],
                                               int base, const bool (&ok)[8], float& m, float& l, float (&acc)[DH]) {
  float s[8];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      float4 kk = *reinterpret_cast<const float4*>(&ks[base + j][c]);
      d = fmaf(q[c], kk.x, d); d = fmaf(q[c + 1], kk.y, d); d = fmaf(q[c + 2], kk.z, d); d = fmaf(q[c + 3], kk.w, d);
    }
    s[j] = ok[j] ? d : -INFINITY;
    mx = fmaxf(mx, s[j]);
  }
  if (mx == -INFINITY) return;
  const float mn = fmaxf(m, mx);
  const float corr = exp2f(m - mn);  // m = -inf on first use -> 0
  l *= corr;
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] *= corr;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (!ok[j]) continue;  // never touch a masked / out-of-range V row (0 * garbage could be NaN)
    const float p = exp2f(s[j] - mn);
    l += p;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      float4 vv = *reinterpret_cast<const float4*>(&vs[base + j][c]);
      acc[c] = fmaf(p, vv.x, acc[c]); acc[c + 1] = fmaf(p, vv.y, acc[c + 1]);
      acc[c + 2] = fmaf(p, vv.z, acc[c + 2]); acc[c + 3] = fmaf(p, vv.w, acc[c + 3]);
    }
  }
  m = mn;
}

__device__ __forceinline__ void load_q(const float* qrow, float (&q)[DH]) {
  const float sc = 0.17677669529663687f * kLog2e;  // d_h^-0.5, folded with log2(e) for exp2f
#pragma unroll
  for (int c = 0; c < DH; c += 4) {
    float4 t = *reinterpret_cast<const float4*>(qrow + c);
    q[c] = t.x * sc; q[c + 1] = t.y * sc; q[c + 2] = t.z * sc; q[c + 3] = t.w * sc;
  }
}

__device__ __forceinline__ void store_o(float* orow, const float (&acc)[DH], float l) {
  const float inv = 1.0f / l;
#pragma unroll
  for (int c = 0; c < DH; c += 4)
    *reinterpret_cast<float4*>(orow + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
}

constexpr int PCH = 112;  // 224 memory tokens = 2 chunks

__global__ void __launch_bounds__(128)
attn_padded_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ Kp, const float* __restrict__ Vp,
                   int ldkv, const uint8_t* __restrict__ key_pad, float* __restrict__ O, int ldo, int Lq, int Lk) {
  __shared__ __align__(16) KVTile<PCH> sm;
  __shared__ uint8_t spad[PCH];
  const int g = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * 128 + threadIdx.x;
  const bool active = row < Lq;
  float q[DH], acc[DH];
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  if (active) load_q(Q + ((size_t)g * Lq + row) * ldq + h * DH, q);
  for (int k0 = 0; k0 < Lk; k0 += PCH) {
    const int nk = min(PCH, Lk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (DH / 4); i += 128) {
      const int r = i >> 3, c = (i & 7) << 2;
      const size_t src = ((size_t)g * Lk + k0 + r) * ldkv + h * DH + c;
      *reinterpret_cast<float4*>(&sm.k[r][c]) = *reinterpret_cast<const float4*>(Kp + src);
      *reinterpret_cast<float4*>(&sm.v[r][c]) = *reinterpret_cast<const float4*>(Vp + src);
    }
    for (int i = threadIdx.x; i < PCH; i += 128) spad[i] = (i < nk) ? key_pad[(size_t)g * Lk + k0 + i] : 1;
    __syncthreads();
    if (active) {
      for (int b = 0; b < nk; b += 8) {
        bool ok[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ok[j] = (b + j < nk) && !spad[b + j];
        online_block8(q, sm.k, sm.v, b, ok, m, l, acc);
      }
    }
  }
  if (active) store_o(O + ((size_t)g * Lq + row) * ldo + h * DH, acc, l);
}

// CTRLSIM_ATTN=simt selects the FP32 FFMA kernels of this file (A/B testing); the default is attention_mma.cu.
static int attn_mode() {  // 0 = simt, 1 = mma.sync, 2 = tcgen05 (default; falls back to mma.sync for tiny query counts)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CTRLSIM_ATTN");
    mode = (e && std::string(e) == "simt") ? 0 : (e && std::string(e) == "mma") ? 1 : 2;
  }
  return mode;
}
static bool use_mma_attn() { return attn_mode() >= 1; }

int launch_attn_padded(const float* Q, int ldq, const float* Kp, const float* Vp, int ldkv, const uint8_t* key_pad,
                       float* O, int ldo, int G, int Lq, int Lk, cudaStream_t st) {
  if (G <= 0 || Lq <= 0) return 0;
  if (attn_mode() == 2 && Lq >= 128 && Lk <= 512 && Vp > Kp && ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(Kp)) & 15) == 0)
    return launch_attn_tc(false, Q, ldq, H, 0, Kp, ldkv, (int)(Vp - Kp) + H, 0, (int)(Vp - Kp), key_pad, O, ldo, G, Lq, Lk, st);
  if (use_mma_attn()) return launch_attn_padded_mma(Q, ldq, Kp, Vp, ldkv, key_pad, O, ldo, G, Lq, Lk, st);
  if (Lk % 8 != 0) return set_error(-2, "attn_padded: Lk=%d must be a multiple of 8", Lk);
  dim3 grid((Lq + 127) / 128, NH, G);
  attn_padded_kernel<<<grid, 128, 0, st>>>(Q, ldq, Kp, Vp, ldkv, key_pad, O, ldo, Lq, Lk);
  CS_CHECK_LAUNCH("attn_padded");
  return 0;
}

// -------------------------------------------------------------------------------------------------------------
// Rule M1: token index = (t*A + a)*3 + k. Query (tq, aq, kq) sees key (tk, ak, kk) iff
//   tk < tq, or tk == tq and (kk == si or (ak == aq and kk <= kq)),  si = position of the state token (0; 1 for the
//   decision transformer's (rtg, state, action) order).
__device__ __forceinline__ bool m1_allowed(int tq, int aq, int kq, int key, int si) {
  const int tk = key / TOK_T;
  if (tk < tq) return true;
  if (tk > tq) return false;
  const int rem = key - tk * TOK_T;
  const int ak = rem / KT, kk = rem - ak * KT;
  return kk == si || (ak == aq && kk <= kq);
}

constexpr int CCH = 64;

// QKV: [G * Lcur, 768] (q | k | v), Lcur = 72 * n_t rows per group. O: [G * Lcur, 256].
__global__ void __launch_bounds__(128)
attn_causal_kernel(const float* __restrict__ QKV, float* __restrict__ O, int Lcur, int si) {
  __shared__ __align__(16) KVTile<CCH> sm;
  const int g = blockIdx.z, h = blockIdx.y;
  const int r0 = blockIdx.x * 128;
  const int row = r0 + threadIdx.x;
  const bool active = row < Lcur;
  const int tq = row / TOK_T, rq = row - tq * TOK_T, aq = rq / KT, kq = rq - aq * KT;
  const int my_end = active ? (tq + 1) * TOK_T : 0;              // keys this row may see lie in [0, my_end)
  const int tile_end = min(Lcur, ((min(r0 + 127, Lcur - 1)) / TOK_T + 1) * TOK_T);
  float q[DH], acc[DH];
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.f;
  const float* base = QKV + (size_t)g * Lcur * (3 * H);
  if (active) load_q(base + (size_t)row * (3 * H) + h * DH, q);
  for (int k0 = 0; k0 < tile_end; k0 += CCH) {
    const int nk = min(CCH, tile_end - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * (DH / 4); i += 128) {
      const int r = i >> 3, c = (i & 7) << 2;
      const float* src = base + (size_t)(k0 + r) * (3 * H) + h * DH + c;
      *reinterpret_cast<float4*>(&sm.k[r][c]) = *reinterpret_cast<const float4*>(src + H);
      *reinterpret_cast<float4*>(&sm.v[r][c]) = *reinterpret_cast<const float4*>(src + 2 * H);
    }
    __syncthreads();
    if (active && k0 < my_end) {
      const int lim = min(nk, my_end - k0);
      for (int b = 0; b < lim; b += 8) {
        bool ok[8];
        if (k0 + b + 8 <= tq * TOK_T) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ok[j] = true;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) ok[j] = (b + j < lim) && m1_allowed(tq, aq, kq, k0 + b + j, si);
        }
        online_block8(q, sm.k, sm.v, b, ok, m, l, acc);
      }
    }
  }
  if (active) store_o(O + ((size_t)g * Lcur + row) * H + h * DH, acc, l);
}

int launch_attn_causal(const float* QKV, float* O, int G, int n_t, cudaStream_t st, int si) {
  if (G <= 0 || n_t <= 0) return 0;
  if (attn_mode() == 2 && n_t * TOK_T >= 128 && (reinterpret_cast<uintptr_t>(QKV) & 15) == 0)
    return launch_attn_tc(true, QKV, 3 * H, 3 * H, 0, QKV, 3 * H, 3 * H, H, 2 * H, nullptr, O, H, G, n_t * TOK_T, n_t * TOK_T, st,
                          0, 0, si);
  if (use_mma_attn()) return launch_attn_causal_mma(QKV, O, G, n_t, st, si);
  const int Lcur = n_t * TOK_T;
  dim3 grid((Lcur + 127) / 128, NH, G);
  attn_causal_kernel<<<grid, 128, 0, st>>>(QKV, O, Lcur, si);
  CS_CHECK_LAUNCH("attn_causal");
  return 0;
}

// Incremental decode (prefix cache): the last n_q rows of every group attend to the group's first Lk cached keys.
int launch_attn_causal_tail(const float* Q, int ldq, const KvView& kv, float* O, int G, int n_q, int Lk, cudaStream_t st,
                            int si) {
  if (G <= 0 || n_q <= 0) return 0;
  if (n_q < 128 || attn_mode() != 2 || ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(kv.base)) & 15))
    return set_error(-2, "attn_causal_tail: needs the tcgen05 kernel and at least 128 query rows (got %d)", n_q);
  return launch_attn_tc(true, Q, ldq, ldq, 0, kv.base, kv.ld, kv.ld, kv.k_off, kv.v_off, nullptr, O, H, G, n_q, Lk, st,
                        Lk - n_q, kv.group_rows, si);
}

// attn_step (the A rows of the current window step against the first pass' K/V rows) lives in attention_step.cu

// -------------------------------------------------------------------------------------------------------------
// map_pool: persistent CTAs, 2-stage TMA bulk pipeline over polylines.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int POOL_THREADS = 256;
constexpr uint32_t TILE_BYTES = NP * H * sizeof(float);  // 102,400

struct PoolSmem {
  float feats[2][NP][H];  // 2 x 100 KB
  float prob[NP][NH];     // scores, then softmax weights, of the current polyline (point-major: float4 reads)
  uint64_t full[2];
};

__global__ void __launch_bounds__(POOL_THREADS, 1)
map_pool_kernel(const float* __restrict__ feats, const uint8_t* __restrict__ pt_valid,
                const uint8_t* __restrict__ poly_valid, const float* __restrict__ U, float* __restrict__ pooled,
                int n_poly) {
  extern __shared__ __align__(128) unsigned char smraw[];
  PoolSmem& sm = *reinterpret_cast<PoolSmem*>(smraw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the folded query/key matrix U lives in registers: lane owns dims [4 lane, 4 lane + 4) and [128 + 4 lane, ...)
  float4 u0[NH], u1[NH];
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    u0[hh] = __ldg(reinterpret_cast<const float4*>(U + hh * H + lane * 4));
    u1[hh] = __ldg(reinterpret_cast<const float4*>(U + hh * H + 128 + lane * 4));
  }
  __syncthreads();
  const int stride = gridDim.x;
  int it = 0;
  if (tid == 0 && (int)blockIdx.x < n_poly) {
    mbar_expect_tx(&sm.full[0], TILE_BYTES);
    tma_bulk_g2s(&sm.feats[0][0][0], feats + (size_t)blockIdx.x * NP * H, TILE_BYTES, &sm.full[0]);
  }
  for (int pid = blockIdx.x; pid < n_poly; pid += stride, ++it) {
    const int s = it & 1;
    const int nxt = pid + stride;
    if (tid == 0 && nxt < n_poly) {  // stage s^1 was fully consumed before the __syncthreads that ended iteration it-1
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ... also as a reduction buffer (generic writes)
      mbar_expect_tx(&sm.full[s ^ 1], TILE_BYTES);
      tma_bulk_g2s(&sm.feats[s ^ 1][0][0], feats + (size_t)nxt * NP * H, TILE_BYTES, &sm.full[s ^ 1]);
    }
    mbar_wait(&sm.full[s], (it >> 1) & 1);
    float* out = pooled + (size_t)pid * (NH * H);
    if (!poly_valid[pid]) {  // padded polyline: never read as a key downstream, keep it finite
      for (int i = tid; i < NH * H; i += POOL_THREADS) out[i] = 0.f;
      __syncthreads();
      continue;
    }
    // scores: one warp per point; the 8 per-head partial dot products are reduced with a transposing butterfly
    // (4 + 2 + 1 + 2 shuffles instead of 8 x 5)
    for (int p = warp; p < NP; p += POOL_THREADS / 32) {
      const float4 f0 = *reinterpret_cast<const float4*>(&sm.feats[s][p][lane * 4]);
      const float4 f1 = *reinterpret_cast<const float4*>(&sm.feats[s][p][128 + lane * 4]);
      float d[NH];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        float x = f0.x * u0[hh].x;
        x = fmaf(f0.y, u0[hh].y, x); x = fmaf(f0.z, u0[hh].z, x); x = fmaf(f0.w, u0[hh].w, x);
        x = fmaf(f1.x, u1[hh].x, x); x = fmaf(f1.y, u1[hh].y, x); x = fmaf(f1.z, u1[hh].z, x);
        d[hh] = fmaf(f1.w, u1[hh].w, x);
      }
      const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
      float e[4], g2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float keep = b16 ? d[i + 4] : d[i], send = b16 ? d[i] : d[i + 4];
        e[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float keep = b8 ? e[i + 2] : e[i], send = b8 ? e[i] : e[i + 2];
        g2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      float r = (b4 ? g2[1] : g2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? g2[0] : g2[1], 4);
      r += __shfl_xor_sync(0xffffffffu, r, 2);
      r += __shfl_xor_sync(0xffffffffu, r, 1);
      if ((lane & 3) == 0) sm.prob[p][lane >> 2] = r;  // head index = lane bits 4..2
    }
    __syncthreads();
    // masked softmax over the 100 points, one warp per head (all-masked rows un-mask point 0, map_encoder.py:31)
    {
      const int hh = warp;  // 8 warps = 8 heads
      const uint8_t* pv = pt_valid + (size_t)pid * NP;
      float sc[4];
      bool ok[4];
      int any = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        ok[j] = p < NP && pv[p];
        any |= ok[j];
      }
      any = __any_sync(0xffffffffu, any);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        if (!any && p == 0) ok[j] = true;
        sc[j] = ok[j] ? sm.prob[p][hh] * kLog2e : -INFINITY;
        mx = fmaxf(mx, sc[j]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { sc[j] = ok[j] ? exp2f(sc[j] - mx) : 0.f; sum += sc[j]; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        if (p < NP) sm.prob[p][hh] = sc[j] * inv;
      }
    }
    __syncthreads();
    // pooled[h][d] = sum_p prob[p][h] * feats[p][d].  Thread = (point group pg, column quad cq): one 128-bit read of the
    // tile and two broadcast reads of the weights feed 32 FMAs (thread-per-column needed the two weight reads for 8);
    // the four point groups are then added through the consumed tile.
    {
      const int pg = tid >> 6, cq = tid & 63;
      float accp[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) accp[hh][0] = accp[hh][1] = accp[hh][2] = accp[hh][3] = 0.f;
#pragma unroll 5
      for (int p = pg; p < NP; p += POOL_THREADS / 64) {
        const float4 f = *reinterpret_cast<const float4*>(&sm.feats[s][p][4 * cq]);
        const float4 pa = *reinterpret_cast<const float4*>(&sm.prob[p][0]);
        const float4 pb = *reinterpret_cast<const float4*>(&sm.prob[p][4]);
        const float w[NH] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          accp[hh][0] = fmaf(w[hh], f.x, accp[hh][0]); accp[hh][1] = fmaf(w[hh], f.y, accp[hh][1]);
          accp[hh][2] = fmaf(w[hh], f.z, accp[hh][2]); accp[hh][3] = fmaf(w[hh], f.w, accp[hh][3]);
        }
      }
      __syncthreads();  // every thread is done with the tile: its first 32 KB become the reduction buffer [4][NH][H]
      float* red = &sm.feats[s][0][0];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
        *reinterpret_cast<float4*>(&red[(pg * NH + hh) * H + 4 * cq]) = make_float4(accp[hh][0], accp[hh][1], accp[hh][2], accp[hh][3]);
      __syncthreads();
      for (int i = tid; i < NH * H / 4; i += POOL_THREADS) {
        float4 a = *reinterpret_cast<const float4*>(&red[4 * i]);
#pragma unroll
        for (int k = 1; k < POOL_THREADS / 64; ++k) {
          const float4 b = *reinterpret_cast<const float4*>(&red[k * NH * H + 4 * i]);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4*>(&out[4 * i]) = a;
      }
    }
    __syncthreads();  // stage s and prob[] free for reuse
  }
}

int launch_map_pool(const float* feats, const uint8_t* pt_valid, const uint8_t* poly_valid, const float* U,
                    float* pooled, int n_poly, int n_sm, cudaStream_t st) {
  if (n_poly <= 0) return 0;
  static bool attr_set = false;
  const int smem = (int)sizeof(PoolSmem);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(map_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "map_pool smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = n_poly < n_sm ? n_poly : n_sm;
  map_pool_kernel<<<grid, POOL_THREADS, smem, st>>>(feats, pt_valid, poly_valid, U, pooled, n_poly);
  CS_CHECK_LAUNCH("map_pool");
  return 0;
}

}  // namespace ctrlsim
