// Polyline encoder front end, fused from the raw points (modules/map_encoder.py:34-45): per point
// MLPLayer(3 -> 256 -> 256), then nn.MultiheadAttention with ONE learned seed query per polyline over its 100 points.
//
// Everything after the first MLP layer is linear in the point features, so it commutes with the attention-weighted
// sum over points (the weights of a head sum to 1; the reference un-masks key 0 of an all-masked polyline):
//     feats_p      = W3 h_p + b3,           h_p = ReLU(LN(W0 x_p + b0))            (road_pts_encoder)
//     score_{p,h}  = feats_p . U_h + const  = h_p . (W3^T U_h) + const'            (const' cancels in the softmax)
//     pooled_h     = sum_p prob_{p,h} feats_p = W3 (sum_p prob_{p,h} h_p) + b3
// so this kernel evaluates h_p on the fly from the three raw floats of a point (x, y, exists), scores it against the
// host-folded U2_h = W3^T U_h (derived.pool_U2), soft-maxes over the polyline's points and accumulates
// P_h = sum_p prob_{p,h} h_p; the value / output projections AND W3 are applied afterwards to the 8 pooled vectors by
// one GEMM with the host-folded matrix derived.pool_W2 (ctrlsim_b200/model.py).  Neither the [points, 256] hidden layer
// nor the [points, 256] feature tensor - 20 MB each per focal group - ever exists: the kernel reads 1.2 KB and writes
// 8 KB per polyline (0.44 MB per group where the unfused chain moved 82 MB) and is FP32-pipe bound.  The unfused
// HBM-streaming kernel (map_pool_kernel, attention.cu) stays exported as ctrlsim_map_pool.
//
// One CTA (256 threads) per polyline at a time, persistent over polylines:
//   scores  warp per point, lane owns 8 channels (the layout and operation order of small_mlp1_kernel: h_p is bit-identical
//           to the unfused path), 8 per-head partial dot products reduced with a transposing butterfly;
//   softmax warp per head over the 100 points;
//   pooling thread = (point group, 4 channels): h_p recomputed from the stored LN statistics (same operation order),
//           32 FMAs per 2 broadcast reads of the weights; the four point groups are added through shared memory.
#include "common.cuh"
#include "kernels.h"
#include "model.h"

namespace ctrlsim {

constexpr int ME_THREADS = 256;
constexpr float kMeLog2e = 1.4426950408889634f;

struct MeSmem {
  float pts[NP][4];        // x, y, exists, -
  float stat[NP][2];       // LayerNorm mean, rstd of the point's hidden vector
  float prob[NP][NH];      // scores, then softmax weights (point-major: float4 reads)
  float red[ME_THREADS / 64][NH][H];
};

__global__ void __launch_bounds__(ME_THREADS, 1)
map_encode_pool_kernel(const float* __restrict__ map_pts, const float* __restrict__ W0, const float* __restrict__ b0,
                       const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ U2,
                       const uint8_t* __restrict__ poly_valid, float* __restrict__ pooled, int n_poly) {
  extern __shared__ __align__(16) unsigned char me_raw[];
  MeSmem& sm = *reinterpret_cast<MeSmem*>(me_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- constants of the scores pass: lane owns channels 4 lane .. + 3 and 128 + 4 lane .. + 3 --------------------------
  float w[8][3], bb[8], gm[8], bt[8];
  float4 u0[NH], u1[NH];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = (i < 4 ? 0 : H / 2) + lane * 4 + (i & 3);
    bb[i] = __ldg(b0 + o); gm[i] = __ldg(gamma + o); bt[i] = __ldg(beta + o);
#pragma unroll
    for (int k = 0; k < 3; ++k) w[i][k] = __ldg(W0 + o * 3 + k);
  }
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    u0[hh] = __ldg(reinterpret_cast<const float4*>(U2 + hh * H + lane * 4));
    u1[hh] = __ldg(reinterpret_cast<const float4*>(U2 + hh * H + 128 + lane * 4));
  }
  // ---- constants of the pooling pass: thread owns channels 4 cq .. + 3 for the points p = pg (mod 4) -------------------
  const int pg = tid >> 6, cq = tid & 63;

  for (int pid = blockIdx.x; pid < n_poly; pid += gridDim.x) {
    float* out = pooled + (size_t)pid * (NH * H);
    if (!poly_valid[pid]) {  // padded polyline: never read as a key downstream, keep it finite
      for (int i = tid; i < NH * H; i += ME_THREADS) out[i] = 0.f;
      continue;
    }
    __syncthreads();  // the previous polyline's readers of pts / stat / prob / red are done
    if (tid < NP) {
      const float* p = map_pts + ((size_t)pid * NP + tid) * 3;
      sm.pts[tid][0] = p[0]; sm.pts[tid][1] = p[1]; sm.pts[tid][2] = p[2];
    }
    __syncthreads();
    // ---- scores: two points per warp iteration (independent shuffle chains overlap) -----------------------------------
    for (int p0 = warp; p0 < NP; p0 += 2 * (ME_THREADS / 32)) {
      const int pq[2] = {p0, p0 + ME_THREADS / 32};
      const bool two = pq[1] < NP;
      float v[2][8], s[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int p = (z == 0 || two) ? pq[z] : pq[0];
        const float x0 = sm.pts[p][0], x1 = sm.pts[p][1], x2 = sm.pts[p][2];
        s[z] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float acc = bb[i];
          acc = fmaf(w[i][0], x0, acc); acc = fmaf(w[i][1], x1, acc); acc = fmaf(w[i][2], x2, acc);
          v[z][i] = acc;
          s[z] += acc;
        }
      }
      float mean[2], rstd[2], q[2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s[0] += __shfl_xor_sync(0xffffffffu, s[0], o); s[1] += __shfl_xor_sync(0xffffffffu, s[1], o); }
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        mean[z] = s[z] * (1.0f / H);
        q[z] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[z][i] - mean[z]; q[z] = fmaf(d, d, q[z]); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { q[0] += __shfl_xor_sync(0xffffffffu, q[0], o); q[1] += __shfl_xor_sync(0xffffffffu, q[1], o); }
      float r[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        rstd[z] = rsqrtf(q[z] * (1.0f / H) + LN_EPS);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[z][i] = fmaxf((v[z][i] - mean[z]) * rstd[z] * gm[i] + bt[i], 0.f);
        float d[NH];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          float x = v[z][0] * u0[hh].x;
          x = fmaf(v[z][1], u0[hh].y, x); x = fmaf(v[z][2], u0[hh].z, x); x = fmaf(v[z][3], u0[hh].w, x);
          x = fmaf(v[z][4], u1[hh].x, x); x = fmaf(v[z][5], u1[hh].y, x); x = fmaf(v[z][6], u1[hh].z, x);
          d[hh] = fmaf(v[z][7], u1[hh].w, x);
        }
        // transposing butterfly: 4 + 2 + 1 + 2 shuffles instead of 8 x 5
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
        float rr = (b4 ? g2[1] : g2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? g2[0] : g2[1], 4);
        rr += __shfl_xor_sync(0xffffffffu, rr, 2);
        rr += __shfl_xor_sync(0xffffffffu, rr, 1);
        r[z] = rr;
      }
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        if (z == 1 && !two) break;
        if ((lane & 3) == 0) sm.prob[pq[z]][lane >> 2] = r[z];  // head index = lane bits 4..2
        if (lane == 0) { sm.stat[pq[z]][0] = mean[z]; sm.stat[pq[z]][1] = rstd[z]; }
      }
    }
    __syncthreads();
    // ---- masked softmax over the points, one warp per head (all-masked rows un-mask point 0, map_encoder.py:31) -----
    {
      const int hh = warp;  // 8 warps = 8 heads
      float sc[4];
      bool ok[4];
      int any = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        ok[j] = p < NP && sm.pts[p < NP ? p : 0][2] != 0.f;
        any |= ok[j];
      }
      any = __any_sync(0xffffffffu, any);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = lane + 32 * j;
        if (!any && p == 0) ok[j] = true;
        sc[j] = ok[j] ? sm.prob[p][hh] * kMeLog2e : -INFINITY;
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
    // ---- pooling: P[h][c] = sum_p prob[p][h] h_p[c] -------------------------------------------------------------------
    {
      float wp[4][3], bp[4], gp[4], tp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = 4 * cq + j;
        bp[j] = __ldg(b0 + o); gp[j] = __ldg(gamma + o); tp[j] = __ldg(beta + o);
#pragma unroll
        for (int k = 0; k < 3; ++k) wp[j][k] = __ldg(W0 + o * 3 + k);
      }
      float accp[NH][4];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) accp[hh][0] = accp[hh][1] = accp[hh][2] = accp[hh][3] = 0.f;
#pragma unroll 5
      for (int p = pg; p < NP; p += ME_THREADS / 64) {
        const float4 pt = *reinterpret_cast<const float4*>(&sm.pts[p][0]);
        const float2 st = *reinterpret_cast<const float2*>(&sm.stat[p][0]);
        const float4 pa = *reinterpret_cast<const float4*>(&sm.prob[p][0]);
        const float4 pb = *reinterpret_cast<const float4*>(&sm.prob[p][4]);
        float f[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float acc = bp[j];
          acc = fmaf(wp[j][0], pt.x, acc); acc = fmaf(wp[j][1], pt.y, acc); acc = fmaf(wp[j][2], pt.z, acc);
          f[j] = fmaxf((acc - st.x) * st.y * gp[j] + tp[j], 0.f);
        }
        const float wgt[NH] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          accp[hh][0] = fmaf(wgt[hh], f[0], accp[hh][0]); accp[hh][1] = fmaf(wgt[hh], f[1], accp[hh][1]);
          accp[hh][2] = fmaf(wgt[hh], f[2], accp[hh][2]); accp[hh][3] = fmaf(wgt[hh], f[3], accp[hh][3]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
        *reinterpret_cast<float4*>(&sm.red[pg][hh][4 * cq]) = make_float4(accp[hh][0], accp[hh][1], accp[hh][2], accp[hh][3]);
      __syncthreads();
      for (int i = tid; i < NH * H / 4; i += ME_THREADS) {
        float4 a = *reinterpret_cast<const float4*>(&sm.red[0][0][0] + 4 * i);
#pragma unroll
        for (int k = 1; k < ME_THREADS / 64; ++k) {
          const float4 b = *reinterpret_cast<const float4*>(&sm.red[k][0][0] + 4 * i);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        *reinterpret_cast<float4*>(&out[4 * i]) = a;
      }
    }
  }
}

int launch_map_encode_pool(const float* map_pts, const MlpW& pts_mlp, const float* U2, const uint8_t* poly_valid,
                           float* pooled, int n_poly, int n_sm, cudaStream_t st) {
  if (n_poly <= 0) return 0;
  static bool attr_set = false;
  const int smem = (int)sizeof(MeSmem);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(map_encode_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "map_encode_pool smem attr: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = n_poly < n_sm ? n_poly : n_sm;  // one persistent CTA per SM (235 registers per thread)
  map_encode_pool_kernel<<<grid, ME_THREADS, smem, st>>>(map_pts, pts_mlp.w0, pts_mlp.b0, pts_mlp.lnw, pts_mlp.lnb, U2,
                                                        poly_valid, pooled, n_poly);
  CS_CHECK_LAUNCH("map_encode_pool");
  return 0;
}

}  // namespace ctrlsim
