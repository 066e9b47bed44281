// Dense linear layers of the policy network: C[M,N] = act( A[M,K] * W[N,K]^T + bias[N] + table[tidx[m], N] ).
//
// Round-1 arithmetic: fp32 FFMA with fp32 accumulation, so that logits stay within 1e-4 of the fp32 CPU reference
// (nn.Linear in modules/encoder.py, modules/decoder.py, utils/layers.py).  128x128x16 tiles, 256 threads, 8x8
// register micro-tile read as float4 from k-major shared tiles, register-staged double buffering of the global loads.
// Both operands are K-contiguous ("TN" layout: activations row-major, nn.Linear weights [out,in]).
// DESIGN.md lists the tcgen05 (kind::tf32, 3-pass split) replacement as the next step for this kernel.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool RELU>
__global__ void __launch_bounds__(256, 2)
gemm_tn_kernel(const float* __restrict__ Aa, const float* __restrict__ W, const float* __restrict__ bias,
               const float* __restrict__ table, const int* __restrict__ tidx, const int* __restrict__ agather,
               float* __restrict__ C, int M, int N, int K, int lda, int ldw, int ldc, int ldt) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int lrow = tid >> 2;         // 0..63
  const int lk = (tid & 3) << 2;     // 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;

  const float* aptr[2];
  const float* wptr[2];
  bool aok[2], wok[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int m = m0 + lrow + 64 * h;
    aok[h] = m < M;
    int src = aok[h] ? (agather ? agather[m] : m) : 0;
    aptr[h] = Aa + (size_t)src * lda + lk;
    int n = n0 + lrow + 64 * h;
    wok[h] = n < N;
    wptr[h] = W + (size_t)(wok[h] ? n : 0) * ldw + lk;
  }
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = aok[h] ? *reinterpret_cast<const float4*>(aptr[h] + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = wok[h] ? __ldg(reinterpret_cast<const float4*>(wptr[h] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + 64 * h;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* trow = table ? table + (size_t)tidx[m] * ldt : nullptr;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (n + j < N) {
          if (bias) x += __ldg(bias + n + j);
          if (trow) x += __ldg(trow + n + j);
        }
        v[j] = RELU ? fmaxf(x, 0.f) : x;
      }
      float* dst = C + (size_t)m * ldc + n;
      if (vec_ok && n + 3 < N) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) dst[j] = v[j];
      }
    }
  }
}

// CTRLSIM_GEMM=simt forces the FP32 FFMA kernel (A/B testing); the default is the tcgen05 3xTF32 kernel.
static bool use_tc() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CTRLSIM_GEMM");
    mode = (e && std::string(e) == "simt") ? 0 : 1;
  }
  return mode == 1;
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0) return 0;
  if (use_tc() && g.K % 32 == 0) return launch_gemm_tc(g, st);
  if (g.K % BK != 0 || (g.lda & 3) || (g.ldw & 3))
    return set_error(-2, "gemm: K=%d must be a multiple of %d and lda/ldw multiples of 4", g.K, BK);
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
  if (g.relu)
    gemm_tn_kernel<true><<<grid, 256, 0, st>>>(g.A, g.W, g.bias, g.table, g.tidx, g.agather, g.C, g.M, g.N, g.K, g.lda,
                                               g.ldw, g.ldc, g.ldt);
  else
    gemm_tn_kernel<false><<<grid, 256, 0, st>>>(g.A, g.W, g.bias, g.table, g.tidx, g.agather, g.C, g.M, g.N, g.K, g.lda,
                                                g.ldw, g.ldc, g.ldt);
  CS_CHECK_LAUNCH("gemm");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Row LayerNorm (eps 1e-5, biased variance, like torch.nn.LayerNorm) over H=256 with optional residual input and ReLU:
//   Y[m] = act( LN(X[m] + R[m]) * gamma + beta )          one warp per row, 8 values per lane, two-pass variance.
template <bool RELU>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ X, const float* __restrict__ R, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ Y, int M, int ldx, int ldr, int ldy) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* x = X + (size_t)warp * ldx;
  float v[8];
  // lane owns columns 4 lane .. 4 lane + 3 and 128 + 4 lane .. + 3: every warp access is one contiguous 512-byte run
  float4 p0 = *reinterpret_cast<const float4*>(x + lane * 4);
  float4 p1 = *reinterpret_cast<const float4*>(x + H / 2 + lane * 4);
  v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
  if (R) {
    const float* r = R + (size_t)warp * ldr;
    float4 q0 = *reinterpret_cast<const float4*>(r + lane * 4);
    float4 q1 = *reinterpret_cast<const float4*>(r + H / 2 + lane * 4);
    v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + LN_EPS);
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (i < 4 ? 0 : H / 2) + lane * 4 + (i & 3);
    float y = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    o[i] = RELU ? fmaxf(y, 0.f) : y;
  }
  float* y = Y + (size_t)warp * ldy;
  *reinterpret_cast<float4*>(y + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(y + H / 2 + lane * 4) = make_float4(o[4], o[5], o[6], o[7]);
}

int launch_layernorm(const float* X, const float* R, const float* gamma, const float* beta, float* Y, int M, int ldx,
                     int ldr, int ldy, bool relu, cudaStream_t st) {
  if (M <= 0) return 0;
  const int blocks = (M + 7) / 8;
  if (relu) layernorm_kernel<true><<<blocks, 256, 0, st>>>(X, R, gamma, beta, Y, M, ldx, ldr, ldy);
  else layernorm_kernel<false><<<blocks, 256, 0, st>>>(X, R, gamma, beta, Y, M, ldx, ldr, ldy);
  CS_CHECK_LAUNCH("layernorm");
  return 0;
}

}  // namespace ctrlsim
