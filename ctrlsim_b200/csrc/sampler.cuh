// Device statement of "sampler contract v1" (see oracle/sampler.py for the normative text).
//
// Replaces the reference's per-vehicle softmax + torch.multinomial (policies/policy.py:122-127,
// policies/autoregressive_policy.py:233-236) by an order-independent, bit-reproducible draw:
//   x_i -> d_i = max(x_i - max x, -80) -> e_i = exp_spec(d_i) -> w_i = floor(e_i * 2^30)  (integers)
//   r = Philox4x32-10(key = seed, counter = (scene, agent, step, component)) ; target = mulhi64(r, sum w)
//   idx = first i whose inclusive integer prefix exceeds target.
// exp_spec uses only separately rounded fp32 multiplies/adds (no FMA), so the weights are bit-identical to numpy's.
#pragma once
#include <stdint.h>

namespace ctrlsim {

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ uint64_t sampler_bits(uint64_t seed, uint32_t scene, uint32_t agent, uint32_t step,
                                                 uint32_t comp) {
  uint32_t c[4] = {scene, agent, step, comp};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (uint64_t)c[0] | ((uint64_t)c[1] << 32);
}

__device__ __forceinline__ float exp_spec(float d) {
  const float k = rintf(__fmul_rn(d, 0x1.715476p+0f));
  float r = __fsub_rn(d, __fmul_rn(k, 0x1.62e4p-1f));
  r = __fsub_rn(r, __fmul_rn(k, 0x1.7f7d1cp-20f));
  float p = 0x1.6c16c2p-10f;  // fp32(1/720), fp32(1/120), fp32(1/24), fp32(1/6): same bits as oracle/sampler.py
  p = __fadd_rn(__fmul_rn(p, r), 0x1.111112p-7f);
  p = __fadd_rn(__fmul_rn(p, r), 0x1.555556p-5f);
  p = __fadd_rn(__fmul_rn(p, r), 0x1.555556p-3f);
  p = __fadd_rn(__fmul_rn(p, r), 0.5f);
  p = __fadd_rn(__fmul_rn(p, r), 1.0f);
  p = __fadd_rn(__fmul_rn(p, r), 1.0f);
  return ldexpf(p, (int)k);
}

__device__ __forceinline__ uint64_t sampler_weight(float x, float xmax) {
  const float d = fmaxf(__fsub_rn(x, xmax), -80.0f);
  return (uint64_t)floorf(__fmul_rn(exp_spec(d), 1073741824.0f));
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp draws one index from n categories. xval(i) returns the fp32 pre-softmax value x_i of category i.
template <class XF>
__device__ __forceinline__ int warp_sample(int n, XF xval, uint64_t bits) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, xval(i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  uint64_t tot = 0;
  for (int i = lane; i < n; i += 32) tot += sampler_weight(xval(i), mx);
  tot = warp_sum_u64(tot);
  const uint64_t target = __umul64hi(bits, tot);
  uint64_t base = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    uint64_t w = i < n ? sampler_weight(xval(i), mx) : 0;
    uint64_t inc = w;  // inclusive scan over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, i < n && base + inc > target);
    if (hit) return i0 + __ffs(hit) - 1;
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
  return n - 1;  // unreachable: the last inclusive prefix equals tot > target
}

// Nucleus (top-p) variant for n <= 1024 categories (oracle/sampler.py, "Nucleus"): the kept set is the shortest prefix,
// in (weight descending, index ascending) order, that reaches p * total.  No sort is needed: with
// S_gt(v) = sum of the weights > v, the kept weight values are those v with S_gt(v) < thr (monotone in v, found by
// bisection), and the ties at the smallest kept value are kept in index order while S_gt + v * rank < thr.
template <class XF>
__device__ __forceinline__ int warp_sample_nucleus(int n, XF xval, uint64_t bits, double p) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, xval(i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  uint32_t w[32];  // category 32 q + lane lives in w[q]; weights are at most 2^30
  uint64_t tot = 0;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const int i = 32 * q + lane;
    w[q] = i < n ? (uint32_t)sampler_weight(xval(i), mx) : 0u;
    tot += w[q];
  }
  tot = warp_sum_u64(tot);
  const double thr = p * (double)tot;
  // smallest v with S_gt(v) < thr; S_gt(2^30) = 0, so hi always qualifies unless thr <= 0 (then only the arg-max is kept)
  uint32_t lo = 0, hi = 1u << 30;
  uint64_t s_gt_hi = 0;
  if (thr > 0.0) {
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      uint64_t sgt = 0;
#pragma unroll
      for (int q = 0; q < 32; ++q) sgt += w[q] > mid ? w[q] : 0u;
      sgt = warp_sum_u64(sgt);
      if ((double)sgt < thr) hi = mid; else lo = mid + 1;
    }
    uint64_t sgt = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) sgt += w[q] > hi ? w[q] : 0u;
    s_gt_hi = warp_sum_u64(sgt);
  }
  const uint32_t vmin = hi;  // weights > vmin are kept entirely; ties at vmin partially, in index order
  // number of ties kept: ranks r = 0, 1, ... while S_gt + vmin * r < thr (rank 0 always: it is what made vmin qualify;
  // with thr <= 0, vmin = 2^30 and the first arg-max tie is the single kept category)
  uint64_t base = 0, kept_total;
  uint32_t ties_seen = 0;
  {
    uint64_t kt = 0;
    uint32_t seen = 0;
    for (int q = 0; q < 32; ++q) {
      const bool tie = w[q] == vmin && vmin > 0u && (32 * q + lane) < n;
      const unsigned tb = __ballot_sync(0xffffffffu, tie);
      const uint32_t rank = seen + __popc(tb & ((1u << lane) - 1u));
      const bool keep = w[q] > vmin || (tie && (rank == 0u || (double)(s_gt_hi + (uint64_t)vmin * rank) < thr));
      kt += keep ? w[q] : 0u;
      seen += __popc(tb);
    }
    kept_total = warp_sum_u64(kt);
  }
  const uint64_t target = __umul64hi(bits, kept_total);
  for (int q = 0; q < 32; ++q) {
    const int i = 32 * q + lane;
    const bool tie = w[q] == vmin && vmin > 0u && i < n;
    const unsigned tb = __ballot_sync(0xffffffffu, tie);
    const uint32_t rank = ties_seen + __popc(tb & ((1u << lane) - 1u));
    const bool keep = w[q] > vmin || (tie && (rank == 0u || (double)(s_gt_hi + (uint64_t)vmin * rank) < thr));
    ties_seen += __popc(tb);
    uint64_t inc = keep ? w[q] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, i < n && base + inc > target);
    if (hit) return 32 * q + __ffs(hit) - 1;
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
  return n - 1;  // unreachable
}

}  // namespace ctrlsim
