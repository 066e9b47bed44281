// Categorical sampling of RTGs and actions (M8/M9) with the explicit sampler of sampler.cuh.
//
// Reference: Policy.process_predicted_rtg (policies/policy.py:108-142): logits[350,3] (bin-major, component-minor)
// + tilt_c * linspace(0,1,350) -> softmax -> multinomial per component; the RTG triple of an agent is drawn ONCE per
// step by the first focal group (in group order) whose context contains it, tilted only if that group serves the agent
// (autoregressive_policy.py:195-207).  Action: softmax(logits / temperature) -> multinomial -> (accel, steer)
// (autoregressive_policy.py:214-240, dataset.py:322-339).
#include "common.cuh"
#include "kernels.h"
#include "model.h"
#include "sampler.cuh"

namespace ctrlsim {

__global__ void sample_rows_kernel(const float* __restrict__ x, int rows, int n, int ld, int stride, uint64_t seed,
                                   const uint32_t* __restrict__ ctr, int* __restrict__ out, int nucleus, double top_p) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const float* row = x + (size_t)r * ld;
  const uint64_t bits = sampler_bits(seed, ctr[r * 4], ctr[r * 4 + 1], ctr[r * 4 + 2], ctr[r * 4 + 3]);
  const auto xf = [&](int i) { return row[(size_t)i * stride]; };
  const int idx = nucleus ? warp_sample_nucleus(n, xf, bits, top_p) : warp_sample(n, xf, bits);
  if ((threadIdx.x & 31) == 0) out[r] = idx;
}

int launch_sample_rows(const float* x, int rows, int n, int ld, int stride, uint64_t seed, const uint32_t* counters,
                       int* out_idx, cudaStream_t st, bool nucleus, double top_p) {
  if (rows <= 0) return 0;
  if (nucleus && n > 1024) return set_error(-2, "sample_rows: nucleus sampling supports at most 1024 categories (got %d)", n);
  sample_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, rows, n, ld, stride, seed, counters, out_idx, nucleus ? 1 : 0, top_p);
  CS_CHECK_LAUNCH("sample_rows");
  return 0;
}

// One warp per (scene, vehicle) of scenes [s0, s1). rtg_logits is chunk-local: row ((group_off[s]+lg-g_base)*A + slot).
__global__ void __launch_bounds__(256)
resolve_rtg_kernel(CtrlSimBatch b, CtrlSimPolicyParams p, int t, int s0, int s1, int g_base, int steps,
                   const float* __restrict__ rtg_logits) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int N = b.max_veh;
  const int s = s0 + w / N, v = w % N;
  if (s >= s1) return;
  int16_t* tr = b.tr_rtg_idx + (((size_t)s * N + v) * steps + t) * 3;
  int16_t* hr = b.hist_rtg + (((size_t)s * N + v) * steps + t) * 3;
  if (v >= b.n_veh[s]) return;
  const int ng = b.n_groups[s];
  int hit_g = -1, hit_k = -1;
  for (int lg = 0; lg < ng && hit_g < 0; ++lg) {
    const int* mem = b.group_members + ((size_t)s * N + lg) * A;
    for (int a0 = 0; a0 < A && hit_g < 0; a0 += 32) {
      const int mv = a0 + lane < A ? mem[a0 + lane] : -1;
      const unsigned m = __ballot_sync(0xffffffffu, mv == v);
      if (m) { hit_g = lg; hit_k = a0 + __ffs(m) - 1; }
    }
  }
  if (hit_g < 0) {  // in no context this step: RTG (0,0,0) is appended (autoregressive_policy.py:246-247) -> bins (0,35,35)
    if (lane == 0) { tr[0] = tr[1] = tr[2] = -1; hr[0] = 0; hr[1] = 35; hr[2] = 35; }
    return;
  }
  const bool tilted = p.tilt_enabled && ((b.group_served[(size_t)s * N + hit_g] >> hit_k) & 1ull);
  const float* row = rtg_logits + ((size_t)(b.group_off[s] + hit_g - g_base) * A + hit_k) * (N_RTG * 3);
  const uint32_t scene = (uint32_t)b.scene_id[s];
  for (int c = 0; c < 3; ++c) {
    const double tilt = tilted ? p.tilt[c] : 0.0;
    const uint64_t bits = sampler_bits(p.seed, scene, (uint32_t)v, (uint32_t)t, (uint32_t)c);
    const int idx = warp_sample(N_RTG, [&](int i) {
      const double lin = i == N_RTG - 1 ? 1.0 : (double)i * (1.0 / (double)(N_RTG - 1));  // np.linspace(0, 1, 350)
      return (float)((double)row[i * 3 + c] + tilt * lin);
    }, bits);
    if (lane == 0) { tr[c] = (int16_t)idx; hr[c] = (int16_t)idx; }
  }
}

int launch_resolve_rtg_range(const CtrlSimBatch& b, const CtrlSimPolicyParams& p, int t, int s0, int s1, int g_base,
                             int steps, const float* rtg_logits, cudaStream_t st) {
  const int warps = (s1 - s0) * b.max_veh;
  if (warps <= 0) return 0;
  resolve_rtg_kernel<<<(warps + 7) / 8, 256, 0, st>>>(b, p, t, s0, s1, g_base, steps, rtg_logits);
  CS_CHECK_LAUNCH("resolve_rtg");
  return 0;
}

// rtg bins of every member slot of groups [g0, g0+ng) at step t -> rtg_new [ng, A, 3] (second-pass rtg tokens)
__global__ void gather_rtg_kernel(CtrlSimBatch b, int g0, int ng, int t, int steps, int* __restrict__ rtg_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng * A) return;
  const int gl = i / A, a = i % A;
  const int g = g0 + gl;
  const int s = b.group_scene[g], lg = b.group_local[g];
  const int N = b.max_veh;
  const int v = b.group_members[((size_t)s * N + lg) * A + a];
  if (v < 0) { rtg_new[i * 3] = rtg_new[i * 3 + 1] = rtg_new[i * 3 + 2] = 0; return; }
  const int16_t* hr = b.hist_rtg + (((size_t)s * N + v) * steps + t) * 3;
  rtg_new[i * 3] = hr[0]; rtg_new[i * 3 + 1] = hr[1]; rtg_new[i * 3 + 2] = hr[2];
}

int launch_gather_rtg_steps(const CtrlSimBatch& b, int g0, int ng, int t, int steps, int* rtg_new, cudaStream_t st) {
  if (ng <= 0) return 0;
  gather_rtg_kernel<<<(ng * A + 255) / 256, 256, 0, st>>>(b, g0, ng, t, steps, rtg_new);
  CS_CHECK_LAUNCH("gather_rtg");
  return 0;
}

// One warp per (group, slot): served members draw their action.
__global__ void __launch_bounds__(256)
sample_actions_kernel(CtrlSimBatch b, CtrlSimPolicyParams p, int g0, int ng, int t, int steps,
                      const float* __restrict__ act_logits, int n_steer, double min_accel, double max_accel,
                      double min_steer, double max_steer) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= ng * A) return;
  const int gl = w / A, a = w % A;
  const int g = g0 + gl;
  const int s = b.group_scene[g], lg = b.group_local[g];
  const int N = b.max_veh;
  if (!((b.group_served[(size_t)s * N + lg] >> a) & 1ull)) return;
  const int v = b.group_members[((size_t)s * N + lg) * A + a];
  const float* row = act_logits + (size_t)w * N_ACT;
  const float temp = p.temperature;
  const uint64_t bits = sampler_bits(p.seed, (uint32_t)b.scene_id[s], (uint32_t)v, (uint32_t)t, 3u);
  const auto xf = [&](int i) { return __fdiv_rn(row[i], temp); };
  const int idx = p.nucleus_sampling ? warp_sample_nucleus(N_ACT, xf, bits, p.nucleus_threshold) : warp_sample(N_ACT, xf, bits);
  if (lane == 0) {
    const int n_acc = N_ACT / n_steer;
    double* na = b.next_action + ((size_t)s * N + v) * 2;
    na[0] = (double)(idx / n_steer) / (double)(n_acc - 1) * (max_accel - min_accel) + min_accel;
    na[1] = (double)(idx % n_steer) / (double)(n_steer - 1) * (max_steer - min_steer) + min_steer;
    b.tr_act_idx[((size_t)s * N + v) * steps + t] = (int16_t)idx;
  }
}

int launch_sample_actions(const CtrlSimBatch& b, const CtrlSimPolicyParams& p, int g0, int ng, int t,
                          const float* act_logits, const ModelCfg& mc, cudaStream_t st) {
  if (ng <= 0) return 0;
  sample_actions_kernel<<<(ng * A + 7) / 8, 256, 0, st>>>(b, p, g0, ng, t, mc.steps, act_logits, mc.n_steer,
                                                         mc.min_accel, mc.max_accel, mc.min_steer, mc.max_steer);
  CS_CHECK_LAUNCH("sample_actions");
  return 0;
}

}  // namespace ctrlsim
