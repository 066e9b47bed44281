// Tokenisation (T2/T3) and embedding front-ends (M2/M3 input stages) of the policy network.
//
// Reference being replaced (all float64 numpy on the host, once per focal group per step):
//   AutoregressivePolicy.get_data          policies/autoregressive_policy.py:51-165 (window slice, RTG clip-normalise)
//   select_relevant_agents / padding       datasets/rl_waymo/dataset.py:278-319
//   discretize_actions / discretize_rtgs   datasets/rl_waymo/dataset.py:365-387   (np.round = half-to-even = rint)
//   normalize_scene + apply_se2_transform  datasets/rl_waymo/dataset.py:390-428, utils/geometry.py:14-47
//   Encoder.forward embeddings             modules/encoder.py:99-153 ; MLPLayer utils/layers.py:10-15
// SE2 math is done in fp64 exactly like numpy and narrowed to fp32 where the reference calls .float().
#include "common.cuh"
#include "kernels.h"
#include "model.h"

namespace ctrlsim {

__device__ __forceinline__ double py_mod(double a, double b) {  // Python float %: result takes the sign of b
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}
__device__ __forceinline__ double angle_sub_d(double cur, double tgt) {  // utils/geometry.py:3-19
  const double two_pi = 6.283185307179586, pi = 3.141592653589793;
  double d = py_mod(tgt - cur, two_pi);
  if (d > pi) d = -(two_pi - d);
  return d;
}

// ---------------------------------------------------------------------------------------------------------------
// tokenize_agents: one block per compact group, one thread per (window step, slot).
__global__ void __launch_bounds__(256)
tokenize_agents_kernel(CtrlSimBatch b, int g0, int t, int n_t, int tok_first, int steps, TokenBufs tk, double min_accel,
                       double max_accel, double min_steer, double max_steer, int n_steer, int rtg_mode, int dt_model,
                       double lo0, double lo1, double lo2, double hi0, double hi1, double hi2) {
  // n_t token steps starting at window index tok_first (0 = whole window; the incremental decode of the prefix cache
  // tokenises only the last two window steps); the normalisation frame is always the focal pose at window index 0.
  const int gl = blockIdx.x, g = g0 + gl;
  const int s = b.group_scene[g], lg = b.group_local[g];
  const int N = b.max_veh;
  const int* members = b.group_members + ((size_t)s * N + lg) * A;
  const int focal = b.group_focal[(size_t)s * N + lg];
  const int t0 = t < T ? 0 : t - (T - 1);
  const double* hs = b.hist_state + (size_t)s * N * steps * 8;
  const double* fo = hs + ((size_t)focal * steps + t0) * 8;
  const double yaw_f = fo[4];
  const double rot = 1.5707963267948966 + (yaw_f > 0 ? -1.0 : (yaw_f < 0 ? 1.0 : 0.0)) * fabs(yaw_f);
  const double cr = cos(rot), sr = sin(rot);
  const double tx = fo[0], ty = fo[1];
  if (threadIdx.x == 0) {
    tk.frame[gl * 4 + 0] = tx; tk.frame[gl * 4 + 1] = ty; tk.frame[gl * 4 + 2] = rot; tk.frame[gl * 4 + 3] = 0.0;
  }
  for (int i = threadIdx.x; i < n_t * A; i += blockDim.x) {
    const int tw = i / A, a = i - tw * A;
    const int v = members[a];
    const size_t row = ((size_t)gl * n_t + tw) * A + a;
    float* f = tk.feat_state + row * 12;
    if (v < 0) {  // padded slot: existence 0, everything it produces is masked (dataset.py:283-288)
#pragma unroll
      for (int k = 0; k < 12; ++k) f[k] = 0.f;
      tk.exist[row] = 0;
      tk.act_idx[row] = 0;
      tk.rtg_idx[row * 3 + 0] = tk.rtg_idx[row * 3 + 1] = tk.rtg_idx[row * 3 + 2] = 0;
      if (dt_model) tk.rtg_val[row * 3 + 0] = tk.rtg_val[row * 3 + 1] = tk.rtg_val[row * 3 + 2] = 0.f;
    } else {
      const double* st = hs + ((size_t)v * steps + t0 + tok_first + tw) * 8;
      const double dx = st[0] - tx, dy = st[1] - ty;
      f[0] = (float)(cr * dx + (-sr) * dy);
      f[1] = (float)(sr * dx + cr * dy);
      f[2] = (float)(cr * st[2] + (-sr) * st[3]);
      f[3] = (float)(sr * st[2] + cr * st[3]);
      f[4] = (float)angle_sub_d(st[4], -rot);
      f[5] = (float)st[5];
      f[6] = (float)st[6];
      f[7] = 0.f; f[8] = 1.f; f[9] = 0.f; f[10] = 0.f; f[11] = 0.f;  // one-hot "vehicle" (utils/data.py:326-328)
      tk.exist[row] = st[7] != 0.0;
      const double* ac = b.hist_action + ((size_t)s * N * steps + (size_t)v * steps + t0 + tok_first + tw) * 2;
      const double a0 = (fmin(fmax(ac[0], min_accel), max_accel) - min_accel) / (max_accel - min_accel);
      const double a1 = (fmin(fmax(ac[1], min_steer), max_steer) - min_steer) / (max_steer - min_steer);
      tk.act_idx[row] = (int)(rint(a0 * (N_ACT / n_steer - 1)) * n_steer + rint(a1 * (n_steer - 1)));
      if (rtg_mode == 0) {
        const int16_t* rt = b.hist_rtg + ((size_t)s * N * steps + (size_t)v * steps + t0 + tok_first + tw) * 3;
        tk.rtg_idx[row * 3 + 0] = rt[0]; tk.rtg_idx[row * 3 + 1] = rt[1]; tk.rtg_idx[row * 3 + 2] = rt[2];
      } else {
        // tracked (real-time) RTGs: clip-normalise in float64 (autoregressive_policy.py:73-78), then either discretise
        // with np.round (dataset.py:382-387) or hand the value to the decision transformer as float32 (encoder.py:95)
        const double* rr = b.rt_rtg + ((size_t)s * N * steps + (size_t)v * steps + t0 + tok_first + tw) * 3;
        const double lo[3] = {lo0, lo1, lo2}, hi[3] = {hi0, hi1, hi2};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double x = (fmin(fmax(rr[c], lo[c]), hi[c]) - lo[c]) / (hi[c] - lo[c]);
          if (dt_model) { tk.rtg_val[row * 3 + c] = (float)x; tk.rtg_idx[row * 3 + c] = 0; }
          else tk.rtg_idx[row * 3 + c] = (int)rint(x * (N_RTG - 1));
        }
      }
    }
  }
  for (int tw = threadIdx.x; tw < n_t; tw += blockDim.x)
    tk.ts[gl * n_t + tw] = (t0 + tok_first + tw <= t) ? t0 + tok_first + tw : 0;  // policy.timesteps[0, window] (policy.py:81)
  for (int a = threadIdx.x; a < A; a += blockDim.x) {
    const int v = members[a];
    float* gf = tk.goal_feat + ((size_t)gl * A + a) * 5;
    if (v < 0) {
      // padded goals are zeros *before* normalize_scene and get transformed like everything else; masked later
      const double dx = 0.0 - tx, dy = 0.0 - ty;
      gf[0] = (float)(cr * dx + (-sr) * dy); gf[1] = (float)(sr * dx + cr * dy);
      gf[2] = 0.f; gf[3] = 0.f; gf[4] = (float)angle_sub_d(0.0, -rot);
    } else {
      const double* go = b.goal + ((size_t)s * N + v) * 4;  // x, y, heading, speed
      const double gvx = go[3] * cos(go[2]), gvy = go[3] * sin(go[2]);
      const double dx = go[0] - tx, dy = go[1] - ty;
      gf[0] = (float)(cr * dx + (-sr) * dy); gf[1] = (float)(sr * dx + cr * dy);
      gf[2] = (float)(cr * gvx + (-sr) * gvy); gf[3] = (float)(sr * gvx + cr * gvy);
      gf[4] = (float)angle_sub_d(go[2], -rot);
    }
  }
}

// tokenize_map: one block per group. Keeps the max_polylines polylines whose farthest valid point is nearest to the
// focal agent, in ascending order of that distance (dataset.py:417-421), else pads (dataset.py:423-427).
// `sel` (optional) lists the chunk-local groups whose map must be (re)built; block i writes map slot i.
__global__ void __launch_bounds__(256)
tokenize_map_kernel(CtrlSimBatch b, int g0, TokenBufs tk, const int* __restrict__ sel) {
  __shared__ double key[1024];
  __shared__ int idx[1024];
  const int gsrc = sel ? sel[blockIdx.x] : blockIdx.x;  // chunk-local group whose frame is used
  const int gl = blockIdx.x, g = g0 + gsrc;             // gl: output map slot
  const int s = b.group_scene[g];
  const int Pm = b.max_poly, np_s = b.n_poly[s];
  const double tx = tk.frame[gsrc * 4 + 0], ty = tk.frame[gsrc * 4 + 1], rot = tk.frame[gsrc * 4 + 2];
  const double cr = cos(rot), sr = sin(rot);
  const double* rxy = b.road_xy + (size_t)s * Pm * NP * 2;
  const uint8_t* rv = b.road_valid + (size_t)s * Pm * NP;
  int nsel = np_s;
  if (np_s > P) {
    int n2 = 1;
    while (n2 < np_s) n2 <<= 1;
    for (int p = threadIdx.x; p < n2; p += blockDim.x) {
      double dmax = 1e300;
      if (p < np_s) {
        dmax = 0.0;
        for (int k = 0; k < NP; ++k) {
          const double dx = rxy[((size_t)p * NP + k) * 2] - tx, dy = rxy[((size_t)p * NP + k) * 2 + 1] - ty;
          const double x = cr * dx + (-sr) * dy, y = sr * dx + cr * dy;
          const double d = sqrt(x * x + y * y) * (rv[(size_t)p * NP + k] ? 1.0 : 0.0);
          dmax = fmax(dmax, d);
        }
      }
      key[p] = dmax;
      idx[p] = p;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool up = (i & k) == 0;
            const bool gt = key[i] > key[ixj] || (key[i] == key[ixj] && idx[i] > idx[ixj]);
            if (gt == up) {
              double tkv = key[i]; key[i] = key[ixj]; key[ixj] = tkv;
              int ti = idx[i]; idx[i] = idx[ixj]; idx[ixj] = ti;
            }
          }
        }
        __syncthreads();
      }
    nsel = P;
  } else {
    for (int p = threadIdx.x; p < P; p += blockDim.x) idx[p] = p;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < P * NP; i += blockDim.x) {
    const int j = i / NP, k = i - j * NP;
    float* o = tk.map_pts + ((size_t)gl * P * NP + i) * 3;
    if (j < nsel) {
      const int p = idx[j];
      const double dx = rxy[((size_t)p * NP + k) * 2] - tx, dy = rxy[((size_t)p * NP + k) * 2 + 1] - ty;
      o[0] = (float)(cr * dx + (-sr) * dy);
      o[1] = (float)(sr * dx + cr * dy);
      o[2] = rv[(size_t)p * NP + k] ? 1.f : 0.f;
    } else {
      o[0] = o[1] = o[2] = 0.f;
    }
  }
  for (int j = threadIdx.x; j < P; j += blockDim.x)
    tk.map_type[gl * P + j] = j < nsel ? (int)b.road_type[(size_t)s * Pm + idx[j]] : -1;
}

int launch_tokenize(const CtrlSimBatch& b, int g0, int ng, int t, int n_t, const TokenBufs& tk, const ModelCfg& mc,
                    cudaStream_t st, const int* map_sel, int n_map, int tok_first, int rtg_mode) {
  if (ng <= 0) return 0;
  if (b.max_poly > 1024) return set_error(-2, "tokenize: at most 1024 polylines per scene (got %d)", b.max_poly);
  if (rtg_mode != 0 && !b.rt_rtg) return set_error(-2, "tokenize: tracked RTGs requested but the batch has no rt_rtg array");
  if (mc.dt_model && rtg_mode == 0) return set_error(-2, "tokenize: the decision transformer needs tracked RTGs (rtg_mode 1)");
  tokenize_agents_kernel<<<ng, 256, 0, st>>>(b, g0, t, n_t, tok_first, mc.steps, tk, mc.min_accel, mc.max_accel,
                                             mc.min_steer, mc.max_steer, mc.n_steer, rtg_mode, mc.dt_model,
                                             mc.rtg_min[0], mc.rtg_min[1], mc.rtg_min[2], mc.rtg_max[0], mc.rtg_max[1],
                                             mc.rtg_max[2]);
  CS_CHECK_LAUNCH("tokenize_agents");
  const int nm = map_sel ? n_map : ng;
  if (nm > 0) {
    tokenize_map_kernel<<<nm, 256, 0, st>>>(b, g0, tk, map_sel);
    CS_CHECK_LAUNCH("tokenize_map");
  }
  return 0;
}

// Reference MotionData layout -> internal token buffers (parity-test entry ctrlsim_forward_tokens).
__global__ void convert_tokens_kernel(int G, int n_t, const float* __restrict__ agent_states,
                                      const float* __restrict__ agent_types, const float* __restrict__ goals,
                                      const int* __restrict__ actions, const void* __restrict__ rtgs_any,
                                      int rtgs_are_float, const int* __restrict__ timesteps,
                                      const float* __restrict__ road_points, const int* __restrict__ road_types,
                                      TokenBufs tk) {
  const int* rtgs = reinterpret_cast<const int*>(rtgs_any);
  const float* rtgs_f = reinterpret_cast<const float*>(rtgs_any);
  const int gl = blockIdx.x;
  for (int i = threadIdx.x; i < n_t * A; i += blockDim.x) {
    const int tw = i / A, a = i - tw * A;
    const size_t row = ((size_t)gl * n_t + tw) * A + a;
    const float* st = agent_states + (((size_t)gl * A + a) * T + tw) * 8;
    float* f = tk.feat_state + row * 12;
    for (int k = 0; k < 7; ++k) f[k] = st[k];
    for (int k = 0; k < 5; ++k) f[7 + k] = agent_types[((size_t)gl * A + a) * 5 + k];
    tk.exist[row] = st[7] != 0.f;
    tk.act_idx[row] = actions[((size_t)gl * A + a) * T + tw];
    for (int c = 0; c < 3; ++c) {
      const size_t src = (((size_t)gl * A + a) * T + tw) * 3 + c;
      if (rtgs_are_float) { tk.rtg_val[row * 3 + c] = rtgs_f[src]; tk.rtg_idx[row * 3 + c] = 0; }
      else tk.rtg_idx[row * 3 + c] = rtgs[src];
    }
  }
  for (int tw = threadIdx.x; tw < n_t; tw += blockDim.x) tk.ts[gl * n_t + tw] = timesteps[gl * T + tw];
  for (int i = threadIdx.x; i < A * 5; i += blockDim.x) tk.goal_feat[(size_t)gl * A * 5 + i] = goals[(size_t)gl * A * 5 + i];
  for (size_t i = threadIdx.x; i < (size_t)P * NP * 3; i += blockDim.x)
    tk.map_pts[(size_t)gl * P * NP * 3 + i] = road_points[(size_t)gl * P * NP * 3 + i];
  for (int j = threadIdx.x; j < P; j += blockDim.x) tk.map_type[gl * P + j] = road_types[gl * P + j];
}

int launch_convert_tokens(int G, int n_t, const float* agent_states, const float* agent_types, const float* goals,
                          const int* actions, const void* rtgs, bool rtgs_are_float, const int* timesteps,
                          const float* road_points, const int* road_types, const TokenBufs& tk, cudaStream_t st) {
  convert_tokens_kernel<<<G, 256, 0, st>>>(G, n_t, agent_states, agent_types, goals, actions, rtgs, rtgs_are_float ? 1 : 0,
                                           timesteps, road_points, road_types, tk);
  CS_CHECK_LAUNCH("convert_tokens");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// First layer of an MLPLayer with a tiny input (3, 5 or 12 features): y = ReLU(LN(W1 x + b1)) ; one warp per row,
// lane owns 8 of the 256 outputs (utils/layers.py:10-15 mlp.0 -> mlp.1 -> mlp.2).
template <int DIN>
__global__ void __launch_bounds__(256)
small_mlp1_kernel(const float* __restrict__ X, const float* __restrict__ W1, const float* __restrict__ b1,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ Y, size_t M) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  // this lane's 8 output channels: weights, bias and LN affine live in registers for all rows the warp processes
  // lane owns channels 4 lane .. 4 lane + 3 and 128 + 4 lane .. + 3: every warp store is one contiguous 512-byte run
  float w[8][DIN], bb[8], gm[8], bt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = (i < 4 ? 0 : H / 2) + lane * 4 + (i & 3);
    bb[i] = __ldg(b1 + o); gm[i] = __ldg(gamma + o); bt[i] = __ldg(beta + o);
#pragma unroll
    for (int k = 0; k < DIN; ++k) w[i][k] = __ldg(W1 + o * DIN + k);
  }
  for (size_t row = warp0; row < M; row += nwarps) {
    float x[DIN];
#pragma unroll
    for (int k = 0; k < DIN; ++k) x[k] = X[row * DIN + k];
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float acc = bb[i];
#pragma unroll
      for (int k = 0; k < DIN; ++k) acc = fmaf(w[i][k], x[k], acc);
      v[i] = acc;
      s += acc;
    }
    const float mean = warp_sum(s) * (1.0f / H);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + LN_EPS);
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o8[i] = fmaxf((v[i] - mean) * rstd * gm[i] + bt[i], 0.f);
    float* y = Y + row * H + lane * 4;
    *reinterpret_cast<float4*>(y) = make_float4(o8[0], o8[1], o8[2], o8[3]);
    *reinterpret_cast<float4*>(y + H / 2) = make_float4(o8[4], o8[5], o8[6], o8[7]);
  }
}

int launch_small_mlp1(int din, const float* X, const MlpW& w, float* Y, size_t M, cudaStream_t st) {
  if (M == 0) return 0;
  // persistent-ish grid: each warp keeps its weight slice in registers and strides over rows
  size_t want = (M + 7) / 8;
  const unsigned blocks = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  if (din == 3) small_mlp1_kernel<3><<<blocks, 256, 0, st>>>(X, w.w0, w.b0, w.lnw, w.lnb, Y, M);
  else if (din == 5) small_mlp1_kernel<5><<<blocks, 256, 0, st>>>(X, w.w0, w.b0, w.lnw, w.lnb, Y, M);
  else if (din == 12) small_mlp1_kernel<12><<<blocks, 256, 0, st>>>(X, w.w0, w.b0, w.lnw, w.lnb, Y, M);
  else return set_error(-2, "small_mlp1: unsupported input width %d", din);
  CS_CHECK_LAUNCH("small_mlp1");
  return 0;
}

// Per-polyline bookkeeping for the pooling kernel: point validity bytes, polyline validity, type index rows.
__global__ void map_flags_kernel(const float* __restrict__ map_pts, uint8_t* __restrict__ pt_valid,
                                 uint8_t* __restrict__ poly_valid, int n_poly) {
  const int pid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pid >= n_poly) return;
  int any = 0;
  for (int k = lane; k < NP; k += 32) {
    const uint8_t v = map_pts[((size_t)pid * NP + k) * 3 + 2] != 0.f;
    pt_valid[(size_t)pid * NP + k] = v;
    any |= v;
  }
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) poly_valid[pid] = (uint8_t)any;
}

int launch_map_flags(const float* map_pts, uint8_t* pt_valid, uint8_t* poly_valid, int n_poly, cudaStream_t st) {
  if (n_poly <= 0) return 0;
  map_flags_kernel<<<(n_poly + 7) / 8, 256, 0, st>>>(map_pts, pt_valid, poly_valid, n_poly);
  CS_CHECK_LAUNCH("map_flags");
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// assemble_tokens: one warp per token row (g, tw, a, k). Builds the pre-LN embedding, applies the existence mask,
// stores the initial-state memory token (pre-LN, window step 0; encoder.py:111-112,136) and LayerNorms into X.
//   state:  sg[row] (= W_sg[:, :H] s_emb + W_sg[:, H:] g_emb + b_sg) + E_ts[ts] + E_id[a]
//   rtg:    Rg[i0] + Rv[i1] + Rr[i2] + b_rtg + E_ts + E_id   (R* = embedding tables folded through embed_rtg)
//   action: E_act[idx] + E_ts + E_id
// Decision transformer (ew.dt): rows are ordered (rtg, state, action) (encoder.py:139-142) and the rtg embedding is
//   r0 L[0] + r1 L[1] + r2 L[2] + b  with L = the three Linear(1, H) maps folded through embed_rtg (model.py).
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(int n_rows, int n_t, const float* __restrict__ sg, TokenBufs tk, EmbedW ew,
                       float* __restrict__ X, float* __restrict__ mem /* [G, MEM, H] */) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int kpos = r % KT;
  const int k = ew.dt ? (kpos == 0 ? 1 : (kpos == 1 ? 0 : 2)) : kpos;  // token kind: 0 state, 1 rtg, 2 action
  const int sa = r / KT;                 // (g, tw, a) flat
  const int a = sa % A;
  const int gtw = sa / A;
  const int tw = gtw % n_t, gl = gtw / n_t;
  const int ts = tk.ts[gl * n_t + tw];
  const float ex = tk.exist[sa] ? 1.f : 0.f;
  float v[8];
  // lane owns channels 4 lane .. + 3 and 128 + 4 lane .. + 3 (fully coalesced 512-byte warp accesses)
#define CH(i) (((i) < 4 ? 0 : H / 2) + lane * 4 + ((i) & 3))
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __ldg(ew.ts + ts * H + CH(i)) + __ldg(ew.id + a * H + CH(i));
  if (k == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = sg[(size_t)sa * H + CH(i)] + v[i];
  } else if (k == 1 && ew.dt) {
    const float r0 = tk.rtg_val[sa * 3], r1 = tk.rtg_val[sa * 3 + 1], r2 = tk.rtg_val[sa * 3 + 2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (((r0 * __ldg(ew.rtg_lin + CH(i)) + r1 * __ldg(ew.rtg_lin + H + CH(i))) +
               r2 * __ldg(ew.rtg_lin + 2 * H + CH(i))) + __ldg(ew.rtg_bias + CH(i))) + v[i];
  } else if (k == 1) {
    const int i0 = tk.rtg_idx[sa * 3], i1 = tk.rtg_idx[sa * 3 + 1], i2 = tk.rtg_idx[sa * 3 + 2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = (((__ldg(ew.rtg_goal + i0 * H + CH(i)) + __ldg(ew.rtg_veh + i1 * H + CH(i))) +
               __ldg(ew.rtg_road + i2 * H + CH(i))) + __ldg(ew.rtg_bias + CH(i))) + v[i];
  } else {
    const int ia = tk.act_idx[sa];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(ew.act + ia * H + CH(i)) + v[i];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] *= ex; s += v[i]; }
  if (k == 0 && tw == 0 && mem) {
    float* m = mem + ((size_t)gl * MEM + P + a) * H + lane * 4;
    *reinterpret_cast<float4*>(m) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(m + H / 2) = make_float4(v[4], v[5], v[6], v[7]);
  }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + LN_EPS);
  float* x = X + (size_t)r * H + lane * 4;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * __ldg(ew.ln_w + CH(i)) + __ldg(ew.ln_b + CH(i));
  *reinterpret_cast<float4*>(x) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(x + H / 2) = make_float4(o[4], o[5], o[6], o[7]);
}

#undef CH
int launch_assemble_tokens(int G, int n_t, const float* sg, const TokenBufs& tk, const EmbedW& ew, float* X, float* mem,
                           cudaStream_t st) {
  const int n_rows = G * n_t * TOK_T;
  if (n_rows <= 0) return 0;
  assemble_tokens_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(n_rows, n_t, sg, tk, ew, X, mem);
  CS_CHECK_LAUNCH("assemble_tokens");
  return 0;
}

// Second-pass rtg tokens of the current window step `ti` with freshly sampled bins: rows (g, a) -> LN'd embedding.
__global__ void __launch_bounds__(256)
assemble_rtg_rows_kernel(int n_rows, int n_t, int ti, const int* __restrict__ rtg_new /* [G,A,3] */, TokenBufs tk,
                         EmbedW ew, float* __restrict__ Xr) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const int a = r % A, gl = r / A;
  const int sa = (gl * n_t + ti) * A + a;
  const int ts = tk.ts[gl * n_t + ti];
  const float ex = tk.exist[sa] ? 1.f : 0.f;
  const int i0 = rtg_new[r * 3], i1 = rtg_new[r * 3 + 1], i2 = rtg_new[r * 3 + 2];
  // lane owns channels 4 lane .. + 3 and 128 + 4 lane .. + 3 (fully coalesced 512-byte warp accesses)
#define CH(i) (((i) < 4 ? 0 : H / 2) + lane * 4 + ((i) & 3))
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float e = __ldg(ew.ts + ts * H + CH(i)) + __ldg(ew.id + a * H + CH(i));
    v[i] = ((((__ldg(ew.rtg_goal + i0 * H + CH(i)) + __ldg(ew.rtg_veh + i1 * H + CH(i))) +
              __ldg(ew.rtg_road + i2 * H + CH(i))) + __ldg(ew.rtg_bias + CH(i))) + e) * ex;
    s += v[i];
  }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + LN_EPS);
  float* x = Xr + (size_t)r * H + lane * 4;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * __ldg(ew.ln_w + CH(i)) + __ldg(ew.ln_b + CH(i));
  *reinterpret_cast<float4*>(x) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(x + H / 2) = make_float4(o[4], o[5], o[6], o[7]);
}

#undef CH
int launch_assemble_rtg_rows(int G, int n_t, int ti, const int* rtg_new, const TokenBufs& tk, const EmbedW& ew,
                             float* Xr, cudaStream_t st) {
  const int n_rows = G * A;
  if (n_rows <= 0) return 0;
  assemble_rtg_rows_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>(n_rows, n_t, ti, rtg_new, tk, ew, Xr);
  CS_CHECK_LAUNCH("assemble_rtg_rows");
  return 0;
}

// RTG bins of window step ti as the tokeniser produced them (tracked RTGs, CtRL-Sim network) -> rtg_new [G, A, 3]
__global__ void gather_rtg_tokens_kernel(int n, int n_t, int ti, const int* __restrict__ rtg_idx, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = i % 3, ga = i / 3, a = ga % A, gl = ga / A;
  out[i] = rtg_idx[(((size_t)gl * n_t + ti) * A + a) * 3 + c];
}
int launch_gather_rtg_tokens(int G, int n_t, int ti, const TokenBufs& tk, int* rtg_new, cudaStream_t st) {
  const int n = G * A * 3;
  if (n <= 0) return 0;
  gather_rtg_tokens_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n_t, ti, tk.rtg_idx, rtg_new);
  CS_CHECK_LAUNCH("gather_rtg_tokens");
  return 0;
}

// Memory tokens: rows [g, 0..P) <- polyline embeddings, key padding mask = !valid (encoder.py:155-166).
// `slot` (optional): group gl reads block slot[gl] of poly_emb / poly_valid (the per-focal map cache) instead of gl.
__global__ void build_memory_kernel(int G, const float* __restrict__ poly_emb, const uint8_t* __restrict__ poly_valid,
                                    const int* __restrict__ slot, TokenBufs tk, int n_t, float* __restrict__ mem,
                                    uint8_t* __restrict__ pad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = (size_t)G * P * (H / 4);
  if (i < n4) {
    const size_t row = i / (H / 4);
    const int c = (int)(i % (H / 4)) * 4;
    const size_t gl = row / P, j = row % P;
    const size_t src = slot ? (size_t)slot[gl] : gl;
    *reinterpret_cast<float4*>(mem + (gl * MEM + j) * H + c) =
        *reinterpret_cast<const float4*>(poly_emb + (src * P + j) * H + c);
  }
  if (i < (size_t)G * MEM) {
    const int gl = (int)(i / MEM), j = (int)(i % MEM);
    const size_t src = slot ? (size_t)slot[gl] : (size_t)gl;
    pad[i] = j < P ? !poly_valid[src * P + j] : !tk.exist[((size_t)gl * n_t + 0) * A + (j - P)];
  }
}

int launch_build_memory(int G, const float* poly_emb, const uint8_t* poly_valid, const int* slot, const TokenBufs& tk,
                        int n_t, float* mem, uint8_t* pad, cudaStream_t st) {
  if (G <= 0) return 0;
  const size_t n = (size_t)G * P * (H / 4);
  build_memory_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(G, poly_emb, poly_valid, slot, tk, n_t, mem, pad);
  CS_CHECK_LAUNCH("build_memory");
  return 0;
}

// Map cache fill: block i of the freshly encoded polyline embeddings -> cache block dst[i].
__global__ void scatter_map_kernel(int n, const float* __restrict__ emb, const uint8_t* __restrict__ valid,
                                   const int* __restrict__ dst, float* __restrict__ cache_emb,
                                   uint8_t* __restrict__ cache_valid) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)P * (H / 4);
  if (i >= (size_t)n * per) return;
  const size_t blk = i / per, r = i % per;
  const size_t d = (size_t)dst[blk];
  reinterpret_cast<float4*>(cache_emb)[d * per + r] = reinterpret_cast<const float4*>(emb)[i];
  if (r < P) cache_valid[d * P + r] = valid[blk * P + r];
}

int launch_scatter_map(int n, const float* emb, const uint8_t* valid, const int* dst, float* cache_emb,
                       uint8_t* cache_valid, cudaStream_t st) {
  if (n <= 0) return 0;
  const size_t tot = (size_t)n * P * (H / 4);
  scatter_map_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, emb, valid, dst, cache_emb, cache_valid);
  CS_CHECK_LAUNCH("scatter_map");
  return 0;
}

__global__ void store_kv_kernel(const float* __restrict__ src, int ld_src, int rows, size_t n4, float* __restrict__ dst,
                                int dst_row0) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)(i % (2 * H / 4)) * 4;
  const size_t r = i / (2 * H / 4);
  const size_t g = r / rows, rr = r % rows;
  *reinterpret_cast<float4*>(dst + (g * L + dst_row0 + rr) * (2 * H) + c) =
      *reinterpret_cast<const float4*>(src + r * ld_src + c);
}
int launch_store_kv(const float* src, int ld_src, int G, int rows, float* dst, int dst_row0, cudaStream_t st) {
  const size_t n4 = (size_t)G * rows * (2 * H / 4);
  if (n4 == 0) return 0;
  store_kv_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(src, ld_src, rows, n4, dst, dst_row0);
  CS_CHECK_LAUNCH("store_kv");
  return 0;
}
int launch_copy_bytes(const uint8_t* src, uint8_t* dst, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  cudaError_t e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return set_error(-5, "copy_bytes: %s", cudaGetErrorString(e));
  return 0;
}

// Row-index tables for the gather GEMMs of the heads: state rows (k=0) or rtg rows (k=1) of window step ti.
__global__ void make_row_index_kernel(int G, int n_t, int ti, int k, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * A) return;
  const int gl = i / A, a = i % A;
  out[i] = ((gl * n_t + ti) * A + a) * KT + k;
}
int launch_make_row_index(int G, int n_t, int ti, int k, int* out, cudaStream_t st) {
  if (G <= 0) return 0;
  make_row_index_kernel<<<(G * A + 255) / 256, 256, 0, st>>>(G, n_t, ti, k, out);
  CS_CHECK_LAUNCH("make_row_index");
  return 0;
}

// Y[i] = X[idx[i]] for H-wide rows (state rows of the current step for the pruned last decoder layer)
__global__ void gather_rows_kernel(int n, const float* __restrict__ X, const int* __restrict__ idx, float* __restrict__ Y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * (H / 4)) return;
  const int r = i / (H / 4), c = (i % (H / 4)) * 4;
  *reinterpret_cast<float4*>(Y + (size_t)r * H + c) = *reinterpret_cast<const float4*>(X + (size_t)idx[r] * H + c);
}
int launch_gather_rows(int n, const float* X, const int* idx, float* Y, cudaStream_t st) {
  if (n <= 0) return 0;
  gather_rows_kernel<<<(n * (H / 4) + 255) / 256, 256, 0, st>>>(n, X, idx, Y);
  CS_CHECK_LAUNCH("gather_rows");
  return 0;
}

// tidx for the goal-part table add of embed_state_goal: row (g, tw, a) -> g*A + a ; and polyline type rows.
__global__ void make_goal_index_kernel(int n, int n_t, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = i % A, gl = i / (A * n_t);
  out[i] = gl * A + a;
}
int launch_make_goal_index(int G, int n_t, int* out, cudaStream_t st) {
  const int n = G * n_t * A;
  if (n <= 0) return 0;
  make_goal_index_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, n_t, out);
  CS_CHECK_LAUNCH("make_goal_index");
  return 0;
}
__global__ void clamp_type_index_kernel(int n, const int* __restrict__ in, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] < 0 ? 8 : in[i];  // row 8 of the type tables = all-(-1) padding input (dataset.py:425)
}
int launch_clamp_type_index(int n, const int* in, int* out, cudaStream_t st) {
  if (n <= 0) return 0;
  clamp_type_index_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, in, out);
  CS_CHECK_LAUNCH("clamp_type_index");
  return 0;
}

}  // namespace ctrlsim
