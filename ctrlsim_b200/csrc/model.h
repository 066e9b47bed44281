// Internal structures shared by the tokeniser, the network forward and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctrlsim_b200.h"

namespace ctrlsim {

struct ModelCfg {
  int steps = 90, hist_steps = 10, n_steer = 50;
  float dt = 0.1f;      // what Simulation.step receives (float)
  double dt_d = 0.1;    // what the python evaluator divides by (float64)
  double agent_dist = 60.0;
  double min_accel = -10, max_accel = 10, min_steer = -0.7, max_steer = 0.7;
  double pos_tol = 1.0, heading_tol = 0.3, speed_tol = 1.0, goal_dist_scaling = 0.2, reward_scaling = 1.0;
  int contacts = 1;     // Box2D contact response between vehicles (sim_contacts.cuh); 0 = contact-free subset
  int dt_model = 0;     // decision-transformer variant of the network (CtrlSimConfig.decision_transformer)
  double rtg_min[3] = {0, -10, -10}, rtg_max[3] = {10, 90, 90};  // clip-normalisation of tracked RTGs (pos, veh, road)
};

struct MlpW {  // utils/layers.py MLPLayer: Linear - LayerNorm - ReLU - Linear
  const float *w0 = nullptr, *b0 = nullptr, *lnw = nullptr, *lnb = nullptr, *w3 = nullptr, *b3 = nullptr;
};
struct LnW { const float *w = nullptr, *b = nullptr; };
struct MhaW { const float *in_w = nullptr, *in_b = nullptr, *out_w = nullptr, *out_b = nullptr; };
struct EncLayerW { MhaW sa; const float *l1w, *l1b, *l2w, *l2b; LnW n1, n2; };
struct DecLayerW { MhaW sa, ca; const float *l1w, *l1b, *l2w, *l2b; LnW n1, n2, n3; };

struct EmbedW {  // what assemble_tokens needs
  const float *ts, *id, *act, *rtg_goal, *rtg_veh, *rtg_road, *rtg_bias, *ln_w, *ln_b;
  // decision transformer (dt = 1): the three Linear(1, H) RTG embeddings folded through embed_rtg: rtg_lin [3, H] holds
  // embed_rtg.weight[:, cH:(c+1)H] @ embed_rtg_c.weight[:, 0]; rtg_bias then also carries the folded Linear(1, H) biases;
  // token order (rtg, state, action) instead of (state, rtg, action)
  const float* rtg_lin = nullptr;
  int dt = 0;
};

struct ModelWeights {
  // map encoder (modules/map_encoder.py)
  MlpW road_pts;            // 3 -> 256 -> 256
  const float* pool_U;      // derived [8,256]: W_k,h^T (q_h * d_h^-0.5)   (unfused chain, ctrlsim_map_pool)
  const float* pool_W;      // derived [256, 8*256]: out_proj . blockdiag(W_v)
  const float* pool_b;      // derived [256]: out_proj . b_v + out_proj.bias
  // the same three with the second layer of road_pts_encoder (W3, b3) folded in (map_encoder.cu): the pooling runs on
  // the hidden layer h = ReLU(LN(W0 x + b0)) and W3 is applied to the 8 pooled vectors instead of the 100 points
  const float* pool_U2;     // derived [8,256]: W3^T U_h
  const float* pool_W2;     // derived [256, 8*256]: block h = pool_W_h W3
  const float* pool_b2;     // derived [256]: pool_b + sum_h pool_W_h b3
  LnW map_n1, map_n2;
  MlpW map_feats;           // 256 -> 256 -> 256
  const float* type_tab2;   // derived [9,256]: road_road_type_encoder.mlp.0 applied to the type half + its bias
  MlpW rr;                  // road_road_type_encoder: w0 is the [256, 512] matrix (first 256 columns used by the GEMM)
  // embeddings (modules/encoder.py)
  MlpW embed_state, embed_goal;
  const float *sg_w, *sg_b;  // embed_state_goal [256, 512]
  EmbedW emb;
  EncLayerW enc[2];
  DecLayerW dec[4];
  MlpW head_action, head_rtg;  // head_rtg is absent (null) in the decision-transformer variant
  bool dt = false;             // decision transformer: state token at position 1 of an agent's step, action head on state rows
};

struct TokenBufs {  // internal token representation of a chunk of groups (time-major)
  float* feat_state;   // [G, n_t, A, 12]
  uint8_t* exist;      // [G, n_t, A]
  float* goal_feat;    // [G, A, 5]
  int* act_idx;        // [G, n_t, A]
  int* rtg_idx;        // [G, n_t, A, 3]
  float* rtg_val;      // [G, n_t, A, 3] clip-normalised continuous RTGs (decision transformer)
  int* ts;             // [G, n_t]
  float* map_pts;      // [G, P, NP, 3]
  int* map_type;       // [G, P]
  double* frame;       // [G, 4] tx, ty, rot
};

// launchers (tokens.cu)
// map_sel (optional, device): chunk-local groups whose polyline tokens are rebuilt (n_map of them, into map slots
// 0..n_map-1); nullptr = every group.
// rtg_mode 0: RTG tokens are the sampled bins of hist_rtg; 1: the tracked series rt_rtg, clip-normalised, then
// discretised (CtRL-Sim network) or kept continuous (mc.dt_model)
int launch_tokenize(const CtrlSimBatch& b, int g0, int ng, int t, int n_t, const TokenBufs& tk, const ModelCfg& mc,
                    cudaStream_t st, const int* map_sel = nullptr, int n_map = 0, int tok_first = 0, int rtg_mode = 0);
// rtgs: int32 bins, or - rtgs_are_float - float32 continuous values (decision transformer)
int launch_convert_tokens(int G, int n_t, const float* agent_states, const float* agent_types, const float* goals,
                          const int* actions, const void* rtgs, bool rtgs_are_float, const int* timesteps,
                          const float* road_points, const int* road_types, const TokenBufs& tk, cudaStream_t st);
int launch_gather_rtg_tokens(int G, int n_t, int ti, const TokenBufs& tk, int* rtg_new, cudaStream_t st);
int launch_small_mlp1(int din, const float* X, const MlpW& w, float* Y, size_t M, cudaStream_t st);
int launch_map_encode_pool(const float* map_pts, const MlpW& pts_mlp, const float* U2, const uint8_t* poly_valid,
                           float* pooled, int n_poly, int n_sm, cudaStream_t st);  // map_encoder.cu
int launch_map_flags(const float* map_pts, uint8_t* pt_valid, uint8_t* poly_valid, int n_poly, cudaStream_t st);
int launch_assemble_tokens(int G, int n_t, const float* sg, const TokenBufs& tk, const EmbedW& ew, float* X, float* mem,
                           cudaStream_t st);
int launch_assemble_rtg_rows(int G, int n_t, int ti, const int* rtg_new, const TokenBufs& tk, const EmbedW& ew,
                             float* Xr, cudaStream_t st);
int launch_build_memory(int G, const float* poly_emb, const uint8_t* poly_valid, const int* slot, const TokenBufs& tk,
                        int n_t, float* mem, uint8_t* pad, cudaStream_t st);
int launch_scatter_map(int n, const float* emb, const uint8_t* valid, const int* dst, float* cache_emb,
                       uint8_t* cache_valid, cudaStream_t st);
// dst[g, dst_row0 + r, 0..2H) = src[(g * rows + r) * ld_src + 0..2H)   (K | V rows into a prefix-cache slot [G, L, 2H])
int launch_store_kv(const float* src, int ld_src, int G, int rows, float* dst, int dst_row0, cudaStream_t st);
int launch_copy_bytes(const uint8_t* src, uint8_t* dst, size_t n, cudaStream_t st);
int launch_make_row_index(int G, int n_t, int ti, int k, int* out, cudaStream_t st);
int launch_make_goal_index(int G, int n_t, int* out, cudaStream_t st);
int launch_gather_rows(int n, const float* X, const int* idx, float* Y, cudaStream_t st);
int launch_clamp_type_index(int n, const int* in, int* out, cudaStream_t st);

// sim.cu
int set_trig_mode(int glibc);  // sim.cu: 1 = sinf / cosf by glibc's algorithm (glibc_trig.h), 0 = fp64 and round once
int launch_sim_reset(const CtrlSimBatch& b, const ModelCfg& mc, cudaStream_t st);
int launch_observe(const CtrlSimBatch& b, int t, const ModelCfg& mc, cudaStream_t st);
int launch_dense_reward(const CtrlSimBatch& b, const CtrlSimRewardParams& rp, int t, const ModelCfg& mc, cudaStream_t st);
int launch_plan_groups(const CtrlSimBatch& b, int t, const ModelCfg& mc, int* n_total, cudaStream_t st);
int launch_sim_step(const CtrlSimBatch& b, int t, const ModelCfg& mc, cudaStream_t st);
int launch_metrics(const CtrlSimBatch& b, const ModelCfg& mc, double* out_scene, long long* out_hist, cudaStream_t st);
int launch_geom_poly_poly(const float* xy1, int n1, const float* xy2, int n2, int* out, cudaStream_t st);
int launch_geom_poly_seg(const float* xy, int n, const float* seg, int* out, cudaStream_t st);

// sample.cu
int launch_sample_rows(const float* x, int rows, int n, int ld, int stride, uint64_t seed, const uint32_t* counters,
                       int* out_idx, cudaStream_t st, bool nucleus = false, double top_p = 1.0);
int launch_sample_actions(const CtrlSimBatch& b, const CtrlSimPolicyParams& p, int g0, int ng, int t,
                          const float* act_logits, const ModelCfg& mc, cudaStream_t st);

}  // namespace ctrlsim
