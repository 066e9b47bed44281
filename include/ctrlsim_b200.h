/* ctrlsim_b200 - C ABI of the B200-native CtRL-Sim closed-loop rollout hot path.
 *
 * The reference exposes this path through Python objects only: the pybind11 simulator binding
 * (nocturne/pybind11/src/{simulation,scenario,object,vehicle}.cc) and the Policy / PolicyEvaluator classes
 * (policies/policy.py:8-154, policies/autoregressive_policy.py:10-274, evaluators/policy_evaluator.py:27-44,426-595).
 * This header is the flat boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): plain pointers and
 * sizes, no torch types.  Unless a parameter says "host", every pointer is a DEVICE pointer owned by the caller and
 * every call is asynchronous on the given cudaStream_t (passed as void*).  Calls return 0 or a negative status;
 * ctrlsim_last_error() gives the message of the last failure on the calling thread.  A handle is bound to one GPU
 * and may be used from one host thread at a time.
 */
#ifndef CTRLSIM_B200_H
#define CTRLSIM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTRLSIM_ABI_VERSION 6
#define CTRLSIM_MAX_VEH 64 /* vehicles per scene supported by the grouping kernel (bitmask width) */

/* Model / episode geometry. The kernels are specialised at compile time: libctrlsim_b200.so to the reference defaults
 * (cfgs/model/base.yaml:1-9, cfgs/dataset/waymo/base.yaml:4,38-43, cfgs/config.yaml:44-46), libctrlsim_b200_wide.so
 * (same sources, -DCTRLSIM_WIDE, same ABI) to max_agents = 64 / max_polylines = 256 - one focal group per 64-vehicle /
 * 256-polyline scene. ctrlsim_create of either library rejects any other geometry. */
typedef struct CtrlSimConfig {
  int32_t abi_version;
  int32_t hidden_dim, num_heads, dim_feedforward, enc_layers, dec_layers;       /* 256, 8, 1024, 2, 4 */
  int32_t max_agents, context_len, max_polylines, pts_per_polyline;              /* 24, 32, 200, 100 (wide: 64, 32, 256, 100) */
  int32_t n_action_bins, n_steer_bins, n_rtg_bins;                               /* 1000, 50, 350 */
  int32_t steps, history_steps;                                                  /* 90, 10 */
  float dt;                                                                      /* 0.1 */
  double agent_dist_threshold;                                                   /* 60.0 */
  double min_accel, max_accel, min_steer, max_steer;                             /* -10, 10, -0.7, 0.7 */
  double pos_tol, heading_tol, speed_tol, goal_dist_scaling, reward_scaling;     /* rew_cfg, cfgs/config.yaml:72-90 */
  /* ABI 6: the decision-transformer variant of the network (cfgs/model/dt.yaml, modules/encoder.py:27-30,116-120,139-142,
   * utils/train_utils.py:85-88): continuous RTG inputs through Linear(1, H), token order (rtg, state, action), the action
   * head reads the STATE rows, no RTG head. 0 = the CtRL-Sim network. */
  int32_t decision_transformer;
  int32_t reserved0;
  double rtg_min[3], rtg_max[3];  /* clip-normalisation of tracked RTGs: pos, veh, road (cfgs/dataset/waymo/base.yaml:27-36) */
} CtrlSimConfig;

/* A batch of S scenes resident in HBM (struct of arrays; N = max_veh, Pm = max_poly, E = max_seg, T1 = steps+1).
 * "static" arrays are written once by the host loader; "state" arrays are owned by the library between calls. */
typedef struct CtrlSimBatch {
  int32_t n_scenes, max_veh, max_poly, max_seg;
  /* ---- static scene data (replaces Scenario::LoadScenario, nocturne/cpp/src/scenario.cc:207-264,893-1057) ---- */
  const int64_t* scene_id;    /* [S] global scene index (sampler counter word 0)                                  */
  const int32_t* n_veh;       /* [S]                                                                               */
  const float* veh_len;       /* [S,N]                                                                             */
  const float* veh_wid;       /* [S,N]                                                                             */
  const double* gt;           /* [S,N,T1,4] log-replay targets x, y, heading, speed: the expert states          */
                              /* (utils/sim.py:20-65, float32-valued) or a scripted track (float64, CAT adversary)  */
  const uint8_t* gt_valid;    /* [S,N,T1]                                                                          */
  const double* goal;         /* [S,N,4] goal x, y, heading, speed (evaluators/evaluator.py:60-76)                 */
  const double* goal_norm;    /* [S,N] initial distance to goal (evaluator.py:79-84)                               */
  const uint8_t* evaluated;   /* [S,N] 1 = policy-controlled                                                       */
  const int32_t* eval_order;  /* [S,N] evaluated vehicle ids by descending GT length, -1 padded                    */
  const double* road_xy;      /* [S,Pm,100,2] polyline points, world frame                                         */
  const uint8_t* road_valid;  /* [S,Pm,100]                                                                        */
  const int8_t* road_type;    /* [S,Pm] index of the one-hot road type, -1 = padding                               */
  const int32_t* n_poly;      /* [S]                                                                               */
  const float* segs;          /* [S,E,4] road_edge segments x0,y0,x1,y1 (scenario.cc:1037-1042)                    */
  const int32_t* n_seg;       /* [S]                                                                               */
  /* ---- simulator state (replaces the Box2D bodies + Nocturne objects) --------------------------------------- */
  float* body;                /* [S,16,N] px,py,cx,cy,lcx,lcy,ang,vx,vy,om,sleep_t,thr,brk,steer,awake,pad         */
  float* obj;                 /* [S,4,N]  x, y, heading, speed as reported by Vehicle::Step                        */
  uint8_t* coll;              /* [S,2,N]  collision_type_veh / collision_type_edge flags of the current step       */
  /* ---- policy state (replaces Policy.reset/update_state buffers, policies/policy.py:45-105) ------------------- */
  double* hist_state;         /* [S,N,steps,8] x,y,vx,vy,yaw,len,wid,exist                                         */
  double* hist_action;        /* [S,N,steps,2] applied accel, steer                                                */
  int16_t* hist_rtg;          /* [S,N,steps,3] RTG bin indices ((0,35,35) where nothing was sampled)               */
  uint64_t* relevant;         /* [S,N] sticky context membership bitmask (relevant_agent_idxs)                     */
  double* next_action;        /* [S,N,2] next_acceleration / next_steering                                         */
  /* ---- trace: everything vehicle_data_dict records (evaluators/policy_evaluator.py:69-96) --------------------- */
  float* tr_pos;              /* [S,N,T1,2] */
  float* tr_vel;              /* [S,N,T1,2] */
  float* tr_heading;          /* [S,N,T1]   */
  uint8_t* tr_exist;          /* [S,N,T1]   */
  double* tr_action;          /* [S,N,T1,2] */
  float* tr_reward;           /* [S,N,T1,8] */
  double* tr_nearest;         /* [S,N,T1,2] simulated / ground-truth nearest-vehicle distance */
  int16_t* tr_rtg_idx;        /* [S,N,steps,3] sampled RTG bins, -1 = not sampled this step */
  int16_t* tr_act_idx;        /* [S,N,steps]   sampled action bin, -1 = none */
  /* ---- per-step focal groups (autoregressive_policy.py:96-138), rebuilt by ctrlsim_plan_groups --------------- */
  int32_t* n_groups;          /* [S]                                   */
  int32_t* group_off;         /* [S+1] exclusive scan of n_groups      */
  int32_t* group_focal;       /* [S,N]      scene-local group -> focal vehicle */
  int32_t* group_members;     /* [S,N,max_agents] vehicle ids ascending, -1 padded */
  uint64_t* group_served;     /* [S,N]      bit k = member slot k is served by this group */
  int32_t* group_scene;       /* [S*N] compact group -> scene          */
  int32_t* group_local;       /* [S*N] compact group -> scene-local id */
  /* ---- Box2D contact state of the vehicle bodies (simulator state, owned by the library) ------------------------ */
  float* cstate;              /* [S, 4 + 8*N + 20*128] per scene: header, proxy AABBs, mass data, contact manifolds with */
                              /* their warm-start impulses (ctrlsim_b200/csrc/sim_contacts.cuh); NULL = contact-free sim */
  /* ---- ABI 6: real-time rewards (policies with real_time_rewards, e.g. the DT baseline of cfgs/policy/dt.yaml) ----- */
  const double* edge_xy;      /* [S,Ep,2] points of the scene's road_edge polylines, concatenated; float32-valued like */
                              /* RoadLine::geometry_points() (evaluators/evaluator.py:143-158, utils/sim.py:67-73)      */
  const int32_t* edge_off;    /* [S,Pe+1] first point of polyline k; edge_off[s][n_edge[s]] = number of points          */
  const int32_t* n_edge;      /* [S] road_edge polylines of the scene                                                   */
  const double* rtg_init;     /* [S,N,3] un-normalised RTGs (pos, veh, road) at t = 0: preproc_data['rtgs'][:, 0, (0,3,4)] */
  double* rt_rtg;             /* [S,N,steps,3] policy state: the tracked RTG series (policy_evaluator.py:123-149)        */
  double* tr_dense;           /* [S,N,T1,3] trace: dense reward goal / veh-veh / veh-edge (evaluator.py:106-140)         */
  int32_t max_edge_pts, max_edge_poly; /* Ep, Pe */
} CtrlSimBatch;

/* Sampling / control knobs of AutoregressivePolicy (cfgs/policy/ctrl_sim.yaml:6-11). */
typedef struct CtrlSimPolicyParams {
  uint64_t seed;
  double tilt[3];        /* goal, veh_veh, veh_edge */
  float temperature;
  int32_t tilt_enabled;
  int32_t nucleus_sampling;   /* 0 / 1: top-p filtering of the action distribution (autoregressive_policy.py:216-230) */
  double nucleus_threshold;   /* p, cfgs/policy/ctrl_sim.yaml:11 */
  /* ABI 6: 0 = RTGs are predicted (RTG head + sampling, predict_rtgs); 1 = RTGs are given: the tracked series rt_rtg of
   * a real_time_rewards policy (policies/policy.py:93-95) - no RTG head, no RTG sampling. The DT network needs 1. */
  int32_t rtg_mode;
  int32_t reserved0;
} CtrlSimPolicyParams;

/* Constants of the dense reward (cfgs/dataset/waymo/base.yaml:18-25,47-49) and how the RTG series starts
 * (policy_evaluator.py:124-143). */
typedef struct CtrlSimRewardParams {
  double max_veh_veh_distance, dist_to_road_edge_scaling_factor;                 /* 15, 15 */
  double veh_veh_collision_rew_multiplier, veh_edge_collision_rew_multiplier;    /* 10, 10 */
  double pos_goal_shaped_min, pos_goal_shaped_max, pos_target_achieved_rew_multiplier; /* 0, 0.2, 10 */
  int32_t remove_shaped_goal, remove_shaped_veh_reward, remove_shaped_edge_reward;     /* 1, 0, 0 */
  int32_t return_mode;  /* 0: rtg_init; 1: max_return (10, 90, 90); 2: min_return (evaluated vehicles (0, -10, -10));
                         * 3: use_rtg = False - rt_rtg stays (0, 0, 0), the dense reward is still traced */
} CtrlSimRewardParams;

typedef struct CtrlSim CtrlSim;

const char* ctrlsim_last_error(void);
int ctrlsim_abi_version(void);
int ctrlsim_create(const CtrlSimConfig* cfg, CtrlSim** out);
void ctrlsim_destroy(CtrlSim* h);

/* Register one weight tensor (fp32, device) under its reference state-dict name (ctrlsim_b200/weights.py) or a
 * "derived.*" name (host-side folding done by ctrlsim_b200/model.py). The handle keeps the pointer, not a copy. */
int ctrlsim_load_weights(CtrlSim* h, const char* name, const float* ptr, int64_t count);
int ctrlsim_finalize_weights(CtrlSim* h);

/* Bytes of scratch ctrlsim_policy_step needs for `max_groups` focal groups processed together. */
int64_t ctrlsim_workspace_bytes(const CtrlSim* h, int32_t max_groups);

/* Optional per-focal cache of the polyline encoder's output (M2). While a group's 32-step window still starts at
 * t = 0 (steps 0..31) its normalisation frame - the focal agent's pose at window index 0, dataset.py:390-428 - does not
 * move, so the encoder output is identical at every step; with a cache attached ctrlsim_policy_step computes it once
 * per (scene, focal) and episode and re-reads it afterwards (results are bit-identical to recomputing). The memory is
 * caller-owned device memory of at least ctrlsim_map_cache_bytes(n_scenes, max_veh); pass NULL to detach. The
 * directory is reset at t = 0, when another batch is stepped, or when steps are not consecutive. */
int64_t ctrlsim_map_cache_bytes(int32_t n_scenes, int32_t max_veh);
int ctrlsim_attach_map_cache(CtrlSim* h, void* mem, int64_t bytes);
void ctrlsim_map_cache_stats(const CtrlSim* h, int64_t* hits, int64_t* misses);

/* Optional prefix cache for the same steps (0..31): decoder keys / values of every token seen so far, cross-attention
 * keys / values and padding mask of the memory tokens, one slot per chunk of `chunk_groups` focal groups. A chunk
 * whose focal groups (scene, focal, member slots) are exactly those that filled its slot at step t-1 runs step t
 * incrementally: only the tokens of window steps t-1 and t (144 rows per group instead of 72 (t+1)) go through the
 * decoder and attend to the cached keys; polyline encoder, scene encoder and cross-attention K/V are not recomputed.
 * Any other chunk runs the full forward and refills its slot. Per-row arithmetic is unchanged, so sampled bins are
 * those of the full forward. Caller-owned, 256-byte aligned device memory of n_slots *
 * ctrlsim_prefix_cache_bytes(chunk_groups, 1); chunks beyond n_slots are not cached; `chunk_groups` must equal the
 * value later passed to ctrlsim_policy_step. The memory is zero-filled on attach. NULL detaches. */
int64_t ctrlsim_prefix_cache_bytes(int32_t chunk_groups, int32_t n_slots);
int ctrlsim_attach_prefix_cache(CtrlSim* h, void* mem, int64_t bytes, int32_t chunk_groups);
void ctrlsim_prefix_cache_stats(const CtrlSim* h, int64_t* incremental_chunks, int64_t* full_chunks);

/* ---- simulator: replaces nocturne_cpp Simulation/Scenario/Vehicle for the evaluator loop ------------------- */
/* S3: Vehicle::CreatePhysicsBody for every vehicle + the load-time UpdateCollision (vehicle.cc:137-179, scenario.cc:263) */
int ctrlsim_sim_reset(CtrlSim* h, CtrlSimBatch* b, void* stream);
/* S5 + T1: update_vehicle_data_dict + Policy.update_state at step t (policy_evaluator.py:99-159, policy.py:68-105) */
int ctrlsim_observe(CtrlSim* h, CtrlSimBatch* b, int32_t t, void* stream);
/* S5': compute_dense_reward + the RTG bookkeeping of update_vehicle_data_dict for step t of a real_time_rewards policy
 * (evaluators/evaluator.py:106-140, policy_evaluator.py:123-156); call right after ctrlsim_observe(t). For every vehicle:
 * signed distance to the nearest road_edge polyline (utils/data.py:152-290, float64 like numpy), nearest-vehicle distance
 * (already in tr_nearest), dense reward -> tr_dense[t]; rt_rtg[t] = start value (t = 0) or rt_rtg[t-1] - tr_dense[t-1].
 * Reference quirks kept: the goal / collision terms are those of STEP 0 (evaluator.py:112-113,136-138 index the reward
 * history with 0), and tr_nearest[t] is rescaled by max_veh_veh_distance (evaluator.py:126-127). */
int ctrlsim_dense_reward(CtrlSim* h, CtrlSimBatch* b, const CtrlSimRewardParams* rp, int32_t t, void* stream);
/* T2: greedy focal grouping of AutoregressivePolicy.get_data; writes group tables and *n_groups_total (device int32) */
int ctrlsim_plan_groups(CtrlSim* h, CtrlSimBatch* b, int32_t t, int32_t* n_groups_total, void* stream);
/* T3 + M2-M9 for compact groups [g0, g0+ng): tokenise, encode map + scene, decode, sample RTGs, second pass, sample
 * actions -> next_action (AutoregressivePolicy.predict, autoregressive_policy.py:168-253). Two-phase because RTG
 * resolution needs pass 1 of every group of a scene: phase 0 = pass 1 (logits), phase 1 = resolve RTGs (whole batch),
 * phase 2 = pass 2 + action sampling. ctrlsim_policy_step runs all phases over all groups in chunks. */
int ctrlsim_policy_step(CtrlSim* h, CtrlSimBatch* b, const CtrlSimPolicyParams* p, int32_t t, int32_t n_groups_total,
                        void* workspace, int64_t workspace_bytes, int32_t chunk_groups, void* stream);
/* S6 + S1-S4: policy.act / apply_gt_action (inverse bicycle) then Scenario::Step (scenario.cc:266-292) */
int ctrlsim_sim_step(CtrlSim* h, CtrlSimBatch* b, int32_t t, void* stream);
/* S7: per-scene partial metrics + histograms (policy_evaluator.py:162-305). out_scene [S,8] double:
 * goal_sum, n_agents, coll_mean, off_mean, has_agents, ade_sum, fde_sum, pad; out_hist [8,200] int64 (sim/gt x 4). */
int ctrlsim_metrics(CtrlSim* h, const CtrlSimBatch* b, double* out_scene, int64_t* out_hist, void* stream);

/* ---- instrumentation (bench.py): kernels launched so far by this library in this process, and optional CUDA-event
 * timing of the hot kernel classes inside ctrlsim_policy_step. out has 12 doubles: for category c in (0 linear GEMM,
 * 1 map_pool, 2 decoder self-attention, 3 decoder cross-attention): total ms, total work (flops; bytes for map_pool),
 * number of launches. */
long long ctrlsim_launch_count(void);
void ctrlsim_debug_gemm(int32_t mode); /* 1: record a clock64 timeline of CTA 0 of the linear-layer GEMM */
void ctrlsim_debug_gemm_trace(int64_t* out_192x4); /* 128 x [k-slab][TMA issued, data landed, lo tiles published, MMAs issued], then 64 x [tile][epilogue events] */
void ctrlsim_debug_attn(int32_t mode); /* bring-up aid for attention_tc.cu; 0 = normal */
void ctrlsim_debug_attn_trace(int64_t* out_8x64); /* mode 4: clock64 timeline of CTA (0,0,0), [tile][event] */
void ctrlsim_profile_enable(int32_t on);
void ctrlsim_profile_read(double* out);

/* ---- building blocks exported for parity tests and for callers that schedule the model themselves ---------- */
int ctrlsim_linear(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                   int32_t relu, void* stream);
int ctrlsim_layernorm(const float* X, const float* R, const float* gamma, const float* beta, float* Y, int32_t M,
                      int32_t relu, void* stream);
/* X[M,256] = LayerNorm(X + A[M,K] W[256,K]^T + bias) * gamma + beta in place - the projection that closes a post-LN
 * transformer sub-block (nn.TransformerDecoderLayer, modules/decoder.py:16-20) with residual add and LayerNorm in the
 * GEMM epilogue when W is a registered weight (else the GEMM and the LayerNorm run one after the other, through
 * `scratch` [M,256]).  Returns 0, or an error code. */
int ctrlsim_linear_res_ln(const float* A, const float* W, const float* bias, float* X, const float* gamma,
                          const float* beta, float* scratch, int32_t M, int32_t K, void* stream);
int ctrlsim_attn_padded(const float* Q, int32_t ldq, const float* K, const float* V, int32_t ldkv,
                        const uint8_t* key_pad, float* O, int32_t G, int32_t Lq, int32_t Lk, void* stream);
int ctrlsim_attn_causal(const float* QKV, float* O, int32_t G, int32_t n_t, void* stream);
/* the same for a given position of the state token inside an agent's (3-token) step: 0 = (state, rtg, action), the
 * CtRL-Sim order; 1 = (rtg, state, action), the decision-transformer order (utils/train_utils.py:85-88) */
int ctrlsim_attn_causal_order(const float* QKV, float* O, int32_t G, int32_t n_t, int32_t state_index, void* stream);
/* decoder self-attention of the max_agents rows of window step ti only (state rows of the last layer / rtg rows of the
 * second pass, policies/autoregressive_policy.py:190-210): queries = columns [0, 256) of qkv_rows [G * max_agents, 768];
 * keys / values = rows of KV [G * group_rows, ld] at column offsets k_off / v_off: every token of steps < ti, the state
 * tokens of step ti and - own_row - the row's own key / value (columns [256, 768) of qkv_rows). O [G * max_agents, 256]. */
int ctrlsim_attn_step(const float* KV, int32_t ld, int32_t k_off, int32_t v_off, int32_t group_rows, const float* qkv_rows,
                      float* O, int32_t G, int32_t ti, int32_t own_row, void* stream);
/* general form: own_mode 0 = nothing but history + the step's state tokens, 1 = + the row's own NEW key / value from
 * qkv_rows (as above), 2 = + the row's own FIRST token of step ti taken from KV (the decision transformer's state rows
 * see their own rtg token); state_index as in ctrlsim_attn_causal_order */
int ctrlsim_attn_step_order(const float* KV, int32_t ld, int32_t k_off, int32_t v_off, int32_t group_rows,
                            const float* qkv_rows, float* O, int32_t G, int32_t ti, int32_t own_mode, int32_t state_index,
                            void* stream);
int ctrlsim_map_pool(const float* feats, const uint8_t* pt_valid, const uint8_t* poly_valid, const float* U,
                     float* pooled, int32_t n_poly, void* stream);
/* the product's polyline front end, fused from the raw points (modules/map_encoder.py:34-45; csrc/map_encoder.cu): first
 * layer of road_pts_encoder per point, scores against the folded seed-query matrix, masked softmax over the polyline's
 * 100 points, attention-weighted sum of the HIDDEN vectors. map_pts [n_poly,100,3] (x, y, exists), poly_valid [n_poly]
 * -> pooled [n_poly, 8, 256]; the remaining linear maps are folded into the next GEMM's weight (derived.pool_W2). */
int ctrlsim_map_encode_pool(CtrlSim* h, const float* map_pts, const uint8_t* poly_valid, float* pooled, int32_t n_poly,
                            void* stream);
/* one categorical draw per row with the explicit sampler: x [rows, n] fp32 (already tilted / tempered) */
int ctrlsim_sample_rows(const float* x, int32_t rows, int32_t n, int32_t ld, int32_t stride, uint64_t seed,
                        const uint32_t* counters /* [rows,4] */, int32_t* out_idx, void* stream);
/* the same with top-p (nucleus) filtering on the sampler's integer weights (n <= 1024), see oracle/sampler.py */
int ctrlsim_sample_rows_nucleus(const float* x, int32_t rows, int32_t n, int32_t ld, int32_t stride, uint64_t seed,
                                const uint32_t* counters /* [rows,4] */, double top_p, int32_t* out_idx, void* stream);
/* full first-pass forward on caller-provided tokens of G groups (parity tests against the reference modules):
 * writes rtg logits [G,24,1050] and, after overwriting the rtg tokens at `ti` with rtg_idx [G,24,3], action logits
 * [G,24,1000]. Token arrays follow the reference MotionData layout (float32 / int32). */
/* decision-transformer handle: one forward on caller-provided tokens; rtgs are the clip-normalised continuous values
 * [G,A,32,3] float32; writes action logits [G,A,1000] read from the state rows of step ti (modules/decoder.py:55-57) */
int ctrlsim_forward_tokens_dt(CtrlSim* h, int32_t G, int32_t n_t, int32_t ti, const float* agent_states,
                              const float* agent_types, const float* goals, const int32_t* actions, const float* rtgs,
                              const int32_t* timesteps, const float* road_points, const int32_t* road_types,
                              float* action_logits, void* workspace, int64_t workspace_bytes, void* stream);
int ctrlsim_forward_tokens(CtrlSim* h, int32_t G, int32_t n_t, int32_t ti, const float* agent_states /*[G,24,32,8]*/,
                           const float* agent_types /*[G,24,5]*/, const float* goals /*[G,24,5]*/,
                           const int32_t* actions /*[G,24,32]*/, const int32_t* rtgs /*[G,24,32,3]*/,
                           const int32_t* timesteps /*[G,32]*/, const float* road_points /*[G,200,100,3]*/,
                           const int32_t* road_types /*[G,200]*/, const int32_t* rtg_idx_pass2 /*[G,24,3]*/,
                           float* rtg_logits, float* action_logits, void* workspace, int64_t workspace_bytes,
                           void* stream);
/* geometry known-answer entry points (nocturne/cpp/tests/src/geometry/{polygon,intersection}_test.cc) */
int ctrlsim_geom_poly_poly(const float* xy1, int32_t n1, const float* xy2, int32_t n2, int32_t* out, void* stream);
int ctrlsim_geom_poly_seg(const float* xy, int32_t n, const float* seg, int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
