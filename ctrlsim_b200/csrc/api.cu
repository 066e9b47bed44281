// extern "C" boundary (include/ctrlsim_b200.h): handle management, weight registry, and the per-step orchestration
// of the rollout hot path.  No torch types cross this file.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "kernels.h"
#include "model.h"
#include "model_ws.h"

namespace {
// NVTX range over the host-side enqueue of one phase (a no-op unless a profiler is attached): nsys / ncu --nvtx
// attribute the kernels launched inside to the phase
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace

namespace ctrlsim {
long long g_launch_count = 0;
thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace ctrlsim

namespace ctrlsim { void set_attn_debug(int v); void read_attn_trace(long long* out); }
using namespace ctrlsim;

struct CtrlSim {
  ModelCfg mc;
  ModelWeights w;
  bool finalized = false;
  int n_sm = 148;
  std::unordered_map<std::string, std::pair<const float*, int64_t>> reg;
  float* w_lo = nullptr;  // lo parts of all registered weights (see ctrlsim_finalize_weights)
  std::vector<int> h_group_off, h_group_focal;
  // per-focal polyline-encoder cache for the steps whose window still starts at t = 0 (see MapPlan, model_ws.h)
  float* mc_emb = nullptr;
  uint8_t* mc_valid = nullptr;
  int64_t mc_blocks = 0;
  std::vector<uint8_t> mc_dir;     // host directory: block (scene * max_veh + focal) holds a valid encoding
  const void* mc_owner = nullptr;  // batch the directory describes (its hist_state pointer)
  int mc_last_t = -1;
  int* h_idx = nullptr;            // pinned staging for the per-chunk index lists
  size_t h_idx_cap = 0;
  long long mc_hits = 0, mc_misses = 0;
  // prefix cache (see PrefixSlot, model_ws.h): one slot per chunk of a step, validated by the chunk's group signature
  char* pc_mem = nullptr;
  int64_t pc_slots = 0;
  int pc_chunk = 0;
  struct PcDir { std::vector<int> sig; int last_t = -1; };
  std::vector<PcDir> pc_dir;
  const void* pc_owner = nullptr;
  int pc_last_t = -1;
  std::vector<int> h_members;
  long long pc_hits = 0, pc_misses = 0;
};

static size_t pc_align(size_t n) { return (n + 255) & ~size_t(255); }
static size_t pc_slot_bytes(int Gc) {
  return N_DEC * pc_align((size_t)Gc * L * 2 * H * sizeof(float)) + N_DEC * pc_align((size_t)Gc * MEM * 2 * H * sizeof(float)) +
         pc_align((size_t)Gc * MEM);
}
static PrefixSlot pc_slot(char* mem, int Gc, int64_t slot) {
  PrefixSlot p;
  char* q = mem + (size_t)slot * pc_slot_bytes(Gc);
  for (int l = 0; l < N_DEC; ++l) { p.KV[l] = reinterpret_cast<float*>(q); q += pc_align((size_t)Gc * L * 2 * H * sizeof(float)); }
  for (int l = 0; l < N_DEC; ++l) { p.KVC[l] = reinterpret_cast<float*>(q); q += pc_align((size_t)Gc * MEM * 2 * H * sizeof(float)); }
  p.PAD = reinterpret_cast<uint8_t*>(q);
  return p;
}

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* ctrlsim_last_error(void) { return g_last_error.c_str(); }
int ctrlsim_abi_version(void) { return CTRLSIM_ABI_VERSION; }

int ctrlsim_create(const CtrlSimConfig* c, CtrlSim** out) {
  if (!c || !out) return set_error(-1, "ctrlsim_create: null argument");
  if (c->abi_version != CTRLSIM_ABI_VERSION) return set_error(-1, "ABI version mismatch: %d vs %d", c->abi_version, CTRLSIM_ABI_VERSION);
  if (c->hidden_dim != H || c->num_heads != NH || c->dim_feedforward != FF || c->enc_layers != N_ENC ||
      c->dec_layers != N_DEC || c->max_agents != A || c->context_len != T || c->max_polylines != P ||
      c->pts_per_polyline != NP || c->n_action_bins != N_ACT || c->n_rtg_bins != N_RTG)
    return set_error(-2, "ctrlsim_create: this library is specialised to H=256, heads=8, FF=1024, 2+4 layers, %d agents, "
                         "32 steps, %dx100 map, 1000/350 bins (got %d agents, %d polylines; the default and the wide "
                         "geometry live in libctrlsim_b200.so / libctrlsim_b200_wide.so)", A, P, c->max_agents, c->max_polylines);
  // the accel_jsd histogram (metrics_kernel) and the action un-discretisation assume the 20 x 50 factorisation
  if (c->n_steer_bins != 50)
    return set_error(-2, "ctrlsim_create: n_steer_bins=%d; the kernels assume the reference's 20 accel x 50 steer bins", c->n_steer_bins);
  int dev = 0, cc_major = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(-5, "no CUDA device: the product path has no CPU fallback");
  cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (cc_major < 10) return set_error(-5, "ctrlsim_b200 is built for sm_100a only (device reports sm_%d)", cc_major * 10);
  CtrlSim* h = new CtrlSim();
  h->n_sm = n_sm;
  h->mc.steps = c->steps; h->mc.hist_steps = c->history_steps; h->mc.n_steer = c->n_steer_bins; h->mc.dt = c->dt;
  h->mc.dt_d = (double)((int)(c->dt * 1000.0f + 0.5f)) / 1000.0;  // 0.1f -> 0.1
  h->mc.agent_dist = c->agent_dist_threshold;
  h->mc.min_accel = c->min_accel; h->mc.max_accel = c->max_accel; h->mc.min_steer = c->min_steer; h->mc.max_steer = c->max_steer;
  h->mc.pos_tol = c->pos_tol; h->mc.heading_tol = c->heading_tol; h->mc.speed_tol = c->speed_tol;
  h->mc.goal_dist_scaling = c->goal_dist_scaling; h->mc.reward_scaling = c->reward_scaling;
  h->mc.dt_model = c->decision_transformer ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    if (!(c->rtg_max[k] > c->rtg_min[k])) { delete h; return set_error(-2, "ctrlsim_create: rtg_min[%d] >= rtg_max[%d]", k, k); }
    h->mc.rtg_min[k] = c->rtg_min[k]; h->mc.rtg_max[k] = c->rtg_max[k];
  }
  { const char* e = getenv("CTRLSIM_CONTACTS"); h->mc.contacts = !(e && e[0] == '0'); }
  {  // process-wide (a __constant__ of the simulator kernels): sinf / cosf / tanf follow glibc's own algorithm
    // (glibc_trig.h) so that the simulator is bit-identical to the reference's through contacts; CTRLSIM_TRIG=fp64
    // selects "evaluate in fp64 and round once" (correctly rounded, 1 ulp off glibc in ~1 % of calls).
    // The constant is statically 1; it is only written when the other mode is asked for (or has to be taken back).
    static bool trig_is_glibc = true;
    const char* e = getenv("CTRLSIM_TRIG");
    const bool want = !(e && strcmp(e, "fp64") == 0);
    if (want != trig_is_glibc) {
      const int rc = set_trig_mode(want ? 1 : 0);
      if (rc) { delete h; return rc; }
      trig_is_glibc = want;
    }
  }
  *out = h;
  return 0;
}

void ctrlsim_destroy(CtrlSim* h) {
  if (h) { gemm_clear_weight_lo(h); if (h->w_lo) cudaFree(h->w_lo); }
  if (h && h->h_idx) cudaFreeHost(h->h_idx); delete h; }

int ctrlsim_load_weights(CtrlSim* h, const char* name, const float* ptr, int64_t count) {
  if (!h || !name || !ptr) return set_error(-1, "ctrlsim_load_weights: null argument");
  h->reg[name] = {ptr, count};
  h->finalized = false;
  return 0;
}

static int need(CtrlSim* h, const std::string& name, int64_t count, const float** out) {
  auto it = h->reg.find(name);
  if (it == h->reg.end()) return set_error(-3, "missing weight '%s'", name.c_str());
  if (it->second.second != count)
    return set_error(-3, "weight '%s' has %lld elements, expected %lld", name.c_str(), (long long)it->second.second, (long long)count);
  *out = it->second.first;
  return 0;
}
#define NEED(name, count, dst)                               \
  do {                                                       \
    int rc__ = need(h, name, count, &(dst));                 \
    if (rc__) return rc__;                                   \
  } while (0)

static int need_mlp(CtrlSim* h, const std::string& p, int din, int dout, MlpW& m) {
  NEED(p + ".mlp.0.weight", (int64_t)H * din, m.w0);
  NEED(p + ".mlp.0.bias", H, m.b0);
  NEED(p + ".mlp.1.weight", H, m.lnw);
  NEED(p + ".mlp.1.bias", H, m.lnb);
  NEED(p + ".mlp.3.weight", (int64_t)dout * H, m.w3);
  NEED(p + ".mlp.3.bias", dout, m.b3);
  return 0;
}
static int need_ln(CtrlSim* h, const std::string& p, LnW& l) {
  NEED(p + ".weight", H, l.w);
  NEED(p + ".bias", H, l.b);
  return 0;
}
static int need_mha(CtrlSim* h, const std::string& p, MhaW& m) {
  NEED(p + ".in_proj_weight", 3 * H * H, m.in_w);
  NEED(p + ".in_proj_bias", 3 * H, m.in_b);
  NEED(p + ".out_proj.weight", H * H, m.out_w);
  NEED(p + ".out_proj.bias", H, m.out_b);
  return 0;
}

int ctrlsim_finalize_weights(CtrlSim* h) {
  if (!h) return set_error(-1, "null handle");
  ModelWeights& w = h->w;
  int rc;
  const std::string me = "encoder.map_encoder";
  if ((rc = need_mlp(h, me + ".road_pts_encoder", 3, H, w.road_pts))) return rc;
  NEED("derived.pool_U", NH * H, w.pool_U);
  NEED("derived.pool_W", (int64_t)H * NH * H, w.pool_W);
  NEED("derived.pool_b", H, w.pool_b);
  NEED("derived.pool_U2", NH * H, w.pool_U2);
  NEED("derived.pool_W2", (int64_t)H * NH * H, w.pool_W2);
  NEED("derived.pool_b2", H, w.pool_b2);
  if ((rc = need_ln(h, me + ".norm1", w.map_n1))) return rc;
  if ((rc = need_ln(h, me + ".norm2", w.map_n2))) return rc;
  if ((rc = need_mlp(h, me + ".map_feats", H, H, w.map_feats))) return rc;
  NEED("derived.type_tab2", 9 * H, w.type_tab2);
  if ((rc = need_mlp(h, me + ".road_road_type_encoder", 2 * H, H, w.rr))) return rc;
  if ((rc = need_mlp(h, "encoder.embed_state", 12, H, w.embed_state))) return rc;
  if ((rc = need_mlp(h, "encoder.embed_goal", 5, H, w.embed_goal))) return rc;
  NEED("encoder.embed_state_goal.weight", 2 * H * H, w.sg_w);
  NEED("encoder.embed_state_goal.bias", H, w.sg_b);
  NEED("encoder.embed_timestep.weight", 90 * H, w.emb.ts);
  NEED("encoder.embed_agent_id.weight", A * H, w.emb.id);
  NEED("encoder.embed_action.weight", N_ACT * H, w.emb.act);
  w.dt = h->mc.dt_model != 0;
  w.emb.dt = w.dt ? 1 : 0;
  if (w.dt) {  // decision transformer: Linear(1, H) RTG embeddings folded through embed_rtg (ctrlsim_b200/model.py)
    NEED("derived.rtg_lin", 3 * H, w.emb.rtg_lin);
    NEED("derived.rtg_lin_bias", H, w.emb.rtg_bias);
    w.emb.rtg_goal = w.emb.rtg_veh = w.emb.rtg_road = nullptr;
  } else {
    NEED("derived.rtg_tab_goal", N_RTG * H, w.emb.rtg_goal);
    NEED("derived.rtg_tab_veh", N_RTG * H, w.emb.rtg_veh);
    NEED("derived.rtg_tab_road", N_RTG * H, w.emb.rtg_road);
    NEED("encoder.embed_rtg.bias", H, w.emb.rtg_bias);
  }
  NEED("encoder.embed_ln.weight", H, w.emb.ln_w);
  NEED("encoder.embed_ln.bias", H, w.emb.ln_b);
  for (int l = 0; l < N_ENC; ++l) {
    const std::string p = "encoder.transformer_encoder.layers." + std::to_string(l);
    EncLayerW& e = w.enc[l];
    if ((rc = need_mha(h, p + ".self_attn", e.sa))) return rc;
    NEED(p + ".linear1.weight", FF * H, e.l1w); NEED(p + ".linear1.bias", FF, e.l1b);
    NEED(p + ".linear2.weight", H * FF, e.l2w); NEED(p + ".linear2.bias", H, e.l2b);
    if ((rc = need_ln(h, p + ".norm1", e.n1))) return rc;
    if ((rc = need_ln(h, p + ".norm2", e.n2))) return rc;
  }
  for (int l = 0; l < N_DEC; ++l) {
    const std::string p = "decoder.transformer_decoder.layers." + std::to_string(l);
    DecLayerW& d = w.dec[l];
    if ((rc = need_mha(h, p + ".self_attn", d.sa))) return rc;
    if ((rc = need_mha(h, p + ".multihead_attn", d.ca))) return rc;
    NEED(p + ".linear1.weight", FF * H, d.l1w); NEED(p + ".linear1.bias", FF, d.l1b);
    NEED(p + ".linear2.weight", H * FF, d.l2w); NEED(p + ".linear2.bias", H, d.l2b);
    if ((rc = need_ln(h, p + ".norm1", d.n1))) return rc;
    if ((rc = need_ln(h, p + ".norm2", d.n2))) return rc;
    if ((rc = need_ln(h, p + ".norm3", d.n3))) return rc;
  }
  if ((rc = need_mlp(h, "decoder.predict_action", H, N_ACT, w.head_action))) return rc;
  if (!w.dt) { if ((rc = need_mlp(h, "decoder.predict_rtg", H, N_RTG * 3, w.head_rtg))) return rc; }
  else w.head_rtg = MlpW();
  // lo-part copies (x - trunc13(x)) of every registered tensor, in one allocation owned by the handle: the GEMM's TMA
  // fetches them as the W_lo operand of the 3xTF32 split. Call ctrlsim_finalize_weights again after changing weights.
  gemm_clear_weight_lo(h);
  if (h->w_lo) { cudaFree(h->w_lo); h->w_lo = nullptr; }
  if (getenv("CTRLSIM_WLO") == nullptr || std::string(getenv("CTRLSIM_WLO")) != "0") {
    size_t total = 0;
    for (auto& kv : h->reg) total += ((size_t)kv.second.second + 63) & ~size_t(63);
    if (cudaMalloc(&h->w_lo, total * sizeof(float)) != cudaSuccess) return set_error(-5, "finalize_weights: cudaMalloc of %zu bytes failed", total * sizeof(float));
    size_t off = 0;
    for (auto& kv : h->reg) {
      const size_t n = (size_t)kv.second.second;
      if ((rc = launch_weight_lo(kv.second.first, h->w_lo + off, n, 0))) return rc;
      gemm_register_weight_lo(h, kv.second.first, n, h->w_lo + off);
      off += (n + 63) & ~size_t(63);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return set_error(-5, "finalize_weights: lo-part kernel failed");
  }
  h->finalized = true;
  return 0;
}

int64_t ctrlsim_workspace_bytes(const CtrlSim* h, int32_t max_groups) {
  (void)h;
  Workspace ws;
  return (int64_t)ws.carve(nullptr, 0, max_groups);
}

int64_t ctrlsim_map_cache_bytes(int32_t n_scenes, int32_t max_veh) {
  return (int64_t)n_scenes * max_veh * ((int64_t)P * H * sizeof(float) + P) + 256;
}
int ctrlsim_attach_map_cache(CtrlSim* h, void* mem, int64_t bytes) {
  if (!h) return set_error(-1, "ctrlsim_attach_map_cache: null handle");
  h->mc_dir.clear(); h->mc_owner = nullptr; h->mc_last_t = -1;
  if (!mem || bytes <= 0) { h->mc_emb = nullptr; h->mc_valid = nullptr; h->mc_blocks = 0; return 0; }
  if (reinterpret_cast<uintptr_t>(mem) & 15) return set_error(-2, "ctrlsim_attach_map_cache: memory must be 16-byte aligned");
  const int64_t per = (int64_t)P * H * sizeof(float) + P;
  h->mc_blocks = (bytes - 256) / per;
  if (h->mc_blocks <= 0) return set_error(-4, "ctrlsim_attach_map_cache: %lld bytes hold no block", (long long)bytes);
  h->mc_emb = reinterpret_cast<float*>(mem);
  h->mc_valid = reinterpret_cast<uint8_t*>(mem) + h->mc_blocks * (int64_t)P * H * sizeof(float);
  return 0;
}
void ctrlsim_map_cache_stats(const CtrlSim* h, int64_t* hits, int64_t* misses) {
  if (hits) *hits = h ? h->mc_hits : 0;
  if (misses) *misses = h ? h->mc_misses : 0;
}

int64_t ctrlsim_prefix_cache_bytes(int32_t chunk_groups, int32_t n_slots) {
  return (int64_t)pc_slot_bytes(chunk_groups) * n_slots;
}
int ctrlsim_attach_prefix_cache(CtrlSim* h, void* mem, int64_t bytes, int32_t chunk_groups) {
  if (!h) return set_error(-1, "ctrlsim_attach_prefix_cache: null handle");
  h->pc_dir.clear(); h->pc_owner = nullptr; h->pc_last_t = -1;
  h->pc_mem = nullptr; h->pc_slots = 0; h->pc_chunk = 0;
  if (!mem || bytes <= 0) return 0;
  if (chunk_groups <= 0) return set_error(-2, "ctrlsim_attach_prefix_cache: chunk_groups must be positive");
  if (reinterpret_cast<uintptr_t>(mem) & 255) return set_error(-2, "ctrlsim_attach_prefix_cache: memory must be 256-byte aligned");
  const int64_t slots = bytes / (int64_t)pc_slot_bytes(chunk_groups);
  if (slots <= 0) return set_error(-4, "ctrlsim_attach_prefix_cache: %lld bytes hold no slot of %d groups (%lld needed)",
                                   (long long)bytes, chunk_groups, (long long)pc_slot_bytes(chunk_groups));
  // key rows past a group's current length are masked, but they are still multiplied: keep them finite
  if (cudaMemset(mem, 0, (size_t)slots * pc_slot_bytes(chunk_groups)) != cudaSuccess) return set_error(-5, "ctrlsim_attach_prefix_cache: cudaMemset failed");
  h->pc_mem = reinterpret_cast<char*>(mem); h->pc_slots = slots; h->pc_chunk = chunk_groups;
  h->pc_dir.assign((size_t)slots, CtrlSim::PcDir());
  return 0;
}
void ctrlsim_prefix_cache_stats(const CtrlSim* h, int64_t* incremental_chunks, int64_t* full_chunks) {
  if (incremental_chunks) *incremental_chunks = h ? h->pc_hits : 0;
  if (full_chunks) *full_chunks = h ? h->pc_misses : 0;
}

int ctrlsim_sim_reset(CtrlSim* h, CtrlSimBatch* b, void* stream) { return launch_sim_reset(*b, h->mc, S(stream)); }
int ctrlsim_observe(CtrlSim* h, CtrlSimBatch* b, int32_t t, void* stream) { return launch_observe(*b, t, h->mc, S(stream)); }
int ctrlsim_dense_reward(CtrlSim* h, CtrlSimBatch* b, const CtrlSimRewardParams* rp, int32_t t, void* stream) {
  if (!h || !b || !rp) return set_error(-1, "ctrlsim_dense_reward: null argument");
  if (rp->return_mode < 0 || rp->return_mode > 3) return set_error(-2, "ctrlsim_dense_reward: return_mode=%d", rp->return_mode);
  return launch_dense_reward(*b, *rp, t, h->mc, S(stream));
}
int ctrlsim_plan_groups(CtrlSim* h, CtrlSimBatch* b, int32_t t, int32_t* n_groups_total, void* stream) {
  return launch_plan_groups(*b, t, h->mc, n_groups_total, S(stream));
}
int ctrlsim_sim_step(CtrlSim* h, CtrlSimBatch* b, int32_t t, void* stream) {
  NvtxRange r("ctrlsim_sim_step");
  return launch_sim_step(*b, t, h->mc, S(stream));
}
int ctrlsim_metrics(CtrlSim* h, const CtrlSimBatch* b, double* out_scene, int64_t* out_hist, void* stream) {
  return launch_metrics(*b, h->mc, out_scene, reinterpret_cast<long long*>(out_hist), S(stream));
}

int ctrlsim_policy_step(CtrlSim* h, CtrlSimBatch* b, const CtrlSimPolicyParams* p, int32_t t, int32_t n_groups_total,
                        void* workspace, int64_t workspace_bytes, int32_t chunk_groups, void* stream) {
  if (!h || !h->finalized) return set_error(-3, "ctrlsim_policy_step: weights not finalized");
  if (n_groups_total <= 0) return 0;
  const int rtg_mode = p->rtg_mode;  // 0: RTGs predicted and sampled; 1: tracked in real time (b->rt_rtg)
  if (rtg_mode != 0 && rtg_mode != 1) return set_error(-2, "ctrlsim_policy_step: rtg_mode=%d", rtg_mode);
  if (h->w.dt && rtg_mode != 1) return set_error(-2, "ctrlsim_policy_step: the decision transformer has no RTG head; set rtg_mode = 1");
  NvtxRange nvtx_step("ctrlsim_policy_step");
  cudaStream_t st = S(stream);
  const int Sn = b->n_scenes;
  h->h_group_off.resize(Sn + 1);
  const int Nv = b->max_veh;
  // the map cache applies while the window starts at t = 0 and the batch fits the attached memory
  const bool use_mc = t < T && h->mc_emb && (int64_t)Sn * Nv <= h->mc_blocks;
  cudaError_t e = cudaMemcpyAsync(h->h_group_off.data(), b->group_off, sizeof(int) * (Sn + 1), cudaMemcpyDeviceToHost, st);
  const bool use_pc = t < T && h->pc_mem && h->pc_chunk == chunk_groups;
  if ((use_mc || use_pc) && e == cudaSuccess) {
    h->h_group_focal.resize((size_t)Sn * Nv);
    e = cudaMemcpyAsync(h->h_group_focal.data(), b->group_focal, sizeof(int) * (size_t)Sn * Nv, cudaMemcpyDeviceToHost, st);
  }
  if (use_pc && e == cudaSuccess) {
    h->h_members.resize((size_t)Sn * Nv * A);
    e = cudaMemcpyAsync(h->h_members.data(), b->group_members, sizeof(int) * (size_t)Sn * Nv * A, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return set_error(-5, "policy_step: reading group offsets: %s", cudaGetErrorString(e));
  const std::vector<int>& off = h->h_group_off;
  if (off[Sn] != n_groups_total) return set_error(-4, "policy_step: n_groups_total=%d but the plan holds %d", n_groups_total, off[Sn]);
  Workspace ws;
  const size_t need_bytes = ws.carve(nullptr, 0, chunk_groups);
  if ((int64_t)need_bytes > workspace_bytes)
    return set_error(-4, "policy_step: workspace of %lld bytes is too small for chunks of %d groups (%lld needed)",
                     (long long)workspace_bytes, chunk_groups, (long long)need_bytes);
  ws.carve(workspace, need_bytes, chunk_groups);
  const int n_t = t < T ? t + 1 : T;
  if (use_mc) {
    // a directory entry stays valid for one episode of one batch, stepped in order
    if (t == 0 || h->mc_owner != (const void*)b->hist_state || t != h->mc_last_t + 1 || h->mc_dir.size() != (size_t)Sn * Nv) {
      h->mc_dir.assign((size_t)Sn * Nv, 0);
      h->mc_owner = (const void*)b->hist_state;
    }
    h->mc_last_t = t;
    const size_t need_idx = 3 * (size_t)n_groups_total;
    if (need_idx > h->h_idx_cap) {
      if (h->h_idx) cudaFreeHost(h->h_idx);
      h->h_idx = nullptr; h->h_idx_cap = 0;
      if (cudaMallocHost(&h->h_idx, need_idx * 2 * sizeof(int)) != cudaSuccess) return set_error(-5, "policy_step: pinned staging allocation failed");
      h->h_idx_cap = need_idx * 2;
    }
  } else {
    h->mc_last_t = -1;
  }
  if (use_pc) {
    if (t == 0 || h->pc_owner != (const void*)b->hist_state || t != h->pc_last_t + 1) {
      for (auto& d : h->pc_dir) d.last_t = -1;
      h->pc_owner = (const void*)b->hist_state;
    }
    h->pc_last_t = t;
  } else {
    h->pc_last_t = -1;
  }
  int s0 = 0;
  int64_t chunk_idx = 0;
  while (s0 < Sn) {
    int s1 = s0;
    while (s1 < Sn && off[s1 + 1] - off[s0] <= chunk_groups) ++s1;
    if (s1 == s0) return set_error(-4, "policy_step: scene %d has %d groups, more than chunk_groups=%d", s0, off[s0 + 1] - off[s0], chunk_groups);
    const int g0 = off[s0], ng = off[s1] - off[s0];
    if (ng > 0) {
      int rc;
      // prefix cache: the chunk runs incrementally if its slot was filled at step t-1 by exactly these groups
      PrefixSlot slot;
      const PrefixSlot* pc = nullptr;
      if (use_pc && chunk_idx < h->pc_slots) {
        CtrlSim::PcDir& d = h->pc_dir[(size_t)chunk_idx];
        std::vector<int> sig;
        sig.reserve((size_t)ng * (A + 2));
        for (int sc = s0; sc < s1; ++sc)
          for (int lg = 0; lg < off[sc + 1] - off[sc]; ++lg) {
            sig.push_back(sc);
            sig.push_back(h->h_group_focal[(size_t)sc * Nv + lg]);
            const int* m = h->h_members.data() + ((size_t)sc * Nv + lg) * A;
            sig.insert(sig.end(), m, m + A);
          }
        slot = pc_slot(h->pc_mem, h->pc_chunk, chunk_idx);
        slot.incr = t >= 1 && d.last_t == t - 1 && d.sig == sig;
        d.sig.swap(sig);
        d.last_t = t;
        pc = &slot;
        if (slot.incr) ++h->pc_hits; else ++h->pc_misses;
      }
      ++chunk_idx;
      MapPlan mp;
      NvtxRange nvtx_chunk(pc && pc->incr ? "chunk (incremental decode)" : "chunk (full window)");
      {
      NvtxRange r_tok("tokenize");
      if (pc && pc->incr) {
        // nothing upstream of the decoder is recomputed; tokenise the last two window steps only
        if ((rc = launch_tokenize(*b, g0, ng, t, 2, ws.tk, h->mc, st, ws.map_sel, 0, t - 1, rtg_mode))) return rc;
      } else if (use_mc) {
        int* slot_l = h->h_idx + 3 * (size_t)g0;  // [ng] slot | [ng] sel | [ng] dst, staged per chunk
        int* sel = slot_l + ng;
        int* dst = sel + ng;
        int n_miss = 0, gl = 0;
        for (int sc = s0; sc < s1; ++sc)
          for (int lg = 0; lg < off[sc + 1] - off[sc]; ++lg, ++gl) {
            const int blk = sc * Nv + h->h_group_focal[(size_t)sc * Nv + lg];
            slot_l[gl] = blk;
            if (!h->mc_dir[blk]) { h->mc_dir[blk] = 1; sel[n_miss] = gl; dst[n_miss] = blk; ++n_miss; }
          }
        h->mc_misses += n_miss; h->mc_hits += ng - n_miss;
        cudaError_t ce = cudaMemcpyAsync(ws.map_slot, slot_l, sizeof(int) * ng, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess && n_miss) ce = cudaMemcpyAsync(ws.map_sel, sel, sizeof(int) * n_miss, cudaMemcpyHostToDevice, st);
        if (ce == cudaSuccess && n_miss) ce = cudaMemcpyAsync(ws.map_dst, dst, sizeof(int) * n_miss, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) return set_error(-5, "policy_step: uploading map-cache lists: %s", cudaGetErrorString(ce));
        mp.n_map = n_miss; mp.slot = ws.map_slot; mp.dst = ws.map_dst; mp.cache_emb = h->mc_emb; mp.cache_valid = h->mc_valid;
        if ((rc = launch_tokenize(*b, g0, ng, t, n_t, ws.tk, h->mc, st, ws.map_sel, n_miss, 0, rtg_mode))) return rc;
      } else {
        if ((rc = launch_tokenize(*b, g0, ng, t, n_t, ws.tk, h->mc, st, nullptr, 0, 0, rtg_mode))) return rc;
      }
      }
      {
        NvtxRange r("pass 1: encoders + decoder -> RTG logits");
        if ((rc = forward_pass1(h->w, ws, ng, n_t, h->n_sm, st, mp, pc))) return rc;
      }
      if (rtg_mode == 0) {
        NvtxRange r("sample RTG");
        if ((rc = launch_resolve_rtg_range(*b, *p, t, s0, s1, g0, h->mc.steps, ws.rtg_logits, st))) return rc;
        if ((rc = launch_gather_rtg_steps(*b, g0, ng, t, h->mc.steps, ws.rtg_new, st))) return rc;
      } else if (!h->w.dt) {  // tracked RTGs, CtRL-Sim network: the bins of the current step as the tokeniser made them
        const bool incr = pc && pc->incr;
        if ((rc = launch_gather_rtg_tokens(ng, incr ? 2 : n_t, incr ? 1 : n_t - 1, ws.tk, ws.rtg_new, st))) return rc;
      }
      if (!h->w.dt) {  // the decision transformer's single pass already produced the action logits (state rows)
        NvtxRange r("pass 2: last-step rows -> action logits");
        if ((rc = forward_pass2(h->w, ws, ng, n_t, st, pc))) return rc;
      }
      NvtxRange r("sample actions");
      if ((rc = launch_sample_actions(*b, *p, g0, ng, t, ws.act_logits, h->mc, st))) return rc;
    } else if (rtg_mode == 0) {
      int rc;  // scenes without any group still owe their vehicles the "(0,0,0) RTG appended" bookkeeping
      if ((rc = launch_resolve_rtg_range(*b, *p, t, s0, s1, g0, h->mc.steps, ws.rtg_logits, st))) return rc;
    }
    s0 = s1;
  }
  g_prof.flush(st);
  return 0;
}

void ctrlsim_debug_gemm(int32_t v) { ctrlsim::set_gemm_debug(v); }
void ctrlsim_debug_gemm_trace(int64_t* out) { ctrlsim::read_gemm_trace(reinterpret_cast<long long*>(out)); }
void ctrlsim_debug_attn(int32_t v) { ctrlsim::set_attn_debug(v); }
void ctrlsim_debug_attn_trace(int64_t* out) { ctrlsim::read_attn_trace(reinterpret_cast<long long*>(out)); }
long long ctrlsim_launch_count(void) { return g_launch_count; }
void ctrlsim_profile_enable(int32_t on) {
  g_prof.on = on != 0;
  if (on) for (int c = 0; c < PROF_NCAT; ++c) { g_prof.tot_ms[c] = 0; g_prof.tot_work[c] = 0; g_prof.count[c] = 0; }
}
/* out[cat*3 + {0,1,2}] = total ms, total work (flops or bytes), launches; cat: 0 gemm, 1 map_pool, 2 attn_causal, 3 attn_cross */
void ctrlsim_profile_read(double* out) {
  for (int c = 0; c < PROF_NCAT; ++c) { out[c * 3] = g_prof.tot_ms[c]; out[c * 3 + 1] = g_prof.tot_work[c]; out[c * 3 + 2] = (double)g_prof.count[c]; }
}

int ctrlsim_linear(const float* Ain, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                   int32_t relu, void* stream) {
  GemmArgs g;
  g.A = Ain; g.W = W; g.bias = bias; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldc = N; g.relu = relu != 0;
  return launch_gemm(g, S(stream));
}
int ctrlsim_layernorm(const float* X, const float* R, const float* gamma, const float* beta, float* Y, int32_t M,
                      int32_t relu, void* stream) {
  return launch_layernorm(X, R, gamma, beta, Y, M, H, H, H, relu != 0, S(stream));
}
int ctrlsim_linear_res_ln(const float* A, const float* W, const float* bias, float* X, const float* gamma,
                          const float* beta, float* scratch, int32_t M, int32_t K, void* stream) {
  const int rc = launch_gemm_res_ln(A, K, W, K, bias, X, H, gamma, beta, M, K, S(stream));
  if (rc != 1) return rc;
  if (!scratch) return set_error(-2, "ctrlsim_linear_res_ln: the fused kernel cannot serve this call and no scratch buffer was given");
  GemmArgs g;
  g.A = A; g.W = W; g.bias = bias; g.C = scratch; g.M = M; g.N = H; g.K = K; g.lda = K; g.ldw = K; g.ldc = H;
  if (int r2 = launch_gemm(g, S(stream))) return r2;
  return launch_layernorm(X, scratch, gamma, beta, X, M, H, H, H, false, S(stream));
}
int ctrlsim_attn_padded(const float* Q, int32_t ldq, const float* K, const float* V, int32_t ldkv,
                        const uint8_t* key_pad, float* O, int32_t G, int32_t Lq, int32_t Lk, void* stream) {
  return launch_attn_padded(Q, ldq, K, V, ldkv, key_pad, O, H, G, Lq, Lk, S(stream));
}
int ctrlsim_attn_causal(const float* QKV, float* O, int32_t G, int32_t n_t, void* stream) {
  return launch_attn_causal(QKV, O, G, n_t, S(stream));
}
int ctrlsim_attn_causal_order(const float* QKV, float* O, int32_t G, int32_t n_t, int32_t state_index, void* stream) {
  if (state_index < 0 || state_index >= KT) return set_error(-2, "ctrlsim_attn_causal_order: state_index=%d", state_index);
  return launch_attn_causal(QKV, O, G, n_t, S(stream), state_index);
}
int ctrlsim_attn_step(const float* KV, int32_t ld, int32_t k_off, int32_t v_off, int32_t group_rows, const float* qkv_rows,
                      float* O, int32_t G, int32_t ti, int32_t own_row, void* stream) {
  if (ti < 0 || ti >= T || (ti + 1) * TOK_T > group_rows) return set_error(-2, "ctrlsim_attn_step: need 0 <= ti < %d and (ti + 1) * %d <= group_rows", T, TOK_T);
  KvView v; v.base = KV; v.ld = ld; v.k_off = k_off; v.v_off = v_off; v.group_rows = group_rows;
  return launch_attn_step(v, qkv_rows, O, G, ti, own_row != 0 ? 1 : 0, S(stream));
}
int ctrlsim_attn_step_order(const float* KV, int32_t ld, int32_t k_off, int32_t v_off, int32_t group_rows,
                            const float* qkv_rows, float* O, int32_t G, int32_t ti, int32_t own_mode, int32_t state_index,
                            void* stream) {
  if (ti < 0 || ti >= T || (ti + 1) * TOK_T > group_rows) return set_error(-2, "ctrlsim_attn_step_order: need 0 <= ti < %d and (ti + 1) * %d <= group_rows", T, TOK_T);
  KvView v; v.base = KV; v.ld = ld; v.k_off = k_off; v.v_off = v_off; v.group_rows = group_rows;
  return launch_attn_step(v, qkv_rows, O, G, ti, own_mode, S(stream), state_index);
}
int ctrlsim_map_pool(const float* feats, const uint8_t* pt_valid, const uint8_t* poly_valid, const float* U,
                     float* pooled, int32_t n_poly, void* stream) {
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return launch_map_pool(feats, pt_valid, poly_valid, U, pooled, n_poly, n_sm, S(stream));
}
int ctrlsim_map_encode_pool(CtrlSim* h, const float* map_pts, const uint8_t* poly_valid, float* pooled, int32_t n_poly,
                            void* stream) {
  if (!h || !h->finalized) return set_error(-3, "ctrlsim_map_encode_pool: weights not finalized");
  return launch_map_encode_pool(map_pts, h->w.road_pts, h->w.pool_U2, poly_valid, pooled, n_poly, h->n_sm, S(stream));
}
int ctrlsim_sample_rows(const float* x, int32_t rows, int32_t n, int32_t ld, int32_t stride, uint64_t seed,
                        const uint32_t* counters, int32_t* out_idx, void* stream) {
  return launch_sample_rows(x, rows, n, ld, stride, seed, counters, out_idx, S(stream));
}

int ctrlsim_sample_rows_nucleus(const float* x, int32_t rows, int32_t n, int32_t ld, int32_t stride, uint64_t seed,
                                const uint32_t* counters, double top_p, int32_t* out_idx, void* stream) {
  return launch_sample_rows(x, rows, n, ld, stride, seed, counters, out_idx, S(stream), true, top_p);
}

int ctrlsim_forward_tokens_dt(CtrlSim* h, int32_t G, int32_t n_t, int32_t ti, const float* agent_states,
                              const float* agent_types, const float* goals, const int32_t* actions, const float* rtgs,
                              const int32_t* timesteps, const float* road_points, const int32_t* road_types,
                              float* action_logits, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !h->finalized) return set_error(-3, "ctrlsim_forward_tokens_dt: weights not finalized");
  if (!h->w.dt) return set_error(-2, "ctrlsim_forward_tokens_dt: the handle holds the CtRL-Sim network, not the decision transformer");
  if (n_t < 1 || n_t > T || ti != n_t - 1) return set_error(-2, "forward_tokens_dt: need 1 <= n_t <= 32 and ti == n_t - 1");
  cudaStream_t st = S(stream);
  Workspace ws;
  const size_t need_bytes = ws.carve(nullptr, 0, G);
  if ((int64_t)need_bytes > workspace_bytes)
    return set_error(-4, "forward_tokens_dt: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need_bytes);
  ws.carve(workspace, need_bytes, G);
  int rc;
  if ((rc = launch_convert_tokens(G, n_t, agent_states, agent_types, goals, actions, rtgs, true, timesteps, road_points, road_types, ws.tk, st))) return rc;
  if ((rc = forward_pass1(h->w, ws, G, n_t, h->n_sm, st))) return rc;
  cudaMemcpyAsync(action_logits, ws.act_logits, sizeof(float) * (size_t)G * A * N_ACT, cudaMemcpyDeviceToDevice, st);
  return 0;
}

int ctrlsim_forward_tokens(CtrlSim* h, int32_t G, int32_t n_t, int32_t ti, const float* agent_states,
                           const float* agent_types, const float* goals, const int32_t* actions, const int32_t* rtgs,
                           const int32_t* timesteps, const float* road_points, const int32_t* road_types,
                           const int32_t* rtg_idx_pass2, float* rtg_logits, float* action_logits, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  if (!h || !h->finalized) return set_error(-3, "ctrlsim_forward_tokens: weights not finalized");
  if (n_t < 1 || n_t > T || ti != n_t - 1) return set_error(-2, "forward_tokens: need 1 <= n_t <= 32 and ti == n_t - 1");
  cudaStream_t st = S(stream);
  Workspace ws;
  const size_t need_bytes = ws.carve(nullptr, 0, G);
  if ((int64_t)need_bytes > workspace_bytes)
    return set_error(-4, "forward_tokens: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)need_bytes);
  ws.carve(workspace, need_bytes, G);
  int rc;
  if (h->w.dt) return set_error(-2, "ctrlsim_forward_tokens: the handle holds the decision transformer; use ctrlsim_forward_tokens_dt");
  if ((rc = launch_convert_tokens(G, n_t, agent_states, agent_types, goals, actions, rtgs, false, timesteps, road_points, road_types, ws.tk, st))) return rc;
  if ((rc = forward_pass1(h->w, ws, G, n_t, h->n_sm, st))) return rc;
  cudaMemcpyAsync(rtg_logits, ws.rtg_logits, sizeof(float) * (size_t)G * A * N_RTG * 3, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(ws.rtg_new, rtg_idx_pass2, sizeof(int) * (size_t)G * A * 3, cudaMemcpyDeviceToDevice, st);
  if ((rc = forward_pass2(h->w, ws, G, n_t, st))) return rc;
  cudaMemcpyAsync(action_logits, ws.act_logits, sizeof(float) * (size_t)G * A * N_ACT, cudaMemcpyDeviceToDevice, st);
  return 0;
}

int ctrlsim_geom_poly_poly(const float* xy1, int32_t n1, const float* xy2, int32_t n2, int32_t* out, void* stream) {
  return launch_geom_poly_poly(xy1, n1, xy2, n2, out, S(stream));
}
int ctrlsim_geom_poly_seg(const float* xy, int32_t n, const float* seg, int32_t* out, void* stream) {
  return launch_geom_poly_seg(xy, n, seg, out, S(stream));
}

}  // extern "C"
