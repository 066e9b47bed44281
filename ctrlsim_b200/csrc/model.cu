// Forward pass of the CtRL-Sim policy network for a chunk of focal groups (M2-M9), as a sequence of kernel launches
// on one stream.  Reference: models/ctrl_sim.py:41-45 -> modules/encoder.py:50-178, modules/map_encoder.py:34-54,
// modules/decoder.py:39-78; the reference runs the whole 2304-token network twice per group per step
// (policies/autoregressive_policy.py:190,210).  Here:
//   pass 1  runs the network over the n_t = min(t+1, 32) window steps that exist (later steps are invisible to the
//           current one under mask rule M1), keeps every decoder layer's K/V rows and the cross-attention K/V of the
//           memory tokens, and evaluates the RTG head on the 24 state rows of the current step only;
//   pass 2  recomputes only the 24 rtg-token rows of the current step (the only rows whose inputs changed after the
//           RTGs were sampled and that the action head reads), attending to the cached K/V.
// Decision-transformer variant (ModelWeights::dt; cfgs/model/dt.yaml): tokens ordered (rtg, state, action) with continuous
// RTG inputs, no RTG head; ONE pass - the action head reads the 24 state rows of the current step, which is exactly what
// pass 1 computes for the last layer.
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <string>

#include "model.h"
#include "model_ws.h"

namespace ctrlsim {

#define CS_TRY(x)            \
  do {                       \
    int rc__ = (x);          \
    if (rc__ != 0) return rc__; \
  } while (0)

// ---- optional per-category CUDA-event timing (bench.py roofline figures) ---------------------------------------------
Prof g_prof;
void Prof::begin(int cat, double work, cudaStream_t st) {
  if (!on) return;
  Rec r;
  r.cat = cat; r.work = work;
  if (pool.size() < 2) { for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); } }
  r.a = pool.back(); pool.pop_back();
  r.b = pool.back(); pool.pop_back();
  cudaEventRecord(r.a, st);
  recs.push_back(r);
}
void Prof::end(cudaStream_t st) {
  if (!on) return;
  cudaEventRecord(recs.back().b, st);
}
void Prof::cancel() {
  if (!on || recs.empty()) return;
  pool.push_back(recs.back().a); pool.push_back(recs.back().b);
  recs.pop_back();
}
void Prof::flush(cudaStream_t st) {
  if (!on || recs.empty()) return;
  cudaStreamSynchronize(st);
  for (auto& r : recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    tot_ms[r.cat] += ms; tot_work[r.cat] += r.work; ++count[r.cat];
    pool.push_back(r.a); pool.push_back(r.b);
  }
  recs.clear();
}

static int gemm(const float* Ain, const float* W, const float* bias, float* C, int M, int N, int K, int lda, int ldw,
                int ldc, bool relu, cudaStream_t st, const float* table = nullptr, const int* tidx = nullptr,
                int ldt = 0, const int* agather = nullptr) {
  GemmArgs g;
  g.A = Ain; g.W = W; g.bias = bias; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldc = ldc;
  g.relu = relu; g.table = table; g.tidx = tidx; g.ldt = ldt; g.agather = agather;
  g_prof.begin(PROF_GEMM, 2.0 * M * (double)N * K, st);
  const int rc = launch_gemm(g, st);
  g_prof.end(st);
  return rc;
}
static int ln(const float* X, const float* R, const LnW& w, float* Y, int M, bool relu, cudaStream_t st) {
  return launch_layernorm(X, R, w.w, w.b, Y, M, H, H, H, relu, st);
}
// X = LayerNorm(X + A W^T + b): the projection that closes a sub-block (post-LN).  One fused kernel when the GEMM can
// take it (registered weights, M large enough to fill the SMs), else GEMM into tmp + the LayerNorm kernel.
static int gemm_res_ln(const float* A, const float* W, const float* bias, float* X, const LnW& w, float* tmp, int M, int K,
                       cudaStream_t st) {
  // The fused kernel (gemm_tc_ta_ln_kernel) is correct and exported (ctrlsim_linear_res_ln) but OFF in the model path:
  // measured at 64 scenes (profiles/r03k_ln_fusion_experiment.txt) the full-window step is 158.6 ms with all three
  // sub-block projections fused and 156.4 ms with FFN2 only, against 155.8-156.1 ms with the separate LayerNorm kernel -
  // with K = 256 the fused epilogue (residual read, pre-norm write, barrier, re-read, normalise, write) outlasts the
  // tile's 8 k-slabs of MMAs, and the step runs at the power cap, where the saved HBM traffic buys no clock.
  // CTRLSIM_LNFUSE=1 switches it on (CTRLSIM_LNFUSE_MINK: smallest K that takes it, default 1024).
  static const bool fuse = getenv("CTRLSIM_LNFUSE") && std::string(getenv("CTRLSIM_LNFUSE")) == "1";
  static const int min_k = getenv("CTRLSIM_LNFUSE_MINK") ? atoi(getenv("CTRLSIM_LNFUSE_MINK")) : 1024;
  if (fuse && M >= 128 * 148 && K >= min_k) {
    g_prof.begin(PROF_GEMM, 2.0 * M * (double)H * K, st);
    const int rc = launch_gemm_res_ln(A, K, W, K, bias, X, H, w.w, w.b, M, K, st);
    if (rc != 1) { g_prof.end(st); return rc; }
    g_prof.cancel();
  }
  CS_TRY(gemm(A, W, bias, tmp, M, H, K, K, K, H, false, st));
  return ln(X, tmp, w, X, M, false, st);
}

size_t Workspace::carve(void* base, size_t bytes, int Gc) {
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n) -> void* {
    off = (off + 255) & ~size_t(255);
    void* r = p ? p + off : nullptr;
    off += n;
    return r;
  };
  const size_t G = (size_t)Gc;
  const size_t R = G * L, Rm = G * MEM, Rp = G * P, Rpt = Rp * NP, Ra = G * A, Rta = G * T * A;
  tk.feat_state = (float*)take(Rta * 12 * 4);
  tk.exist = (uint8_t*)take(Rta);
  tk.goal_feat = (float*)take(Ra * 5 * 4);
  tk.act_idx = (int*)take(Rta * 4);
  tk.rtg_idx = (int*)take(Rta * 3 * 4);
  tk.rtg_val = (float*)take(Rta * 3 * 4);
  tk.ts = (int*)take(G * T * 4);
  tk.map_pts = (float*)take(Rpt * 3 * 4);
  tk.map_type = (int*)take(Rp * 4);
  tk.frame = (double*)take(G * 4 * 8);
  h1 = (float*)take(Rpt * H * 4);
  feats = (float*)take(Rpt * H * 4);
  pt_valid = (uint8_t*)take(Rpt);
  poly_valid = (uint8_t*)take(Rp);
  pooled = (float*)take(Rp * NH * H * 4);
  pe_a = (float*)take(Rp * H * 4);
  pe_b = (float*)take(Rp * H * 4);
  pe_c = (float*)take(Rp * H * 4);
  type_idx = (int*)take(Rp * 4);
  map_slot = (int*)take(G * 4);
  map_sel = (int*)take(G * 4);
  map_dst = (int*)take(G * 4);
  s1 = (float*)take(Rta * H * 4);
  s2 = (float*)take(Rta * H * 4);
  sg = (float*)take(Rta * H * 4);
  g1 = (float*)take(Ra * H * 4);
  g2 = (float*)take(Ra * H * 4);
  gpart = (float*)take(Ra * H * 4);
  goal_idx = (int*)take(Rta * 4);
  mem = (float*)take(Rm * H * 4);
  pad = (uint8_t*)take(Rm);
  qkv_m = (float*)take(Rm * 3 * H * 4);
  att_m = (float*)take(Rm * H * 4);
  tmp_m = (float*)take(Rm * H * 4);
  ff_m = (float*)take(Rm * FF * 4);
  X = (float*)take(R * H * 4);
  for (int l = 0; l < N_DEC; ++l) QKV[l] = (float*)take(R * 3 * H * 4);
  for (int l = 0; l < N_DEC; ++l) kv_c[l] = (float*)take(Rm * 2 * H * 4);
  att = (float*)take(R * H * 4);
  tmp = (float*)take(R * H * 4);
  q_c = (float*)take(R * H * 4);
  ff = (float*)take(R * FF * 4);
  row_idx = (int*)take(Ra * 4);
  hd1 = (float*)take(Ra * H * 4);
  rtg_logits = (float*)take(Ra * N_RTG * 3 * 4);
  act_logits = (float*)take(Ra * N_ACT * 4);
  rtg_new = (int*)take(Ra * 3 * 4);
  xr = (float*)take(Ra * H * 4);
  qkv_r = (float*)take(Ra * 3 * H * 4);
  att_r = (float*)take(Ra * H * 4);
  tmp_r = (float*)take(Ra * H * 4);
  qc_r = (float*)take(Ra * H * 4);
  ff_r = (float*)take(Ra * FF * 4);
  off = (off + 255) & ~size_t(255);
  (void)bytes;
  return off;
}

// ---- pieces shared by the full forward and the incremental (prefix-cache) forward -----------------------------------
// M3: embeddings of n_tok token steps per group -> X [G * n_tok * 72, H] (and the initial-state memory rows if asked)
static int embed_tokens(const ModelWeights& w, Workspace& ws, int G, int n_tok, bool with_mem, cudaStream_t st) {
  const int Ra = G * A, Rta = G * n_tok * A;
  CS_TRY(launch_small_mlp1(12, ws.tk.feat_state, w.embed_state, ws.s1, (size_t)Rta, st));
  CS_TRY(gemm(ws.s1, w.embed_state.w3, w.embed_state.b3, ws.s2, Rta, H, H, H, H, H, false, st));
  CS_TRY(launch_small_mlp1(5, ws.tk.goal_feat, w.embed_goal, ws.g1, (size_t)Ra, st));
  CS_TRY(gemm(ws.g1, w.embed_goal.w3, w.embed_goal.b3, ws.g2, Ra, H, H, H, H, H, false, st));
  CS_TRY(gemm(ws.g2, w.sg_w + H, w.sg_b, ws.gpart, Ra, H, H, H, 2 * H, H, false, st));
  CS_TRY(launch_make_goal_index(G, n_tok, ws.goal_idx, st));
  CS_TRY(gemm(ws.s2, w.sg_w, nullptr, ws.sg, Rta, H, H, H, 2 * H, H, false, st, ws.gpart, ws.goal_idx, H));
  CS_TRY(launch_assemble_tokens(G, n_tok, ws.sg, ws.tk, w.emb, ws.X, with_mem ? ws.mem : nullptr, st));
  return 0;
}

// everything of a decoder layer after self-attention, on R rows of ws.X (attention output in ws.att)
static int decoder_layer_rest(const DecLayerW& d, Workspace& ws, int G, int rows_per_group, const float* kvc,
                              const uint8_t* pad, cudaStream_t st) {
  const int R = G * rows_per_group;
  CS_TRY(gemm_res_ln(ws.att, d.sa.out_w, d.sa.out_b, ws.X, d.n1, ws.tmp, R, H, st));
  CS_TRY(gemm(ws.X, d.ca.in_w, d.ca.in_b, ws.q_c, R, H, H, H, H, H, false, st));
  g_prof.begin(PROF_ATTN_CROSS, (double)G * NH * rows_per_group * MEM * 4.0 * DH, st);
  CS_TRY(launch_attn_padded(ws.q_c, H, kvc, kvc + H, 2 * H, pad, ws.att, H, G, rows_per_group, MEM, st));
  g_prof.end(st);
  CS_TRY(gemm_res_ln(ws.att, d.ca.out_w, d.ca.out_b, ws.X, d.n2, ws.tmp, R, H, st));
  CS_TRY(gemm(ws.X, d.l1w, d.l1b, ws.ff, R, FF, H, H, H, FF, true, st));
  CS_TRY(gemm_res_ln(ws.ff, d.l2w, d.l2b, ws.X, d.n3, ws.tmp, R, FF, st));
  return 0;
}

// Last decoder layer + RTG head for the 24 state rows of the current step only (the only rows of the last layer's
// output that are read): token step ti_tok of the n_tok tokenised steps, window index ti_abs.
static int last_layer_state_rows(const ModelWeights& w, Workspace& ws, int G, int n_tok, int ti_tok, int ti_abs,
                                 const KvView& kv, const float* kvc, const uint8_t* pad, cudaStream_t st) {
  const int Ra = G * A;
  const DecLayerW& d = w.dec[N_DEC - 1];
  // the state rows: position 0 of an agent's step, position 1 in the decision transformer's (rtg, state, action) order,
  // where a state row additionally sees its own rtg token (rule M1: own tokens up to the row itself)
  const int si = w.dt ? 1 : 0;
  CS_TRY(launch_make_row_index(G, n_tok, ti_tok, si, ws.row_idx, st));
  CS_TRY(launch_gather_rows(Ra, ws.X, ws.row_idx, ws.xr, st));
  CS_TRY(gemm(ws.xr, d.sa.in_w, d.sa.in_b, ws.qkv_r, Ra, H, H, H, H, 3 * H, false, st));
  CS_TRY(launch_attn_step(kv, ws.qkv_r, ws.att_r, G, ti_abs, w.dt ? 2 : 0, st, si));
  CS_TRY(gemm(ws.att_r, d.sa.out_w, d.sa.out_b, ws.tmp_r, Ra, H, H, H, H, H, false, st));
  CS_TRY(ln(ws.xr, ws.tmp_r, d.n1, ws.xr, Ra, false, st));
  CS_TRY(gemm(ws.xr, d.ca.in_w, d.ca.in_b, ws.qc_r, Ra, H, H, H, H, H, false, st));
  CS_TRY(launch_attn_padded(ws.qc_r, H, kvc, kvc + H, 2 * H, pad, ws.att_r, H, G, A, MEM, st));
  CS_TRY(gemm(ws.att_r, d.ca.out_w, d.ca.out_b, ws.tmp_r, Ra, H, H, H, H, H, false, st));
  CS_TRY(ln(ws.xr, ws.tmp_r, d.n2, ws.xr, Ra, false, st));
  CS_TRY(gemm(ws.xr, d.l1w, d.l1b, ws.ff_r, Ra, FF, H, H, H, FF, true, st));
  CS_TRY(gemm(ws.ff_r, d.l2w, d.l2b, ws.tmp_r, Ra, H, FF, FF, FF, H, false, st));
  CS_TRY(ln(ws.xr, ws.tmp_r, d.n3, ws.xr, Ra, false, st));
  if (w.dt) {  // decision transformer: the ACTION head reads the state rows (modules/decoder.py:55-57), no RTG head
    CS_TRY(gemm(ws.xr, w.head_action.w0, w.head_action.b0, ws.hd1, Ra, H, H, H, H, H, false, st));
    CS_TRY(launch_layernorm(ws.hd1, nullptr, w.head_action.lnw, w.head_action.lnb, ws.hd1, Ra, H, H, H, true, st));
    CS_TRY(gemm(ws.hd1, w.head_action.w3, w.head_action.b3, ws.act_logits, Ra, N_ACT, H, H, H, N_ACT, false, st));
    return 0;
  }
  // ---- M8 RTG head ------------------------------------------------------------------------------------------------
  CS_TRY(gemm(ws.xr, w.head_rtg.w0, w.head_rtg.b0, ws.hd1, Ra, H, H, H, H, H, false, st));
  CS_TRY(launch_layernorm(ws.hd1, nullptr, w.head_rtg.lnw, w.head_rtg.lnb, ws.hd1, Ra, H, H, H, true, st));
  CS_TRY(gemm(ws.hd1, w.head_rtg.w3, w.head_rtg.b3, ws.rtg_logits, Ra, N_RTG * 3, H, H, H, N_RTG * 3, false, st));
  return 0;
}

static KvView ws_view(const Workspace& ws, int l, int Lcur) {
  KvView v; v.base = ws.QKV[l]; v.ld = 3 * H; v.k_off = H; v.v_off = 2 * H; v.group_rows = Lcur; return v;
}
static KvView cache_view(const PrefixSlot& pc, int l) {
  KvView v; v.base = pc.KV[l]; v.ld = 2 * H; v.k_off = 0; v.v_off = H; v.group_rows = L; return v;
}

// Incremental first pass at step t (1 <= t < 32) of a chunk whose groups are unchanged since step t-1: only the
// tokens of window steps t-1 (whose rtg / action tokens got their final values after step t-1) and t are embedded and
// run through the decoder; they attend to the cached keys / values of all earlier tokens.  Nothing upstream of the
// decoder is recomputed: polyline encoder, scene encoder and cross-attention K/V do not change while the window
// still starts at t = 0.
static int forward_incr(const ModelWeights& w, Workspace& ws, int G, int t, cudaStream_t st, const PrefixSlot& pc) {
  const int n_tok = 2, rows = n_tok * TOK_T, R = G * rows, Lk = (t + 1) * TOK_T, row0 = (t - 1) * TOK_T;
  CS_TRY(embed_tokens(w, ws, G, n_tok, false, st));
  for (int l = 0; l < N_DEC - 1; ++l) {
    const DecLayerW& d = w.dec[l];
    CS_TRY(gemm(ws.X, d.sa.in_w, d.sa.in_b, ws.QKV[l], R, 3 * H, H, H, H, 3 * H, false, st));
    CS_TRY(launch_store_kv(ws.QKV[l] + H, 3 * H, G, rows, pc.KV[l], row0, st));
    {
      double vis = 0;
      for (int tw = t - 1; tw <= t; ++tw) vis += (double)TOK_T * (TOK_T * tw + A) + 3.0 * A;
      g_prof.begin(PROF_ATTN_CAUSAL, (double)G * NH * vis * 4.0 * DH, st);
    }
    CS_TRY(launch_attn_causal_tail(ws.QKV[l], 3 * H, cache_view(pc, l), ws.att, G, rows, Lk, st, w.dt ? 1 : 0));
    g_prof.end(st);
    CS_TRY(decoder_layer_rest(d, ws, G, rows, pc.KVC[l], pc.PAD, st));
  }
  const int l = N_DEC - 1;
  const DecLayerW& d = w.dec[l];
  CS_TRY(gemm(ws.X, d.sa.in_w + (size_t)H * H, d.sa.in_b + H, ws.QKV[l] + H, R, 2 * H, H, H, H, 3 * H, false, st));
  CS_TRY(launch_store_kv(ws.QKV[l] + H, 3 * H, G, rows, pc.KV[l], row0, st));
  return last_layer_state_rows(w, ws, G, n_tok, n_tok - 1, t, cache_view(pc, l), pc.KVC[l], pc.PAD, st);
}

int forward_pass1(const ModelWeights& w, Workspace& ws, int G, int n_t, int n_sm, cudaStream_t st, const MapPlan& mp,
                  const PrefixSlot* pc) {
  if (pc && pc->incr) return forward_incr(w, ws, G, n_t - 1, st, *pc);
  const int Gm = mp.n_map < 0 ? G : mp.n_map;  // groups whose polylines are encoded this step
  const int Rp = Gm * P, Rpt = Rp * NP, Ra = G * A, Rta = G * n_t * A, Rm = G * MEM;
  const int Lcur = n_t * TOK_T, R = G * Lcur, ti = n_t - 1;
  // ---- M2 polyline encoder --------------------------------------------------------------------------------------
  if (Gm > 0) {
  CS_TRY(launch_map_flags(ws.tk.map_pts, ws.pt_valid, ws.poly_valid, Rp, st));
  // "encoder-attn" of BASELINE.json, fused from the raw points (map_encoder.cu): point MLP layer 1 -> scores -> softmax
  // -> pooled hidden vectors; algorithmic bytes per polyline: 100 x 3 fp32 points + validity in, 8 x 256 fp32 out
  g_prof.begin(PROF_MAP_POOL, (double)Rp * (NP * 3 * 4.0 + 1.0 + NH * H * 4.0), st);
  CS_TRY(launch_map_encode_pool(ws.tk.map_pts, w.road_pts, w.pool_U2, ws.poly_valid, ws.pooled, Rp, n_sm, st));
  g_prof.end(st);
  CS_TRY(gemm(ws.pooled, w.pool_W2, w.pool_b2, ws.pe_a, Rp, H, NH * H, NH * H, NH * H, H, false, st));
  CS_TRY(ln(ws.pe_a, nullptr, w.map_n1, ws.pe_b, Rp, false, st));
  CS_TRY(gemm(ws.pe_b, w.map_feats.w0, w.map_feats.b0, ws.pe_a, Rp, H, H, H, H, H, false, st));
  CS_TRY(launch_layernorm(ws.pe_a, nullptr, w.map_feats.lnw, w.map_feats.lnb, ws.pe_a, Rp, H, H, H, true, st));
  CS_TRY(gemm(ws.pe_a, w.map_feats.w3, w.map_feats.b3, ws.pe_c, Rp, H, H, H, H, H, false, st));
  CS_TRY(ln(ws.pe_b, ws.pe_c, w.map_n2, ws.pe_a, Rp, false, st));
  CS_TRY(launch_clamp_type_index(Rp, ws.tk.map_type, ws.type_idx, st));
  CS_TRY(gemm(ws.pe_a, w.rr.w0, nullptr, ws.pe_b, Rp, H, H, H, 2 * H, H, false, st, w.type_tab2, ws.type_idx, H));
  CS_TRY(launch_layernorm(ws.pe_b, nullptr, w.rr.lnw, w.rr.lnb, ws.pe_b, Rp, H, H, H, true, st));
  CS_TRY(gemm(ws.pe_b, w.rr.w3, w.rr.b3, ws.pe_c, Rp, H, H, H, H, H, false, st));
  if (mp.cache_emb) CS_TRY(launch_scatter_map(Gm, ws.pe_c, ws.poly_valid, mp.dst, mp.cache_emb, mp.cache_valid, st));
  }
  // ---- M3 token embeddings --------------------------------------------------------------------------------------
  CS_TRY(embed_tokens(w, ws, G, n_t, true, st));
  if (mp.cache_emb) CS_TRY(launch_build_memory(G, mp.cache_emb, mp.cache_valid, mp.slot, ws.tk, n_t, ws.mem, ws.pad, st));
  else CS_TRY(launch_build_memory(G, ws.pe_c, ws.poly_valid, nullptr, ws.tk, n_t, ws.mem, ws.pad, st));
  if (pc) CS_TRY(launch_copy_bytes(ws.pad, pc->PAD, (size_t)G * MEM, st));
  // ---- M4 scene encoder -----------------------------------------------------------------------------------------
  for (int l = 0; l < N_ENC; ++l) {
    const EncLayerW& e = w.enc[l];
    CS_TRY(gemm(ws.mem, e.sa.in_w, e.sa.in_b, ws.qkv_m, Rm, 3 * H, H, H, H, 3 * H, false, st));
    CS_TRY(launch_attn_padded(ws.qkv_m, 3 * H, ws.qkv_m + H, ws.qkv_m + 2 * H, 3 * H, ws.pad, ws.att_m, H, G, MEM, MEM, st));
    CS_TRY(gemm(ws.att_m, e.sa.out_w, e.sa.out_b, ws.tmp_m, Rm, H, H, H, H, H, false, st));
    CS_TRY(ln(ws.mem, ws.tmp_m, e.n1, ws.mem, Rm, false, st));
    CS_TRY(gemm(ws.mem, e.l1w, e.l1b, ws.ff_m, Rm, FF, H, H, H, FF, true, st));
    CS_TRY(gemm(ws.ff_m, e.l2w, e.l2b, ws.tmp_m, Rm, H, FF, FF, FF, H, false, st));
    CS_TRY(ln(ws.mem, ws.tmp_m, e.n2, ws.mem, Rm, false, st));
  }
  // ---- M5-M7 decoder --------------------------------------------------------------------------------------------
  // Layers 0..2 run on every row (their outputs are the next layer's keys/values).  Of the last layer's output only
  // the 24 state rows of the current step are read (RTG head), so it computes K/V for every row (the second pass
  // needs them) but queries / attention / cross-attention / FFN for those 24 rows only.
  // With a prefix-cache slot the cross-attention K/V go straight into the slot and every layer's K | V rows are
  // copied into it, so that the next step can run incrementally.
  for (int l = 0; l < N_DEC; ++l) {
    const DecLayerW& d = w.dec[l];
    float* kvc = pc ? pc->KVC[l] : ws.kv_c[l];
    CS_TRY(gemm(ws.mem, d.ca.in_w + (size_t)H * H, d.ca.in_b + H, kvc, Rm, 2 * H, H, H, H, 2 * H, false, st));
    if (l < N_DEC - 1) {
      CS_TRY(gemm(ws.X, d.sa.in_w, d.sa.in_b, ws.QKV[l], R, 3 * H, H, H, H, 3 * H, false, st));
    } else {  // K | V of every row into columns [256, 768) of QKV[l]
      CS_TRY(gemm(ws.X, d.sa.in_w + (size_t)H * H, d.sa.in_b + H, ws.QKV[l] + H, R, 2 * H, H, H, H, 3 * H, false, st));
    }
    if (pc) CS_TRY(launch_store_kv(ws.QKV[l] + H, 3 * H, G, Lcur, pc->KV[l], 0, st));
    if (l == N_DEC - 1) break;
    {  // useful flops: 4 * d_h per (row, head, visible key); visible keys per step tw: 72*tw + 24 (+0/1/2 own tokens)
      double vis = 0;
      for (int tw = 0; tw < n_t; ++tw) vis += (double)TOK_T * (TOK_T * tw + A) + 3.0 * A;
      g_prof.begin(PROF_ATTN_CAUSAL, (double)G * NH * vis * 4.0 * DH, st);
    }
    CS_TRY(launch_attn_causal(ws.QKV[l], ws.att, G, n_t, st, w.dt ? 1 : 0));
    g_prof.end(st);
    CS_TRY(decoder_layer_rest(d, ws, G, Lcur, kvc, ws.pad, st));
  }
  return last_layer_state_rows(w, ws, G, n_t, ti, ti, ws_view(ws, N_DEC - 1, Lcur),
                               pc ? pc->KVC[N_DEC - 1] : ws.kv_c[N_DEC - 1], ws.pad, st);
}

// Second pass: the 24 rtg rows of the current step with the sampled RTGs, against the first pass' keys / values
// (workspace or prefix-cache slot).
int forward_pass2(const ModelWeights& w, Workspace& ws, int G, int n_t, cudaStream_t st, const PrefixSlot* pc) {
  if (w.dt) return set_error(-2, "forward_pass2: the decision transformer has a single pass (its action head reads the state rows)");
  const bool incr = pc && pc->incr;
  const int Ra = G * A, ti = n_t - 1, n_tok = incr ? 2 : n_t, Lcur = n_t * TOK_T;
  CS_TRY(launch_assemble_rtg_rows(G, n_tok, n_tok - 1, ws.rtg_new, ws.tk, w.emb, ws.xr, st));
  for (int l = 0; l < N_DEC; ++l) {
    const DecLayerW& d = w.dec[l];
    const float* kvc = pc ? pc->KVC[l] : ws.kv_c[l];
    const uint8_t* pad = incr ? pc->PAD : ws.pad;
    CS_TRY(gemm(ws.xr, d.sa.in_w, d.sa.in_b, ws.qkv_r, Ra, 3 * H, H, H, H, 3 * H, false, st));
    CS_TRY(launch_attn_step(incr ? cache_view(*pc, l) : ws_view(ws, l, Lcur), ws.qkv_r, ws.att_r, G, ti, 1, st));
    CS_TRY(gemm(ws.att_r, d.sa.out_w, d.sa.out_b, ws.tmp_r, Ra, H, H, H, H, H, false, st));
    CS_TRY(ln(ws.xr, ws.tmp_r, d.n1, ws.xr, Ra, false, st));
    CS_TRY(gemm(ws.xr, d.ca.in_w, d.ca.in_b, ws.qc_r, Ra, H, H, H, H, H, false, st));
    CS_TRY(launch_attn_padded(ws.qc_r, H, kvc, kvc + H, 2 * H, pad, ws.att_r, H, G, A, MEM, st));
    CS_TRY(gemm(ws.att_r, d.ca.out_w, d.ca.out_b, ws.tmp_r, Ra, H, H, H, H, H, false, st));
    CS_TRY(ln(ws.xr, ws.tmp_r, d.n2, ws.xr, Ra, false, st));
    CS_TRY(gemm(ws.xr, d.l1w, d.l1b, ws.ff_r, Ra, FF, H, H, H, FF, true, st));
    CS_TRY(gemm(ws.ff_r, d.l2w, d.l2b, ws.tmp_r, Ra, H, FF, FF, FF, H, false, st));
    CS_TRY(ln(ws.xr, ws.tmp_r, d.n3, ws.xr, Ra, false, st));
  }
  // ---- M9 action head on the rtg rows ---------------------------------------------------------------------------
  CS_TRY(gemm(ws.xr, w.head_action.w0, w.head_action.b0, ws.hd1, Ra, H, H, H, H, H, false, st));
  CS_TRY(launch_layernorm(ws.hd1, nullptr, w.head_action.lnw, w.head_action.lnb, ws.hd1, Ra, H, H, H, true, st));
  CS_TRY(gemm(ws.hd1, w.head_action.w3, w.head_action.b3, ws.act_logits, Ra, N_ACT, H, H, H, N_ACT, false, st));
  return 0;
}

}  // namespace ctrlsim
