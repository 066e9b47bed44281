// Scratch layout of one chunk of focal groups (all device memory, carved from the caller's workspace).
#pragma once
#include <vector>

#include "model.h"

namespace ctrlsim {

struct Workspace {
  TokenBufs tk;
  // polyline encoder
  float *h1, *feats, *pooled, *pe_a, *pe_b, *pe_c;
  uint8_t *pt_valid, *poly_valid;
  int* type_idx;
  int *map_slot, *map_sel, *map_dst;  // map-cache index lists of the chunk (api.cu)
  // embeddings
  float *s1, *s2, *sg, *g1, *g2, *gpart;
  int* goal_idx;
  // scene encoder
  float *mem, *qkv_m, *att_m, *tmp_m, *ff_m;
  uint8_t* pad;
  // decoder (QKV and kv_c of every layer are kept for the second pass)
  float *X, *QKV[N_DEC], *kv_c[N_DEC], *att, *tmp, *q_c, *ff;
  // heads + second pass
  int *row_idx, *rtg_new;
  float *hd1, *rtg_logits, *act_logits, *xr, *qkv_r, *att_r, *tmp_r, *qc_r, *ff_r;

  // Assign pointers for Gc groups inside [base, base+bytes); base == nullptr only measures. Returns bytes needed.
  size_t carve(void* base, size_t bytes, int Gc);
};

enum { PROF_GEMM, PROF_MAP_POOL, PROF_ATTN_CAUSAL, PROF_ATTN_CROSS, PROF_NCAT };
struct Prof {
  struct Rec { int cat; double work; cudaEvent_t a, b; };
  bool on = false;
  std::vector<cudaEvent_t> pool;
  std::vector<Rec> recs;
  double tot_ms[PROF_NCAT] = {0, 0, 0, 0}, tot_work[PROF_NCAT] = {0, 0, 0, 0};
  long long count[PROF_NCAT] = {0, 0, 0, 0};
  void begin(int cat, double work, cudaStream_t st);
  void end(cudaStream_t st);
  void cancel();  // drop the record begin() just opened (the work was not launched)
  void flush(cudaStream_t st);
};
extern Prof g_prof;

// Which polyline-encoder outputs a chunk uses.  While the window still starts at t = 0 (t < 32) the normalisation
// frame of a focal group is the focal agent's pose at t = 0, so the polyline encoder's output is the same at every
// step; it is computed once per (scene, focal) and kept in a caller-provided cache (ctrlsim_attach_map_cache).
struct MapPlan {
  int n_map = -1;                 // groups to encode this step (their tokens sit in map slots 0..n_map-1); -1 = all G
  const int* slot = nullptr;      // [G] cache block of every group of the chunk; nullptr = group g uses ws block g
  const int* dst = nullptr;       // [n_map] cache block each freshly encoded map is stored to
  float* cache_emb = nullptr;     // [n_blocks, P, H]
  uint8_t* cache_valid = nullptr; // [n_blocks, P]
};

// One chunk slot of the caller-attached prefix cache (ctrlsim_attach_prefix_cache): decoder keys / values of every
// token seen so far, cross-attention keys / values of the memory tokens and the memory padding mask of the chunk's
// groups.  Valid while the 32-step window still starts at t = 0 and the chunk's focal groups do not change.
struct PrefixSlot {
  float* KV[N_DEC];    // [Gc, L, 2H]   K | V rows of decoder layer l
  float* KVC[N_DEC];   // [Gc, MEM, 2H] cross-attention K | V of layer l
  uint8_t* PAD;        // [Gc, MEM]
  bool incr = false;   // true: step t reuses the slot (incremental first pass); false: a full forward (re)fills it
};

int forward_pass1(const ModelWeights& w, Workspace& ws, int G, int n_t, int n_sm, cudaStream_t st,
                  const MapPlan& mp = MapPlan(), const PrefixSlot* pc = nullptr);
int forward_pass2(const ModelWeights& w, Workspace& ws, int G, int n_t, cudaStream_t st, const PrefixSlot* pc = nullptr);

int launch_resolve_rtg_range(const CtrlSimBatch& b, const CtrlSimPolicyParams& p, int t, int s0, int s1, int g_base,
                             int steps, const float* rtg_logits, cudaStream_t st);
int launch_gather_rtg_steps(const CtrlSimBatch& b, int g0, int ng, int t, int steps, int* rtg_new, cudaStream_t st);

}  // namespace ctrlsim
