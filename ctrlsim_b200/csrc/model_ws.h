// Scratch layout of one chunk of focal groups (all device memory, carved from the caller's workspace).
#pragma once
#include "model.h"

namespace ctrlsim {

struct Workspace {
  TokenBufs tk;
  // polyline encoder
  float *h1, *feats, *pooled, *pe_a, *pe_b, *pe_c;
  uint8_t *pt_valid, *poly_valid;
  int* type_idx;
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

int forward_pass1(const ModelWeights& w, Workspace& ws, int G, int n_t, int n_sm, cudaStream_t st);
int forward_pass2(const ModelWeights& w, Workspace& ws, int G, int n_t, cudaStream_t st);

int launch_resolve_rtg_range(const CtrlSimBatch& b, const CtrlSimPolicyParams& p, int t, int s0, int s1, int g_base,
                             int steps, const float* rtg_logits, cudaStream_t st);
int launch_gather_rtg_steps(const CtrlSimBatch& b, int g0, int ng, int t, int steps, int* rtg_new, cudaStream_t st);

}  // namespace ctrlsim
