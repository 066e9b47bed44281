// Internal launcher declarations (host side). Every launcher returns 0 or a negative status and records a message
// retrievable through ctrlsim_last_error().
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ctrlsim {

struct GemmArgs {
  const float* A = nullptr;   // [M,K] row-major, leading dim lda (row gather through agather if set)
  const float* W = nullptr;   // [N,K] row-major (nn.Linear weight), leading dim ldw
  const float* bias = nullptr;
  const float* table = nullptr;  // optional additive rows: C[m] += table[tidx[m]]
  const int* tidx = nullptr;
  const int* agather = nullptr;
  float* C = nullptr;
  int M = 0, N = 0, K = 0, lda = 0, ldw = 0, ldc = 0, ldt = 0;
  bool relu = false;
};
int launch_gemm(const GemmArgs& g, cudaStream_t st);     // dispatches to the tcgen05 kernel unless CTRLSIM_GEMM=simt
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);  // gemm_tc.cu
// X = LayerNorm(X + A W^T + bias) * gamma + beta, N = H = 256, fused (gemm_tc.cu); 1 = not applicable, run the two steps
int launch_gemm_res_ln(const float* A, int lda, const float* W, int ldw, const float* bias, float* X, int ldx,
                       const float* gamma, const float* beta, int M, int K, cudaStream_t st);
// gemm_tc.cu: lo-part copies of registered weights (fetched by TMA instead of being derived per tile)
void gemm_register_weight_lo(const void* owner, const float* base, size_t count, const float* lo);
void gemm_clear_weight_lo(const void* owner);
int launch_weight_lo(const float* w, float* lo, size_t n, cudaStream_t st);
void set_gemm_debug(int v);
void read_gemm_trace(long long* out);
int launch_layernorm(const float* X, const float* R, const float* gamma, const float* beta, float* Y, int M, int ldx,
                     int ldr, int ldy, bool relu, cudaStream_t st);

// attention.cu
int launch_attn_padded(const float* Q, int ldq, const float* Kp, const float* Vp, int ldkv, const uint8_t* key_pad,
                       float* O, int ldo, int G, int Lq, int Lk, cudaStream_t st);
// si: position of the state token inside an agent's step (0 CtRL-Sim order, 1 decision-transformer order)
int launch_attn_causal(const float* QKV, float* O, int G, int n_t, cudaStream_t st, int si = 0);
int launch_attn_padded_mma(const float* Q, int ldq, const float* Kp, const float* Vp, int ldkv, const uint8_t* key_pad,
                           float* O, int ldo, int G, int Lq, int Lk, cudaStream_t st);  // attention_mma.cu
int launch_attn_causal_mma(const float* QKV, float* O, int G, int n_t, cudaStream_t st, int si = 0);
int launch_attn_tc(bool causal, const float* Qbase, int ldq, int q_cols, int q_col0, const float* KVbase, int ldkv,
                   int kv_cols, int k_col0, int v_col0, const uint8_t* key_pad, float* O, int ldo, int G, int Lq, int Lk,
                   cudaStream_t st, int q_pos0 = 0, int kv_group_rows = 0, int si = 0);  // attention_tc.cu
// Where the decoder self-attention keys / values of a chunk live: the first-pass QKV buffer of the workspace
// (ld 768, K at column 256, V at 512, Lcur rows per group) or the prefix cache (ld 512, K at 0, V at 256, 2304 rows).
struct KvView { const float* base = nullptr; int ld = 0, k_off = 0, v_off = 0, group_rows = 0; };
// causal attention of the LAST n_q rows of each group's Lk-row sequence (queries compact in Q: [G * n_q, ldq])
int launch_attn_causal_tail(const float* Q, int ldq, const KvView& kv, float* O, int G, int n_q, int Lk, cudaStream_t st,
                            int si = 0);
// own_mode 0: history + the step's state tokens; 1: + the row's own new key / value (columns [H, 3H) of qkv_rows);
// 2: + the row's own first token of step ti from the K/V buffer (decision transformer: state rows see their rtg token)
int launch_attn_step(const KvView& kv, const float* qkv_rows, float* O, int G, int ti, int own_mode, cudaStream_t st,
                     int si = 0);
int launch_map_pool(const float* feats, const uint8_t* pt_valid, const uint8_t* poly_valid, const float* U,
                    float* pooled, int n_poly, int n_sm, cudaStream_t st);

}  // namespace ctrlsim
