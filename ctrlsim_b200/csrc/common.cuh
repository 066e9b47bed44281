// Shared helpers for the sm_100a kernels of the rollout hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace ctrlsim {

// Fixed model geometry of the reference default config (cfgs/model/base.yaml:1-9, cfgs/dataset/waymo/base.yaml:4,38-43).
// The kernels are specialised to these; ctrlsim_create() rejects any other configuration loudly.
constexpr int H = 256;         // hidden_dim
constexpr int NH = 8;          // num_heads
constexpr int DH = 32;         // head dim
constexpr int FF = 1024;       // dim_feedforward
constexpr int A = 24;          // max_num_agents per focal group
constexpr int T = 32;          // train_context_length
constexpr int KT = 3;          // token types (state, rtg, action)
constexpr int TOK_T = A * KT;  // 72 tokens per timestep
constexpr int L = T * TOK_T;   // 2304 decoder tokens
constexpr int P = 200;         // max_num_road_polylines
constexpr int NP = 100;        // points per polyline
constexpr int MEM = P + A;     // 224 memory tokens
constexpr int N_ACT = 1000;    // 20 x 50 action bins
constexpr int N_RTG = 350;     // rtg bins per component
constexpr int N_ENC = 2;       // encoder layers
constexpr int N_DEC = 4;       // decoder layers
constexpr float LN_EPS = 1e-5f;

extern thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...);

extern long long g_launch_count;  // kernels launched by this library (bench.py reports it as gpu_launches)

#define CS_CHECK_LAUNCH(name)                                                          \
  do {                                                                                 \
    ++::ctrlsim::g_launch_count;                                                       \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) return ::ctrlsim::set_error(-5, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ctrlsim
