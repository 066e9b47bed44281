// Shared helpers for the sm_100a kernels of the rollout hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace ctrlsim {

// Fixed model geometry of the reference default config (cfgs/model/base.yaml:1-9, cfgs/dataset/waymo/base.yaml:4,38-43).
// The kernels are specialised to these; ctrlsim_create() rejects any other configuration loudly.
// The library is built twice from the same sources: libctrlsim_b200.so with the reference caps (24 agents, 200
// polylines per focal group) and - with -DCTRLSIM_WIDE - libctrlsim_b200_wide.so with the caps of cfgs/dataset/waymo/
// base.yaml:38-39 raised to 64 agents and 256 polylines, so that ONE focal group covers a whole 64-vehicle /
// 256-polyline scene ("wide" variant: 6144 decoder tokens, 320 memory tokens).
constexpr int H = 256;         // hidden_dim
constexpr int NH = 8;          // num_heads
constexpr int DH = 32;         // head dim
constexpr int FF = 1024;       // dim_feedforward
#ifdef CTRLSIM_WIDE
constexpr int A = 64;          // max_num_agents per focal group (wide variant)
#else
constexpr int A = 24;          // max_num_agents per focal group
#endif
constexpr int T = 32;          // train_context_length
constexpr int KT = 3;          // token types (state, rtg, action)
constexpr int TOK_T = A * KT;  // 72 (192) tokens per timestep
constexpr int L = T * TOK_T;   // 2304 (6144) decoder tokens
#ifdef CTRLSIM_WIDE
constexpr int P = 256;         // max_num_road_polylines (wide variant)
#else
constexpr int P = 200;         // max_num_road_polylines
#endif
constexpr int NP = 100;        // points per polyline
constexpr int MEM = P + A;     // 224 (320) memory tokens
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

// Explicit shared-state-space accesses.  The kernels carve dynamic shared memory through an aligned uintptr_t, after
// which the compiler no longer knows the address space and emits GENERIC loads / stores (LD.E / ST.E) - slower and
// queued with global traffic.  Hot shared-memory paths use these instead (addr = __cvta_generic_to_shared(ptr)).
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

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
