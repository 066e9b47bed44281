// Micro-benchmark: issue rate of tcgen05.mma kind::tf32 (K = 8) and kind::f16 (K = 16) with shared-memory operands,
// M = 128, one CTA per SM: cycles per instruction for N = 256 / 128 / 64, back to back on resident operand tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind: 0 tf32 (a/b format 2), 1 bf16 (a/b format 1)
__device__ __forceinline__ uint32_t mk_idesc(int kind, int M, int N) {
  const uint32_t fmt = kind == 0 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int KIND>
__global__ void __launch_bounds__(128, 1) k(int N, int n_acc, int iters, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  float* A = (float*)base;                 // 128 rows x 128 B
  float* B = (float*)(base + 16384);       // 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint64_t da = mk_desc(s32(A)), db = mk_desc(s32(B));
    const uint32_t id = mk_idesc(KIND, 128, N);
    t0 = clock64();
    const uint32_t d0 = tb, d1 = tb + (uint32_t)((n_acc - 1) * N);
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // no address arithmetic on the issue path: 8 instructions, constants only
        const uint32_t d = (u & 1) ? d1 : d0;
        const uint64_t o = (uint64_t)(2 * (u & 3));
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da + o), "l"(db + o), "r"(id), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da + o), "l"(db + o), "r"(id), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 16384 + 32768 + 1024, iters = 4096;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int kind = 0; kind < 2; ++kind)
    for (int N : {256, 128, 64, 32})
      for (int n_acc : {1, 2}) {
        if (N * n_acc > 512) continue;
        if (kind == 0) k<0><<<148, 128, smem>>>(N, n_acc, iters, d); else k<1><<<148, 128, smem>>>(N, n_acc, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double clk = (double)h[0] / iters;
        const int K = kind == 0 ? 8 : 16;
        printf("%s M=128 N=%3d K=%2d, %d accumulator(s): %7.1f clk/MMA -> %7.0f MAC/clk/SM  (%s)\n", kind == 0 ? "tf32" : "bf16", N, K, n_acc, clk, 128.0 * N * K / clk, cudaGetErrorString(e));
      }
  return 0;
}
