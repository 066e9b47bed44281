// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (clock64 around a loop, 1 CTA per SM, 4 or 8 warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                       \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                       \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),       \
                 "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),     \
                 "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),     \
                 "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
               : "r"(taddr))
#define ST32(taddr, r)                                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                 \
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                      \
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr), \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),   \
               "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),  \
               "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

// mode 0: ld + wait each; 1: two ld then wait; 2: st + wait each; 3: ld, 32 FADDs, st (softmax-like)
__global__ void __launch_bounds__(256, 1) k(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tb + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
  uint32_t r[32], q[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r[i] = threadIdx.x + i; q[i] = i; }
  for (int c = 0; c < 8; ++c) { ST32(base + 32 * c, r); }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + 32 * (it & 3);
    if (mode == 0) {
      LD32(a, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
    } else if (mode == 1) {
      LD32(a, r);
      LD32(a + 128, q);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(r[0]) + __uint_as_float(q[31]);
    } else if (mode == 2) {
      ST32(a, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      LD32(a, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + 1.0f);
      ST32(a + 128, r);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; float* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 4096;
  const char* names[4] = {"ld.x32 + wait", "2 x ld.x32 + wait", "st.x32 + wait", "ld, 32 FADD, st"};
  for (int threads : {128, 256})
    for (int mode = 0; mode < 4; ++mode) {
      k<<<148, threads>>>(mode, iters, d, s);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_instr = (double)h[0] / iters;
      const int warps = threads / 32;
      const double bytes = 32.0 * 32 * 4 * warps * (mode == 1 ? 2 : mode == 3 ? 2 : 1);  // per iteration per SM
      printf("%d warps, %-20s: %8.1f clk/iter  -> %7.1f B/clk/SM  (%s)\n", warps, names[mode], per_instr, bytes / per_instr, cudaGetErrorString(e));
    }
  return 0;
}
