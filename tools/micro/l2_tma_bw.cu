// Micro-benchmark: how many bytes per clock can ONE SM pull into shared memory with TMA bulk copies when all 148 SMs
// do the same - the resource the GEMM / attention operand rings compete for.
//   mode 0: every SM streams its OWN 512 KB region again and again (L2-resident, 74 MB in total)
//   mode 1: every SM streams the SAME 512 KB region (weights-like: one W tile wanted by everybody)
//   mode 2: every SM streams its own 64 MB region once (HBM)
//   mode 3: as mode 1, clusters of 2: rank 0 fetches the first half, rank 1 the second half, each MULTICAST to both
//   mode 4: half own (L2), half shared - the GEMM's A + W mix
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_tma_bw l2_tma_bw.cu && ./l2_tma_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int STAGES = 6, CHUNK = 32768;  // 192 KB ring

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "h"(mask) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64, 1) k(const char* __restrict__ buf, size_t own_stride, int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[STAGES];
  const int tid = threadIdx.x;
  uint32_t rank = 0;
  if (MODE == 3) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (MODE == 3) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  const char* own = buf + (size_t)blockIdx.x * own_stride;
  const size_t region = MODE == 2 ? own_stride : 512 * 1024;
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) wait(&full[s], ((it / STAGES) - 1) & 1);  // consume: the ring slot is free again
      if (MODE == 3) {  // both CTAs of the pair must have consumed the slot before either overwrites it
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
      }
      if (it < iters) {
        const size_t off = ((size_t)it * CHUNK) % region;
        expect_tx(&full[s], CHUNK);
        if (MODE == 0 || MODE == 2) bulk(smem + s * CHUNK, own + off, CHUNK, &full[s]);
        else if (MODE == 1) bulk(smem + s * CHUNK, buf + off, CHUNK, &full[s]);
        else if (MODE == 3) bulk_mc(smem + s * CHUNK + rank * (CHUNK / 2), buf + off + rank * (CHUNK / 2), CHUNK / 2, &full[s], 3);
        else { bulk(smem + s * CHUNK, own + off, CHUNK / 2, &full[s]); bulk(smem + s * CHUNK + CHUNK / 2, buf + off, CHUNK / 2, &full[s]); }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

template <int MODE>
void run(const char* name, const char* buf, size_t stride, int iters, long long* d_out, int n_sm) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(MODE == 3 ? (n_sm / 2) * 2 : n_sm);
  cfg.blockDim = dim3(MODE == 3 ? 1 : 64);  // mode 3: one thread per CTA, so the cluster barriers are trivially aligned
  cfg.dynamicSmemBytes = STAGES * CHUNK;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MODE == 3 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, buf, stride, iters, d_out);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("%s: %s / %s\n", name, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[160]; cudaMemcpy(h, d_out, sizeof(long long) * cfg.gridDim.x, cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (unsigned i = 0; i < cfg.gridDim.x; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
    avg /= cfg.gridDim.x;
    const double bytes = (double)iters * CHUNK;
    if (rep == 2)
      printf("%-44s %7.1f B/clk/SM (avg), %7.1f (slowest SM), chip %.2f TB/s by events (%d SMs)\n", name, bytes / avg, bytes / mx,
             bytes * cfg.gridDim.x / (ms * 1e-3) / 1e12, cfg.gridDim.x);
  }
}

int main() {
  int n_sm; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)64 << 20;
  char* buf; cudaMalloc(&buf, big * n_sm); cudaMemset(buf, 1, big * n_sm);
  long long* d_out; cudaMalloc(&d_out, sizeof(long long) * 160);
  run<0>("own 512 KB region per SM (L2)", buf, 512 * 1024, 4096, d_out, n_sm);
  run<1>("same 512 KB region for all SMs (L2)", buf, 0, 4096, d_out, n_sm);
  run<2>("own 64 MB region per SM (HBM)", buf, big, 2048, d_out, n_sm);
  run<3>("same region, cluster of 2, halves multicast", buf, 0, 4096, d_out, n_sm);
  run<4>("half own + half shared (GEMM-like)", buf, 512 * 1024, 4096, d_out, n_sm);
  return 0;
}
