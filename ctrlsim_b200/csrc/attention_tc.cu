// tcgen05 flash attention (8 heads x d_h = 32, fp32 in / out, fp32-class accuracy via the 3xTF32 split).
//   causal  decoder self-attention with mask rule M1 (utils/train_utils.py:82-130, evaluated arithmetically)
//   padded  attention over the 224 memory tokens with a key-padding mask (encoder self-attn, decoder cross-attn)
//
// One CTA = one (group, head, 128-query tile); key tiles of 64 stream through a 2-stage shared-memory ring.
//   warp 7      TMA: Q tile once, then raw K tiles (cp.async.bulk.tensor.2d, SWIZZLE_128B).  The raw fp32 tiles are
//               used directly as the "hi" operands (kind::tf32 ignores the 13 low mantissa bits).
//   warps 4-5   derive the lo = x - trunc13(x) tiles (Q once, K per tile) and stage V^T hi / lo from global memory
//               (bank-conflict-free: the four values a thread holds are stored in a rotated order)
//   warp 6      MMA issuer.  S = Q K^T: A = Q (smem, K-major), B = K tile (smem, K-major), M=128 N=64, 4 k-steps x 3
//               split products into ONE TMEM accumulator.  The tensor core adds into its fp32 accumulator with
//               truncation, so the order matters: the 8 small cross products (Q_lo K_hi, Q_hi K_lo; 2^-11 of the
//               result) are issued first and the 4 main products last - only 4 truncations happen at full magnitude,
//               the same error as a separate cross accumulator at half the TMEM traffic.  O_tile = P V: A = P read
//               FROM TMEM (FlashAttention-4 style, P never touches shared memory), B = V^T tile (K-major; the producer
//               warps transpose V while splitting it - an MN-major tf32 B operand needs the SWIZZLE_128B_BASE32B
//               layout, which this kernel avoids), 8 k-steps of two instructions: [O_a | O_b] += P_hi [V_hi ; V_lo]
//               (N=64) and O_b += P_lo V_hi (N=32) - P_hi is read from TMEM once instead of twice and the cross products
//               have their own accumulator; fresh accumulators per key tile.  S is double-buffered and the issue order is QK(0) QK(1) PV(0) QK(2) PV(1)
//               ..., so the scores of tile j+1 are ready while the softmax warps still work on tile j; K tiles and V^T
//               tiles are released separately (K right after its QK).
//   warps 0-3   softmax: thread = query row = TMEM lane.  ONE sweep over S per tile: p = ex2(s * scale - ref) against a
//               lazily updated INTEGER reference exponent (FlashAttention-4's conditional rescale): ref only moves when
//               a score exceeds it by more than 2^16; moving it multiplies l, o and the part of P already written by an
//               exact power of two.  P_hi overwrites S in place, P_lo has its own columns.  The tile's P V product is
//               added to the register-resident running output with a rounded fp32 add - ONE TILE LATE: P V of tile j-1
//               runs on the tensor pipe while tile j is swept and is collected right before p_ready(j) is signalled
//               (scaled by the reference moves of that sweep), so a CTA's chain per tile is max(sweep, P V) instead of
//               sweep + P V.  The single P_lo buffer is handed back early: P V issues its P_lo products first and
//               commits them on plo_free.
// TMEM (256 columns per CTA, two CTAs per SM): [0,64) S0 / P_hi, [64,128) S1 / P_hi, [128,192) P_lo, [192,224) O_a (main
// products), [224,256) O_b (cross products).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ctrlsim {

constexpr int AT_QT = 128, AT_KT = 64, AT_STAGES = 2, AT_THREADS = 256;
constexpr uint32_t AT_TMEM_COLS = 256;
constexpr uint32_t AT_S0 = 0, AT_PLO = 128, AT_O = 192;  // S1 = AT_S0 + 64
constexpr float AT_LAZY = 16.f;  // log2 head-room before the softmax reference is moved

struct alignas(1024) AtKV {
  float k_raw[AT_KT * DH];   // [64 keys][32 dims], K-major (TMA, SWIZZLE_128B)
  float k_lo[AT_KT * DH];
  // V^T, K-major, written by warps 4-5: two 8 KB blocks of 32 keys each; block cb = [hi: 32 dims x 32 keys (4 KB)]
  // [lo: 32 dims x 32 keys (4 KB)], so one N = 64 B-operand descriptor spans [V_hi ; V_lo] of a key block
  float vt[2 * DH * AT_KT];
};
struct AtSmem {
  float q_raw[AT_QT * DH];
  float q_lo[AT_QT * DH];
  AtKV kv[AT_STAGES];
  uint64_t q_full, q_lo_ready, k_full[AT_STAGES], k_ready[AT_STAGES], k_empty[AT_STAGES], v_ready[AT_STAGES],
      v_empty[AT_STAGES], s_full[2], p_ready, o_full, plo_free;
  uint32_t tmem_base;
  unsigned long long pad_mask[8];  // padded mode: bit c of word j = key 64 j + c is valid and inside Lk (Lk <= 512)
};

__device__ __forceinline__ uint32_t at_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(at_u32(bar)), "r"(count));
}
__device__ __forceinline__ void at_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(at_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(at_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void at_tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(at_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t at_desc(uint32_t saddr) {  // SWIZZLE_128B, 8-row groups of 1024 B (see gemm_tc.cu)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t at_idesc(int M, int N, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void at_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void at_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void at_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(at_u32(bar)) : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// two 32-column loads in flight before one wait (a tcgen05.ld round trip is ~115 cycles)
__device__ __forceinline__ void at_ld32x2(uint32_t ta, uint32_t (&r)[32], uint32_t tb, uint32_t (&q)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(ta));
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
        "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
        "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
        "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
      : "r"(tb));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void at_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void at_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// multiply 32 TMEM columns of this thread's lane by f (rare softmax-reference move; small pieces keep registers free)
__device__ __noinline__ void at_scale32(uint32_t taddr, float f) {
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    uint32_t t[8];
    at_ld8(taddr + 8 * q, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * f);
    at_st8(taddr + 8 * q, t);
  }
}
__device__ __forceinline__ void at_lo_tile(float* lo, const float* raw, int n_float4, int tid, int nthreads) {
  const uint32_t lo_a = at_u32(lo), raw_a = at_u32(raw);
  for (int i = tid; i < n_float4; i += nthreads) {
    const float4 v = lds128(raw_a + 16u * (uint32_t)i);
    float4 l;
    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    sts128(lo_a + 16u * (uint32_t)i, l);
  }
}
__device__ __forceinline__ float at_ex2(float x) {  // 2^x, flush-to-zero; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ int g_attn_debug = 0;  // 4: record a timeline of CTA (0, 0, 0)
__device__ long long g_attn_trace[8 * 64];  // mode 4: clock64() at 8 events x up to 64 key tiles of CTA (0, 0, 0)
#define AT_TRACE(slot) do { if (trace_on && j < 64) g_attn_trace[j * 8 + (slot)] = clock64(); } while (0)

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, int q_col0, int k_col0,
               const float* __restrict__ Vbase, int ldkv, const uint8_t* __restrict__ key_pad, float* __restrict__ O,
               int ldo, int Lq, int Lk, int q_pos0, int kv_group_rows, int si) {
  // si: position of the state token inside an agent's 3-token step (0 = CtRL-Sim order, 1 = decision transformer)
  // Lq query rows per group (tensor-map rows g * Lq + row); causal mode: row i sits at sequence position q_pos0 + i.
  // Lk keys are valid; the K / V rows of group g start at row g * kv_group_rows of their buffer.
  extern __shared__ unsigned char at_raw[];
  AtSmem& sm = *reinterpret_cast<AtSmem*>((reinterpret_cast<uintptr_t>(at_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = blockIdx.z, h = blockIdx.y;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;  // heavy (late) causal tiles first
  const int r0 = qt * AT_QT;
  int kend = Lk;
  if (CAUSAL) kend = min(Lk, ((q_pos0 + min(r0 + AT_QT - 1, Lq - 1)) / TOK_T + 1) * TOK_T);
  const int n_tiles = (kend + AT_KT - 1) / AT_KT;

  if (tid == 0) {
    at_mbar_init(&sm.q_full, 1); at_mbar_init(&sm.q_lo_ready, 2);
    for (int s = 0; s < AT_STAGES; ++s) {
      at_mbar_init(&sm.k_full[s], 1); at_mbar_init(&sm.k_ready[s], 2); at_mbar_init(&sm.k_empty[s], 1);
      at_mbar_init(&sm.v_ready[s], 2); at_mbar_init(&sm.v_empty[s], 1); at_mbar_init(&sm.s_full[s], 1);
    }
    at_mbar_init(&sm.p_ready, 128); at_mbar_init(&sm.o_full, 1); at_mbar_init(&sm.plo_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!CAUSAL) {  // key-padding bits: thread = key (Lk <= 512), one ballot per warp = one 32-bit half of a tile's mask
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      const int key = tid + rep * AT_THREADS;
      const bool ok = key < Lk && !key_pad[(size_t)g * Lk + key];
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) reinterpret_cast<unsigned*>(sm.pad_mask)[warp + 8 * rep] = bal;
    }
  }
  if (warp == 6) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_u32(&sm.tmem_base)), "r"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = sm.tmem_base;
  const bool trace_on = g_attn_debug == 4 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                        (warp == 0 || warp == 4 || warp == 6);

  if (warp < 4) {
    // ------------------------------------------------------------------ softmax + output (thread = query row)
    const int row = r0 + 32 * warp + lane;
    const bool row_ok = row < Lq;
    int tq = 0;
    // rule M1 for the TOK_T keys of the row's own timestep: every state token (offsets si, si + 3, si + 6, ...), plus the
    // row's own agent's tokens up to the row itself (offsets 3 aq .. 3 aq + kq)
    int aq = 0, kq = 0;
    if (CAUSAL) {
      tq = (q_pos0 + row) / TOK_T;
      const int rem = (q_pos0 + row) - tq * TOK_T;
      aq = rem / KT; kq = rem - aq * KT;
    }
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    const float scale = 0.17677669529663687f * 1.4426950408889634f;  // d_h^-0.5 * log2(e)
    // ref: reference exponent (log2 domain, always an INTEGER so that moving it rescales by an exact power of two) of
    // every p, l and o accumulated so far
    float ref = -INFINITY, l = 0.f, o[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int k0 = j * AT_KT;
      const uint32_t s_addr = lane_addr + AT_S0 + (uint32_t)(j & 1) * AT_KT;  // S of this tile, overwritten by P_hi
      // visibility of the 64 keys of this tile for this row, as a bit mask (bit c <-> key k0 + c)
      unsigned long long okm;
      if (CAUSAL) {
        const int base = tq * TOK_T - k0;  // position of the row's own timestep relative to the tile
        if (!row_ok || base <= -TOK_T) okm = 0ull;           // padding row / the tile lies past the row's timestep
        else if (base >= AT_KT) okm = ~0ull;                  // the tile lies entirely in earlier timesteps
        else {
          const unsigned long long every3 = 0x9249249249249249ull;  // bits 0, 3, 6, ..., 63
          // keys of earlier timesteps (bits below base) are visible; state tokens of the own timestep sit at offsets
          // = si (mod 3) from base
          const int p0 = base + si;                            // first state token of the own timestep, tile-relative
          okm = (base > 0 ? ((1ull << base) - 1ull) : 0ull) |
                (p0 >= AT_KT ? 0ull : (p0 >= 0 ? (every3 << p0) : (every3 >> ((-p0) % 3))));
          const int end = base + TOK_T;                        // first key of the next timestep (> 0 here)
          if (end < AT_KT) okm &= (1ull << end) - 1ull;
          for (int kk = 0; kk <= kq; ++kk) {
            const int c = base + 3 * aq + kk;
            if (c >= 0 && c < AT_KT) okm |= 1ull << c;
          }
        }
      } else {
        okm = sm.pad_mask[j];
      }
      const bool fast = __all_sync(0xffffffffu, okm == ~0ull);
      AT_TRACE(7);
      at_wait(&sm.s_full[j & 1], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      AT_TRACE(0);
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      float fpend = 1.f;  // product of the reference moves of this sweep: the not yet collected O of tile j-1 needs it too
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32], rl[32];
        at_ld32(s_addr + 32 * c, r);
        if (!fast) {
          const unsigned bits = (unsigned)(okm >> (32 * c));
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = ((bits >> i) & 1u) ? r[i] : 0xff800000u;  // -inf: p = 0, ignored by max
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
        const float cmx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale;
        // lazy reference update: only when a score exceeds the reference by more than 2^AT_LAZY (or none exists yet)
        const bool bump = cmx > ref + AT_LAZY || (ref == -INFINITY && cmx > -INFINITY);
        if (__any_sync(0xffffffffu, bump)) {
          const float nref = bump ? ceilf(cmx) : ref;
          const float dref = ref - nref;  // 0, a negative integer, or -inf (no reference before)
          const float f = dref == 0.f ? 1.f : (dref < -126.f ? 0.f : __int_as_float((127 + (int)dref) << 23));  // 2^dref, exact
          l *= f;
          fpend *= f;
#pragma unroll
          for (int i = 0; i < 4; ++i) s4[i] *= f;
#pragma unroll
          for (int i = 0; i < DH; ++i) o[i] *= f;
          ref = nref;
          if (c == 1) {  // the first half of this tile's P was written against the old reference: rescale it in place
            at_scale32(s_addr, f);
            at_scale32(lane_addr + AT_PLO, f);
          }
        }
        const float rr = ref == -INFINITY ? 0.f : ref;  // no visible key yet: every p below is ex2(-inf) = 0
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float p = at_ex2(fmaf(__uint_as_float(r[i]), scale, -rr));
          s4[i & 3] += p;
          const uint32_t hi = __float_as_uint(p) & 0xFFFFE000u;
          r[i] = hi;
          rl[i] = __float_as_uint(p - __uint_as_float(hi));
        }
        at_st32(s_addr + 32 * c, r);
        if (c == 0 && j > 0) {  // the single P_lo buffer is free once the P_lo products of tile j-1 (issued first) are done
          at_wait(&sm.plo_free, (j - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        at_st32(lane_addr + AT_PLO + 32 * c, rl);
      }
      l += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      AT_TRACE(1);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (j > 0) {  // collect P V of tile j-1 (it ran while this tile was swept); PV(j) may overwrite O only afterwards
        at_wait(&sm.o_full, (j - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        AT_TRACE(3);
        uint32_t r[32], rb[32];
        at_ld32x2(lane_addr + AT_O, r, lane_addr + AT_O + DH, rb);
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = fmaf(__uint_as_float(r[i]) + __uint_as_float(rb[i]), fpend, o[i]);
        AT_TRACE(4);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      at_arrive(&sm.p_ready);
      AT_TRACE(2);
    }
    if (n_tiles > 0) {  // P V of the last tile
      at_wait(&sm.o_full, (n_tiles - 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32], rb[32];
      at_ld32x2(lane_addr + AT_O, r, lane_addr + AT_O + DH, rb);
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] += __uint_as_float(r[i]) + __uint_as_float(rb[i]);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (row_ok) {
      const float inv = 1.0f / l;
      float* dst = O + ((size_t)g * Lq + row) * ldo + h * DH;
#pragma unroll
      for (int i = 0; i < DH; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ lo producers (64 threads)
    const int pt = tid - 128;
    at_wait(&sm.q_full, 0);
    at_lo_tile(sm.q_lo, sm.q_raw, AT_QT * DH / 4, pt, 64);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) at_arrive(&sm.q_lo_ready);
    const int u = pt & 7, hrot = u >> 1;  // this thread's 4 dims 4u..4u+3 of V; store order rotated by hrot (bank spread)
    const float* vsrc = Vbase + (size_t)g * kv_group_rows * ldkv + h * DH + u * 4;
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % AT_STAGES;
      float4 vv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // V rows of this tile: key = j*64 + pt/8 + 8 i, dims 4u .. 4u+3
        const int key = j * AT_KT + (pt >> 3) + 8 * i;
        vv[i] = key < Lk ? *reinterpret_cast<const float4*>(vsrc + (size_t)key * ldkv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      at_wait(&sm.k_full[s], (j / AT_STAGES) & 1);
      at_lo_tile(sm.kv[s].k_lo, sm.kv[s].k_raw, AT_KT * DH / 4, pt, 64);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_arrive(&sm.k_ready[s]);
      at_wait(&sm.v_empty[s], ((j / AT_STAGES) & 1) ^ 1);
      const uint32_t vt_a = at_u32(sm.kv[s].vt);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int kl = (pt >> 3) + 8 * i;                 // key within the tile
        const int cb = kl >> 5, kk = kl & 31;
        // rotate the four values by hrot so that in every store instruction the 8 threads sharing a key hit 8
        // different (d & 7) classes: with the 128-byte swizzle below the warp's 32 stores land in 32 distinct banks
        float x0 = vv[i].x, x1 = vv[i].y, x2 = vv[i].z, x3 = vv[i].w;
        if (hrot & 1) { const float t0 = x0; x0 = x1; x1 = x2; x2 = x3; x3 = t0; }
        if (hrot & 2) { const float t0 = x0, t1 = x1; x0 = x2; x1 = x3; x2 = t0; x3 = t1; }
        const float x[4] = {x0, x1, x2, x3};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int d = u * 4 + ((n + hrot) & 3);
          const int off = cb * (2 * DH * 32) + d * 32 + ((((kk >> 2) ^ (d & 7)) << 2) | (kk & 3));
          sts32(vt_a + 4u * (uint32_t)off, x[n]);
          sts32(vt_a + 4u * (uint32_t)(off + DH * 32), x[n] - __uint_as_float(__float_as_uint(x[n]) & 0xFFFFE000u));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) at_arrive(&sm.v_ready[s]);
      AT_TRACE(5);
    }
  } else if (warp == 6) {
    // ------------------------------------------------------------------ MMA issuer
    // Issue order QK(0), QK(1), PV(0), QK(2), PV(1), ...: S is double-buffered, so the scores of tile j+1 are computed
    // while the softmax warps work on tile j.
    const uint32_t id_qk = at_idesc(AT_QT, AT_KT, false), id_pv = at_idesc(AT_QT, DH, false), id_pv2 = at_idesc(AT_QT, 2 * DH, false);
    at_wait(&sm.q_lo_ready, 0);
    const uint64_t dqh = at_desc(at_u32(sm.q_raw)), dql = at_desc(at_u32(sm.q_lo));
    auto issue_qk = [&](int i) {
      const int s = i % AT_STAGES;
      at_wait(&sm.k_ready[s], (i / AT_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t d_s = tmem + AT_S0 + (uint32_t)(i & 1) * AT_KT;
        const uint64_t dkh = at_desc(at_u32(sm.kv[s].k_raw)), dkl = at_desc(at_u32(sm.kv[s].k_lo));
#pragma unroll
        for (int ks = 0; ks < DH / 8; ++ks) {  // small cross products first (see the header comment)
          const uint64_t o2 = (uint64_t)(2 * ks);
          at_mma_ss(d_s, dql + o2, dkh + o2, id_qk, ks > 0 ? 1u : 0u);
          at_mma_ss(d_s, dqh + o2, dkl + o2, id_qk, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < DH / 8; ++ks) {
          const uint64_t o2 = (uint64_t)(2 * ks);
          at_mma_ss(d_s, dqh + o2, dkh + o2, id_qk, 1u);
        }
        at_commit(&sm.k_empty[s]);
        at_commit(&sm.s_full[i & 1]);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % AT_STAGES;
      if (j + 1 < n_tiles) issue_qk(j + 1);
      at_wait(&sm.v_ready[s], (j / AT_STAGES) & 1);
      at_wait(&sm.p_ready, j & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      AT_TRACE(6);
      if (lane == 0) {
        const uint32_t a_hi = tmem + AT_S0 + (uint32_t)(j & 1) * AT_KT, a_lo = tmem + AT_PLO;
        const uint64_t dv = at_desc(at_u32(sm.kv[s].vt));
        // V^T is K-major: 8 keys = 32 bytes inside a 128-byte atom row; key block cb (32 keys) = 8 KB [V_hi ; V_lo].
        // Per k-step two instructions: [O_a | O_b] += P_hi x [V_hi ; V_lo] (N = 64, P_hi read once) and O_b += P_lo x V_hi
        // (N = 32): O_a collects the 8 main products, O_b the 16 small cross products; they are added in registers.
        // Order: the first N = 64 product zeroes both accumulators, then ALL P_lo products (committed on plo_free, so
        // that the softmax warps can write the P_lo of tile j+1 while the remaining products still run), then the rest.
        auto vt_off = [](int ks) { return (uint64_t)((ks >> 2) * ((2 * DH * 128) >> 4) + (ks & 3) * 2); };
        at_mma_ts(tmem + AT_O, a_hi, dv + vt_off(0), id_pv2, 0u);
#pragma unroll
        for (int ks = 0; ks < AT_KT / 8; ++ks) at_mma_ts(tmem + AT_O + DH, a_lo + 8 * ks, dv + vt_off(ks), id_pv, 1u);
        at_commit(&sm.plo_free);
#pragma unroll
        for (int ks = 1; ks < AT_KT / 8; ++ks) at_mma_ts(tmem + AT_O, a_hi + 8 * ks, dv + vt_off(ks), id_pv2, 1u);
        at_commit(&sm.v_empty[s]);
        at_commit(&sm.o_full);
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ------------------------------------------------------------------ TMA issuer
    if (lane == 0) {
      at_expect_tx(&sm.q_full, AT_QT * DH * 4);
      at_tma_2d(sm.q_raw, &tmQ, q_col0 + h * DH, g * Lq + r0, &sm.q_full);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % AT_STAGES;
        at_wait(&sm.k_empty[s], ((j / AT_STAGES) & 1) ^ 1);
        at_expect_tx(&sm.k_full[s], AT_KT * DH * 4);
        at_tma_2d(sm.kv[s].k_raw, &tmKV, k_col0 + h * DH, g * kv_group_rows + j * AT_KT, &sm.k_full[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 6) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(AT_TMEM_COLS) : "memory");
  }
}

void set_attn_debug(int v) { cudaMemcpyToSymbol(g_attn_debug, &v, sizeof(int)); }
void read_attn_trace(long long* out) { cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * 8 * 64); }

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int at_make_map(AtEncodeFn enc, CUtensorMap* map, const float* base, long long rows, int cols, int ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)DH, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-5, "attn_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// Q rows: [G*Lq, q_cols] at Qbase (ldq), head h at columns q_col0 + 32h; K / V rows: [G*Lk, kv_cols] at KVbase (ldkv).
int launch_attn_tc(bool causal, const float* Qbase, int ldq, int q_cols, int q_col0, const float* KVbase, int ldkv,
                   int kv_cols, int k_col0, int v_col0, const uint8_t* key_pad, float* O, int ldo, int G, int Lq, int Lk,
                   cudaStream_t st, int q_pos0, int kv_group_rows, int si) {
  if (kv_group_rows <= 0) kv_group_rows = Lk;
  static AtEncodeFn enc = nullptr;
  const int smem = (int)sizeof(AtSmem) + 1024;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return set_error(-5, "attn_tc: cuTensorMapEncodeTiled unavailable");
    e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(-5, "attn_tc smem attr: %s", cudaGetErrorString(e));
    enc = reinterpret_cast<AtEncodeFn>(fn);
  }
  if (!causal && Lk > 512) return set_error(-2, "attn_tc: padded mode supports at most 512 keys");
  CUtensorMap tmQ, tmKV;
  int rc;
  if ((rc = at_make_map(enc, &tmQ, Qbase, (long long)G * Lq, q_cols, ldq, AT_QT))) return rc;
  if ((rc = at_make_map(enc, &tmKV, KVbase, (long long)G * kv_group_rows, kv_cols, ldkv, AT_KT))) return rc;
  dim3 grid((Lq + AT_QT - 1) / AT_QT, NH, G);
  if (causal)
    attn_tc_kernel<true><<<grid, AT_THREADS, smem, st>>>(tmQ, tmKV, q_col0, k_col0, KVbase + v_col0, ldkv, key_pad, O, ldo, Lq, Lk, q_pos0, kv_group_rows, si);
  else
    attn_tc_kernel<false><<<grid, AT_THREADS, smem, st>>>(tmQ, tmKV, q_col0, k_col0, KVbase + v_col0, ldkv, key_pad, O, ldo, Lq, Lk, 0, kv_group_rows, 0);
  CS_CHECK_LAUNCH("attn_tc");
  return 0;
}

}  // namespace ctrlsim
