// tc_common.cuh -- inline-PTX wrappers for the Blackwell tensor-core path (sm_100a):
// mbarrier, proxy fences, TMEM allocation, tcgen05.mma / commit / ld, UMMA descriptors.
#pragma once
#include <cstdint>
#include <cstdio>

#include "common.cuh"

namespace mog {
namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // bf16 elements per smem stage row = one 128-byte swizzle row
constexpr int NPROD = 128;       // producer / epilogue threads (warps 0-3)
constexpr int NTHREADS = 160;    // + warp 4: MMA issuer and TMEM owner
constexpr int MAX_STAGES = 6;

struct TcWeightLayout {
  int BN, ntiles, Npad, K, Kpad, planes;
  size_t plane_elems;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("libmog: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// ---- cp.async (16-byte global -> shared copies; src_bytes < 16 zero-fills the rest) ------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- fences -----------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------
// Executed by one full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp receives TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Asynchronous variant: the registers are valid only after tmem_ld_wait(r) (which names them as in/out operands
// so the compiler cannot schedule a use above the wait).
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// Epilogue math on a group of 16 consecutive output channels of one row: v = act(acc + prev + bias).
// The activation switch is uniform and sits OUTSIDE the element loop (a per-element switch costs ~20
// instructions and an indirect branch per value and made the epilogue the bottleneck of the conv kernels).
__device__ __forceinline__ void epi_act16(float (&o)[16], int act) {
  switch (act) {
    case MOG_ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : 0.2f * o[j];
      break;
    case MOG_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
      break;
    case MOG_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = tanhf(o[j]);
      break;
    case MOG_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = 1.0f / (1.0f + expf(-o[j]));
      break;
    default:
      break;
  }
}
// Stores one 16-channel group of an output row: dptr points at channel c of the row, `nvalid` channels exist from
// there on (>= 16 for interior groups).  vec: the row pitch and c allow 16-byte accesses.
__device__ __forceinline__ void epi_store16(const uint32_t (&acc)[16], float* dptr, const float* bias_c, int nvalid, bool vec,
                                            bool accum_dst, int act) {
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(acc[j]);
  const bool full = vec && nvalid >= 16;
  if (accum_dst) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(dptr + j);
        o[j] += pv.x; o[j + 1] += pv.y; o[j + 2] += pv.z; o[j + 3] += pv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] += dptr[j];
    }
  }
  if (bias_c) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias_c + j));
        o[j] += bv.x; o[j + 1] += bv.y; o[j + 2] += bv.z; o[j + 3] += bv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] += __ldg(bias_c + j);
    }
  }
  if (act != MOG_ACT_NONE) epi_act16(o, act);
  if (full) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dptr + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dptr[j] = o[j];
  }
}

// ---- UMMA descriptors ----------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, densely packed 8-row atoms:
//   [0,14) start address >> 4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) |
//   [32,46) SBO>>4 (= 1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand, SWIZZLE_128B: 64 MN-elements (128 B) contiguous per k row, 8 k rows per 1024-byte
// atom; LBO = byte distance between 64-element MN blocks, SBO = byte distance between 8-k groups.
__device__ __forceinline__ uint64_t make_desc_sw128_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16, A=B=bf16, D=fp32, dense:
//   [4,6) c_format=1 (f32) | [7,10) a_format=1 (bf16) | [10,13) b_format=1 | [15] a_major | [16] b_major |
//   [17,23) N>>3 | [24,29) M>>4
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// ---- warp-convergent issue ---------------------------------------------------------------------
// The MMA warp runs its loops with ALL lanes (warp-uniform values stay in uniform registers, which is what
// UTCHMMA reads); one elected lane executes the tcgen05 instruction itself.  Wrapping the loop in
// `if (lane == 0)` instead makes the compiler move every descriptor / TMEM address from vector to uniform
// registers through an ELECT / R2UR.BROADCAST / BRA.U.ANY loop per MMA (~100+ cycles: issue-bound at N <= 128).
// 64-bit descriptors are passed as (lo, hi) words: hi is constant per layout, lo = start address field | LBO.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
// hi word: SBO>>4 | version 1 (bit 46) | layout (bits 61-63; 2 = SWIZZLE_128B)
__device__ __forceinline__ uint32_t desc_hi_sw128(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05 async ops of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace tc
}  // namespace mog
