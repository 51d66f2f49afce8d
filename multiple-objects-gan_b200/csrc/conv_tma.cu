// conv_tma.cu -- persistent, warp-specialised tcgen05 convolution with TMA staging (sm_100a).
//
// Serves the unit-stride gather-GEMM problems (stride-1 forward convs, every stride phase of a data
// gradient) whose operand is available as pre-split bf16 planes [N][H][W][C8]:
//
//   * im2col by coordinates: an M-tile of 128 output pixels is a (bw x bh x bn) box of the
//     (W, H, N) grid; for tap (kh, kw) and channel chunk c0 ONE `cp.async.bulk.tensor.4d` loads the
//     box {64 ch, bw, bh, bn} at (c0, w0+kw-pad, h0+kh-pad, n0) straight into the K-major
//     SWIZZLE_128B layout the tensor core reads; out-of-range coordinates (the conv padding) are
//     zero-filled by the TMA unit.  No index arithmetic, no per-thread copies.
//   * weights: one 2-D TMA box {64 k, BN rows} per plane per stage.
//   * warp 0 = TMA producer (one thread), warp 1 = tcgen05.mma issuer (one thread) + TMEM owner,
//     warps 2-5 = epilogue.  smem full/empty mbarrier ring (expect_tx / tcgen05.commit); the fp32
//     accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
//     the main loop of tile i+1; CTAs are persistent (grid = #SMs, static tile schedule).
//   * MOG_PREC_BF16X3: three MMAs per k-step (hi*hi, lo*hi, hi*lo) on the hi/lo planes.
#include <cuda.h>
#include <cuda_bf16.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mog {
namespace tc {

constexpr int TMA_THREADS = 192;   // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

struct TmaConvParams {
  int bw, bh, bn;                 // pixel box of one M tile: bw*bh*bn == 128
  int tiles_w, tiles_h, tiles_n;  // tile grid over (W, H, N)
  int n_ntiles, BN;               // output-channel tiles
  int nth, ntw, nchunk;           // taps and 64-channel chunks per tap
  int off_h[8], off_w[8];
  int passes, stages, tmem_cols;
  float* dst;
  const float* bias;
  int act, accum_dst;
  int N, Cd, Hd, Wd, dsh, doh, dsw, dow;
  long long total_tiles;
};

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__global__ void __launch_bounds__(TMA_THREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TmaConvParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nplanes = p.passes == 3 ? 2 : 1;
  const int a_plane = BM * 128;
  const int b_plane = p.BN * 128;
  const int stage_bytes = nplanes * (a_plane + b_plane);
  unsigned char* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tfull = empty + MAX_STAGES;   // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmBh);
    if (nplanes == 2) {
      tma_prefetch_desc(&tmAl);
      tma_prefetch_desc(&tmBl);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nk = p.nth * p.ntw * p.nchunk;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0) {
    // ===================== TMA producer ==========================================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ntile = (int)(tile % p.n_ntiles);
        long long mt = tile / p.n_ntiles;
        const int tw_ = (int)(mt % p.tiles_w);
        const int th_ = (int)((mt / p.tiles_w) % p.tiles_h);
        const int tn_ = (int)(mt / tiles_per_img);
        const int w0 = tw_ * p.bw, h0 = th_ * p.bh, n0 = tn_ * p.bn;
        int kc = 0;
        for (int a = 0; a < p.nth; ++a) {
          for (int b = 0; b < p.ntw; ++b) {
            for (int ch = 0; ch < p.nchunk; ++ch, ++kc) {
              mbar_wait(&empty[s], ph ^ 1u);
              const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
              mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
              tma_load_4d(st, &tmAh, &full[s], ch * 64, w0 + p.off_w[b], h0 + p.off_h[a], n0);
              if (nplanes == 2) tma_load_4d(st + a_plane, &tmAl, &full[s], ch * 64, w0 + p.off_w[b], h0 + p.off_h[a], n0);
              const uint32_t sb = st + nplanes * a_plane;
              tma_load_2d(sb, &tmBh, &full[s], kc * 64, ntile * p.BN);
              if (nplanes == 2) tma_load_2d(sb + b_plane, &tmBl, &full[s], kc * 64, ntile * p.BN);
              if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ============================================================
    const uint32_t idesc = make_idesc_bf16(BM, p.BN);
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aph ^ 1u);       // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(as * p.BN);
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full[s], ph);
        tcgen05_fence_after();
        // all lanes run this (uniform registers); one elected lane issues inside umma_bf16_elect
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = desc_lo(st, 16), a_lo = desc_lo(st + a_plane, 16);
        const uint32_t b_hi = desc_lo(st + nplanes * a_plane, 16), b_lo = desc_lo(st + nplanes * a_plane + b_plane, 16);
        const uint32_t dhi = desc_hi_sw128(1024);
#pragma unroll
        for (int k16 = 0; k16 < BK / 16; ++k16) umma_bf16_elect(tacc, a_hi + 2 * k16, dhi, b_hi + 2 * k16, dhi, idesc, (kc | k16) != 0);
        if (p.passes == 3) {
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) umma_bf16_elect(tacc, a_lo + 2 * k16, dhi, b_hi + 2 * k16, dhi, idesc, 1u);   // lo*hi
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) umma_bf16_elect(tacc, a_hi + 2 * k16, dhi, b_lo + 2 * k16, dhi, idesc, 1u);   // hi*lo
        }
        umma_commit_elect(&empty[s]);
        if (kc == nk - 1) umma_commit_elect(&tfull[as]);
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  } else {
    // ===================== epilogue (warps 2-5) ==================================================
    const int q4 = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;       // tile row
    const int wl = r % p.bw;
    const int hl = (r / p.bw) % p.bh;
    const int nl = r / (p.bw * p.bh);
    int as = 0;
    uint32_t aph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int ntile = (int)(tile % p.n_ntiles);
      long long mt = tile / p.n_ntiles;
      const int tw_ = (int)(mt % p.tiles_w);
      const int th_ = (int)((mt / p.tiles_w) % p.tiles_h);
      const int tn_ = (int)(mt / tiles_per_img);
      const int n = tn_ * p.bn + nl, rh = th_ * p.bh + hl, rw = tw_ * p.bw + wl;
      const bool ok = n < p.N;
      const int n0c = ntile * p.BN;
      float* dptr = nullptr;
      if (ok) {
        const size_t pix = ((size_t)n * p.Hd + (rh * p.dsh + p.doh)) * p.Wd + (rw * p.dsw + p.dow);
        dptr = p.dst + pix * p.Cd + n0c;
      }
      mbar_wait(&tfull[as], aph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * p.BN);
      // software-pipelined TMEM reads: the load of group g+1 is in flight while group g is stored
      const int ngrp = p.BN / 16;
      const bool vec = (p.Cd & 3) == 0;
      uint32_t acc[2][16];
      tmem_ld16_async(taddr, acc[0]);
#pragma unroll 1
      for (int gq = 0; gq < ngrp; gq += 2) {
        tmem_ld_wait(acc[0]);
        if (gq + 1 < ngrp) tmem_ld16_async(taddr + (uint32_t)((gq + 1) * 16), acc[1]);
        if (ok && n0c + gq * 16 < p.Cd)
          epi_store16(acc[0], dptr + gq * 16, p.bias ? p.bias + n0c + gq * 16 : nullptr, p.Cd - n0c - gq * 16, vec, p.accum_dst != 0, p.act);
        if (gq + 1 < ngrp) {
          tmem_ld_wait(acc[1]);
          if (gq + 2 < ngrp) tmem_ld16_async(taddr + (uint32_t)((gq + 2) * 16), acc[0]);
          if (ok && n0c + (gq + 1) * 16 < p.Cd)
            epi_store16(acc[1], dptr + (gq + 1) * 16, p.bias ? p.bias + n0c + (gq + 1) * 16 : nullptr, p.Cd - n0c - (gq + 1) * 16, vec,
                        p.accum_dst != 0, p.act);
        }
      }
      // all TMEM reads of this warp are complete (every load was waited for): release the accumulator
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aph ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// pitch (in channels) of one tap in the K dimension of the weights packed for this kernel
int tma_tap_pitch(int Cs) { return ceil_div(Cs, 64) * 64; }

// shape-only test (the weight packing depends on it); Cs is the channel count of the gathered tensor
bool tma_shape_eligible(const IGemmParams& g) {
  if (g.rs != 1 || g.up2x) return false;
  if (g.vstep > 1 && ((g.Hp % g.vstep) || (g.Wp % g.vstep))) return false;
  if (!is_pow2(g.Hr) || !is_pow2(g.Wr)) return false;
  if (g.Wr > 128 && (g.Wr % 128)) return false;
  if (g.nth < 1 || g.ntw < 1) return false;
  if (g.Cd < 1) return false;
  if (g.M < 64LL * 128) return false;   // small problems: the split-K kernel fills the machine better
  return encode_fn() != nullptr;
}

int launch_igemm_tma(const IGemmParams& g, const void* packed, int passes, cudaStream_t st) {
  const int accum_dst = g.accum_dst;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (!g.src_planes || (g.Cs % 8)) return fail(MOG_ERR_BAD_ARG, "conv (TMA): needs pre-split planes with a channel pitch multiple of 8");
  const int nplanes = passes == 3 ? 2 : 1;
  TmaConvParams p{};
  p.bw = g.Wr < 128 ? g.Wr : 128;
  p.bh = g.Hr < 128 / p.bw ? g.Hr : 128 / p.bw;
  p.bn = 128 / (p.bw * p.bh);
  p.tiles_w = g.Wr / p.bw;
  p.tiles_h = g.Hr / p.bh;
  p.tiles_n = ceil_div(g.N, p.bn);
  p.BN = tc_bn_for(g.Cd);
  p.n_ntiles = ceil_div(g.Cd, p.BN);
  p.nth = g.nth; p.ntw = g.ntw;
  p.nchunk = ceil_div(g.Cs, 64);
  for (int i = 0; i < 8; ++i) { p.off_h[i] = g.off_h[i]; p.off_w[i] = g.off_w[i]; }
  p.passes = passes;
  p.dst = g.dst; p.bias = g.bias; p.act = g.act; p.accum_dst = accum_dst;
  p.N = g.N; p.Cd = g.Cd; p.Hd = g.Hd; p.Wd = g.Wd; p.dsh = g.dsh; p.doh = g.doh; p.dsw = g.dsw; p.dow = g.dow;
  p.total_tiles = (long long)p.tiles_n * p.tiles_h * p.tiles_w * p.n_ntiles;
  const int stage_bytes = nplanes * (BM * 128 + p.BN * 128);
  int stages = (208 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return fail(MOG_ERR_UNSUPPORTED, "conv (TMA): stage of %d bytes does not fit twice", stage_bytes);
  p.stages = stages;
  int cols = 32;
  while (cols < 2 * p.BN) cols *= 2;
  p.tmem_cols = cols;

  // ---- tensor maps
  const int Kpad = g.nth * g.ntw * p.nchunk * 64;
  const int Npad = p.n_ntiles * p.BN;
  CUtensorMap tmA[2], tmB[2];
  const __nv_bfloat16* xa = static_cast<const __nv_bfloat16*>(g.src_planes);
  const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(packed);
  for (int pl = 0; pl < 2; ++pl) {
    const int src = pl < nplanes ? pl : 0;   // unused maps alias plane 0 (never dereferenced)
    {
      // (strided) view of the physical [N][Hp][Wp][Cs] plane: logical (h, w) = physical (h*vs + voh, w*vs + vow)
      const int vs = g.vstep > 0 ? g.vstep : 1, Hp = g.vstep > 0 ? g.Hp : g.Hs, Wp = g.vstep > 0 ? g.Wp : g.Ws;
      cuuint64_t dims[4] = {(cuuint64_t)g.Cs, (cuuint64_t)g.Ws, (cuuint64_t)g.Hs, (cuuint64_t)g.N};
      cuuint64_t strides[3] = {(cuuint64_t)vs * g.Cs * 2, (cuuint64_t)vs * Wp * g.Cs * 2, (cuuint64_t)Hp * Wp * g.Cs * 2};
      cuuint32_t box[4] = {64u, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
      cuuint32_t es[4] = {1, 1, 1, 1};
      void* base = const_cast<__nv_bfloat16*>(xa + (size_t)src * g.src_plane_elems + ((size_t)g.voh * Wp + g.vow) * g.Cs);
      CUresult r = enc(&tmA[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
      cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
      cuuint32_t box[2] = {64u, (cuuint32_t)p.BN};
      cuuint32_t es[2] = {1, 1};
      void* base = const_cast<__nv_bfloat16*>(wb + (size_t)src * Npad * Kpad);
      CUresult r = enc(&tmB[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MOG_ERR_CUDA, "conv_tma_kernel smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  long long grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  conv_tma_kernel<<<(unsigned)grid, TMA_THREADS, smem, st>>>(tmA[0], tmA[1], tmB[0], tmB[1], p);
  return check_launch("conv_tma_kernel");
}

}  // namespace mog
