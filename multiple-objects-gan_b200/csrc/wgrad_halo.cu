// wgrad_halo.cu -- weight gradient with *tap reuse* (tcgen05 + TMA, sm_100a).
//
//   dW[kh][kw][ci][co] = sum_pixels x[p*s + (kh, kw) - pad][ci] * dy[p][co]
//
// The generic weight-gradient kernel (conv_tc_wgrad.cu) re-gathers x once per filter tap with per-thread
// cp.async copies and re-reads dy once per 128-row slice of (tap, ci): it moves ~110 B of L2 traffic per
// tensor-core cycle and is bound by the L2->SM fabric (tensor pipe ~13 %).  Here the operands are staged
// by TMA as *pixel tiles with a halo* and every tap of a filter column is just a different start address
// into the same shared-memory tile:
//
//   * pixel tile = 8 x 8 pixels of the iteration grid (k = 64 = 4 MMA k-steps).  The TMA box
//     {64 ch, 8 w, 8 (+halo) h, 1 n} lands in the canonical MN-major SWIZZLE_128B layout: one 128-byte row
//     per pixel, 8 pixels (= one image row of the tile) per 1024-byte swizzle atom.  A shift by one image
//     row is therefore a shift by one whole atom: tap (kh, kw) reads the x box loaded for column shift kw
//     at byte offset (kh - kh_min) * 1024.  x is read once per filter COLUMN instead of once per tap.
//   * work item (CTA) = (group, 128-channel block of the M operand, N block, pixel split); a group is one
//     (view pair, column shift) with its <= 4 row taps, whose accumulators sit side by side in TMEM
//     (ntaps_h x N <= 512 fp32 columns).
//   * views: upsample(2x)+conv is handled as its 4 sub-pixel phases (dy read through a stride-2 parity
//     view, 2 x 2 pre-summed taps folded back onto the 3 x 3 filter by the reduce kernel); a stride-2 conv
//     reads x through its 4 parity views.  Views are plain strided tensor maps - no data movement.
//   * warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue; full/empty mbarrier
//     ring over the smem stages; bf16x3 = three MMAs per k-step on the hi/lo planes.
//   * fp32 partials [split][group, tap][co][ci] go to the workspace; wgrad_halo_reduce_kernel sums the
//     splits deterministically, folds phases and writes OIHW.
#include <cuda.h>
#include <cuda_bf16.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mog {
namespace tc {

constexpr int WH_THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue
constexpr int PT = 8;             // pixel tile edge: 8 x 8 = 64 pixels = 4 MMA k-steps
constexpr int WH_MAXGROUPS = 16;
constexpr int WH_MAXVIEWS = 4;

struct WHGroup {
  int xview, yview;     // tensor-map (view) indices of x and dy
  int w_off, h_org;     // origin of the x box relative to the pixel tile origin (view coordinates)
  int nth;              // row taps of this group
  int shift[4];         // per tap: image rows between the box origin and the tap's window
};

struct WHParams {
  WHGroup grp[WH_MAXGROUPS];
  int ngroups, maxtaps;
  int HH;                 // rows of the x box (PT + largest tap shift)
  int swap;               // 0: A (M = 128) = dy channels, B (N) = x channels;  1: the reverse
  int Nmma;               // N of one MMA (multiple of 16)
  int nblkA, nblkB;       // 64-channel blocks per operand tile (A: 2)
  int n_mb, n_nb;         // channel blocks of the A / B operand
  int Cx8, Cy8;           // channel pitch of the x / dy planes
  int Cin, Cout;
  int tiles_w, tiles_h;
  int tw, th, tn;         // pixel tile: tw x th pixels of tn images = 64 pixels (8 x 8 x 1, or 4 x 4 x 4 on 4 x 4 grids)
  long long total_tiles, tiles_per_split;
  int splits, passes, stages, tmem_cols;
  float* ws;              // [split][group * maxtaps + tap][co][ci]
};

struct WHMaps {
  CUtensorMap x[WH_MAXVIEWS][2];    // [view][hi / lo plane]
  CUtensorMap dy[WH_MAXVIEWS][2];
};

__device__ __forceinline__ void wh_tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void wh_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(WH_THREADS, 1) wgrad_halo_kernel(const __grid_constant__ WHMaps maps, const __grid_constant__ WHParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nplanes = p.passes == 3 ? 2 : 1;
  const int blk_plain = PT * PT * 128;        // one 64-channel block of the un-shifted operand (dy)
  const int blk_halo = p.tw * p.HH * p.tn * 128;   // ... of the shifted operand (x)
  const int blkA = p.swap ? blk_halo : blk_plain, blkB = p.swap ? blk_plain : blk_halo;
  const int regA = p.nblkA * blkA, regB = p.nblkB * blkB;
  const int plane_bytes = regA + regB;
  const int stage_bytes = nplanes * plane_bytes;
  unsigned char* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accum = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  // work decomposition: blockIdx.x -> (split, group, mb, nb)
  int bid = blockIdx.x;
  const int nb = bid % p.n_nb; bid /= p.n_nb;
  const int mb = bid % p.n_mb; bid /= p.n_mb;
  const int gi = bid % p.ngroups; bid /= p.ngroups;
  const int split = bid;
  const WHGroup& g = p.grp[gi];
  const long long tile_begin = (long long)split * p.tiles_per_split;
  long long tile_end = tile_begin + p.tiles_per_split;
  if (tile_end > p.total_tiles) tile_end = p.total_tiles;
  const int ntile = tile_begin < tile_end ? (int)(tile_end - tile_begin) : 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // channel origin of this CTA's operand blocks in the x / dy planes
  const int cA0 = mb * 128, cB0 = nb * p.Nmma;
  const int cx0 = p.swap ? cA0 : cB0, cy0 = p.swap ? cB0 : cA0;
  const int nblk_x = p.swap ? p.nblkA : p.nblkB, nblk_y = p.swap ? p.nblkB : p.nblkA;
  const int off_x = p.swap ? 0 : regA, off_y = p.swap ? regA : 0;   // region offsets inside a plane

  if (warp == 0) {
    // ===================== TMA producer ==========================================================
    if (lane == 0) {
      // blocks that start inside the tensor are loaded (partly out-of-range boxes are zero-filled by the
      // TMA unit); blocks entirely beyond the channel count are never read into a stored output
      int lx = 0, ly = 0;
      for (int b = 0; b < nblk_x; ++b) lx += (cx0 + 64 * b < p.Cx8);
      for (int b = 0; b < nblk_y; ++b) ly += (cy0 + 64 * b < p.Cy8);
      const uint32_t tx = (uint32_t)(nplanes * (lx * blk_halo + ly * blk_plain));
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < ntile; ++it) {
        const long long tile = tile_begin + it;
        const int tw_ = (int)(tile % p.tiles_w);
        const int th_ = (int)((tile / p.tiles_w) % p.tiles_h);
        const int n = (int)(tile / tiles_per_img) * p.tn;    // (images beyond N: the TMA unit zero-fills)
        mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        wh_expect_tx(&full[s], tx);
        for (int pl = 0; pl < nplanes; ++pl) {
          const uint32_t pb = st + (uint32_t)(pl * plane_bytes);
          for (int b = 0; b < lx; ++b)
            wh_tma_load_4d(pb + off_x + b * blk_halo, &maps.x[g.xview][pl], &full[s], cx0 + 64 * b, tw_ * p.tw + g.w_off,
                           th_ * p.th + g.h_org, n);
          for (int b = 0; b < ly; ++b)
            wh_tma_load_4d(pb + off_y + b * blk_plain, &maps.dy[g.yview][pl], &full[s], cy0 + 64 * b, tw_ * p.tw, th_ * p.th, n);
        }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ============================================================
    const uint32_t idesc = make_idesc_bf16(BM, p.Nmma, 1, 1);   // both operands MN-major
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < ntile; ++it) {
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      {
        // all lanes run this (uniform registers); one elected lane issues inside umma_bf16_elect
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t dhi = desc_hi_sw128(1024);
        const uint32_t a0 = desc_lo(st, (uint32_t)blkA), a1 = desc_lo(st + plane_bytes, (uint32_t)blkA);
        const uint32_t b0 = desc_lo(st + regA, (uint32_t)blkB), b1 = desc_lo(st + plane_bytes + regA, (uint32_t)blkB);
        for (int a = 0; a < g.nth; ++a) {
          const uint32_t tacc = tmem_base + (uint32_t)(a * p.Nmma);
          const uint32_t sh = (uint32_t)g.shift[a] * 64u;   // image rows -> 1024-byte atoms, in 16-byte units
          const uint32_t sa = p.swap ? sh : 0u, sb = p.swap ? 0u : sh;
          // 16 pixels per MMA = two image rows of the tile = two atoms (2048 bytes = 128 units)
#pragma unroll
          for (int k16 = 0; k16 < (PT * PT) / 16; ++k16)
            umma_bf16_elect(tacc, a0 + sa + 128 * k16, dhi, b0 + sb + 128 * k16, dhi, idesc, (it | k16) != 0);   // hi*hi
          if (p.passes == 3) {
#pragma unroll
            for (int k16 = 0; k16 < (PT * PT) / 16; ++k16)
              umma_bf16_elect(tacc, a1 + sa + 128 * k16, dhi, b0 + sb + 128 * k16, dhi, idesc, 1u);               // lo*hi
#pragma unroll
            for (int k16 = 0; k16 < (PT * PT) / 16; ++k16)
              umma_bf16_elect(tacc, a0 + sa + 128 * k16, dhi, b1 + sb + 128 * k16, dhi, idesc, 1u);               // hi*lo
          }
        }
        umma_commit_elect(&empty[s]);
        if (it == ntile - 1) umma_commit_elect(accum);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: TMEM lane = A-operand channel, columns = (tap, B-operand channel) ====
    const int q4 = warp & 3;
    const int chA = cA0 + q4 * 32 + lane;
    if (ntile > 0) {
      mbar_wait(accum, 0);
      tcgen05_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int CA = p.swap ? p.Cin : p.Cout, CB = p.swap ? p.Cout : p.Cin;
    const size_t per_tap = (size_t)p.Cin * p.Cout;
    for (int a = 0; a < g.nth; ++a) {
      float* out = p.ws + ((size_t)split * p.ngroups * p.maxtaps + (size_t)gi * p.maxtaps + a) * per_tap;
      for (int c0 = 0; c0 < p.Nmma; c0 += 16) {
        uint32_t acc[16];
        if (ntile > 0) {
          tmem_ld16(taddr + (uint32_t)(a * p.Nmma + c0), acc);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0u;
        }
        if (chA < CA) {
          const int cb = cB0 + c0;
          if (!p.swap) {
            // row co = chA, consecutive ci: contiguous per thread
            float* o = out + (size_t)chA * p.Cin + cb;
            if (cb + 15 < CB && (p.Cin & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                               __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cb + j < CB) o[j] = __uint_as_float(acc[j]);
            }
          } else {
            // lane = ci, column = co: consecutive lanes write consecutive ci
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cb + j < CB) out[(size_t)(cb + j) * p.Cin + chA] = __uint_as_float(acc[j]);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dw_oihw[co][ci][k] = sum_{split} sum_{e in src[k]} ws[split][e][co][ci]
struct WHReduceArgs {
  const float* ws;
  float* dw;
  int splits, nent, Cin, Cout, KHW;
  int nsrc[16];
  int src[16][4];
};
// Many splits (the generator's layers: up to ~300 partials per element; the patch-matrix GEMMs of the thin layers: 32 x 48
// outputs, 296 partials each).  A thread per output walking its partials 8 at a time was pure latency: a dozen blocks and
// ~110 us for the thin GEMMs, 60 - 75 us for the 96 x 96 x 9 layers (ncu r2q).  Here 16 split lanes share an output: lane l
// sums partials l, l + 16, ... (ascending, 4 loads in flight), the 16 lane sums are combined in lane order through shared
// memory -- a fixed order, deterministic.  Block = 32 outputs x 16 lanes.
constexpr int WRS_LANES = 16;
__global__ void __launch_bounds__(32 * WRS_LANES) wgrad_halo_reduce_lanes_kernel(const WHReduceArgs a) {
  __shared__ float part[WRS_LANES][33];
  const size_t per_tap = (size_t)a.Cin * a.Cout;
  const int ol = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const size_t idx = (size_t)blockIdx.x * 32 + ol;   // co * Cin + ci
  const int k = blockIdx.y;
  const size_t per_split = (size_t)a.nent * per_tap;
  const int ns = a.nsrc[k];
  const int total = a.splits * ns;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (idx < per_tap) {
    for (int i0 = sl; i0 < total; i0 += 4 * WRS_LANES) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * WRS_LANES;
        if (i < total) {
          const int z = i / ns, u = i - z * ns;
          acc[j] += __ldg(a.ws + (size_t)z * per_split + (size_t)a.src[k][u] * per_tap + idx);
        }
      }
    }
  }
  part[sl][ol] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  __syncthreads();
  if (sl == 0 && idx < per_tap) {
    float r = part[0][ol];
#pragma unroll
    for (int l = 1; l < WRS_LANES; ++l) r += part[l][ol];
    a.dw[idx * a.KHW + k] = r;
  }
}

// One block per (WR_CO output channels, 64-channel chunk of ci): the KHW x 64 sums of each co are formed with reads coalesced
// along ci (partials added in ascending order: deterministic), staged in shared memory and written as ONE contiguous
// 64 * KHW float run of the OIHW tensor per co.  A thread owns up to 4 taps x WR_CO channels and issues their loads together:
// 8 independent 4-byte loads in flight per thread.  (History: a thread per (k, co, ci) writing dw[(co*Cin + ci)*KHW + k]
// -- a 4-byte store every KHW floats, 8x write amplification -- took 895 us for the 1536 -> 3072 4x4 layer, 0.67 TB/s; one
// (co, chunk) per block with ONE load in flight per thread: 362 us, 1.5 TB/s -- latency bound, ncu r2j.)
constexpr int WR_CI = 64;
constexpr int WR_CO = 2;
__global__ void __launch_bounds__(256) wgrad_halo_reduce_kernel(const WHReduceArgs a, const int maxtotal) {
  __shared__ float tile[WR_CO][16][WR_CI + 1];
  const size_t per_tap = (size_t)a.Cin * a.Cout;
  const size_t per_split = (size_t)a.nent * per_tap;
  const int co0 = blockIdx.y * WR_CO, c0 = blockIdx.x * WR_CI;
  const int cl = threadIdx.x % WR_CI, kq = threadIdx.x / WR_CI;      // 4 tap lanes (uniform per warp)
  const int ci = c0 + cl;
  float acc[WR_CO][4];
  int ns[4], tot[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = kq + 4 * j;
    ns[j] = k < a.KHW ? a.nsrc[k] : 1;
    tot[j] = k < a.KHW ? a.splits * ns[j] : 0;
#pragma unroll
    for (int c = 0; c < WR_CO; ++c) acc[c][j] = 0.f;
  }
  if (ci < a.Cin) {
    for (int i0 = 0; i0 < maxtotal; i0 += 2) {     // two partials per step: up to 16 loads in flight per thread
      float v[2][WR_CO][4];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + h;
          const bool on = i < tot[j];
          const int k = kq + 4 * j;
          const int z = on ? i / ns[j] : 0, u = on ? i - z * ns[j] : 0;
          const float* src = a.ws + (size_t)z * per_split + (size_t)a.src[on ? k : 0][u] * per_tap + ci;
#pragma unroll
          for (int c = 0; c < WR_CO; ++c) v[h][c][j] = (on && co0 + c < a.Cout) ? __ldg(src + (size_t)(co0 + c) * a.Cin) : 0.f;
        }
#pragma unroll
      for (int h = 0; h < 2; ++h)      // ascending partial index: the summation order is fixed
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < WR_CO; ++c) acc[c][j] += v[h][c][j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = kq + 4 * j;
    if (k < a.KHW) {
#pragma unroll
      for (int c = 0; c < WR_CO; ++c) tile[c][k][cl] = acc[c][j];
    }
  }
  __syncthreads();
  const int nci = min(WR_CI, a.Cin - c0);
#pragma unroll
  for (int c = 0; c < WR_CO; ++c) {
    if (co0 + c >= a.Cout) break;
    float* o = a.dw + ((size_t)(co0 + c) * a.Cin + c0) * a.KHW;
    for (int e = threadIdx.x; e < nci * a.KHW; e += blockDim.x) o[e] = tile[c][e % a.KHW][e / a.KHW];
  }
}

}  // namespace tc

using namespace tc;

typedef CUresult (*WhEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WhEncodeFn wh_encode_fn() {
  static WhEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WhEncodeFn>(ptr);
  }
  return fn;
}

// 4-D (C, W, H, N) map of a bf16 NHWC plane [N][Hp][Wp][C8], optionally its stride-`vs` parity view (voh, vow);
// box = {64 channels, 8 pixels, box_h rows, 1 image}, SWIZZLE_128B
static int wh_encode(CUtensorMap* tm, const __nv_bfloat16* base, int N, int Hp, int Wp, int C8, int vs, int voh, int vow, int box_w,
                     int box_h, int box_n) {
  WhEncodeFn enc = wh_encode_fn();
  if (!enc) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int Hv = (Hp - voh + vs - 1) / vs, Wv = (Wp - vow + vs - 1) / vs;
  cuuint64_t dims[4] = {(cuuint64_t)C8, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)vs * C8 * 2, (cuuint64_t)vs * Wp * C8 * 2, (cuuint64_t)Hp * Wp * C8 * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t es[4] = {1, 1, 1, 1};
  void* ptr = const_cast<__nv_bfloat16*>(base + ((size_t)voh * Wp + vow) * C8);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo) failed: %d", (int)r);
  return MOG_OK;
}

static int whp8(int c) { return ceil_div(c, 8) * 8; }
static int wh_fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int wh_pmod(int a, int b) { return ((a % b) + b) % b; }

struct WHView { int vs, voh, vow; };
struct WHPlan {
  WHParams p;
  int nxv, nyv;
  WHView xv[WH_MAXVIEWS], yv[WH_MAXVIEWS];
  int Hg, Wg;                 // iteration grid (view coordinates)
  int nsrc[16], src[16][4];   // filter tap k -> workspace entries (group * maxtaps + tap)
  size_t smem;
};

// one axis: local taps (distinct offsets) and the filter taps folded into each
static int wh_axis(int mode, int a, int K, int pad, int s, int* offs, int (*mem)[2]) {
  // mode 0: plain stride-1: offset k - pad.  mode 1: sub-pixel phase a of upsample+conv: offset floor((a + k - pad)/2).
  // mode 2: parity view a of x for a stride-s conv: taps with (k - pad) mod s == a at offset floor((k - pad - a)/s).
  int n = 0;
  for (int k = 0; k < K; ++k) {
    int dd;
    if (mode == 0) dd = k - pad;
    else if (mode == 1) dd = wh_fdiv(a + k - pad, 2);
    else {
      if (wh_pmod(k - pad, s) != a) continue;
      dd = wh_fdiv(k - pad - a, s);
    }
    int j = 0;
    for (; j < n; ++j)
      if (offs[j] == dd) break;
    if (j == n) {
      if (n == 4) return -1;
      offs[n] = dd; mem[n][0] = mem[n][1] = -1; ++n;
    }
    if (mem[j][0] < 0) mem[j][0] = k;
    else if (mem[j][1] < 0) mem[j][1] = k;
    else return -1;
  }
  return n;
}

static bool wh_make_plan(const MogConvDesc& d, int Ho, int Wo, int passes, WHPlan* pl) {
  if (!wh_encode_fn()) return false;
  if (d.KH > 4 || d.KW > 4) return false;
  int mode;
  if (d.up2x) {
    if (d.stride != 1 || Ho != 2 * d.H || Wo != 2 * d.W) return false;
    mode = 1;
  } else if (d.stride == 1) {
    mode = 0;
  } else if (d.stride == 2 && (d.H % 2) == 0 && (d.W % 2) == 0) {
    mode = 2;
  } else {
    return false;
  }
  WHParams& p = pl->p;
  p = WHParams{};
  pl->Hg = mode == 1 ? d.H : Ho;
  pl->Wg = mode == 1 ? d.W : Wo;
  // 4 x 4 grids (the deep discriminator layers: img_code_s64*, jointConv): a tile packs 4 x 4 pixels of 4 images; rows of
  // different images are not a uniform stride apart, so there is no halo -- every filter tap is its own group (own box)
  const bool small = pl->Hg == 4 && pl->Wg == 4 && mode != 1;
  if (!small && (pl->Hg < PT || pl->Wg < PT)) return false;   // other tiny grids: an 8 x 8 tile would be mostly padding
  p.tw = small ? 4 : PT; p.th = small ? 4 : PT; p.tn = small ? 4 : 1;
  for (int k = 0; k < 16; ++k) pl->nsrc[k] = 0;
  const int nview = mode == 0 ? 1 : 2;   // per axis
  pl->nxv = mode == 2 ? 4 : 1;
  pl->nyv = mode == 1 ? 4 : 1;
  for (int v = 0; v < WH_MAXVIEWS; ++v) {
    pl->xv[v] = WHView{mode == 2 ? 2 : 1, mode == 2 ? v >> 1 : 0, mode == 2 ? v & 1 : 0};
    pl->yv[v] = WHView{mode == 1 ? 2 : 1, mode == 1 ? v >> 1 : 0, mode == 1 ? v & 1 : 0};
  }
  int ng = 0, maxtaps = 0, maxshift = 0;
  // first pass: groups
  for (int va = 0; va < nview; ++va)
    for (int vb = 0; vb < nview; ++vb) {
      int oh[4], ow[4], mh[4][2], mw[4][2];
      const int nth = wh_axis(mode, va, d.KH, d.pad, d.stride, oh, mh);
      const int ntw = wh_axis(mode, vb, d.KW, d.pad, d.stride, ow, mw);
      if (nth < 0 || ntw < 0) return false;
      if (nth == 0 || ntw == 0) continue;
      if (small) {
        for (int i = 0; i < nth; ++i)
          for (int j = 0; j < ntw; ++j) {
            if (ng == WH_MAXGROUPS) return false;
            WHGroup& g = p.grp[ng];
            g.xview = mode == 2 ? va * 2 + vb : 0;
            g.yview = 0;
            g.w_off = ow[j]; g.h_org = oh[i]; g.nth = 1; g.shift[0] = 0;
            for (int x = 0; x < 2; ++x)
              for (int y = 0; y < 2; ++y) {
                const int kh = mh[i][x], kw = mw[j][y];
                if (kh < 0 || kw < 0) continue;
                const int k = kh * d.KW + kw;
                if (pl->nsrc[k] == 4) return false;
                pl->src[k][pl->nsrc[k]++] = ng * 4;   // provisional stride 4, fixed below
              }
            if (maxtaps < 1) maxtaps = 1;
            ++ng;
          }
        continue;
      }
      for (int j = 0; j < ntw; ++j) {
        if (ng == WH_MAXGROUPS) return false;
        WHGroup& g = p.grp[ng];
        g.xview = mode == 2 ? va * 2 + vb : 0;
        g.yview = mode == 1 ? va * 2 + vb : 0;
        g.w_off = ow[j];
        int hmin = oh[0];
        for (int i = 1; i < nth; ++i) hmin = oh[i] < hmin ? oh[i] : hmin;
        g.h_org = hmin;
        g.nth = nth;
        for (int i = 0; i < nth; ++i) {
          g.shift[i] = oh[i] - hmin;
          if (g.shift[i] > maxshift) maxshift = g.shift[i];
          for (int x = 0; x < 2; ++x)
            for (int y = 0; y < 2; ++y) {
              const int kh = mh[i][x], kw = mw[j][y];
              if (kh < 0 || kw < 0) continue;
              const int k = kh * d.KW + kw;
              if (pl->nsrc[k] == 4) return false;
              pl->src[k][pl->nsrc[k]++] = ng * 4 + i;   // provisional stride 4, fixed below
            }
        }
        if (nth > maxtaps) maxtaps = nth;
        ++ng;
      }
    }
  if (ng == 0) return false;
  p.ngroups = ng;
  p.maxtaps = maxtaps;
  for (int k = 0; k < d.KH * d.KW; ++k)
    for (int u = 0; u < pl->nsrc[k]; ++u) pl->src[k][u] = (pl->src[k][u] / 4) * maxtaps + (pl->src[k][u] % 4);
  p.HH = p.th + maxshift;
  p.Cx8 = whp8(d.Cin); p.Cy8 = whp8(d.Cout);
  p.Cin = d.Cin; p.Cout = d.Cout;
  // operand roles: the M operand is padded to 128 channels, the N operand to 16
  auto effM = [](int c) { return (double)c / (ceil_div(c, 128) * 128); };
  auto effN = [](int c) { return (double)c / (ceil_div(c, 16) * 16); };
  p.swap = effM(d.Cin) * effN(d.Cout) > effM(d.Cout) * effN(d.Cin) ? 1 : 0;
  const int CA = p.swap ? d.Cin : d.Cout, CB = p.swap ? d.Cout : d.Cin;
  int maxN = (512 / maxtaps) / 16 * 16;
  if (maxN > 192) maxN = 192;
  const int cb16 = ceil_div(CB, 16) * 16;
  p.n_nb = ceil_div(cb16, maxN);
  p.Nmma = ceil_div(ceil_div(cb16, p.n_nb), 16) * 16;
  p.n_mb = ceil_div(CA, 128);
  p.nblkA = 2;
  p.nblkB = ceil_div(p.Nmma, 64);
  p.tiles_w = ceil_div(pl->Wg, p.tw);
  p.tiles_h = ceil_div(pl->Hg, p.th);
  p.total_tiles = (long long)ceil_div(d.N, p.tn) * p.tiles_w * p.tiles_h;
  const long long per_split_ctas = (long long)ng * p.n_mb * p.n_nb;
  long long splits = (2 * kNumSMs) / per_split_ctas;
  if (splits < 1) splits = 1;
  const long long max_splits = p.total_tiles / 4 > 0 ? p.total_tiles / 4 : 1;   // >= 4 tiles per CTA
  if (splits > max_splits) splits = max_splits;
  p.tiles_per_split = ceil_div_ll(p.total_tiles, splits);
  p.splits = (int)ceil_div_ll(p.total_tiles, p.tiles_per_split);
  p.passes = passes;
  const int nplanes = passes == 3 ? 2 : 1;
  const int blk_plain = PT * PT * 128, blk_halo = p.tw * p.HH * p.tn * 128;
  const int regA = p.nblkA * (p.swap ? blk_halo : blk_plain), regB = p.nblkB * (p.swap ? blk_plain : blk_halo);
  const int stage_bytes = nplanes * (regA + regB);
  int stages = (222 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return false;
  p.stages = stages;
  int cols = 32;
  while (cols < maxtaps * p.Nmma) cols *= 2;
  if (cols > 512) return false;
  p.tmem_cols = cols;
  pl->smem = (size_t)stages * stage_bytes + 1024 + 256;
  return true;
}

bool wgrad_halo_eligible(const MogConvDesc& d, int Ho, int Wo, int passes) {
  WHPlan pl;
  return wh_make_plan(d, Ho, Wo, passes, &pl);
}

size_t wgrad_halo_workspace_bytes(const MogConvDesc& d, int Ho, int Wo, int passes) {
  WHPlan pl;
  if (!wh_make_plan(d, Ho, Wo, passes, &pl)) return 0;
  return (size_t)pl.p.splits * pl.p.ngroups * pl.p.maxtaps * d.Cin * d.Cout * sizeof(float);
}

int launch_wgrad_halo(const MogConvDesc& d, int Ho, int Wo, const void* x_planes, const void* dy_planes, float* dw, float* ws,
                      int passes, cudaStream_t st) {
  WHPlan pl;
  if (!wh_make_plan(d, Ho, Wo, passes, &pl)) return fail(MOG_ERR_UNSUPPORTED, "wgrad (halo): shape not eligible");
  WHParams& p = pl.p;
  p.ws = ws;
  const int nplanes = passes == 3 ? 2 : 1;
  WHMaps maps;
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x_planes);
  const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(dy_planes);
  const size_t x_elems = (size_t)d.N * d.H * d.W * p.Cx8, y_elems = (size_t)d.N * Ho * Wo * p.Cy8;
  for (int v = 0; v < WH_MAXVIEWS; ++v)
    for (int plane = 0; plane < 2; ++plane) {
      const int src = plane < nplanes ? plane : 0;   // unused maps alias plane 0 (never dereferenced)
      const WHView& xv = pl.xv[v < pl.nxv ? v : 0];
      const WHView& yv = pl.yv[v < pl.nyv ? v : 0];
      int rc = wh_encode(&maps.x[v][plane], xb + src * x_elems, d.N, d.H, d.W, p.Cx8, xv.vs, xv.voh, xv.vow, p.tw, p.HH, p.tn);
      if (rc) return rc;
      rc = wh_encode(&maps.dy[v][plane], yb + src * y_elems, d.N, Ho, Wo, p.Cy8, yv.vs, yv.voh, yv.vow, p.tw, p.th, p.tn);
      if (rc) return rc;
    }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MOG_ERR_CUDA, "wgrad_halo_kernel smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const unsigned grid = (unsigned)((long long)p.splits * p.ngroups * p.n_mb * p.n_nb);
  wgrad_halo_kernel<<<grid, WH_THREADS, pl.smem, st>>>(maps, p);
  int rc = check_launch("wgrad_halo_kernel");
  if (rc) return rc;
  WHReduceArgs ra{};
  ra.ws = ws; ra.dw = dw; ra.splits = p.splits; ra.nent = p.ngroups * p.maxtaps;
  ra.Cin = d.Cin; ra.Cout = d.Cout; ra.KHW = d.KH * d.KW;
  for (int k = 0; k < d.KH * d.KW; ++k) {
    ra.nsrc[k] = pl.nsrc[k];
    for (int u = 0; u < 4; ++u) ra.src[k][u] = pl.src[k][u];
  }
  int max_src = 1;
  for (int k = 0; k < d.KH * d.KW; ++k) max_src = pl.nsrc[k] > max_src ? pl.nsrc[k] : max_src;
  if (p.splits * max_src <= 8) {
    // few partials, large tensors (the deep discriminator layers): the pass is a layout change -> coalesced OIHW runs
    wgrad_halo_reduce_kernel<<<dim3((unsigned)ceil_div(d.Cin, WR_CI), (unsigned)ceil_div(d.Cout, WR_CO)), 256, 0, st>>>(ra, p.splits * max_src);
  } else {
    const size_t total = (size_t)d.Cin * d.Cout;
    wgrad_halo_reduce_lanes_kernel<<<dim3((unsigned)ceil_div_ll((long long)total, 32), (unsigned)(d.KH * d.KW)), 32 * WRS_LANES, 0, st>>>(ra);
  }
  return check_launch("wgrad_halo_reduce_kernel");
}

}  // namespace mog
