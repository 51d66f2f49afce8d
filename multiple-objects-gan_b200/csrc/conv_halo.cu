// conv_halo.cu -- persistent, warp-specialised tcgen05 convolution with halo-tile TMA staging (sm_100a).
//
// Serves the unit-stride gather-GEMM problems (stride-1 forward convs, sub-pixel phases of upsample+conv,
// parity views of strided convs, every stride phase of a data gradient) whose operand exists as pre-split
// bf16 planes [N][H][W][C8].  Measured on B200 the previous per-tap TMA kernel was bound by the bytes each
// SM can ingest from L2 (~37 B/clk/SM): it staged 128 px x 64 ch of A and BN x 64 of B per k-chunk and used
// each byte once.  This kernel raises the reuse of both operands:
//
//   * A, halo tiles: an M sub-tile is 8 (w) x 16 (h) output pixels.  For filter COLUMN kw and a 32-channel
//     chunk ONE TMA box {32 ch, 8 w, 16 + (KH-1) h} lands in the K-major SWIZZLE_64B layout (one 64-byte row
//     per pixel, one 512-byte swizzle atom per 8-pixel image row), so the row taps kh are the same box read
//     at start address + (kh - kh_min) * 512 bytes: A is staged once per filter column, not once per tap.
//   * B, two sub-tiles per CTA: every weight chunk {32 k, BN} staged in shared memory feeds the MMAs of two
//     M sub-tiles (two TMEM accumulators), halving the weight traffic per output pixel.
//   * 32-channel chunks: K per tap is the channel count rounded up to 32 (96-channel layers waste nothing;
//     the 64-channel chunks of the previous kernel padded them to 128).
//   * several *views* in one K loop: the parity views of a strided conv (or of dy in the data gradient of
//     upsample+conv) are separate tensor maps over the same planes; their taps accumulate in TMEM instead of
//     through read-modify-write passes over the fp32 output.
//   * separate full/empty mbarrier rings for A boxes and B chunks, each fed by its own producer thread (warp 0:
//     A; warp 11: B; explicit cp.async.bulk.prefetch of the next tile pair into L2 was measured and made both
//     this kernel and the weight-gradient kernel SLOWER - the prefetches compete with the loads); warps 1-2 = MMA
//     issuers, one per sub-tile accumulator (a single issuing thread needs ~40-50 cycles per tcgen05.mma and
//     would bound the N <= 128 shapes; warp-convergent loops, one elected lane issues); warps 3-10 = epilogue,
//     four per sub-tile; accumulators double-buffered in TMEM when 4 x BN <= 512 columns; CTAs persistent
//     (grid = #SMs, static schedule).
//   * epilogue through TMA stores: scattered 16-byte st.global at a pixel pitch of Cd*4 bytes kept the LSU busy
//     ~2 clk per half-filled sector (25k clk per tile pair at BN=192, more than the MMAs took).  Each epilogue
//     group now converts TMEM -> registers (bias / activation) -> a swizzled [128 px][16 ch] fp32 slab in
//     shared memory (double-buffered) and one thread issues cp.async.bulk.tensor stores of the box
//     {16 ch, 8 w, 16 h} into the (strided, for sub-pixel phases) NHWC destination; ragged edges, channel
//     tails and padding sub-tiles are clipped by the TMA unit.
//   * MOG_PREC_BF16X3: three MMAs per k-step (hi*hi, lo*hi, hi*lo) on the hi/lo planes.
//   * small-grid form (pixel grids up to 8 x 8: the deep discriminator layers, the 8 x 8 stage of the image encoder): a
//     sub-tile packs several images (8 x 8 px of 2 images, or 4 x 4 px of 8 images), every filter tap is its own box (no
//     halo: rows of different images are not a uniform stride apart), the N tile grows to 256 columns (single-buffered
//     TMEM: two 128-row sub-tiles share each 256-column weight chunk), and the few output tiles are spread over the SMs
//     by split-K: a work item is (sub-tile pair, N tile, K range); raw fp32 partials leave through a 5-D TMA store into
//     [split][N][H][W][C] and splitk_reduce sums them in a fixed order (bias / activation applied there).
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mog {
namespace tc {

constexpr int HALO_THREADS = 384;          // A producer + 2 MMA warps + 8 epilogue warps + B producer
constexpr int HT_W = 8, HT_H = 16;          // sub-tile: 8 x 16 = 128 output pixels
constexpr int HCH = 32;                     // channels per chunk (64-byte rows, SWIZZLE_64B)
constexpr int HALO_MAXGROUPS = 16;
constexpr int HALO_MAXPROBS = 4;
constexpr int HALO_MAXRING = 8;

struct HaloGroup {
  int prob;          // which problem (view + weight block)
  int w_off, h_org;  // box origin relative to the sub-tile origin (view coordinates)
  int nth;
  int shift[4];      // image rows between the box origin and the window of each row tap
  int slot[4];       // weight tap slot of each row tap (k0 = slot * pitch + chunk * 32)
};

struct HaloParams {
  HaloGroup grp[HALO_MAXGROUPS];
  int ngroups, nchunk, pitch;
  int HH;                          // rows of an A box
  int subA;                        // bytes of one A sub-tile box (1024-aligned)
  int tw, th, tn;                  // sub-tile geometry: tw x th pixels of tn images = 128 GEMM rows
  int tiles_w, tiles_h, tiles_n;
  int ksplit, kper, niter;         // split-K: niter = ngroups * nchunk (group, chunk) iterations, kper per split
  int split_store;                 // epilogue stores raw fp32 partials [ksplit][N][Hr][Wr][Cd] (5-D map), reduced by splitk_reduce
  long long nsub, npairs, total_tiles;
  int n_ntiles, BN, nbuf;
  int passes, stagesA, stagesB, tmem_cols;
  float* dst;
  const float* bias;
  int act, accum_dst;
  int tma_store;                   // epilogue through shared memory + TMA stores (Cd % 4 == 0, no accumulate)
  int single;                      // 1: a work item is ONE sub-tile (the second MMA warp idles): small problems that would leave
                                   // more than half of the SMs without a tile pair
  int N, Hr, Wr, Cd, Hd, Wd, dsh, doh, dsw, dow;
};

struct HaloMaps {
  CUtensorMap a[HALO_MAXPROBS][2];   // [problem / view][hi, lo]
  CUtensorMap b[HALO_MAXPROBS][2];
  CUtensorMap d;                     // fp32 destination pixel grid of this problem (a strided view for sub-pixel phases), or the
                                     // 5-D split-K partial buffer
};

__device__ __forceinline__ void halo_tma_4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void halo_tma_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void halo_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void halo_tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void halo_tma_store_5d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tm), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int HALO_OUT_SLAB = 128 * 16 * 4;   // [128 px][16 ch] fp32

// one 16-channel group of one epilogue thread (= one pixel row of the sub-tile): bias / activation in registers,
// then the 64-byte row goes to the SWIZZLE_64B slab (16-byte chunk index ^ bits 7-8 of the byte address)
__device__ __forceinline__ void halo_epi_slab16(const uint32_t (&acc)[16], const float* bias_c, int nvalid, int act, unsigned char* slab, int r) {
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(acc[j]);
  if (bias_c) {
    if (nvalid >= 16) {
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
  const int sw = (r >> 1) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(slab + r * 64 + ((c ^ sw) << 4)) = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
}

// hi word of a K-major SWIZZLE_64B descriptor: SBO = 512 bytes (8 rows of 64 bytes), version 1, layout 4
__device__ __forceinline__ uint32_t desc_hi_sw64() { return (512u >> 4) | (1u << 14) | (4u << 29); }

// sub-tile index -> first image / pixel of the sub-tile (tn images x th x tw pixels)
__device__ __forceinline__ void halo_decode(const HaloParams& p, long long sub, int* n, int* h0, int* w0) {
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int tw_ = (int)(sub % p.tiles_w);
  const int th_ = (int)((sub / p.tiles_w) % p.tiles_h);
  *n = (int)(sub / tiles_per_img) * p.tn;   // >= N for the padding sub-tile of an odd count: TMA zero-fills, nothing is stored
  *h0 = th_ * p.th;
  *w0 = tw_ * p.tw;
}
// sub-tile s (0 / 1) of a work item's pair; in single mode the second sub-tile does not exist (decodes beyond N)
__device__ __forceinline__ long long halo_sub(const HaloParams& p, long long pair, int s) {
  if (!p.single) return pair * 2 + s;
  return s == 0 ? pair : (long long)p.tiles_n * p.tiles_w * p.tiles_h;   // first index past the last sub-tile: n >= N
}
// work item -> (sub-tile pair, N tile, K split)
__device__ __forceinline__ void halo_item(const HaloParams& p, long long tile, long long* pair, int* ntile, int* ks) {
  *ks = (int)(tile % p.ksplit);
  const long long q = tile / p.ksplit;
  *ntile = (int)(q % p.n_ntiles);
  *pair = q / p.n_ntiles;
}

__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nplanes = p.passes == 3 ? 2 : 1;
  const int planeA = 2 * p.subA;                 // two sub-tiles
  const int slotA = nplanes * planeA;
  const int planeB = p.BN * (HCH * 2);
  const int slotB = nplanes * planeB;
  unsigned char* ringA = smem;
  unsigned char* ringB = smem + (size_t)p.stagesA * slotA;
  unsigned char* outbuf = ringB + (size_t)p.stagesB * slotB;   // [2 sub-tiles][2 buffers] slabs (TMA-store epilogue only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(outbuf + (p.tma_store ? 4 * HALO_OUT_SLAB : 0));
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + HALO_MAXRING;
  uint64_t* fullB = emptyA + HALO_MAXRING;
  uint64_t* emptyB = fullB + HALO_MAXRING;
  uint64_t* tfull = emptyB + HALO_MAXRING;   // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stagesA; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 2); }   // empty: one commit per MMA warp
    for (int s = 0; s < p.stagesB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 2); mbar_init(&tempty[s], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer of A (activation halo boxes) =============================
    if (lane == 0) {
      int sa = 0;
      uint32_t pha = 0;
      const uint32_t bytesA = (uint32_t)(nplanes * 2 * (HCH * 2 * p.tw * p.HH * p.tn));
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        long long pair;
        int ntile_, ks;
        halo_item(p, tile, &pair, &ntile_, &ks);
        int n[2], h0[2], w0[2];
        halo_decode(p, halo_sub(p, pair, 0), &n[0], &h0[0], &w0[0]);
        halo_decode(p, halo_sub(p, pair, 1), &n[1], &h0[1], &w0[1]);
        const int it_end = min((ks + 1) * p.kper, p.niter);
        for (int it = ks * p.kper; it < it_end; ++it) {
          const int gi = it / p.nchunk, c = it - gi * p.nchunk;
          const HaloGroup& g = p.grp[gi];
          {
            mbar_wait(&emptyA[sa], pha ^ 1u);
            const uint32_t stA = smem_u32(ringA + (size_t)sa * slotA);
            halo_expect_tx(&fullA[sa], bytesA);
            for (int pl = 0; pl < nplanes; ++pl)
              for (int s = 0; s < 2; ++s)
                halo_tma_4d(stA + pl * planeA + s * p.subA, &maps.a[g.prob][pl], &fullA[sa], c * HCH, w0[s] + g.w_off, h0[s] + g.h_org, n[s]);
            if (++sa == p.stagesA) { sa = 0; pha ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 11) {
    // ===================== TMA producer of B (weight chunks): its own thread, so that the deep B ring keeps
    // running ahead while the A producer waits for a free A slot ===================================
    if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      const uint32_t bytesB = (uint32_t)slotB;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        long long pair;
        int ntile, ks;
        halo_item(p, tile, &pair, &ntile, &ks);
        const int it_end = min((ks + 1) * p.kper, p.niter);
        for (int it = ks * p.kper; it < it_end; ++it) {
          const int gi = it / p.nchunk, c = it - gi * p.nchunk;
          const HaloGroup& g = p.grp[gi];
          {
            for (int a = 0; a < g.nth; ++a) {
              mbar_wait(&emptyB[sb], phb ^ 1u);
              const uint32_t stB = smem_u32(ringB + (size_t)sb * slotB);
              halo_expect_tx(&fullB[sb], bytesB);
              for (int pl = 0; pl < nplanes; ++pl)
                halo_tma_2d(stB + pl * planeB, &maps.b[g.prob][pl], &fullB[sb], g.slot[a] * p.pitch + c * HCH, ntile * p.BN);
              if (++sb == p.stagesB) { sb = 0; phb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp <= 2) {
    // ===================== MMA issuers: warp 1 -> sub-tile 0, warp 2 -> sub-tile 1 ================
    // (all lanes run the loop so that addresses stay in uniform registers; one elected lane issues)
    const int sub = warp - 1;
    const bool active = !(p.single && sub == 1);     // single mode: this warp only keeps the barrier protocol going
    const uint32_t idesc = make_idesc_bf16(BM, p.BN);
    const uint32_t dhi = desc_hi_sw64();
    int sa = 0, sb = 0, as = 0;
    uint32_t pha = 0, phb = 0, aph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aph ^ 1u);       // epilogue has drained this accumulator pair
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)((as * 2 + sub) * p.BN);
      uint32_t first = 0u;                     // becomes 1 after the first k-step of the tile
      const int ks = (int)(tile % p.ksplit);
      const int it_end = min((ks + 1) * p.kper, p.niter);
      for (int it = ks * p.kper; it < it_end; ++it) {
        const HaloGroup& g = p.grp[it / p.nchunk];
        {
          mbar_wait(&fullA[sa], pha);
          const uint32_t stA = smem_u32(ringA + (size_t)sa * slotA) + (uint32_t)(sub * p.subA);
          for (int a = 0; a < g.nth; ++a) {
            mbar_wait(&fullB[sb], phb);
            tcgen05_fence_after();
            const uint32_t stB = smem_u32(ringB + (size_t)sb * slotB);
            const uint32_t sh = (uint32_t)g.shift[a] * (uint32_t)(p.tw * HCH * 2);
            const uint32_t ah = desc_lo(stA + sh, 16), al = desc_lo(stA + planeA + sh, 16);
            const uint32_t bh = desc_lo(stB, 16), bl = desc_lo(stB + planeB, 16);
            if (active) {
#pragma unroll
            for (int k16 = 0; k16 < HCH / 16; ++k16) umma_bf16_elect(tacc, ah + 2 * k16, dhi, bh + 2 * k16, dhi, idesc, first | (uint32_t)k16);   // hi*hi
            }
            if (active && p.passes == 3) {
#pragma unroll
              for (int k16 = 0; k16 < HCH / 16; ++k16) umma_bf16_elect(tacc, al + 2 * k16, dhi, bh + 2 * k16, dhi, idesc, 1u);   // lo*hi
#pragma unroll
              for (int k16 = 0; k16 < HCH / 16; ++k16) umma_bf16_elect(tacc, ah + 2 * k16, dhi, bl + 2 * k16, dhi, idesc, 1u);   // hi*lo
            }
            first = 1u;
            umma_commit_elect(&emptyB[sb]);
            if (++sb == p.stagesB) { sb = 0; phb ^= 1u; }
          }
          umma_commit_elect(&emptyA[sa]);
          if (++sa == p.stagesA) { sa = 0; pha ^= 1u; }
        }
      }
      umma_commit_elect(&tfull[as]);
      __syncwarp();
      if (++as == p.nbuf) { as = 0; aph ^= 1u; }
    }
  } else if (warp <= 10) {
    // ===================== epilogue: warps 3-6 drain sub-tile 0, warps 7-10 sub-tile 1 ==============
    const int s = (warp - 3) >> 2;      // sub-tile of this warp
    const int q4 = warp & 3;            // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int r = q4 * 32 + lane;       // sub-tile row = (image nl, pixel hl, wl) of the sub-tile, wl fastest
    const int wl = r % p.tw, hl = (r / p.tw) % p.th, nl = r / (p.tw * p.th);
    const bool vec = (p.Cd & 3) == 0;
    int as = 0;
    uint32_t aph = 0;
    int ob = 0;                                   // output slab buffer (alternates across groups and tiles)
    const bool leader = ((warp - 3) & 3) == 0 && lane == 0;   // issues this group's TMA stores
    unsigned char* slabs = outbuf + (size_t)s * 2 * HALO_OUT_SLAB;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long pair;
      int ntile, ks;
      halo_item(p, tile, &pair, &ntile, &ks);
      const int n0c = ntile * p.BN;
      mbar_wait(&tfull[as], aph);
      tcgen05_fence_after();
      if (p.tma_store) {
        int n, h0, w0;
        halo_decode(p, halo_sub(p, pair, s), &n, &h0, &w0);
        if (n < p.N) {   // (padding sub-tile of an odd count / idle half of single mode: nothing to store)
          const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((as * 2 + s) * p.BN);
          const int ngrp = min(p.BN, p.Cd - n0c + 15) / 16;   // groups that hold at least one real channel
          uint32_t acc[2][16];
          tmem_ld16_async(taddr, acc[0]);
#pragma unroll 1
          for (int gq = 0; gq < ngrp; gq += 2) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int gg = gq + u;
              if (gg < ngrp) {
                tmem_ld_wait(acc[u]);
                if (gg + 1 < ngrp) tmem_ld16_async(taddr + (uint32_t)((gg + 1) * 16), acc[u ^ 1]);
                if (leader) bulk_wait_read<1>();   // the store that read slab `ob` two groups ago is done with it
                named_bar_sync(1 + s, 128);
                unsigned char* slab = slabs + ob * HALO_OUT_SLAB;
                halo_epi_slab16(acc[u], p.bias ? p.bias + n0c + gg * 16 : nullptr, p.Cd - n0c - gg * 16, p.act, slab, r);
                fence_proxy_async();               // generic-proxy smem writes -> visible to the TMA unit
                named_bar_sync(1 + s, 128);
                if (leader) {
                  if (p.split_store) halo_tma_store_5d(&maps.d, smem_u32(slab), n0c + gg * 16, w0, h0, n, ks);
                  else halo_tma_store_4d(&maps.d, smem_u32(slab), n0c + gg * 16, w0, h0, n);
                  bulk_commit();
                }
                ob ^= 1;
              }
            }
          }
        }
      } else {
        int n, h0, w0;
        halo_decode(p, halo_sub(p, pair, s), &n, &h0, &w0);
        const int rh = h0 + hl, rw = w0 + wl, rn = n + nl;
        const bool ok = rn < p.N && rh < p.Hr && rw < p.Wr;
        float* dptr = nullptr;
        if (ok) {
          const size_t pix = ((size_t)rn * p.Hd + (rh * p.dsh + p.doh)) * p.Wd + (rw * p.dsw + p.dow);
          dptr = p.dst + pix * p.Cd + n0c;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((as * 2 + s) * p.BN);
        // software-pipelined TMEM reads: the load of group g+1 is in flight while group g is stored
        const int ngrp = p.BN / 16;
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
      }
      // every TMEM load of this warp was waited for: release the accumulator pair
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
      if (++as == p.nbuf) { as = 0; aph ^= 1u; }
    }
    if (p.tma_store && leader) bulk_wait_read<0>();   // shared memory must outlive the reads of the last stores
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

typedef CUresult (*HaloEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static HaloEncodeFn halo_encode_fn() {
  static HaloEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<HaloEncodeFn>(ptr);
  }
  return fn;
}

// K pitch of one tap slot in the weights packed for this kernel
int halo_tap_pitch(int Cs) { return ceil_div(Cs, HCH) * HCH; }

// ---- sub-tile geometry -------------------------------------------------------------------------------------
// Large pixel grids use 8 x 16 pixel sub-tiles of one image with halo boxes (row taps = shifted windows of one box per
// filter column).  Small grids (the deep discriminator layers at 8 x 8 / 4 x 4, the 8 x 8 stage of the image encoder)
// pack several images into one 128-row sub-tile and load one box per filter tap (no halo: the rows of different images
// are not a uniform stride apart); their few output tiles are spread over the SMs by split-K.
struct HaloGeom { int tw, th, tn, halo; };
static HaloGeom halo_geom(const IGemmParams& g) {
  if (g.Hr >= 12 && g.Wr >= HT_W) return {HT_W, HT_H, 1, 1};
  if (g.Hr > 4 || g.Wr > 4) return {8, 8, 2, 0};
  return {4, 4, 8, 0};
}

// shape-only test (the weight packing depends on it)
bool halo_shape_eligible(const IGemmParams& g) {
  if (g.rs != 1 || g.up2x) return false;
  if (g.vstep > 1 && ((g.Hp % g.vstep) || (g.Wp % g.vstep))) return false;
  if (g.nth < 1 || g.ntw < 1) return false;
  if (g.Cd < 1) return false;
  if (halo_encode_fn() == nullptr) return false;
  const HaloGeom ge = halo_geom(g);
  if (ge.halo) {
    if (g.nth > 4 || g.ntw > HALO_MAXGROUPS) return false;
    if (g.M < 32LL * 128) return false;             // small problems on big grids: the split-K gather kernel fills the machine better
    int lo = g.off_h[0], hi = g.off_h[0];
    for (int i = 1; i < g.nth; ++i) { lo = g.off_h[i] < lo ? g.off_h[i] : lo; hi = g.off_h[i] > hi ? g.off_h[i] : hi; }
    return hi - lo <= 8;
  }
  // small grids: one group per tap; split-K partials go through TMA stores (channel count multiple of 4, no accumulate chain)
  if (g.nth * g.ntw > HALO_MAXGROUPS) return false;
  if ((g.Cd & 3) || g.Cd < 64) return false;        // (narrow outputs: the gather kernel's N tile wastes less)
  if (g.Hr > 8 || g.Wr > 8) return false;           // 9..11-row grids stay on the gather kernel
  if (g.M < 256 || g.Cs < 64) return false;
  return true;
}

// plan of one (possibly merged) launch: everything but the pointers
static int halo_plan(const IGemmParams* gs, int n, int passes, HaloParams* out) {
  if (n < 1 || n > HALO_MAXPROBS) return fail(MOG_ERR_BAD_ARG, "conv (halo): %d problems", n);
  const IGemmParams& g0 = gs[0];
  const int nplanes = passes == 3 ? 2 : 1;
  const HaloGeom ge = halo_geom(g0);
  HaloParams p{};
  int maxshift = 0, ng = 0;
  for (int i = 0; i < n; ++i) {
    const IGemmParams& g = gs[i];
    if (g.Cs % 8) return fail(MOG_ERR_BAD_ARG, "conv (halo): needs pre-split planes with a channel pitch multiple of 8");
    if (g.Cs != g0.Cs || g.Hr != g0.Hr || g.Wr != g0.Wr || g.N != g0.N || g.Cd != g0.Cd)
      return fail(MOG_ERR_BAD_ARG, "conv (halo): merged problems must share the output grid");
    if (ge.halo) {
      int hmin = g.off_h[0];
      for (int a = 1; a < g.nth; ++a) hmin = g.off_h[a] < hmin ? g.off_h[a] : hmin;
      for (int b = 0; b < g.ntw; ++b) {
        if (ng == HALO_MAXGROUPS) return fail(MOG_ERR_UNSUPPORTED, "conv (halo): too many filter columns");
        HaloGroup& q = p.grp[ng++];
        q.prob = i; q.w_off = g.off_w[b]; q.h_org = hmin; q.nth = g.nth;
        for (int a = 0; a < g.nth; ++a) {
          q.shift[a] = g.off_h[a] - hmin;
          q.slot[a] = a * g.ntw + b;
          if (q.shift[a] > maxshift) maxshift = q.shift[a];
        }
      }
    } else {
      for (int a = 0; a < g.nth; ++a)
        for (int b = 0; b < g.ntw; ++b) {
          if (ng == HALO_MAXGROUPS) return fail(MOG_ERR_UNSUPPORTED, "conv (halo): too many filter taps for the small-grid form");
          HaloGroup& q = p.grp[ng++];
          q.prob = i; q.w_off = g.off_w[b]; q.h_org = g.off_h[a]; q.nth = 1;
          q.shift[0] = 0; q.slot[0] = a * g.ntw + b;
        }
    }
  }
  p.ngroups = ng;
  p.nchunk = ceil_div(g0.Cs, HCH);
  p.pitch = p.nchunk * HCH;
  p.niter = p.ngroups * p.nchunk;
  p.tw = ge.tw; p.th = ge.th; p.tn = ge.tn;
  p.HH = ge.th + maxshift;
  p.subA = ceil_div(ge.tw * p.HH * ge.tn * HCH * 2, 1024) * 1024;
  p.tiles_w = ceil_div(g0.Wr, ge.tw);
  p.tiles_h = ceil_div(g0.Hr, ge.th);
  p.tiles_n = ceil_div(g0.N, ge.tn);
  p.nsub = (long long)p.tiles_n * p.tiles_w * p.tiles_h;
  p.npairs = (p.nsub + 1) / 2;
  p.single = 0;
  p.passes = passes;
  p.tma_store = ((g0.Cd & 3) == 0 && !g0.accum_dst) ? 1 : 0;
  // N tile: <= 128 columns on the big grids so that two sub-tiles x two accumulator buffers fit the 512 TMEM columns
  // (the epilogue overlaps the MMAs of the next tile; N = 192 is issued as 2 x 96); the small-grid form takes up to 256
  // columns (single-buffered: few tiles per CTA, and the weight chunk is then shared by 2 x 128 rows x 256 columns)
  {
    const int cpad = ceil_div(g0.Cd, 16) * 16;
    static int halo_bn_max = 0;
    if (!halo_bn_max) {
      const char* e = getenv("MOG_HALO_BN_MAX");     // tuning knob: widest N tile of the halo form (default 128: double-buffered TMEM)
      halo_bn_max = e ? atoi(e) : 128;
      if (halo_bn_max < 16 || halo_bn_max > 256) halo_bn_max = 128;
    }
    const int tiles = ceil_div(cpad, ge.halo ? halo_bn_max : 256);
    p.BN = ceil_div(ceil_div(cpad, tiles), 16) * 16;
  }
  p.n_ntiles = ceil_div(g0.Cd, p.BN);
  p.nbuf = 4 * p.BN <= 512 ? 2 : 1;
  // split-K (small-grid form only): aim at one work item per SM, at least 4 (group, chunk) iterations per split
  p.ksplit = 1;
  if (!ge.halo && p.tma_store) {
    const long long items = p.npairs * p.n_ntiles;
    if (items < kNumSMs) {
      static int per_sm = 0;
      if (!per_sm) {
        // tuning knob: work items per SM the split aims at.  Measured on the full step: 1 -> 72.8 ms, 2 -> 77.8 ms (the
        // fp32 partials of a second wave cost more than the idle SMs of a single one)
        const char* e = getenv("MOG_HALO_SPLIT_WAVES");
        per_sm = e ? atoi(e) : 1;
        if (per_sm < 1 || per_sm > 4) per_sm = 1;
      }
      long long want = ((long long)per_sm * kNumSMs) / items, maxs = p.niter / 4;
      long long ks = want < maxs ? want : maxs;
      if (ks > 32) ks = 32;
      if (ks > 1) p.ksplit = (int)ks;
    }
  }
  if (ge.halo && p.ksplit == 1 && p.nsub > 1) {
    // small problems on the halo form (the 17 x 17 / 35 x 35 stages of the image encoder at B = 32: 37 tile pairs x 2 N
    // tiles for 148 SMs): one sub-tile per work item doubles the CTAs; each finishes in about half the time (tensor work per
    // CTA halves, the weight chunk is no longer shared)
    static int knob = -1;
    if (knob < 0) {
      const char* e = getenv("MOG_HALO_SINGLE");     // tuning knob: 0 disables
      knob = e ? atoi(e) : 1;
    }
    // (a cost model that also switched mid-size problems -- rounds x 0.55 per single item -- was measured SLOWER: 63.2 vs
    // 61.6 ms per step; singles pay unshared weight chunks.  Only problems whose pairs leave half of the SMs idle switch.)
    if (knob && p.npairs * p.n_ntiles * 2 <= kNumSMs + kNumSMs / 4) {
      p.single = 1;
      p.npairs = p.nsub;
      // still under half a wave (e.g. 73 sub-tiles x 1 N tile): split the N tile too (A is re-read per N tile: tiny here)
      static int knob2 = -1;
      if (knob2 < 0) {
        const char* e = getenv("MOG_HALO_SINGLE_NSPLIT");
        knob2 = e ? atoi(e) : 1;
      }
      while (knob2 && p.npairs * p.n_ntiles * 2 <= kNumSMs && p.BN >= 64) {
        p.BN = ceil_div(p.BN / 2, 16) * 16;
        p.n_ntiles = ceil_div(g0.Cd, p.BN);
      }
      p.nbuf = 4 * p.BN <= 512 ? 2 : 1;
    }
  }
  p.kper = ceil_div(p.niter, p.ksplit);
  p.ksplit = ceil_div(p.niter, p.kper);
  p.split_store = p.ksplit > 1 ? 1 : 0;
  p.total_tiles = p.npairs * p.n_ntiles * p.ksplit;
  int cols = 32;
  while (cols < 2 * p.nbuf * p.BN) cols *= 2;
  p.tmem_cols = cols;
  const int slotA = nplanes * 2 * p.subA, slotB = nplanes * p.BN * HCH * 2;
  const int budget = 224 * 1024 - 1024 - (p.tma_store ? 4 * HALO_OUT_SLAB : 0);
  int stagesA = ge.halo ? 3 : 2;     // (small-grid form: the weight ring is the one that must run deep)
  {
    static int knob = -1;
    if (knob < 0) {
      const char* e = getenv("MOG_HALO_STAGES_A");   // tuning knob: depth of the activation ring of the halo form
      knob = e ? atoi(e) : 0;
    }
    if (ge.halo && knob >= 2 && knob <= 6) stagesA = knob;
  }
  int stagesB = (budget - stagesA * slotA) / slotB;
  if (stagesB < 2) { stagesA = 2; stagesB = (budget - stagesA * slotA) / slotB; }
  if (stagesB < 2) return fail(MOG_ERR_UNSUPPORTED, "conv (halo): BN=%d does not fit the shared-memory rings", p.BN);
  if (stagesB > HALO_MAXRING) stagesB = HALO_MAXRING;
  p.stagesA = stagesA; p.stagesB = stagesB;
  p.N = g0.N; p.Hr = g0.Hr; p.Wr = g0.Wr; p.Cd = g0.Cd; p.Hd = g0.Hd; p.Wd = g0.Wd;
  p.dsh = g0.dsh; p.doh = g0.doh; p.dsw = g0.dsw; p.dow = g0.dow;
  *out = p;
  return MOG_OK;
}

// split-K workspace of a launch (0 when it does not split); gs[i].Cs must be the channel pitch of the planes
size_t halo_workspace_bytes(const IGemmParams* gs, int n, int passes) {
  HaloParams p;
  if (halo_plan(gs, n, passes, &p) != MOG_OK || p.ksplit <= 1) return 0;
  return (size_t)p.ksplit * gs[0].M * gs[0].Cd * sizeof(float);
}

// gs[0..n): problems that write the same destination pixels (n > 1: parity views accumulated in TMEM);
// packed[i]: weight block of problem i (layout of tc_pack_pitch with pitch = halo_tap_pitch(Cs))
int launch_igemm_halo(const IGemmParams* gs, int n, const void* const* packed, int passes, void* workspace, size_t ws_bytes,
                      cudaStream_t st) {
  HaloEncodeFn enc = halo_encode_fn();
  if (!enc) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  HaloParams p;
  int rc = halo_plan(gs, n, passes, &p);
  if (rc) return rc;
  const IGemmParams& g0 = gs[0];
  for (int i = 0; i < n; ++i)
    if (!gs[i].src_planes) return fail(MOG_ERR_BAD_ARG, "conv (halo): needs pre-split planes");
  const int nplanes = passes == 3 ? 2 : 1;
  const int stagesA = p.stagesA, stagesB = p.stagesB;
  const int slotA = nplanes * 2 * p.subA, slotB = nplanes * p.BN * HCH * 2;
  float* partial = nullptr;
  if (p.split_store) {
    const size_t need = (size_t)p.ksplit * g0.M * g0.Cd * sizeof(float);
    if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "conv (halo split-K): workspace %zu < %zu", ws_bytes, need);
    partial = static_cast<float*>(workspace);
    p.dst = partial; p.bias = nullptr; p.act = MOG_ACT_NONE; p.accum_dst = 0;   // raw partials; bias / activation in the reduce
  } else {
    p.dst = g0.dst; p.bias = g0.bias; p.act = g0.act; p.accum_dst = g0.accum_dst;
  }

  HaloMaps maps;
  const int Npad = ceil_div(g0.Cd, tc_bn_for(g0.Cd)) * tc_bn_for(g0.Cd);   // rows of the packed weight planes (tc_pack_pitch)
  for (int i = 0; i < HALO_MAXPROBS; ++i) {
    const IGemmParams& g = gs[i < n ? i : 0];
    const __nv_bfloat16* xa = static_cast<const __nv_bfloat16*>(g.src_planes);
    const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(packed[i < n ? i : 0]);
    const int Kpad = ceil_div(g.nth * g.ntw * p.pitch, 64) * 64;
    for (int pl = 0; pl < 2; ++pl) {
      const int src = pl < nplanes ? pl : 0;   // unused maps alias plane 0 (never dereferenced)
      {
        // (strided) view of the physical [N][Hp][Wp][Cs] plane: logical (h, w) = physical (h*vs + voh, w*vs + vow)
        const int vs = g.vstep > 0 ? g.vstep : 1, Hp = g.vstep > 0 ? g.Hp : g.Hs, Wp = g.vstep > 0 ? g.Wp : g.Ws;
        cuuint64_t dims[4] = {(cuuint64_t)g.Cs, (cuuint64_t)g.Ws, (cuuint64_t)g.Hs, (cuuint64_t)g.N};
        cuuint64_t strides[3] = {(cuuint64_t)vs * g.Cs * 2, (cuuint64_t)vs * Wp * g.Cs * 2, (cuuint64_t)Hp * Wp * g.Cs * 2};
        cuuint32_t box[4] = {(cuuint32_t)HCH, (cuuint32_t)p.tw, (cuuint32_t)p.HH, (cuuint32_t)p.tn};
        cuuint32_t es[4] = {1, 1, 1, 1};
        void* base = const_cast<__nv_bfloat16*>(xa + (size_t)src * g.src_plane_elems + ((size_t)g.voh * Wp + g.vow) * g.Cs);
        CUresult r = enc(&maps.a[i][pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(halo A) failed: %d", (int)r);
      }
      {
        cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
        cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
        cuuint32_t box[2] = {(cuuint32_t)HCH, (cuuint32_t)p.BN};
        cuuint32_t es[2] = {1, 1};
        void* base = const_cast<__nv_bfloat16*>(wb + (size_t)src * Npad * Kpad);
        CUresult r = enc(&maps.b[i][pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(halo B) failed: %d", (int)r);
      }
    }
  }
  if (p.split_store) {
    // split-K partials [ksplit][N][Hr][Wr][Cd] fp32 (= [ksplit][M][Cd], the layout splitk_reduce reads); images beyond N clip
    cuuint64_t dims[5] = {(cuuint64_t)g0.Cd, (cuuint64_t)g0.Wr, (cuuint64_t)g0.Hr, (cuuint64_t)g0.N, (cuuint64_t)p.ksplit};
    cuuint64_t strides[4] = {(cuuint64_t)g0.Cd * 4, (cuuint64_t)g0.Wr * g0.Cd * 4, (cuuint64_t)g0.Hr * g0.Wr * g0.Cd * 4,
                             (cuuint64_t)g0.M * g0.Cd * 4};
    cuuint32_t box[5] = {16u, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn, 1u};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&maps.d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, partial, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(halo split-K partials) failed: %d", (int)r);
  } else if (p.tma_store) {
    // destination pixel grid of this problem: pixel (rh, rw) -> physical (rh*dsh + doh, rw*dsw + dow) of [N][Hd][Wd][Cd] fp32
    cuuint64_t dims[4] = {(cuuint64_t)g0.Cd, (cuuint64_t)g0.Wr, (cuuint64_t)g0.Hr, (cuuint64_t)g0.N};
    cuuint64_t strides[3] = {(cuuint64_t)g0.dsw * g0.Cd * 4, (cuuint64_t)g0.dsh * g0.Wd * g0.Cd * 4, (cuuint64_t)g0.Hd * g0.Wd * g0.Cd * 4};
    cuuint32_t box[4] = {16u, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    void* base = g0.dst + ((size_t)g0.doh * g0.Wd + g0.dow) * g0.Cd;
    CUresult r = enc(&maps.d, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(halo D) failed: %d", (int)r);
  } else {
    maps.d = maps.a[0][0];
  }
  const size_t smem = (size_t)stagesA * slotA + (size_t)stagesB * slotB + (p.tma_store ? 4 * HALO_OUT_SLAB : 0) + 1024 /*alignment*/ +
                      512 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MOG_ERR_CUDA, "conv_halo_kernel smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const long long grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  conv_halo_kernel<<<(unsigned)grid, HALO_THREADS, smem, st>>>(maps, p);
  rc = check_launch("conv_halo_kernel");
  if (rc || !p.split_store) return rc;
  IGemmParams gr = g0;          // destination / bias / activation of the (merged) problem
  return launch_splitk_reduce(partial, gr, p.ksplit, st);
}

}  // namespace mog
