// conv_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a.
//
//   D[M x N] = A[M x K] * B[K x N]      M = output pixels (gathered rows), N = output channels,
//                                        K = taps x input channels
//
// * A is never materialised: 8 producer warps gather the fp32 NHWC pixels of each (tap, channel
//   chunk), split them on the fly into bf16 hi (+ lo for the 3-pass mode) and store them into
//   shared memory in the canonical K-major SWIZZLE_128B UMMA layout (8-row x 128-byte atoms).
//   Nearest-2x upsampling, stride, padding and the per-phase tap subsets of the data gradient
//   are all just index arithmetic of the gather (IGemmParams).
// * B (weights) is pre-packed once per optimiser step into bf16 hi/lo planes [N][K] (K-major),
//   copied by the producers into the same swizzled layout (all loads of a stage issued before
//   the first store, so a stage costs one memory round trip).
// * One elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16 per instruction, fp32
//   accumulator in TMEM).  MOG_PREC_BF16X3 issues three MMAs per k-step
//   (a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) into the same accumulator: fp32-equivalent products
//   at one third of the bf16 rate; MOG_PREC_BF16 issues only the first.
// * mbarrier pipeline: full[s] (256 producer arrivals) / empty[s] (tcgen05.commit) over `stages`
//   shared-memory stages; accum barrier (tcgen05.commit) hands the TMEM tile to the epilogue.
// * Split-K (grid.z) for problems whose M x N tiling cannot fill 148 SMs (the deep discriminator
//   layers: M = 512, K up to 27648): fp32 partial tiles go to a workspace and a deterministic
//   reduce kernel applies bias/activation and scatters to the NHWC destination.
// * Epilogue: the 8 producer warps read TMEM with tcgen05.ld (32x32b.x16; warp w and w+4 share a
//   lane quarter and split the columns), add bias / apply the activation, store fp32 NHWC rows.
//
// The weight gradient (reduction over pixels, both operands MN-major) is in conv_tc_wgrad.cu.
#include <cstdlib>

#include <cuda.h>
#include <cuda_bf16.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mog {
namespace tc {

constexpr int CPROD = 256;             // producer threads (8 warps)
constexpr int CTHREADS = CPROD + 32;   // + MMA warp (warp 8)

struct TcParams {
  IGemmParams g;
  const __nv_bfloat16* Xhi;  // gathered operand as pre-split bf16 planes [pixels][Cs] (PLANES kernel) or nullptr
  const __nv_bfloat16* Xlo;
  const __nv_bfloat16* Bhi;  // [Npad][Kpad]
  const __nv_bfloat16* Blo;  // [Npad][Kpad] (3-pass mode) or nullptr
  float* partial;            // split-K workspace [splits][M][Cd] or nullptr
  int Kpad;
  int BN;
  int passes;
  int stages;
  int tmem_cols;
  int splits;      // grid.z
  int kc_per_split;
};

// B (packed weights [Npad][Kpad], K-major) as TMA tensor maps, hi and lo plane: the PLANES kernel stages every B chunk
// with ONE bulk tensor copy per plane (box {64 k, BN rows}, SWIZZLE_128B = the layout the MMA descriptors expect) instead
// of BN*8 16-byte cp.async per plane: with B = 2/3 of a stage's bytes the producers' LDGSTS issue rate (~8 clk per warp
// instruction) had capped the tensor pipe at ~22 % on the weight-heavy discriminator layers.
struct TcMaps {
  CUtensorMap b[2];
};

__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// raises the barrier's pending transaction count without arriving (the issuing thread still arrives like every producer)
__device__ __forceinline__ void tc_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <bool PLANES>
__global__ void __launch_bounds__(CTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned base (SWIZZLE_128B atoms)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const IGemmParams& g = p.g;
  const int nplanes = p.passes == 3 ? 2 : 1;
  const int a_plane = BM * 128;        // bytes of one A plane (128 rows x 64 bf16)
  const int b_plane = p.BN * 128;      // bytes of one B plane
  const int stage_bytes = nplanes * (a_plane + b_plane);
  unsigned char* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accum = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  if (warp == 8) tmem_alloc(tmem_slot, p.tmem_cols);
  if (t == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], CPROD);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * p.BN;
  const int nk_total = p.Kpad / BK;
  const int kc_begin = blockIdx.z * p.kc_per_split;
  int kc_end = kc_begin + p.kc_per_split;
  if (kc_end > nk_total) kc_end = nk_total;
  const int nk = kc_end - kc_begin;   // >= 1 by construction

  if (warp < 8) {
    // ===================== producers: gather + split + swizzled store ======================
    const int r = t & 127;   // tile row
    const int half = t >> 7; // which 4 of the 8 channel chunks of a k-chunk
    const long long m = m0 + r;
    const bool row_ok = m < g.M;
    int rn = 0, h0 = 0, w0 = 0, rh = 0, rw = 0;
    if (row_ok) {
      rw = (int)(m % g.Wr);
      long long q = m / g.Wr;
      rh = (int)(q % g.Hr);
      rn = (int)(q / g.Hr);
      h0 = rh * g.rs;
      w0 = rw * g.rs;
    }
    const int HL = g.Hs << g.up2x, WL = g.Ws << g.up2x;
    const int vs = g.vstep > 0 ? g.vstep : 1, Hp = g.vstep > 0 ? g.Hp : g.Hs, Wp = g.vstep > 0 ? g.Wp : g.Ws;
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    const int rx = r & 7;
    // running decode of k -> (tap, channel); chunks of 8 channels never straddle a tap (Cs % 8 == 0)
    int c, th, tw;
    {
      const long long k0 = (long long)kc_begin * BK + half * 32;
      int tap = (int)(k0 / g.Cs);
      c = (int)(k0 - (long long)tap * g.Cs);
      th = tap / g.ntw;
      tw = tap - th * g.ntw;
    }
    const int nB = p.BN * 8;   // 16-byte chunks of one B plane per stage
    if (PLANES) {
      // ---- pre-split bf16 planes: pure 16-byte async copies (cp.async, zero-fill for padding), no ALU work.
      // Completion is signalled LAG stages late so several stages of copies stay in flight per thread.
      const int LAG = p.stages >= 3 ? 2 : 1;
      for (int it = 0; it < nk; ++it) {
        const int kc = kc_begin + it;
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + a_plane;
        const uint32_t b_hi = st + nplanes * a_plane, b_lo = b_hi + b_plane;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool inb = false;
          size_t off = 0;
          if (row_ok && th < g.nth) {
            const int sh = h0 + g.off_h[th], sw = w0 + g.off_w[tw];
            if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
              inb = true;
              off = (((size_t)rn * Hp + ((sh >> g.up2x) * vs + g.voh)) * Wp + ((sw >> g.up2x) * vs + g.vow)) * g.Cs + c;
            }
          }
          const uint32_t soff = row_off + (uint32_t)(((half * 4 + j) ^ rx) << 4);
          cp_async16(a_hi + soff, p.Xhi + off, inb ? 16u : 0u);
          if (nplanes == 2) cp_async16(a_lo + soff, p.Xlo + off, inb ? 16u : 0u);
          c += 8;
          if (c >= g.Cs) {
            c = 0;
            if (++tw == g.ntw) { tw = 0; ++th; }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // skip the other half's 4 chunks
          c += 8;
          if (c >= g.Cs) {
            c = 0;
            if (++tw == g.ntw) { tw = 0; ++th; }
          }
        }
        if (t == 0) {
          // B chunk of this stage: one bulk tensor copy per plane, completing on the stage's full barrier
          tc_expect_tx(&full[s], (uint32_t)(nplanes * b_plane));
          tc_tma_2d(b_hi, &maps.b[0], &full[s], kc * BK, n0);
          if (nplanes == 2) tc_tma_2d(b_lo, &maps.b[1], &full[s], kc * BK, n0);
        }
        cp_async_commit();
        if (it >= LAG) {
          if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
          fence_proxy_async();
          mbar_arrive(&full[(it - LAG) % p.stages]);
        }
      }
      if (LAG == 2 && nk >= 2) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(&full[(nk - 2) % p.stages]);
      }
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(&full[(nk - 1) % p.stages]);
    } else {
    for (int it = 0; it < nk; ++it) {
      const int kc = kc_begin + it;
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)((it / p.stages) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* st = smem + (size_t)s * stage_bytes;
      unsigned char* a_hi = st;
      unsigned char* a_lo = st + a_plane;                 // valid only when nplanes == 2
      unsigned char* b_hi = st + nplanes * a_plane;
      unsigned char* b_lo = b_hi + b_plane;
      // ---- issue all global loads of this stage first: A (4 chunks of 8 fp32) ...
      float4 v[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        v[j][1] = v[j][0];
        if (row_ok && th < g.nth) {
          const int sh = h0 + g.off_h[th], sw = w0 + g.off_w[tw];
          if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
            const size_t off = (((size_t)rn * Hp + ((sh >> g.up2x) * vs + g.voh)) * Wp + ((sw >> g.up2x) * vs + g.vow)) * g.Cs + c;
            const float4* sp = reinterpret_cast<const float4*>(g.src + off);
            v[j][0] = __ldg(sp);
            v[j][1] = __ldg(sp + 1);
          }
        }
        c += 8;
        if (c >= g.Cs) {
          c = 0;
          if (++tw == g.ntw) { tw = 0; ++th; }
        }
      }
      // skip the other half's 4 chunks
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c += 8;
        if (c >= g.Cs) {
          c = 0;
          if (++tw == g.ntw) { tw = 0; ++th; }
        }
      }
      // ... and B (bf16 weights): up to 8 chunks per thread per plane
      const size_t kbase = (size_t)kc * BK;
      uint4 bh[8], bl[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * CPROD;
        if (i < nB) {
          const int row = i >> 3, ch = i & 7;
          const size_t goff = (size_t)(n0 + row) * p.Kpad + kbase + ch * 8;
          bh[u] = __ldg(reinterpret_cast<const uint4*>(p.Bhi + goff));
          if (nplanes == 2) bl[u] = __ldg(reinterpret_cast<const uint4*>(p.Blo + goff));
        }
      }
      // ---- convert + store A
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f[8] = {v[j][0].x, v[j][0].y, v[j][0].z, v[j][0].w, v[j][1].x, v[j][1].y, v[j][1].z, v[j][1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
          hi[e] = *reinterpret_cast<uint32_t*>(&h2);
          if (nplanes == 2) {
            float2 hf = __bfloat1622float2(h2);
            __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
            lo[e] = *reinterpret_cast<uint32_t*>(&l2);
          }
        }
        const uint32_t off = row_off + (uint32_t)(((half * 4 + j) ^ rx) << 4);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (nplanes == 2) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      // ---- store B
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * CPROD;
        if (i < nB) {
          const int row = i >> 3, ch = i & 7;
          const uint32_t soff = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(b_hi + soff) = bh[u];
          if (nplanes == 2) *reinterpret_cast<uint4*>(b_lo + soff) = bl[u];
        }
      }
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(&full[s]);
    }

    }

    // ===================== epilogue: TMEM -> registers -> global ===========================
    mbar_wait(accum, 0);
    tcgen05_fence_after();
    const int q4 = warp & 3;               // TMEM lane quarter of this warp
    const int er = q4 * 32 + lane;         // tile row handled in the epilogue
    const long long em = m0 + er;
    const bool e_ok = em < g.M;
    float* dptr = nullptr;
    if (e_ok) {
      if (p.partial) {
        dptr = p.partial + ((size_t)blockIdx.z * g.M + em) * g.Cd + n0;
      } else {
        int ew = (int)(em % g.Wr);
        long long q = em / g.Wr;
        int eh = (int)(q % g.Hr);
        int en = (int)(q / g.Hr);
        size_t pix = ((size_t)en * g.Hd + (eh * g.dsh + g.doh)) * g.Wd + (ew * g.dsw + g.dow);
        dptr = g.dst + pix * g.Cd + n0;
      }
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int ncol16 = p.BN / 16;
    const int cbeg = (warp >> 2) ? (ncol16 + 1) / 2 : 0;     // warps 4-7 take the upper column half
    const int cend = (warp >> 2) ? ncol16 : (ncol16 + 1) / 2;
    const bool raw = p.partial != nullptr;
    const bool vec = (g.Cd & 3) == 0;
    for (int cb = cbeg; cb < cend; ++cb) {
      const int c0 = cb * 16;
      uint32_t acc[16];
      tmem_ld16(taddr + (uint32_t)c0, acc);
      if (e_ok && n0 + c0 < g.Cd)
        epi_store16(acc, dptr + c0, (!raw && g.bias) ? g.bias + n0 + c0 : nullptr, g.Cd - n0 - c0, vec, !raw && g.accum_dst != 0,
                    raw ? (int)MOG_ACT_NONE : g.act);
    }
  } else {
    // ===================== MMA issuer (warp 8) ==============================================
    const uint32_t idesc = make_idesc_bf16(BM, p.BN);
    for (int it = 0; it < nk; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)((it / p.stages) & 1);
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + a_plane;
        const uint32_t b_hi = st + nplanes * a_plane, b_lo = b_hi + b_plane;
        for (int pass = 0; pass < p.passes; ++pass) {
          const uint32_t ab = pass == 1 ? a_lo : a_hi;   // 0: hi*hi   1: lo*hi   2: hi*lo
          const uint32_t bb = pass == 2 ? b_lo : b_hi;
#pragma unroll
          for (int k16 = 0; k16 < BK / 16; ++k16) {
            const uint64_t da = make_desc_sw128(ab + k16 * 32);
            const uint64_t db = make_desc_sw128(bb + k16 * 32);
            umma_bf16(tmem_base, da, db, idesc, (it | pass | k16) != 0);
          }
        }
        umma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
        if (it == nk - 1) umma_commit(accum);
      }
      __syncwarp();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, p.tmem_cols);
}

__device__ __forceinline__ float epi_act(float v, int act) {
  if (act == MOG_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == MOG_ACT_TANH) return tanhf(v);
  if (act == MOG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == MOG_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

// split-K reduce: dst[pix(m), n] = act(sum_z partial[z][m][n] + bias[n])
// V = 4: four consecutive channels per thread with 128-bit accesses (Cd % 4 == 0); the split loop is unrolled by 4 so
// that four partial loads are in flight (fixed summation order: z ascending)
template <int V>
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, const IGemmParams g, int splits) {
  const size_t total = (size_t)g.M * g.Cd;
  const size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (idx >= total) return;
  const int n = (int)(idx % g.Cd);
  const long long m = (long long)(idx / g.Cd);
  float s[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s[j] = 0.f;
  if (V == 4) {
    int z = 0;
    for (; z + 4 <= splits; z += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(partial + (size_t)(z + u) * total + idx));
#pragma unroll
      for (int u = 0; u < 4; ++u) { s[0] += v[u].x; s[1 % V] += v[u].y; s[2 % V] += v[u].z; s[3 % V] += v[u].w; }
    }
    for (; z < splits; ++z) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (size_t)z * total + idx));
      s[0] += v.x; s[1 % V] += v.y; s[2 % V] += v.z; s[3 % V] += v.w;
    }
  } else {
    for (int z = 0; z < splits; ++z) s[0] += partial[(size_t)z * total + idx];
  }
  int rw = (int)(m % g.Wr);
  long long q = m / g.Wr;
  int rh = (int)(q % g.Hr);
  int rn = (int)(q / g.Hr);
  size_t pix = ((size_t)rn * g.Hd + (rh * g.dsh + g.doh)) * g.Wd + (rw * g.dsw + g.dow);
  float* dp = g.dst + pix * g.Cd + n;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    if (g.bias) s[j] += __ldg(g.bias + n + j);
    if (g.accum_dst) s[j] += dp[j];
    s[j] = epi_act(s[j], g.act);
  }
  if (V == 4) *reinterpret_cast<float4*>(dp) = make_float4(s[0], s[1 % V], s[2 % V], s[3 % V]);
  else dp[0] = s[0];
}

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 OIHW -> bf16 hi/lo planes [Npad][Kpad], k = local_tap * Cs + c
// ---------------------------------------------------------------------------------------------
struct PackArgs {
  const float* w;  // OIHW
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;  // may be null
  int Cout, Cin, KHW;
  int transpose;  // 0: n = co, c = ci (forward)    1: n = ci, c = co (data gradient)
  int ntaps;
  int taps[64][4];   // up to 4 filter taps summed into one local tap (sub-pixel phases of upsample+conv); -1 = unused
  int Nreal, Npad, Cs, CsReal, K, Kpad;   // Cs: channel pitch of k (multiple of 8), CsReal: channels that exist
};

// Tiled through shared memory so that both the OIHW reads and the [N][K] writes are coalesced: a block owns
// 16 output rows n x 32 channels c (all filter taps).  (The previous one-thread-per-output version read with a
// stride of KH*KW floats (forward) or Cin*KH*KW floats (data gradient) between consecutive threads and took
// 3.8 ms per training step for the 236 M parameters of the four networks.)
constexpr int PK_N = 16, PK_C = 32;
// KHW_T: filter taps as a compile-time constant (index arithmetic without integer division); 0 = generic (<= 16 taps)
// NR: rows n per block, TP: tap pitch of the tile (odd: conflict-free; >= taps).  <.., 16, 17> serves filters up to 16 taps,
// <25, 8, 25> the 5x5 filters of the image encoder.
template <int KHW_T, int NR = PK_N, int TP = 17>
__global__ void __launch_bounds__(256) pack_tc_kernel(const PackArgs a) {
  __shared__ float tile[NR * PK_C * TP];   // [n][c][tap]
  const int n0 = blockIdx.y * NR, c0 = blockIdx.x * PK_C;
  const int KHW = KHW_T ? KHW_T : a.KHW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  // ---- load: element (nl, cl, t) <- w[co][ci][t]
  if (!a.transpose) {
    // n = co, c = ci: for a fixed n the (c, t) block of PK_C * KHW floats is contiguous in memory
    for (int nl = ty; nl < NR; nl += 8) {
      const int n = n0 + nl;
      const float* src = a.w + ((size_t)n * a.Cin + c0) * KHW;
      for (int j = tx; j < PK_C * KHW; j += 32) {
        const int cl = j / KHW, t = j - cl * KHW;
        float v = 0.f;
        if (n < a.Nreal && c0 + cl < a.CsReal) v = __ldg(src + j);
        tile[(nl * PK_C + cl) * TP + t] = v;
      }
    }
  } else {
    // n = ci, c = co: for a fixed c the (n, t) block of PK_N * KHW floats is contiguous in memory
    for (int cl = ty; cl < PK_C; cl += 8) {
      const int c = c0 + cl;
      const float* src = a.w + ((size_t)c * a.Cin + n0) * KHW;
      for (int j = tx; j < NR * KHW; j += 32) {
        const int nl = j / KHW, t = j - nl * KHW;
        float v = 0.f;
        if (n0 + nl < a.Nreal && c < a.CsReal) v = __ldg(src + j);
        tile[(nl * PK_C + cl) * TP + t] = v;
      }
    }
  }
  __syncthreads();
  // ---- store: out[n][tl * Cs + c] = sum of the filter taps folded into local tap tl; 32 consecutive c per warp
  const int c = c0 + tx;
  if (c < a.Cs) {
    for (int nl = ty; nl < NR; nl += 8) {
      const int n = n0 + nl;
      if (n >= a.Npad) break;
      for (int tl = 0; tl < a.ntaps; ++tl) {
        float v = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int tp = a.taps[tl][u];
          if (tp >= 0) v += tile[(nl * PK_C + tx) * TP + tp];
        }
        const size_t o = (size_t)n * a.Kpad + (size_t)tl * a.Cs + c;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        a.hi[o] = h;
        if (a.lo) a.lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
  // ---- zero the K tail [K, Kpad) of this block's rows (once per row block)
  if (blockIdx.x == 0) {
    const int tail = a.Kpad - a.K;
    for (int nl = ty; nl < NR; nl += 8) {
      const int n = n0 + nl;
      if (n >= a.Npad) break;
      for (int k = a.K + tx; k < a.Kpad; k += 32) {
        a.hi[(size_t)n * a.Kpad + k] = __float2bfloat16_rn(0.f);
        if (a.lo) a.lo[(size_t)n * a.Kpad + k] = __float2bfloat16_rn(0.f);
      }
    }
    (void)tail;
  }
}

// Multi-tensor form: one launch repacks every (weight, problem) entry of a device-resident table (mog_pack_multi).
// Same tiling as pack_tc_kernel<0, 16, 17>; block -> entry by binary search over the entries' first blocks.
constexpr int PM_ROW = PK_C * 17 + 1;

// load the block's source tile [16 n][32 c][taps] of entry a (all filter taps); KHW_T: compile-time tap count
template <int KHW_T>
__device__ __forceinline__ void pack_multi_load(const MogPackEntry& a, int bx, int by, float* tile) {
  const int n0 = by * PK_N, c0 = bx * PK_C;
  const int KHW = KHW_T ? KHW_T : a.KHW;      // no integer division by a runtime value in the copy loops
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!a.transpose) {
    for (int nl = ty; nl < PK_N; nl += 8) {
      const int nn = n0 + nl;
      const float* src = a.w + ((size_t)nn * a.Cin + c0) * KHW;
      // interior tiles: the PK_C * KHW floats of a row are one contiguous run -> 128-bit loads (ncu: the scalar loop kept
      // 8 x 4 bytes per thread in flight and ran at 24 % of the HBM peak, long-scoreboard bound)
      if (nn < a.Nreal && c0 + PK_C <= a.CsReal && ((PK_C * KHW) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll 4
        for (int j4 = tx; j4 < (PK_C * KHW) >> 2; j4 += 32) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            const int cl = j / KHW, t = j - cl * KHW;
            tile[nl * PM_ROW + cl * 17 + t] = vv[e];
          }
        }
        continue;
      }
#pragma unroll 8
      for (int j = tx; j < PK_C * KHW; j += 32) {
        const int cl = j / KHW, t = j - cl * KHW;
        float v = 0.f;
        if (nn < a.Nreal && c0 + cl < a.CsReal) v = __ldg(src + j);
        tile[nl * PM_ROW + cl * 17 + t] = v;
      }
    }
  } else {
    for (int cl = ty; cl < PK_C; cl += 8) {
      const int c = c0 + cl;
      const float* src = a.w + ((size_t)c * a.Cin + n0) * KHW;
      if (c < a.CsReal && n0 + PK_N <= a.Nreal && ((PK_N * KHW) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll 4
        for (int j4 = tx; j4 < (PK_N * KHW) >> 2; j4 += 32) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j4);
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            const int nl = j / KHW, t = j - nl * KHW;
            tile[nl * PM_ROW + cl * 17 + t] = vv[e];
          }
        }
        continue;
      }
#pragma unroll 8
      for (int j = tx; j < PK_N * KHW; j += 32) {
        const int nl = j / KHW, t = j - nl * KHW;
        float v = 0.f;
        if (n0 + nl < a.Nreal && c < a.CsReal) v = __ldg(src + j);
        tile[nl * PM_ROW + cl * 17 + t] = v;
      }
    }
  }
}

// write the tile into entry a: an item = (local tap tl, row nl, group of 8 consecutive channels) -> one 16-byte store per
// plane; the n rows of the tile are PM_ROW = 32 * 17 + 1 floats apart so that the 32 items of a warp read 32 different banks
__device__ __forceinline__ void pack_multi_store(const MogPackEntry& a, int bx, int by, const float* tile) {
  const int n0 = by * PK_N, c0 = bx * PK_C;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __nv_bfloat16* phi = static_cast<__nv_bfloat16*>(a.hi);
  __nv_bfloat16* plo = static_cast<__nv_bfloat16*>(a.lo);
  const int ngr = PK_C / 8, per_tap = PK_N * ngr;
  for (int it = threadIdx.x; it < a.ntaps * per_tap; it += 256) {
    const int tl = it / per_tap, rem = it - tl * per_tap;
    const int nl = rem / ngr, gq = rem - nl * ngr;
    const int nn = n0 + nl, cg = c0 + gq * 8;
    if (nn >= a.Npad || cg >= a.Cs) continue;
    const int t0 = a.taps[tl][0], t1 = a.taps[tl][1], t2 = a.taps[tl][2], t3 = a.taps[tl][3];
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float* tp = tile + nl * PM_ROW + (gq * 8 + 2 * e + q) * 17;
        float x = tp[t0];
        if (t1 >= 0) x += tp[t1];
        if (t2 >= 0) x += tp[t2];
        if (t3 >= 0) x += tp[t3];
        v[q] = x;
      }
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[0], v[1]);
      h[e] = *reinterpret_cast<uint32_t*>(&h2);
      const float2 hf = __bfloat1622float2(h2);
      __nv_bfloat162 l2 = __floats2bfloat162_rn(v[0] - hf.x, v[1] - hf.y);
      l[e] = *reinterpret_cast<uint32_t*>(&l2);
    }
    const size_t o = (size_t)nn * a.Kpad + (size_t)tl * a.Cs + cg;
    *reinterpret_cast<uint4*>(phi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (plo) *reinterpret_cast<uint4*>(plo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  if (bx == 0) {     // zero the K tail [K, Kpad) of this block's rows
    for (int nl = ty; nl < PK_N; nl += 8) {
      const int nn = n0 + nl;
      if (nn >= a.Npad) break;
      for (int k = a.K + tx; k < a.Kpad; k += 32) {
        phi[(size_t)nn * a.Kpad + k] = __float2bfloat16_rn(0.f);
        if (plo) plo[(size_t)nn * a.Kpad + k] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

__global__ void __launch_bounds__(256, 4) pack_multi_kernel(const MogPackEntry* __restrict__ T, const MogPackGroup* __restrict__ G, int ng) {
  __shared__ float tile[PK_N * PM_ROW];
  __shared__ MogPackEntry ent;
  int lo = 0, hi = ng - 1;
  const int b = (int)blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&G[mid].block_start) <= b) lo = mid; else hi = mid - 1;
  }
  const int first = __ldg(&G[lo].first), count = __ldg(&G[lo].count), nxb = __ldg(&G[lo].nxb);
  const int local = b - __ldg(&G[lo].block_start);
  const int bx = local % nxb, by = local / nxb;
  for (int gi = 0; gi < count; ++gi) {
    // the entry (336 bytes) into shared memory: its taps table and geometry are read many times
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(sizeof(MogPackEntry) / 4); i += blockDim.x)
      reinterpret_cast<int*>(&ent)[i] = __ldg(reinterpret_cast<const int*>(&T[first + gi]) + i);
    __syncthreads();
    if (gi == 0) {   // the source tile is the same for every entry of the group
      switch (ent.KHW) {
        case 1: pack_multi_load<1>(ent, bx, by, tile); break;
        case 9: pack_multi_load<9>(ent, bx, by, tile); break;
        case 16: pack_multi_load<16>(ent, bx, by, tile); break;
        default: pack_multi_load<0>(ent, bx, by, tile); break;
      }
      __syncthreads();
    }
    pack_multi_store(ent, bx, by, tile);
  }
}

// Data gradient of a non-overlapping strided conv (stride == KH == KW, pad == 0, e.g. the 4x4/s4 logit heads,
// model.py:627,640) with a handful of output channels: every dx pixel sees exactly one tap,
//   dx[n, ho*s + kh, wo*s + kw, ci] = sum_co dy[n, ho, wo, co] * w[co, ci, kh, kw]
// One CUDA-core pass instead of s*s tensor-core launches with M = N*Ho*Wo rows each.  The weights are read
// from the per-phase packed buffer (hi + lo planes [Npad][Kpad], phase = kh*s + kw, k = co).
struct PatchDgradArgs {
  const float* dy;
  const __nv_bfloat16* packed;
  float* dx;
  int N, Ho, Wo, Cout, Cin, s;
  int Kpad;
  size_t plane_elems, phase_elems;   // elements per plane / per phase block (all planes, incl. alignment)
  int planes;
};
__global__ void patch_dgrad_kernel(const PatchDgradArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int H = a.Ho * a.s, W = a.Wo * a.s;
  const size_t total = (size_t)a.N * H * W * a.Cin;
  if (idx >= total) return;
  const int ci = (int)(idx % a.Cin);
  size_t r = idx / a.Cin;
  const int w = (int)(r % W); r /= W;
  const int h = (int)(r % H);
  const int n = (int)(r / H);
  const int ho = h / a.s, kh = h - ho * a.s, wo = w / a.s, kw = w - wo * a.s;
  const __nv_bfloat16* wp = a.packed + (size_t)(kh * a.s + kw) * a.phase_elems + (size_t)ci * a.Kpad;
  const float* dyp = a.dy + (((size_t)n * a.Ho + ho) * a.Wo + wo) * a.Cout;
  float acc = 0.f;
  for (int co = 0; co < a.Cout; ++co) {
    float wv = __bfloat162float(wp[co]);
    if (a.planes == 2) wv += __bfloat162float(wp[a.plane_elems + co]);
    acc += __ldg(dyp + co) * wv;
  }
  a.dx[idx] = acc;
}

// fp32 [rows][C] -> bf16 planes [rows][CP] (hi, and lo = bf16(x - hi) when nplanes == 2); CP = C rounded up
// to 8, pad channels zero.  One thread per 8-channel chunk.
// With y != nullptr the value split is x * act'(y) (x = upstream gradient, y = activation OUTPUT of a conv epilogue):
// the backward of the epilogue activation fused into the split, so dz never exists in fp32.
__global__ void split_planes_kernel(const float* __restrict__ x, long long rows, int C, int CP, __nv_bfloat16* hi,
                                    __nv_bfloat16* lo, const float* __restrict__ y, int act) {
  const int cpr = CP >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cpr) return;
  const long long r = idx / cpr;
  const int c0 = (int)(idx - r * cpr) * 8;
  float f[8];
  const float* xp = x + (size_t)r * C + c0;
  if (c0 + 7 < C && (C & 3) == 0) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(xp)), b = __ldg(reinterpret_cast<const float4*>(xp) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < C) ? __ldg(xp + j) : 0.f;
  }
  if (y) {
    const float* yp = y + (size_t)r * C + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = (c0 + j < C) ? __ldg(yp + j) : 0.f;
      switch (act) {
        case MOG_ACT_RELU: f[j] = v > 0.f ? f[j] : 0.f; break;
        case MOG_ACT_LRELU: f[j] = v > 0.f ? f[j] : 0.2f * f[j]; break;
        case MOG_ACT_TANH: f[j] *= 1.f - v * v; break;
        case MOG_ACT_SIGMOID: f[j] *= v * (1.f - v); break;
        default: break;
      }
    }
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    h[e] = *reinterpret_cast<uint32_t*>(&h2);
    float2 hf = __bfloat1622float2(h2);
    __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    l[e] = *reinterpret_cast<uint32_t*>(&l2);
  }
  const size_t o = (size_t)r * CP + c0;
  *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
}

// fp32 NHWC [N][H][W][C] -> bf16 planes of its im2col matrix [N*Ho*Wo][KP]: k = (kh*KW + kw)*C + c, KP = KH*KW*C rounded
// up to 8 (pad columns and out-of-image taps zero).  For convolutions with a handful of input channels (RGB images, 3-channel
// image gradients): with C = 3 every filter tap would be its own K = 16 MMA step and its own TMA box, 3 of 64 channels real;
// on the patch matrix the same conv is a 1x1 conv with KP channels.  One thread per 8-column chunk of a patch row.
__global__ void patch_planes_kernel(const float* __restrict__ x, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                                    int Ho, int Wo, int K, int KP, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  // grid: x = image row of the patch grid (n * Ho + ho), y = blocks over (wo, 8-column chunk): 32-bit index arithmetic only
  // (the flat 64-bit decomposition idx -> (n, ho, wo, chunk) cost four 64-bit divisions per thread: the kernel was
  // instruction bound at 1.7 TB/s of plane writes, ncu r2u)
  const int cpr = KP >> 3;
  const int item = blockIdx.y * blockDim.x + threadIdx.x;     // wo * cpr + chunk
  if (item >= Wo * cpr) return;
  const int wo = item / cpr;
  const int k0 = (item - wo * cpr) * 8;
  const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
  const long long r = (long long)blockIdx.x * Wo + wo;
  const int h0 = ho * stride - pad, w0 = wo * stride - pad;
  const float* xn = x + (size_t)n * H * W * C;
  float f[8];
  int tap = k0 / C, c = k0 - tap * C;
  int kh = tap / KW, kw = tap - kh * KW;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = 0.f;
    if (k0 + j < K) {
      const int hh = h0 + kh, ww = w0 + kw;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(xn + ((size_t)hh * W + ww) * C + c);
    }
    f[j] = v;
    if (++c == C) { c = 0; if (++kw == KW) { kw = 0; ++kh; } }
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    h[e] = *reinterpret_cast<uint32_t*>(&h2);
    float2 hf = __bfloat1622float2(h2);
    __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    l[e] = *reinterpret_cast<uint32_t*>(&l2);
  }
  const size_t o = (size_t)r * KP + k0;
  *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
}

// The adjoint of patch extraction (col2im) with the conv epilogue:
//   y[n, h, w, c] = act(bias[c] + sum over taps (kh, kw) with (h + pad - kh, w + pad - kw) = stride * (ho, wo), (ho, wo) inside
//                       the Ho x Wo grid, of z[(n, ho, wo)][(kh*KW + kw)*C + c])
// z = the output of a 1x1 problem with KP >= KH*KW*C columns (row pitch ldz).  Two uses, both for C <= 4:
//   * the data gradient of a conv with <= 4 INPUT channels: z = dy . W2, W2[(kh, kw, ci)][co] = w[co][ci][kh][kw] -- instead of
//     KH*KW/stride^2-tap stride phases whose N tile holds 3 real columns (D_NET256 layer 1: 4 x 89 us at 9.6 % tensor pipe);
//   * the FORWARD of a 'same' conv with <= 4 output channels (GET_IMAGE_G): z[q][(a, b, co)] = sum_ci x[q][ci] w[co][ci][KH-1-a][KW-1-b]
//     (0.42 ms at 12 % tensor pipe as a 9-tap conv with a 16-column N tile).
// One thread per output pixel, all C channels; the <= 16 taps are loaded up front (fixed summation order: tap ascending).
struct Col2imArgs {
  const float* z;
  const float* bias;
  float* y;
  int N, H, W, C, KH, KW, s, p, Ho, Wo, ldz, act;
  int nh, nw, pitch;   // rows / columns of z rows staged per tile, shared-memory pitch of a staged row (odd)
};
constexpr int C2I_TH = 8, C2I_TW = 32;   // output tile: 8 x 32 pixels, one per thread
// first / last patch-grid index along one axis that touches output positions [o0, o0 + T): (o + p - k) = s * g, 0 <= k < K
__host__ __device__ __forceinline__ int c2i_lo(int o0, int p, int K, int s) {
  const int v = o0 + p - (K - 1);
  return v <= 0 ? 0 : (v + s - 1) / s;
}
// One block per 8 x 32 tile of output pixels: the z rows the tile gathers from (a (8 + KH - 1)/s x (32 + KW - 1)/s patch of the
// Ho x Wo grid; consecutive grid columns are consecutive rows of z, so every staged line is one contiguous run) are loaded with
// 128-bit accesses into shared memory, then every thread sums the taps of its pixel from there (odd pitch: conflict free).
// (A first version gathered straight from global memory: 12-byte pieces of 9 different 128-byte rows per pixel -- 0.76 ms per
// step, no faster than the convolutions it replaced.)
__global__ void __launch_bounds__(C2I_TH * C2I_TW) col2im_act_kernel(const Col2imArgs a) {
  extern __shared__ float c2i_sm[];
  const int w0 = blockIdx.x * C2I_TW, h0 = blockIdx.y * C2I_TH, n = blockIdx.z;
  const int K = a.KH * a.KW * a.C;
  const int gh0 = c2i_lo(h0, a.p, a.KH, a.s), gw0 = c2i_lo(w0, a.p, a.KW, a.s);
  const float* zn = a.z + (size_t)n * a.Ho * a.Wo * a.ldz;
  // ---- stage: lines gh0 .. gh0 + nh - 1, columns gw0 .. gw0 + nw - 1 (clipped to the grid)
  const int f4_per_row = a.ldz >> 2;
  const int cols = min(a.nw, a.Wo - gw0), lines = min(a.nh, a.Ho - gh0);
  const int per_line = cols * f4_per_row, total = lines > 0 && cols > 0 ? lines * per_line : 0;
  const float4* zb = reinterpret_cast<const float4*>(zn);
  constexpr int NT = C2I_TH * C2I_TW;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * NT) {     // four 128-bit loads in flight per thread
    float4 v[4];
    int line[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * NT;
      line[u] = i / per_line;
      r[u] = i - line[u] * per_line;
      if (i < total) v[u] = __ldg(zb + ((size_t)(gh0 + line[u]) * a.Wo + gw0) * f4_per_row + r[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * NT >= total) break;
      const int col = r[u] / f4_per_row, k = (r[u] - col * f4_per_row) * 4;
      float* d = c2i_sm + ((size_t)line[u] * a.nw + col) * a.pitch + k;
      if (k < K) d[0] = v[u].x;
      if (k + 1 < K) d[1] = v[u].y;
      if (k + 2 < K) d[2] = v[u].z;
      if (k + 3 < K) d[3] = v[u].w;
    }
  }
  __syncthreads();
  // ---- gather
  const int tw = threadIdx.x % C2I_TW, th = threadIdx.x / C2I_TW;
  const int h = h0 + th, w = w0 + tw;
  if (h >= a.H || w >= a.W) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int kh = 0; kh < a.KH; ++kh) {
    const int hh = h + a.p - kh;
    if (hh < 0) break;
    const int gh = a.s == 1 ? hh : (a.s == 2 ? hh >> 1 : hh / a.s);     // (no integer division for the strides that occur)
    if (gh * a.s != hh || gh >= a.Ho) continue;
    const float* line = c2i_sm + (size_t)(gh - gh0) * a.nw * a.pitch;
    for (int kw = 0; kw < a.KW; ++kw) {
      const int ww = w + a.p - kw;
      if (ww < 0) break;
      const int gw = a.s == 1 ? ww : (a.s == 2 ? ww >> 1 : ww / a.s);
      if (gw * a.s != ww || gw >= a.Wo) continue;
      const float* e = line + (gw - gw0) * a.pitch + (kh * a.KW + kw) * a.C;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < a.C) acc[c] += e[c];
    }
  }
  float* yp = a.y + (((size_t)n * a.H + h) * a.W + w) * a.C;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < a.C) yp[c] = epi_act(acc[c] + (a.bias ? __ldg(a.bias + c) : 0.f), a.act);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using namespace tc;

int launch_split_planes(const float* x, long long rows, int C, int CP, void* planes, int nplanes, cudaStream_t st, const float* y,
                        int act) {
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(planes);
  __nv_bfloat16* lo = nplanes == 2 ? hi + (size_t)rows * CP : nullptr;
  const long long n = rows * (CP / 8);
  tc::split_planes_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(x, rows, C, CP, hi, lo, y, act);
  return check_launch("split_planes_kernel");
}

int launch_patch_planes(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo,
                        void* planes, int nplanes, cudaStream_t st) {
  const int K = KH * KW * C, KP = ceil_div(K, 8) * 8;
  const long long rows = (long long)N * Ho * Wo;
  __nv_bfloat16* hi = static_cast<__nv_bfloat16*>(planes);
  __nv_bfloat16* lo = nplanes == 2 ? hi + (size_t)rows * KP : nullptr;
  const int per_line = Wo * (KP / 8);
  if ((long long)N * Ho > 0x7fffffffLL || ceil_div(per_line, 256) > 65535) return fail(MOG_ERR_UNSUPPORTED, "mog_patch_planes: problem too large");
  tc::patch_planes_kernel<<<dim3((unsigned)(N * Ho), (unsigned)ceil_div(per_line, 256)), 256, 0, st>>>(x, N, H, W, C, KH, KW, stride, pad, Ho,
                                                                                                       Wo, K, KP, hi, lo);
  return check_launch("patch_planes_kernel");
}

int launch_col2im_act(const float* z, int ldz, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo,
                      const float* bias, int act, float* y, cudaStream_t st) {
  tc::Col2imArgs a{z, bias, y, N, H, W, C, KH, KW, stride, pad, Ho, Wo, ldz, act, 0, 0, 0};
  // staged patch of the grid per tile: indices g with s*g in [o0 + p - (K-1), o0 + T - 1 + p]
  a.nh = (tc::C2I_TH - 1 + KH - 1) / stride + 2;
  a.nw = (tc::C2I_TW - 1 + KW - 1) / stride + 2;
  const int K = KH * KW * C;
  a.pitch = K | 1;
  const size_t smem = sizeof(float) * (size_t)a.nh * a.nw * a.pitch;
  if (smem > 200 * 1024) return fail(MOG_ERR_UNSUPPORTED, "mog_col2im_act: %zu bytes of shared memory", smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(tc::col2im_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MOG_ERR_UNSUPPORTED, "mog_col2im_act: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
  }
  if (N > 65535) return fail(MOG_ERR_UNSUPPORTED, "mog_col2im_act: batch too large");
  dim3 grid((unsigned)ceil_div(W, tc::C2I_TW), (unsigned)ceil_div(H, tc::C2I_TH), (unsigned)N);
  tc::col2im_act_kernel<<<grid, tc::C2I_TH * tc::C2I_TW, smem, st>>>(a);
  return check_launch("col2im_act_kernel");
}

int tc_bn_for(int Cd) {
  static int cap = 0;
  if (!cap) {
    const char* e = getenv("MOG_TC_BN_CAP");   // tuning knob: largest N tile of the generic kernel (default 256)
    cap = e ? atoi(e) : 256;
    if (cap < 16 || cap > 256) cap = 256;
  }
  int cpad = ceil_div(Cd, 16) * 16;
  int tiles = ceil_div(cpad, cap);
  int bn = ceil_div(ceil_div(cpad, tiles), 16) * 16;
  return bn;
}


// layout of the packed buffer of one gather-GEMM problem (K = ntaps * Cs, N = Cd)
static TcWeightLayout tc_weight_layout(int ntaps, int Cs, int Cd, int passes) {
  TcWeightLayout L;
  L.BN = tc_bn_for(Cd);
  L.ntiles = ceil_div(Cd, L.BN);
  L.Npad = L.ntiles * L.BN;
  L.K = ntaps * Cs;
  L.Kpad = ceil_div(L.K > 0 ? L.K : 1, BK) * BK;
  L.plane_elems = (size_t)L.Npad * L.Kpad;
  L.planes = passes == 3 ? 2 : 1;
  return L;
}

size_t tc_packed_bytes(int ntaps, int Cs, int Cd, int passes) {
  TcWeightLayout L = tc_weight_layout(ntaps, Cs, Cd, passes);
  size_t b = L.plane_elems * L.planes * 2;
  return (b + 255) / 256 * 256;  // keep every phase 256-byte aligned
}


bool patch_dgrad_eligible(const MogConvDesc& d) {
  return !d.up2x && d.stride > 1 && d.stride == d.KH && d.stride == d.KW && d.pad == 0 && d.Cout <= 8 && (d.H % d.stride) == 0 &&
         (d.W % d.stride) == 0 && (long long)d.N * (d.H / d.stride) * (d.W / d.stride) < 8192;   // (small M: packed with pitch p8)
}

// packed: the dgrad weight buffer of mog_pack_weight (one block per stride phase, each one tap, pitch p8(Cout))
int launch_patch_dgrad(const MogConvDesc& d, const float* dy, const void* packed, float* dx, int passes, cudaStream_t st) {
  const int pitch = ceil_div(d.Cout, 8) * 8;
  TcWeightLayout L = tc_weight_layout(1, pitch, d.Cin, passes);
  PatchDgradArgs a;
  a.dy = dy; a.packed = static_cast<const __nv_bfloat16*>(packed); a.dx = dx;
  a.N = d.N; a.Ho = d.H / d.stride; a.Wo = d.W / d.stride; a.Cout = d.Cout; a.Cin = d.Cin; a.s = d.stride;
  a.Kpad = L.Kpad; a.plane_elems = L.plane_elems; a.planes = L.planes;
  a.phase_elems = tc_packed_bytes(1, pitch, d.Cin, passes) / 2;
  const size_t total = (size_t)d.N * d.H * d.W * d.Cin;
  patch_dgrad_kernel<<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, st>>>(a);
  return check_launch("patch_dgrad_kernel");
}

// k = local_tap * pitch + c  (pitch >= channel count, multiple of 8; channels beyond the real count are zero)
int tc_pack_pitch(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int transpose, int ntaps,
                  const int (*taps)[4], int pitch, int passes, cudaStream_t st) {
  const int CsReal = transpose ? Cout : Cin, Cd = transpose ? Cin : Cout;
  const int Cs = pitch;
  TcWeightLayout L = tc_weight_layout(ntaps, Cs, Cd, passes);
  PackArgs a;
  a.w = w_oihw;
  a.hi = static_cast<__nv_bfloat16*>(out);
  a.lo = L.planes == 2 ? a.hi + L.plane_elems : nullptr;
  a.Cout = Cout; a.Cin = Cin; a.KHW = KH * KW; a.transpose = transpose; a.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i)
    for (int u = 0; u < 4; ++u) a.taps[i][u] = taps[i][u];
  a.Nreal = Cd; a.Npad = L.Npad; a.Cs = Cs; a.CsReal = CsReal; a.K = L.K; a.Kpad = L.Kpad;
  if (KH * KW > 25) return fail(MOG_ERR_UNSUPPORTED, "weight packing (tcgen05): filters with more than 25 taps are not supported");
  if (KH * KW > 16) {
    pack_tc_kernel<0, 8, 25><<<dim3((unsigned)ceil_div(Cs, PK_C), (unsigned)ceil_div(L.Npad, 8)), 256, 0, st>>>(a);
    return check_launch("pack_tc_kernel");
  }
  const dim3 grid((unsigned)ceil_div(Cs, PK_C), (unsigned)ceil_div(L.Npad, PK_N));
  switch (KH * KW) {
    case 1: pack_tc_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 9: pack_tc_kernel<9><<<grid, 256, 0, st>>>(a); break;
    case 16: pack_tc_kernel<16><<<grid, 256, 0, st>>>(a); break;
    default: pack_tc_kernel<0><<<grid, 256, 0, st>>>(a); break;
  }
  return check_launch("pack_tc_kernel");
}

// table entry of the launch tc_pack_pitch would make (multi-tensor repacking)
int tc_pack_entry(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int transpose, int ntaps,
                  const int (*taps)[4], int pitch, int passes, MogPackEntry* e) {
  if (KH * KW > 16 || ntaps > 16) return MOG_ERR_UNSUPPORTED;
  const int CsReal = transpose ? Cout : Cin, Cd = transpose ? Cin : Cout;
  TcWeightLayout L = tc_weight_layout(ntaps, pitch, Cd, passes);
  e->w = w_oihw;
  e->hi = out;
  e->lo = L.planes == 2 ? static_cast<void*>(static_cast<__nv_bfloat16*>(out) + L.plane_elems) : nullptr;
  e->Cout = Cout; e->Cin = Cin; e->KHW = KH * KW; e->transpose = transpose; e->ntaps = ntaps;
  e->Nreal = Cd; e->Npad = L.Npad; e->Cs = pitch; e->CsReal = CsReal; e->K = L.K; e->Kpad = L.Kpad;
  e->nxb = ceil_div(pitch, PK_C); e->nyb = ceil_div(L.Npad, PK_N);
  e->block_start = 0;
  for (int i = 0; i < 16; ++i)
    for (int u = 0; u < 4; ++u) e->taps[i][u] = i < ntaps ? taps[i][u] : -1;
  return MOG_OK;
}

int launch_pack_multi(const MogPackEntry* entries_dev, const MogPackGroup* groups_dev, int ngroups, int total_blocks, cudaStream_t st) {
  tc::pack_multi_kernel<<<(unsigned)total_blocks, 256, 0, st>>>(entries_dev, groups_dev, ngroups);
  return check_launch("pack_multi_kernel");
}

// split-K factor for one gather-GEMM problem
static int tc_splits(long long M, int ntiles, int nk) {
  const long long ctas = ceil_div_ll(M, BM) * ntiles;
  if (ctas > kNumSMs || nk < 8) return 1;   // (96 CTAs of a K = 12288 layer on 148 SMs took 610 us in one under-filled wave)
  long long want = (2 * kNumSMs) / ctas;   // floor: whole waves
  long long maxs = nk / 4;
  long long s = want < maxs ? want : maxs;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

size_t tc_igemm_workspace_bytes(long long M, int ntaps, int Cs, int Cd, int passes) {
  TcWeightLayout L = tc_weight_layout(ntaps, Cs, Cd, passes);
  const int nk = L.Kpad / BK;
  int splits = tc_splits(M, L.ntiles, nk);
  if (splits <= 1) return 0;
  const int per = ceil_div(nk, splits);
  splits = ceil_div(nk, per);
  return splits > 1 ? (size_t)splits * M * Cd * sizeof(float) : 0;
}

typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeFn tc_encode_fn() {
  static TcEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TcEncodeFn>(ptr);
  }
  return fn;
}

int launch_igemm_tc(const IGemmParams& g, const void* packed, int passes, void* workspace, size_t ws_bytes,
                    cudaStream_t st) {
  TcWeightLayout L = tc_weight_layout(g.nth * g.ntw, g.Cs, g.Cd, passes);
  TcParams p;
  p.g = g;
  p.Bhi = static_cast<const __nv_bfloat16*>(packed);
  p.Blo = L.planes == 2 ? p.Bhi + L.plane_elems : nullptr;
  p.Kpad = L.Kpad;
  p.BN = L.BN;
  p.passes = passes;
  const int nk = L.Kpad / BK;
  int splits = tc_splits(g.M, L.ntiles, nk);
  int per = ceil_div(nk, splits);
  splits = ceil_div(nk, per);
  p.splits = splits;
  p.kc_per_split = per;
  p.partial = nullptr;
  if (splits > 1) {
    const size_t need = (size_t)splits * g.M * g.Cd * sizeof(float);
    if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "conv (tcgen05 split-K): workspace %zu < %zu", ws_bytes, need);
    p.partial = static_cast<float*>(workspace);
  }
  const int stage_bytes = L.planes * (BM * 128 + L.BN * 128);
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  int cols = 32;
  while (cols < L.BN) cols *= 2;
  p.tmem_cols = cols;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MOG_ERR_CUDA, "conv_tc_kernel smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div_ll(g.M, BM), (unsigned)L.ntiles, (unsigned)splits);
  TcMaps maps{};
  if (g.src_planes) {
    TcEncodeFn enc = tc_encode_fn();
    if (!enc) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    for (int pl = 0; pl < 2; ++pl) {
      const __nv_bfloat16* base = (pl == 1 && p.Blo) ? p.Blo : p.Bhi;   // (lo map aliases hi in single-pass mode: never used)
      cuuint64_t dims[2] = {(cuuint64_t)L.Kpad, (cuuint64_t)L.Npad};
      cuuint64_t strides[1] = {(cuuint64_t)L.Kpad * 2};
      cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)L.BN};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&maps.b[pl], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(MOG_ERR_CUDA, "cuTensorMapEncodeTiled(conv_tc B) failed: %d", (int)r);
    }
    p.Xhi = static_cast<const __nv_bfloat16*>(g.src_planes);
    p.Xlo = p.Xhi + g.src_plane_elems;
    conv_tc_kernel<true><<<grid, CTHREADS, smem, st>>>(maps, p);
  } else {
    p.Xhi = p.Xlo = nullptr;
    conv_tc_kernel<false><<<grid, CTHREADS, smem, st>>>(maps, p);
  }
  int rc = check_launch("conv_tc_kernel");
  if (rc || splits == 1) return rc;
  return launch_splitk_reduce(p.partial, g, splits, st);
}

// dst[pix(m), n] = act(sum_z partial[z][m][n] + bias[n]) (fixed summation order)
int launch_splitk_reduce(const float* partial, const IGemmParams& g, int splits, cudaStream_t st) {
  const size_t total = (size_t)g.M * g.Cd;
  if ((g.Cd & 3) == 0 && (reinterpret_cast<uintptr_t>(partial) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.dst) & 15) == 0)
    splitk_reduce_kernel<4><<<(unsigned)ceil_div_ll((long long)(total / 4), 256), 256, 0, st>>>(partial, g, splits);
  else
    splitk_reduce_kernel<1><<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, st>>>(partial, g, splits);
  return check_launch("splitk_reduce_kernel");
}

}  // namespace mog
