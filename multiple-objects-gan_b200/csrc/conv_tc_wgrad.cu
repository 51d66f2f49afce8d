// conv_tc_wgrad.cu -- tcgen05 weight gradient.
//
//   dW[kf, co] = sum_p  X[pix(p, tap(kf)), c(kf)] * dY[p, co]        kf = tap*Cin + c,  p = (n, ho, wo)
//
// GEMM view: M = 128 consecutive kf, N = BN output channels, reduction over pixels (64 per
// shared-memory stage, 16 per tcgen05.mma).  In NHWC both operands are contiguous along their
// M/N index for a fixed pixel, i.e. *MN-major*: they are staged in the canonical MN-major
// SWIZZLE_128B layout (128-byte rows of 64 M/N elements, 8 pixel rows per 1024-byte atom;
// LBO = stride between 64-element blocks, SBO = stride between 8-pixel groups) with no transpose.
// The pixel range is split across CTAs (grid.z); every CTA writes an fp32 partial tile to the
// workspace and the deterministic reduce kernel of conv_ffma.cu sums the splits and emits OIHW.
// Producers: 8 warps gather + split fp32 -> bf16 hi/lo; warp 8 issues the MMAs; warps 0-3 drain TMEM.
#include <cuda_bf16.h>

#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mog {
namespace tc {

constexpr int WPROD = 256;            // producer threads
constexpr int WTHREADS = WPROD + 32;  // + MMA warp
constexpr int PIX = 64;               // pixels per stage
constexpr int BLK_BYTES = 8192;       // one 64-element MN block x 64 pixels
constexpr int A_PLANE = 2 * BLK_BYTES;  // 128 kf

struct WgParams {
  const float* x;
  const float* dy;
  const __nv_bfloat16* xhi;   // PLANES kernel: pre-split bf16 planes [pixels][Cin] / [P][CoutP]
  const __nv_bfloat16* xlo;
  const __nv_bfloat16* dyhi;
  const __nv_bfloat16* dylo;
  int CoutP;                  // channel pitch of the dy planes (Cout rounded up to 8); == Cout for fp32 input
  float* ws;  // [splits][K][Cout]
  int N, H, W, Cin, up2x, Ho, Wo, Cout, KH, KW, stride, pad;
  long long P, chunk;
  int K, BN, nblkB, passes, stages, tmem_cols;
};

__device__ __forceinline__ void split_store(const float4& v0, const float4& v1, unsigned char* hi_p, unsigned char* lo_p,
                                            uint32_t off, bool two) {
  const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    hi[e] = *reinterpret_cast<uint32_t*>(&h2);
    float2 hf = __bfloat1622float2(h2);
    __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    lo[e] = *reinterpret_cast<uint32_t*>(&l2);
  }
  *reinterpret_cast<uint4*>(hi_p + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (two) *reinterpret_cast<uint4*>(lo_p + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <bool PLANES>
__global__ void __launch_bounds__(WTHREADS, 1) wgrad_tc_kernel(const WgParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool two = p.passes == 3;
  const int nplanes = two ? 2 : 1;
  const int b_plane = p.nblkB * BLK_BYTES;
  const int stage_bytes = nplanes * (A_PLANE + b_plane);
  unsigned char* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* accum = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  if (warp == 8) tmem_alloc(tmem_slot, p.tmem_cols);
  if (t == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], WPROD);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kf0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * p.BN;
  const long long p_begin = (long long)blockIdx.z * p.chunk;
  long long p_end = p_begin + p.chunk;
  if (p_end > p.P) p_end = p.P;
  const int nst = p_begin < p_end ? (int)((p_end - p_begin + PIX - 1) / PIX) : 0;

  if (warp < 8) {
    // ===================== producers ==========================================================
    const int pk = t & 63;   // pixel within the stage
    const int q = t >> 6;    // 0..3: which quarter of the chunk list
    const int kgrp = pk >> 3, kin = pk & 7;
    const int HL = p.H << p.up2x, WL = p.W << p.up2x;
    // A: this thread's 4 chunks (8 kf each): cm = q*4 + j ; fixed (tap, c) per chunk
    int a_oh[4], a_ow[4], a_c[4];
    bool a_ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kf = kf0 + (q * 4 + j) * 8;
      a_ok[j] = kf < p.K;
      int kk = a_ok[j] ? kf : 0;
      int tap = kk / p.Cin;
      a_c[j] = kk - tap * p.Cin;
      int kh = tap / p.KW, kw = tap - kh * p.KW;
      a_oh[j] = kh - p.pad;
      a_ow[j] = kw - p.pad;
    }
    const int nchB = p.BN / 8;  // 16-byte chunks per pixel of the B tile
    if (PLANES) {
      const int LAG = p.stages >= 3 ? 2 : 1;
      for (int it = 0; it < nst; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)((it / p.stages) & 1);
        mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + A_PLANE;
        const uint32_t b_hi = st + nplanes * A_PLANE, b_lo = b_hi + b_plane;
        const long long pp = p_begin + (long long)it * PIX + pk;
        const bool pix_ok = pp < p_end;
        int n = 0, ho = 0, wo = 0;
        if (pix_ok) {
          wo = (int)(pp % p.Wo);
          long long qq = pp / p.Wo;
          ho = (int)(qq % p.Ho);
          n = (int)(qq / p.Ho);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          bool inb = false;
          size_t off = 0;
          if (pix_ok && a_ok[j]) {
            const int sh = ho * p.stride + a_oh[j], sw = wo * p.stride + a_ow[j];
            if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
              inb = true;
              off = (((size_t)n * p.H + (sh >> p.up2x)) * p.W + (sw >> p.up2x)) * p.Cin + a_c[j];
            }
          }
          const int cm = q * 4 + j;
          const uint32_t soff = (uint32_t)((cm >> 3) * BLK_BYTES + kgrp * 1024 + kin * 128 + (((cm & 7) ^ kin) << 4));
          cp_async16(a_hi + soff, p.xhi + off, inb ? 16u : 0u);
          if (two) cp_async16(a_lo + soff, p.xlo + off, inb ? 16u : 0u);
        }
        for (int cb = q; cb < nchB; cb += 4) {
          const int co = n0 + cb * 8;
          const bool ok = pix_ok && co < p.CoutP;
          const size_t off = ok ? (size_t)pp * p.CoutP + co : 0;
          const uint32_t soff = (uint32_t)((cb >> 3) * BLK_BYTES + kgrp * 1024 + kin * 128 + (((cb & 7) ^ kin) << 4));
          cp_async16(b_hi + soff, p.dyhi + off, ok ? 16u : 0u);
          if (two) cp_async16(b_lo + soff, p.dylo + off, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (it >= LAG) {
          if (LAG == 2) cp_async_wait<2>(); else cp_async_wait<1>();
          fence_proxy_async();
          mbar_arrive(&full[(it - LAG) % p.stages]);
        }
      }
      if (nst > 0) {
        if (LAG == 2 && nst >= 2) {
          cp_async_wait<1>();
          fence_proxy_async();
          mbar_arrive(&full[(nst - 2) % p.stages]);
        }
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(&full[(nst - 1) % p.stages]);
      }
    } else {
    for (int it = 0; it < nst; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)((it / p.stages) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* st = smem + (size_t)s * stage_bytes;
      unsigned char* a_hi = st;
      unsigned char* a_lo = st + A_PLANE;
      unsigned char* b_hi = st + nplanes * A_PLANE;
      unsigned char* b_lo = b_hi + b_plane;
      const long long pp = p_begin + (long long)it * PIX + pk;
      const bool pix_ok = pp < p_end;
      int n = 0, ho = 0, wo = 0;
      if (pix_ok) {
        wo = (int)(pp % p.Wo);
        long long qq = pp / p.Wo;
        ho = (int)(qq % p.Ho);
        n = (int)(qq / p.Ho);
      }
      // ---- A (gathered input pixels)
      float4 v[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        v[j][1] = v[j][0];
        if (pix_ok && a_ok[j]) {
          const int sh = ho * p.stride + a_oh[j], sw = wo * p.stride + a_ow[j];
          if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
            const size_t off = (((size_t)n * p.H + (sh >> p.up2x)) * p.W + (sw >> p.up2x)) * p.Cin + a_c[j];
            const float4* sp = reinterpret_cast<const float4*>(p.x + off);
            v[j][0] = __ldg(sp);
            v[j][1] = __ldg(sp + 1);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cm = q * 4 + j;
        const uint32_t off = (uint32_t)((cm >> 3) * BLK_BYTES + kgrp * 1024 + kin * 128 + (((cm & 7) ^ kin) << 4));
        split_store(v[j][0], v[j][1], a_hi, a_lo, off, two);
      }
      // ---- B (dy rows)
      for (int cb = q; cb < nchB; cb += 4) {
        float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
        const int co = n0 + cb * 8;
        if (pix_ok && co < p.Cout) {   // Cout % 8 == 0 => whole chunk valid
          const float4* sp = reinterpret_cast<const float4*>(p.dy + (size_t)pp * p.Cout + co);
          w0 = __ldg(sp);
          w1 = __ldg(sp + 1);
        }
        const uint32_t off = (uint32_t)((cb >> 3) * BLK_BYTES + kgrp * 1024 + kin * 128 + (((cb & 7) ^ kin) << 4));
        split_store(w0, w1, b_hi, b_lo, off, two);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
    }
    // ===================== epilogue (warps 0-3) ===============================================
    if (warp < 4) {
      if (nst > 0) {
        mbar_wait(accum, 0);
        tcgen05_fence_after();
      }
      const int kf = kf0 + t;
      float* out = p.ws + ((size_t)blockIdx.z * p.K + kf) * p.Cout + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t acc[16];
        if (nst > 0) {
          tmem_ld16(taddr + (uint32_t)c0, acc);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0u;
        }
        if (kf < p.K) {
          if (n0 + c0 + 15 < p.Cout && (p.Cout & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(out + c0 + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                                      __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < p.Cout) out[c0 + j] = __uint_as_float(acc[j]);
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer (warp 8) ================================================
    const uint32_t idesc = make_idesc_bf16(BM, p.BN, 1, 1);  // both operands MN-major
    for (int it = 0; it < nst; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)((it / p.stages) & 1);
      mbar_wait(&full[s], ph);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + A_PLANE;
        const uint32_t b_hi = st + nplanes * A_PLANE, b_lo = b_hi + b_plane;
        for (int pass = 0; pass < p.passes; ++pass) {
          const uint32_t ab = pass == 1 ? a_lo : a_hi;
          const uint32_t bb = pass == 2 ? b_lo : b_hi;
#pragma unroll
          for (int k16 = 0; k16 < PIX / 16; ++k16) {
            const uint64_t da = make_desc_sw128_mn(ab + k16 * 2048, BLK_BYTES, 1024);
            const uint64_t db = make_desc_sw128_mn(bb + k16 * 2048, BLK_BYTES, 1024);
            umma_bf16(tmem_base, da, db, idesc, (it | pass | k16) != 0);
          }
        }
        umma_commit(&empty[s]);
        if (it == nst - 1) umma_commit(accum);
      }
      __syncwarp();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace tc

using namespace tc;

// fp32 operands need whole 8-channel chunks; pre-split planes are padded to multiples of 8 by the splitter
bool tc_wgrad_eligible(const MogConvDesc& d, bool planes) { return planes || ((d.Cin % 8) == 0 && (d.Cout % 8) == 0); }

static void wg_tiling(const MogConvDesc& d, int Ho, int Wo, int* BN, int* ntn, int* splits, long long* chunk) {
  *BN = tc_bn_for(d.Cout);
  *ntn = ceil_div(d.Cout, *BN);
  const long long P = (long long)d.N * Ho * Wo;
  const int K = d.KH * d.KW * (ceil_div(d.Cin, 8) * 8);
  long long tiles = (long long)ceil_div(K, BM) * (*ntn);
  long long want = (2 * kNumSMs) / tiles;   // floor: two full waves of CTAs, no ragged third wave
  long long maxs = ceil_div_ll(P, 4 * PIX);
  long long s = want < maxs ? want : maxs;
  if (s < 1) s = 1;
  if (s > 2048) s = 2048;
  long long c = ceil_div_ll(P, s);
  c = ceil_div_ll(c, PIX) * PIX;
  *chunk = c;
  *splits = (int)ceil_div_ll(P, c);
}

size_t tc_wgrad_workspace_bytes(const MogConvDesc& d, int Ho, int Wo) {
  int BN, ntn, splits;
  long long chunk;
  wg_tiling(d, Ho, Wo, &BN, &ntn, &splits, &chunk);
  return (size_t)splits * d.KH * d.KW * (ceil_div(d.Cin, 8) * 8) * d.Cout * sizeof(float);
}

int launch_wgrad_tc(const MogConvDesc& d, int Ho, int Wo, const float* x, const float* dy, const void* x_planes,
                    size_t x_plane_elems, const void* dy_planes, size_t dy_plane_elems, float* ws, int passes,
                    int* splits_out, cudaStream_t st) {
  // with planes the channel counts seen by the kernel are the padded ones (multiples of 8)
  const bool planes = x_planes != nullptr;
  const int CinP = planes ? ceil_div(d.Cin, 8) * 8 : d.Cin;
  WgParams p;
  p.x = x; p.dy = dy; p.ws = ws;
  p.xhi = static_cast<const __nv_bfloat16*>(x_planes);
  p.xlo = p.xhi ? p.xhi + x_plane_elems : nullptr;
  p.dyhi = static_cast<const __nv_bfloat16*>(dy_planes);
  p.dylo = p.dyhi ? p.dyhi + dy_plane_elems : nullptr;
  p.CoutP = planes ? ceil_div(d.Cout, 8) * 8 : d.Cout;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = CinP; p.up2x = d.up2x; p.Ho = Ho; p.Wo = Wo; p.Cout = d.Cout;
  p.KH = d.KH; p.KW = d.KW; p.stride = d.stride; p.pad = d.pad;
  p.P = (long long)d.N * Ho * Wo;
  p.K = d.KH * d.KW * CinP;
  int ntn, splits;
  wg_tiling(d, Ho, Wo, &p.BN, &ntn, &splits, &p.chunk);
  p.nblkB = ceil_div(p.BN, 64);
  p.passes = passes;
  const int nplanes = passes == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (A_PLANE + p.nblkB * BLK_BYTES);
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  int cols = 32;
  while (cols < p.BN) cols *= 2;
  p.tmem_cols = cols;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(MOG_ERR_CUDA, "wgrad_tc_kernel smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(ceil_div(p.K, BM), ntn, splits);
  if (planes)
    wgrad_tc_kernel<true><<<grid, WTHREADS, smem, st>>>(p);
  else
    wgrad_tc_kernel<false><<<grid, WTHREADS, smem, st>>>(p);
  *splits_out = splits;
  return check_launch("wgrad_tc_kernel");
}

}  // namespace mog
