// conv_ffma.cu -- exact-fp32 implicit-GEMM convolution on the CUDA cores (MOG_PREC_FP32).
//
// This is the precision reference path of libmog (and the kernel for the awkward shapes the
// tensor-core path does not take: Cin=3, Cout in {1,3}, ...).  One gather-GEMM kernel serves the
// forward conv and the data gradient (which, per stride phase, is a stride-1 gather with
// per-tap offsets), a second one the weight gradient (reduction over pixels, split across CTAs,
// deterministic two-stage reduce that also writes the OIHW layout of the state_dict).
//
// Layouts: activations NHWC fp32; forward B operand [KH*KW*Cin][Cout]; dgrad B operand
// [KH*KW*Cout][Cin].
#include "common.cuh"
#include "conv_common.cuh"

namespace mog {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == MOG_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
  if (act == MOG_ACT_TANH) return tanhf(v);
  if (act == MOG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == MOG_ACT_SIGMOID) return sigmoidf_(v);
  return v;
}

// ------------------------------------------------------------------------------------------
// gather-GEMM:  dst[row, n] = act( sum_{t,c} src[pix(row, t), c] * wmat[tapw[t]*Cs + c, n] + bias[n] )
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(NT, 2) igemm_kernel(const IGemmParams p) {
  __shared__ __align__(16) float As[BK][AS_LD];
  __shared__ __align__(16) float Bs[BK][BN];

  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A loader mapping: 4 k-quads x 64 rows, two rows per thread
  const int a_kq = t & 3;
  const int a_row = t >> 2;  // 0..63 (+64)
  int a_n[2], a_h[2], a_w[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    long long m = m0 + a_row + i * 64;
    a_ok[i] = m < p.M;
    long long mm = a_ok[i] ? m : 0;
    int rw = (int)(mm % p.Wr);
    long long q = mm / p.Wr;
    int rh = (int)(q % p.Hr);
    a_n[i] = (int)(q / p.Hr);
    a_h[i] = rh * p.rs;
    a_w[i] = rw * p.rs;
  }
  const int HL = p.Hs << p.up2x, WL = p.Ws << p.up2x;

  // ---- B loader mapping: 16 k-rows x 16 column-quads
  const int b_k = t >> 4;
  const int b_c = (t & 15) * 4;
  const bool b_vec = (p.Cd & 3) == 0;

  float a_reg[2][4];
  float b_reg[4];

  auto load_tiles = [&](int k0) {
    // A
    {
      int kf = k0 + a_kq * 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a_reg[i][j] = 0.f;
      }
      if (VEC == 4) {
        if (kf < p.K) {
          int tap = kf / p.Cs, c = kf - tap * p.Cs;
          int th = tap / p.ntw, tw = tap - th * p.ntw;
          int oh = p.off_h[th], ow = p.off_w[tw];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            int sh = a_h[i] + oh, sw = a_w[i] + ow;
            if (a_ok[i] && sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
              size_t off = (((size_t)a_n[i] * p.Hs + (sh >> p.up2x)) * p.Ws + (sw >> p.up2x)) * p.Cs + c;
              float4 v = __ldg(reinterpret_cast<const float4*>(p.src + off));
              a_reg[i][0] = v.x; a_reg[i][1] = v.y; a_reg[i][2] = v.z; a_reg[i][3] = v.w;
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = kf + j;
          if (k < p.K) {
            int tap = k / p.Cs, c = k - tap * p.Cs;
            int th = tap / p.ntw, tw = tap - th * p.ntw;
            int oh = p.off_h[th], ow = p.off_w[tw];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              int sh = a_h[i] + oh, sw = a_w[i] + ow;
              if (a_ok[i] && sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
                size_t off = (((size_t)a_n[i] * p.Hs + (sh >> p.up2x)) * p.Ws + (sw >> p.up2x)) * p.Cs + c;
                a_reg[i][j] = __ldg(p.src + off);
              }
            }
          }
        }
      }
    }
    // B
    {
      int k = k0 + b_k;
#pragma unroll
      for (int j = 0; j < 4; ++j) b_reg[j] = 0.f;
      if (k < p.K) {
        int tap = k / p.Cs, c = k - tap * p.Cs;
        size_t row = (size_t)p.tapw[tap] * p.Cs + c;
        const float* wp = p.wmat + row * p.Cd + n0 + b_c;
        if (b_vec && n0 + b_c + 3 < p.Cd) {
          float4 v = __ldg(reinterpret_cast<const float4*>(wp));
          b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + b_c + j < p.Cd) b_reg[j] = __ldg(wp + j);
        }
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) As[a_kq * 4 + j][a_row + i * 64] = a_reg[i][j];
    *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = ceil_div(p.K, BK);
  load_tiles(0);
  for (int kt = 0; kt < nk; ++kt) {
    store_tiles();
    __syncthreads();
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + tx * 4 + j < p.Cd) bias[j] = __ldg(p.bias + n0 + tx * 4 + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    int rw = (int)(m % p.Wr);
    long long q = m / p.Wr;
    int rh = (int)(q % p.Hr);
    int n = (int)(q / p.Hr);
    size_t pix = ((size_t)n * p.Hd + (rh * p.dsh + p.doh)) * p.Wd + (rw * p.dsw + p.dow);
    float* dp = p.dst + pix * p.Cd + n0 + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(acc[i][j] + bias[j], p.act);
    if (b_vec && n0 + tx * 4 + 3 < p.Cd) {
      *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j < p.Cd) dp[j] = v[j];
    }
  }
}

int launch_igemm_ffma(const IGemmParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div_ll(p.M, BM), (unsigned)ceil_div(p.Cd, BN));
  if ((p.Cs & 3) == 0)
    igemm_kernel<4><<<grid, NT, 0, st>>>(p);
  else
    igemm_kernel<1><<<grid, NT, 0, st>>>(p);
  return check_launch("igemm_kernel");
}

// ------------------------------------------------------------------------------------------
// weight gradient:  ws[z][kf, co] = sum_{p in split z} X[pix(p, tap(kf)), c(kf)] * dY[p, co]
// ------------------------------------------------------------------------------------------
struct WgradParams {
  const float* x;
  const float* dy;
  float* ws;  // [splits][K][Cout]
  int N, H, W, Cin, up2x;
  int Ho, Wo, Cout, KH, KW, stride, pad;
  long long P;  // N*Ho*Wo
  int K;        // KH*KW*Cin
  long long chunk;
};

template <int VEC>
__global__ void __launch_bounds__(NT, 2) wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float As[BK][AS_LD];  // [pixel][kflat]
  __shared__ __align__(16) float Bs[BK][BN];     // [pixel][co]
  const int t = threadIdx.x;
  const int kf0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long p_begin = (long long)blockIdx.z * p.chunk;
  long long p_end = p_begin + p.chunk;
  if (p_end > p.P) p_end = p.P;

  // A loader: 32 k-quads x 8 pixel rows (two per thread)
  const int a_kq = t & 31;
  const int a_pp = t >> 5;  // 0..7 (+8)
  int a_off_h[4], a_off_w[4], a_c[4];
  bool a_kok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int kf = kf0 + a_kq * 4 + j;
    a_kok[j] = kf < p.K;
    int kk = a_kok[j] ? kf : 0;
    int tap = kk / p.Cin;
    a_c[j] = kk - tap * p.Cin;
    int kh = tap / p.KW, kw = tap - kh * p.KW;
    a_off_h[j] = kh - p.pad;
    a_off_w[j] = kw - p.pad;
  }
  const int HL = p.H << p.up2x, WL = p.W << p.up2x;
  const int b_pp = t >> 4;
  const int b_c = (t & 15) * 4;
  const bool b_vec = (p.Cout & 3) == 0;

  float a_reg[2][4], b_reg[4];
  auto load_tiles = [&](long long p0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a_reg[i][j] = 0.f;
      long long pp = p0 + a_pp + i * 8;
      if (pp < p_end) {
        int wo = (int)(pp % p.Wo);
        long long q = pp / p.Wo;
        int ho = (int)(q % p.Ho);
        int n = (int)(q / p.Ho);
        if (VEC == 4) {
          if (a_kok[0]) {
            int sh = ho * p.stride + a_off_h[0], sw = wo * p.stride + a_off_w[0];
            if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
              size_t off = (((size_t)n * p.H + (sh >> p.up2x)) * p.W + (sw >> p.up2x)) * p.Cin + a_c[0];
              float4 v = __ldg(reinterpret_cast<const float4*>(p.x + off));
              a_reg[i][0] = v.x; a_reg[i][1] = v.y; a_reg[i][2] = v.z; a_reg[i][3] = v.w;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!a_kok[j]) continue;
            int sh = ho * p.stride + a_off_h[j], sw = wo * p.stride + a_off_w[j];
            if (sh >= 0 && sh < HL && sw >= 0 && sw < WL) {
              size_t off = (((size_t)n * p.H + (sh >> p.up2x)) * p.W + (sw >> p.up2x)) * p.Cin + a_c[j];
              a_reg[i][j] = __ldg(p.x + off);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) b_reg[j] = 0.f;
    long long pp = p0 + b_pp;
    if (pp < p_end) {
      const float* dp = p.dy + (size_t)pp * p.Cout + n0 + b_c;
      if (b_vec && n0 + b_c + 3 < p.Cout) {
        float4 v = __ldg(reinterpret_cast<const float4*>(dp));
        b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n0 + b_c + j < p.Cout) b_reg[j] = __ldg(dp + j);
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      *reinterpret_cast<float4*>(&As[a_pp + i * 8][a_kq * 4]) =
          make_float4(a_reg[i][0], a_reg[i][1], a_reg[i][2], a_reg[i][3]);
    *reinterpret_cast<float4*>(&Bs[b_pp][b_c]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (p_begin < p_end) {
    load_tiles(p_begin);
    for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
      store_tiles();
      __syncthreads();
      if (p0 + BK < p_end) load_tiles(p0 + BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* wsz = p.ws + (size_t)blockIdx.z * p.K * p.Cout;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int kf = kf0 + ty * 8 + i;
    if (kf >= p.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co < p.Cout) wsz[(size_t)kf * p.Cout + co] = acc[i][j];
    }
  }
}

// dw_oihw[co][ci][tap] = sum_z ws[z][tap*CinP + ci][co]      (deterministic split reduction + layout change)
// Block = 32 output channels x 16 input channels x all taps: reads are coalesced along co (128 B per kf row),
// the tile is transposed through shared memory and written as contiguous (ci, tap) runs per co.
constexpr int RCO = 32, RCI = 16;
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits,
                                                           int K, int Cout, int Cin, int CinP, int KHW) {
  extern __shared__ float tile[];   // [RCO][RCI*KHW + 1]
  const int ld = RCI * KHW + 1;
  const int ci0 = blockIdx.x * RCI, co0 = blockIdx.y * RCO;
  const size_t total = (size_t)K * Cout;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;   // lane -> co, 8 warps -> (tap, ci) rows
  for (int r = wrp; r < RCI * KHW; r += 8) {
    const int tap = r / RCI, cil = r - tap * RCI;
    const int ci = ci0 + cil, co = co0 + lane;
    float s = 0.f;
    if (ci < Cin && co < Cout) {
      const size_t idx = (size_t)(tap * CinP + ci) * Cout + co;
      for (int z = 0; z < splits; ++z) s += ws[(size_t)z * total + idx];
    }
    tile[lane * ld + cil * KHW + tap] = s;
  }
  __syncthreads();
  const int nci = min(RCI, Cin - ci0);
  const int run = nci * KHW;   // contiguous floats per co in the OIHW tensor
  for (int c = wrp; c < RCO; c += 8) {
    const int co = co0 + c;
    if (co >= Cout) break;
    float* out = dw + ((size_t)co * Cin + ci0) * KHW;
    for (int i = lane; i < run; i += 32) out[i] = tile[c * ld + i];
  }
}

__global__ void colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long M, int C) {
  // one block per channel group of 32; deterministic tree over rows
  int c = blockIdx.x * 32 + (threadIdx.x & 31);
  int r0 = threadIdx.x >> 5;  // 8 row lanes
  float s = 0.f;
  if (c < C)
    for (long long r = r0; r < M; r += 8) s += x[(size_t)r * C + c];
  __shared__ float sm[8][33];
  sm[r0][threadIdx.x & 31] = s;
  __syncthreads();
  if (r0 == 0 && c < C) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += sm[i][threadIdx.x & 31];
    out[c] = tot;
  }
}

__global__ void sumpool2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W,
                                  int C) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)N * H * W * C;
  if (idx >= total) return;
  int c = (int)(idx % C);
  size_t q = idx / C;
  int w = (int)(q % W);
  q /= W;
  int h = (int)(q % H);
  int n = (int)(q / H);
  const size_t rs = (size_t)2 * W * C;
  const float* s = src + (((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C + c;
  dst[idx] = (s[0] + s[C]) + (s[rs] + s[rs + C]);
}

__global__ void pack_fwd_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW) {
  // out[(tap*Cin+ci)*Cout + co] = w[(co*Cin+ci)*KHW + tap]
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)Cout * Cin * KHW;
  if (idx >= total) return;
  int co = (int)(idx % Cout);
  size_t q = idx / Cout;
  int ci = (int)(q % Cin);
  int tap = (int)(q / Cin);
  out[idx] = w[((size_t)co * Cin + ci) * KHW + tap];
}
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW) {
  // out[(tap*Cout+co)*Cin + ci] = w[(co*Cin+ci)*KHW + tap]
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)Cout * Cin * KHW;
  if (idx >= total) return;
  int ci = (int)(idx % Cin);
  size_t q = idx / Cin;
  int co = (int)(q % Cout);
  int tap = (int)(q / Cout);
  out[idx] = w[((size_t)co * Cin + ci) * KHW + tap];
}

int wgrad_splits(const MogConvDesc& d, int Ho, int Wo) {
  long long P = (long long)d.N * Ho * Wo;
  int K = d.KH * d.KW * d.Cin;
  long long tiles = (long long)ceil_div(K, BM) * ceil_div(d.Cout, BN);
  long long want = ceil_div_ll(4 * kNumSMs, tiles);
  long long maxs = ceil_div_ll(P, 4 * BK);
  long long s = want < maxs ? want : maxs;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

int launch_wgrad_ffma_partial(const MogConvDesc& d, int Ho, int Wo, const float* x, const float* dy, float* ws,
                              int* splits_out, cudaStream_t st) {
  WgradParams p;
  p.x = x; p.dy = dy; p.ws = ws;
  p.N = d.N; p.H = d.H; p.W = d.W; p.Cin = d.Cin; p.up2x = d.up2x;
  p.Ho = Ho; p.Wo = Wo; p.Cout = d.Cout; p.KH = d.KH; p.KW = d.KW; p.stride = d.stride; p.pad = d.pad;
  p.P = (long long)d.N * Ho * Wo;
  p.K = d.KH * d.KW * d.Cin;
  int splits = wgrad_splits(d, Ho, Wo);
  long long chunk = ceil_div_ll(p.P, splits);
  chunk = ceil_div_ll(chunk, BK) * BK;
  splits = (int)ceil_div_ll(p.P, chunk);
  p.chunk = chunk;
  dim3 grid(ceil_div(p.K, BM), ceil_div(d.Cout, BN), splits);
  if ((d.Cin & 3) == 0)
    wgrad_kernel<4><<<grid, NT, 0, st>>>(p);
  else
    wgrad_kernel<1><<<grid, NT, 0, st>>>(p);
  *splits_out = splits;
  return check_launch("wgrad_kernel");
}

size_t wgrad_ffma_workspace_bytes(const MogConvDesc& d, int Ho, int Wo) {
  return (size_t)wgrad_splits(d, Ho, Wo) * d.KH * d.KW * d.Cin * d.Cout * sizeof(float);
}

int launch_wgrad_reduce(const float* ws, float* dw, int splits, int K, int Cout, int Cin, int CinP, int KHW,
                        cudaStream_t st) {
  dim3 grid(ceil_div(Cin, RCI), ceil_div(Cout, RCO));
  const size_t smem = sizeof(float) * RCO * (RCI * KHW + 1);
  if (smem > 48 * 1024) {   // 5x5 filters: 51 KB (found by tests/test_gpu_fullsize.py: the launch failed with "invalid argument")
    cudaError_t e = cudaFuncSetAttribute(wgrad_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MOG_ERR_UNSUPPORTED, "wgrad_reduce_kernel: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
  }
  wgrad_reduce_kernel<<<grid, 256, smem, st>>>(ws, dw, splits, K, Cout, Cin, CinP, KHW);
  return check_launch("wgrad_reduce_kernel");
}

}  // namespace mog

using namespace mog;

extern "C" int mog_sumpool2x2(const float* src, float* dst, int N, int H, int W, int C, void* stream) {
  MOG_REQUIRE(src && dst && N > 0 && H > 0 && W > 0 && C > 0, "mog_sumpool2x2: bad argument");
  size_t total = (size_t)N * H * W * C;
  sumpool2x2_kernel<<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, as_stream(stream)>>>(src, dst, N, H, W, C);
  return check_launch("sumpool2x2_kernel");
}

namespace mog {
int launch_colsum(const float* x, float* out, long long M, int C, cudaStream_t st) {
  colsum_kernel<<<ceil_div(C, 32), 256, 0, st>>>(x, out, M, C);
  return check_launch("colsum_kernel");
}
int launch_sumpool(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st) {
  size_t total = (size_t)N * H * W * C;
  sumpool2x2_kernel<<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, st>>>(src, dst, N, H, W, C);
  return check_launch("sumpool2x2_kernel");
}
}  // namespace mog
