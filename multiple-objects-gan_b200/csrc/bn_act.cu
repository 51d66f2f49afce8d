// bn_act.cu -- train-mode BatchNorm statistics, fused affine + activation (GLU / ReLU /
// LeakyReLU / none, optional residual) forward, and the two-pass backward.  HBM-bound streaming
// kernels over [S*M][C] row-major (NHWC) tensors: per-channel reductions run over rows with
// 32 consecutive channels per warp row (128-byte coalesced), partials are combined with
// warp/smem reductions into one fp64 partial per (block, channel); a second kernel adds the partials in a fixed order.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace mog {

constexpr int RED_ROWS = 8;      // row lanes per block (blockDim = 32 x 8)
constexpr int ROWS_PER_BLOCK = 256;
// rows handled per block: 256, grown only if the grid's y dimension would overflow
inline int rows_per_block(int M) {
  int r = ROWS_PER_BLOCK;
  while (ceil_div(M, r) > 65535) r *= 2;
  return r;
}

// ---- statistics -------------------------------------------------------------------------
// Partial sums: every reducing kernel leaves ONE partial per (block row p, quantity k, segment, channel) in
// part[p][k][S][C] (fp64); the consumer (bn_finalize / bn_bwd_sum) adds the P partials in a fixed order, so the
// statistics are bit-reproducible run to run (the first version ended in fp64 atomics, which are not) and no atomics
// serialise in L2.
__device__ __forceinline__ double* part_slot(double* part, int p, int k, int S, int C, int s, int c) {
  return part + (((size_t)p * 2 + k) * S + s) * C + c;
}

__global__ void bn_stats_kernel(const float* __restrict__ x, int M, int C, int rpb, double* __restrict__ part, int S) {
  // grid: (ceil(C/32), P row-chunk groups, S); block p takes the row chunks p, p + P, ...
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int s = blockIdx.z;
  // Shifted sums: the fp32 partials accumulate d = x - x0 (x0 = the channel's value in the first row of the segment), so
  // that sum(x^2)/M - mean^2 does not cancel for channels whose mean is large against their spread (measured: mean/std = 300
  // lost 3 digits of the variance with plain fp32 partials; torch uses Welford).  The shift is folded back exactly in double.
  double da = 0.0, db = 0.0;
  float x0 = 0.f;
  if (c < C) {
    const float* xp = x + ((size_t)s * M) * C + c;
    x0 = __ldg(xp);
    for (long long r_begin = (long long)blockIdx.y * rpb; r_begin < M; r_begin += (long long)gridDim.y * rpb) {
      long long r_end = r_begin + rpb;
      if (r_end > M) r_end = M;
      float a = 0.f, b = 0.f;
      int n = 0;
      for (long long r = r_begin + rl; r < r_end; r += RED_ROWS) {
        float v = __ldg(xp + (size_t)r * C) - x0;
        a += v;
        b = fmaf(v, v, b);
        ++n;
      }
      const double s0 = (double)x0;
      da += (double)a + (double)n * s0;
      db += (double)b + 2.0 * s0 * (double)a + (double)n * s0 * s0;
    }
  }
  __shared__ double sa[RED_ROWS][33], sb[RED_ROWS][33];
  sa[rl][lane] = da;
  sb[rl][lane] = db;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < RED_ROWS; ++i) {
      ta += sa[i][lane];
      tb += sb[i][lane];
    }
    *part_slot(part, blockIdx.y, 0, S, C, s, c) = ta;
    *part_slot(part, blockIdx.y, 1, S, C, s, c) = tb;
  }
}

// One warp per channel: the lanes add the P partials (lane-strided, then a fixed xor-shuffle tree: the order never
// changes, so the result is bit-reproducible), lane 0 finalises.  (One THREAD per channel looping over up to ~600 partials
// took 30 us per layer.)
__device__ __forceinline__ void part_sum2(const double* __restrict__ part, int P, int S, int C, int s, int c, int lane, double* o0,
                                          double* o1) {
  double a = 0.0, b = 0.0;
  for (int p = lane; p < P; p += 32) {
    a += part[(((size_t)p * 2 + 0) * S + s) * C + c];
    b += part[(((size_t)p * 2 + 1) * S + s) * C + c];
  }
  *o0 = warp_sum_d(a);
  *o1 = warp_sum_d(b);
}

__global__ void bn_finalize_kernel(const double* __restrict__ part, int P, int S, int M,
                                   int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean,
                                   float* invstd, float* scale, float* shift) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f;
  float rv = running_var ? running_var[c] : 0.f;
  for (int s = 0; s < S; ++s) {
    double sum, sqsum;
    part_sum2(part, P, S, C, s, c, lane, &sum, &sqsum);
    double mu = sum / (double)M;
    double var = sqsum / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    float is = (float)(1.0 / sqrt(var + (double)eps));
    float muf = (float)mu;
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * is;
    if (lane == 0) {
      mean[(size_t)s * C + c] = muf;
      invstd[(size_t)s * C + c] = is;
      scale[(size_t)s * C + c] = sc;
      shift[(size_t)s * C + c] = b - muf * sc;
    }
    // nn.BatchNorm: running = (1-m)*running + m*stat, unbiased variance for running_var
    double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    rm = (1.f - momentum) * rm + momentum * muf;
    rv = (1.f - momentum) * rv + momentum * (float)unb;
  }
  if (lane == 0) {
    if (running_mean) running_mean[c] = rm;
    if (running_var) running_var[c] = rv;
  }
}

// ---- forward ------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float z, int act) {
  switch (act) {
    case MOG_ACT_RELU: return fmaxf(z, 0.f);
    case MOG_ACT_LRELU: return z > 0.f ? z : 0.2f * z;
    case MOG_ACT_TANH: return tanhf(z);
    case MOG_ACT_SIGMOID: return sigmoidf_(z);
    default: return z;
  }
}

// 4 output channels per thread (C_out % 4 == 0 fast path), else scalar
template <int VEC>
__global__ void affine_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      const float* __restrict__ shift, const float* __restrict__ res,
                                      float* __restrict__ y, int M, int C, int act, size_t total_out,
                                      __nv_bfloat16* __restrict__ phi, __nv_bfloat16* __restrict__ plo, int CP) {
  const int Co = act == MOG_ACT_GLU ? C / 2 : C;
  size_t idx = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (idx >= total_out) return;
  const size_t row = idx / Co;
  const int c = (int)(idx - row * Co);
  const int s = (int)(row / M);
  const float* xr = x + row * C;
  const float* sc = scale ? scale + (size_t)s * C : nullptr;
  const float* sh = shift ? shift + (size_t)s * C : nullptr;
  float out[VEC];
  if (VEC == 4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(xr + c));
    float av[4] = {a.x, a.y, a.z, a.w};
    if (sc) {
      float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c)), h4 = __ldg(reinterpret_cast<const float4*>(sh + c));
      av[0] = fmaf(av[0], s4.x, h4.x); av[1] = fmaf(av[1], s4.y, h4.y);
      av[2] = fmaf(av[2], s4.z, h4.z); av[3] = fmaf(av[3], s4.w, h4.w);
    }
    if (act == MOG_ACT_GLU) {
      float4 g = __ldg(reinterpret_cast<const float4*>(xr + Co + c));
      float gv[4] = {g.x, g.y, g.z, g.w};
      if (sc) {
        float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + Co + c)), h4 = __ldg(reinterpret_cast<const float4*>(sh + Co + c));
        gv[0] = fmaf(gv[0], s4.x, h4.x); gv[1] = fmaf(gv[1], s4.y, h4.y);
        gv[2] = fmaf(gv[2], s4.z, h4.z); gv[3] = fmaf(gv[3], s4.w, h4.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = av[j] * sigmoidf_(gv[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) out[j] = act_fwd(av[j], act);
    }
    if (res) {
      float4 r = __ldg(reinterpret_cast<const float4*>(res + idx));
      out[0] += r.x; out[1] += r.y; out[2] += r.z; out[3] += r.w;
    }
    *reinterpret_cast<float4*>(y + idx) = make_float4(out[0], out[1], out[2], out[3]);
    if (phi) {
      // bf16 hi (+ lo = bf16(v - hi)) planes [rows][CP] for the tensor-core convolutions; pad channels zero
      __nv_bfloat162 h01 = __floats2bfloat162_rn(out[0], out[1]), h23 = __floats2bfloat162_rn(out[2], out[3]);
      const size_t o = row * CP + c;
      *reinterpret_cast<uint2*>(phi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
      if (plo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        __nv_bfloat162 l01 = __floats2bfloat162_rn(out[0] - f01.x, out[1] - f01.y), l23 = __floats2bfloat162_rn(out[2] - f23.x, out[3] - f23.y);
        *reinterpret_cast<uint2*>(plo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
      }
      if (c + 4 == Co && CP > Co) {
        *reinterpret_cast<uint2*>(phi + o + 4) = make_uint2(0u, 0u);
        if (plo) *reinterpret_cast<uint2*>(plo + o + 4) = make_uint2(0u, 0u);
      }
    }
  } else {
    float a = xr[c];
    if (sc) a = fmaf(a, sc[c], sh[c]);
    float o;
    if (act == MOG_ACT_GLU) {
      float g = xr[Co + c];
      if (sc) g = fmaf(g, sc[Co + c], sh[Co + c]);
      o = a * sigmoidf_(g);
    } else {
      o = act_fwd(a, act);
    }
    if (res) o += res[idx];
    y[idx] = o;
  }
}

// ---- backward -----------------------------------------------------------------------------
struct BnBwdArgs {
  const float* x; const float* dy; const float* mean; const float* invstd; const float* gamma; const float* beta;
  int S, M, C, act, rows_per_block;
};

// Per-thread channel constants: affine of the own channel and (GLU) of the partner channel.
struct ChanConst {
  float sc, sh, mu, is;   // own channel: z = x*sc + sh, xhat = (x-mu)*is
  float psc, psh;         // partner channel (GLU only)
  int partner;            // channel index of the partner, or -1
  int dyc;                // dy channel feeding this input channel
  bool gate;              // GLU: this channel is the gate half
};

__device__ __forceinline__ ChanConst chan_const(const BnBwdArgs& a, int s, int c) {
  ChanConst k;
  const int Co = a.act == MOG_ACT_GLU ? a.C / 2 : a.C;
  auto aff = [&](int ch, float& sc, float& sh, float& mu, float& is) {
    if (!a.mean) { sc = 1.f; sh = 0.f; mu = 0.f; is = 1.f; return; }  // plain activation, no BN
    float g = a.gamma ? a.gamma[ch] : 1.f, b = a.beta ? a.beta[ch] : 0.f;
    mu = a.mean[(size_t)s * a.C + ch];
    is = a.invstd[(size_t)s * a.C + ch];
    sc = g * is;
    sh = b - mu * sc;
  };
  aff(c, k.sc, k.sh, k.mu, k.is);
  k.partner = -1; k.psc = 1.f; k.psh = 0.f; k.gate = false; k.dyc = c;
  if (a.act == MOG_ACT_GLU) {
    k.gate = c >= Co;
    k.partner = k.gate ? c - Co : c + Co;
    k.dyc = k.gate ? c - Co : c;
    float m_, i_;
    aff(k.partner, k.psc, k.psh, m_, i_);
  }
  return k;
}

// dz for this thread's channel on one row (xr: x row, dyr: dy row)
__device__ __forceinline__ float dz_of(const ChanConst& k, float xv, const float* __restrict__ xr,
                                       const float* __restrict__ dyr, int act) {
  const float z = fmaf(xv, k.sc, k.sh);
  const float g = __ldg(dyr + k.dyc);
  if (act == MOG_ACT_GLU) {
    const float zp = fmaf(__ldg(xr + k.partner), k.psc, k.psh);
    if (!k.gate) return g * sigmoidf_(zp);      // d/d(value half): sigmoid(gate)
    const float sg = sigmoidf_(z);
    return g * zp * sg * (1.f - sg);            // d/d(gate half): value * sigmoid'(gate)
  }
  switch (act) {
    case MOG_ACT_RELU: return z > 0.f ? g : 0.f;
    case MOG_ACT_LRELU: return z > 0.f ? g : 0.2f * g;
    case MOG_ACT_TANH: { float t = tanhf(z); return g * (1.f - t * t); }
    case MOG_ACT_SIGMOID: { float t = sigmoidf_(z); return g * t * (1.f - t); }
    default: return g;
  }
}

// part[p][0] = sum dz (d beta), part[p][1] = sum dz * xhat (d gamma)
__global__ void bn_bwd_reduce_kernel(BnBwdArgs a, double* __restrict__ part) {
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int s = blockIdx.z;
  const int Co = a.act == MOG_ACT_GLU ? a.C / 2 : a.C;
  double d1 = 0.0, d2 = 0.0;
  if (c < a.C) {
    const ChanConst k = chan_const(a, s, c);
    for (long long r_begin = (long long)blockIdx.y * a.rows_per_block; r_begin < a.M; r_begin += (long long)gridDim.y * a.rows_per_block) {
      long long r_end = r_begin + a.rows_per_block;
      if (r_end > a.M) r_end = a.M;
      float g1 = 0.f, g2 = 0.f;
      for (long long r = r_begin + rl; r < r_end; r += RED_ROWS) {
        size_t row = (size_t)s * a.M + r;
        const float* xr = a.x + row * a.C;
        float xv = __ldg(xr + c);
        float dz = dz_of(k, xv, xr, a.dy + row * Co, a.act);
        g1 += dz;
        g2 = fmaf(dz, (xv - k.mu) * k.is, g2);
      }
      d1 += (double)g1;
      d2 += (double)g2;
    }
  }
  __shared__ double sa[RED_ROWS][33], sb[RED_ROWS][33];
  sa[rl][lane] = d1;
  sb[rl][lane] = d2;
  __syncthreads();
  if (rl == 0 && c < a.C) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < RED_ROWS; ++i) {
      t1 += sa[i][lane];
      t2 += sb[i][lane];
    }
    *part_slot(part, blockIdx.y, 0, a.S, a.C, s, c) = t1;
    *part_slot(part, blockIdx.y, 1, a.S, a.C, s, c) = t2;
  }
}

// adds the P partials of the backward sums in a fixed order -> dbeta_seg / dgamma_seg [S][C] (what the apply pass reads),
// and the parameter gradients (sum over segments) in the same launch
__global__ void bn_bwd_sum_kernel(const double* __restrict__ part, int P, int S, int C, double* __restrict__ dgamma_seg,
                                  double* __restrict__ dbeta_seg, float* dgamma, float* dbeta) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double gt = 0.0, bt = 0.0;
  for (int s = 0; s < S; ++s) {
    double b, g;
    part_sum2(part, P, S, C, s, c, lane, &b, &g);
    if (lane == 0) {
      dbeta_seg[(size_t)s * C + c] = b;
      dgamma_seg[(size_t)s * C + c] = g;
    }
    gt += g;
    bt += b;
  }
  if (lane == 0) {
    if (dgamma) dgamma[c] = (float)gt;
    if (dbeta) dbeta[c] = (float)bt;
  }
}

__global__ void bn_bwd_apply_kernel(BnBwdArgs a, const double* __restrict__ dgamma, const double* __restrict__ dbeta,
                                    float* __restrict__ dx) {
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int s = blockIdx.z;
  if (c >= a.C) return;
  const int Co = a.act == MOG_ACT_GLU ? a.C / 2 : a.C;
  const long long r_begin = (long long)blockIdx.y * a.rows_per_block;
  long long r_end = r_begin + a.rows_per_block;
  if (r_end > a.M) r_end = a.M;
  const ChanConst k = chan_const(a, s, c);
  const float k1 = dbeta ? (float)(dbeta[(size_t)s * a.C + c] / (double)a.M) : 0.f;
  const float k2 = dgamma ? (float)(dgamma[(size_t)s * a.C + c] / (double)a.M) : 0.f;
  for (long long r = r_begin + rl; r < r_end; r += RED_ROWS) {
    size_t row = (size_t)s * a.M + r;
    const float* xr = a.x + row * a.C;
    float xv = __ldg(xr + c);
    float dz = dz_of(k, xv, xr, a.dy + row * Co, a.act);
    float xh = (xv - k.mu) * k.is;
    dx[row * a.C + c] = k.sc * (dz - k1 - xh * k2);
  }
}

__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                               size_t n, int act) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = dy[i], v = y[i], o;
  switch (act) {
    case MOG_ACT_RELU: o = v > 0.f ? g : 0.f; break;
    case MOG_ACT_LRELU: o = v > 0.f ? g : 0.2f * g; break;
    case MOG_ACT_TANH: o = g * (1.f - v * v); break;
    case MOG_ACT_SIGMOID: o = g * v * (1.f - v); break;
    default: o = g;
  }
  dz[i] = o;
}


// ---------------------------------------------------------------------------------------------
// Vectorised variants (channel width % 4 == 0): one thread owns 4 consecutive channels (for GLU: 4 value
// channels AND their 4 gate partners), a block is GB column groups x R row lanes, every load / store is a
// 16-byte access and consecutive threads cover a contiguous row segment; 4 rows are in flight per thread.
// The scalar kernels above remain for odd channel counts.
// ---------------------------------------------------------------------------------------------
struct V4Geom { int Cv, GB, R, nxb, rpb, nyb; };
// reduce = true: the reducing kernels (statistics, backward sums) leave one fp64 partial per channel per block, so they run a
// fixed, small number of blocks (a few per SM over all column groups and segments), each looping over its row chunks with the
// sums kept in registers.  (History: the first version ended in fp64 atomics, which serialise in L2 at ~55 ns each -- 6554
// blocks -> 320 us for a pass that streams in 125 us -- and are not bit-reproducible.)
// which (reducing kernels only): 0 = forward statistics, 1 = backward reduce.  The grid of a reducing kernel is capped at the
// number of blocks that are RESIDENT at once (each block strides over the rows and leaves one partial): bn_stats_v4_kernel
// holds 3 blocks per SM (80 registers), and a grid of 4 per SM ran as one full wave plus a one-third wave -- 4.3 TB/s instead
// of the ~6 TB/s of the two-wave-exact backward reduce (ncu r2j).
static V4Geom v4_geom(int width, int M, int S = 1, bool reduce = false, int which = 1) {
  V4Geom g;
  g.Cv = width / 4;
  g.nxb = ceil_div(g.Cv, 32);
  g.GB = ceil_div(g.Cv, g.nxb);
  g.R = 256 / g.GB;
  if (g.R > 32) g.R = 32;
  g.rpb = g.R * 32;
  while (ceil_div(M, g.rpb) > 65535) g.rpb *= 2;
  g.nyb = ceil_div(M, g.rpb);
  if (reduce) {
    static int per_sm[2] = {-1, -1};
    if (per_sm[0] < 0) {
      const char* e0 = getenv("MOG_BN_STATS_BLOCKS_PER_SM");   // tuning knobs: reducing blocks per SM
      const char* e1 = getenv("MOG_BN_BLOCKS_PER_SM");
      per_sm[0] = e0 ? atoi(e0) : 3;
      per_sm[1] = e1 ? atoi(e1) : 4;
      if (per_sm[0] < 1) per_sm[0] = 1;
      if (per_sm[1] < 1) per_sm[1] = 1;
    }
    int cap = (per_sm[which ? 1 : 0] * kNumSMs) / (g.nxb * (S > 0 ? S : 1));
    if (cap < 1) cap = 1;
    if (g.nyb > cap) g.nyb = cap;
  }
  return g;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void to_arr(const float4& v, float (&a)[4]) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }

// block-level sum over the R row lanes of NV per-thread values (4 channels each), then one partial store per channel
template <int NV>
__device__ __forceinline__ void v4_block_reduce(double (&v)[NV][4], int GB, int R, int cgl, int rl, bool valid, double* const (&dst)[NV]) {
  __shared__ double red[NV][256 * 4];
  const int t = rl * GB + cgl;
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[k][t * 4 + j] = v[k][j];
  __syncthreads();
  if (rl == 0 && valid) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double acc = 0.0;
        for (int i = 0; i < R; ++i) acc += red[k][(i * GB + cgl) * 4 + j];
        dst[k][j] = acc;
      }
  }
}

__global__ void __launch_bounds__(256, 3) bn_stats_v4_kernel(const float* __restrict__ x, int M, int C, V4Geom g, double* __restrict__ part, int S) {
  const int cgl = threadIdx.x % g.GB, rl = threadIdx.x / g.GB;
  const int cg = blockIdx.x * g.GB + cgl;
  const bool valid = cg < g.Cv;
  const int c = cg * 4, s = blockIdx.z;
  double v[2][4] = {{0., 0., 0., 0.}, {0., 0., 0., 0.}};
  if (valid) {
    const float* xp = x + ((size_t)s * M) * C + c;
    // shifted sums (see bn_stats_kernel): fp32 partials of d = x - x0, folded back exactly in double
    float x0[4];
    to_arr(ld4(xp), x0);
    for (long long r_begin = (long long)blockIdx.y * g.rpb; r_begin < M; r_begin += (long long)gridDim.y * g.rpb) {
      long long r_end = r_begin + g.rpb;
      if (r_end > M) r_end = M;
      float f[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // fp32 over <= rpb / R rows, then folded into the doubles
      int n = 0;
#pragma unroll 8
      for (long long r = r_begin + rl; r < r_end; r += g.R) {
        float a[4];
        to_arr(ld4(xp + (size_t)r * C), a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float dd = a[j] - x0[j]; f[0][j] += dd; f[1][j] = fmaf(dd, dd, f[1][j]); }
        ++n;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double s0 = (double)x0[j], fa = (double)f[0][j];
        v[0][j] += fa + (double)n * s0;
        v[1][j] += (double)f[1][j] + 2.0 * s0 * fa + (double)n * s0 * s0;
      }
    }
  }
  double* const dst[2] = {part_slot(part, blockIdx.y, 0, S, C, s, c), part_slot(part, blockIdx.y, 1, S, C, s, c)};
  v4_block_reduce<2>(v, g.GB, g.R, cgl, rl, valid, dst);
}

// per-thread constants of 4 channels
struct Aff4 { float sc[4], sh[4], mu[4], is[4]; };
__device__ __forceinline__ void load_aff4(const BnBwdArgs& a, int s, int c, Aff4& k) {
  if (!a.mean) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { k.sc[j] = 1.f; k.sh[j] = 0.f; k.mu[j] = 0.f; k.is[j] = 1.f; }
    return;
  }
  float g[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.gamma) to_arr(ld4(a.gamma + c), g);
  if (a.beta) to_arr(ld4(a.beta + c), b);
  to_arr(ld4(a.mean + (size_t)s * a.C + c), k.mu);
  to_arr(ld4(a.invstd + (size_t)s * a.C + c), k.is);
#pragma unroll
  for (int j = 0; j < 4; ++j) { k.sc[j] = g[j] * k.is[j]; k.sh[j] = b[j] - k.mu[j] * k.sc[j]; }
}

template <int ACT>
__device__ __forceinline__ float dact(float z, float g) {
  if (ACT == MOG_ACT_RELU) return z > 0.f ? g : 0.f;
  if (ACT == MOG_ACT_LRELU) return z > 0.f ? g : 0.2f * g;
  if (ACT == MOG_ACT_TANH) { const float t = tanhf(z); return g * (1.f - t * t); }
  if (ACT == MOG_ACT_SIGMOID) { const float t = sigmoidf_(z); return g * t * (1.f - t); }
  return g;
}

// dz of the 4 value channels (and, GLU, of the 4 gate channels) of one row
template <int ACT>
__device__ __forceinline__ void dz_row(const Aff4& kv, const Aff4& kg, const float (&xv)[4], const float (&xg)[4], const float (&gy)[4],
                                       float (&dzv)[4], float (&dzg)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float zv = fmaf(xv[j], kv.sc[j], kv.sh[j]);
    if (ACT == MOG_ACT_GLU) {
      const float zg = fmaf(xg[j], kg.sc[j], kg.sh[j]);
      const float sg = sigmoidf_(zg);
      dzv[j] = gy[j] * sg;
      dzg[j] = gy[j] * zv * sg * (1.f - sg);
    } else {
      dzv[j] = dact<ACT>(zv, gy[j]);
      dzg[j] = 0.f;
    }
  }
}

// (ncu: 160 registers per thread left ONE 240-thread block per SM, 12 % of the warps active; bounded to 2 blocks per SM)
template <int ACT>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_v4_kernel(BnBwdArgs a, V4Geom g, double* __restrict__ part) {
  constexpr bool GLU = ACT == MOG_ACT_GLU;
  const int Co = GLU ? a.C / 2 : a.C;
  const int cgl = threadIdx.x % g.GB, rl = threadIdx.x / g.GB;
  const int cg = blockIdx.x * g.GB + cgl;
  const bool valid = cg < g.Cv;
  const int c = cg * 4, s = blockIdx.z;
  constexpr int NV = GLU ? 4 : 2;
  double v[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[k][j] = 0.;
  if (valid) {
    Aff4 kv, kg;
    load_aff4(a, s, c, kv);
    if (GLU) load_aff4(a, s, c + Co, kg); else kg = kv;
    const float* xp = a.x + ((size_t)s * a.M) * a.C + c;
    const float* yp = a.dy + ((size_t)s * a.M) * Co + c;
    for (long long r_begin = (long long)blockIdx.y * g.rpb; r_begin < a.M; r_begin += (long long)gridDim.y * g.rpb) {
      long long r_end = r_begin + g.rpb;
      if (r_end > a.M) r_end = a.M;
      float f[NV][4];
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) f[k][j] = 0.f;
      // rows in batches of 4 with all 8-12 loads issued before the first use (the compiler did not hoist them across the
      // sigmoid / division code on its own: 3 loads in flight per thread held the GLU variant at 1.9 TB/s)
      auto accumulate = [&](const float4& xa, const float4& xb, const float4& ya) {
        float xv[4], xg[4] = {0.f, 0.f, 0.f, 0.f}, gy[4], dzv[4], dzg[4];
        to_arr(xa, xv);
        if (GLU) to_arr(xb, xg);
        to_arr(ya, gy);
        dz_row<ACT>(kv, kg, xv, xg, gy, dzv, dzg);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[0][j] += dzv[j];
          f[1][j] = fmaf(dzv[j], (xv[j] - kv.mu[j]) * kv.is[j], f[1][j]);
          if (GLU) {
            f[2 % NV][j] += dzg[j];
            f[3 % NV][j] = fmaf(dzg[j], (xg[j] - kg.mu[j]) * kg.is[j], f[3 % NV][j]);
          }
        }
      };
      long long r = r_begin + rl;
      for (; r + 3LL * g.R < r_end; r += 4LL * g.R) {
        float4 xa[4], xb[4], ya[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t rr = (size_t)(r + (long long)u * g.R);
          xa[u] = ld4(xp + rr * a.C);
          xb[u] = GLU ? ld4(xp + rr * a.C + Co) : make_float4(0.f, 0.f, 0.f, 0.f);
          ya[u] = ld4(yp + rr * Co);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) accumulate(xa[u], xb[u], ya[u]);
      }
      for (; r < r_end; r += g.R) {
        const float4 xa = ld4(xp + (size_t)r * a.C);
        const float4 xb = GLU ? ld4(xp + (size_t)r * a.C + Co) : make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate(xa, xb, ld4(yp + (size_t)r * Co));
      }
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[k][j] += (double)f[k][j];
    }
  }
  const int p = blockIdx.y;
  if constexpr (GLU) {
    double* const dst[4] = {part_slot(part, p, 0, a.S, a.C, s, c), part_slot(part, p, 1, a.S, a.C, s, c),
                            part_slot(part, p, 0, a.S, a.C, s, c + Co), part_slot(part, p, 1, a.S, a.C, s, c + Co)};
    v4_block_reduce<4>(v, g.GB, g.R, cgl, rl, valid, dst);
  } else {
    double* const dst[2] = {part_slot(part, p, 0, a.S, a.C, s, c), part_slot(part, p, 1, a.S, a.C, s, c)};
    v4_block_reduce<2>(v, g.GB, g.R, cgl, rl, valid, dst);
  }
}

__device__ __forceinline__ void emit_planes4(const float (&o)[4], __nv_bfloat16* hi, __nv_bfloat16* lo) {
  __nv_bfloat162 h01 = __floats2bfloat162_rn(o[0], o[1]), h23 = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
  if (lo) {
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    __nv_bfloat162 l01 = __floats2bfloat162_rn(o[0] - f01.x, o[1] - f01.y), l23 = __floats2bfloat162_rn(o[2] - f23.x, o[3] - f23.y);
    *reinterpret_cast<uint2*>(lo) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
  }
}

template <int ACT>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_v4_kernel(BnBwdArgs a, V4Geom g, const double* __restrict__ dgamma, const double* __restrict__ dbeta,
                                       float* __restrict__ dx, __nv_bfloat16* __restrict__ phi, __nv_bfloat16* __restrict__ plo) {
  constexpr bool GLU = ACT == MOG_ACT_GLU;
  const int Co = GLU ? a.C / 2 : a.C;
  const int cgl = threadIdx.x % g.GB, rl = threadIdx.x / g.GB;
  const int cg = blockIdx.x * g.GB + cgl;
  if (cg >= g.Cv) return;
  const int c = cg * 4, s = blockIdx.z;
  const long long r_begin = (long long)blockIdx.y * g.rpb;
  long long r_end = r_begin + g.rpb;
  if (r_end > a.M) r_end = a.M;
  Aff4 kv, kg;
  load_aff4(a, s, c, kv);
  if (GLU) load_aff4(a, s, c + Co, kg); else kg = kv;
  float k1v[4], k2v[4], k1g[4], k2g[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k1v[j] = dbeta ? (float)(dbeta[(size_t)s * a.C + c + j] / (double)a.M) : 0.f;
    k2v[j] = dgamma ? (float)(dgamma[(size_t)s * a.C + c + j] / (double)a.M) : 0.f;
    k1g[j] = (GLU && dbeta) ? (float)(dbeta[(size_t)s * a.C + c + Co + j] / (double)a.M) : 0.f;
    k2g[j] = (GLU && dgamma) ? (float)(dgamma[(size_t)s * a.C + c + Co + j] / (double)a.M) : 0.f;
  }
  const float* xp = a.x + ((size_t)s * a.M) * a.C + c;
  const float* yp = a.dy + ((size_t)s * a.M) * Co + c;
  float* dp = dx + ((size_t)s * a.M) * a.C + c;
  auto apply_row = [&](long long r, const float4& xa, const float4& xb, const float4& ya) {
    float xv[4], xg[4] = {0.f, 0.f, 0.f, 0.f}, gy[4], dzv[4], dzg[4], o[4];
    to_arr(xa, xv);
    if (GLU) to_arr(xb, xg);
    to_arr(ya, gy);
    dz_row<ACT>(kv, kg, xv, xg, gy, dzv, dzg);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = kv.sc[j] * (dzv[j] - k1v[j] - (xv[j] - kv.mu[j]) * kv.is[j] * k2v[j]);
    *reinterpret_cast<float4*>(dp + (size_t)r * a.C) = make_float4(o[0], o[1], o[2], o[3]);
    // dx also as the bf16 hi (+ lo) planes the data / weight gradient kernels of the producing convolution read (C % 8 == 0)
    const size_t prow = ((size_t)s * a.M + (size_t)r) * a.C + c;
    if (phi) emit_planes4(o, phi + prow, plo ? plo + prow : nullptr);
    if (GLU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = kg.sc[j] * (dzg[j] - k1g[j] - (xg[j] - kg.mu[j]) * kg.is[j] * k2g[j]);
      *reinterpret_cast<float4*>(dp + (size_t)r * a.C + Co) = make_float4(o[0], o[1], o[2], o[3]);
      if (phi) emit_planes4(o, phi + prow + Co, plo ? plo + prow + Co : nullptr);
    }
  };
  long long r = r_begin + rl;
  for (; r + 3LL * g.R < r_end; r += 4LL * g.R) {   // 4 rows per batch, loads first
    float4 xa[4], xb[4], ya[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t rr = (size_t)(r + (long long)u * g.R);
      xa[u] = ld4(xp + rr * a.C);
      xb[u] = GLU ? ld4(xp + rr * a.C + Co) : make_float4(0.f, 0.f, 0.f, 0.f);
      ya[u] = ld4(yp + rr * Co);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) apply_row(r + (long long)u * g.R, xa[u], xb[u], ya[u]);
  }
  for (; r < r_end; r += g.R) {
    const float4 xa = ld4(xp + (size_t)r * a.C);
    const float4 xb = GLU ? ld4(xp + (size_t)r * a.C + Co) : make_float4(0.f, 0.f, 0.f, 0.f);
    apply_row(r, xa, xb, ld4(yp + (size_t)r * Co));
  }
}

template <int ACT>
static void launch_bwd_v4(bool reduce, const BnBwdArgs& a, const V4Geom& g, double* dgamma, double* dbeta, float* dx, cudaStream_t st,
                          __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr) {
  dim3 grid(g.nxb, g.nyb, a.S);
  if (reduce)
    bn_bwd_reduce_v4_kernel<ACT><<<grid, g.GB * g.R, 0, st>>>(a, g, dgamma /* = the partial buffer */);
  else
    bn_bwd_apply_v4_kernel<ACT><<<grid, g.GB * g.R, 0, st>>>(a, g, dgamma, dbeta, dx, phi, plo);
}
static bool bwd_v4(bool reduce, const BnBwdArgs& a, double* dgamma, double* dbeta, float* dx, cudaStream_t st,
                   __nv_bfloat16* phi = nullptr, __nv_bfloat16* plo = nullptr) {
  const int width = a.act == MOG_ACT_GLU ? a.C / 2 : a.C;
  if ((width & 3) || (a.C & 3) || a.S > 65535) return false;
  const V4Geom g = v4_geom(width, a.M, a.S, reduce);
  switch (a.act) {
    case MOG_ACT_NONE: launch_bwd_v4<MOG_ACT_NONE>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    case MOG_ACT_RELU: launch_bwd_v4<MOG_ACT_RELU>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    case MOG_ACT_LRELU: launch_bwd_v4<MOG_ACT_LRELU>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    case MOG_ACT_GLU: launch_bwd_v4<MOG_ACT_GLU>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    case MOG_ACT_TANH: launch_bwd_v4<MOG_ACT_TANH>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    case MOG_ACT_SIGMOID: launch_bwd_v4<MOG_ACT_SIGMOID>(reduce, a, g, dgamma, dbeta, dx, st, phi, plo); break;
    default: return false;
  }
  return true;
}

}  // namespace mog

using namespace mog;

// number of partial slots P of the reducing kernels for a problem (which = 0: forward statistics, 1: backward sums)
static int bn_parts_of(int S, int M, int C, int act, int which) {
  const int width = (which == 1 && act == MOG_ACT_GLU) ? C / 2 : C;
  if ((C & 3) == 0 && (width & 3) == 0 && S <= 65535) return v4_geom(width, M, S, true, which).nyb;
  const int chunks = ceil_div(M, rows_per_block(M));
  return chunks < 64 ? chunks : 64;
}

extern "C" int mog_bn_parts(int S, int M, int C, int act, int which) {
  if (S <= 0 || M <= 0 || C <= 0) return 0;
  return bn_parts_of(S, M, C, act, which);
}

extern "C" int mog_bn_stats(const float* x, int S, int M, int C, double* part, int nparts, void* stream) {
  MOG_REQUIRE(x && part && S > 0 && M > 0 && C > 0, "mog_bn_stats: bad argument");
  MOG_REQUIRE(nparts == bn_parts_of(S, M, C, MOG_ACT_NONE, 0), "mog_bn_stats: nparts must be mog_bn_parts(S, M, C, act, 0)");
  cudaStream_t st = as_stream(stream);
  MOG_REQUIRE(S <= 65535, "mog_bn_stats: too many segments");
  if ((C & 3) == 0) {
    const V4Geom g = v4_geom(C, M, S, true, 0);
    bn_stats_v4_kernel<<<dim3(g.nxb, g.nyb, S), g.GB * g.R, 0, st>>>(x, M, C, g, part, S);
    return check_launch("bn_stats_v4_kernel");
  }
  const int rpb = rows_per_block(M);
  dim3 grid(ceil_div(C, 32), nparts, S);
  bn_stats_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(x, M, C, rpb, part, S);
  return check_launch("bn_stats_kernel");
}

extern "C" int mog_bn_finalize(const double* part, int nparts, int S, int M, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean,
                               float* running_var, float* mean, float* invstd, float* scale, float* shift,
                               void* stream) {
  MOG_REQUIRE(part && nparts > 0 && mean && invstd && scale && shift && S > 0 && M > 0 && C > 0, "mog_bn_finalize: bad argument");
  bn_finalize_kernel<<<ceil_div(C, 8), 256, 0, as_stream(stream)>>>(part, nparts, S, M, C, gamma, beta, eps, momentum,
                                                                     running_mean, running_var, mean, invstd, scale, shift);
  return check_launch("bn_finalize_kernel");
}

extern "C" int mog_affine_act_fwd(const float* x, const float* scale, const float* shift, const float* residual,
                                  float* y, int S, int M, int C, int act, void* stream) {
  return mog_affine_act_fwd_planes(x, scale, shift, residual, y, nullptr, MOG_PREC_FP32, S, M, C, act, stream);
}

extern "C" int mog_affine_act_fwd_planes(const float* x, const float* scale, const float* shift, const float* residual,
                                         float* y, void* y_planes, int precision, int S, int M, int C, int act, void* stream) {
  MOG_REQUIRE(x && y && S > 0 && M > 0 && C > 0, "mog_affine_act_fwd: bad argument");
  MOG_REQUIRE((scale == nullptr) == (shift == nullptr), "mog_affine_act_fwd: scale/shift must both be given or NULL");
  MOG_REQUIRE(act != MOG_ACT_GLU || (C % 2) == 0, "mog_affine_act_fwd: GLU needs an even channel count");
  const int Co = act == MOG_ACT_GLU ? C / 2 : C;
  size_t total = (size_t)S * M * Co;
  cudaStream_t st = as_stream(stream);
  __nv_bfloat16 *phi = nullptr, *plo = nullptr;
  const int CP = ceil_div(Co, 8) * 8;
  if (y_planes) {
    MOG_REQUIRE(precision == MOG_PREC_BF16X3 || precision == MOG_PREC_BF16, "mog_affine_act_fwd_planes: precision must be a tcgen05 mode");
    MOG_REQUIRE((Co & 3) == 0 && (C & 3) == 0, "mog_affine_act_fwd_planes: channel counts must be multiples of 4");
    phi = static_cast<__nv_bfloat16*>(y_planes);
    if (precision == MOG_PREC_BF16X3) plo = phi + (size_t)S * M * CP;
  }
  if ((Co & 3) == 0 && (C & 3) == 0) {
    size_t nthreads = total / 4;
    affine_act_fwd_kernel<4><<<(unsigned)ceil_div_ll((long long)nthreads, 256), 256, 0, st>>>(x, scale, shift, residual, y, M, C, act, total, phi, plo, CP);
  } else {
    affine_act_fwd_kernel<1><<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, st>>>(x, scale, shift, residual, y, M, C, act, total, nullptr, nullptr, CP);
  }
  return check_launch("affine_act_fwd_kernel");
}

extern "C" int mog_bn_act_bwd_reduce(const float* x, const float* dy, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int S, int M, int C, int act,
                                     double* part, int nparts, double* dgamma_seg, double* dbeta_seg, float* dgamma,
                                     float* dbeta, void* stream) {
  MOG_REQUIRE(x && dy && mean && invstd && part && dgamma_seg && dbeta_seg && S > 0 && M > 0 && C > 0, "mog_bn_act_bwd_reduce: bad argument");
  MOG_REQUIRE(nparts == bn_parts_of(S, M, C, act, 1), "mog_bn_act_bwd_reduce: nparts must be mog_bn_parts(S, M, C, act, 1)");
  cudaStream_t st = as_stream(stream);
  const int rpb = rows_per_block(M);
  BnBwdArgs a{x, dy, mean, invstd, gamma, beta, S, M, C, act, rpb};
  int rc;
  if (bwd_v4(true, a, part, nullptr, nullptr, st)) {
    rc = check_launch("bn_bwd_reduce_v4_kernel");
  } else {
    dim3 grid(ceil_div(C, 32), nparts, S);
    MOG_REQUIRE(grid.z <= 65535, "mog_bn_act_bwd_reduce: too many segments");
    bn_bwd_reduce_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(a, part);
    rc = check_launch("bn_bwd_reduce_kernel");
  }
  if (rc) return rc;
  // partials -> per-segment sums (read by the apply pass) + parameter gradients, fixed order
  bn_bwd_sum_kernel<<<ceil_div(C, 8), 256, 0, st>>>(part, nparts, S, C, dgamma_seg, dbeta_seg, dgamma, dbeta);
  return check_launch("bn_bwd_sum_kernel");
}

extern "C" int mog_bn_act_bwd_apply(const float* x, const float* dy, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const double* dgamma_seg,
                                    const double* dbeta_seg, int S, int M, int C, int act, float* dx, void* stream) {
  return mog_bn_act_bwd_apply_planes(x, dy, mean, invstd, gamma, beta, dgamma_seg, dbeta_seg, S, M, C, act, dx, nullptr, MOG_PREC_FP32,
                                     stream);
}

extern "C" int mog_bn_act_bwd_apply_planes(const float* x, const float* dy, const float* mean, const float* invstd,
                                           const float* gamma, const float* beta, const double* dgamma_seg,
                                           const double* dbeta_seg, int S, int M, int C, int act, float* dx, void* dx_planes,
                                           int precision, void* stream) {
  MOG_REQUIRE(x && dy && dx && S > 0 && M > 0 && C > 0, "mog_bn_act_bwd_apply: bad argument");
  __nv_bfloat16* phi = nullptr;
  __nv_bfloat16* plo = nullptr;
  if (dx_planes) {
    const int width = act == MOG_ACT_GLU ? C / 2 : C;
    MOG_REQUIRE(precision == MOG_PREC_BF16X3 || precision == MOG_PREC_BF16, "mog_bn_act_bwd_apply_planes: precision must be a tcgen05 mode");
    MOG_REQUIRE((C & 7) == 0 && (width & 3) == 0, "mog_bn_act_bwd_apply_planes: plane emission needs C %% 8 == 0 (got %d)", C);
    phi = static_cast<__nv_bfloat16*>(dx_planes);
    plo = precision == MOG_PREC_BF16X3 ? phi + (size_t)S * M * C : nullptr;
  }
  const bool has_bn = mean != nullptr;
  MOG_REQUIRE(has_bn == (invstd != nullptr) && has_bn == (dgamma_seg != nullptr) && has_bn == (dbeta_seg != nullptr),
              "mog_bn_act_bwd_apply: mean/invstd/dgamma_seg/dbeta_seg must all be given or all be NULL");
  cudaStream_t st = as_stream(stream);
  const int rpb = rows_per_block(M);
  BnBwdArgs a{x, dy, mean, invstd, gamma, beta, S, M, C, act, rpb};
  int rc;
  if (bwd_v4(false, a, const_cast<double*>(dgamma_seg), const_cast<double*>(dbeta_seg), dx, st, phi, plo)) {
    rc = check_launch("bn_bwd_apply_v4_kernel");
  } else {
    MOG_REQUIRE(!dx_planes, "mog_bn_act_bwd_apply_planes: this shape runs on the scalar kernel, which does not emit planes");
    dim3 grid(ceil_div(C, 32), ceil_div(M, rpb), S);
    MOG_REQUIRE(grid.z <= 65535, "mog_bn_act_bwd_apply: too many segments");
    bn_bwd_apply_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(a, dgamma_seg, dbeta_seg, dx);
    rc = check_launch("bn_bwd_apply_kernel");
  }
  return rc;
}

extern "C" int mog_act_bwd(const float* dy, const float* y, float* dz, size_t n, int act, void* stream) {
  MOG_REQUIRE(dy && y && dz && n > 0, "mog_act_bwd: bad argument");
  act_bwd_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, as_stream(stream)>>>(dy, y, dz, n, act);
  return check_launch("act_bwd_kernel");
}
