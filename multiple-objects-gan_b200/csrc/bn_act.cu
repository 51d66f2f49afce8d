// bn_act.cu -- train-mode BatchNorm statistics, fused affine + activation (GLU / ReLU /
// LeakyReLU / none, optional residual) forward, and the two-pass backward.  HBM-bound streaming
// kernels over [S*M][C] row-major (NHWC) tensors: per-channel reductions run over rows with
// 32 consecutive channels per warp row (128-byte coalesced), partials are combined with
// warp/smem reductions and one double atomicAdd per (block, channel).
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
__global__ void bn_stats_kernel(const float* __restrict__ x, int M, int C, int rpb, double* __restrict__ sum,
                                double* __restrict__ sqsum) {
  // grid: (ceil(C/32), row chunks, S)
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int s = blockIdx.z;
  const long long r_begin = (long long)blockIdx.y * rpb;
  long long r_end = r_begin + rpb;
  if (r_end > M) r_end = M;
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float* xp = x + ((size_t)s * M) * C + c;
    for (long long r = r_begin + rl; r < r_end; r += RED_ROWS) {
      float v = __ldg(xp + (size_t)r * C);
      a += v;
      b = fmaf(v, v, b);
    }
  }
  __shared__ float sa[RED_ROWS][33], sb[RED_ROWS][33];
  sa[rl][lane] = a;
  sb[rl][lane] = b;
  __syncthreads();
  if (rl == 0 && c < C) {
    double ta = 0.0, tb = 0.0;
#pragma unroll
    for (int i = 0; i < RED_ROWS; ++i) {
      ta += (double)sa[i][lane];
      tb += (double)sb[i][lane];
    }
    atomicAdd(sum + (size_t)s * C + c, ta);
    atomicAdd(sqsum + (size_t)s * C + c, tb);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sqsum, int S, int M,
                                   int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean,
                                   float* invstd, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f;
  float rv = running_var ? running_var[c] : 0.f;
  for (int s = 0; s < S; ++s) {
    double mu = sum[(size_t)s * C + c] / (double)M;
    double var = sqsum[(size_t)s * C + c] / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    float is = (float)(1.0 / sqrt(var + (double)eps));
    float muf = (float)mu;
    mean[(size_t)s * C + c] = muf;
    invstd[(size_t)s * C + c] = is;
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * is;
    scale[(size_t)s * C + c] = sc;
    shift[(size_t)s * C + c] = b - muf * sc;
    // nn.BatchNorm: running = (1-m)*running + m*stat, unbiased variance for running_var
    double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
    rm = (1.f - momentum) * rm + momentum * muf;
    rv = (1.f - momentum) * rv + momentum * (float)unb;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
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
                                      float* __restrict__ y, int M, int C, int act, size_t total_out) {
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

__global__ void bn_bwd_reduce_kernel(BnBwdArgs a, double* __restrict__ dgamma, double* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int s = blockIdx.z;
  const int Co = a.act == MOG_ACT_GLU ? a.C / 2 : a.C;
  const long long r_begin = (long long)blockIdx.y * a.rows_per_block;
  long long r_end = r_begin + a.rows_per_block;
  if (r_end > a.M) r_end = a.M;
  float g1 = 0.f, g2 = 0.f;
  if (c < a.C) {
    const ChanConst k = chan_const(a, s, c);
    for (long long r = r_begin + rl; r < r_end; r += RED_ROWS) {
      size_t row = (size_t)s * a.M + r;
      const float* xr = a.x + row * a.C;
      float xv = __ldg(xr + c);
      float dz = dz_of(k, xv, xr, a.dy + row * Co, a.act);
      g1 += dz;
      g2 = fmaf(dz, (xv - k.mu) * k.is, g2);
    }
  }
  __shared__ float sa[RED_ROWS][33], sb[RED_ROWS][33];
  sa[rl][lane] = g1;
  sb[rl][lane] = g2;
  __syncthreads();
  if (rl == 0 && c < a.C) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int i = 0; i < RED_ROWS; ++i) {
      t1 += (double)sa[i][lane];
      t2 += (double)sb[i][lane];
    }
    atomicAdd(dbeta + (size_t)s * a.C + c, t1);
    atomicAdd(dgamma + (size_t)s * a.C + c, t2);
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

__global__ void bn_bwd_param_kernel(const double* __restrict__ dgamma_seg, const double* __restrict__ dbeta_seg,
                                    int S, int C, float* dgamma, float* dbeta) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double g = 0.0, b = 0.0;
  for (int s = 0; s < S; ++s) {
    g += dgamma_seg[(size_t)s * C + c];
    b += dbeta_seg[(size_t)s * C + c];
  }
  if (dgamma) dgamma[c] = (float)g;
  if (dbeta) dbeta[c] = (float)b;
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

}  // namespace mog

using namespace mog;

extern "C" int mog_bn_stats(const float* x, int S, int M, int C, double* sum, double* sqsum, void* stream) {
  MOG_REQUIRE(x && sum && sqsum && S > 0 && M > 0 && C > 0, "mog_bn_stats: bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(sum, 0, sizeof(double) * S * C, st);
  cudaMemsetAsync(sqsum, 0, sizeof(double) * S * C, st);
  const int rpb = rows_per_block(M);
  dim3 grid(ceil_div(C, 32), ceil_div(M, rpb), S);
  MOG_REQUIRE(grid.z <= 65535, "mog_bn_stats: too many segments");
  bn_stats_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(x, M, C, rpb, sum, sqsum);
  return check_launch("bn_stats_kernel");
}

extern "C" int mog_bn_finalize(const double* sum, const double* sqsum, int S, int M, int C, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean,
                               float* running_var, float* mean, float* invstd, float* scale, float* shift,
                               void* stream) {
  MOG_REQUIRE(sum && sqsum && mean && invstd && scale && shift && S > 0 && M > 0 && C > 0, "mog_bn_finalize: bad argument");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(sum, sqsum, S, M, C, gamma, beta, eps, momentum,
                                                                     running_mean, running_var, mean, invstd, scale, shift);
  return check_launch("bn_finalize_kernel");
}

extern "C" int mog_affine_act_fwd(const float* x, const float* scale, const float* shift, const float* residual,
                                  float* y, int S, int M, int C, int act, void* stream) {
  MOG_REQUIRE(x && y && S > 0 && M > 0 && C > 0, "mog_affine_act_fwd: bad argument");
  MOG_REQUIRE((scale == nullptr) == (shift == nullptr), "mog_affine_act_fwd: scale/shift must both be given or NULL");
  MOG_REQUIRE(act != MOG_ACT_GLU || (C % 2) == 0, "mog_affine_act_fwd: GLU needs an even channel count");
  const int Co = act == MOG_ACT_GLU ? C / 2 : C;
  size_t total = (size_t)S * M * Co;
  cudaStream_t st = as_stream(stream);
  if ((Co & 3) == 0 && (C & 3) == 0) {
    size_t nthreads = total / 4;
    affine_act_fwd_kernel<4><<<(unsigned)ceil_div_ll((long long)nthreads, 256), 256, 0, st>>>(x, scale, shift, residual, y, M, C, act, total);
  } else {
    affine_act_fwd_kernel<1><<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, st>>>(x, scale, shift, residual, y, M, C, act, total);
  }
  return check_launch("affine_act_fwd_kernel");
}

extern "C" int mog_bn_act_bwd_reduce(const float* x, const float* dy, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int S, int M, int C, int act,
                                     double* dgamma_seg, double* dbeta_seg, void* stream) {
  MOG_REQUIRE(x && dy && mean && invstd && dgamma_seg && dbeta_seg && S > 0 && M > 0 && C > 0, "mog_bn_act_bwd_reduce: bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(dgamma_seg, 0, sizeof(double) * S * C, st);
  cudaMemsetAsync(dbeta_seg, 0, sizeof(double) * S * C, st);
  const int rpb = rows_per_block(M);
  BnBwdArgs a{x, dy, mean, invstd, gamma, beta, S, M, C, act, rpb};
  dim3 grid(ceil_div(C, 32), ceil_div(M, rpb), S);
  MOG_REQUIRE(grid.z <= 65535, "mog_bn_act_bwd_reduce: too many segments");
  bn_bwd_reduce_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(a, dgamma_seg, dbeta_seg);
  return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int mog_bn_act_bwd_apply(const float* x, const float* dy, const float* mean, const float* invstd,
                                    const float* gamma, const float* beta, const double* dgamma_seg,
                                    const double* dbeta_seg, int S, int M, int C, int act, float* dx, float* dgamma,
                                    float* dbeta, void* stream) {
  MOG_REQUIRE(x && dy && dx && S > 0 && M > 0 && C > 0, "mog_bn_act_bwd_apply: bad argument");
  const bool has_bn = mean != nullptr;
  MOG_REQUIRE(has_bn == (invstd != nullptr) && has_bn == (dgamma_seg != nullptr) && has_bn == (dbeta_seg != nullptr),
              "mog_bn_act_bwd_apply: mean/invstd/dgamma_seg/dbeta_seg must all be given or all be NULL");
  cudaStream_t st = as_stream(stream);
  const int rpb = rows_per_block(M);
  BnBwdArgs a{x, dy, mean, invstd, gamma, beta, S, M, C, act, rpb};
  dim3 grid(ceil_div(C, 32), ceil_div(M, rpb), S);
  MOG_REQUIRE(grid.z <= 65535, "mog_bn_act_bwd_apply: too many segments");
  bn_bwd_apply_kernel<<<grid, 32 * RED_ROWS, 0, st>>>(a, dgamma_seg, dbeta_seg, dx);
  int rc = check_launch("bn_bwd_apply_kernel");
  if (rc) return rc;
  if (has_bn && (dgamma || dbeta)) {
    bn_bwd_param_kernel<<<ceil_div(C, 128), 128, 0, st>>>(dgamma_seg, dbeta_seg, S, C, dgamma, dbeta);
    return check_launch("bn_bwd_param_kernel");
  }
  return MOG_OK;
}

extern "C" int mog_act_bwd(const float* dy, const float* y, float* dz, size_t n, int act, void* stream) {
  MOG_REQUIRE(dy && y && dz && n > 0, "mog_act_bwd: bad argument");
  act_bwd_kernel<<<(unsigned)ceil_div_ll((long long)n, 256), 256, 0, as_stream(stream)>>>(dy, y, dz, n, act);
  return check_launch("act_bwd_kernel");
}
