// pool.cu -- pooling and bilinear resize of the DAMSM image encoder (NHWC fp32, HBM-bound streaming kernels).
//
// replaces (code/coco/attngan/model.py): nn.Upsample(size=(299, 299), mode='bilinear') (:256), F.max_pool2d(x, 3, 2)
// (:264,271 and inside Mixed_6a / Mixed_7a), F.avg_pool2d(x, 3, 1, 1) of the Inception branch_pool paths,
// F.avg_pool2d(x, 8) (:301) -- and their autograd (the encoder is frozen, but the gradient w.r.t. the generated image
// flows through every one of them, losses.py:205-224).
//
// One thread per output element, channel fastest (coalesced 4-byte accesses over C).  Backward passes are deterministic
// gathers: every input element sums the contributions of the (few) output windows that cover it; max pooling recomputes
// each covering window's arg-max with torch's tie rule (first maximum in (kh, kw) scan order wins).
#include <cfloat>

#include "common.cuh"

namespace mog {

struct PoolArgs {
  const float* x;
  const float* dy;
  float* out;
  unsigned char* idx;   // max pooling: window position kh*k + kw of the arg-max per output element (forward writes, backward reads)
  int N, H, W, C, Ho, Wo, k, s, p, mode;   // mode 0: max, 1: average (divisor k*k, count_include_pad)
  long long total;
};

// V = channels per thread (4: 128-bit accesses when C % 4 == 0; 1: any C)
template <int V>
struct Vec { float v[V]; };
template <int V>
__device__ __forceinline__ Vec<V> ldv(const float* p) {
  Vec<V> r;
  if (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1 % V] = t.y; r.v[2 % V] = t.z; r.v[3 % V] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const Vec<V>& r) {
  if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1 % V], r.v[2 % V], r.v[3 % V]);
  else *p = r.v[0];
}

// K: compile-time window edge (3 = every pooling layer of Inception-v3 but the final 8x8 average; 0 = run-time a.k).  With K known the
// K*K taps are loaded up front -- nine independent 128-bit loads in flight per thread; the run-time loop issued one load per
// iteration and waited for it (ncu r2j: 2.6 TB/s on the 147^2 x 64 max pool, a pure latency bound).
template <int V, int K>
__global__ void pool_fwd_kernel(PoolArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // element index / V
  if (i >= a.total) return;
  const int Cv = a.C / V;
  const int c = (int)(i % Cv) * V;
  long long q = i / Cv;
  const int wo = (int)(q % a.Wo);
  q /= a.Wo;
  const int ho = (int)(q % a.Ho);
  const int n = (int)(q / a.Ho);
  const float* xn = a.x + (size_t)n * a.H * a.W * a.C + c;
  Vec<V> acc;
  int best[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { acc.v[j] = a.mode == 0 ? -FLT_MAX : 0.f; best[j] = -1; }
  if (K > 0) {
    constexpr int KK = K > 0 ? K * K : 1;
    Vec<V> v[KK];
    bool ok[KK];
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      const int h = ho * a.s - a.p + t / (K > 0 ? K : 1), w = wo * a.s - a.p + t % (K > 0 ? K : 1);
      ok[t] = h >= 0 && h < a.H && w >= 0 && w < a.W;
      if (ok[t]) v[t] = ldv<V>(xn + ((size_t)h * a.W + w) * a.C);
      else v[t] = Vec<V>{};
    }
#pragma unroll
    for (int t = 0; t < KK; ++t) {
      if (!ok[t]) continue;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        if (a.mode == 0) {
          if (v[t].v[j] > acc.v[j] || v[t].v[j] != v[t].v[j] || best[j] < 0) { acc.v[j] = v[t].v[j]; best[j] = t; }   // first maximum wins (torch)
        } else {
          acc.v[j] += v[t].v[j];
        }
      }
    }
  } else {
    for (int kh = 0; kh < a.k; ++kh) {
      const int h = ho * a.s - a.p + kh;
      if (h < 0 || h >= a.H) continue;
      for (int kw = 0; kw < a.k; ++kw) {
        const int w = wo * a.s - a.p + kw;
        if (w < 0 || w >= a.W) continue;
        const Vec<V> v = ldv<V>(xn + ((size_t)h * a.W + w) * a.C);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          if (a.mode == 0) {
            if (v.v[j] > acc.v[j] || v.v[j] != v.v[j] || best[j] < 0) { acc.v[j] = v.v[j]; best[j] = kh * a.k + kw; }   // first maximum wins (torch)
          } else {
            acc.v[j] += v.v[j];
          }
        }
      }
    }
  }
  if (a.mode == 1) {
#pragma unroll
    for (int j = 0; j < V; ++j) acc.v[j] = acc.v[j] / (float)(a.k * a.k);
  }
  stv<V>(a.out + i * V, acc);
  if (a.idx) {
    if (V == 4) {
      *reinterpret_cast<uchar4*>(a.idx + i * V) = make_uchar4((unsigned char)best[0], (unsigned char)best[1 % V], (unsigned char)best[2 % V],
                                                               (unsigned char)best[3 % V]);
    } else {
      a.idx[i] = (unsigned char)best[0];
    }
  }
}

__device__ __forceinline__ int fdiv_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// WM: compile-time bound of the covering windows per axis (3 for k = 3; 0 = run-time loops).  As in the forward kernel the
// window gradients (and recorded arg-max positions) are loaded up front, WM*WM independent loads in flight.
template <int V, int WM>
__global__ void pool_bwd_kernel(PoolArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  const int Cv = a.C / V;
  const int c = (int)(i % Cv) * V;
  long long q = i / Cv;
  const int w = (int)(q % a.W);
  q /= a.W;
  const int h = (int)(q % a.H);
  const int n = (int)(q / a.H);
  // output windows covering (h, w): ho*s - p <= h <= ho*s - p + k - 1
  int ho_lo = fdiv_i(h + a.p - a.k + a.s, a.s), ho_hi = fdiv_i(h + a.p, a.s);
  int wo_lo = fdiv_i(w + a.p - a.k + a.s, a.s), wo_hi = fdiv_i(w + a.p, a.s);
  ho_lo = max(ho_lo, 0); wo_lo = max(wo_lo, 0);
  ho_hi = min(ho_hi, a.Ho - 1); wo_hi = min(wo_hi, a.Wo - 1);
  const float* xn = a.x ? a.x + (size_t)n * a.H * a.W * a.C + c : nullptr;
  const float* dyn = a.dy + (size_t)n * a.Ho * a.Wo * a.C + c;
  Vec<V> acc;
#pragma unroll
  for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
  if (WM > 0 && (a.mode == 1 || a.idx)) {
    constexpr int WW = WM > 0 ? WM * WM : 1;
    Vec<V> g[WW];
    unsigned int pos[WW];
    bool ok[WW];
#pragma unroll
    for (int t = 0; t < WW; ++t) {
      const int ho = ho_lo + t / (WM > 0 ? WM : 1), wo = wo_lo + t % (WM > 0 ? WM : 1);
      ok[t] = ho <= ho_hi && wo <= wo_hi;
      pos[t] = 0u;
      if (ok[t]) {
        g[t] = ldv<V>(dyn + ((size_t)ho * a.Wo + wo) * a.C);
        if (a.mode == 0) {
          const unsigned char* ip = a.idx + ((size_t)n * a.Ho * a.Wo + (size_t)ho * a.Wo + wo) * a.C + c;
          pos[t] = V == 4 ? *reinterpret_cast<const unsigned int*>(ip) : (unsigned int)*ip;
        }
      } else {
        g[t] = Vec<V>{};
      }
    }
#pragma unroll
    for (int t = 0; t < WW; ++t) {
      if (!ok[t]) continue;
      const int ho = ho_lo + t / (WM > 0 ? WM : 1), wo = wo_lo + t % (WM > 0 ? WM : 1);
      const int me = (h - (ho * a.s - a.p)) * a.k + (w - (wo * a.s - a.p));   // this pixel's position inside the window
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (a.mode == 1 || (int)((pos[t] >> (8 * j)) & 0xffu) == me) acc.v[j] += g[t].v[j];
    }
  } else {
    for (int ho = ho_lo; ho <= ho_hi; ++ho)
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const Vec<V> g = ldv<V>(dyn + ((size_t)ho * a.Wo + wo) * a.C);
        if (a.mode == 1) {
#pragma unroll
          for (int j = 0; j < V; ++j) acc.v[j] += g.v[j];
        } else if (a.idx) {
          // the forward recorded each window's arg-max position
          const unsigned char* ip = a.idx + ((size_t)n * a.Ho * a.Wo + (size_t)ho * a.Wo + wo) * a.C + c;
          unsigned char pos[V];
          if (V == 4) {
            const uchar4 t = *reinterpret_cast<const uchar4*>(ip);
            pos[0] = t.x; pos[1 % V] = t.y; pos[2 % V] = t.z; pos[3 % V] = t.w;
          } else {
            pos[0] = *ip;
          }
          const int me = (h - (ho * a.s - a.p)) * a.k + (w - (wo * a.s - a.p));   // this pixel's position inside the window
#pragma unroll
          for (int j = 0; j < V; ++j)
            if ((int)pos[j] == me) acc.v[j] += g.v[j];
        } else {
          // arg-max of this window recomputed from x, first maximum wins (torch: `val > maxval || isnan(val)`)
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float best = -FLT_MAX;
            int bh = -1, bw = -1;
            for (int kh = 0; kh < a.k; ++kh) {
              const int hh = ho * a.s - a.p + kh;
              if (hh < 0 || hh >= a.H) continue;
              for (int kw = 0; kw < a.k; ++kw) {
                const int ww = wo * a.s - a.p + kw;
                if (ww < 0 || ww >= a.W) continue;
                const float v = __ldg(xn + ((size_t)hh * a.W + ww) * a.C + j);
                if (v > best || v != v || bh < 0) { best = v; bh = hh; bw = ww; }
              }
            }
            if (bh == h && bw == w) acc.v[j] += g.v[j];
          }
        }
      }
  }
  if (a.mode == 1) {
#pragma unroll
    for (int j = 0; j < V; ++j) acc.v[j] = acc.v[j] / (float)(a.k * a.k);
  }
  stv<V>(a.out + i * V, acc);
}

// ---- bilinear resize (torch upsample_bilinear2d) ------------------------------------------------
struct ResizeArgs {
  const float* src;
  float* dst;
  int N, Hi, Wi, C, Ho, Wo, align;
  float sh, sw;   // source step per output pixel
  long long total;
};

// source coordinate of output index o: (index of the first tap, 0/1 offset of the second tap, weight of the second tap)
__device__ __forceinline__ void bil_src(int o, float scale, int align, int in, int* i0, int* ip, float* l1) {
  float s = align ? scale * (float)o : scale * ((float)o + 0.5f) - 0.5f;
  if (!align && s < 0.f) s = 0.f;
  int i = (int)s;
  if (i > in - 1) i = in - 1;
  *i0 = i;
  *ip = i < in - 1 ? 1 : 0;
  *l1 = s - (float)i;
}

__global__ void resize_fwd_kernel(ResizeArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  const int c = (int)(i % a.C);
  long long q = i / a.C;
  const int wo = (int)(q % a.Wo);
  q /= a.Wo;
  const int ho = (int)(q % a.Ho);
  const int n = (int)(q / a.Ho);
  int h0, hp, w0, wp;
  float lh, lw;
  bil_src(ho, a.sh, a.align, a.Hi, &h0, &hp, &lh);
  bil_src(wo, a.sw, a.align, a.Wi, &w0, &wp, &lw);
  const float* p = a.src + (((size_t)n * a.Hi + h0) * a.Wi + w0) * a.C + c;
  const size_t dw = (size_t)wp * a.C, dh = (size_t)hp * a.Wi * a.C;
  const float h0l = 1.f - lh, w0l = 1.f - lw;
  a.dst[i] = h0l * (w0l * __ldg(p) + lw * __ldg(p + dw)) + lh * (w0l * __ldg(p + dh) + lw * __ldg(p + dh + dw));
}

// total weight with which output index o reads input index t along one axis
__device__ __forceinline__ float bil_weight(int o, int t, float scale, int align, int in) {
  int i0, ip;
  float l1;
  bil_src(o, scale, align, in, &i0, &ip, &l1);
  float w = 0.f;
  if (i0 == t) w += 1.f - l1;
  if (i0 + ip == t) w += l1;
  return w;
}

__global__ void resize_bwd_kernel(ResizeArgs a) {
  // a.src = dy [N,Ho,Wo,C], a.dst = dx [N,Hi,Wi,C]; one thread per input element gathers its outputs
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  const int c = (int)(i % a.C);
  long long q = i / a.C;
  const int w = (int)(q % a.Wi);
  q /= a.Wi;
  const int h = (int)(q % a.Hi);
  const int n = (int)(q / a.Hi);
  // outputs whose first tap is h-1 .. h: o in about [(h - 1 + 0.5)/scale - 0.5, (h + 1 + 0.5)/scale - 0.5]; two extra on each side
  const float ish = 1.f / a.sh, isw = 1.f / a.sw;
  int ho_lo = (int)floorf(((float)h - 1.f) * ish) - 2, ho_hi = (int)ceilf(((float)h + 1.5f) * ish) + 2;
  int wo_lo = (int)floorf(((float)w - 1.f) * isw) - 2, wo_hi = (int)ceilf(((float)w + 1.5f) * isw) + 2;
  ho_lo = max(ho_lo, 0); wo_lo = max(wo_lo, 0);
  ho_hi = min(ho_hi, a.Ho - 1); wo_hi = min(wo_hi, a.Wo - 1);
  const float* dyn = a.src + (size_t)n * a.Ho * a.Wo * a.C + c;
  float acc = 0.f;
  for (int ho = ho_lo; ho <= ho_hi; ++ho) {
    const float wh = bil_weight(ho, h, a.sh, a.align, a.Hi);
    if (wh == 0.f) continue;
    float row = 0.f;
    for (int wo = wo_lo; wo <= wo_hi; ++wo) {
      const float ww = bil_weight(wo, w, a.sw, a.align, a.Wi);
      if (ww != 0.f) row += ww * __ldg(dyn + ((size_t)ho * a.Wo + wo) * a.C);
    }
    acc += wh * row;
  }
  a.dst[i] = acc;
}

}  // namespace mog

using namespace mog;

static int pool_out(int H, int k, int s, int p) { return (H + 2 * p - k) / s + 1; }

extern "C" int mog_pool2d_out_hw(int H, int W, int k, int stride, int pad, int* Ho, int* Wo) {
  MOG_REQUIRE(H > 0 && W > 0 && k > 0 && stride > 0 && pad >= 0 && 2 * pad <= k && H + 2 * pad >= k && W + 2 * pad >= k,
              "mog_pool2d_out_hw: bad geometry");
  if (Ho) *Ho = pool_out(H, k, stride, pad);
  if (Wo) *Wo = pool_out(W, k, stride, pad);
  return MOG_OK;
}

extern "C" int mog_pool2d_fwd(const float* x, float* y, unsigned char* argmax, int N, int H, int W, int C, int k, int stride, int pad,
                              int mode, void* stream) {
  MOG_REQUIRE(x && y && N > 0 && C > 0 && (mode == 0 || mode == 1), "mog_pool2d_fwd: bad argument");
  MOG_REQUIRE(!argmax || (mode == 0 && k * k <= 255), "mog_pool2d_fwd: arg-max positions exist for max pooling with k*k <= 255 only");
  int Ho, Wo;
  int rc = mog_pool2d_out_hw(H, W, k, stride, pad, &Ho, &Wo);
  if (rc) return rc;
  const int V = (C & 3) == 0 ? 4 : 1;
  PoolArgs a{x, nullptr, y, argmax, N, H, W, C, Ho, Wo, k, stride, pad, mode, (long long)N * Ho * Wo * C / V};
  const unsigned grid = (unsigned)ceil_div_ll(a.total, 256);
  if (V == 4 && k == 3) pool_fwd_kernel<4, 3><<<grid, 256, 0, as_stream(stream)>>>(a);
  else if (V == 4) pool_fwd_kernel<4, 0><<<grid, 256, 0, as_stream(stream)>>>(a);
  else pool_fwd_kernel<1, 0><<<grid, 256, 0, as_stream(stream)>>>(a);
  return check_launch("pool_fwd_kernel");
}

extern "C" int mog_pool2d_bwd(const float* x, const unsigned char* argmax, const float* dy, float* dx, int N, int H, int W, int C,
                              int k, int stride, int pad, int mode, void* stream) {
  MOG_REQUIRE(dy && dx && N > 0 && C > 0 && (mode == 0 || mode == 1), "mog_pool2d_bwd: bad argument");
  MOG_REQUIRE(mode == 1 || x || argmax, "mog_pool2d_bwd: max pooling needs the forward input or the recorded arg-max positions");
  int Ho, Wo;
  int rc = mog_pool2d_out_hw(H, W, k, stride, pad, &Ho, &Wo);
  if (rc) return rc;
  const int V = (C & 3) == 0 ? 4 : 1;
  PoolArgs a{x, dy, dx, const_cast<unsigned char*>(argmax), N, H, W, C, Ho, Wo, k, stride, pad, mode, (long long)N * H * W * C / V};
  const unsigned grid = (unsigned)ceil_div_ll(a.total, 256);
  if (V == 4 && k == 3) pool_bwd_kernel<4, 3><<<grid, 256, 0, as_stream(stream)>>>(a);   // <= 3 covering windows per axis (stride >= 1)
  else if (V == 4) pool_bwd_kernel<4, 0><<<grid, 256, 0, as_stream(stream)>>>(a);
  else pool_bwd_kernel<1, 0><<<grid, 256, 0, as_stream(stream)>>>(a);
  return check_launch("pool_bwd_kernel");
}

static float resize_scale(int in, int out, int align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

extern "C" int mog_resize_bilinear_fwd(const float* x, float* y, int N, int Hi, int Wi, int C, int Ho, int Wo, int align_corners,
                                       void* stream) {
  MOG_REQUIRE(x && y && N > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0, "mog_resize_bilinear_fwd: bad argument");
  ResizeArgs a{x, y, N, Hi, Wi, C, Ho, Wo, align_corners ? 1 : 0, resize_scale(Hi, Ho, align_corners), resize_scale(Wi, Wo, align_corners),
               (long long)N * Ho * Wo * C};
  resize_fwd_kernel<<<(unsigned)ceil_div_ll(a.total, 256), 256, 0, as_stream(stream)>>>(a);
  return check_launch("resize_fwd_kernel");
}

extern "C" int mog_resize_bilinear_bwd(const float* dy, float* dx, int N, int Hi, int Wi, int C, int Ho, int Wo, int align_corners,
                                       void* stream) {
  MOG_REQUIRE(dy && dx && N > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0, "mog_resize_bilinear_bwd: bad argument");
  MOG_REQUIRE(Ho >= Hi && Wo >= Wi, "mog_resize_bilinear_bwd: implemented for upsampling (the 256 -> 299 resize of the image encoder)");
  ResizeArgs a{dy, dx, N, Hi, Wi, C, Ho, Wo, align_corners ? 1 : 0, resize_scale(Hi, Ho, align_corners), resize_scale(Wi, Wo, align_corners),
               (long long)N * Hi * Wi * C};
  resize_bwd_kernel<<<(unsigned)ceil_div_ll(a.total, 256), 256, 0, as_stream(stream)>>>(a);
  return check_launch("resize_bwd_kernel");
}
