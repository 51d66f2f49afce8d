// optim.cu -- fused multi-tensor Adam (+ optional EMA of the updated parameters) for the G/D training step.
//
// replaces: torch.optim.Adam(...).step() per network (attngan/trainer.py:141-148,326,340; stackgan/trainer.py:136-137,
// 218,234) and the EMA loop `avg_p.mul_(0.999).add_(0.001, p.data)` (attngan/trainer.py:341-342).
//
// HBM-bound: per parameter element it reads p, g, m, v (16 B) and writes p, m, v (12 B), +8 B for the EMA copy;
// 235.7 M parameters => 6.6 GB per step.  One launch updates up to ADAM_MAX_TENSORS tensors (their pointers travel in
// the kernel parameter block, like torch's multi_tensor_apply, so nothing is staged in device memory); each block
// owns ADAM_CHUNK consecutive elements of one tensor and streams them with 128-bit loads/stores.
// The arithmetic follows torch's non-capturable Adam exactly (lerp form of the first moment, sqrt(v)/sqrt(bc2) + eps,
// addcdiv with step_size = lr / bc1), with the scalars computed in double on the host.
#include "common.cuh"

namespace mog {

constexpr int ADAM_MAX_TENSORS = 56;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_CHUNK = ADAM_THREADS * 4 * 4;   // 4 float4 per thread

struct AdamTensors {
  float* p[ADAM_MAX_TENSORS];
  const float* g[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  float* ema[ADAM_MAX_TENSORS];
  long long numel[ADAM_MAX_TENSORS];
  int block_start[ADAM_MAX_TENSORS + 1];   // first block of each tensor
  int n;
};

struct AdamScalars {
  float w1;          // 1 - beta1 (lerp weight of the gradient)
  float beta2, omb2; // beta2, 1 - beta2
  float bc2_sqrt;    // sqrt(1 - beta2^t)
  float eps;
  float neg_step;    // -lr / (1 - beta1^t)
  float ema_decay, ema_in;   // avg = avg * ema_decay + ema_in * p
  float grad_scale;  // gradients are multiplied by this first (1/world after an all-reduce(sum)); 1 = none
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float* ema, const AdamScalars& s) {
  g *= s.grad_scale;
  // torch lerp: weight < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
  const float d = g - m;
  m = s.w1 < 0.5f ? m + s.w1 * d : g - d * (1.0f - s.w1);
  v = v * s.beta2 + s.omb2 * g * g;
  const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;
  p = p + s.neg_step * (m / denom);
  if (ema) *ema = *ema * s.ema_decay + s.ema_in * p;
}

// Hyper-parameters of the device-step form: the step count lives in device memory (a CUDA graph replays the launch with
// fixed arguments, so nothing that changes from step to step may be a kernel parameter), the bias corrections are formed
// from it in double by one thread per block -- the same arithmetic the host form does.
struct AdamHyper {
  double lr, beta1, beta2;
  const double* step;   // null: use the host-computed scalars
};

__global__ void adam_step_inc_kernel(double* step) { *step += 1.0; }

template <bool VEC>
__global__ void __launch_bounds__(ADAM_THREADS) adam_multi_kernel(const __grid_constant__ AdamTensors T, const __grid_constant__ AdamScalars S0,
                                                                  const __grid_constant__ AdamHyper H) {
  __shared__ AdamScalars sS;
  if (H.step != nullptr) {
    if (threadIdx.x == 0) {
      AdamScalars q = S0;
      const double t = *H.step;
      const double bc1 = 1.0 - pow(H.beta1, t), bc2 = 1.0 - pow(H.beta2, t);
      q.bc2_sqrt = (float)sqrt(bc2);
      q.neg_step = (float)(-H.lr / bc1);
      sS = q;
    }
    __syncthreads();
  }
  const AdamScalars& S = H.step != nullptr ? sS : S0;
  // which tensor does this block belong to (uniform scan over <= 56 entries)
  int t = 0;
  while (t + 1 < T.n && (int)blockIdx.x >= T.block_start[t + 1]) ++t;
  const long long base = (long long)((int)blockIdx.x - T.block_start[t]) * ADAM_CHUNK;
  const long long n = T.numel[t];
  float* __restrict__ p = T.p[t];
  const float* __restrict__ g = T.g[t];
  float* __restrict__ m = T.m[t];
  float* __restrict__ v = T.v[t];
  float* __restrict__ e = T.ema[t];
  if (VEC) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const long long i = base + ((long long)it * ADAM_THREADS + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = *reinterpret_cast<float4*>(p + i);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
        float4 mv = *reinterpret_cast<float4*>(m + i);
        float4 vv = *reinterpret_cast<float4*>(v + i);
        float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e) ev = *reinterpret_cast<float4*>(e + i);
        adam_elem(pv.x, gv.x, mv.x, vv.x, e ? &ev.x : nullptr, S);
        adam_elem(pv.y, gv.y, mv.y, vv.y, e ? &ev.y : nullptr, S);
        adam_elem(pv.z, gv.z, mv.z, vv.z, e ? &ev.z : nullptr, S);
        adam_elem(pv.w, gv.w, mv.w, vv.w, e ? &ev.w : nullptr, S);
        *reinterpret_cast<float4*>(p + i) = pv;
        *reinterpret_cast<float4*>(m + i) = mv;
        *reinterpret_cast<float4*>(v + i) = vv;
        if (e) *reinterpret_cast<float4*>(e + i) = ev;
      } else {
        for (long long j = i; j < n && j < i + 4; ++j) {
          float pj = p[j], mj = m[j], vj = v[j];
          adam_elem(pj, g[j], mj, vj, e ? e + j : nullptr, S);
          p[j] = pj; m[j] = mj; v[j] = vj;
        }
      }
    }
  } else {
    for (int it = 0; it < 16; ++it) {
      const long long j = base + (long long)it * ADAM_THREADS + threadIdx.x;
      if (j < n) {
        float pj = p[j], mj = m[j], vj = v[j];
        adam_elem(pj, g[j], mj, vj, e ? e + j : nullptr, S);
        p[j] = pj; m[j] = mj; v[j] = vj;
      }
    }
  }
}

}  // namespace mog

using namespace mog;

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

static int adam_multi_impl(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* ema,
                           const long long* numel, double lr, double beta1, double beta2, double eps, long long step, double* step_dev,
                           double ema_decay, float grad_scale, void* stream) {
  MOG_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v && numel)), "mog_adam_multi: null array");
  MOG_REQUIRE(step_dev || step >= 1, "mog_adam_multi: step must be >= 1 (the step being taken)");
  MOG_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "mog_adam_multi: bad hyper-parameter");
  cudaStream_t st = as_stream(stream);
  AdamScalars S;
  AdamHyper H;
  H.lr = lr; H.beta1 = beta1; H.beta2 = beta2; H.step = step_dev;
  if (step_dev) {
    step = 1;   // (placeholder for the host-side scalars below; the kernel recomputes them from *step_dev)
    adam_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
    int rc0 = check_launch("adam_step_inc_kernel");
    if (rc0) return rc0;
  }
  // hyper-parameters arrive as doubles (Python floats): 1 - beta must be formed in double like torch does, not from
  // the rounded fp32 beta (1 - float(0.999) is off by 1.3e-5 relative)
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  S.w1 = (float)(1.0 - beta1);
  S.beta2 = (float)beta2;
  S.omb2 = (float)(1.0 - beta2);
  S.bc2_sqrt = (float)sqrt(bc2);
  S.eps = (float)eps;
  S.neg_step = (float)(-lr / bc1);
  S.ema_decay = (float)ema_decay;
  S.ema_in = (float)(1.0 - ema_decay);
  S.grad_scale = grad_scale;
  int i = 0;
  while (i < n) {
    AdamTensors T;
    int k = 0, blocks = 0;
    bool vec = true;
    for (; i < n && k < ADAM_MAX_TENSORS; ++i) {
      if (numel[i] == 0) continue;
      MOG_REQUIRE(numel[i] > 0 && p[i] && g[i] && m[i] && v[i], "mog_adam_multi: tensor %d: null pointer or negative size", i);
      const long long nb = ceil_div_ll(numel[i], ADAM_CHUNK);
      if (nb + blocks > 0x7fffffffLL) break;
      T.p[k] = p[i]; T.g[k] = g[i]; T.m[k] = m[i]; T.v[k] = v[i];
      T.ema[k] = ema ? ema[i] : nullptr;
      T.numel[k] = numel[i];
      T.block_start[k] = blocks;
      vec = vec && aligned16(p[i]) && aligned16(g[i]) && aligned16(m[i]) && aligned16(v[i]) && (!T.ema[k] || aligned16(T.ema[k]));
      blocks += (int)nb;
      ++k;
    }
    if (k == 0) break;
    T.block_start[k] = blocks;
    T.n = k;
    if (vec)
      adam_multi_kernel<true><<<(unsigned)blocks, ADAM_THREADS, 0, st>>>(T, S, H);
    else
      adam_multi_kernel<false><<<(unsigned)blocks, ADAM_THREADS, 0, st>>>(T, S, H);
    int rc = check_launch("adam_multi_kernel");
    if (rc) return rc;
  }
  return MOG_OK;
}

extern "C" int mog_adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* ema,
                              const long long* numel, double lr, double beta1, double beta2, double eps, long long step,
                              double ema_decay, float grad_scale, void* stream) {
  return adam_multi_impl(n, p, g, m, v, ema, numel, lr, beta1, beta2, eps, step, nullptr, ema_decay, grad_scale, stream);
}

extern "C" int mog_adam_multi_dev(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* ema,
                                  const long long* numel, double lr, double beta1, double beta2, double eps, double* step_dev,
                                  double ema_decay, float grad_scale, void* stream) {
  MOG_REQUIRE(step_dev != nullptr, "mog_adam_multi_dev: null step counter");
  return adam_multi_impl(n, p, g, m, v, ema, numel, lr, beta1, beta2, eps, 0, step_dev, ema_decay, grad_scale, stream);
}
