// stn.cu -- spatial transformer for the object pathway: affine grid + bilinear sampling with
// zero padding (model.py:17-21), fused over the per-object loops of the reference:
//   mode 0 "scatter-sum": y[b] = sum_s stn(x[s*B+b], theta[b,s])      (model.py:393-401, 107-112, 691-693)
//   mode 1 "crop":        y[s*B+b] = [stn(x[b], theta[b,s]) | extra[b,s,:]]   (model.py:686-689)
// The backward is a deterministic gather (no atomics): theta produced by the reference is
// axis-aligned (miscc/utils.py:28-29), so the sampling weights separate into a row table and
// a column table per (b, s), staged in shared memory.
// Coordinates reproduce torch's own arithmetic (linspace base grid scaled by (W-1)/W for
// align_corners=False, then ((g+1)*size-1)/2) so fp32 results agree to rounding.
#include "common.cuh"

namespace mog {

__device__ __forceinline__ float base_coord(int j, int W, int align) {
  if (W <= 1) return 0.f;  // linspace(-1,1,1) = -1, scaled by (W-1)/W = 0; align: single point -1 -> torch gives -1
  const float step = 2.0f / (float)(W - 1);
  float v = (j < W / 2) ? (-1.0f + step * (float)j) : (1.0f - step * (float)(W - 1 - j));
  if (!align) v = v * (float)(W - 1) / (float)W;
  return v;
}
__device__ __forceinline__ float unnormalize(float g, int size, int align) {
  return align ? ((g + 1.f) / 2.f) * (float)(size - 1) : ((g + 1.f) * (float)size - 1.f) / 2.f;
}

struct StnArgs {
  const float* x; const float* theta; const float* extra; float* y;
  int mode, B, S, Hi, Wi, C, Ho, Wo, Cy, align;
};

// forward: one thread per (output pixel, channel); channels fastest => coalesced NHWC
__global__ void stn_fwd_kernel(StnArgs a) {
  const size_t n_out_img = a.mode == 0 ? (size_t)a.B : (size_t)a.S * a.B;
  const size_t total = n_out_img * a.Ho * a.Wo * a.Cy;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % a.Cy);
  size_t q = idx / a.Cy;
  const int wo = (int)(q % a.Wo);
  q /= a.Wo;
  const int ho = (int)(q % a.Ho);
  const int img = (int)(q / a.Ho);
  const float xo = base_coord(wo, a.Wo, a.align), yo = base_coord(ho, a.Ho, a.align);
  float acc = 0.f;
  const int s_begin = a.mode == 0 ? 0 : img / a.B;
  const int s_end = a.mode == 0 ? a.S : s_begin + 1;
  const int b = a.mode == 0 ? img : img % a.B;
  if (a.mode == 1 && c >= a.C) {
    a.y[idx] = a.extra ? a.extra[((size_t)b * a.S + s_begin) * (a.Cy - a.C) + (c - a.C)] : 0.f;
    return;
  }
  for (int s = s_begin; s < s_end; ++s) {
    const float* th = a.theta + ((size_t)b * a.S + s) * 6;
    const float xs = fmaf(th[0], xo, fmaf(th[1], yo, th[2]));
    const float ys = fmaf(th[4], yo, fmaf(th[3], xo, th[5]));
    const float ix = unnormalize(xs, a.Wi, a.align), iy = unnormalize(ys, a.Hi, a.align);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    // reject far-out coordinates before the int conversion (empty slots map to ~ -40)
    if (!(fx0 >= -1.f && fx0 < (float)a.Wi && fy0 >= -1.f && fy0 < (float)a.Hi)) continue;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float tx = ix - fx0, ty = iy - fy0;
    const int src_img = a.mode == 0 ? s * a.B + b : b;
    const float* xp = a.x + (size_t)src_img * a.Hi * a.Wi * a.C + c;
    auto at = [&](int yy, int xx) -> float {
      return (yy >= 0 && yy < a.Hi && xx >= 0 && xx < a.Wi) ? __ldg(xp + ((size_t)yy * a.Wi + xx) * a.C) : 0.f;
    };
    // same association as torch's grid_sampler: nw*w_nw + ne*w_ne + sw*w_sw + se*w_se
    const float w_nw = (1.f - tx) * (1.f - ty), w_ne = tx * (1.f - ty), w_sw = (1.f - tx) * ty, w_se = tx * ty;
    acc += at(y0, x0) * w_nw + at(y0, x0 + 1) * w_ne + at(y0 + 1, x0) * w_sw + at(y0 + 1, x0 + 1) * w_se;
  }
  a.y[idx] = acc;
}

// backward gather.  grid: (pixel tiles, source images); block: 256 threads.
// smem tables per (b, s): x0[Wo], tx[Wo], y0[Ho], ty[Ho]
struct StnBwdArgs {
  const float* dy; const float* theta; float* dx;
  int mode, B, S, Hi, Wi, C, Ho, Wo, Cy, align;
};

__global__ void stn_bwd_kernel(StnBwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // the source image this block produces gradients for
  const int img = blockIdx.y;            // mode 0: s*B + b ; mode 1: b
  const int b = a.mode == 0 ? img % a.B : img;
  const int s_begin = a.mode == 0 ? img / a.B : 0;
  const int s_end = a.mode == 0 ? s_begin + 1 : a.S;
  const int ns = s_end - s_begin;
  int* tx0 = reinterpret_cast<int*>(smem_raw);                 // [ns][Wo]
  float* ttx = reinterpret_cast<float*>(tx0 + ns * a.Wo);      // [ns][Wo]
  int* ty0 = reinterpret_cast<int*>(ttx + ns * a.Wo);          // [ns][Ho]
  float* tty = reinterpret_cast<float*>(ty0 + ns * a.Ho);      // [ns][Ho]
  for (int i = threadIdx.x; i < ns * a.Wo; i += blockDim.x) {
    int s = s_begin + i / a.Wo, wo = i % a.Wo;
    const float* th = a.theta + ((size_t)b * a.S + s) * 6;
    float ix = unnormalize(fmaf(th[0], base_coord(wo, a.Wo, a.align), th[2]), a.Wi, a.align);
    float f = floorf(ix);
    bool ok = f >= -1.f && f < (float)a.Wi;
    tx0[i] = ok ? (int)f : -1000000;
    ttx[i] = ix - f;
  }
  for (int i = threadIdx.x; i < ns * a.Ho; i += blockDim.x) {
    int s = s_begin + i / a.Ho, ho = i % a.Ho;
    const float* th = a.theta + ((size_t)b * a.S + s) * 6;
    float iy = unnormalize(fmaf(th[4], base_coord(ho, a.Ho, a.align), th[5]), a.Hi, a.align);
    float f = floorf(iy);
    bool ok = f >= -1.f && f < (float)a.Hi;
    ty0[i] = ok ? (int)f : -1000000;
    tty[i] = iy - f;
  }
  __syncthreads();
  const size_t per_img = (size_t)a.Hi * a.Wi * a.C;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < per_img; e += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % a.C);
    const int wi = (int)((e / a.C) % a.Wi);
    const int hi = (int)(e / ((size_t)a.C * a.Wi));
    float acc = 0.f;
    for (int si = 0; si < ns; ++si) {
      const int s = s_begin + si;
      const int dy_img = a.mode == 0 ? b : s * a.B + b;
      const float* dyp = a.dy + (size_t)dy_img * a.Ho * a.Wo * a.Cy + c;
      for (int ho = 0; ho < a.Ho; ++ho) {
        const int y0 = ty0[si * a.Ho + ho];
        float wy;
        if (y0 == hi) wy = 1.f - tty[si * a.Ho + ho];
        else if (y0 + 1 == hi) wy = tty[si * a.Ho + ho];
        else continue;
        for (int wo = 0; wo < a.Wo; ++wo) {
          const int x0 = tx0[si * a.Wo + wo];
          float wx;
          if (x0 == wi) wx = 1.f - ttx[si * a.Wo + wo];
          else if (x0 + 1 == wi) wx = ttx[si * a.Wo + wo];
          else continue;
          acc = fmaf(__ldg(dyp + ((size_t)ho * a.Wo + wo) * a.Cy), wx * wy, acc);
        }
      }
    }
    a.dx[(size_t)img * per_img + e] = acc;
  }
}

}  // namespace mog

using namespace mog;

extern "C" int mog_stn_fwd(const float* x, const float* theta, const float* extra, float* y, int mode, int B, int S,
                           int Hi, int Wi, int C, int Ho, int Wo, int Cy, int align_corners, void* stream) {
  MOG_REQUIRE(x && theta && y, "mog_stn_fwd: null tensor");
  MOG_REQUIRE(mode == 0 || mode == 1, "mog_stn_fwd: mode must be 0 (scatter-sum) or 1 (crop)");
  MOG_REQUIRE(B > 0 && S > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0, "mog_stn_fwd: non-positive dims");
  MOG_REQUIRE(Cy >= C && (mode == 1 || Cy == C), "mog_stn_fwd: Cy must equal C (scatter) or be >= C (crop)");
  StnArgs a{x, theta, extra, y, mode, B, S, Hi, Wi, C, Ho, Wo, Cy, align_corners};
  size_t total = (size_t)(mode == 0 ? B : S * B) * Ho * Wo * Cy;
  stn_fwd_kernel<<<(unsigned)ceil_div_ll((long long)total, 256), 256, 0, as_stream(stream)>>>(a);
  return check_launch("stn_fwd_kernel");
}

extern "C" int mog_stn_bwd(const float* dy, const float* theta, float* dx, int mode, int B, int S, int Hi, int Wi,
                           int C, int Ho, int Wo, int Cy, int align_corners, void* stream) {
  MOG_REQUIRE(dy && theta && dx, "mog_stn_bwd: null tensor");
  MOG_REQUIRE(mode == 0 || mode == 1, "mog_stn_bwd: mode must be 0 (scatter-sum) or 1 (crop)");
  MOG_REQUIRE(B > 0 && S > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0 && Cy >= C, "mog_stn_bwd: bad dims");
  StnBwdArgs a{dy, theta, dx, mode, B, S, Hi, Wi, C, Ho, Wo, Cy, align_corners};
  const int ns = mode == 0 ? 1 : S;
  size_t smem = (size_t)ns * (Wo + Ho) * 8;
  MOG_REQUIRE(smem <= 48 * 1024, "mog_stn_bwd: output grid too large for the weight tables");
  size_t per_img = (size_t)Hi * Wi * C;
  unsigned gx = (unsigned)ceil_div_ll((long long)per_img, 256);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, mode == 0 ? S * B : B);
  stn_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(a);
  return check_launch("stn_bwd_kernel");
}
