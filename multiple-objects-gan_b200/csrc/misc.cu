// misc.cu -- layout conversion at the NCHW API boundary and the fused sigmoid + BCE loss head.
#include "common.cuh"

namespace mog {

// tiled transpose of the (C, H*W) plane of every sample: NCHW <-> NHWC
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // src: [N][rows][cols] -> dst: [N][cols][rows]
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[base + (size_t)r * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) dst[base + (size_t)c2 * rows + r2] = tile[threadIdx.x][j];
  }
}

static int launch_transpose(const float* src, float* dst, int N, int rows, int cols, cudaStream_t st) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), N);
  if (grid.y > 65535 || grid.z > 65535) return fail(MOG_ERR_UNSUPPORTED, "transpose: grid too large");
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, rows, cols);
  return check_launch("transpose_kernel");
}

// BCE(sigmoid(z), t) = -(t*max(log p, -100) + (1-t)*max(log(1-p), -100))   (torch.nn.BCELoss)
// with_logits: BCEWithLogitsLoss, max(z,0) - z*t + log1p(exp(-|z|)) (stackgan/miscc/utils.py:77)
__global__ void sigmoid_bce_fwd_kernel(const float* __restrict__ z, const float* __restrict__ tgt, float weight, int n,
                                       float* prob, float* loss, int accumulate, int with_logits) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float p = sigmoidf_(z[i]);
    if (prob) prob[i] = p;
    const float target = tgt[i];
    if (with_logits) {
      const float zz = z[i];
      s += fmaxf(zz, 0.f) - zz * target + log1pf(expf(-fabsf(zz)));
    } else {
      float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
      s -= target * lp + (1.f - target) * lq;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      float r = weight * v / (float)n;
      loss[0] = accumulate ? loss[0] + r : r;
    }
  }
}

__global__ void sigmoid_bce_bwd_kernel(const float* __restrict__ z, const float* __restrict__ tgt, float weight, int n,
                                       const float* __restrict__ gscale, float* __restrict__ dz, int with_logits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p = sigmoidf_(z[i]);
  // torch: grad_p = (p - t) / max((1-p)*p, 1e-12); sigmoid': p*(1-p)
  float pq = (1.f - p) * p;
  float g = with_logits ? (p - tgt[i]) : (p - tgt[i]) / fmaxf(pq, 1e-12f) * pq;
  dz[i] = g * weight * (gscale ? gscale[0] : 1.f) / (float)n;
}

}  // namespace mog

using namespace mog;

extern "C" int mog_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
  MOG_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "mog_nchw_to_nhwc: bad argument");
  return launch_transpose(src, dst, N, C, H * W, as_stream(stream));
}
extern "C" int mog_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
  MOG_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "mog_nhwc_to_nchw: bad argument");
  return launch_transpose(src, dst, N, H * W, C, as_stream(stream));
}

extern "C" int mog_sigmoid_bce_fwd(const float* z, const float* target, float weight, int n, float* prob, float* loss,
                                   int accumulate, int with_logits, void* stream) {
  MOG_REQUIRE(z && target && loss && n > 0, "mog_sigmoid_bce_fwd: bad argument");
  sigmoid_bce_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(z, target, weight, n, prob, loss, accumulate, with_logits);
  return check_launch("sigmoid_bce_fwd_kernel");
}
extern "C" int mog_sigmoid_bce_bwd(const float* z, const float* target, float weight, int n, const float* gscale,
                                   float* dz, int with_logits, void* stream) {
  MOG_REQUIRE(z && target && dz && n > 0, "mog_sigmoid_bce_bwd: bad argument");
  sigmoid_bce_bwd_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(z, target, weight, n, gscale, dz, with_logits);
  return check_launch("sigmoid_bce_bwd_kernel");
}
