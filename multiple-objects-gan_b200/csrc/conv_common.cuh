// Shared between the CUDA-core and the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace mog {

// One gather-GEMM problem:  dst[row, :] = act(sum_{tap, c} src[pix(row, tap), c] * wmat[tapw[tap]*Cs + c, :] + bias)
// rows enumerate (n, rh, rw) over an Hr x Wr grid; source pixel = (rh*rs + off_h[th], rw*rs + off_w[tw])
// in the logical (optionally 2x nearest-upsampled) source; destination pixel = (rh*dsh + doh, rw*dsw + dow).
// The forward conv is one such problem; the data gradient is one per stride phase.
struct IGemmParams {
  const float* src;
  const float* wmat;
  const float* bias;
  float* dst;
  int N, Hs, Ws, Cs;
  int up2x;
  int Hr, Wr, rs;
  int nth, ntw;
  int off_h[8], off_w[8];
  int tapw[64];
  int Cd;
  int Hd, Wd, dsh, doh, dsw, dow;
  int act;
  long long M;
  int K;
};

int launch_igemm_ffma(const IGemmParams& p, cudaStream_t st);
int launch_wgrad_ffma(const MogConvDesc& d, int Ho, int Wo, const float* x, const float* dy, float* dw,
                      float* ws, cudaStream_t st);
int wgrad_splits(const MogConvDesc& d, int Ho, int Wo);
int launch_colsum(const float* x, float* out, long long M, int C, cudaStream_t st);
int launch_sumpool(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);

}  // namespace mog
