// Shared between the CUDA-core and the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace mog {

// One gather-GEMM problem:  dst[row, :] = act(sum_{tap, c} src[pix(row, tap), c] * wmat[tapw[tap]*Cs + c, :] + bias)
// rows enumerate (n, rh, rw) over an Hr x Wr grid; source pixel = (rh*rs + off_h[th], rw*rs + off_w[tw])
// in the logical (optionally 2x nearest-upsampled) source; destination pixel = (rh*dsh + doh, rw*dsw + dow).
// The forward conv is one such problem; the data gradient is one per stride phase.
struct IGemmParams {
  const float* src;            // fp32 NHWC source (CUDA-core path, or tcgen05 path without planes)
  const void* src_planes;      // tcgen05 path: bf16 hi plane [pixels][Cs] followed by the lo plane, or nullptr
  size_t src_plane_elems;      // elements per plane
  const float* wmat;
  const float* bias;
  float* dst;
  int N, Hs, Ws, Cs;           // logical source grid (a strided *view* of the physical tensor when vstep > 1)
  int vstep, voh, vow, Hp, Wp; // view: logical pixel (h, w) = physical (h*vstep + voh, w*vstep + vow) of an Hp x Wp tensor;
                               // vstep == 0 means "no view" (physical == logical)
  int accum_dst;               // epilogue adds to dst instead of overwriting (sums over parity views)
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
int launch_wgrad_ffma_partial(const MogConvDesc& d, int Ho, int Wo, const float* x, const float* dy, float* ws,
                              int* splits_out, cudaStream_t st);
size_t wgrad_ffma_workspace_bytes(const MogConvDesc& d, int Ho, int Wo);
int launch_wgrad_reduce(const float* ws, float* dw, int splits, int K, int Cout, int Cin, int CinP, int KHW,
                        cudaStream_t st);
int launch_colsum(const float* x, float* out, long long M, int C, cudaStream_t st);
int launch_sumpool(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);

// tcgen05 path (conv_tc.cu, conv_tc_wgrad.cu)
namespace tc { struct TcWeightLayout; }
int tc_bn_for(int Cd);
size_t tc_packed_bytes(int ntaps, int Cs, int Cd, int passes);
int launch_igemm_tc(const IGemmParams& g, const void* packed, int passes, void* workspace, size_t ws_bytes,
                    cudaStream_t st);
size_t tc_igemm_workspace_bytes(long long M, int ntaps, int Cs, int Cd, int passes);
bool tc_wgrad_eligible(const MogConvDesc& d, bool planes);
bool patch_dgrad_eligible(const MogConvDesc& d);
int launch_patch_dgrad(const MogConvDesc& d, const float* dy, const void* packed, float* dx, int passes, cudaStream_t st);
// halo-tile TMA persistent kernel (conv_halo.cu)
bool halo_shape_eligible(const IGemmParams& g);
int halo_tap_pitch(int Cs);
int launch_igemm_halo(const IGemmParams* gs, int n, const void* const* packed, int passes, void* workspace, size_t ws_bytes,
                      cudaStream_t st);
size_t halo_workspace_bytes(const IGemmParams* gs, int n, int passes);
int launch_splitk_reduce(const float* partial, const IGemmParams& g, int splits, cudaStream_t st);
int tc_pack_pitch(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int transpose, int ntaps,
                  const int (*taps)[4], int pitch, int passes, cudaStream_t st);
int tc_pack_entry(const float* w_oihw, void* out, int Cout, int Cin, int KH, int KW, int transpose, int ntaps,
                  const int (*taps)[4], int pitch, int passes, MogPackEntry* e);
int launch_pack_multi(const MogPackEntry* entries_dev, const MogPackGroup* groups_dev, int ngroups, int total_blocks, cudaStream_t st);
size_t tc_wgrad_workspace_bytes(const MogConvDesc& d, int Ho, int Wo);
int launch_wgrad_tc(const MogConvDesc& d, int Ho, int Wo, const float* x, const float* dy, const void* x_planes,
                    size_t x_plane_elems, const void* dy_planes, size_t dy_plane_elems, float* ws, int passes,
                    int* splits_out, cudaStream_t st);
// halo-tile weight gradient (wgrad_halo.cu)
bool wgrad_halo_eligible(const MogConvDesc& d, int Ho, int Wo, int passes);
size_t wgrad_halo_workspace_bytes(const MogConvDesc& d, int Ho, int Wo, int passes);
int launch_wgrad_halo(const MogConvDesc& d, int Ho, int Wo, const void* x_planes, const void* dy_planes, float* dw, float* ws,
                      int passes, cudaStream_t st);
int launch_patch_planes(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo,
                        void* planes, int nplanes, cudaStream_t st);
int launch_col2im_act(const float* z, int ldz, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo,
                      const float* bias, int act, float* y, cudaStream_t st);
int launch_split_planes(const float* x, long long rows, int C, int CP, void* planes, int nplanes, cudaStream_t st,
                        const float* y = nullptr, int act = 0);

}  // namespace mog
