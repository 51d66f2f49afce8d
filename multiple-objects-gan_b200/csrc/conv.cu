// conv.cu -- C-ABI entry points of the convolution family; builds the gather-GEMM problems and
// dispatches on MogConvDesc.precision.
#include "conv_common.cuh"

using namespace mog;

static int validate(const MogConvDesc* d, const char* who) {
  MOG_REQUIRE(d, "%s: null descriptor", who);
  MOG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: non-positive dims", who);
  MOG_REQUIRE(d->KH > 0 && d->KW > 0 && d->KH <= 8 && d->KW <= 8 && d->KH * d->KW <= 64, "%s: filter %dx%d unsupported", who, d->KH, d->KW);
  MOG_REQUIRE(d->stride >= 1 && d->pad >= 0, "%s: bad stride/pad", who);
  MOG_REQUIRE(d->up2x == 0 || d->up2x == 1, "%s: up2x must be 0/1", who);
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  MOG_REQUIRE(HL + 2 * d->pad >= d->KH && WL + 2 * d->pad >= d->KW, "%s: filter larger than padded input", who);
  return MOG_OK;
}

extern "C" int mog_conv_out_hw(const MogConvDesc* d, int* Ho, int* Wo) {
  int rc = validate(d, "mog_conv_out_hw");
  if (rc) return rc;
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  if (Ho) *Ho = (HL + 2 * d->pad - d->KH) / d->stride + 1;
  if (Wo) *Wo = (WL + 2 * d->pad - d->KW) / d->stride + 1;
  return MOG_OK;
}

extern "C" size_t mog_conv_workspace_bytes(const MogConvDesc* d, int which) {
  if (validate(d, "mog_conv_workspace_bytes")) return 0;
  int Ho, Wo;
  mog_conv_out_hw(d, &Ho, &Wo);
  if (which == 1) return d->up2x ? (size_t)d->N * (2 * d->H) * (2 * d->W) * d->Cin * sizeof(float) : 0;
  if (which == 2) return (size_t)wgrad_splits(*d, Ho, Wo) * d->KH * d->KW * d->Cin * d->Cout * sizeof(float);
  return 0;
}

extern "C" int mog_conv2d_fwd(const MogConvDesc* d, const float* x, const float* w, const float* bias, float* y,
                              void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_fwd");
  if (rc) return rc;
  MOG_REQUIRE(x && w && y, "mog_conv2d_fwd: null tensor");
  (void)workspace; (void)ws_bytes;
  int Ho, Wo;
  mog_conv_out_hw(d, &Ho, &Wo);
  IGemmParams p{};
  p.src = x; p.wmat = w; p.bias = bias; p.dst = y;
  p.N = d->N; p.Hs = d->H; p.Ws = d->W; p.Cs = d->Cin; p.up2x = d->up2x;
  p.Hr = Ho; p.Wr = Wo; p.rs = d->stride;
  p.nth = d->KH; p.ntw = d->KW;
  for (int i = 0; i < d->KH; ++i) p.off_h[i] = i - d->pad;
  for (int i = 0; i < d->KW; ++i) p.off_w[i] = i - d->pad;
  for (int i = 0; i < d->KH * d->KW; ++i) p.tapw[i] = i;
  p.Cd = d->Cout; p.Hd = Ho; p.Wd = Wo; p.dsh = 1; p.doh = 0; p.dsw = 1; p.dow = 0;
  p.act = d->act;
  p.M = (long long)d->N * Ho * Wo;
  p.K = d->KH * d->KW * d->Cin;
  if (d->precision != MOG_PREC_FP32) return fail(MOG_ERR_UNSUPPORTED, "mog_conv2d_fwd: precision %d not built", d->precision);
  return launch_igemm_ffma(p, as_stream(stream));
}

extern "C" int mog_conv2d_dgrad(const MogConvDesc* d, const float* dy, const float* wt, float* dx, void* workspace,
                                size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_dgrad");
  if (rc) return rc;
  MOG_REQUIRE(dy && wt && dx, "mog_conv2d_dgrad: null tensor");
  if (d->precision != MOG_PREC_FP32) return fail(MOG_ERR_UNSUPPORTED, "mog_conv2d_dgrad: precision %d not built", d->precision);
  int Ho, Wo;
  mog_conv_out_hw(d, &Ho, &Wo);
  const int HL = d->H << d->up2x, WL = d->W << d->up2x;  // logical input grid of the conv
  float* target = dx;
  if (d->up2x) {
    size_t need = mog_conv_workspace_bytes(d, 1);
    if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_dgrad: workspace %zu < %zu", ws_bytes, need);
    target = static_cast<float*>(workspace);
  }
  const int s = d->stride;
  cudaStream_t st = as_stream(stream);
  // one gather-GEMM per stride phase: input pixels (hi, wi) with hi%s==ph, wi%s==pw only see the
  // taps kh with (ph + pad - kh) % s == 0, at dy row hi/s + (ph + pad - kh)/s.
  for (int ph = 0; ph < s; ++ph) {
    for (int pw = 0; pw < s; ++pw) {
      IGemmParams p{};
      p.src = dy; p.wmat = wt; p.bias = nullptr; p.dst = target;
      p.N = d->N; p.Hs = Ho; p.Ws = Wo; p.Cs = d->Cout; p.up2x = 0;
      p.Hr = (HL - ph + s - 1) / s; p.Wr = (WL - pw + s - 1) / s; p.rs = 1;
      if (p.Hr <= 0 || p.Wr <= 0) continue;
      int nth = 0, ntw = 0, khs[8], kws[8];
      for (int kh = 0; kh < d->KH; ++kh)
        if (((ph + d->pad - kh) % s + s) % s == 0) { khs[nth] = kh; p.off_h[nth] = (ph + d->pad - kh) / s; ++nth; }
      for (int kw = 0; kw < d->KW; ++kw)
        if (((pw + d->pad - kw) % s + s) % s == 0) { kws[ntw] = kw; p.off_w[ntw] = (pw + d->pad - kw) / s; ++ntw; }
      p.nth = nth; p.ntw = ntw;
      for (int a = 0; a < nth; ++a)
        for (int b = 0; b < ntw; ++b) p.tapw[a * ntw + b] = khs[a] * d->KW + kws[b];
      p.Cd = d->Cin; p.Hd = HL; p.Wd = WL; p.dsh = s; p.doh = ph; p.dsw = s; p.dow = pw;
      p.act = MOG_ACT_NONE;
      p.M = (long long)d->N * p.Hr * p.Wr;
      p.K = nth * ntw * d->Cout;
      if (p.K == 0) {
        // no tap reaches this phase: gradient is zero there. Handled by a K=0 GEMM (writes zeros).
        p.nth = 0; p.ntw = 1;
      }
      rc = launch_igemm_ffma(p, st);
      if (rc) return rc;
    }
  }
  if (d->up2x) return launch_sumpool(target, dx, d->N, d->H, d->W, d->Cin, st);
  return MOG_OK;
}

extern "C" int mog_conv2d_wgrad(const MogConvDesc* d, const float* x, const float* dy, float* dw, float* dbias,
                                void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_wgrad");
  if (rc) return rc;
  MOG_REQUIRE(x && dy && dw, "mog_conv2d_wgrad: null tensor");
  if (d->precision != MOG_PREC_FP32) return fail(MOG_ERR_UNSUPPORTED, "mog_conv2d_wgrad: precision %d not built", d->precision);
  int Ho, Wo;
  mog_conv_out_hw(d, &Ho, &Wo);
  size_t need = mog_conv_workspace_bytes(d, 2);
  if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_wgrad: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  rc = launch_wgrad_ffma(*d, Ho, Wo, x, dy, dw, static_cast<float*>(workspace), st);
  if (rc) return rc;
  if (dbias) return launch_colsum(dy, dbias, (long long)d->N * Ho * Wo, d->Cout, st);
  return MOG_OK;
}
