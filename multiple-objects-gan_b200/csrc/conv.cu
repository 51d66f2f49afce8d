// conv.cu -- C-ABI entry points of the convolution family.  Builds the gather-GEMM problems
// (forward: one; data gradient: one per stride phase) and dispatches on MogConvDesc.precision:
//   MOG_PREC_FP32              -> CUDA-core kernels (conv_ffma.cu), fp32 operands
//   MOG_PREC_BF16X3 / _BF16    -> tcgen05 kernels (conv_tc.cu, conv_tc_wgrad.cu), always.  The
//       gathered operands are either pre-split bf16 planes (mog_split_planes; channel pitch rounded
//       up to 8, so every channel count is accepted) or fp32 tensors split on the fly (only when
//       the channel count is a multiple of 8).
#include "conv_common.cuh"

using namespace mog;

namespace mog {
// conv_ffma.cu
__global__ void pack_fwd_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW);
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW);
}  // namespace mog

static int validate(const MogConvDesc* d, const char* who) {
  MOG_REQUIRE(d, "%s: null descriptor", who);
  MOG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: non-positive dims", who);
  MOG_REQUIRE(d->KH > 0 && d->KW > 0 && d->KH <= 8 && d->KW <= 8 && d->KH * d->KW <= 64, "%s: filter %dx%d unsupported", who, d->KH, d->KW);
  MOG_REQUIRE(d->stride >= 1 && d->stride <= 8 && d->pad >= 0, "%s: bad stride/pad", who);
  MOG_REQUIRE(d->up2x == 0 || d->up2x == 1, "%s: up2x must be 0/1", who);
  MOG_REQUIRE(d->precision >= MOG_PREC_FP32 && d->precision <= MOG_PREC_BF16, "%s: unknown precision %d", who, d->precision);
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  MOG_REQUIRE(HL + 2 * d->pad >= d->KH && WL + 2 * d->pad >= d->KW, "%s: filter larger than padded input", who);
  return MOG_OK;
}

static void out_hw(const MogConvDesc* d, int* Ho, int* Wo) {
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  *Ho = (HL + 2 * d->pad - d->KH) / d->stride + 1;
  *Wo = (WL + 2 * d->pad - d->KW) / d->stride + 1;
}

static int passes_of(const MogConvDesc* d) { return d->precision == MOG_PREC_BF16X3 ? 3 : 1; }
static bool use_tc(const MogConvDesc* d) { return d->precision != MOG_PREC_FP32; }
static int p8(int c) { return ceil_div(c, 8) * 8; }

// ---- problem builders -----------------------------------------------------------------------
static IGemmParams fwd_problem(const MogConvDesc* d) {
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  IGemmParams p{};
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
  return p;
}

// One gather-GEMM per stride phase: input pixels (hi, wi) with hi%s==ph, wi%s==pw only see the taps
// kh with (ph + pad - kh) % s == 0, at dy row hi/s + (ph + pad - kh)/s.  Returns false for an
// empty phase (no rows).
static bool dgrad_problem(const MogConvDesc* d, int ph, int pw, IGemmParams* out) {
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  const int HL = d->H << d->up2x, WL = d->W << d->up2x;
  const int s = d->stride;
  IGemmParams p{};
  p.N = d->N; p.Hs = Ho; p.Ws = Wo; p.Cs = d->Cout; p.up2x = 0;
  p.Hr = (HL - ph + s - 1) / s; p.Wr = (WL - pw + s - 1) / s; p.rs = 1;
  if (p.Hr <= 0 || p.Wr <= 0) return false;
  int nth = 0, ntw = 0, khs[8], kws[8];
  for (int kh = 0; kh < d->KH; ++kh)
    if (((ph + d->pad - kh) % s + s) % s == 0) { khs[nth] = kh; p.off_h[nth] = (ph + d->pad - kh) / s; ++nth; }
  for (int kw = 0; kw < d->KW; ++kw)
    if (((pw + d->pad - kw) % s + s) % s == 0) { kws[ntw] = kw; p.off_w[ntw] = (pw + d->pad - kw) / s; ++ntw; }
  if (nth == 0 || ntw == 0) { nth = 0; ntw = 1; }  // no tap reaches this phase: K = 0, zeros are written
  p.nth = nth; p.ntw = ntw;
  for (int a = 0; a < nth; ++a)
    for (int b = 0; b < ntw; ++b) p.tapw[a * ntw + b] = khs[a] * d->KW + kws[b];
  p.Cd = d->Cin; p.Hd = HL; p.Wd = WL; p.dsh = s; p.doh = ph; p.dsw = s; p.dow = pw;
  p.act = MOG_ACT_NONE;
  p.M = (long long)d->N * p.Hr * p.Wr;
  p.K = nth * ntw * d->Cout;
  *out = p;
  return true;
}

// tcgen05 path: k runs over (tap, channel); the channel pitch per tap is Cs rounded up to 8 for the
// generic gather kernel and to 64 for the TMA kernel (whole 128-byte swizzle rows per box).
static int tap_pitch(const IGemmParams& shape, int CsReal) { return tma_shape_eligible(shape) ? tma_tap_pitch(CsReal) : p8(CsReal); }
static size_t tc_bytes(const IGemmParams& shape, int CsReal, int passes) {
  return tc_packed_bytes(shape.nth * shape.ntw, tap_pitch(shape, CsReal), shape.Cd, passes);
}

// dgrad workspace = [hi-res gradient of the fused upsample][split-K partials of the largest phase]
static size_t dgrad_up_bytes(const MogConvDesc* d) {
  size_t b = d->up2x ? (size_t)d->N * (2 * d->H) * (2 * d->W) * d->Cin * sizeof(float) : 0;
  return (b + 255) / 256 * 256;
}
static size_t dgrad_split_bytes(const MogConvDesc* d) {
  if (!use_tc(d)) return 0;
  size_t mx = 0;
  for (int ph = 0; ph < d->stride; ++ph)
    for (int pw = 0; pw < d->stride; ++pw) {
      IGemmParams p;
      if (!dgrad_problem(d, ph, pw, &p)) continue;
      if (tma_shape_eligible(p)) continue;
      size_t b = tc_igemm_workspace_bytes(p.M, p.nth * p.ntw, p8(d->Cout), d->Cin, passes_of(d));
      if (b > mx) mx = b;
    }
  return mx;
}

// attach the gathered operand (planes or fp32) of a tcgen05 problem; Cs becomes the channel pitch
static int attach_source(IGemmParams* p, const float* src_f32, const void* planes, long long pixels, const char* who) {
  const int CsReal = p->Cs;
  if (planes) {
    p->Cs = p8(CsReal);
    p->src = nullptr;
    p->src_planes = planes;
    p->src_plane_elems = (size_t)pixels * p->Cs;
  } else {
    if (CsReal % 8) return fail(MOG_ERR_UNSUPPORTED, "%s: %d channels need pre-split planes (mog_split_planes) in tcgen05 precision", who, CsReal);
    if (!src_f32) return fail(MOG_ERR_BAD_ARG, "%s: neither an fp32 tensor nor planes given", who);
    p->src = src_f32;
    p->src_planes = nullptr;
    p->src_plane_elems = 0;
  }
  p->K = p->nth * p->ntw * p->Cs;
  return MOG_OK;
}

// ---- public API ---------------------------------------------------------------------------------
extern "C" int mog_conv_out_hw(const MogConvDesc* d, int* Ho, int* Wo) {
  int rc = validate(d, "mog_conv_out_hw");
  if (rc) return rc;
  int a, b;
  out_hw(d, &a, &b);
  if (Ho) *Ho = a;
  if (Wo) *Wo = b;
  return MOG_OK;
}

extern "C" size_t mog_planes_bytes(long long rows, int C, int precision) {
  if (rows <= 0 || C <= 0 || precision == MOG_PREC_FP32) return 0;
  return (size_t)rows * p8(C) * 2 * (precision == MOG_PREC_BF16X3 ? 2 : 1);
}

extern "C" int mog_split_planes(const float* x, long long rows, int C, int precision, void* planes, void* stream) {
  MOG_REQUIRE(x && planes && rows > 0 && C > 0, "mog_split_planes: bad argument");
  MOG_REQUIRE(precision == MOG_PREC_BF16X3 || precision == MOG_PREC_BF16, "mog_split_planes: precision must be a tcgen05 mode");
  return launch_split_planes(x, rows, C, p8(C), planes, precision == MOG_PREC_BF16X3 ? 2 : 1, as_stream(stream));
}

extern "C" size_t mog_packed_weight_bytes(const MogConvDesc* d, int which) {
  if (validate(d, "mog_packed_weight_bytes")) return 0;
  const size_t dense = (size_t)d->KH * d->KW * d->Cin * d->Cout * sizeof(float);
  if (!use_tc(d)) return dense;
  if (which == 0) return tc_bytes(fwd_problem(d), d->Cin, passes_of(d));
  if (which == 1) {
    size_t tot = 0;
    for (int ph = 0; ph < d->stride; ++ph)
      for (int pw = 0; pw < d->stride; ++pw) {
        IGemmParams p;
        if (!dgrad_problem(d, ph, pw, &p)) continue;
        tot += tc_bytes(p, d->Cout, passes_of(d));
      }
    return tot;
  }
  return 0;
}

// Identifies the packed layout chosen for (d, which) so callers can cache packed weights per layout:
// bit i set = stride phase i (forward: bit 0) uses the TMA kernel's 64-channel tap pitch.
extern "C" int mog_packed_weight_layout(const MogConvDesc* d, int which) {
  if (validate(d, "mog_packed_weight_layout")) return -1;
  if (!use_tc(d)) return 0;
  if (which == 0) return tma_shape_eligible(fwd_problem(d)) ? 1 : 0;
  int tag = 0, i = 0;
  for (int ph = 0; ph < d->stride; ++ph)
    for (int pw = 0; pw < d->stride; ++pw, ++i) {
      IGemmParams p;
      if (!dgrad_problem(d, ph, pw, &p)) continue;
      if (tma_shape_eligible(p)) tag |= 1 << (i & 30);
    }
  return tag;
}

extern "C" int mog_pack_weight(const MogConvDesc* d, int which, const float* w, void* out, void* stream) {
  int rc = validate(d, "mog_pack_weight");
  if (rc) return rc;
  MOG_REQUIRE(w && out && (which == 0 || which == 1), "mog_pack_weight: bad argument");
  cudaStream_t st = as_stream(stream);
  const size_t total = (size_t)d->Cout * d->Cin * d->KH * d->KW;
  const unsigned blocks = (unsigned)ceil_div_ll((long long)total, 256);
  if (!use_tc(d)) {
    if (which == 0)
      pack_fwd_kernel<<<blocks, 256, 0, st>>>(w, static_cast<float*>(out), d->Cout, d->Cin, d->KH * d->KW);
    else
      pack_dgrad_kernel<<<blocks, 256, 0, st>>>(w, static_cast<float*>(out), d->Cout, d->Cin, d->KH * d->KW);
    return check_launch("pack_kernel");
  }
  if (which == 0) {
    int taps[64];
    for (int i = 0; i < d->KH * d->KW; ++i) taps[i] = i;
    return tc_pack_pitch(w, out, d->Cout, d->Cin, d->KH, d->KW, 0, d->KH * d->KW, taps, tap_pitch(fwd_problem(d), d->Cin),
                         passes_of(d), st);
  }
  unsigned char* o = static_cast<unsigned char*>(out);
  for (int ph = 0; ph < d->stride; ++ph)
    for (int pw = 0; pw < d->stride; ++pw) {
      IGemmParams p;
      if (!dgrad_problem(d, ph, pw, &p)) continue;
      rc = tc_pack_pitch(w, o, d->Cout, d->Cin, d->KH, d->KW, 1, p.nth * p.ntw, p.tapw, tap_pitch(p, d->Cout), passes_of(d), st);
      if (rc) return rc;
      o += tc_bytes(p, d->Cout, passes_of(d));
    }
  return MOG_OK;
}

extern "C" size_t mog_conv_workspace_bytes(const MogConvDesc* d, int which) {
  if (validate(d, "mog_conv_workspace_bytes")) return 0;
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  if (which == 0) {
    if (!use_tc(d) || tma_shape_eligible(fwd_problem(d))) return 0;
    return tc_igemm_workspace_bytes((long long)d->N * Ho * Wo, d->KH * d->KW, p8(d->Cin), d->Cout, passes_of(d));
  }
  if (which == 1) return dgrad_up_bytes(d) + dgrad_split_bytes(d);
  if (which == 2) return use_tc(d) ? tc_wgrad_workspace_bytes(*d, Ho, Wo) : wgrad_ffma_workspace_bytes(*d, Ho, Wo);
  return 0;
}

extern "C" int mog_conv2d_fwd(const MogConvDesc* d, const float* x, const void* x_planes, const void* w,
                              const float* bias, float* y, void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_fwd");
  if (rc) return rc;
  MOG_REQUIRE((x || x_planes) && w && y, "mog_conv2d_fwd: null tensor");
  IGemmParams p = fwd_problem(d);
  p.bias = bias; p.dst = y;
  if (use_tc(d)) {
    const bool tma = tma_shape_eligible(p);
    if (tma && !x_planes) return fail(MOG_ERR_BAD_ARG, "mog_conv2d_fwd: this shape runs on the TMA kernel and needs x_planes");
    rc = attach_source(&p, x, x_planes, (long long)d->N * d->H * d->W, "mog_conv2d_fwd");
    if (rc) return rc;
    if (tma) return launch_igemm_tma(p, w, passes_of(d), 0, as_stream(stream));
    return launch_igemm_tc(p, w, passes_of(d), workspace, ws_bytes, as_stream(stream));
  }
  MOG_REQUIRE(x, "mog_conv2d_fwd: fp32 precision needs the fp32 input");
  p.src = x;
  p.wmat = static_cast<const float*>(w);
  return launch_igemm_ffma(p, as_stream(stream));
}

extern "C" int mog_conv2d_dgrad(const MogConvDesc* d, const float* dy, const void* dy_planes, const void* wt,
                                float* dx, void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_dgrad");
  if (rc) return rc;
  MOG_REQUIRE((dy || dy_planes) && wt && dx, "mog_conv2d_dgrad: null tensor");
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  float* target = dx;
  const size_t need = mog_conv_workspace_bytes(d, 1);
  if (need && (!workspace || ws_bytes < need)) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_dgrad: workspace %zu < %zu", ws_bytes, need);
  if (d->up2x) target = static_cast<float*>(workspace);
  unsigned char* split_ws = workspace ? static_cast<unsigned char*>(workspace) + dgrad_up_bytes(d) : nullptr;
  const size_t split_bytes = need - dgrad_up_bytes(d);
  cudaStream_t st = as_stream(stream);
  const unsigned char* wp = static_cast<const unsigned char*>(wt);
  for (int ph = 0; ph < d->stride; ++ph) {
    for (int pw = 0; pw < d->stride; ++pw) {
      IGemmParams p;
      if (!dgrad_problem(d, ph, pw, &p)) continue;
      p.bias = nullptr; p.dst = target;
      if (use_tc(d)) {
        const bool tma = tma_shape_eligible(p);
        const size_t wbytes = tc_bytes(p, d->Cout, passes_of(d));
        if (tma && !dy_planes) return fail(MOG_ERR_BAD_ARG, "mog_conv2d_dgrad: this shape runs on the TMA kernel and needs dy_planes");
        rc = attach_source(&p, dy, dy_planes, (long long)d->N * Ho * Wo, "mog_conv2d_dgrad");
        if (rc) return rc;
        rc = tma ? launch_igemm_tma(p, wp, passes_of(d), 0, st) : launch_igemm_tc(p, wp, passes_of(d), split_ws, split_bytes, st);
        wp += wbytes;
      } else {
        MOG_REQUIRE(dy, "mog_conv2d_dgrad: fp32 precision needs the fp32 gradient");
        p.src = dy;
        p.wmat = static_cast<const float*>(wt);
        rc = launch_igemm_ffma(p, st);
      }
      if (rc) return rc;
    }
  }
  if (d->up2x) return launch_sumpool(target, dx, d->N, d->H, d->W, d->Cin, st);
  return MOG_OK;
}

extern "C" int mog_conv2d_wgrad(const MogConvDesc* d, const float* x, const void* x_planes, const float* dy,
                                const void* dy_planes, float* dw, float* dbias, void* workspace, size_t ws_bytes,
                                void* stream) {
  int rc = validate(d, "mog_conv2d_wgrad");
  if (rc) return rc;
  MOG_REQUIRE((x || x_planes) && (dy || dy_planes) && dw, "mog_conv2d_wgrad: null tensor");
  MOG_REQUIRE(!dbias || dy, "mog_conv2d_wgrad: the bias gradient needs the fp32 dy");
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  size_t need = mog_conv_workspace_bytes(d, 2);
  if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_wgrad: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = as_stream(stream);
  int splits = 1;
  float* ws = static_cast<float*>(workspace);
  int CinP = d->Cin;
  if (use_tc(d)) {
    const bool planes = x_planes && dy_planes;
    if (!planes) {
      MOG_REQUIRE(x && dy, "mog_conv2d_wgrad: give both operands as planes or both as fp32 tensors");
      if ((d->Cin % 8) || (d->Cout % 8))
        return fail(MOG_ERR_UNSUPPORTED, "mog_conv2d_wgrad: %d/%d channels need pre-split planes in tcgen05 precision", d->Cin, d->Cout);
    }
    CinP = planes ? p8(d->Cin) : d->Cin;
    rc = launch_wgrad_tc(*d, Ho, Wo, x, dy, planes ? x_planes : nullptr, (size_t)d->N * d->H * d->W * p8(d->Cin),
                         planes ? dy_planes : nullptr, (size_t)d->N * Ho * Wo * p8(d->Cout), ws, passes_of(d), &splits, st);
  } else {
    MOG_REQUIRE(x && dy, "mog_conv2d_wgrad: fp32 precision needs fp32 tensors");
    rc = launch_wgrad_ffma_partial(*d, Ho, Wo, x, dy, ws, &splits, st);
  }
  if (rc) return rc;
  rc = launch_wgrad_reduce(ws, dw, splits, d->KH * d->KW * CinP, d->Cout, d->Cin, CinP, d->KH * d->KW, st);
  if (rc) return rc;
  if (dbias) return launch_colsum(dy, dbias, (long long)d->N * Ho * Wo, d->Cout, st);
  return MOG_OK;
}
