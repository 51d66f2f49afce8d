// conv.cu -- C-ABI entry points of the convolution family.
//
// A convolution (forward or data gradient) is decomposed into one or more *gather-GEMM problems*
// (IGemmParams): rows = a grid of output pixels, K = (local tap, channel), unit or strided gather
// from an NHWC source (possibly a parity *view* of it), destination pixels on a strided sub-grid.
//   MOG_PREC_FP32            -> one problem per conv / stride phase on the CUDA cores (conv_ffma.cu)
//   MOG_PREC_BF16X3 / _BF16  -> tcgen05 kernels, always:
//       * upsample(2x nearest)+conv  = 4 sub-pixel phases: each a stride-1 conv with 2x2 taps whose
//         weights are pre-summed filter taps (2.25x fewer MACs than the dense 3x3 on the 4x grid);
//         its data gradient = the 4 transposed phase convs over the parity views of dy, accumulated.
//       * stride-s conv = s*s unit-stride convs over the parity views of x, accumulated (when the
//         grid is big enough for the TMA kernel), its data gradient = one unit-stride problem per
//         stride phase of dx.
//       * each problem runs on the persistent halo-tile TMA kernel (conv_halo.cu) when its shape allows
//         (problems that accumulate into the same pixels - parity views - are merged into ONE launch whose
//         K loop runs over all views), else on the generic gather kernel with split-K (conv_tc.cu).
//     Operands are pre-split bf16 planes (mog_split_planes) or, for the generic kernel only, fp32
//     tensors split on the fly (channel count multiple of 8).
#include "conv_common.cuh"

using namespace mog;

namespace mog {
// conv_ffma.cu
__global__ void pack_fwd_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW);
__global__ void pack_dgrad_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int KHW);
}  // namespace mog

// padding along W: MogConvDesc.pad_w1 = 1 + pad_w, or 0 for "same as pad" (the 1x7 / 7x1 / 1x3 / 3x1 filters of the DAMSM image encoder)
static int padw(const MogConvDesc* d) { return d->pad_w1 > 0 ? d->pad_w1 - 1 : d->pad; }

static int validate(const MogConvDesc* d, const char* who) {
  MOG_REQUIRE(d, "%s: null descriptor", who);
  MOG_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: non-positive dims", who);
  MOG_REQUIRE(d->KH > 0 && d->KW > 0 && d->KH <= 8 && d->KW <= 8 && d->KH * d->KW <= 64, "%s: filter %dx%d unsupported", who, d->KH, d->KW);
  MOG_REQUIRE(d->stride >= 1 && d->stride <= 4 && d->pad >= 0 && d->pad_w1 >= 0, "%s: bad stride/pad", who);
  MOG_REQUIRE(d->up2x == 0 || d->up2x == 1, "%s: up2x must be 0/1", who);
  MOG_REQUIRE(d->precision >= MOG_PREC_FP32 && d->precision <= MOG_PREC_BF16, "%s: unknown precision %d", who, d->precision);
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  MOG_REQUIRE(HL + 2 * d->pad >= d->KH && WL + 2 * padw(d) >= d->KW, "%s: filter larger than padded input", who);
  return MOG_OK;
}

static void out_hw(const MogConvDesc* d, int* Ho, int* Wo) {
  int HL = d->H << d->up2x, WL = d->W << d->up2x;
  *Ho = (HL + 2 * d->pad - d->KH) / d->stride + 1;
  *Wo = (WL + 2 * padw(d) - d->KW) / d->stride + 1;
}

static int passes_of(const MogConvDesc* d) { return d->precision == MOG_PREC_BF16X3 ? 3 : 1; }
static bool use_tc(const MogConvDesc* d) { return d->precision != MOG_PREC_FP32; }
static int p8(int c) { return ceil_div(c, 8) * 8; }
static int fdiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }   // floor division, b > 0
static int pmod(int a, int b) { return ((a % b) + b) % b; }

// ---- problems ---------------------------------------------------------------------------------
struct Problem {
  IGemmParams g;      // shape; src/dst/bias pointers are attached at call time
  int taps[64][4];    // filter taps (kh*KW + kw) summed into each local tap, -1 = unused
  int transpose;      // weight operand orientation: 0 n=co,c=ci (forward)   1 n=ci,c=co (data gradient)
  int CsReal;         // channels of the gathered tensor
  int src_pixels_h, src_pixels_w;   // physical source dims
};

static void clear_taps(Problem* q) {
  for (int i = 0; i < 64; ++i)
    for (int u = 0; u < 4; ++u) q->taps[i][u] = -1;
}

// the single dense problem of a forward conv (also the CUDA-core formulation)
static void fwd_single(const MogConvDesc* d, Problem* q) {
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  IGemmParams p{};
  p.N = d->N; p.Hs = d->H; p.Ws = d->W; p.Cs = d->Cin; p.up2x = d->up2x;
  p.Hr = Ho; p.Wr = Wo; p.rs = d->stride;
  p.nth = d->KH; p.ntw = d->KW;
  for (int i = 0; i < d->KH; ++i) p.off_h[i] = i - d->pad;
  for (int i = 0; i < d->KW; ++i) p.off_w[i] = i - padw(d);
  for (int i = 0; i < d->KH * d->KW; ++i) p.tapw[i] = i;
  p.Cd = d->Cout; p.Hd = Ho; p.Wd = Wo; p.dsh = 1; p.doh = 0; p.dsw = 1; p.dow = 0;
  p.act = d->act;
  p.M = (long long)d->N * Ho * Wo;
  p.K = d->KH * d->KW * d->Cin;
  q->g = p;
  clear_taps(q);
  for (int i = 0; i < d->KH * d->KW; ++i) q->taps[i][0] = i;
  q->transpose = 0; q->CsReal = d->Cin; q->src_pixels_h = d->H; q->src_pixels_w = d->W;
}

// One gather-GEMM per stride phase of dx: input pixels (hi, wi) with hi%s==ph, wi%s==pw only see the taps
// kh with (ph + pad - kh) % s == 0, at dy row hi/s + (ph + pad - kh)/s.  Returns false for an empty phase.
static bool dgrad_phase(const MogConvDesc* d, int ph, int pw, Problem* q) {
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
    if (pmod(ph + d->pad - kh, s) == 0) { khs[nth] = kh; p.off_h[nth] = (ph + d->pad - kh) / s; ++nth; }
  for (int kw = 0; kw < d->KW; ++kw)
    if (pmod(pw + padw(d) - kw, s) == 0) { kws[ntw] = kw; p.off_w[ntw] = (pw + padw(d) - kw) / s; ++ntw; }
  if (nth == 0 || ntw == 0) { nth = 0; ntw = 1; }  // no tap reaches this phase: K = 0, zeros are written
  p.nth = nth; p.ntw = ntw;
  clear_taps(q);
  for (int a = 0; a < nth; ++a)
    for (int b = 0; b < ntw; ++b) {
      p.tapw[a * ntw + b] = khs[a] * d->KW + kws[b];
      q->taps[a * ntw + b][0] = khs[a] * d->KW + kws[b];
    }
  p.Cd = d->Cin; p.Hd = HL; p.Wd = WL; p.dsh = s; p.doh = ph; p.dsw = s; p.dow = pw;
  p.act = MOG_ACT_NONE;
  p.M = (long long)d->N * p.Hr * p.Wr;
  p.K = nth * ntw * d->Cout;
  q->g = p;
  q->transpose = 1; q->CsReal = d->Cout; q->src_pixels_h = Ho; q->src_pixels_w = Wo;
  return true;
}

// upsample+conv as 4 sub-pixel phases: output row 2i+a reads low-res rows i + floor((a + kh - pad)/2)
static bool up2x_phases_ok(const MogConvDesc* d) {
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  return d->up2x && d->stride == 1 && Ho == 2 * d->H && Wo == 2 * d->W && d->KH <= 3 && d->KW <= 3;
}
// distinct low-res offsets of phase a along one axis and the filter taps that map to each
static int phase_axis(int a, int K, int pad, int* offs, int (*members)[2]) {
  int n = 0;
  for (int k = 0; k < K; ++k) {
    const int dd = fdiv(a + k - pad, 2);
    int j = 0;
    for (; j < n; ++j)
      if (offs[j] == dd) break;
    if (j == n) { offs[n] = dd; members[n][0] = members[n][1] = -1; ++n; }
    if (members[j][0] < 0) members[j][0] = k; else members[j][1] = k;
  }
  return n;
}
static void up2x_phase(const MogConvDesc* d, int a, int b, bool dgrad, Problem* q) {
  int oh[4], ow[4], mh[4][2], mw[4][2];
  const int nth = phase_axis(a, d->KH, d->pad, oh, mh), ntw = phase_axis(b, d->KW, padw(d), ow, mw);
  IGemmParams p{};
  p.N = d->N; p.up2x = 0; p.rs = 1; p.nth = nth; p.ntw = ntw;
  p.Hr = d->H; p.Wr = d->W;
  p.M = (long long)d->N * d->H * d->W;
  clear_taps(q);
  for (int i = 0; i < nth; ++i)
    for (int j = 0; j < ntw; ++j) {
      int u = 0;
      for (int x = 0; x < 2; ++x)
        for (int y = 0; y < 2; ++y)
          if (mh[i][x] >= 0 && mw[j][y] >= 0) q->taps[i * ntw + j][u++] = mh[i][x] * d->KW + mw[j][y];
    }
  if (!dgrad) {
    // y[2i+a, 2j+b] = sum_taps x[i + oh, j + ow] * W'
    p.Hs = d->H; p.Ws = d->W; p.Cs = d->Cin;
    for (int i = 0; i < nth; ++i) p.off_h[i] = oh[i];
    for (int j = 0; j < ntw; ++j) p.off_w[j] = ow[j];
    p.Cd = d->Cout; p.Hd = 2 * d->H; p.Wd = 2 * d->W; p.dsh = 2; p.doh = a; p.dsw = 2; p.dow = b;
    p.act = d->act;
    q->transpose = 0; q->CsReal = d->Cin; q->src_pixels_h = d->H; q->src_pixels_w = d->W;
  } else {
    // dx[i, j] += sum_taps dy[2(i - oh) + a, 2(j - ow) + b] * W'^T : parity view (a, b) of dy, offsets -oh
    p.Hs = d->H; p.Ws = d->W; p.Cs = d->Cout;
    p.vstep = 2; p.voh = a; p.vow = b; p.Hp = 2 * d->H; p.Wp = 2 * d->W;
    for (int i = 0; i < nth; ++i) p.off_h[i] = -oh[i];
    for (int j = 0; j < ntw; ++j) p.off_w[j] = -ow[j];
    p.Cd = d->Cin; p.Hd = d->H; p.Wd = d->W; p.dsh = 1; p.doh = 0; p.dsw = 1; p.dow = 0;
    p.act = MOG_ACT_NONE;
    p.accum_dst = (a | b) != 0;
    q->transpose = 1; q->CsReal = d->Cout; q->src_pixels_h = 2 * d->H; q->src_pixels_w = 2 * d->W;
  }
  p.K = nth * ntw * p.Cs;
  q->g = p;
}

// stride-s forward conv over the parity view (a, b) of x: taps with (k - pad) % s == a at view offset (k - pad - a)/s
static bool strided_view(const MogConvDesc* d, int a, int b, Problem* q) {
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  const int s = d->stride;
  IGemmParams p{};
  p.N = d->N; p.up2x = 0; p.rs = 1;
  p.vstep = s; p.voh = a; p.vow = b; p.Hp = d->H; p.Wp = d->W;
  p.Hs = d->H / s; p.Ws = d->W / s; p.Cs = d->Cin;
  p.Hr = Ho; p.Wr = Wo;
  int nth = 0, ntw = 0, khs[8], kws[8];
  for (int kh = 0; kh < d->KH; ++kh)
    if (pmod(kh - d->pad, s) == a) { khs[nth] = kh; p.off_h[nth] = fdiv(kh - d->pad - a, s); ++nth; }
  for (int kw = 0; kw < d->KW; ++kw)
    if (pmod(kw - padw(d), s) == b) { kws[ntw] = kw; p.off_w[ntw] = fdiv(kw - padw(d) - b, s); ++ntw; }
  if (nth == 0 || ntw == 0) return false;
  p.nth = nth; p.ntw = ntw;
  clear_taps(q);
  for (int i = 0; i < nth; ++i)
    for (int j = 0; j < ntw; ++j) q->taps[i * ntw + j][0] = khs[i] * d->KW + kws[j];
  p.Cd = d->Cout; p.Hd = Ho; p.Wd = Wo; p.dsh = 1; p.doh = 0; p.dsw = 1; p.dow = 0;
  p.act = MOG_ACT_NONE;
  p.M = (long long)d->N * Ho * Wo;
  p.K = nth * ntw * p.Cs;
  q->g = p;
  q->transpose = 0; q->CsReal = d->Cin; q->src_pixels_h = d->H; q->src_pixels_w = d->W;
  return true;
}

// problems of the forward conv in tcgen05 precision; returns the count (<= 16)
static int build_fwd(const MogConvDesc* d, Problem* out) {
  if (up2x_phases_ok(d)) {
    int n = 0;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) up2x_phase(d, a, b, false, &out[n++]);
    return n;
  }
  if (d->stride > 1 && !d->up2x && (d->H % d->stride) == 0 && (d->W % d->stride) == 0) {
    Problem probe;
    if (strided_view(d, pmod(-d->pad, d->stride), pmod(-padw(d), d->stride), &probe) && halo_shape_eligible(probe.g)) {
      int n = 0;
      for (int a = 0; a < d->stride; ++a)
        for (int b = 0; b < d->stride; ++b)
          if (strided_view(d, a, b, &out[n])) ++n;
      for (int i = 0; i < n; ++i) out[i].g.accum_dst = i > 0;
      out[n - 1].g.act = d->act;   // bias (attached at call time to the last problem) and activation once, at the end
      return n;
    }
  }
  fwd_single(d, &out[0]);
  return 1;
}

// problems of the data gradient; *hires = 1 when they produce the gradient of the upsampled input
// (fallback path: caller sum-pools it)
static int build_dgrad(const MogConvDesc* d, Problem* out, int* hires) {
  *hires = 0;
  if (use_tc(d) && up2x_phases_ok(d)) {
    int n = 0;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) up2x_phase(d, a, b, true, &out[n++]);
    return n;
  }
  int n = 0;
  for (int ph = 0; ph < d->stride; ++ph)
    for (int pw = 0; pw < d->stride; ++pw)
      if (dgrad_phase(d, ph, pw, &out[n])) ++n;
  *hires = d->up2x;
  return n;
}

// tcgen05 path: k runs over (tap, channel); the channel pitch per tap is Cs rounded up to 8 for the
// generic gather kernel and to 32 for the halo kernel (whole 64-byte swizzle rows per box).
static int tap_pitch(const Problem& q) { return halo_shape_eligible(q.g) ? halo_tap_pitch(q.CsReal) : p8(q.CsReal); }
static size_t tc_bytes(const Problem& q, int passes) {
  return tc_packed_bytes(q.g.nth * q.g.ntw, tap_pitch(q), q.g.Cd, passes);
}
static size_t split_bytes(const Problem& q, int passes) {
  if (halo_shape_eligible(q.g)) {
    IGemmParams g = q.g;
    g.Cs = p8(q.CsReal);
    return halo_workspace_bytes(&g, 1, passes);     // (small-grid form: split-K partials)
  }
  return tc_igemm_workspace_bytes(q.g.M, q.g.nth * q.g.ntw, p8(q.CsReal), q.g.Cd, passes);
}

static size_t dgrad_up_bytes(const MogConvDesc* d, int hires) {
  size_t b = hires ? (size_t)d->N * (2 * d->H) * (2 * d->W) * d->Cin * sizeof(float) : 0;
  return (b + 255) / 256 * 256;
}

// attach the gathered operand (planes or fp32) of a tcgen05 problem; Cs becomes the channel pitch
static int attach_source(Problem* q, const float* src_f32, const void* planes, int N, const char* who) {
  IGemmParams* p = &q->g;
  if (planes) {
    p->Cs = p8(q->CsReal);
    p->src = nullptr;
    p->src_planes = planes;
    p->src_plane_elems = (size_t)N * q->src_pixels_h * q->src_pixels_w * p->Cs;
  } else {
    if (q->CsReal % 8) return fail(MOG_ERR_UNSUPPORTED, "%s: %d channels need pre-split planes (mog_split_planes) in tcgen05 precision", who, q->CsReal);
    if (!src_f32) return fail(MOG_ERR_BAD_ARG, "%s: neither an fp32 tensor nor planes given", who);
    p->Cs = q->CsReal;
    p->src = src_f32;
    p->src_planes = nullptr;
    p->src_plane_elems = 0;
  }
  p->K = p->nth * p->ntw * p->Cs;
  return MOG_OK;
}

// problems [0, n) write the same pixels through an accumulate chain (parity views): one merged halo launch?
static bool mergeable(const Problem* q, int n) {
  if (n < 2 || n > 4) return false;
  for (int i = 0; i < n; ++i) {
    const IGemmParams& g = q[i].g;
    const IGemmParams& g0 = q[0].g;
    if (!halo_shape_eligible(g)) return false;
    if (g.accum_dst != (i > 0)) return false;
    if (g.Hr != g0.Hr || g.Wr != g0.Wr || g.dsh != g0.dsh || g.doh != g0.doh || g.dsw != g0.dsw || g.dow != g0.dow ||
        g.Cd != g0.Cd || q[i].CsReal != q[0].CsReal)
      return false;
  }
  return true;
}

// runs the n problems of one conv (forward or data gradient) in a tcgen05 precision; `wp` = packed weight blocks back to back
static int run_tc_problems(Problem* probs, int n, const float* src_f32, const void* planes, int N, const unsigned char* wp, int passes,
                           const float* bias, bool bias_each, float* dst, void* ws, size_t ws_bytes, cudaStream_t st, const char* who) {
  const void* blocks[16];
  for (int i = 0; i < n; ++i) {
    blocks[i] = wp;
    wp += tc_bytes(probs[i], passes);
  }
  if (mergeable(probs, n)) {
    if (!planes) return fail(MOG_ERR_BAD_ARG, "%s: this shape runs on the halo kernel and needs pre-split planes", who);
    IGemmParams gs[4];
    for (int i = 0; i < n; ++i) {
      int rc = attach_source(&probs[i], src_f32, planes, N, who);
      if (rc) return rc;
      gs[i] = probs[i].g;
    }
    gs[0].dst = dst; gs[0].bias = bias; gs[0].act = probs[n - 1].g.act; gs[0].accum_dst = 0;
    return launch_igemm_halo(gs, n, blocks, passes, ws, ws_bytes, st);
  }
  for (int i = 0; i < n; ++i) {
    Problem& q = probs[i];
    q.g.dst = dst;
    // bias: problems writing their own pixels (sub-pixel phases) add it each; accumulate chains add it once, at the end
    q.g.bias = bias_each ? bias : (i == n - 1 ? bias : nullptr);
    const bool halo = halo_shape_eligible(q.g);
    if (halo && !planes) return fail(MOG_ERR_BAD_ARG, "%s: this shape runs on the halo kernel and needs pre-split planes", who);
    int rc = attach_source(&q, src_f32, planes, N, who);
    if (rc) return rc;
    rc = halo ? launch_igemm_halo(&q.g, 1, &blocks[i], passes, ws, ws_bytes, st) : launch_igemm_tc(q.g, blocks[i], passes, ws, ws_bytes, st);
    if (rc) return rc;
  }
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

extern "C" int mog_split_planes_act(const float* dy, const float* y, int act, long long rows, int C, int precision, void* planes,
                                    void* stream) {
  MOG_REQUIRE(dy && y && planes && rows > 0 && C > 0, "mog_split_planes_act: bad argument");
  MOG_REQUIRE(precision == MOG_PREC_BF16X3 || precision == MOG_PREC_BF16, "mog_split_planes_act: precision must be a tcgen05 mode");
  MOG_REQUIRE(act == MOG_ACT_RELU || act == MOG_ACT_LRELU || act == MOG_ACT_TANH || act == MOG_ACT_SIGMOID,
              "mog_split_planes_act: activation %d has no output-based derivative", act);
  return launch_split_planes(dy, rows, C, p8(C), planes, precision == MOG_PREC_BF16X3 ? 2 : 1, as_stream(stream), y, act);
}

extern "C" int mog_patch_planes(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int precision,
                               void* planes, void* stream) {
  MOG_REQUIRE(x && planes && N > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, "mog_patch_planes: bad argument");
  MOG_REQUIRE(precision == MOG_PREC_BF16X3 || precision == MOG_PREC_BF16, "mog_patch_planes: precision must be a tcgen05 mode");
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  MOG_REQUIRE(H + 2 * pad >= KH && W + 2 * pad >= KW, "mog_patch_planes: filter larger than the padded image");
  return launch_patch_planes(x, N, H, W, C, KH, KW, stride, pad, Ho, Wo, planes, precision == MOG_PREC_BF16X3 ? 2 : 1, as_stream(stream));
}

extern "C" int mog_col2im_act(const float* z, int ldz, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                             const float* bias, int act, float* y, void* stream) {
  MOG_REQUIRE(z && y && N > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, "mog_col2im_act: bad argument");
  MOG_REQUIRE(C <= 4 && KH * KW <= 16, "mog_col2im_act: implemented for C <= 4 channels and filters of at most 16 taps (C=%d, %dx%d)", C, KH, KW);
  MOG_REQUIRE(ldz >= KH * KW * C && (ldz & 3) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0,
              "mog_col2im_act: row pitch %d must be >= KH*KW*C and a multiple of 4, z 16-byte aligned", ldz);
  MOG_REQUIRE(H + 2 * pad >= KH && W + 2 * pad >= KW, "mog_col2im_act: filter larger than the padded image");
  MOG_REQUIRE(act == MOG_ACT_NONE || act == MOG_ACT_RELU || act == MOG_ACT_LRELU || act == MOG_ACT_TANH || act == MOG_ACT_SIGMOID,
              "mog_col2im_act: activation %d", act);
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  return launch_col2im_act(z, ldz, N, H, W, C, KH, KW, stride, pad, Ho, Wo, bias, act, y, as_stream(stream));
}

static int build(const MogConvDesc* d, int which, Problem* probs, int* hires) {
  *hires = 0;
  return which == 0 ? build_fwd(d, probs) : build_dgrad(d, probs, hires);
}

extern "C" size_t mog_packed_weight_bytes(const MogConvDesc* d, int which) {
  if (validate(d, "mog_packed_weight_bytes") || (which != 0 && which != 1)) return 0;
  if (!use_tc(d)) return (size_t)d->KH * d->KW * d->Cin * d->Cout * sizeof(float);
  Problem probs[16];
  int hires;
  const int n = build(d, which, probs, &hires);
  size_t tot = 0;
  for (int i = 0; i < n; ++i) tot += tc_bytes(probs[i], passes_of(d));
  return tot;
}

// Identifies the packed layout chosen for (d, which) so callers can cache packed weights per layout:
// low bits: which problems use the TMA kernel's 64-channel tap pitch; high bits: the decomposition.
extern "C" int mog_packed_weight_layout(const MogConvDesc* d, int which) {
  if (validate(d, "mog_packed_weight_layout") || (which != 0 && which != 1)) return -1;
  if (!use_tc(d)) return 0;
  Problem probs[16];
  int hires;
  const int n = build(d, which, probs, &hires);
  int tag = n << 20;
  for (int i = 0; i < n; ++i)
    if (halo_shape_eligible(probs[i].g)) tag |= 1 << i;
  return tag;
}

extern "C" int mog_pack_weight(const MogConvDesc* d, int which, const float* w, void* out, void* stream) {
  int rc = validate(d, "mog_pack_weight");
  if (rc) return rc;
  MOG_REQUIRE(w && out && (which == 0 || which == 1), "mog_pack_weight: bad argument");
  cudaStream_t st = as_stream(stream);
  if (!use_tc(d)) {
    const size_t total = (size_t)d->Cout * d->Cin * d->KH * d->KW;
    const unsigned blocks = (unsigned)ceil_div_ll((long long)total, 256);
    if (which == 0)
      pack_fwd_kernel<<<blocks, 256, 0, st>>>(w, static_cast<float*>(out), d->Cout, d->Cin, d->KH * d->KW);
    else
      pack_dgrad_kernel<<<blocks, 256, 0, st>>>(w, static_cast<float*>(out), d->Cout, d->Cin, d->KH * d->KW);
    return check_launch("pack_kernel");
  }
  Problem probs[16];
  int hires;
  const int n = build(d, which, probs, &hires);
  unsigned char* o = static_cast<unsigned char*>(out);
  for (int i = 0; i < n; ++i) {
    const Problem& q = probs[i];
    rc = tc_pack_pitch(w, o, d->Cout, d->Cin, d->KH, d->KW, q.transpose, q.g.nth * q.g.ntw, q.taps, tap_pitch(q), passes_of(d), st);
    if (rc) return rc;
    o += tc_bytes(q, passes_of(d));
  }
  return MOG_OK;
}

extern "C" int mog_pack_plan(const MogConvDesc* d, int which, const float* w, void* out, MogPackEntry* entries, int capacity) {
  int rc = validate(d, "mog_pack_plan");
  if (rc) return rc;
  MOG_REQUIRE(w && out && entries && capacity > 0 && (which == 0 || which == 1), "mog_pack_plan: bad argument");
  if (!use_tc(d)) return fail(MOG_ERR_UNSUPPORTED, "mog_pack_plan: tcgen05 precisions only");
  Problem probs[16];
  int hires;
  const int n = build(d, which, probs, &hires);
  if (n > capacity) return fail(MOG_ERR_BAD_ARG, "mog_pack_plan: %d entries > capacity %d", n, capacity);
  unsigned char* o = static_cast<unsigned char*>(out);
  for (int i = 0; i < n; ++i) {
    const Problem& q = probs[i];
    rc = tc_pack_entry(w, o, d->Cout, d->Cin, d->KH, d->KW, q.transpose, q.g.nth * q.g.ntw, q.taps, tap_pitch(q), passes_of(d), &entries[i]);
    if (rc) return fail(rc, "mog_pack_plan: filter %dx%d / %d taps not supported by the multi-tensor form", d->KH, d->KW, q.g.nth * q.g.ntw);
    o += tc_bytes(q, passes_of(d));
  }
  return n;
}

extern "C" int mog_pack_multi(const MogPackEntry* entries_dev, const MogPackGroup* groups_dev, int ngroups, int total_blocks,
                              void* stream) {
  MOG_REQUIRE(entries_dev && groups_dev && ngroups > 0 && total_blocks > 0, "mog_pack_multi: bad argument");
  return launch_pack_multi(entries_dev, groups_dev, ngroups, total_blocks, as_stream(stream));
}

extern "C" size_t mog_conv_workspace_bytes(const MogConvDesc* d, int which) {
  if (validate(d, "mog_conv_workspace_bytes")) return 0;
  int Ho, Wo;
  out_hw(d, &Ho, &Wo);
  if (which == 2) {
    if (!use_tc(d)) return wgrad_ffma_workspace_bytes(*d, Ho, Wo);
    const size_t a = tc_wgrad_workspace_bytes(*d, Ho, Wo), b = wgrad_halo_workspace_bytes(*d, Ho, Wo, passes_of(d));
    return a > b ? a : b;
  }
  if (which != 0 && which != 1) return 0;
  if (!use_tc(d)) return which == 1 ? dgrad_up_bytes(d, d->up2x) : 0;
  Problem probs[16];
  int hires;
  const int n = build(d, which, probs, &hires);
  size_t mx = 0;
  for (int i = 0; i < n; ++i) {
    size_t b = split_bytes(probs[i], passes_of(d));
    if (b > mx) mx = b;
  }
  if (mergeable(probs, n)) {      // parity views merged into one launch: its own split-K plan
    IGemmParams gs[4];
    for (int i = 0; i < n; ++i) { gs[i] = probs[i].g; gs[i].Cs = p8(probs[i].CsReal); }
    gs[0].accum_dst = 0;
    size_t b = halo_workspace_bytes(gs, n, passes_of(d));
    if (b > mx) mx = b;
  }
  return dgrad_up_bytes(d, hires) + mx;
}

extern "C" int mog_conv2d_fwd(const MogConvDesc* d, const float* x, const void* x_planes, const void* w,
                              const float* bias, float* y, void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_fwd");
  if (rc) return rc;
  MOG_REQUIRE((x || x_planes) && w && y, "mog_conv2d_fwd: null tensor");
  cudaStream_t st = as_stream(stream);
  if (!use_tc(d)) {
    MOG_REQUIRE(x, "mog_conv2d_fwd: fp32 precision needs the fp32 input");
    Problem q;
    fwd_single(d, &q);
    q.g.src = x; q.g.bias = bias; q.g.dst = y;
    q.g.wmat = static_cast<const float*>(w);
    return launch_igemm_ffma(q.g, st);
  }
  Problem probs[16];
  const int n = build_fwd(d, probs);
  const size_t need = mog_conv_workspace_bytes(d, 0);
  if (need && (!workspace || ws_bytes < need)) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_fwd: workspace %zu < %zu", ws_bytes, need);
  // bias: every sub-pixel phase writes its own pixels (bias each); accumulated parity views add it once, at the end
  const bool bias_each = !(n > 1 && (probs[1].g.accum_dst || probs[0].g.vstep > 1));
  return run_tc_problems(probs, n, x, x_planes, d->N, static_cast<const unsigned char*>(w), passes_of(d), bias, bias_each, y, workspace,
                         ws_bytes, st, "mog_conv2d_fwd");
}

extern "C" int mog_conv2d_dgrad(const MogConvDesc* d, const float* dy, const void* dy_planes, const void* wt,
                                float* dx, void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(d, "mog_conv2d_dgrad");
  if (rc) return rc;
  MOG_REQUIRE((dy || dy_planes) && wt && dx, "mog_conv2d_dgrad: null tensor");
  cudaStream_t st = as_stream(stream);
  if (use_tc(d) && dy && patch_dgrad_eligible(*d)) return launch_patch_dgrad(*d, dy, wt, dx, passes_of(d), st);
  Problem probs[16];
  int hires;
  const int n = build_dgrad(d, probs, &hires);
  const size_t need = mog_conv_workspace_bytes(d, 1);
  if (need && (!workspace || ws_bytes < need)) return fail(MOG_ERR_WORKSPACE, "mog_conv2d_dgrad: workspace %zu < %zu", ws_bytes, need);
  float* target = hires ? static_cast<float*>(workspace) : dx;
  unsigned char* sws = workspace ? static_cast<unsigned char*>(workspace) + dgrad_up_bytes(d, hires) : nullptr;
  const size_t sbytes = need - dgrad_up_bytes(d, hires);
  if (use_tc(d)) {
    rc = run_tc_problems(probs, n, dy, dy_planes, d->N, static_cast<const unsigned char*>(wt), passes_of(d), nullptr, true, target, sws,
                         sbytes, st, "mog_conv2d_dgrad");
    if (rc) return rc;
  } else {
    MOG_REQUIRE(dy, "mog_conv2d_dgrad: fp32 precision needs the fp32 gradient");
    for (int i = 0; i < n; ++i) {
      Problem& q = probs[i];
      q.g.bias = nullptr; q.g.dst = target;
      q.g.src = dy;
      q.g.wmat = static_cast<const float*>(wt);
      rc = launch_igemm_ffma(q.g, st);
      if (rc) return rc;
    }
  }
  if (hires) return launch_sumpool(target, dx, d->N, d->H, d->W, d->Cin, st);
  return MOG_OK;
}

extern "C" int mog_conv2d_wgrad(const MogConvDesc* d, const float* x, const void* x_planes, const float* dy,
                                const void* dy_planes, float* dw, float* dbias, void* workspace, size_t ws_bytes,
                                void* stream) {
  int rc = validate(d, "mog_conv2d_wgrad");
  if (rc) return rc;
  MOG_REQUIRE((x || x_planes) && (dy || dy_planes) && dw, "mog_conv2d_wgrad: null tensor");
  MOG_REQUIRE(!dbias || dy, "mog_conv2d_wgrad: the bias gradient needs the fp32 dy");
  if (padw(d) != d->pad) return fail(MOG_ERR_UNSUPPORTED, "mog_conv2d_wgrad: different padding along H and W is implemented for the forward conv and the data gradient only");
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
    if (planes && wgrad_halo_eligible(*d, Ho, Wo, passes_of(d))) {
      // halo-tile kernel: x staged once per filter column, reduce + OIHW scatter included
      rc = launch_wgrad_halo(*d, Ho, Wo, x_planes, dy_planes, dw, ws, passes_of(d), st);
      if (rc) return rc;
      if (dbias) return launch_colsum(dy, dbias, (long long)d->N * Ho * Wo, d->Cout, st);
      return MOG_OK;
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
