/*
 * mog.h -- C ABI of libmog.so, the sm_100a implementation of the generator / discriminator
 * forward+backward hot path of tohinz/multiple-objects-gan (AttnGAN variant first).
 *
 * The reference has no native/FFI layer: the path is a set of torch library calls issued from
 * code/coco/attngan/{model.py,GlobalAttention.py,miscc/losses.py}.  Each entry point below names
 * the reference call site(s) it replaces ("replaces:").  All tensors are caller-owned device
 * buffers; activations are NHWC fp32 (a logically-NCHW torch tensor in channels_last memory
 * format is exactly that), matrices are row-major.  Every function enqueues work on `stream`
 * (a cudaStream_t passed as void*), never synchronises, never allocates, and returns 0 or a
 * negative MOG_ERR_* code; mog_last_error() gives the message of the calling thread's last
 * failure.  There is no CPU fallback.
 */
#ifndef MOG_H_
#define MOG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOG_VERSION 100

enum {
  MOG_OK = 0,
  MOG_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, inconsistent shapes */
  MOG_ERR_UNSUPPORTED = -2,  /* shape / mode not implemented by the selected kernel */
  MOG_ERR_WORKSPACE = -3,    /* workspace too small */
  MOG_ERR_CUDA = -4          /* a CUDA call or launch failed */
};

/* activation / epilogue codes */
enum {
  MOG_ACT_NONE = 0,
  MOG_ACT_RELU = 1,
  MOG_ACT_LRELU = 2, /* negative slope 0.2 (model.py:93,579,589) */
  MOG_ACT_GLU = 3,   /* x[:, :C] * sigmoid(x[:, C:]) (model.py:24-32): 2C channels in, C out */
  MOG_ACT_TANH = 4,  /* model.py:470 */
  MOG_ACT_SIGMOID = 5
};

/* operand precision of the convolution kernels */
enum {
  MOG_PREC_FP32 = 0,   /* CUDA-core FFMA implicit GEMM, fp32 operands and accumulate            */
  MOG_PREC_BF16X3 = 1, /* tcgen05: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, fp32 accumulate in TMEM   */
  MOG_PREC_BF16 = 2    /* tcgen05: single-pass bf16 operands, fp32 accumulate in TMEM           */
};

typedef struct MogConvDesc {
  int32_t N, H, W, Cin;  /* input tensor NHWC (before the optional fused upsample)               */
  int32_t Cout, KH, KW;  /* filter                                                               */
  int32_t stride, pad;
  int32_t up2x;          /* 1: the conv reads nearest-neighbour 2x upsampled input (upBlock)     */
  int32_t act;           /* fused epilogue: MOG_ACT_NONE / LRELU / TANH (bias added first)       */
  int32_t precision;     /* MOG_PREC_*                                                           */
  int32_t pad_w1;        /* 0: `pad` applies to both axes; else 1 + padding along W (`pad` = along H):
                            the 1x7 / 7x1 / 1x3 / 3x1 filters of the DAMSM image encoder (forward + dgrad) */
} MogConvDesc;

int mog_version(void);
const char* mog_last_error(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py: gpu_launches) */
unsigned long long mog_launch_count(void);

/* ---- layout ---------------------------------------------------------------------------- */
/* replaces: nothing (the reference is NCHW throughout); boundary helpers for NCHW callers.  */
int mog_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, void* stream);
int mog_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, void* stream);
/* Packs an OIHW fp32 weight (state_dict layout) into the GEMM B operand of the forward conv
 * (which = 0) or of the data gradient (which = 1) for the kernel selected by d->precision:
 * fp32 [K][N] for the CUDA-core path; bf16 hi(/lo) planes [Npad][Kpad], K-major, one block per
 * stride phase for the tcgen05 path.  The buffer is opaque; size it with mog_packed_weight_bytes.
 * Re-pack after every optimiser step (weights changed). */
size_t mog_packed_weight_bytes(const MogConvDesc* d, int which);
/* Tag of the packed layout chosen for (d, which): it depends on the problem shape (the TMA-staged
 * kernel uses a 64-channel tap pitch), so cache packed weights per (parameter version, tag). */
int mog_packed_weight_layout(const MogConvDesc* d, int which);
int mog_pack_weight(const MogConvDesc* d, int which, const float* w_oihw, void* w_packed, void* stream);

/* Multi-tensor repacking: after an optimiser step every trainable weight needs its packed operands again
 * (~350 small launches per step when done one problem at a time).  mog_pack_plan describes the launches
 * mog_pack_weight(d, which, w, out) would make as table entries; the caller concatenates the entries of all weights of a
 * network, fills `block_start` (exclusive prefix sum of nxb * nyb), keeps the table in device memory and refreshes all
 * packs with ONE mog_pack_multi launch per step.  tcgen05 precisions, filters of at most 16 taps. */
typedef struct MogPackEntry {
  const float* w;          /* OIHW fp32 source                                                   */
  void* hi;                /* destination planes of this problem                                  */
  void* lo;                /* NULL in single-pass mode                                            */
  int32_t Cout, Cin, KHW, transpose, ntaps, Nreal, Npad, Cs, CsReal, K, Kpad;
  int32_t nxb, nyb;        /* tile grid of the entry: ceil(Cs / 32) x ceil(Npad / 16) blocks      */
  int32_t block_start;     /* first block of the entry within the launch (filled by the caller)   */
  int32_t taps[16][4];     /* filter taps summed into each local tap, -1 = unused                 */
} MogPackEntry;
/* Entries that read the same source tiles (the problems of one weight: sub-pixel phases, parity views, stride phases --
 * same w, orientation, channel pitch and row padding) form a group: a block loads its 16 x 32 x taps tile ONCE and writes it
 * into every entry of the group.  block_start (exclusive prefix sum of nxb * nyb over the groups) is filled by the caller. */
typedef struct MogPackGroup {
  int32_t first, count;    /* entries [first, first + count) of the entry table                   */
  int32_t block_start;     /* first block of the group within the launch                          */
  int32_t nxb;             /* blocks along the channel axis (= the entries' nxb)                   */
} MogPackGroup;
/* returns the number of entries written (<= capacity), or a negative MOG_ERR_* (MOG_ERR_UNSUPPORTED: use mog_pack_weight) */
int mog_pack_plan(const MogConvDesc* d, int which, const float* w, void* out, MogPackEntry* entries, int capacity);
int mog_pack_multi(const MogPackEntry* entries_dev, const MogPackGroup* groups_dev, int ngroups, int total_blocks, void* stream);

/* ---- convolution ----------------------------------------------------------------------- */
/* Operand formats.  MOG_PREC_FP32: fp32 NHWC tensors.  tcgen05 precisions: the gathered operands
 * (x for forward/wgrad, dy for dgrad/wgrad) are preferably passed as pre-split bf16 *planes*
 * made by mog_split_planes -- [rows][C8] bf16 "hi" plane followed (BF16X3 only) by the "lo" plane,
 * C8 = C rounded up to 8, pad channels zero -- so a tensor is split once and reused by every
 * kernel that reads it; when the planes pointer is NULL the fp32 tensor is split on the fly
 * (channel count must then be a multiple of 8). */
size_t mog_planes_bytes(long long rows, int C, int precision);
int mog_split_planes(const float* x, long long rows, int C, int precision, void* planes, void* stream);
/* Same for the gradient of a conv whose epilogue applied an activation: splits dy * act'(y) (y = the activation OUTPUT;
 * RELU / LRELU / TANH / SIGMOID), i.e. mog_act_bwd fused into the split -- the fp32 dz is never materialised. */
int mog_split_planes_act(const float* dy, const float* y, int act, long long rows, int C, int precision, void* planes,
                         void* stream);

/* Planes of the im2col (patch) matrix of x: [N*Ho*Wo][KP] in the layout of mog_split_planes, k = (kh*KW + kw)*C + c,
 * KP = KH*KW*C rounded up to 8; taps outside the image and the pad columns are zero.  For the thin ends of the networks
 * (RGB inputs of the discriminators / the image encoder, model.py:595-599,258; the 3-channel image gradient entering
 * GET_IMAGE_G, model.py:464-474): on the patch matrix the conv is a 1x1 conv with KP input channels
 * (mog_conv2d_fwd / mog_conv2d_wgrad with x == NULL and these planes). */
int mog_patch_planes(const float* x, int N, int H, int W, int C, int KH, int KW, int stride, int pad, int precision,
                     void* planes, void* stream);

/* The adjoint of mog_patch_planes with the conv epilogue (col2im): z is the fp32 output [N*Ho*Wo][ldz] of a 1x1 problem whose
 * column k = (kh*KW + kw)*C + c is the contribution of tap (kh, kw) to channel c;
 *   y[n,h,w,c] = act(bias[c] + sum of z[(n,ho,wo)][k] over the taps with (h + pad - kh, w + pad - kw) = stride * (ho, wo)).
 * C <= 4, KH*KW <= 16; Ho, Wo are the conv's output size for an H x W input.  Used for the data gradient of convs with <= 4
 * input channels (z = dy . w) and the forward of 'same' convs with <= 4 output channels (GET_IMAGE_G, model.py:464-474). */
int mog_col2im_act(const float* z, int ldz, int N, int H, int W, int C, int KH, int KW, int stride, int pad,
                   const float* bias, int act, float* y, void* stream);

/* replaces: nn.Conv2d forward incl. a preceding nn.Upsample(2,'nearest') (model.py:41-55,
 * 587-609,626,664-677; GlobalAttention.py:25-28) and nn.Linear (H=W=KH=KW=1; model.py:324,
 * 365,371).  y: [N,Ho,Wo,Cout]; bias may be NULL. */
int mog_conv_out_hw(const MogConvDesc* d, int* Ho, int* Wo);
size_t mog_conv_workspace_bytes(const MogConvDesc* d, int which /*0 fwd, 1 dgrad, 2 wgrad*/);
int mog_conv2d_fwd(const MogConvDesc* d, const float* x, const void* x_planes, const void* w_packed_fwd,
                   const float* bias, float* y, void* workspace, size_t ws_bytes, void* stream);
/* replaces: autograd of the above w.r.t. the input.  dy: [N,Ho,Wo,Cout] (already multiplied by
 * the epilogue derivative, see mog_act_bwd), dx: [N,H,W,Cin]. */
int mog_conv2d_dgrad(const MogConvDesc* d, const float* dy, const void* dy_planes, const void* w_packed_dgrad,
                     float* dx, void* workspace, size_t ws_bytes, void* stream);
/* replaces: autograd w.r.t. the weight; writes dw in OIHW (state_dict layout), deterministic
 * split reduction through the workspace.  dbias (may be NULL; needs the fp32 dy): [Cout]. */
int mog_conv2d_wgrad(const MogConvDesc* d, const float* x, const void* x_planes, const float* dy,
                     const void* dy_planes, float* dw_oihw, float* dbias, void* workspace, size_t ws_bytes,
                     void* stream);

/* ---- BatchNorm (train mode) + activation ------------------------------------------------ */
/* x is [S*M][C] (S segments of M rows; the object pathway calls the same BN once per object
 * with separate statistics, model.py:393-401,685-693).
 * replaces: nn.BatchNorm2d/1d train-mode forward (model.py:53,62,72,75,96,100,366,372,...). */
/* Reductions are two-stage and bit-reproducible: the streaming kernel leaves one fp64 partial per (block row p, quantity,
 * segment, channel) in part[P][2][S][C] (P = mog_bn_parts(S, M, C, act, which): which = 0 forward statistics, 1 backward
 * sums), the consumer adds the P partials in a fixed order.  The fp32 partial sums are taken relative to the channel's
 * first value, so the variance does not cancel for channels whose mean is large against their spread. */
int mog_bn_parts(int S, int M, int C, int act, int which);
int mog_bn_stats(const float* x, int S, int M, int C, double* part /*[P][2][S][C]: sum, sum of squares*/, int nparts,
                 void* stream);
/* Finalises statistics: mean, invstd [S][C]; scale = gamma*invstd, shift = beta - mean*scale;
 * running stats updated once per segment in order (momentum, unbiased variance) when not NULL. */
int mog_bn_finalize(const double* part, int nparts, int S, int M, int C, const float* gamma,
                    const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                    float* mean, float* invstd, float* scale, float* shift, void* stream);
/* y = act(x*scale + shift) (+ residual).  scale/shift may be NULL (plain activation).  For
 * MOG_ACT_GLU x has C channels and y (and residual) C/2.
 * replaces: BN apply + GLU / ReLU / LeakyReLU / residual add (model.py:24-32,54,63,73,80,...). */
int mog_affine_act_fwd(const float* x, const float* scale, const float* shift, const float* residual,
                       float* y, int S, int M, int C, int act, void* stream);
/* Same, and additionally emits y as the pre-split bf16 planes the tcgen05 convolutions read
 * (layout of mog_split_planes: hi [rows][C8], then lo for MOG_PREC_BF16X3), saving the separate
 * split pass over y when the consumer is a convolution.  Output channel count must be a multiple of 4. */
int mog_affine_act_fwd_planes(const float* x, const float* scale, const float* shift, const float* residual,
                              float* y, void* y_planes, int precision, int S, int M, int C, int act, void* stream);
/* Backward of BN(train)+act: pass 1 reduces dgamma=sum(dz*xhat), dbeta=sum(dz) per (segment,
 * channel) in double; pass 2 writes dx.  dy: [S*M][C or C/2].  With mean == invstd ==
 * dgamma_seg == dbeta_seg == NULL, mog_bn_act_bwd_apply is the backward of a plain activation
 * of the pre-activation x (dx = dy * act'(x); used for GLU without BN, model.py:328). */
int mog_bn_act_bwd_reduce(const float* x, const float* dy, const float* mean, const float* invstd,
                          const float* gamma, const float* beta, int S, int M, int C, int act,
                          double* part /*[P][2][S][C]: sum dz, sum dz*xhat*/, int nparts, double* dgamma_seg /*[S][C]*/,
                          double* dbeta_seg /*[S][C]*/, float* dgamma /*[C], summed over segments, or NULL*/, float* dbeta,
                          void* stream);
int mog_bn_act_bwd_apply(const float* x, const float* dy, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const double* dgamma_seg,
                         const double* dbeta_seg, int S, int M, int C, int act, float* dx, void* stream);
/* Same, and additionally emits dx as the pre-split bf16 planes (hi [rows][C], then lo for MOG_PREC_BF16X3) that the data /
 * weight gradient kernels of the producing convolution read, saving the separate split pass over dx.  C % 8 == 0. */
int mog_bn_act_bwd_apply_planes(const float* x, const float* dy, const float* mean, const float* invstd,
                                const float* gamma, const float* beta, const double* dgamma_seg,
                                const double* dbeta_seg, int S, int M, int C, int act, float* dx, void* dx_planes,
                                int precision, void* stream);
/* dz = dy * act'(.) given the activation OUTPUT y (LRELU / RELU / TANH / SIGMOID). */
int mog_act_bwd(const float* dy, const float* y, float* dz, size_t n, int act, void* stream);

/* ---- small elementwise helpers ----------------------------------------------------------- */
/* 2x2 sum pooling: backward of the nearest 2x upsample fused into mog_conv2d_fwd. */
int mog_sumpool2x2(const float* src /*[N,2H,2W,C]*/, float* dst /*[N,H,W,C]*/, int N, int H, int W, int C, void* stream);

/* ---- spatial transformer (bbox crop / scatter) ------------------------------------------- */
/* Bilinear sampling with an affine grid, zero padding, align_corners=0|1.
 * replaces: stn() = F.affine_grid + F.grid_sample (model.py:17-21) and the per-object loops
 * around it (model.py:107-112, 393-401, 685-693).
 *   mode 0 "scatter-sum": x [S*B,Hi,Wi,C] (segment-major), theta [B,S,2,3] -> y [B,Ho,Wo,C] = sum_s stn(x[s*B+b], theta[b,s])
 *   mode 1 "crop":        x [B,Hi,Wi,C], theta [B,S,2,3] -> y [S*B,Ho,Wo,Cy]; channels [0,C) sampled,
 *                         channels [C,Cy) filled with extra[b,s,:] (label planes, model.py:686-689; may be NULL, Cy==C) */
int mog_stn_fwd(const float* x, const float* theta, const float* extra, float* y, int mode, int B, int S,
                int Hi, int Wi, int C, int Ho, int Wo, int Cy, int align_corners, void* stream);
/* Input gradient of the above (gather form, deterministic; theta must be axis-aligned, i.e.
 * theta[0][1]==theta[1][0]==0 as produced by miscc/utils.py:16-49). dy has Cy channels per pixel. */
int mog_stn_bwd(const float* dy, const float* theta, float* dx, int mode, int B, int S,
                int Hi, int Wi, int C, int Ho, int Wo, int Cy, int align_corners, void* stream);

/* ---- word attention ----------------------------------------------------------------------- */
/* replaces: GlobalAttentionGeneral.forward after the 1x1 conv (GlobalAttention.py:95-123):
 * bmm -> mask -> softmax over words -> bmm.  h [B,Q,D] (NHWC pixels), src [B,T,D] (NHWC conv_context
 * output), mask [B,T] uint8 or NULL.  mask_quirk=1 reproduces the reference's tiled mask
 * (row (b*Q+q) uses mask[(b*Q+q) % B], GlobalAttention.py:104-108); 0 = per-sample mask.
 * out [B,Q,D]; attn (may be NULL) [B,T,Q] as returned by the reference. */
int mog_word_attention_fwd(const float* h, const float* src, const uint8_t* mask, float* out, float* attn,
                           int B, int Q, int D, int T, int mask_quirk, void* stream);
/* dh [B,Q,D]; dsrc [B,T,D].  With a workspace of mog_word_attention_bwd_workspace_bytes (D % 4 == 0) every block leaves its
 * partial [T][D] there and a second kernel sums them in a fixed order (deterministic; dsrc is written, not accumulated).
 * Without one (workspace NULL, or D % 4 != 0) dsrc must be zero-initialised by the caller and is accumulated atomically. */
size_t mog_word_attention_bwd_workspace_bytes(int B, int Q, int D, int T);
int mog_word_attention_bwd(const float* h, const float* src, const uint8_t* mask, const float* dout,
                           float* dh, float* dsrc, int B, int Q, int D, int T, int mask_quirk, void* workspace,
                           size_t ws_bytes, void* stream);

/* ---- DAMSM word-region matching (AttnGAN) ---------------------------------------------------- */
/* replaces: the per-caption loop of words_loss, each a func_attention call + cosine similarity +
 * log-sum-exp (miscc/losses.py:72-112, GlobalAttention.py:31-69).  ctx [B][R][D] region features
 * (NHWC), words [NI][D][Tw] (reference layout), lens [NI] int32.  sims[b][i] = log sum_t exp(gamma2 *
 * cos(word_t, attended context)); paired=1 evaluates only pairs (b, b) (func_attention's contract) and
 * then sims is [B]; wei_out [pairs][D][Tw] and attn_out [pairs][Tw][R] are optional outputs. */
int mog_damsm_words_fwd(const float* ctx, const float* words, const int* lens, float* sims, float* wei_out,
                        float* attn_out, int B, int NI, int R, int D, int Tw, int paired, float gamma1, float gamma2,
                        void* stream);
/* gradient w.r.t. the region features (the word embeddings come from the frozen text encoder).  One CTA per
 * (image, caption) pair writes its contribution into the caller's workspace [NI][B][R][D] (fp32,
 * mog_damsm_bwd_workspace_bytes); a second kernel sums the NI slices in a fixed order.  D % 4 == 0. */
size_t mog_damsm_bwd_workspace_bytes(int B, int NI, int R, int D);
int mog_damsm_words_bwd(const float* ctx, const float* words, const int* lens, const float* dsims, float* dctx,
                        int B, int NI, int R, int D, int Tw, float gamma1, float gamma2, void* workspace, size_t ws_bytes,
                        void* stream);

/* ---- loss heads --------------------------------------------------------------------------- */
/* loss[0] (+)= weight * mean_i BCE(sigmoid(z_i), target_i) with torch's log clamp at -100;
 * prob (may be NULL) receives sigmoid(z).  replaces: nn.Sigmoid + nn.BCELoss
 * (model.py:627, miscc/losses.py:156-171,195-200).  with_logits=1: nn.BCEWithLogitsLoss of the
 * StackGAN / CLEVR / Multi-MNIST programs (stackgan/miscc/utils.py:77, no clamp, stable form). */
int mog_sigmoid_bce_fwd(const float* z, const float* target /*[n]*/, float weight, int n, float* prob, float* loss,
                        int accumulate, int with_logits, void* stream);
/* dz_i = gscale[0] * weight * dBCE/dz_i / n */
int mog_sigmoid_bce_bwd(const float* z, const float* target /*[n]*/, float weight, int n, const float* gscale,
                        float* dz, int with_logits, void* stream);

/* ---- optimiser ----------------------------------------------------------------------------- */
/* Fused multi-tensor Adam step (+ optional exponential moving average of the updated parameters).
 * replaces: optim.Adam(betas=(0.5, 0.999)).step() per network (attngan/trainer.py:141-148,326,340;
 *   stackgan/trainer.py:136-137,218,234; multi-mnist/trainer.py:103-104; clevr/trainer.py:100-101) and the EMA
 *   loop `avg_p.mul_(0.999).add_(0.001, p.data)` (attngan/trainer.py:341-342).
 * p/g/m/v/ema: HOST arrays of n device pointers (fp32 tensors of numel[i] elements; ema == NULL or ema[i] == NULL: no EMA);
 * `step` is the 1-based index of the step being taken (bias corrections 1 - beta^step); gradients are multiplied by
 * grad_scale first (1/world after an all-reduce(sum)).  Same arithmetic as torch's Adam (no weight decay, no amsgrad):
 *   m = lerp(m, g, 1-beta1); v = beta2 v + (1-beta2) g^2; p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps);
 *   ema = ema_decay * ema + (1 - ema_decay) * p.
 * The pointer tables travel in kernel parameters (<= 56 tensors per launch): nothing is staged or allocated. */
int mog_adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* ema,
                   const long long* numel, double lr, double beta1, double beta2, double eps, long long step,
                   double ema_decay, float grad_scale, void* stream);
/* Same with the step count in DEVICE memory (one double, the number of steps taken so far): the call first increments
 * *step_dev, then forms the bias corrections from it on the device, in double, like the host form.  No argument changes from
 * step to step, so a CUDA graph that captured the call replays the whole optimiser step (trainer.py:326,340 inside the graph). */
int mog_adam_multi_dev(int n, float* const* p, const float* const* g, float* const* m, float* const* v, float* const* ema,
                       const long long* numel, double lr, double beta1, double beta2, double eps, double* step_dev,
                       double ema_decay, float grad_scale, void* stream);

/* ---- pooling / resize of the DAMSM image encoder -------------------------------------------- */
/* replaces: F.max_pool2d(x, 3, 2) (attngan/model.py:264,271; Mixed_6a/7a of torchvision's Inception-v3),
 *   F.avg_pool2d(x, 3, 1, 1) (Inception branch_pool; divisor k*k, count_include_pad), F.avg_pool2d(x, 8) (model.py:301)
 *   and their autograd.  NHWC fp32; mode 0 = max, 1 = average; output size floor((H + 2 pad - k) / stride) + 1.
 * Backward is a deterministic gather; for max pooling it recomputes each window's arg-max from the forward input x
 * with torch's tie rule (first maximum in scan order). */
int mog_pool2d_out_hw(int H, int W, int k, int stride, int pad, int* Ho, int* Wo);
int mog_pool2d_fwd(const float* x, float* y, unsigned char* argmax /*optional, max pooling: [N,Ho,Wo,C] window positions*/,
                   int N, int H, int W, int C, int k, int stride, int pad, int mode, void* stream);
int mog_pool2d_bwd(const float* x /*max pooling without argmax*/, const unsigned char* argmax /*or NULL*/, const float* dy,
                   float* dx, int N, int H, int W, int C, int k, int stride, int pad, int mode, void* stream);
/* replaces: nn.Upsample(size=(299, 299), mode='bilinear')(x) (attngan/model.py:256) = torch upsample_bilinear2d
 *   (source index scale*(o + 0.5) - 0.5 clamped at 0 for align_corners = 0) and its autograd (upsampling only). */
int mog_resize_bilinear_fwd(const float* x, float* y, int N, int Hi, int Wi, int C, int Ho, int Wo, int align_corners,
                            void* stream);
int mog_resize_bilinear_bwd(const float* dy, float* dx, int N, int Hi, int Wi, int C, int Ho, int Wo, int align_corners,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOG_H_ */
