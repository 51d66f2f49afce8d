// damsm.cu -- fused DAMSM word-region attention similarity (AttnGAN matching loss).
//
// replaces: the Python loop of words_loss over captions, each iteration a func_attention call
// (bmm -> softmax over words -> *gamma1 -> softmax over regions -> bmm) followed by a cosine
// similarity and a log-sum-exp (miscc/losses.py:72-112, GlobalAttention.py:31-69):
//
//   for image b, caption i (n_i words):
//     S[r,t]   = <ctx[b,r,:], w_i[:,t]>                       r: regions (17*17), t < n_i
//     a1[r,:]  = softmax_t S[r,:]
//     a2[t,:]  = softmax_r (gamma1 * a1[:,t])
//     v[:,t]   = sum_r a2[t,r] ctx[b,r,:]
//     sim[b,i] = log sum_t exp(gamma2 * cos(w_i[:,t], v[:,t]))
//
// One CTA per (image, caption) pair, forward and backward; the whole pair stays on chip (scores in shared
// memory, per-thread channel accumulators), ctx[b] is streamed from L2.  The backward of a pair writes its
// contribution to d ctx[b] into its own slice of a workspace [NI][B][R][D]; a second kernel sums the NI slices in
// a fixed order (deterministic, no atomics).  (A first version ran one CTA per image looping over the captions:
// 32 of 148 SMs busy, 22 ms at B = 32.)
// FLOPs are small (5.4 GFLOP at B=32); this kernel exists to remove ~15 launches x B iterations of latency.
#include "common.cuh"

namespace mog {

constexpr int DT = 256;     // threads
constexpr int TMAXW = 32;   // max words per caption
constexpr int MAXCPT = 2;   // channels per thread (template parameter CPT) -> D <= 512

struct DamsmArgs {
  const float* ctx;    // [B][R][D]   region features, NHWC
  const float* words;  // [NI][D][Tw] word embeddings (reference layout: batch x nef x seq_len)
  const int* lens;     // [NI]
  float* sims;         // [B][NI]
  float* wei_out;      // optional [B*NI][D][Tw]
  float* attn_out;     // optional [B*NI][Tw][R]
  const float* dsims;  // backward: [B][NI]
  float* dctx;         // backward: [B][R][D]
  int B, NI, R, D, Tw, paired;
  float g1, g2, eps;
};

// shared state of one (b, i) pair
struct PairSmem {
  float* w;    // [D][NT+2]   (pitch 2 x odd: see pair_scores)
  float* S;    // [R][TMAXW+1]  scores -> a1
  float* A2;   // [TMAXW][R+1]
  float* red;  // [3][TMAXW][DT/32 + 1]
  float* cosv; // [TMAXW] cos, [TMAXW] |w|, [TMAXW] |v|, [TMAXW] wv
};

// tp: word pitch of the layout (= the kernel's NT: Tw rounded up to 8; 18 words -> 24 keeps the forward at ~84 KB, two CTAs per SM)
// Region pitch of the [word][region] arrays: R rounded up to 4 so that four regions are one 128-bit shared-memory load.
__host__ __device__ __forceinline__ int lda_of(int R) { return (R + 3) & ~3; }
__host__ __device__ __forceinline__ size_t al4(size_t n) { return (n + 3) & ~size_t(3); }
__device__ __forceinline__ PairSmem carve(float* sm, int R, int D, int tp) {
  PairSmem p;
  p.w = sm;
  p.S = p.w + al4((size_t)D * (tp + 2));
  p.A2 = p.S + al4((size_t)R * (tp + 1));
  p.red = p.A2 + (size_t)tp * lda_of(R);
  p.cosv = p.red + al4(3 * tp * (DT / 32 + 1));
  return p;
}
static size_t pair_smem_bytes(int R, int D, int tp, bool bwd) {
  size_t f = al4((size_t)D * (tp + 2)) + al4((size_t)R * (tp + 1)) + (size_t)tp * lda_of(R) + al4(3 * tp * (DT / 32 + 1)) + 4 * tp;
  if (bwd) f += (size_t)tp * lda_of(R);   // dA (dv reuses the word-vector buffer)
  return sizeof(float) * f;
}

// out(r, t) = sum_c ctx[r][c] * m[c][t]  for all regions r and words t < n; m = word vectors (scores) or dv (d a2) in shared
// memory, pitch NT + 2.  One warp per FOUR regions; the two half-warps take the even / the odd words and 16 channels each per
// step, so a word-vector element read from shared memory feeds four FMAs (the loop was bound by shared-memory loads at one
// load per two FMAs: 24 wavefronts against 12 issue cycles per step).  Pitch NT + 2 = 2 x odd: the 16 channels of a half land
// in 16 distinct even banks, the other half (word + 1) in the odd ones -- conflict free.  The 16-lane sums are formed by
// shuffles (4 steps instead of 5).  transposed: store to out[t * ld + r] (d a2) instead of out[r * ld + t] (scores).
template <int NT, bool TRANSPOSED>
__device__ __forceinline__ void pair_scores(const float* __restrict__ ctx, const float* m, int R, int D, int n, float* out, int ld) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  constexpr int NU = NT / 2;
  const int ldw = NT + 2;
  for (int r = 4 * wrp; r < R; r += 4 * (DT / 32)) {
    float acc[4][NU];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int u = 0; u < NU; ++u) acc[q][u] = 0.f;
    const float* c0 = ctx + (size_t)r * D;
    const bool h1 = r + 1 < R, h2 = r + 2 < R, h3 = r + 3 < R;
#pragma unroll 2
    for (int c = hl; c < D; c += 16) {
      float x[4];
      x[0] = __ldg(c0 + c);
      x[1] = h1 ? __ldg(c0 + D + c) : 0.f;
      x[2] = h2 ? __ldg(c0 + 2 * D + c) : 0.f;
      x[3] = h3 ? __ldg(c0 + 3 * D + c) : 0.f;
      const float* mr = m + c * ldw + half;
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (2 * u < n) {        // (word 2u + half; a word n <= t < NT of the odd half multiplies stale data that is never stored)
          const float wv = mr[2 * u];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q][u] = fmaf(x[q], wv, acc[q][u]);
        }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      if (2 * u < n) {
        const int t = 2 * u + half;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float z = acc[q][u];
          z += __shfl_xor_sync(0xffffffffu, z, 8);
          z += __shfl_xor_sync(0xffffffffu, z, 4);
          z += __shfl_xor_sync(0xffffffffu, z, 2);
          z += __shfl_xor_sync(0xffffffffu, z, 1);
          if (hl == 0 && t < n && r + q < R) {
            if (TRANSPOSED) out[t * ld + r + q] = z;
            else out[(r + q) * ld + t] = z;
          }
        }
      }
    }
  }
}

// forward of one pair up to v (kept in registers: v[k][t] for channel c = tid + k*DT) and cos/|w|/|v|/wv in smem
// NT: compile-time bound of the word loops (Tw rounded up to 8): the unrolled loops issue NT, not TMAXW, predicated slots.
// The loops over regions / channel chunks that start with a global load are unrolled so that several loads are in flight
// (one CTA of 8 warps per SM: an exposed L2 round trip per iteration made a pair take 240 us).
template <int CPT, int NT>
__device__ void pair_forward(const DamsmArgs& a, const PairSmem& s, const float* ctx, const float* wi, int n,
                             float (&v)[CPT][NT]) {
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int ldw = NT + 2, ldS = NT + 1, ldA = lda_of(a.R);
  for (int i = tid; i < a.D * n; i += DT) {
    int c = i / n, t = i - c * n;
    s.w[c * ldw + t] = __ldg(wi + (size_t)c * a.Tw + t);
  }
  __syncthreads();
  pair_scores<NT, false>(ctx, s.w, a.R, a.D, n, s.S, ldS);
  __syncthreads();
  // a1 = softmax over words (per region)
  for (int r = tid; r < a.R; r += DT) {
    float mx = -INFINITY;
    for (int t = 0; t < n; ++t) mx = fmaxf(mx, s.S[r * ldS + t]);
    float sum = 0.f;
    for (int t = 0; t < n; ++t) {
      float e = expf(s.S[r * ldS + t] - mx);
      s.S[r * ldS + t] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int t = 0; t < n; ++t) s.S[r * ldS + t] *= inv;
  }
  __syncthreads();
  // a2 = softmax over regions of gamma1 * a1 (per word): one warp per word
  for (int t = wrp; t < n; t += DT / 32) {
    float mx = -INFINITY;
    for (int r = lane; r < a.R; r += 32) mx = fmaxf(mx, a.g1 * s.S[r * ldS + t]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int r = lane; r < a.R; r += 32) {
      float e = expf(a.g1 * s.S[r * ldS + t] - mx);
      s.A2[t * ldA + r] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int r = lane; r < a.R; r += 32) s.A2[t * ldA + r] *= inv;
    if (lane < ldA - a.R) s.A2[t * ldA + a.R + lane] = 0.f;     // pad regions: read by the 128-bit loads below
  }
  __syncthreads();
  // v[c][t] = sum_r a2[t][r] ctx[r][c]: thread per channel (coalesced over c)
#pragma unroll
  for (int k = 0; k < CPT; ++k)
#pragma unroll
    for (int t = 0; t < NT; ++t) v[k][t] = 0.f;
#pragma unroll 2
  for (int r = 0; r < a.R; r += 4) {     // four regions per step: one 128-bit load of a2[t][r..r+3] feeds four FMAs
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * DT;
      if (c < a.D) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = (r + u < a.R) ? __ldg(ctx + (size_t)(r + u) * a.D + c) : 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t)
          if (t < n) {
            const float4 q = *reinterpret_cast<const float4*>(s.A2 + t * ldA + r);
            v[k][t] = fmaf(x[0], q.x, fmaf(x[1], q.y, fmaf(x[2], q.z, fmaf(x[3], q.w, v[k][t]))));
          }
      }
    }
  }
  // cos(w_t, v_t): block reductions of <w,v>, |w|^2, |v|^2 over channels
  const int ldr = DT / 32 + 1;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    if (t < n) {
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int c = tid + k * DT;
        if (c < a.D) {
          const float wv = s.w[c * ldw + t];
          p0 = fmaf(wv, v[k][t], p0);
          p1 = fmaf(wv, wv, p1);
          p2 = fmaf(v[k][t], v[k][t], p2);
        }
      }
      p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2);
      if (lane == 0) {
        s.red[(0 * NT + t) * ldr + wrp] = p0;
        s.red[(1 * NT + t) * ldr + wrp] = p1;
        s.red[(2 * NT + t) * ldr + wrp] = p2;
      }
    }
  }
  __syncthreads();
  if (tid < n) {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    for (int j = 0; j < DT / 32; ++j) {
      p0 += s.red[(0 * NT + tid) * ldr + j];
      p1 += s.red[(1 * NT + tid) * ldr + j];
      p2 += s.red[(2 * NT + tid) * ldr + j];
    }
    const float nw = sqrtf(p1), nv = sqrtf(p2);
    s.cosv[tid] = p0 / fmaxf(nw * nv, a.eps);   // miscc/losses.py:11-17
    s.cosv[NT + tid] = nw;
    s.cosv[2 * NT + tid] = nv;
    s.cosv[3 * NT + tid] = p0;
  }
  __syncthreads();
}

template <int CPT, int NT>
__global__ void __launch_bounds__(DT) damsm_fwd_kernel(DamsmArgs a) {
  extern __shared__ float sm[];
  const PairSmem s = carve(sm, a.R, a.D, NT);
  const int b = blockIdx.x, i = a.paired ? blockIdx.x : blockIdx.y;
  const int n = min(a.lens[i], a.Tw);
  const float* ctx = a.ctx + (size_t)b * a.R * a.D;
  const float* wi = a.words + (size_t)i * a.D * a.Tw;
  float v[CPT][NT];
  pair_forward<CPT, NT>(a, s, ctx, wi, n, v);
  const int tid = threadIdx.x;
  const size_t pair = a.paired ? (size_t)b : (size_t)b * a.NI + i;
  if (tid == 0 && a.sims) {
    float sum = 0.f;
    for (int t = 0; t < n; ++t) sum += expf(a.g2 * s.cosv[t]);
    a.sims[pair] = logf(sum);
  }
  if (a.wei_out) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * DT;
      if (c < a.D) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
          if (t < n) a.wei_out[(pair * a.D + c) * a.Tw + t] = v[k][t];
      }
    }
  }
  if (a.attn_out) {
    const int ldA = lda_of(a.R);
    for (int j = tid; j < n * a.R; j += DT) {
      int t = j / a.R, r = j - t * a.R;
      a.attn_out[(pair * a.Tw + t) * a.R + r] = s.A2[t * ldA + r];
    }
  }
}

// backward w.r.t. ctx: one CTA per (image b, caption i); its d ctx[b] contribution goes to slice i of the workspace
// (a.dctx = workspace [NI][B][R][D])
template <int CPT, int NT>
__global__ void __launch_bounds__(DT, CPT == 1 && NT <= 24 ? 2 : 1) damsm_bwd_kernel(DamsmArgs a) {
  extern __shared__ float sm[];
  const PairSmem s = carve(sm, a.R, a.D, NT);
  float* dA = s.cosv + 4 * NT;   // [NT][R+1]: d a2 -> d(gamma1*a1) -> reused
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const float* ctx = a.ctx + (size_t)b * a.R * a.D;
  const int ldw = NT + 2, ldS = NT + 1, ldA = lda_of(a.R);
  {
    const int i = blockIdx.y;
    float* dctx = a.dctx + ((size_t)i * a.B + b) * a.R * a.D;
    const float gsim = a.dsims[(size_t)b * a.NI + i];
    const int n = min(a.lens[i], a.Tw);
    const float* wi = a.words + (size_t)i * a.D * a.Tw;
    float v[CPT][NT];
    __syncthreads();
    pair_forward<CPT, NT>(a, s, ctx, wi, n, v);
    // d cos_t = gsim * gamma2 * softmax_t(gamma2 cos);  dv_t[c] = dcos_t (w_t/(|w||v|) - cos_t v_t/|v|^2)
    float esum = 0.f;
    for (int t = 0; t < n; ++t) esum += expf(a.g2 * s.cosv[t]);
    float dv[CPT][NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      float dcos = 0.f, inv_wv = 0.f, c_over_v2 = 0.f;
      if (t < n) {
        dcos = gsim * a.g2 * expf(a.g2 * s.cosv[t]) / esum;
        const float nw = s.cosv[NT + t], nv = s.cosv[2 * NT + t];
        if (nw * nv > a.eps) {   // clamp inactive (always, in practice)
          inv_wv = 1.f / (nw * nv);
          c_over_v2 = s.cosv[t] / (nv * nv);
        }
      }
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int c = tid + k * DT;
        dv[k][t] = (t < n && c < a.D) ? dcos * (s.w[c * ldw + t] * inv_wv - c_over_v2 * v[k][t]) : 0.f;
      }
    }
    // dv goes to shared memory for the d a2 dot products below (lanes over channels).  It takes the place of the word vectors:
    // a thread reads and writes only its own channel rows here, and from now on nobody reads another row of w -- the
    // final phase needs just the thread's own row, kept in registers (wr).  Without a separate dv buffer the pair state
    // is 111 KB: two CTAs per SM.
    float wr[CPT][NT];
    float* dvs = s.w;   // [D][NT+1]
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * DT;
#pragma unroll
      for (int t = 0; t < NT; ++t) {     // (static indices: registers)
        wr[k][t] = (t < n && c < a.D) ? s.w[c * ldw + t] : 0.f;
        if (t < n && c < a.D) dvs[c * ldw + t] = dv[k][t];
      }
    }
    __syncthreads();
    pair_scores<NT, true>(ctx, dvs, a.R, a.D, n, dA, ldA);   // d a2[t][r] = sum_c dv_t[c] ctx[r][c]
    __syncthreads();
    // softmax-over-regions backward: d(gamma1 a1[r][t]) = a2 (da2 - <a2, da2>)  -> da1 = gamma1 * that
    for (int t = wrp; t < n; t += DT / 32) {
      float dot = 0.f;
      for (int r = lane; r < a.R; r += 32) dot = fmaf(s.A2[t * ldA + r], dA[t * ldA + r], dot);
      dot = warp_sum(dot);
      for (int r = lane; r < a.R; r += 32)
        dA[t * ldA + r] = a.g1 * s.A2[t * ldA + r] * (dA[t * ldA + r] - dot);   // = d a1[r][t]
    }
    __syncthreads();
    // softmax-over-words backward (per region): dS[r][t] = a1 (da1 - <a1, da1>), stored back into dA[t][r]
    for (int r = tid; r < a.R; r += DT) {
      float dot = 0.f;
      for (int t = 0; t < n; ++t) dot = fmaf(s.S[r * ldS + t], dA[t * ldA + r], dot);
      for (int t = 0; t < n; ++t) dA[t * ldA + r] = s.S[r * ldS + t] * (dA[t * ldA + r] - dot);
    }
    __syncthreads();
    // d ctx[r][c] = sum_t dv_t[c] a2[t][r] + dS[r][t] w[c][t]: four regions per step (128-bit loads of a2 / dS), the
    // thread's word-vector row w[c][:] in registers, its dv row read back from shared memory
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int c = tid + k * DT;
      if (c < a.D) {
        float dvr[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) dvr[t] = t < n ? dvs[c * ldw + t] : 0.f;
        for (int r = 0; r < a.R; r += 4) {
          float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll
          for (int t = 0; t < NT; ++t)
            if (t < n) {
              const float4 q = *reinterpret_cast<const float4*>(s.A2 + t * ldA + r);
              const float4 e = *reinterpret_cast<const float4*>(dA + t * ldA + r);
              g0 = fmaf(dvr[t], q.x, fmaf(e.x, wr[k][t], g0));
              g1 = fmaf(dvr[t], q.y, fmaf(e.y, wr[k][t], g1));
              g2 = fmaf(dvr[t], q.z, fmaf(e.z, wr[k][t], g2));
              g3 = fmaf(dvr[t], q.w, fmaf(e.w, wr[k][t], g3));
            }
          dctx[(size_t)r * a.D + c] = g0;
          if (r + 1 < a.R) dctx[(size_t)(r + 1) * a.D + c] = g1;
          if (r + 2 < a.R) dctx[(size_t)(r + 2) * a.D + c] = g2;
          if (r + 3 < a.R) dctx[(size_t)(r + 3) * a.D + c] = g3;
        }
      }
    }
  }
}

// d ctx[e] = sum_i part[i][e], i ascending (fixed order)
__global__ void damsm_bwd_sum_kernel(const float* __restrict__ part, float* __restrict__ dctx, int NI, size_t n4) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n4) return;
  float4 acc = __ldg(reinterpret_cast<const float4*>(part) + e);
  for (int i = 1; i < NI; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part) + (size_t)i * n4 + e);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(dctx)[e] = acc;
}

}  // namespace mog

using namespace mog;

static int check_damsm(int B, int NI, int R, int D, int Tw, const char* who) {
  MOG_REQUIRE(B > 0 && NI > 0 && R > 0 && D > 0 && Tw > 0, "%s: non-positive dims", who);
  MOG_REQUIRE(Tw <= TMAXW && D <= DT * MAXCPT && R <= 2048, "%s: Tw=%d (<=32), D=%d (<=512), R=%d (<=2048) out of range", who, Tw, D, R);
  MOG_REQUIRE(B <= 65535 && NI <= 65535, "%s: batch too large", who);
  return MOG_OK;
}

extern "C" int mog_damsm_words_fwd(const float* ctx, const float* words, const int* lens, float* sims, float* wei_out,
                                   float* attn_out, int B, int NI, int R, int D, int Tw, int paired, float gamma1,
                                   float gamma2, void* stream) {
  int rc = check_damsm(B, NI, R, D, Tw, "mog_damsm_words_fwd");
  if (rc) return rc;
  MOG_REQUIRE(ctx && words && lens && (sims || wei_out || attn_out), "mog_damsm_words_fwd: null tensor");
  MOG_REQUIRE(!paired || B == NI, "mog_damsm_words_fwd: paired mode needs B == NI");
  DamsmArgs a{ctx, words, lens, sims, wei_out, attn_out, nullptr, nullptr, B, NI, R, D, Tw, paired, gamma1, gamma2, 1e-8f};
  const size_t smem = pair_smem_bytes(R, D, ((Tw - 1) / 8 + 1) * 8, false);
  typedef void (*Kern)(DamsmArgs);
  static const Kern table[2][4] = {{damsm_fwd_kernel<1, 8>, damsm_fwd_kernel<1, 16>, damsm_fwd_kernel<1, 24>, damsm_fwd_kernel<1, 32>},
                                   {damsm_fwd_kernel<2, 8>, damsm_fwd_kernel<2, 16>, damsm_fwd_kernel<2, 24>, damsm_fwd_kernel<2, 32>}};
  Kern kern = table[D <= DT ? 0 : 1][(Tw - 1) / 8];
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MOG_ERR_UNSUPPORTED, "mog_damsm_words_fwd: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
  dim3 grid(B, paired ? 1 : NI);
  kern<<<grid, DT, smem, as_stream(stream)>>>(a);
  return check_launch("damsm_fwd_kernel");
}

extern "C" size_t mog_damsm_bwd_workspace_bytes(int B, int NI, int R, int D) {
  if (B <= 0 || NI <= 0 || R <= 0 || D <= 0) return 0;
  return sizeof(float) * (size_t)NI * B * R * D;
}

extern "C" int mog_damsm_words_bwd(const float* ctx, const float* words, const int* lens, const float* dsims, float* dctx,
                                   int B, int NI, int R, int D, int Tw, float gamma1, float gamma2, void* workspace,
                                   size_t ws_bytes, void* stream) {
  int rc = check_damsm(B, NI, R, D, Tw, "mog_damsm_words_bwd");
  if (rc) return rc;
  MOG_REQUIRE(ctx && words && lens && dsims && dctx, "mog_damsm_words_bwd: null tensor");
  MOG_REQUIRE((D & 3) == 0, "mog_damsm_words_bwd: D must be a multiple of 4");
  const size_t need = mog_damsm_bwd_workspace_bytes(B, NI, R, D);
  if (!workspace || ws_bytes < need) return fail(MOG_ERR_WORKSPACE, "mog_damsm_words_bwd: workspace %zu < %zu", ws_bytes, need);
  DamsmArgs a{ctx, words, lens, nullptr, nullptr, nullptr, dsims, static_cast<float*>(workspace), B, NI, R, D, Tw, 0, gamma1, gamma2, 1e-8f};
  const size_t smem = pair_smem_bytes(R, D, ((Tw - 1) / 8 + 1) * 8, true);
  typedef void (*Kern)(DamsmArgs);
  static const Kern table[2][4] = {{damsm_bwd_kernel<1, 8>, damsm_bwd_kernel<1, 16>, damsm_bwd_kernel<1, 24>, damsm_bwd_kernel<1, 32>},
                                   {damsm_bwd_kernel<2, 8>, damsm_bwd_kernel<2, 16>, damsm_bwd_kernel<2, 24>, damsm_bwd_kernel<2, 32>}};
  Kern kern = table[D <= DT ? 0 : 1][(Tw - 1) / 8];
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(MOG_ERR_UNSUPPORTED, "mog_damsm_words_bwd: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
  cudaStream_t st = as_stream(stream);
  kern<<<dim3(B, NI), DT, smem, st>>>(a);
  rc = check_launch("damsm_bwd_kernel");
  if (rc) return rc;
  const size_t n4 = (size_t)B * R * D / 4;
  damsm_bwd_sum_kernel<<<(unsigned)ceil_div_ll((long long)n4, 256), 256, 0, st>>>(static_cast<const float*>(workspace), dctx, NI, n4);
  return check_launch("damsm_bwd_sum_kernel");
}
