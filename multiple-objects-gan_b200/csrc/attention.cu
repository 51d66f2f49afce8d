// attention.cu -- fused word attention of the generator's refinement stages
// (GlobalAttention.py:95-123): per pixel q of sample b
//     score[t] = <h[b,q,:], src[b,t,:]>,  masked softmax over the T words,  out[b,q,:] = sum_t attn[t] src[b,t,:]
// in ONE pass over h (NHWC pixels: the reference's three .contiguous() transposes disappear).
// FLOPs are negligible (2*2*D*T per pixel); the kernel is HBM-bound: it reads h once
// (D floats/pixel) and writes out (D floats/pixel) [+ the attention map if requested].
// Each block owns 128 consecutive pixels of one sample; src (T x D) and the pixel tile are
// staged in shared memory, one thread per pixel keeps the T scores in registers.
#include "common.cuh"

namespace mog {

constexpr int APIX = 128;

struct AttnArgs {
  const float* h; const float* src; const uint8_t* mask; const float* dout;
  float* out; float* attn; float* dh; float* dsrc;
  int B, Q, D, T, quirk;
};

template <int TMAX>
__device__ __forceinline__ void scores_softmax(const AttnArgs& a, const float* hs, const float* srcs, int b, int q,
                                               bool valid, float (&p)[TMAX]) {
  const int ld = a.D + 1;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) p[t] = 0.f;
  const float* hr = hs + threadIdx.x * ld;
  for (int c = 0; c < a.D; ++c) {
    const float hv = hr[c];
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) p[t] = fmaf(hv, srcs[t * a.D + c], p[t]);
  }
  // mask row: the reference tiles the (B,T) mask queryL times against a batch-major view
  // (GlobalAttention.py:104-108) => row (b*Q+q) uses mask[(b*Q+q) % B].
  int mb = b;
  if (a.quirk) mb = (int)(((long long)b * a.Q + q) % a.B);
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < a.T) {
      if (a.mask && valid && a.mask[(size_t)mb * a.T + t]) p[t] = -INFINITY;
      mx = fmaxf(mx, p[t]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < a.T) {
      p[t] = expf(p[t] - mx);
      sum += p[t];
    }
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < a.T) p[t] *= inv;
}

__device__ __forceinline__ void load_tile(const float* __restrict__ g, float* s, int rows_valid, int D) {
  // g: [rows][D] contiguous -> s: [APIX][D+1]
  const int n = rows_valid * D;
  for (int i = threadIdx.x; i < n; i += APIX) s[(i / D) * (D + 1) + (i % D)] = __ldg(g + i);
}
__device__ __forceinline__ void store_tile(float* __restrict__ g, const float* s, int rows_valid, int D) {
  const int n = rows_valid * D;
  for (int i = threadIdx.x; i < n; i += APIX) g[i] = s[(i / D) * (D + 1) + (i % D)];
}

template <int TMAX>
__global__ void __launch_bounds__(APIX) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  float* srcs = sm;                 // [T][D]
  float* hs = sm + a.T * a.D;       // [APIX][D+1]
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * APIX;
  const int rows = min(APIX, a.Q - q0);
  for (int i = threadIdx.x; i < a.T * a.D; i += APIX) srcs[i] = __ldg(a.src + (size_t)b * a.T * a.D + i);
  load_tile(a.h + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D);
  __syncthreads();
  const int q = q0 + threadIdx.x;
  const bool valid = threadIdx.x < rows;
  float p[TMAX];
  scores_softmax<TMAX>(a, hs, srcs, b, q, valid, p);
  if (a.attn && valid) {
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) a.attn[((size_t)b * a.T + t) * a.Q + q] = p[t];
  }
  // weighted context, written back through the (now free) pixel row of this thread
  float* hr = hs + threadIdx.x * (a.D + 1);
  for (int c = 0; c < a.D; ++c) {
    float o = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) o = fmaf(p[t], srcs[t * a.D + c], o);
    hr[c] = o;
  }
  __syncthreads();
  store_tile(a.out + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D);
}

template <int TMAX>
__global__ void __launch_bounds__(APIX) attn_bwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  const int ld = a.D + 1, lt = TMAX + 1;
  float* srcs = sm;                       // [T][D]
  float* hs = srcs + a.T * a.D;           // [APIX][D+1]
  float* gs = hs + APIX * ld;             // [APIX][D+1]  dout tile, later dh tile
  float* ps = gs + APIX * ld;             // [APIX][TMAX+1] attn
  float* ds = ps + APIX * lt;             // [APIX][TMAX+1] dscore
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * APIX;
  const int rows = min(APIX, a.Q - q0);
  for (int i = threadIdx.x; i < a.T * a.D; i += APIX) srcs[i] = __ldg(a.src + (size_t)b * a.T * a.D + i);
  load_tile(a.h + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D);
  load_tile(a.dout + ((size_t)b * a.Q + q0) * a.D, gs, rows, a.D);
  __syncthreads();
  const int q = q0 + threadIdx.x;
  const bool valid = threadIdx.x < rows;
  float p[TMAX];
  scores_softmax<TMAX>(a, hs, srcs, b, q, valid, p);
  // dattn[t] = <dout, src[t]>;  dscore = p * (dattn - <p, dattn>)
  float da[TMAX];
#pragma unroll
  for (int t = 0; t < TMAX; ++t) da[t] = 0.f;
  const float* gr = gs + threadIdx.x * ld;
  for (int c = 0; c < a.D; ++c) {
    const float gv = gr[c];
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) da[t] = fmaf(gv, srcs[t * a.D + c], da[t]);
  }
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < a.T) dot = fmaf(p[t], da[t], dot);
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    float v = 0.f, pv = 0.f;
    if (t < a.T && valid) {
      pv = p[t];
      v = pv * (da[t] - dot);
    }
    da[t] = v;
    ps[threadIdx.x * lt + t] = pv;
    ds[threadIdx.x * lt + t] = v;
  }
  __syncthreads();  // ps/ds complete; gs (dout) still needed for dsrc below
  // dsrc[t][c] += sum_q dout[q][c]*attn[q][t] + h[q][c]*dscore[q][t]   (block partial -> atomics)
  for (int i = threadIdx.x; i < a.T * a.D; i += APIX) {
    const int t = i / a.D, c = i % a.D;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r)
      acc = fmaf(gs[r * ld + c], ps[r * lt + t], fmaf(hs[r * ld + c], ds[r * lt + t], acc));
    atomicAdd(a.dsrc + (size_t)b * a.T * a.D + i, acc);
  }
  __syncthreads();
  // dh[q][c] = sum_t dscore[t] * src[t][c]  (overwrites the dout tile row of this thread)
  float* orow = gs + threadIdx.x * ld;
  for (int c = 0; c < a.D; ++c) {
    float o = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) o = fmaf(da[t], srcs[t * a.D + c], o);
    orow[c] = o;
  }
  __syncthreads();
  store_tile(a.dh + ((size_t)b * a.Q + q0) * a.D, gs, rows, a.D);
}


// ---------------------------------------------------------------------------------------------
// Vectorised kernels (D % 4 == 0: every configuration of the reference).  The scalar kernels above issued one
// shared-memory load per FMA (measured: 11 % / 5 % of the HBM roofline forward / backward at stage 3); here every
// shared-memory access is a 128-bit load feeding 4 FMAs (src rows are warp-uniform broadcasts, the pixel row of a
// thread sits at a pitch of D + 4 floats: conflict-free for 128-bit accesses), tiles move with 128-bit global accesses,
// and the d src reduction of the backward leaves the block as a partial [T][D] in a workspace that a second kernel sums in
// a fixed order (deterministic; the float atomics of the first version were not).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile4(const float* __restrict__ g, float* s, int rows_valid, int D, int pitch) {
  const int d4 = D >> 2, n4 = rows_valid * d4;
  for (int i = threadIdx.x; i < n4; i += APIX) {
    const int r = i / d4, c = i - r * d4;
    *reinterpret_cast<float4*>(s + r * pitch + 4 * c) = __ldg(reinterpret_cast<const float4*>(g) + i);
  }
}
__device__ __forceinline__ void store_tile4(float* __restrict__ g, const float* s, int rows_valid, int D, int pitch) {
  const int d4 = D >> 2, n4 = rows_valid * d4;
  for (int i = threadIdx.x; i < n4; i += APIX) {
    const int r = i / d4, c = i - r * d4;
    reinterpret_cast<float4*>(g)[i] = *reinterpret_cast<const float4*>(s + r * pitch + 4 * c);
  }
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}

// p[t] = <row, src[t]> for the thread's row (shared memory, pitch-aligned), then the masked softmax
template <int TMAX>
__device__ __forceinline__ void scores_softmax4(const AttnArgs& a, const float* row, const float* srcs, int b, int q, bool valid,
                                                float (&p)[TMAX]) {
#pragma unroll
  for (int t = 0; t < TMAX; ++t) p[t] = 0.f;
  const int d4 = a.D >> 2;
  for (int c = 0; c < d4; ++c) {
    const float4 hv = *reinterpret_cast<const float4*>(row + 4 * c);
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) p[t] = dot4(hv, *reinterpret_cast<const float4*>(srcs + t * a.D + 4 * c), p[t]);
  }
  int mb = b;
  if (a.quirk) mb = (int)(((long long)b * a.Q + q) % a.B);
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < a.T) {
      if (a.mask && valid && a.mask[(size_t)mb * a.T + t]) p[t] = -INFINITY;
      mx = fmaxf(mx, p[t]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < a.T) {
      p[t] = expf(p[t] - mx);
      sum += p[t];
    }
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < a.T) p[t] *= inv;
}

// row[c] = sum_t w[t] * src[t][c]
template <int TMAX>
__device__ __forceinline__ void weighted_rows4(const AttnArgs& a, const float (&w)[TMAX], const float* srcs, float* row) {
  const int d4 = a.D >> 2;
  for (int c = 0; c < d4; ++c) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) {
        const float4 sv = *reinterpret_cast<const float4*>(srcs + t * a.D + 4 * c);
        o.x = fmaf(w[t], sv.x, o.x); o.y = fmaf(w[t], sv.y, o.y); o.z = fmaf(w[t], sv.z, o.z); o.w = fmaf(w[t], sv.w, o.w);
      }
    *reinterpret_cast<float4*>(row + 4 * c) = o;
  }
}

template <int TMAX>
__global__ void __launch_bounds__(APIX) attn_fwd4_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int pitch = a.D + 4;
  float* srcs = sm;                 // [T][D]
  float* hs = sm + a.T * a.D;       // [APIX][D+4]
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * APIX;
  const int rows = min(APIX, a.Q - q0);
  for (int i = threadIdx.x; i < (a.T * a.D) >> 2; i += APIX)
    reinterpret_cast<float4*>(srcs)[i] = __ldg(reinterpret_cast<const float4*>(a.src + (size_t)b * a.T * a.D) + i);
  load_tile4(a.h + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D, pitch);
  __syncthreads();
  const int q = q0 + threadIdx.x;
  const bool valid = threadIdx.x < rows;
  float p[TMAX];
  float* row = hs + threadIdx.x * pitch;
  scores_softmax4<TMAX>(a, row, srcs, b, q, valid, p);
  if (a.attn && valid) {
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
      if (t < a.T) a.attn[((size_t)b * a.T + t) * a.Q + q] = p[t];
  }
  weighted_rows4<TMAX>(a, p, srcs, row);     // weighted context, through the (now free) pixel row of this thread
  __syncthreads();
  store_tile4(a.out + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D, pitch);
}

// partial: [B][gridDim.x][T][D]
template <int TMAX>
__global__ void __launch_bounds__(APIX) attn_bwd4_kernel(AttnArgs a, float* __restrict__ partial) {
  extern __shared__ __align__(16) float sm[];
  const int pitch = a.D + 4, lt = TMAX + 1;
  float* srcs = sm;                       // [T][D]
  float* hs = srcs + a.T * a.D;           // [APIX][D+4]
  float* gs = hs + APIX * pitch;          // [APIX][D+4]  dout tile, later dh tile
  float* ps = gs + APIX * pitch;          // [APIX][TMAX+1] attn
  float* ds = ps + APIX * lt;             // [APIX][TMAX+1] dscore
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * APIX;
  const int rows = min(APIX, a.Q - q0);
  for (int i = threadIdx.x; i < (a.T * a.D) >> 2; i += APIX)
    reinterpret_cast<float4*>(srcs)[i] = __ldg(reinterpret_cast<const float4*>(a.src + (size_t)b * a.T * a.D) + i);
  load_tile4(a.h + ((size_t)b * a.Q + q0) * a.D, hs, rows, a.D, pitch);
  load_tile4(a.dout + ((size_t)b * a.Q + q0) * a.D, gs, rows, a.D, pitch);
  __syncthreads();
  const int q = q0 + threadIdx.x;
  const bool valid = threadIdx.x < rows;
  float p[TMAX];
  scores_softmax4<TMAX>(a, hs + threadIdx.x * pitch, srcs, b, q, valid, p);
  // dattn[t] = <dout, src[t]>;  dscore = p * (dattn - <p, dattn>)
  float da[TMAX];
#pragma unroll
  for (int t = 0; t < TMAX; ++t) da[t] = 0.f;
  {
    const float* gr = gs + threadIdx.x * pitch;
    const int d4 = a.D >> 2;
    for (int c = 0; c < d4; ++c) {
      const float4 gv = *reinterpret_cast<const float4*>(gr + 4 * c);
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < a.T) da[t] = dot4(gv, *reinterpret_cast<const float4*>(srcs + t * a.D + 4 * c), da[t]);
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t)
    if (t < a.T) dot = fmaf(p[t], da[t], dot);
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    float v = 0.f, pv = 0.f;
    if (t < a.T && valid) {
      pv = p[t];
      v = pv * (da[t] - dot);
    }
    da[t] = v;
    ps[threadIdx.x * lt + t] = pv;
    ds[threadIdx.x * lt + t] = v;
  }
  __syncthreads();  // ps/ds complete; gs (dout) still needed for dsrc below
  // block partial of dsrc[t][c4] = sum_q dout[q][c4]*attn[q][t] + h[q][c4]*dscore[q][t]: one (t, 4 channels) item per thread
  {
    const int d4 = a.D >> 2, items = a.T * d4;
    float* dstp = partial + ((size_t)b * gridDim.x + blockIdx.x) * a.T * a.D;
    for (int i = threadIdx.x; i < items; i += APIX) {
      const int t = i / d4, c = i - t * d4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        const float4 gv = *reinterpret_cast<const float4*>(gs + r * pitch + 4 * c);
        const float4 hv = *reinterpret_cast<const float4*>(hs + r * pitch + 4 * c);
        const float pv = ps[r * lt + t], dv = ds[r * lt + t];
        acc.x = fmaf(gv.x, pv, fmaf(hv.x, dv, acc.x)); acc.y = fmaf(gv.y, pv, fmaf(hv.y, dv, acc.y));
        acc.z = fmaf(gv.z, pv, fmaf(hv.z, dv, acc.z)); acc.w = fmaf(gv.w, pv, fmaf(hv.w, dv, acc.w));
      }
      *reinterpret_cast<float4*>(dstp + t * a.D + 4 * c) = acc;
    }
  }
  __syncthreads();
  // dh[q][c] = sum_t dscore[t] * src[t][c]  (overwrites the dout tile row of this thread)
  weighted_rows4<TMAX>(a, da, srcs, gs + threadIdx.x * pitch);
  __syncthreads();
  store_tile4(a.dh + ((size_t)b * a.Q + q0) * a.D, gs, rows, a.D, pitch);
}

// dsrc[b][i] = sum over the nblk block partials, in block order
__global__ void attn_dsrc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dsrc, int nblk, int TD) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TD) return;
  const float* p = partial + (size_t)b * nblk * TD + i;
  float acc = 0.f;
  for (int k = 0; k < nblk; ++k) acc += p[(size_t)k * TD];
  dsrc[(size_t)b * TD + i] = acc;
}

template <int TMAX>
int launch_attn(const AttnArgs& a, bool bwd, float* partial, cudaStream_t st) {
  dim3 grid(ceil_div(a.Q, APIX), a.B);
  if ((a.D & 3) == 0 && (!bwd || partial)) {
    if (!bwd) {
      size_t smem = sizeof(float) * ((size_t)a.T * a.D + (size_t)APIX * (a.D + 4));
      cudaFuncSetAttribute(attn_fwd4_kernel<TMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attn_fwd4_kernel<TMAX><<<grid, APIX, smem, st>>>(a);
      return check_launch("attn_fwd4_kernel");
    }
    size_t smem = sizeof(float) * ((size_t)a.T * a.D + 2 * (size_t)APIX * (a.D + 4) + 2 * (size_t)APIX * (TMAX + 1));
    cudaFuncSetAttribute(attn_bwd4_kernel<TMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_bwd4_kernel<TMAX><<<grid, APIX, smem, st>>>(a, partial);
    int rc = check_launch("attn_bwd4_kernel");
    if (rc) return rc;
    const int TD = a.T * a.D;
    attn_dsrc_reduce_kernel<<<dim3(ceil_div(TD, 128), a.B), 128, 0, st>>>(partial, a.dsrc, grid.x, TD);
    return check_launch("attn_dsrc_reduce_kernel");
  }
  if (!bwd) {
    size_t smem = sizeof(float) * ((size_t)a.T * a.D + (size_t)APIX * (a.D + 1));
    cudaFuncSetAttribute(attn_fwd_kernel<TMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn_fwd_kernel<TMAX><<<grid, APIX, smem, st>>>(a);
    return check_launch("attn_fwd_kernel");
  }
  size_t smem = sizeof(float) * ((size_t)a.T * a.D + 2 * (size_t)APIX * (a.D + 1) + 2 * (size_t)APIX * (TMAX + 1));
  cudaFuncSetAttribute(attn_bwd_kernel<TMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_bwd_kernel<TMAX><<<grid, APIX, smem, st>>>(a);
  return check_launch("attn_bwd_kernel");
}

static int dispatch(const AttnArgs& a, bool bwd, float* partial, cudaStream_t st) {
  if (a.T <= 8) return launch_attn<8>(a, bwd, partial, st);
  if (a.T <= 16) return launch_attn<16>(a, bwd, partial, st);
  if (a.T <= 24) return launch_attn<24>(a, bwd, partial, st);
  return launch_attn<32>(a, bwd, partial, st);
}

}  // namespace mog

using namespace mog;

extern "C" int mog_word_attention_fwd(const float* h, const float* src, const uint8_t* mask, float* out, float* attn,
                                      int B, int Q, int D, int T, int mask_quirk, void* stream) {
  MOG_REQUIRE(h && src && out, "mog_word_attention_fwd: null tensor");
  MOG_REQUIRE(B > 0 && Q > 0 && D > 0 && T > 0, "mog_word_attention_fwd: non-positive dims");
  MOG_REQUIRE(T <= 32 && D <= 160, "mog_word_attention_fwd: T=%d (<=32) / D=%d (<=160) out of range", T, D);
  MOG_REQUIRE(B <= 65535, "mog_word_attention_fwd: batch too large");
  AttnArgs a{h, src, mask, nullptr, out, attn, nullptr, nullptr, B, Q, D, T, mask_quirk};
  return dispatch(a, false, nullptr, as_stream(stream));
}

extern "C" size_t mog_word_attention_bwd_workspace_bytes(int B, int Q, int D, int T) {
  if (B <= 0 || Q <= 0 || D <= 0 || T <= 0 || (D & 3)) return 0;
  return sizeof(float) * (size_t)B * ceil_div(Q, APIX) * T * D;
}

extern "C" int mog_word_attention_bwd(const float* h, const float* src, const uint8_t* mask, const float* dout,
                                      float* dh, float* dsrc, int B, int Q, int D, int T, int mask_quirk,
                                      void* workspace, size_t ws_bytes, void* stream) {
  MOG_REQUIRE(h && src && dout && dh && dsrc, "mog_word_attention_bwd: null tensor");
  MOG_REQUIRE(B > 0 && Q > 0 && D > 0 && T > 0, "mog_word_attention_bwd: non-positive dims");
  MOG_REQUIRE(T <= 32 && D <= 160, "mog_word_attention_bwd: T=%d (<=32) / D=%d (<=160) out of range", T, D);
  MOG_REQUIRE(B <= 65535, "mog_word_attention_bwd: batch too large");
  AttnArgs a{h, src, mask, dout, nullptr, nullptr, dh, dsrc, B, Q, D, T, mask_quirk};
  // with a workspace (D % 4 == 0) dsrc is WRITTEN by a deterministic two-stage reduction; without one the scalar kernel
  // accumulates into dsrc with atomics (the caller must have zeroed it)
  const size_t need = mog_word_attention_bwd_workspace_bytes(B, Q, D, T);
  float* partial = (workspace && need && ws_bytes >= need) ? static_cast<float*>(workspace) : nullptr;
  return dispatch(a, true, partial, as_stream(stream));
}
