// umma_rate.cu -- cycles per tcgen05.mma (kind::f16, bf16 operands, M=128) by operand major-ness and N.
// One CTA per SM issues REPS back-to-back MMAs on (uninitialised) shared memory and times them with clock64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../multiple-objects-gan_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace mog::tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int a_mn, int b_mn, int reps, int nacc, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operands (NaN garbage could change data-dependent power, not timing; zero anyway)
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 48 * 1024;
    long long t0 = 0, t1 = 0;
    if (lane == 0) {
      uint64_t da[4], db[4];
      for (uint32_t k = 0; k < 4; ++k) {
        da[k] = a_mn ? make_desc_sw128_mn(sa + k * 2048, 8192, 1024) : make_desc_sw128(sa + k * 32);
        db[k] = b_mn ? make_desc_sw128_mn(sb + k * 2048, 8192, 1024) : make_desc_sw128(sb + k * 32);
      }
      umma_bf16(tmem, da[0], db[0], idesc, false);
      umma_bf16(tmem + N, da[0], db[0], idesc, false);
      t0 = clock64();
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) umma_bf16(tmem + (nacc > 1 ? (u & 1) * N : 0), da[u & 3], db[u & 3], idesc, true);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane == 0) {
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 4096;
  const int Ns[] = {16, 32, 48, 64, 96, 128, 160, 192, 256};
  printf("cycles per MMA (M=128, K=16, bf16), median over 148 CTAs; ideal = N/2\n");
  printf("%6s %10s %10s %10s %10s   (A major, B major)\n", "N", "K,K", "MN,K", "K,MN", "MN,MN");
  for (int N : Ns) {
    printf("%6d", N);
    for (int combo = 0; combo < 4; ++combo) {
      const int a_mn = combo & 1, b_mn = combo >> 1;
      int nacc = 512 / N; if (nacc > 4) nacc = 4;
      rate_kernel<<<148, 128, 97 * 1024>>>(N, a_mn, b_mn, reps, nacc, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" err:%s", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 148; ++i) for (int j = i + 1; j < 148; ++j) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
      printf(" %10.1f", (double)h[74] / reps);
    }
    printf("\n");
  }
  return 0;
}
