// tma_rate.cu -- per-SM ingest rate of TMA loads by box shape (all 148 SMs active, source mostly L2-resident).
// Each CTA keeps RING loads in flight (mbarrier per slot) and reports clocks per load and bytes per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../multiple-objects-gan_b200/csrc tma_rate.cu -o tma_rate -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace mog::tc;

constexpr int RING = 4;   // power of two: the single issuing thread must not spend its time on integer division

__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
               "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}

// mode 0: 2-D box (inner, rows)   mode 1: 1-D bulk copy of `bytes`   mode 2: 4-D box {inner, 8, hh, 1} of an NHWC plane
__global__ void __launch_bounds__(32, 1) rate_kernel(const __grid_constant__ CUtensorMap tm, const unsigned char* src, int mode, uint32_t bytes,
                                                     int reps, int span_rows, int hh, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[RING];
  const uint32_t slot = (bytes + 1023) / 1024 * 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  uint32_t j = blockIdx.x * 131u;
  for (int it = 0; it < reps + RING; ++it) {
    const int s = it & (RING - 1);
    if (it >= RING) mbar_wait(&bar[s], (uint32_t)(((it >> 2) - 1) & 1));
    if (it < reps) {
      expect_tx(&bar[s], bytes);
      const uint32_t dst = smem_u32(smem) + s * slot;
      j = (j + 7u) & 1023u;
      if (mode == 0) tma2d(dst, &tm, &bar[s], 0, (int)((j * 16u) & 4095u));
      else if (mode == 1) bulk1d(dst, src + (size_t)j * 32768, bytes, &bar[s]);
      else tma4d(dst, &tm, &bar[s], 0, (int)(j & 7u) * 8, (int)((j >> 3) & 3u) * 16, (int)((j >> 5) & 15u));
    }
  }
  out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncFn enc = (EncFn)fp;
  unsigned char* buf;
  const size_t BYTES = 64ull << 20;
  cudaMalloc(&buf, BYTES);
  cudaMemset(buf, 1, BYTES);
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 2000;
  struct Case { const char* name; int mode; int inner_bytes; int rows; int hh; CUtensorMapSwizzle sw; };
  Case cases[] = {
      {"2-D box  64 B x 256 rows (SW64), row pitch 6 KB ", 0, 64, 256, 0, CU_TENSOR_MAP_SWIZZLE_64B},
      {"2-D box 128 B x 128 rows (SW128), row pitch 6 KB", 0, 128, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B},
      {"2-D box 128 B x 256 rows (SW128), row pitch 6 KB", 0, 128, 256, 0, CU_TENSOR_MAP_SWIZZLE_128B},
      {"2-D box  32 B x 256 rows (SW32), row pitch 6 KB ", 0, 32, 256, 0, CU_TENSOR_MAP_SWIZZLE_32B},
      {"1-D bulk copy 16 KB contiguous                  ", 1, 16384, 1, 0, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"1-D bulk copy  8 KB contiguous                  ", 1, 8192, 1, 0, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"1-D bulk copy 24 KB contiguous                  ", 1, 24576, 1, 0, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"4-D box {64 B, 8 w, 18 h} NHWC 192 B/pixel (SW64)", 2, 64, 144, 18, CU_TENSOR_MAP_SWIZZLE_64B},
      {"4-D box {128 B, 8 w, 18 h} NHWC 192 B/pixel (SW128)", 2, 128, 144, 18, CU_TENSOR_MAP_SWIZZLE_128B},
      {"4-D box {128 B, 8 w, 18 h} NHWC 128 B/pixel (SW128)", 3, 128, 144, 18, CU_TENSOR_MAP_SWIZZLE_128B},
  };
  printf("%-52s %10s %10s %10s\n", "case", "clk/load", "B/clk/SM", "B/clk 1SM");
  for (auto& c : cases) {
    CUtensorMap tm;
    uint32_t bytes;
    int mode = c.mode == 3 ? 2 : c.mode;
    if (c.mode == 0) {
      cuuint64_t dims[2] = {3072, 8192};           // bf16 [8192 rows][3072], 6 KB pitch, 48 MB
      cuuint64_t strides[1] = {6144};
      cuuint32_t box[2] = {(cuuint32_t)c.inner_bytes / 2, (cuuint32_t)c.rows};
      cuuint32_t es[2] = {1, 1};
      enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      bytes = c.inner_bytes * c.rows;
    } else if (c.mode >= 2) {
      const int C = c.mode == 3 ? 64 : 96;         // bf16 NHWC plane [32][128][128][C]
      cuuint64_t dims[4] = {(cuuint64_t)C, 128, 128, 16};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)128 * C * 2, (cuuint64_t)128 * 128 * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)c.inner_bytes / 2, 8, (cuuint32_t)c.hh, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) printf("encode failed %d\n", (int)r);
      bytes = c.inner_bytes * 8 * c.hh;
    } else {
      bytes = c.inner_bytes;
      tm = CUtensorMap{};
    }
    double res[2];
    for (int gsel = 0; gsel < 2; ++gsel) {
      const int grid = gsel ? 148 : 1;
      rate_kernel<<<grid, 32, 198 * 1024>>>(tm, buf, mode, bytes, reps, 8192 - 256, c.hh, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < grid; ++i) for (int j = i + 1; j < grid; ++j) if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
      res[gsel] = (double)h[grid / 2] / reps;
    }
    printf("%-52s %10.1f %10.1f %10.1f\n", c.name, res[1], bytes / res[1], bytes / res[0]);
  }
  return 0;
}
