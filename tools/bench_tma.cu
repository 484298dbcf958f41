// bench_tma.cu — how many bytes per clock one SM can pull in through TMA, with and without cluster multicast.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I metatransformer_b200/csrc tools/bench_tma.cu -o tools/bench_tma.bin -lcuda
// Clusters of 2 CTAs (one per SM), a 5-deep ring of 32 KB stages per CTA like the GEMM's operand ring, one thread per CTA
// producing and consuming.  mode 0: every CTA loads its own two 16 KB boxes (128 rows x 64 bf16, 128B swizzle) per stage;
// mode 1: every CTA issues ONE box per stage, multicast to both CTAs — each SM still receives 32 KB per stage but its TMA
// unit issues half the requests.  Source region 25 MB (L2 resident after the first pass).
#include "mt_ptx.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
using namespace mt;

constexpr int kStages = 5;
constexpr int kBox = 16384;

MT_DEVINL void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
tma_rate_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_packed, int iters, int mode, int rows_total, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar0 = smem_base + kStages * 2 * kBox;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), mode == 1 ? 2 : 1);   // multicast: both CTAs' consumers release a stage to both producers
    }
    fence_mbar_init();
  }
  cluster_sync_all();
  if (threadIdx.x == 0) {
    const int pair = blockIdx.x >> 1;
    const int npairs = gridDim.x >> 1;
    const int rows_per_pair = rows_total / npairs;
    const int row_base = pair * rows_per_pair;
    long long t0 = clock64();
    int consumed = 0;
    for (int it = 0; it < iters + kStages - 1; ++it) {
      if (it < iters) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(full(s), 2 * kBox);
        const int kb = it % 12, rb = (it / 12) % (rows_per_pair / 256);
        const uint32_t dst = smem_base + s * 2 * kBox;
        if (mode == 2) {
          // packed source: the same bytes viewed as [rows * 12][64] (row pitch 128 B), a 128-row box = 16 KB contiguous
          const int r0 = ((row_base + rb * 256 + int(rank) * 128) * 12 + kb * 128) % (rows_total * 12 - 256);
          tma_load_2d(dst, &tm_packed, full(s), 0, r0, kEvictNormal);
          tma_load_2d(dst + kBox, &tm_packed, full(s), 0, r0 + 128, kEvictNormal);
        } else if (mode == 0) {
          tma_load_2d(dst, &tm, full(s), kb * 64, row_base + rb * 256 + int(rank) * 128, kEvictNormal);
          tma_load_2d(dst + kBox, &tm, full(s), kb * 64, row_base + (rb * 256 + 128 + int(rank) * 128) % rows_per_pair, kEvictNormal);
        } else {
          // box `rank` of the stage, delivered to both CTAs (same smem offset, same barrier offset in each)
          tma_load_2d_mc(dst + rank * kBox, &tm, full(s), kb * 64, row_base + rb * 256 + int(rank) * 128, uint16_t(3));
        }
      }
      if (it >= kStages - 1) {
        const int s = consumed % kStages;
        const uint32_t ph = (consumed / kStages) & 1;
        mbar_wait(full(s), ph);
        mbar_arrive(empty(s));
        if (mode == 1) mbar_arrive_remote(empty(s), rank ^ 1u);
        ++consumed;
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  cluster_sync_all();
}

int main() {
  const int rows = 16384, cols = 768;
  void* d;
  cudaMalloc(&d, size_t(rows) * cols * 2);
  cudaMemset(d, 0, size_t(rows) * cols * 2);
  CUtensorMap tm;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); return 1; }
  }
  CUtensorMap tm_packed;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    cuuint64_t dims[2] = {64, cuuint64_t(rows) * 12};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm_packed, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 2 failed %d\n", int(r)); return 1; }
  }
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 256);
  const int smem = kStages * 2 * kBox + 256 + 1024;
  cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  for (int grid : {2, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        tma_rate_kernel<<<grid, 128, smem>>>(tm, tm_packed, iters, mode, rows, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[256];
      cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      printf("grid=%3d %s: %7.1f cycles per 32 KB stage -> %5.1f B/clk received per SM\n", grid,
             mode == 1 ? "multicast (1 box issued / SM)" : (mode == 2 ? "unicast, 16 KB-contiguous boxes " : "unicast   (2 boxes issued / SM)"), double(mx) / iters,
             32768.0 * iters / double(mx));
    }
  }
  return 0;
}
