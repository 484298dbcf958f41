// bench_tmem.cu — micro-benchmark of tcgen05.ld / tcgen05.st throughput (TMEM <-> registers) on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I metatransformer_b200/csrc tools/bench_tmem.cu -o tools/bench_tmem.bin
#include "mt_ptx.cuh"
#include <cstdio>
using namespace mt;

__global__ void __launch_bounds__(256, 1) tmem_rate_kernel(int iters, int mode, long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmem_ptr), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_ptr + (uint32_t(32 * (warp & 3)) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {          // loads only
    for (int it = 0; it < iters; ++it) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(base + uint32_t((it & 7) * 32), r);
      tmem_wait_ld();
      acc += __uint_as_float(r[it & 31]);
    }
  } else if (mode == 1) {   // 4 loads in flight before the wait (128 columns, like a softmax row)
    for (int it = 0; it < iters; it += 4) {
      uint32_t r[128];
      tmem_ld_32x32b_x32_at<0>(base, r);
      tmem_ld_32x32b_x32_at<32>(base + 32u, r);
      tmem_ld_32x32b_x32_at<64>(base + 64u, r);
      tmem_ld_32x32b_x32_at<96>(base + 96u, r);
      tmem_wait_ld();
      acc += __uint_as_float(r[it & 127]);
    }
  } else {                  // load + store round trip (O rescale pattern)
    for (int it = 0; it < iters; ++it) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(base + uint32_t(256 + (it & 3) * 32), r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * 0.5f);
      tmem_st_32x32b_x32(base + uint32_t(256 + (it & 3) * 32), r);
      tmem_wait_st();
    }
  }
  long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  if (acc == 12345.f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_ptr, 512);
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, sizeof(long long) * 8 * 148);
  cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int threads : {32, 128, 256}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        tmem_rate_kernel<<<148, threads>>>(iters, mode, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long h[8 * 148];
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < threads / 32; ++i) if (h[i] > mx) mx = h[i];
      const double cyc = double(mx) / iters;
      const double bytes_per_cyc_sm = (threads / 32) * 4096.0 / cyc * (mode == 2 ? 2 : 1);
      printf("warps=%d mode=%s: %.1f cycles per x32 op per warp -> %.0f B/cycle/SM\n", threads / 32,
             mode == 0 ? "ld+wait" : (mode == 1 ? "4xld+wait" : "ld+st"), cyc, bytes_per_cyc_sm);
    }
  }
  return 0;
}
