// bench_umma.cu — micro-benchmark of the raw tcgen05.mma issue/execute rate on one SM (or one SM pair).
// No TMA, operands are whatever sits in shared memory (zeros): measures cycles per MMA for
// cta_group in {1,2} x N in {64,128,256} x (all k-steps from one 128B-swizzled stage).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I metatransformer_b200/csrc tools/bench_umma.cu -o /tmp/bench_umma
#include "mt_ptx.cuh"
#include <cstdio>
#include <cstdlib>

using namespace mt;

template <int CG>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int n_dim, int iters, int commit_every, long long* out, int uniform) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  // layout: 4 stages x (A 16KB + B 32KB) then barrier + tmem ptr
  const uint32_t bar = smem_base + 4 * 49152;
  const uint32_t tmem_ptr_addr = bar + 32;
  for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_gen)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 0) tmem_alloc<CG>(tmem_ptr_addr, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 4 * 49152 + 32);
  if (uniform && warp == 1 && rank == 0) {
    // warp-uniform issue loop: every lane runs the loop, one elected lane issues (no divergent single-lane region)
    const uint32_t idesc = make_idesc_bf16(128 * CG, n_dim);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = smem_base + (it & 3) * 49152;
      const uint64_t adesc = make_smem_desc_sw128(sa);
      const uint64_t bdesc = make_smem_desc_sw128(sa + 16384);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss<CG>(tmem_base + uint32_t((it & 1) * 256), adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc,
                          (k > 0) ? 1u : 0u);
        if (it + 1 == iters || (commit_every != 1 && (it + 1) % commit_every == 0)) {
          if constexpr (CG == 2) umma_commit_cg2_mc(bar, 0x1); else umma_commit(bar);
        } else if (commit_every == 1) {
          if constexpr (CG == 2) umma_commit_cg2_mc(bar + 8, 0x3); else umma_commit(bar + 8);
        }
      }
      __syncwarp();
      if (it + 1 == iters || (commit_every != 1 && (it + 1) % commit_every == 0)) {
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * CG, n_dim);
    uint32_t phase = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = smem_base + (it & 3) * 49152;
      const uint64_t adesc = make_smem_desc_sw128(sa);
      const uint64_t bdesc = make_smem_desc_sw128(sa + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss<CG>(tmem_base + uint32_t((it & 1) * 256), adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc,
                        (k > 0) ? 1u : 0u);
      if (it + 1 == iters) {
        if constexpr (CG == 2) umma_commit_cg2_mc(bar, 0x1); else umma_commit(bar);
        mbar_wait(bar, phase);
      } else if (commit_every == 1) {
        // commit to a barrier nobody waits on (like the per-stage "smem slot free" commits of the GEMM main loop)
        if constexpr (CG == 2) umma_commit_cg2_mc(bar + 8, 0x3); else umma_commit(bar + 8);
      } else if ((it + 1) % commit_every == 0) {
        if constexpr (CG == 2) umma_commit_cg2_mc(bar, 0x1); else umma_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG>
static void run(int n_dim, int iters, int commit_every, int grid, int uniform = 0) {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 256);
  cudaMemset(d_out, 0, sizeof(long long) * 256);
  const int smem = 4 * 49152 + 64 + 1024;
  cudaFuncSetAttribute(umma_rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, umma_rate_kernel<CG>, n_dim, iters, commit_every, d_out, uniform);
    if (e != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sync error %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[256];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  const double per_mma = double(mx) / (4.0 * iters);
  const double macs = 128.0 * n_dim * 16;  // per CTA per MMA
  printf("cg=%d N=%3d grid=%3d commit_every=%4d %s: %8.1f cycles/MMA  -> %6.1f MAC/cycle/SM (peak 4096)\n", CG, n_dim, grid,
         commit_every, uniform ? "uniform" : "lane0  ", per_mma, macs / per_mma);
  cudaFree(d_out);
}

// dependent-accumulation study: MMA i accumulates into accumulator (i % streams); consecutive MMAs on the same
// accumulator are a dependent chain, different accumulators are independent work for the tensor pipe
__global__ void __launch_bounds__(128, 1) umma_streams_kernel(int n_dim, int iters, int streams, long long* out, int ts) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = smem_base + 4 * 49152;
  const uint32_t tmem_ptr_addr = bar + 32;
  for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_gen)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 0) tmem_alloc<1>(tmem_ptr_addr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + 4 * 49152 + 32);
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n_dim);
    const uint64_t adesc = make_smem_desc_sw128(smem_base);
    const uint64_t bdesc = make_smem_desc_sw128(smem_base + 16384);
    long long t0 = clock64();
    int st = 0;
    for (int it = 0; it < iters; ++it) {
      if (ts) umma_f16_ts(tmem_base + uint32_t(st * n_dim), tmem_base + 480u + uint32_t(8 * (it & 3)), bdesc + uint64_t(2 * (it & 3)), idesc, 1u);
      else umma_f16_ss<1>(tmem_base + uint32_t(st * n_dim), adesc + uint64_t(2 * (it & 3)), bdesc + uint64_t(2 * (it & 3)), idesc, 1u);
      if (++st == streams) st = 0;
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_base, 512);
}

static void run_streams(int n_dim, int iters, int streams, int ts = 0) {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 256);
  const int smem = 4 * 49152 + 64 + 1024;
  cudaFuncSetAttribute(umma_streams_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    umma_streams_kernel<<<2, 128, smem>>>(n_dim, iters, streams, d_out, ts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("streams error %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d accumulators=%d A from %s: %8.1f cycles/MMA (ideal %d)\n", n_dim, streams, ts ? "TMEM" : "smem", double(h[0]) / iters, n_dim / 2);
  cudaFree(d_out);
}

int main() {
  for (int n : {16, 32, 64, 128})
    for (int st : {1, 2})
      if (st * n <= 448) { run_streams(n, 4000, st, 0); run_streams(n, 4000, st, 1); }
  const int iters = 2000;
  for (int n : {32, 64, 128, 256}) {
    for (int ce : {2000, 4, 1}) {
      run<1>(n, iters, ce, 2, 0);
      run<1>(n, iters, ce, 2, 1);
    }
  }
  run<2>(256, iters, 4, 2, 0);
  run<2>(256, iters, 4, 2, 1);
  for (int grid : {148}) {
    for (int n : {64, 128, 256}) {
      run<1>(n, iters, iters, grid);
      run<2>(n, iters, iters, grid);
    }
    run<1>(256, iters, 1, grid);
    run<2>(256, iters, 1, grid);
    run<1>(256, iters, 12, grid);
    run<2>(256, iters, 12, grid);
  }
  return 0;
}
