// mt_fp32.cu — the fp32-accuracy path of the encoder: callers that run the reference in plain
// fp32 (X-Ray/train.py:52-60, README.md:125-149 outside autocast) get fp32-grade results (gate 1e-4 against the fp32
// oracle after all L layers) instead of bf16-operand math.
//
//   * Linear layers: "3xTF32" on the 5th-gen tensor cores.  Every fp32 operand is kept as a PAIR of fp32 arrays
//     big = x with the low 13 mantissa bits cleared (exactly representable in TF32, so the tensor core's own fp32 -> tf32
//     conversion cannot change it) and small = x - big (exact in fp32); x.w is accumulated in fp32 as
//     small_x.big_w + big_x.small_w + big_x.big_w (the small.small term is 2^-22 relative).  tcgen05.mma.kind::tf32,
//     operands staged by TMA (128B swizzle, 32 fp32 of K per row), accumulators in TMEM, 128 x 128 tiles, 3-stage ring.
//     Producers write the pair directly (LayerNorm, GELU epilogue, attention output), weights are split once.
//   * LayerNorm: two-pass fp32 statistics, fp32 output pair.
//   * Attention core: fp32 SIMT flash-style kernel (4 % of the FLOPs): one query row per thread, K / V tiles in shared memory.
// The residual stream is plain fp32.  Roughly 6-8x the time of the bf16 path (3 MMAs at half the bf16 rate, twice the
// operand bytes): an accuracy option, not the throughput path.
#define MT_WAIT_QUIET 1
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <stdlib.h>

namespace mt {

// ---------------------------------------------------------------------------------------------
// split helpers
// ---------------------------------------------------------------------------------------------
MT_DEVINL void split_tf32(float v, float& big, float& small) {
  big = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  small = v - big;
}

__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ big, float* __restrict__ small, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x * 4;
  for (size_t i = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i + 3 < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    float4 b, s;
    split_tf32(v.x, b.x, s.x); split_tf32(v.y, b.y, s.y); split_tf32(v.z, b.z, s.z); split_tf32(v.w, b.w, s.w);
    *reinterpret_cast<float4*>(big + i) = b;
    *reinterpret_cast<float4*>(small + i) = s;
  }
}

// LayerNorm, fp32 in, fp32 pair out; one warp per row, two-pass statistics (mt_layernorm.cu is the bf16-output twin)
__global__ void __launch_bounds__(256)
layernorm_f32_split_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* __restrict__ big, float* __restrict__ small, int M, int d, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const float* xr = x + size_t(row) * d;
  const int n4 = d >> 2;
  float s = 0.f;
  for (int c = lane; c < n4; c += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[c];
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / float(d);
  float q = 0.f;
  for (int c = lane; c < n4; c += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[c];
    const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, e = v.w - mean;
    q += (a * a + b * b) + (cc * cc + e * e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / float(d) + eps);
  for (int c = lane; c < n4; c += 32) {
    const float4 v = reinterpret_cast<const float4*>(xr)[c];
    const float4 g = reinterpret_cast<const float4*>(gamma)[c];
    const float4 bt = reinterpret_cast<const float4*>(beta)[c];
    float4 b, sm;
    split_tf32(fmaf((v.x - mean) * rstd, g.x, bt.x), b.x, sm.x);
    split_tf32(fmaf((v.y - mean) * rstd, g.y, bt.y), b.y, sm.y);
    split_tf32(fmaf((v.z - mean) * rstd, g.z, bt.z), b.z, sm.z);
    split_tf32(fmaf((v.w - mean) * rstd, g.w, bt.w), b.w, sm.w);
    reinterpret_cast<float4*>(big + size_t(row) * d)[c] = b;
    reinterpret_cast<float4*>(small + size_t(row) * d)[c] = sm;
  }
}

// ---------------------------------------------------------------------------------------------
// 3xTF32 GEMM:  y[M,N] = act(x[M,K] . w[N,K]^T + bias) (+ residual), x / w as (big, small) fp32 pairs
// ---------------------------------------------------------------------------------------------
constexpr int kFBM = 128, kFBN = 128, kFBK = 32;              // 32 fp32 = 128 B = one swizzle row
constexpr int kFTile = kFBM * kFBK * 4;                        // 16 KB: any 128 x 32 fp32 tile
constexpr int kFStageBytes = 4 * kFTile;                       // A big, A small, B big, B small
constexpr int kFStages = 3;
constexpr int kFEpiWarps = 4;
constexpr int kFThreads = 128 + 32 * kFEpiWarps;               // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int kFTmemCols = 2 * kFBN;
constexpr int kFSmemBar = kFStages * kFStageBytes;
constexpr int kFSmemTmemPtr = kFSmemBar + (2 * kFStages + 4) * 8;
constexpr int kFSmemBytes = kFSmemTmemPtr + 16 + 1024;

struct GemmF32Params {
  const float* bias;
  const float* resid;       // fp32 [M,N] or nullptr
  float* y;                 // fp32 [M,N]: the result (y_small == nullptr) or its big half
  float* y_small;           // fp32 [M,N] small half, or nullptr
  int M, N, K;
  int act;                  // 0 none, 1 exact-erf GELU
  int num_m_tiles, num_n_tiles;
};

MT_DEVINL void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor for kind::tf32: tf32 x tf32 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__global__ void __launch_bounds__(kFThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_ab, const __grid_constant__ CUtensorMap tm_as,
                   const __grid_constant__ CUtensorMap tm_bb, const __grid_constant__ CUtensorMap tm_bs,
                   const GemmF32Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_base = smem_base + kFSmemBar;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kFStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kFStages + 2 + s); };
  const uint32_t tmem_ptr_addr = smem_base + kFSmemTmemPtr;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_ab); tma_prefetch_desc(&tm_as); tma_prefetch_desc(&tm_bb); tma_prefetch_desc(&tm_bs);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kFStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kFEpiWarps); }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 2) tmem_alloc<1>(tmem_ptr_addr, kFTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kFSmemTmemPtr);
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + kFBK - 1) / kFBK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int row_a = (t / p.num_n_tiles) * kFBM, row_b = (t % p.num_n_tiles) * kFBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t s0 = smem_base + stage * kFStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), kFStageBytes);
          tma_load_2d(s0, &tm_ab, full_bar(stage), kb * kFBK, row_a, kEvictNormal);
          tma_load_2d(s0 + kFTile, &tm_as, full_bar(stage), kb * kFBK, row_a, kEvictNormal);
          tma_load_2d(s0 + 2 * kFTile, &tm_bb, full_bar(stage), kb * kFBK, row_b, kEvictLast);
          tma_load_2d(s0 + 3 * kFTile, &tm_bs, full_bar(stage), kb * kFBK, row_b, kEvictLast);
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kFBM, kFBN);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * kFBN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t s0 = smem_base + stage * kFStageBytes;
          const uint64_t ab = make_smem_desc_sw128(s0), asd = make_smem_desc_sw128(s0 + kFTile);
          const uint64_t bb = make_smem_desc_sw128(s0 + 2 * kFTile), bs = make_smem_desc_sw128(s0 + 3 * kFTile);
#pragma unroll
          for (int k = 0; k < kFBK / 8; ++k) {     // 8 fp32 = 32 B of K per instruction: +2 in the (addr >> 4) field
            const uint64_t o = uint64_t(2 * k);
            umma_tf32_ss(d_tmem, asd + o, bb + o, idesc, (kb > 0 || k > 0) ? 1u : 0u);   // small_x . big_w
            umma_tf32_ss(d_tmem, ab + o, bs + o, idesc, 1u);                             // big_x . small_w
            umma_tf32_ss(d_tmem, ab + o, bb + o, idesc, 1u);                             // big_x . big_w
          }
          umma_commit(empty_bar(stage));
          if (++stage == kFStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    int as = 0; uint32_t aphase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int row = (t / p.num_n_tiles) * kFBM + 32 * q + lane;
      const int col0 = (t % p.num_n_tiles) * kFBN;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(32 * q) << 16) + uint32_t(as * kFBN);
#pragma unroll 1
      for (int piece = 0; piece < kFBN / 16; ++piece) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(t_row + uint32_t(16 * piece), acc);
        tmem_wait_ld();
        if (piece == kFBN / 16 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        const int c_p = col0 + 16 * piece;
        if (row < p.M && c_p < p.N) {        // N % 16 == 0 (checked on the host)
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c_p + i));
            v[i] = __uint_as_float(acc[i]) + b4.x; v[i + 1] = __uint_as_float(acc[i + 1]) + b4.y;
            v[i + 2] = __uint_as_float(acc[i + 2]) + b4.z; v[i + 3] = __uint_as_float(acc[i + 3]) + b4.w;
          }
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
          }
          if (p.resid != nullptr) {
            const float4* rp = reinterpret_cast<const float4*>(p.resid + size_t(row) * p.N + c_p);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 r4 = __ldg(rp + i);
              v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
            }
          }
          float4* yp = reinterpret_cast<float4*>(p.y + size_t(row) * p.N + c_p);
          if (p.y_small == nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) yp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            float4* sp = reinterpret_cast<float4*>(p.y_small + size_t(row) * p.N + c_p);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 b, s;
              split_tf32(v[4 * i], b.x, s.x); split_tf32(v[4 * i + 1], b.y, s.y);
              split_tf32(v[4 * i + 2], b.z, s.z); split_tf32(v[4 * i + 3], b.w, s.w);
              yp[i] = b;
              sp[i] = s;
            }
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, kFTmemCols);
}

// ---------------------------------------------------------------------------------------------
// fp32 attention core (SIMT): one query row per thread, 64 queries per block, K / V tiles of 32 keys in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kFAQ = 64, kFAK = 32;

template <int DH>
__global__ void __launch_bounds__(kFAQ)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out_big, float* __restrict__ out_small, int B, int N,
                     int H, float scale) {
  __shared__ float ks[kFAK][DH];
  __shared__ float vs[kFAK][DH];
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int qi = qt * kFAQ + threadIdx.x;
  const size_t row_stride = size_t(3) * H * DH;
  float q[DH], o[DH];
  if (qi < N) {
    const float* qp = qkv + (size_t(b) * N + qi) * row_stride + size_t(h) * DH;
#pragma unroll
    for (int d4 = 0; d4 < DH; d4 += 4) {
      const float4 v = *reinterpret_cast<const float4*>(qp + d4);
      q[d4] = v.x * scale; q[d4 + 1] = v.y * scale; q[d4 + 2] = v.z * scale; q[d4 + 3] = v.w * scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < N; k0 += kFAK) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kFAK * (DH / 4); idx += kFAQ) {
      const int kk = idx / (DH / 4), d4 = (idx % (DH / 4)) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + kk < N) {
        const float* base = qkv + (size_t(b) * N + k0 + kk) * row_stride + size_t(h) * DH + d4;
        kv = *reinterpret_cast<const float4*>(base + size_t(H) * DH);
        vv = *reinterpret_cast<const float4*>(base + size_t(2) * H * DH);
      }
      *reinterpret_cast<float4*>(&ks[kk][d4]) = kv;
      *reinterpret_cast<float4*>(&vs[kk][d4]) = vv;
    }
    __syncthreads();
    const int kn = min(kFAK, N - k0);
    for (int kk = 0; kk < kn; ++kk) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[kk][d], s);
      const float m_new = fmaxf(m, s);
      const float alpha = __expf(m - m_new);       // exp(-inf) = 0 on the first key
      const float pr = __expf(s - m_new);
      l = fmaf(l, alpha, pr);
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha, pr * vs[kk][d]);
      m = m_new;
    }
  }
  if (qi < N) {
    const float inv = 1.0f / l;
    float* ob = out_big + (size_t(b) * N + qi) * H * DH + size_t(h) * DH;
    float* os = out_small + (size_t(b) * N + qi) * H * DH + size_t(h) * DH;
#pragma unroll
    for (int d4 = 0; d4 < DH; d4 += 4) {
      float4 bb, ss;
      split_tf32(o[d4] * inv, bb.x, ss.x); split_tf32(o[d4 + 1] * inv, bb.y, ss.y);
      split_tf32(o[d4 + 2] * inv, bb.z, ss.z); split_tf32(o[d4 + 3] * inv, bb.w, ss.w);
      *reinterpret_cast<float4*>(ob + d4) = bb;
      *reinterpret_cast<float4*>(os + d4) = ss;
    }
  }
}

}  // namespace mt

extern "C" int mt_split_tf32(const float* src, float* big, float* small, size_t n, void* stream) {
  using namespace mt;
  MT_REQUIRE(src && big && small, MT_ERR_BAD_SHAPE, "mt_split_tf32: null pointer");
  MT_REQUIRE(n % 4 == 0, MT_ERR_BAD_SHAPE, "mt_split_tf32: n must be a multiple of 4");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(big) | reinterpret_cast<uintptr_t>(small)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_split_tf32: pointers must be 16-byte aligned");
  if (n == 0) return MT_OK;
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<unsigned(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, big, small, n);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_layernorm_f32_split(const float* x, const float* gamma, const float* beta, float* big, float* small, int M,
                                      int d, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && gamma && beta && big && small, MT_ERR_BAD_SHAPE, "mt_layernorm_f32_split: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 4 == 0, MT_ERR_BAD_SHAPE, "mt_layernorm_f32_split: need M > 0 and d %% 4 == 0");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
               reinterpret_cast<uintptr_t>(big) | reinterpret_cast<uintptr_t>(small)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_layernorm_f32_split: pointers must be 16-byte aligned");
  layernorm_f32_split_kernel<<<unsigned(ceil_div(M, 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, gamma, beta, big,
                                                                                                          small, M, d, eps);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_gemm_f32x3(const float* x_big, const float* x_small, const float* w_big, const float* w_small,
                             const float* bias, const float* residual, float* y, float* y_small, int M, int N, int K, int act,
                             void* stream) {
  using namespace mt;
  MT_REQUIRE(x_big && x_small && w_big && w_small && y, MT_ERR_BAD_SHAPE, "mt_gemm_f32x3: null pointer");
  MT_REQUIRE(M > 0 && N > 0 && K > 0 && K % 4 == 0 && N % 16 == 0, MT_ERR_BAD_SHAPE,
             "mt_gemm_f32x3: need M, N, K > 0, K %% 4 == 0, N %% 16 == 0 (got %d,%d,%d)", M, N, K);
  MT_REQUIRE(act == 0 || act == 1, MT_ERR_UNSUPPORTED, "mt_gemm_f32x3: act must be 0 or 1 (gelu)");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x_big) | reinterpret_cast<uintptr_t>(x_small) | reinterpret_cast<uintptr_t>(w_big) |
               reinterpret_cast<uintptr_t>(w_small) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(residual) |
               reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(y_small)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_gemm_f32x3: pointers must be 16-byte aligned");
  CUtensorMap tm[4];
  const void* ptrs[4] = {x_big, x_small, w_big, w_small};
  for (int i = 0; i < 4; ++i) {
    uint64_t dims[2] = {uint64_t(K), uint64_t(i < 2 ? M : N)};
    uint64_t strides[1] = {uint64_t(K) * 4};
    uint32_t box[2] = {uint32_t(kFBK), uint32_t(i < 2 ? kFBM : kFBN)};
    int rc = encode_tmap(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptrs[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  GemmF32Params p;
  p.bias = bias; p.resid = residual; p.y = y; p.y_small = y_small;
  p.M = M; p.N = N; p.K = K; p.act = act;
  p.num_m_tiles = ceil_div(M, kFBM);
  p.num_n_tiles = ceil_div(N, kFBN);
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  int grid = p.num_m_tiles * p.num_n_tiles;
  if (grid > sms) grid = sms;
  {
    static bool attr_set[64] = {false};
    int dev = 0;
    MT_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      MT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  gemm_tf32x3_kernel<<<unsigned(grid), kFThreads, kFSmemBytes, reinterpret_cast<cudaStream_t>(stream)>>>(tm[0], tm[1], tm[2],
                                                                                                           tm[3], p);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_attention_f32(const float* qkv, float* out_big, float* out_small, int B, int N, int H, int dh, float scale,
                                void* stream) {
  using namespace mt;
  MT_REQUIRE(qkv && out_big && out_small, MT_ERR_BAD_SHAPE, "mt_attention_f32: null pointer");
  MT_REQUIRE(B > 0 && N > 0 && H > 0 && B < 65536 && H < 65536, MT_ERR_BAD_SHAPE, "mt_attention_f32: bad sizes");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out_big) | reinterpret_cast<uintptr_t>(out_small)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_attention_f32: pointers must be 16-byte aligned");
  const dim3 grid(unsigned(ceil_div(N, kFAQ)), unsigned(H), unsigned(B));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (dh) {
    case 64: attention_f32_kernel<64><<<grid, kFAQ, 0, s>>>(qkv, out_big, out_small, B, N, H, scale); break;
    case 32: attention_f32_kernel<32><<<grid, kFAQ, 0, s>>>(qkv, out_big, out_small, B, N, H, scale); break;
    case 24: attention_f32_kernel<24><<<grid, kFAQ, 0, s>>>(qkv, out_big, out_small, B, N, H, scale); break;
    case 16: attention_f32_kernel<16><<<grid, kFAQ, 0, s>>>(qkv, out_big, out_small, B, N, H, scale); break;
    default:
      set_error("mt_attention_f32: head_dim %d not instantiated (64, 32, 24, 16)", dh);
      return MT_ERR_UNSUPPORTED;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}
