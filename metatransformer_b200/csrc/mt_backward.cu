// mt_backward.cu — element-wise / row-wise backward pieces of the Block (the tokenizers are
// trainable under a frozen encoder, Audio/src/models/ast_models.py:66-70, so dX must flow through every Block).
//   * mt_layernorm_bwd: dx = add + rstd * (g - mean_d(g) - xhat * mean_d(g * xhat)),  g = dy * gamma, xhat = (x-mean)*rstd
//   * mt_gelu:          h = gelu(u) (training-mode forward keeps the pre-activation u for the backward pass)
// Both are HBM-bound streaming kernels: one warp per row / 16-byte vectors.
#include "mt_ptx.cuh"
#include "mt_common.cuh"

namespace mt {

constexpr int kBwWarps = 8;

MT_DEVINL void ld8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
MT_DEVINL float warp_sum_bw(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV > 0: d == NV * 256 exactly and the row's dy / x chunks stay in registers between the two passes (one HBM read);
// NV == 0: generic d (d % 8 == 0), the row is re-read for the second pass (L1/L2 hits; rows are <= 16 KB).
template <int NV>
__global__ void __launch_bounds__(kBwWarps * 32)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const float2* __restrict__ stats, const float* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx, int M, int d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kBwWarps + warp;
  if (row >= M) return;
  const size_t base = size_t(row) * d;
  const float2 st = stats[row];
  const float mean = st.x, rstd = st.y;
  const float inv_d = 1.0f / float(d);
  constexpr int NI = NV > 0 ? NV : 1;
  float gr[NI][8], xh[NI][8];
  float s1 = 0.f, s2 = 0.f;
  const int nchunks = d >> 3;
  if constexpr (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      ld8_bf16(dy + base + c * 8, gr[i]);
      ld8_bf16(x + base + c * 8, xh[i]);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (gamma != nullptr) gr[i][j] *= gamma[c * 8 + j];
        xh[i][j] = (xh[i][j] - mean) * rstd;
        s1 += gr[i][j];
        s2 = fmaf(gr[i][j], xh[i][j], s2);
      }
    }
  } else {
    for (int c = lane; c < nchunks; c += 32) {
      float g[8], xv[8];
      ld8_bf16(dy + base + c * 8, g);
      ld8_bf16(x + base + c * 8, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (gamma != nullptr) g[j] *= gamma[c * 8 + j];
        s1 += g[j];
        s2 = fmaf(g[j], (xv[j] - mean) * rstd, s2);
      }
    }
  }
  const float m1 = warp_sum_bw(s1) * inv_d, m2 = warp_sum_bw(s2) * inv_d;
  auto emit = [&](int c, const float (&g)[8], const float (&xhat)[8]) {
    float a[8];
    if (add != nullptr) ld8_bf16(add + base + c * 8, a);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] - m1 - xhat[j] * m2) + (add != nullptr ? a[j] : 0.f);
    uint4 w;
    w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
    w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(dx + base + c * 8) = w;
  };
  if constexpr (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) emit(lane + 32 * i, gr[i], xh[i]);
  } else {
    for (int c = lane; c < nchunks; c += 32) {
      float g[8], xv[8];
      ld8_bf16(dy + base + c * 8, g);
      ld8_bf16(x + base + c * 8, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (gamma != nullptr) g[j] *= gamma[c * 8 + j];
        xv[j] = (xv[j] - mean) * rstd;
      }
      emit(c, g, xv);
    }
  }
}

MT_DEVINL float gelu_erf_ref(float x) {
  const float t = fminf(fabsf(x), 12.0f);
  float p = -0.00048811757005751133f;
  p = fmaf(p, t, 0.007198806386440992f);
  p = fmaf(p, t, -0.052146803587675095f);
  p = fmaf(p, t, -0.4595957100391388f);
  p = fmaf(p, t, -1.1510006189346313f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));
  return fmaf(t * e, -0.5f, fmaxf(x, 0.0f));
}

__global__ void __launch_bounds__(256) gelu_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ h,
                                                    size_t n8) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    ld8_bf16(u + i * 8, v);
    uint4 w;
    w.x = pack_bf16x2(gelu_erf_ref(v[0]), gelu_erf_ref(v[1])); w.y = pack_bf16x2(gelu_erf_ref(v[2]), gelu_erf_ref(v[3]));
    w.z = pack_bf16x2(gelu_erf_ref(v[4]), gelu_erf_ref(v[5])); w.w = pack_bf16x2(gelu_erf_ref(v[6]), gelu_erf_ref(v[7]));
    *reinterpret_cast<uint4*>(h + i * 8) = w;
  }
}

// dst[c, r] = src[r, c]; dst has leading dimension ldd >= R (columns R..ldd-1 are zero-filled) so that the result can
// feed the TN GEMM (row pitch must be a multiple of 16 bytes).  64x64 tiles: 16-byte global loads and stores on both
// sides, the transposition itself happens through shared memory (row pitch 66 elements = 33 words -> conflict-free
// column reads).  Requires C % 8 == 0 and ldd % 8 == 0 for the vector path; a scalar path covers the rest.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src,
                                                              __nv_bfloat16* __restrict__ dst, int R, int C, int ldd,
                                                              int vec) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int t = threadIdx.x;
  if (vec) {
    // load: 64 rows x 8 vectors of 8 elements; thread -> (row = t / 8 + 32 * k, vector = t % 8)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rr = (t >> 3) + 32 * k, cv = (t & 7) * 8;
      const int r = r0 + rr, c = c0 + cv;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < R && c < C) v = *reinterpret_cast<const uint4*>(src + size_t(r) * C + c);
      const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[rr][cv + j] = e[j];
    }
    __syncthreads();
    // store: output row = source column; thread -> (out row = t / 8 + 32 * k, vector of 8 source rows = t % 8)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int cc = (t >> 3) + 32 * k, rv = (t & 7) * 8;
      const int c = c0 + cc, r = r0 + rv;
      if (c < C && r < ldd) {
        __nv_bfloat16 e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = tile[rv + j][cc];   // rows >= R were loaded as zeros
        *reinterpret_cast<uint4*>(dst + size_t(c) * ldd + r) = *reinterpret_cast<const uint4*>(e);
      }
    }
  } else {
    for (int i = t; i < 64 * 64; i += 256) {
      const int rr = i >> 6, cc = i & 63;
      const int r = r0 + rr, c = c0 + cc;
      tile[rr][cc] = (r < R && c < C) ? src[size_t(r) * C + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = t; i < 64 * 64; i += 256) {
      const int cc = i >> 6, rr = i & 63;
      const int c = c0 + cc, r = r0 + rr;
      if (c < C && r < ldd) dst[size_t(c) * ldd + r] = tile[rr][cc];
    }
  }
}

// out[n] += sum_m x[m, n]   (out must be zero-initialised; fp32 atomics across row chunks)
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int M,
                                                      int N) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= N) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < M; r += gridDim.y) s += __bfloat162float(x[size_t(r) * N + col]);
  atomicAdd(out + col, s);
}

// Vectorised form for N % 8 == 0: a block covers 256 columns (32 lanes x 8 bf16 = 512 contiguous bytes per row) and a
// contiguous range of rows; its 8 warps take every 8th row, partial sums meet in shared memory, one fp32 atomic per
// (block, column).  16-byte loads, fully coalesced: HBM-bound instead of the 2-byte-per-thread kernel's ~1 TB/s.
__global__ void __launch_bounds__(256) colsum_vec8_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                           int M, int N, int rows_per_block) {
  __shared__ float part[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {   // four independent 16-byte loads in flight per thread
      float v0[8], v1[8], v2[8], v3[8];
      ld8_bf16(x + size_t(r) * N + col, v0);
      ld8_bf16(x + size_t(r + 8) * N + col, v1);
      ld8_bf16(x + size_t(r + 16) * N + col, v2);
      ld8_bf16(x + size_t(r + 24) * N + col, v3);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
    }
    for (; r < r1; r += 8) {
      float v[8];
      ld8_bf16(x + size_t(r) * N + col, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// Un-folding of the weight gradient of a LayerNorm-folded Linear (forward: y = xhat W'^T + b', W' = W*gamma, b' = b + W beta):
//   dW[n,k]   = dWf[n,k] * gamma[k] + dyb[n] * beta[k]        (dWf = dY^T xhat, dyb = column sums of dY = db)
//   dgamma[k] = sum_n dWf[n,k] * W[n,k]
//   dbeta[k]  = sum_n dyb[n]   * W[n,k]
// Threads run along k (coalesced), blocks split n; dgamma / dbeta are accumulated with fp32 atomics (zero-initialised).
template <typename TW>
__global__ void __launch_bounds__(256)
ln_fold_bwd_kernel(const float* dwf, const TW* __restrict__ w, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ dyb, float* dw,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int K, int rows_per_block) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  const int n0 = blockIdx.y * rows_per_block;
  const int n1 = min(N, n0 + rows_per_block);
  const float g = gamma[k], bt = beta[k];
  float sg = 0.f, sb = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float d = dwf[size_t(n) * K + k];
    float wv;
    if constexpr (sizeof(TW) == 4) wv = w[size_t(n) * K + k];
    else wv = __bfloat162float(w[size_t(n) * K + k]);
    const float yb = dyb[n];
    dw[size_t(n) * K + k] = fmaf(d, g, yb * bt);
    sg = fmaf(d, wv, sg);
    sb = fmaf(yb, wv, sb);
  }
  atomicAdd(dgamma + k, sg);
  atomicAdd(dbeta + k, sb);
}

}  // namespace mt

extern "C" int mt_ln_fold_bwd(const void* dw_folded, const void* w, int w_dtype, const float* gamma, const float* beta,
                              const float* dyb, float* dw, float* dgamma, float* dbeta, int N, int K, void* stream) {
  using namespace mt;
  MT_REQUIRE(dw_folded && w && gamma && beta && dyb && dw && dgamma && dbeta, MT_ERR_BAD_SHAPE, "mt_ln_fold_bwd: null pointer");
  MT_REQUIRE(N > 0 && K > 0, MT_ERR_BAD_SHAPE, "mt_ln_fold_bwd: N and K must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MT_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, size_t(K) * 4, s));
  MT_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, size_t(K) * 4, s));
  const int rpb = 32;
  dim3 grid(unsigned(ceil_div(K, 256)), unsigned(ceil_div(N, rpb)));
  if (w_dtype == MT_DT_F32)
    ln_fold_bwd_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(dw_folded),
                                                   reinterpret_cast<const float*>(w), gamma, beta, dyb, dw, dgamma, dbeta,
                                                   N, K, rpb);
  else if (w_dtype == MT_DT_BF16)
    ln_fold_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(dw_folded),
                                                           reinterpret_cast<const __nv_bfloat16*>(w), gamma, beta, dyb, dw,
                                                           dgamma, dbeta, N, K, rpb);
  else {
    set_error("mt_ln_fold_bwd: w_dtype must be fp32 or bf16 (got %d)", w_dtype);
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_transpose_bf16(const void* src, void* dst, int R, int C, int ldd, void* stream) {
  using namespace mt;
  MT_REQUIRE(src && dst, MT_ERR_BAD_SHAPE, "mt_transpose_bf16: null pointer");
  MT_REQUIRE(R > 0 && C > 0 && ldd >= R, MT_ERR_BAD_SHAPE, "mt_transpose_bf16: need R, C > 0 and ldd >= R");
  dim3 grid(unsigned(ceil_div(C, 64)), unsigned(ceil_div(ldd, 64)));
  const int vec = (C % 8 == 0 && ldd % 8 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) ? 1 : 0;
  transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), reinterpret_cast<__nv_bfloat16*>(dst), R, C, ldd, vec);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_colsum(const void* x, float* out, int M, int N, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && out, MT_ERR_BAD_SHAPE, "mt_colsum: null pointer");
  MT_REQUIRE(M > 0 && N > 0, MT_ERR_BAD_SHAPE, "mt_colsum: M and N must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  MT_CHECK_CUDA(cudaMemsetAsync(out, 0, size_t(N) * 4, s));
  if (N % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    const int col_blocks = ceil_div(N, 256);
    int row_blocks = ceil_div(4 * sms, col_blocks);                 // ~4 blocks per SM in total
    int rows_per_block = ceil_div(M, row_blocks);
    rows_per_block = ceil_div(rows_per_block < 64 ? 64 : rows_per_block, 8) * 8;
    row_blocks = ceil_div(M, rows_per_block);
    const dim3 grid(static_cast<unsigned>(col_blocks), static_cast<unsigned>(row_blocks), 1u);
    colsum_vec8_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, M, N, rows_per_block);
  } else {
    int chunks = M < 512 ? 1 : (M / 256 > 592 ? 592 : M / 256);
    dim3 grid(unsigned(ceil_div(N, 256)), unsigned(chunks));
    colsum_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, M, N);
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_layernorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const void* add,
                                void* dx, int M, int d, void* stream) {
  using namespace mt;
  MT_REQUIRE(dy && x && stats && dx, MT_ERR_BAD_SHAPE, "mt_layernorm_bwd: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_layernorm_bwd: need M > 0 and d %% 8 == 0");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
               reinterpret_cast<uintptr_t>(add)) & 15) == 0, MT_ERR_BAD_ALIGN,
             "mt_layernorm_bwd: tensors must be 16-byte aligned");
  const unsigned grid = unsigned(ceil_div(M, kBwWarps));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const float2* sp = reinterpret_cast<const float2*>(stats);
  const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(add);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(dx);
  if (d == 768) layernorm_bwd_kernel<3><<<grid, kBwWarps * 32, 0, s>>>(dyp, xp, sp, gamma, ap, op, M, d);
  else if (d == 1024) layernorm_bwd_kernel<4><<<grid, kBwWarps * 32, 0, s>>>(dyp, xp, sp, gamma, ap, op, M, d);
  else layernorm_bwd_kernel<0><<<grid, kBwWarps * 32, 0, s>>>(dyp, xp, sp, gamma, ap, op, M, d);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_gelu(const void* u, void* h, size_t n, void* stream) {
  using namespace mt;
  MT_REQUIRE(u && h, MT_ERR_BAD_SHAPE, "mt_gelu: null pointer");
  MT_REQUIRE(n % 8 == 0, MT_ERR_BAD_SHAPE, "mt_gelu: n must be a multiple of 8");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(h)) & 15) == 0, MT_ERR_BAD_ALIGN,
             "mt_gelu: tensors must be 16-byte aligned");
  if (n == 0) return MT_OK;
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > size_t(sms) * 16) blocks = size_t(sms) * 16;
  gelu_kernel<<<unsigned(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(u), reinterpret_cast<__nv_bfloat16*>(h), n / 8);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}
