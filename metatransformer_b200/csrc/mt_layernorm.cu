// mt_layernorm.cu — LayerNorm over the feature dim (E2 in SURVEY.md §8a): norm1 / norm2 of the timm Block
// (PointCloud/openpoints/models/layers/attention.py:46,50,56-57; eps = 1e-5 because README.md:131 passes a bare
// nn.LayerNorm).  Statistics in fp32 with the two-pass (mean, then centred second moment) formulation, biased
// variance, exactly as F.layer_norm; output bf16 (it only ever feeds a bf16 tensor-core GEMM).
//
// HBM-bound: one warp owns one row, the row stays in registers between the passes, 16-byte loads/stores,
// algorithmic traffic = M*d*(sizeof(in) + 2) bytes.
#include "mt_ptx.cuh"
#include "mt_common.cuh"

namespace mt {

constexpr int kLnWarps = 8;

template <typename TIn>
MT_DEVINL void load8(const TIn* p, float (&v)[8]);
template <>
MT_DEVINL void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
template <>
MT_DEVINL void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

MT_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NV > 0: d == NV*256 exactly, row held in registers.  NV == 0: any d % 8 == 0, row re-read (L1/L2 hits).
template <int NV, typename TIn>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_kernel(const TIn* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ y, int M, int d, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= M) return;
  const TIn* xr = x + size_t(row) * d;
  __nv_bfloat16* yr = y + size_t(row) * d;
  const float inv_d = 1.0f / float(d);

  if constexpr (NV > 0) {
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      load8<TIn>(xr + (lane + 32 * i) * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float c = v[i][j] - mean; q = fmaf(c, c, q); }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (lane + 32 * i) * 8;
      float g[8], b[8];
      load8<float>(gamma + c0, g);
      load8<float>(beta + c0, b);
      uint4 o;
      o.x = pack_bf16x2(fmaf((v[i][0] - mean) * rstd, g[0], b[0]), fmaf((v[i][1] - mean) * rstd, g[1], b[1]));
      o.y = pack_bf16x2(fmaf((v[i][2] - mean) * rstd, g[2], b[2]), fmaf((v[i][3] - mean) * rstd, g[3], b[3]));
      o.z = pack_bf16x2(fmaf((v[i][4] - mean) * rstd, g[4], b[4]), fmaf((v[i][5] - mean) * rstd, g[5], b[5]));
      o.w = pack_bf16x2(fmaf((v[i][6] - mean) * rstd, g[6], b[6]), fmaf((v[i][7] - mean) * rstd, g[7], b[7]));
      *reinterpret_cast<uint4*>(yr + c0) = o;
    }
  } else {
    const int nchunks = d >> 3;
    float s = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      float v[8];
      load8<TIn>(xr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      float v[8];
      load8<TIn>(xr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = v[j] - mean; q = fmaf(t, t, q); }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    for (int c = lane; c < nchunks; c += 32) {
      float v[8], g[8], b[8];
      load8<TIn>(xr + c * 8, v);
      load8<float>(gamma + c * 8, g);
      load8<float>(beta + c * 8, b);
      uint4 o;
      o.x = pack_bf16x2(fmaf((v[0] - mean) * rstd, g[0], b[0]), fmaf((v[1] - mean) * rstd, g[1], b[1]));
      o.y = pack_bf16x2(fmaf((v[2] - mean) * rstd, g[2], b[2]), fmaf((v[3] - mean) * rstd, g[3], b[3]));
      o.z = pack_bf16x2(fmaf((v[4] - mean) * rstd, g[4], b[4]), fmaf((v[5] - mean) * rstd, g[5], b[5]));
      o.w = pack_bf16x2(fmaf((v[6] - mean) * rstd, g[6], b[6]), fmaf((v[7] - mean) * rstd, g[7], b[7]));
      *reinterpret_cast<uint4*>(yr + c * 8) = o;
    }
  }
}

// Row statistics only: stats[row] = (mean, rstd).  Feeds the LN-folded GEMM epilogue (mt_gemm_bf16_ln), which
// makes the normalised activations never touch HBM.  Algorithmic traffic: M*d*sizeof(in) + 8*M bytes.
template <int NV, typename TIn>
__global__ void __launch_bounds__(kLnWarps * 32)
row_stats_kernel(const TIn* __restrict__ x, float2* __restrict__ stats, int M, int d, float eps,
                 unsigned long long* tl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  pdl_launch_dependents();
  pdl_wait();
  timeline_start(tl);
  if (row >= M) { timeline_end(tl); return; }
  const TIn* xr = x + size_t(row) * d;
  const float inv_d = 1.0f / float(d);
  float mean, q = 0.f;
  if constexpr (NV > 0) {
    // d == NV*256: all of the row's 16-byte loads are issued up front and the row stays in registers
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) load8<TIn>(xr + (lane + 32 * i) * 8, v[i]);
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    mean = warp_sum(s) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = v[i][j] - mean; q = fmaf(t, t, q); }
  } else {
    const int nchunks = d >> 3;
    float s = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      float v[8];
      load8<TIn>(xr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    mean = warp_sum(s) * inv_d;
    for (int c = lane; c < nchunks; c += 32) {   // second pass hits L1 (row <= 16 KB)
      float v[8];
      load8<TIn>(xr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = v[j] - mean; q = fmaf(t, t, q); }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
  __syncthreads();
  timeline_end(tl);
}

// Entry of the whole-encoder driver: per row  c = mean(x),  xc = x - c  ->  split pair hi = bf16(xc), lo = bf16(xc - hi),
// plus the LayerNorm statistics of the CENTRED row (mean 0, rstd).  LayerNorm is shift-invariant and the residual adds
// commute with a per-row constant, so the whole Block stack runs on the centred stream unchanged and c is added back in
// the last epilogue.  Why: the LN-folded GEMMs consume bf16(x) instead of bf16(LN(x)), whose rounding error scales with
// |x| instead of |x - mean|; centring removes the row offset before anything is rounded (rows with |mean| >> std).
// One warp per row, 16-byte loads / 8-byte-pair stores; traffic = M*d*(sizeof(in) + 4) bytes.
template <typename TIn>
__global__ void __launch_bounds__(kLnWarps * 32)
center_split_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                    float* __restrict__ c_row, float2* __restrict__ stats, int M, int d, float eps, unsigned long long* tl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  timeline_start(tl);
  if (row < M) {
    const TIn* xr = x + size_t(row) * d;
    const float inv_d = 1.0f / float(d);
    const int nchunks = d >> 3;
    float s = 0.f;
    for (int c = lane; c < nchunks; c += 32) {
      float v[8];
      load8<TIn>(xr + c * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
    for (int c = lane; c < nchunks; c += 32) {   // second pass hits L1 (row <= 16 KB)
      float v[8];
      load8<TIn>(xr + c * 8, v);
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float a = v[j] - mean, b = v[j + 1] - mean;
        q = fmaf(a, a, fmaf(b, b, q));
        const uint32_t hh = pack_bf16x2(a, b);
        h[j >> 1] = hh;
        l[j >> 1] = pack_bf16x2(a - bf16_lo(hh), b - bf16_hi(hh));
      }
      *reinterpret_cast<uint4*>(hi + size_t(row) * d + c * 8) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(lo + size_t(row) * d + c * 8) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) {
      c_row[row] = mean;
      stats[row] = make_float2(0.f, rstd);
    }
  }
  __syncthreads();
  timeline_end(tl);
}

// LayerNorm statistics records of a residual GEMM ([M][S][4]: shift, sum(v - shift), sum((v - shift)^2), -; one per
// 64-column slice, see mt_gemm.cu) -> (mean, rstd) per row.  One thread per row, the S <= 64 records of a row are
// contiguous (S * 16 bytes); every slice is re-referenced to the first slice's shift, so nothing cancels.
__global__ void __launch_bounds__(256)
ln_stats_finalize_kernel(const float4* __restrict__ rec, float2* __restrict__ stats, int M, int S, int d, float eps,
                         unsigned long long* tl) {
  timeline_start(tl);
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < M) {
    const float4* ps = rec + size_t(row) * S;
    const float ref = ps[0].x;
    float S1 = 0.f, S2 = 0.f;
#pragma unroll 4
    for (int i = 0; i < S; ++i) {
      const float4 r = ps[i];
      const float n_i = float(min(64, d - 64 * i));
      const float dd = r.x - ref;
      S1 += fmaf(n_i, dd, r.y);
      S2 += fmaf(dd, fmaf(n_i, dd, 2.0f * r.y), r.z);
    }
    const float inv_d = 1.0f / float(d);
    const float m1 = S1 * inv_d;
    stats[row] = make_float2(ref + m1, rsqrtf(fmaxf(fmaf(S2, inv_d, -m1 * m1), 0.f) + eps));
  }
  __syncthreads();
  timeline_end(tl);
}

// same for a split (hi + lo) row: the statistics of hi + lo (residual stream of the per-Block path)
__global__ void __launch_bounds__(kLnWarps * 32)
row_stats_split_kernel(const __nv_bfloat16* __restrict__ xh, const __nv_bfloat16* __restrict__ xl, float2* __restrict__ stats,
                       int M, int d, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kLnWarps + warp;
  if (row >= M) return;
  const __nv_bfloat16* hr = xh + size_t(row) * d;
  const __nv_bfloat16* lr = xl + size_t(row) * d;
  const float inv_d = 1.0f / float(d);
  const int nchunks = d >> 3;
  float s = 0.f;
  for (int c = lane; c < nchunks; c += 32) {
    float v[8], w[8];
    load8<__nv_bfloat16>(hr + c * 8, v);
    load8<__nv_bfloat16>(lr + c * 8, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j] + w[j];
  }
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
  for (int c = lane; c < nchunks; c += 32) {   // second pass hits L1 (row <= 16 KB)
    float v[8], w[8];
    load8<__nv_bfloat16>(hr + c * 8, v);
    load8<__nv_bfloat16>(lr + c * 8, w);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float t = (v[j] - mean) + w[j]; q = fmaf(t, t, q); }
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

template <typename TIn>
static int launch_row_stats(const void* x, float* stats, int M, int d, float eps, cudaStream_t s) {
  const TIn* xp = reinterpret_cast<const TIn*>(x);
  float2* sp = reinterpret_cast<float2*>(stats);
  const unsigned grid = unsigned(ceil_div(M, kLnWarps));
  const unsigned block = kLnWarps * 32;
  unsigned long long* tl = timeline_next_slot();
  if (d == 768) MT_CHECK_CUDA(launch_pdl(row_stats_kernel<3, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, sp, M, d, eps, tl));
  else if (d == 1024) MT_CHECK_CUDA(launch_pdl(row_stats_kernel<4, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, sp, M, d, eps, tl));
  else MT_CHECK_CUDA(launch_pdl(row_stats_kernel<0, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, sp, M, d, eps, tl));
  return MT_OK;
}

// One-time weight preparation for the folded form of  Linear(LayerNorm(x)):
//   LN(x) W^T + b = rstd * (x W'^T - mean * colsum) + b',   W' = W * gamma (per input channel, rounded to bf16),
//   colsum[n] = sum_k W'[n,k] (of the ROUNDED values, so the mean term cancels exactly),  b' = b + W beta.
template <typename TW>
__global__ void __launch_bounds__(kLnWarps * 32)
fold_ln_kernel(const TW* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ w_out, float* __restrict__ colsum,
               float* __restrict__ bias_out, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * kLnWarps + warp;
  if (n >= N) return;
  float cs = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    float wv;
    if constexpr (sizeof(TW) == 4) wv = w[size_t(n) * K + k];
    else wv = __bfloat162float(w[size_t(n) * K + k]);
    const __nv_bfloat16 wf = __float2bfloat16_rn(wv * gamma[k]);
    w_out[size_t(n) * K + k] = wf;
    cs += __bfloat162float(wf);
    bb = fmaf(wv, beta[k], bb);
  }
  cs = warp_sum(cs);
  bb = warp_sum(bb);
  if (lane == 0) {
    colsum[n] = cs;
    bias_out[n] = bb + (bias ? bias[n] : 0.f);
  }
}

template <typename TIn>
static int launch_ln(const void* x, const float* gamma, const float* beta, void* y, int M, int d, float eps,
                     cudaStream_t s) {
  const TIn* xp = reinterpret_cast<const TIn*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const unsigned grid = unsigned(ceil_div(M, kLnWarps));
  const unsigned block = kLnWarps * 32;
  if (d == 768) MT_CHECK_CUDA(launch_pdl(layernorm_kernel<3, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, gamma, beta, yp, M, d, eps));
  else if (d == 1024) MT_CHECK_CUDA(launch_pdl(layernorm_kernel<4, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, gamma, beta, yp, M, d, eps));
  else MT_CHECK_CUDA(launch_pdl(layernorm_kernel<0, TIn>, dim3(grid), dim3(block), 0, s, 1u, xp, gamma, beta, yp, M, d, eps));
  return MT_OK;
}

}  // namespace mt

extern "C" int mt_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int M, int d,
                            float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && gamma && beta && y, MT_ERR_BAD_SHAPE, "mt_layernorm: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_layernorm: need M > 0 and d %% 8 == 0 (M=%d d=%d)", M,
             d);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
               reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_layernorm: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == MT_DT_BF16) return launch_ln<__nv_bfloat16>(x, gamma, beta, y, M, d, eps, s);
  if (x_dtype == MT_DT_F32) return launch_ln<float>(x, gamma, beta, y, M, d, eps, s);
  set_error("mt_layernorm: x_dtype must be bf16 or fp32 (got %d)", x_dtype);
  return MT_ERR_BAD_DTYPE;
}

extern "C" int mt_row_stats(const void* x, int x_dtype, float* stats, int M, int d, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && stats, MT_ERR_BAD_SHAPE, "mt_row_stats: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_row_stats: need M > 0 and d %% 8 == 0 (M=%d d=%d)", M, d);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
             MT_ERR_BAD_ALIGN, "mt_row_stats: x must be 16-byte and stats 8-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == MT_DT_BF16) return launch_row_stats<__nv_bfloat16>(x, stats, M, d, eps, s);
  if (x_dtype == MT_DT_F32) return launch_row_stats<float>(x, stats, M, d, eps, s);
  set_error("mt_row_stats: x_dtype must be bf16 or fp32 (got %d)", x_dtype);
  return MT_ERR_BAD_DTYPE;
}

extern "C" int mt_center_split(const void* x, int x_dtype, void* hi, void* lo, float* c_row, float* stats, int M, int d,
                               float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && hi && lo && c_row && stats, MT_ERR_BAD_SHAPE, "mt_center_split: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_center_split: need M > 0 and d %% 8 == 0 (M=%d d=%d)", M, d);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(c_row) & 3) == 0,
             MT_ERR_BAD_ALIGN, "mt_center_split: x / hi / lo must be 16-byte, stats 8-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = unsigned(ceil_div(M, kLnWarps));
  unsigned long long* tl = timeline_next_slot();
  __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* lp = reinterpret_cast<__nv_bfloat16*>(lo);
  float2* sp = reinterpret_cast<float2*>(stats);
  if (x_dtype == MT_DT_BF16)
    center_split_kernel<__nv_bfloat16><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), hp, lp, c_row,
                                                                      sp, M, d, eps, tl);
  else if (x_dtype == MT_DT_F32)
    center_split_kernel<float><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const float*>(x), hp, lp, c_row, sp, M, d, eps, tl);
  else {
    set_error("mt_center_split: x_dtype must be bf16 or fp32 (got %d)", x_dtype);
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_ln_stats_finalize(const float* records, float* stats, int M, int d, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(records && stats, MT_ERR_BAD_SHAPE, "mt_ln_stats_finalize: null pointer");
  MT_REQUIRE(M > 0 && d > 0, MT_ERR_BAD_SHAPE, "mt_ln_stats_finalize: M and d must be positive");
  MT_REQUIRE((reinterpret_cast<uintptr_t>(records) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
             MT_ERR_BAD_ALIGN, "mt_ln_stats_finalize: records 16-byte, stats 8-byte aligned");
  ln_stats_finalize_kernel<<<unsigned(ceil_div(M, 256)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(records), reinterpret_cast<float2*>(stats), M, (d + 63) / 64, d, eps,
      timeline_next_slot());
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_row_stats_split(const void* x_hi, const void* x_lo, float* stats, int M, int d, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x_hi && x_lo && stats, MT_ERR_BAD_SHAPE, "mt_row_stats_split: null pointer");
  MT_REQUIRE(M > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_row_stats_split: need M > 0 and d %% 8 == 0 (M=%d d=%d)", M, d);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x_hi) | reinterpret_cast<uintptr_t>(x_lo)) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(stats) & 7) == 0,
             MT_ERR_BAD_ALIGN, "mt_row_stats_split: x must be 16-byte and stats 8-byte aligned");
  row_stats_split_kernel<<<unsigned(ceil_div(M, kLnWarps)), kLnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_hi), reinterpret_cast<const __nv_bfloat16*>(x_lo),
      reinterpret_cast<float2*>(stats), M, d, eps);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_fold_layernorm(const void* w, int w_dtype, const float* gamma, const float* beta, const float* bias,
                                 void* w_folded, float* colsum, float* bias_folded, int N, int K, void* stream) {
  using namespace mt;
  MT_REQUIRE(w && gamma && beta && w_folded && colsum && bias_folded, MT_ERR_BAD_SHAPE, "mt_fold_layernorm: null pointer");
  MT_REQUIRE(N > 0 && K > 0, MT_ERR_BAD_SHAPE, "mt_fold_layernorm: N and K must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const unsigned grid = unsigned(ceil_div(N, kLnWarps));
  __nv_bfloat16* wo = reinterpret_cast<__nv_bfloat16*>(w_folded);
  if (w_dtype == MT_DT_F32)
    fold_ln_kernel<float><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const float*>(w), gamma, beta, bias, wo, colsum,
                                                         bias_folded, N, K);
  else if (w_dtype == MT_DT_BF16)
    fold_ln_kernel<__nv_bfloat16><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(w), gamma, beta,
                                                                 bias, wo, colsum, bias_folded, N, K);
  else {
    set_error("mt_fold_layernorm: w_dtype must be fp32 or bf16 (got %d)", w_dtype);
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}
