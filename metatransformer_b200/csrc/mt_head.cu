// mt_head.cu — the steps either side of the encoder stack and the training-time dropout /
// stochastic-depth pass (§8f rank 2).  All HBM-bound streaming kernels: 16-byte vectors, fp32 arithmetic.
//   * mt_assemble_tokens:    out[b] = [prefix rows ; tok[b]] + pos   (cls / dist tokens + positional embedding in one
//                            pass: Audio/src/models/ast_models.py:157-159, Hyper-spectrum/metatransformer.py:155-157)
//   * mt_pool_layernorm:     final LayerNorm fused with the pooling each fork applies before its head:
//                            norm(x)[:,0] (Video/models/modeling_finetune.py:449, Hyper-spectrum/metatransformer.py:163-165),
//                            fc_norm(x.mean(1)) (modeling_finetune.py:446-447), (norm(x)[:,0] + norm(x)[:,1]) / 2
//                            (ast_models.py:163-164)
//   * mt_pool_layernorm_bwd: its autograd (dx for all tokens, d gamma, d beta)
//   * mt_dropout:            out = resid + row_scale[row / rows_per_scale] * keep(i) / (1 - p) * x  — nn.Dropout of
//                            Mlp / Attention.proj_drop (mlp.py:31-34, attention.py:37) and timm DropPath
//                            (attention.py:48,56-57) with a counter-based generator, so the backward regenerates the
//                            mask from (seed, offset) instead of storing it
//   * mt_pad_cols:           zero padding of text features to the encoder width (Data2Seq/Text.py:15-19)
#include "mt_ptx.cuh"
#include "mt_common.cuh"

namespace mt {

MT_DEVINL void ld8_bf16_h(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
MT_DEVINL uint4 pack8_bf16(const float (&o)[8]) {
  uint4 w;
  w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
  w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
  return w;
}

// ------------------------------------------------------------------------------------------------
// token assembly
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok, const __nv_bfloat16* __restrict__ prefix,
                       const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ out, int B, int n, int k,
                       int d8) {
  const int T = n + k;
  const long long total = (long long)B * T * d8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = int(idx % d8);
    const long long r = idx / d8;
    const int i = int(r % T);
    const long long b = r / T;
    uint4 v = (i < k) ? reinterpret_cast<const uint4*>(prefix)[size_t(i) * d8 + c]
                      : reinterpret_cast<const uint4*>(tok)[(b * n + (i - k)) * d8 + c];
    if (pos != nullptr) {
      float a[8], q[8];
      ld8_bf16_h(reinterpret_cast<const __nv_bfloat16*>(&v), a);
      ld8_bf16_h(pos + (size_t(i) * d8 + c) * 8, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += q[j];
      v = pack8_bf16(a);
    }
    reinterpret_cast<uint4*>(out)[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// embedding-bag sum: out[r, :] = pos[r % pos_period, :] + sum_f table[idx[r, f] + offs[f], :]
// (calendar embeddings of the time-series tokenizer, Data2Seq/Time_Series.py:64-92; atom / bond feature embeddings of
// the graph tokenizer, Data2Seq/Graph.py:254-255).  Gather-bound: one thread per (row, 4 columns), fp32 tables.
// ------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256)
embed_sum_kernel(const float* __restrict__ table, const int* __restrict__ idx, const int* __restrict__ offs, int F,
                 const float* __restrict__ pos, int pos_period, TOut* __restrict__ out, long long R, int d4) {
  const long long total = R * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % d4);
    const long long r = i / d4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos != nullptr) acc = reinterpret_cast<const float4*>(pos)[(r % pos_period) * d4 + c];
    for (int f = 0; f < F; ++f) {
      const long long row = (long long)idx[r * F + f] + (offs ? offs[f] : 0);
      const float4 t = reinterpret_cast<const float4*>(table)[row * d4 + c];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if constexpr (sizeof(TOut) == 4) {
      reinterpret_cast<float4*>(out)[i] = acc;
    } else {
      uint2 o;
      o.x = pack_bf16x2(acc.x, acc.y);
      o.y = pack_bf16x2(acc.z, acc.w);
      reinterpret_cast<uint2*>(out)[i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// final LayerNorm + pooling
// ------------------------------------------------------------------------------------------------
constexpr int kPoolThreads = 256;

MT_DEVINL float block_sum(float v, float* red) {  // red: 8 floats of shared memory; all threads get the sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kPoolThreads / 32; ++w) t += red[w];
  return t;
}

// Gathers the pooled pre-norm row(s) of sample b into shared memory (fp32): mode 0 -> token 0; mode 1 -> mean over the
// N tokens; mode 2 -> tokens 0 and 1.  rows[r * d + c]; returns the number of rows (1 or 2).
MT_DEVINL int pooled_rows(const __nv_bfloat16* __restrict__ xb, int N, int d, int mode, float* rows, float* scratch) {
  const int d8 = d >> 3;
  if (mode != 1) {
    const int nr = (mode == 2) ? 2 : 1;
    for (int i = threadIdx.x; i < nr * d8; i += kPoolThreads) {
      float v[8];
      ld8_bf16_h(xb + size_t(i) * 8, v);   // rows 0 (and 1) are contiguous at the start of the sample
#pragma unroll
      for (int j = 0; j < 8; ++j) rows[i * 8 + j] = v[j];
    }
    __syncthreads();
    return nr;
  }
  // mean over tokens: thread = (column vector cv, row group rg); partial sums per row group in `scratch`
  for (int cv0 = 0; cv0 < d8; cv0 += kPoolThreads) {   // d8 > 256 (d > 2048): several column passes
    const int span = min(d8 - cv0, kPoolThreads);
    const int g = (span < kPoolThreads) ? max(1, min(kPoolThreads / span, 8)) : 1;
    const int cv = cv0 + int(threadIdx.x) % span, rg = int(threadIdx.x) / span;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (rg < g) {
      for (int t = rg; t < N; t += g) {
        float v[8];
        ld8_bf16_h(xb + (size_t(t) * d8 + cv) * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) scratch[size_t(rg) * d + cv * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < span * 8; c += kPoolThreads) {
      float t = 0.f;
      for (int q = 0; q < g; ++q) t += scratch[size_t(q) * d + cv0 * 8 + c];
      rows[cv0 * 8 + c] = t / float(N);
    }
    __syncthreads();
  }
  return 1;
}

// two-pass mean / rstd of a row held in shared memory
MT_DEVINL void row_mean_rstd(const float* row, int d, float eps, float* red, float& mean, float& rstd) {
  float s = 0.f;
  for (int c = threadIdx.x; c < d; c += kPoolThreads) s += row[c];
  mean = block_sum(s, red) / float(d);
  float q = 0.f;
  for (int c = threadIdx.x; c < d; c += kPoolThreads) {
    const float t = row[c] - mean;
    q = fmaf(t, t, q);
  }
  rstd = rsqrtf(block_sum(q, red) / float(d) + eps);
}

__global__ void __launch_bounds__(kPoolThreads)
pool_layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                      const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int N, int d, int mode,
                      float eps) {
  extern __shared__ float sm[];
  float* rows = sm;                 // [2][d]
  float* scratch = sm + 2 * d;      // [8][d] (mode 1 only)
  __shared__ float red[8];
  const int b = blockIdx.x;
  const __nv_bfloat16* xb = x + size_t(b) * N * d;
  const int nr = pooled_rows(xb, N, d, mode, rows, scratch);
  float mean[2], rstd[2];
  for (int r = 0; r < nr; ++r) row_mean_rstd(rows + r * d, d, eps, red, mean[r], rstd[r]);
  const float w = 1.0f / float(nr);
  for (int c = threadIdx.x; c < d; c += kPoolThreads) {
    float o = 0.f;
    for (int r = 0; r < nr; ++r) o += w * ((rows[r * d + c] - mean[r]) * rstd[r] * gamma[c] + beta[c]);
    out[size_t(b) * d + c] = __float2bfloat16_rn(o);
  }
}

__global__ void __launch_bounds__(kPoolThreads)
pool_layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                          const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, int N, int d, int mode, float eps) {
  extern __shared__ float sm[];
  float* rows = sm;                 // [2][d]: pooled rows, overwritten by their input gradients
  float* scratch = sm + 2 * d;
  __shared__ float red[8];
  const int b = blockIdx.x;
  const __nv_bfloat16* xb = x + size_t(b) * N * d;
  const int nr = pooled_rows(xb, N, d, mode, rows, scratch);
  const float w = 1.0f / float(nr);
  for (int r = 0; r < nr; ++r) {
    float* row = rows + r * d;
    float mean, rstd;
    row_mean_rstd(row, d, eps, red, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
    for (int c = threadIdx.x; c < d; c += kPoolThreads) {
      const float gy = w * __bfloat162float(dy[size_t(b) * d + c]);
      const float xh = (row[c] - mean) * rstd;
      const float g = gy * gamma[c];
      s1 += g;
      s2 = fmaf(g, xh, s2);
      if (dgamma != nullptr) atomicAdd(dgamma + c, gy * xh);
      if (dbeta != nullptr && r == 0) atomicAdd(dbeta + c, gy * float(nr));   // d beta = dy (both rows share it)
    }
    const float m1 = block_sum(s1, red) / float(d), m2 = block_sum(s2, red) / float(d);
    for (int c = threadIdx.x; c < d; c += kPoolThreads) {
      const float gy = w * __bfloat162float(dy[size_t(b) * d + c]);
      const float xh = (row[c] - mean) * rstd;
      float v = rstd * (gy * gamma[c] - m1 - xh * m2);
      if (mode == 1) v /= float(N);
      row[c] = v;
    }
    __syncthreads();
  }
  // write the (B, N, d) input gradient: pooled rows get their gradient, every other token zero (mode 1: all rows share it)
  const int d8 = d >> 3;
  __nv_bfloat16* dxb = dx + size_t(b) * N * d;
  for (int i = threadIdx.x; i < N * d8; i += kPoolThreads) {
    const int t = i / d8, cv = i - t * d8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int src = (mode == 1) ? 0 : (t < nr ? t : -1);
    if (src >= 0) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rows[src * d + cv * 8 + j];
      v = pack8_bf16(o);
    }
    reinterpret_cast<uint4*>(dxb)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// dropout / stochastic depth with a counter-based generator (Philox4x32-10)
// ------------------------------------------------------------------------------------------------
MT_DEVINL uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(kM0, ctr.x), lo0 = kM0 * ctr.x;
    const uint32_t hi1 = __umulhi(kM1, ctr.z), lo1 = kM1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += kW0;
    key.y += kW1;
  }
  return ctr;
}

// one thread = 8 consecutive elements = one Philox call (eight 16-bit lanes); keep iff lane < thresh (= (1-p) * 65536)
__global__ void __launch_bounds__(256)
dropout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ resid,
               const float* __restrict__ row_scale, int rows_per_scale, __nv_bfloat16* __restrict__ out,
               long long n8, int d8, uint32_t thresh, float inv_keep, uint2 seed, unsigned long long offset) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    float v[8];
    ld8_bf16_h(x + idx * 8, v);
    float s = inv_keep;
    if (row_scale != nullptr) s *= row_scale[(idx / d8) / rows_per_scale];
    if (thresh < 65536u) {
      const unsigned long long c = offset + (unsigned long long)idx;
      const uint4 r = philox4x32_10(make_uint4(uint32_t(c), uint32_t(c >> 32), 0u, 0u), seed);
      const uint32_t lanes[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[2 * j + 0] = ((lanes[j] & 0xffffu) < thresh) ? v[2 * j + 0] * s : 0.f;
        v[2 * j + 1] = ((lanes[j] >> 16) < thresh) ? v[2 * j + 1] * s : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= s;
    }
    if (resid != nullptr) {
      float a[8];
      ld8_bf16_h(resid + idx * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += a[j];
    }
    reinterpret_cast<uint4*>(out)[idx] = pack8_bf16(v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pad_cols_kernel(const T* __restrict__ x, T* __restrict__ out, long long rows,
                                                       int c_in, int c_out) {
  const long long total = rows * c_out;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / c_out;
    const int c = int(idx - r * c_out);
    out[idx] = (c < c_in) ? x[r * c_in + c] : T(0.f);
  }
}

static int grid_for(long long work_items, int* out_blocks) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  long long blocks = (work_items + 255) / 256;
  if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  if (blocks < 1) blocks = 1;
  *out_blocks = int(blocks);
  return MT_OK;
}

}  // namespace mt

extern "C" int mt_assemble_tokens(const void* tok, const void* prefix, const void* pos, void* out, int B, int n, int k,
                                  int d, void* stream) {
  using namespace mt;
  MT_REQUIRE(tok && out, MT_ERR_BAD_SHAPE, "mt_assemble_tokens: null pointer");
  MT_REQUIRE(B > 0 && n > 0 && k >= 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE,
             "mt_assemble_tokens: need B, n > 0, k >= 0 and d %% 8 == 0");
  MT_REQUIRE(k == 0 || prefix != nullptr, MT_ERR_BAD_SHAPE, "mt_assemble_tokens: k > 0 needs the prefix rows");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(tok) | reinterpret_cast<uintptr_t>(prefix) | reinterpret_cast<uintptr_t>(pos) |
               reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_assemble_tokens: pointers must be 16-byte aligned");
  int blocks = 0;
  int rc = grid_for((long long)B * (n + k) * (d / 8), &blocks);
  if (rc) return rc;
  assemble_tokens_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(tok), reinterpret_cast<const __nv_bfloat16*>(prefix),
      reinterpret_cast<const __nv_bfloat16*>(pos), reinterpret_cast<__nv_bfloat16*>(out), B, n, k, d / 8);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

static int pool_check(const char* who, int B, int N, int d, int mode) {
  using namespace mt;
  MT_REQUIRE(B > 0 && N > 0 && d > 0 && d % 8 == 0 && d <= 4096, MT_ERR_BAD_SHAPE,
             "%s: need B, N > 0 and d %% 8 == 0, d <= 4096 (got %d,%d,%d)", who, B, N, d);
  MT_REQUIRE(mode == MT_POOL_FIRST || mode == MT_POOL_MEAN || mode == MT_POOL_FIRST_TWO, MT_ERR_UNSUPPORTED,
             "%s: unknown pooling mode %d", who, mode);
  MT_REQUIRE(mode != MT_POOL_FIRST_TWO || N >= 2, MT_ERR_BAD_SHAPE, "%s: MT_POOL_FIRST_TWO needs N >= 2", who);
  return MT_OK;
}

static size_t pool_smem(int d, int mode) { return size_t(mode == MT_POOL_MEAN ? 10 : 2) * d * sizeof(float); }

extern "C" int mt_pool_layernorm(const void* x, const float* gamma, const float* beta, void* out, int B, int N, int d,
                                 int mode, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && gamma && beta && out, MT_ERR_BAD_SHAPE, "mt_pool_layernorm: null pointer");
  int rc = pool_check("mt_pool_layernorm", B, N, d, mode);
  if (rc) return rc;
  MT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, MT_ERR_BAD_ALIGN, "mt_pool_layernorm: x must be 16-byte aligned");
  const size_t smem = pool_smem(d, mode);
  if (smem > 48 * 1024)
    MT_CHECK_CUDA(cudaFuncSetAttribute(pool_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  pool_layernorm_kernel<<<B, kPoolThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), N, d, mode, eps);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_pool_layernorm_bwd(const void* dy, const void* x, const float* gamma, void* dx, float* dgamma,
                                     float* dbeta, int B, int N, int d, int mode, float eps, void* stream) {
  using namespace mt;
  MT_REQUIRE(dy && x && gamma && dx, MT_ERR_BAD_SHAPE, "mt_pool_layernorm_bwd: null pointer");
  int rc = pool_check("mt_pool_layernorm_bwd", B, N, d, mode);
  if (rc) return rc;
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, MT_ERR_BAD_ALIGN,
             "mt_pool_layernorm_bwd: x and dx must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dgamma) MT_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, size_t(d) * sizeof(float), s));
  if (dbeta) MT_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, size_t(d) * sizeof(float), s));
  const size_t smem = pool_smem(d, mode);
  if (smem > 48 * 1024)
    MT_CHECK_CUDA(cudaFuncSetAttribute(pool_layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  pool_layernorm_bwd_kernel<<<B, kPoolThreads, smem, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), gamma,
      reinterpret_cast<__nv_bfloat16*>(dx), dgamma, dbeta, N, d, mode, eps);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_dropout(const void* x, const void* residual, const float* row_scale, int rows_per_scale, void* out,
                          long long rows, int d, float p, unsigned long long seed, unsigned long long offset,
                          void* stream) {
  using namespace mt;
  MT_REQUIRE(x && out, MT_ERR_BAD_SHAPE, "mt_dropout: null pointer");
  MT_REQUIRE(rows > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_dropout: need rows > 0 and d %% 8 == 0");
  MT_REQUIRE(p >= 0.f && p < 1.f, MT_ERR_BAD_SHAPE, "mt_dropout: p must be in [0, 1) (got %f)", double(p));
  MT_REQUIRE(row_scale == nullptr || rows_per_scale > 0, MT_ERR_BAD_SHAPE, "mt_dropout: rows_per_scale must be positive");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_dropout: pointers must be 16-byte aligned");
  const long long n8 = rows * (d / 8);
  int blocks = 0;
  int rc = grid_for(n8, &blocks);
  if (rc) return rc;
  // 16-bit lanes: keep probability quantised to 1/65536; the rescale uses the quantised value so E[out] = x exactly
  uint32_t thresh = uint32_t(double(1.0f - p) * 65536.0 + 0.5);
  if (p == 0.f || thresh > 65536u) thresh = 65536u;
  MT_REQUIRE(thresh > 0u, MT_ERR_BAD_SHAPE, "mt_dropout: p too close to 1");
  const float inv_keep = float(65536.0 / double(thresh));
  dropout_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(residual), row_scale,
      rows_per_scale, reinterpret_cast<__nv_bfloat16*>(out), n8, d / 8, thresh, inv_keep,
      make_uint2(uint32_t(seed), uint32_t(seed >> 32)), offset);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_embed_sum(const float* table, const int* idx, const int* offsets, int F, const float* pos, int pos_period,
                            void* out, int out_dtype, long long rows, int d, void* stream) {
  using namespace mt;
  MT_REQUIRE(table && idx && out, MT_ERR_BAD_SHAPE, "mt_embed_sum: null pointer");
  MT_REQUIRE(rows > 0 && F > 0 && d > 0 && d % 4 == 0, MT_ERR_BAD_SHAPE, "mt_embed_sum: need rows, F > 0 and d %% 4 == 0");
  MT_REQUIRE(pos == nullptr || pos_period > 0, MT_ERR_BAD_SHAPE, "mt_embed_sum: pos needs pos_period > 0");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_embed_sum: table / pos / out must be 16-byte aligned");
  const long long total = rows * (d / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == MT_DT_F32)
    embed_sum_kernel<float><<<unsigned(blocks), 256, 0, s>>>(table, idx, offsets, F, pos, pos_period,
                                                             reinterpret_cast<float*>(out), rows, d / 4);
  else if (out_dtype == MT_DT_BF16)
    embed_sum_kernel<__nv_bfloat16><<<unsigned(blocks), 256, 0, s>>>(table, idx, offsets, F, pos, pos_period,
                                                                     reinterpret_cast<__nv_bfloat16*>(out), rows, d / 4);
  else {
    set_error("mt_embed_sum: out_dtype must be bf16 or fp32");
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_pad_cols(const void* x, void* out, int dtype, long long rows, int c_in, int c_out, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && out, MT_ERR_BAD_SHAPE, "mt_pad_cols: null pointer");
  MT_REQUIRE(rows > 0 && c_in > 0 && c_out >= c_in, MT_ERR_BAD_SHAPE,
             "mt_pad_cols: need rows > 0 and 0 < c_in <= c_out (got %d -> %d)", c_in, c_out);
  int blocks = 0;
  int rc = grid_for(rows * c_out, &blocks);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == MT_DT_F32)
    pad_cols_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(x), reinterpret_cast<float*>(out), rows,
                                                  c_in, c_out);
  else if (dtype == MT_DT_BF16)
    pad_cols_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                          reinterpret_cast<__nv_bfloat16*>(out), rows, c_in, c_out);
  else {
    set_error("mt_pad_cols: dtype must be fp32 or bf16 (got %d)", dtype);
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}
