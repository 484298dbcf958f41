// mt_patch_embed.cu — the gather half of the Data2Seq conv tokenizers (T1/T2/T3 in SURVEY.md §8a).
// A Conv2d/Conv3d whose kernel tiles the input (Data2Seq/Image.py:16 Conv2d(3,d,16,16);
// Video/models/modeling_finetune.py:285-289 Conv3d(3,d,(2,16,16),(2,16,16)); Audio/src/models/ast_models.py:86
// Conv2d(1,d,(16,16),stride=(fstride,tstride))) is a GEMM over im2row patches:
//   tokens[m, :] = rows[m, :] @ weight.view(d, -1)^T + bias,  m = ((b*Tp + tp)*Hp + hp)*Wp + wp,
//   rows[m, ((c*kt + dt)*kh + dh)*kw + dw] = x[b, c, tp*st + dt, hp*sh + dh, wp*sw + dw]
// which also yields the (B, N, d) token layout of `.flatten(2).transpose(1, 2)` (Image.py:26) directly.
// This kernel builds `rows` in bf16 (cast fused into the gather); the GEMM + bias runs in mt_gemm_bf16.
//
// HBM-bound: algorithmic bytes = B*C*T*H*W*sizeof(in) (read once when stride == kernel) + M*K*2 (write).
// Mapping: one thread per 8 consecutive dw (16-byte bf16 store; 2x16-byte fp32 loads), consecutive threads
// walk K fastest so stores are fully coalesced and loads are whole 32/64-byte segments.
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <stdlib.h>

namespace mt {

struct Im2RowParams {
  int B, C, T, H, W;
  int kt, kh, kw, st, sh, sw;
  int Tp, Hp, Wp;
  int K;            // C*kt*kh*kw
  long long M;      // B*Tp*Hp*Wp
};

template <typename TIn>
MT_DEVINL float ld_as_f32(const TIn* p);
template <> MT_DEVINL float ld_as_f32<float>(const float* p) { return __ldg(p); }
template <> MT_DEVINL float ld_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

// VEC = true: kw % 8 == 0 and every 8-pixel segment is 16-byte (bf16) / 32-byte (fp32) aligned.
template <typename TIn, bool VEC>
__global__ void __launch_bounds__(256)
im2row_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ rows, const Im2RowParams p) {
  const int k8_per_row = p.K >> 3;
  const long long total = p.M * k8_per_row;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long m = idx / k8_per_row;
    const int k0 = int(idx - m * k8_per_row) << 3;
    // token coordinates
    long long r = m;
    const int wp = int(r % p.Wp); r /= p.Wp;
    const int hp = int(r % p.Hp); r /= p.Hp;
    const int tp = int(r % p.Tp);
    const int b = int(r / p.Tp);
    uint4 o;
    if constexpr (VEC) {
      // k0 -> (c, dt, dh, dw0), dw0 multiple of 8
      int kk = k0;
      const int dw0 = kk % p.kw; kk /= p.kw;
      const int dh = kk % p.kh; kk /= p.kh;
      const int dt = kk % p.kt;
      const int c = kk / p.kt;
      const size_t off = (((size_t(b) * p.C + c) * p.T + (tp * p.st + dt)) * p.H + (hp * p.sh + dh)) * size_t(p.W) +
                         size_t(wp * p.sw + dw0);
      if constexpr (sizeof(TIn) == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + off));
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(x + off) + 1);
        o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
        o.z = pack_bf16x2(c4.x, c4.y); o.w = pack_bf16x2(c4.z, c4.w);
      } else {
        o = __ldg(reinterpret_cast<const uint4*>(x + off));
      }
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int kk = k0 + j;
        const int dw = kk % p.kw; kk /= p.kw;
        const int dh = kk % p.kh; kk /= p.kh;
        const int dt = kk % p.kt;
        const int c = kk / p.kt;
        const size_t off = (((size_t(b) * p.C + c) * p.T + (tp * p.st + dt)) * p.H + (hp * p.sh + dh)) * size_t(p.W) +
                           size_t(wp * p.sw + dw);
        v[j] = ld_as_f32<TIn>(x + off);
      }
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(rows + size_t(m) * p.K + k0) = o;
  }
}

// Vector path with the index arithmetic tabled per block (ncu, round 2: the kernel above is issue-bound — 73 % issue-slot
// utilisation at 57 % of DRAM peak, six integer divisions per 16-byte store).  One element per thread and a huge grid,
// like above (two "run per thread" variants with fewer, longer-lived threads measured 1.6-4x slower on C1), but a 256-thread block
// = rpb token rows x tpr K-segments whose K offsets and token base offsets are decoded once per block into shared memory.
constexpr int kTabTok = 8;   // tokens per thread row and block iteration (independent loads in flight per thread)

template <typename TIn>
__global__ void __launch_bounds__(256)
im2row_tab_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ rows, const Im2RowParams p, int tpr, int rpb) {
  __shared__ long long koff_s[256];
  __shared__ long long base_s[256];
  const int k8_per_row = p.K >> 3;
  const int tid = int(threadIdx.x);
  const int rl = tid / tpr;
  const int kt = tid - rl * tpr;
  const int segs = (k8_per_row + tpr - 1) / tpr;          // K segments per thread (1 unless K > 2048)
  const long long plane = (long long)p.H * p.W;
  const int tok_per_iter = rpb * kTabTok;                 // <= 256
  auto decode_k = [&](int sg) {
    if (tid < tpr) {
      int kk = (tid + tpr * sg) << 3;
      const int dw0 = kk % p.kw; kk /= p.kw;
      const int dh = kk % p.kh; kk /= p.kh;
      const int dt = kk % p.kt;
      const int c = kk / p.kt;
      koff_s[tid] = ((long long)(c * p.T + dt) * p.H + dh) * p.W + dw0;
    }
  };
  if (segs == 1) decode_k(0);
  for (long long m0 = (long long)blockIdx.x * tok_per_iter; m0 < p.M; m0 += (long long)gridDim.x * tok_per_iter) {
    __syncthreads();
    if (tid < tok_per_iter) {
      long long r = m0 + tid;
      const int wp = int(r % p.Wp); r /= p.Wp;
      const int hp = int(r % p.Hp); r /= p.Hp;
      const int tp = int(r % p.Tp);
      const long long b = r / p.Tp;
      base_s[tid] = (b * p.C * p.T + (long long)tp * p.st) * plane + (long long)(hp * p.sh) * p.W + wp * p.sw;
    }
    for (int sg = 0; sg < segs; ++sg) {
      if (segs > 1) {
        __syncthreads();
        decode_k(sg);
      }
      __syncthreads();
      const int k8 = kt + tpr * sg;
      if (rl < rpb && k8 < k8_per_row) {
        const long long ko = koff_s[kt];
        uint4 o[kTabTok];
#pragma unroll
        for (int u = 0; u < kTabTok; ++u) {
          const int tl = rl + rpb * u;                     // token of this thread inside the iteration
          if (m0 + tl < p.M) {
            const long long off = base_s[tl] + ko;
            if constexpr (sizeof(TIn) == 4) {
              const float4* src = reinterpret_cast<const float4*>(x + off);
              const float4 a = __ldg(src);
              const float4 c4 = __ldg(src + 1);
              o[u].x = pack_bf16x2(a.x, a.y); o[u].y = pack_bf16x2(a.z, a.w);
              o[u].z = pack_bf16x2(c4.x, c4.y); o[u].w = pack_bf16x2(c4.z, c4.w);
            } else {
              o[u] = __ldg(reinterpret_cast<const uint4*>(x + off));
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kTabTok; ++u) {
          const int tl = rl + rpb * u;
          if (m0 + tl < p.M) *reinterpret_cast<uint4*>(rows + size_t(m0 + tl) * p.K + (k8 << 3)) = o[u];
        }
      }
    }
  }
}

template <typename TIn>
static int launch_im2row(const void* x, void* rows, const Im2RowParams& p, cudaStream_t s) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long total = p.M * (p.K >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  if (blocks < 1) blocks = 1;
  const int elt = int(sizeof(TIn));
  const int seg_align = 16 / elt * ((elt == 4) ? 1 : 1);  // elements per 16 bytes
  // vector path needs each 8-pixel segment start (wp*sw + dw0, row stride W) to sit on a 16-byte boundary
  const bool vec = (p.kw % 8 == 0) && (p.W % seg_align == 0) && (p.sw % seg_align == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (8 % seg_align == 0);
  const TIn* xp = reinterpret_cast<const TIn*>(x);
  __nv_bfloat16* rp = reinterpret_cast<__nv_bfloat16*>(rows);
  // The tabled kernel wins on long patch rows (C5 tubelets, K = 1536: 284 vs 364 us for the whole tokenizer call, same box)
  // and loses on short ones (C1 image patches, K = 768: 235 vs 133 us): default by K, MT_B200_IM2ROW_TAB=0|1 forces one.
  static int use_tab = -2;
  if (use_tab == -2) {
    const char* e = getenv("MT_B200_IM2ROW_TAB");
    use_tab = (e && (e[0] == '0' || e[0] == '1')) ? (e[0] - '0') : -1;
  }
  const bool tab = use_tab >= 0 ? (use_tab == 1) : (p.K >= 1024);
  if (vec && tab) {
    const int k8 = p.K >> 3;
    const int tpr = k8 < 256 ? k8 : 256;
    const int rpb = 256 / tpr;
    long long nb = (p.M + rpb * kTabTok - 1) / (rpb * kTabTok);
    if (nb > (long long)sms * 32) nb = (long long)sms * 32;
    im2row_tab_kernel<TIn><<<unsigned(nb), 256, 0, s>>>(xp, rp, p, tpr, rpb);
  } else if (vec) im2row_kernel<TIn, true><<<unsigned(blocks), 256, 0, s>>>(xp, rp, p);
  else im2row_kernel<TIn, false><<<unsigned(blocks), 256, 0, s>>>(xp, rp, p);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

// Circular 1-D im2row of TokenEmbedding (Data2Seq/Time_Series.py:29-42): Conv1d(c_in, d, k=3, padding=1,
// padding_mode='circular') over time.  rows[b*T + t, c*3 + kk] = x[b, (t + kk - 1) mod T, c]  (K order (c, kk) =
// weight.view(d, -1)); columns 3*C .. Kp-1 are zero so that the row pitch is a multiple of 16 bytes.
template <typename TIn>
__global__ void __launch_bounds__(256) im2row1d_circular_kernel(const TIn* __restrict__ x,
                                                                 __nv_bfloat16* __restrict__ rows, int B, int T, int C,
                                                                 int Kp) {
  const long long total = (long long)B * T * Kp;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = int(idx % Kp);
    const long long m = idx / Kp;
    float v = 0.f;
    if (k < 3 * C) {
      const int c = k / 3, kk = k % 3;
      const int t = int(m % T), b = int(m / T);
      int ts = t + kk - 1;
      ts = ts < 0 ? ts + T : (ts >= T ? ts - T : ts);
      v = ld_as_f32<TIn>(x + (size_t(b) * T + ts) * C + c);
    }
    rows[idx] = __float2bfloat16_rn(v);
  }
}

// out[b, 0, :] = first[:], out[b, 1 + i, :] = tok[b, i, :]  — prepends a (cls + pos) row to every sequence
// (Hyper-spectrum/metatransformer.py:155-157: cat(cls_tokens, x) on the token axis).  16-byte vectors.
__global__ void __launch_bounds__(256) prepend_row_kernel(const __nv_bfloat16* __restrict__ tok,
                                                           const __nv_bfloat16* __restrict__ first,
                                                           __nv_bfloat16* __restrict__ out, int B, int n, int d8) {
  const long long total = (long long)B * (n + 1) * d8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = int(idx % d8);
    const long long r = idx / d8;
    const int i = int(r % (n + 1));
    const long long b = r / (n + 1);
    const uint4 v = (i == 0) ? reinterpret_cast<const uint4*>(first)[c]
                             : reinterpret_cast<const uint4*>(tok)[(b * n + (i - 1)) * d8 + c];
    reinterpret_cast<uint4*>(out)[idx] = v;
  }
}

}  // namespace mt

extern "C" int mt_prepend_row(const void* tok, const void* first, void* out, int B, int n, int d, void* stream) {
  using namespace mt;
  MT_REQUIRE(tok && first && out, MT_ERR_BAD_SHAPE, "mt_prepend_row: null pointer");
  MT_REQUIRE(B > 0 && n > 0 && d > 0 && d % 8 == 0, MT_ERR_BAD_SHAPE, "mt_prepend_row: need B, n > 0 and d %% 8 == 0");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(tok) | reinterpret_cast<uintptr_t>(first) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_prepend_row: pointers must be 16-byte aligned");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long total = (long long)B * (n + 1) * (d / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  prepend_row_kernel<<<unsigned(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(tok), reinterpret_cast<const __nv_bfloat16*>(first),
      reinterpret_cast<__nv_bfloat16*>(out), B, n, d / 8);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_im2row1d_circular(const void* x, int x_dtype, void* rows, int B, int T, int C, int Kp, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && rows, MT_ERR_BAD_SHAPE, "mt_im2row1d_circular: null pointer");
  MT_REQUIRE(B > 0 && T > 0 && C > 0 && Kp >= 3 * C && Kp % 8 == 0, MT_ERR_BAD_SHAPE,
             "mt_im2row1d_circular: need B,T,C > 0 and Kp >= 3*C with Kp %% 8 == 0");
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const long long total = (long long)B * T * Kp;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == MT_DT_F32)
    im2row1d_circular_kernel<float><<<unsigned(blocks), 256, 0, s>>>(reinterpret_cast<const float*>(x),
                                                                     reinterpret_cast<__nv_bfloat16*>(rows), B, T, C, Kp);
  else if (x_dtype == MT_DT_BF16)
    im2row1d_circular_kernel<__nv_bfloat16><<<unsigned(blocks), 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(rows), B, T, C, Kp);
  else {
    set_error("mt_im2row1d_circular: x_dtype must be fp32 or bf16 (got %d)", x_dtype);
    return MT_ERR_BAD_DTYPE;
  }
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

extern "C" int mt_im2row(const void* x, int x_dtype, void* rows, int B, int C, int T, int H, int W, int kt, int kh,
                         int kw, int st, int sh, int sw, void* stream) {
  using namespace mt;
  MT_REQUIRE(x && rows, MT_ERR_BAD_SHAPE, "mt_im2row: null pointer");
  MT_REQUIRE(B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && kt > 0 && kh > 0 && kw > 0 && st > 0 && sh > 0 && sw > 0,
             MT_ERR_BAD_SHAPE, "mt_im2row: all sizes must be positive");
  MT_REQUIRE(T >= kt && H >= kh && W >= kw, MT_ERR_BAD_SHAPE, "mt_im2row: kernel (%d,%d,%d) larger than input (%d,%d,%d)",
             kt, kh, kw, T, H, W);
  Im2RowParams p;
  p.B = B; p.C = C; p.T = T; p.H = H; p.W = W;
  p.kt = kt; p.kh = kh; p.kw = kw; p.st = st; p.sh = sh; p.sw = sw;
  p.Tp = (T - kt) / st + 1; p.Hp = (H - kh) / sh + 1; p.Wp = (W - kw) / sw + 1;
  p.K = C * kt * kh * kw;
  p.M = (long long)B * p.Tp * p.Hp * p.Wp;
  MT_REQUIRE(p.K % 8 == 0, MT_ERR_BAD_SHAPE, "mt_im2row: C*kt*kh*kw must be a multiple of 8 (got %d)", p.K);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(rows) & 15) == 0, MT_ERR_BAD_ALIGN, "mt_im2row: rows must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (x_dtype == MT_DT_F32) return launch_im2row<float>(x, rows, p, s);
  if (x_dtype == MT_DT_BF16) return launch_im2row<__nv_bfloat16>(x, rows, p, s);
  set_error("mt_im2row: x_dtype must be fp32 or bf16 (got %d)", x_dtype);
  return MT_ERR_BAD_DTYPE;
}
