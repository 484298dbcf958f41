// mt_attention_generic.cu — softmax attention for head dims other than 64 (API completeness, SURVEY.md §3.4 / §8f
// rank 2: Graph/metatransformer/modules/tokengt_graph_encoder.py:194-203 builds Block(dim=768, num_heads=32), i.e.
// head_dim 24).  Same math as mt_attention.cu (attention.py:31-35), plain SIMT: one query row per thread with q and the
// running output in fp32 registers, K/V tiles of 32 keys staged in shared memory as fp32, online softmax.
// Not a tensor-core kernel: the tuned tcgen05 path covers head_dim 64 (both published encoders); this one exists so
// that other configurations run on the GPU instead of failing.  head_dim must be a multiple of 8 and <= 64.
#include "mt_ptx.cuh"
#include "mt_common.cuh"

namespace mt {

constexpr int kGaThreads = 128;   // queries per block
constexpr int kGaKeys = 32;       // keys per smem tile

template <int DH>
__global__ void __launch_bounds__(kGaThreads)
attention_generic_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int N, int H,
                         float scale_log2) {
  __shared__ float ks[kGaKeys][DH];
  __shared__ float vs[kGaKeys][DH];
  const int bh = blockIdx.y, h = bh % H, b = bh / H;
  const int qi = blockIdx.x * kGaThreads + threadIdx.x;
  const size_t row_stride = size_t(3) * H * DH;
  const __nv_bfloat16* base = qkv + size_t(b) * N * row_stride;
  float q[DH], o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; o[d] = 0.f; }
  if (qi < N) {
    const __nv_bfloat16* qp = base + size_t(qi) * row_stride + size_t(h) * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(qp + d);
      q[d] = bf16_lo(u.x); q[d + 1] = bf16_hi(u.x); q[d + 2] = bf16_lo(u.y); q[d + 3] = bf16_hi(u.y);
      q[d + 4] = bf16_lo(u.z); q[d + 5] = bf16_hi(u.z); q[d + 6] = bf16_lo(u.w); q[d + 7] = bf16_hi(u.w);
    }
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < N; j0 += kGaKeys) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kGaKeys * DH; idx += kGaThreads) {
      const int j = idx / DH, d = idx % DH;
      const int kj = j0 + j;
      float kv = 0.f, vv = 0.f;
      if (kj < N) {
        const __nv_bfloat16* rp = base + size_t(kj) * row_stride + size_t(h) * DH + d;
        kv = __bfloat162float(rp[size_t(H) * DH]);
        vv = __bfloat162float(rp[size_t(2) * H * DH]);
      }
      ks[j][d] = kv;
      vs[j][d] = vv;
    }
    __syncthreads();
    const int nk = min(kGaKeys, N - j0);
    for (int j = 0; j < nk; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], ks[j][d], s);
      s *= scale_log2;
      const float m_new = fmaxf(m, s);
      const float alpha = exp2f(m - m_new);
      const float pj = exp2f(s - m_new);
      l = fmaf(l, alpha, pj);
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(o[d], alpha, pj * vs[j][d]);
      m = m_new;
    }
  }
  if (qi < N) {
    const float inv = 1.0f / l;
    __nv_bfloat16* op = out + (size_t(b) * N + qi) * H * DH + size_t(h) * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 8) {
      uint4 w;
      w.x = pack_bf16x2(o[d] * inv, o[d + 1] * inv); w.y = pack_bf16x2(o[d + 2] * inv, o[d + 3] * inv);
      w.z = pack_bf16x2(o[d + 4] * inv, o[d + 5] * inv); w.w = pack_bf16x2(o[d + 6] * inv, o[d + 7] * inv);
      *reinterpret_cast<uint4*>(op + d) = w;
    }
  }
}

template <int DH>
static int launch_generic(const void* qkv, void* out, int B, int N, int H, float scale, cudaStream_t s) {
  dim3 grid(unsigned(ceil_div(N, kGaThreads)), unsigned(B * H));
  attention_generic_kernel<DH><<<grid, kGaThreads, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                           reinterpret_cast<__nv_bfloat16*>(out), B, N, H,
                                                           scale * 1.4426950408889634f);
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}

int attention_generic_fwd(const void* qkv, void* out, int B, int N, int H, int dh, float scale, cudaStream_t s) {
  MT_REQUIRE(B * H <= 65535, MT_ERR_BAD_SHAPE, "attention (generic head_dim path): B*H must be <= 65535 (got %d)", B * H);
  switch (dh) {
    case 8: return launch_generic<8>(qkv, out, B, N, H, scale, s);
    case 16: return launch_generic<16>(qkv, out, B, N, H, scale, s);
    case 24: return launch_generic<24>(qkv, out, B, N, H, scale, s);
    case 32: return launch_generic<32>(qkv, out, B, N, H, scale, s);
    case 40: return launch_generic<40>(qkv, out, B, N, H, scale, s);
    case 48: return launch_generic<48>(qkv, out, B, N, H, scale, s);
    case 56: return launch_generic<56>(qkv, out, B, N, H, scale, s);
    default:
      set_error("mt_attention_fwd: head_dim %d not supported (64 = tcgen05 path; 8..56 in steps of 8 = generic path)", dh);
      return MT_ERR_UNSUPPORTED;
  }
}

}  // namespace mt
