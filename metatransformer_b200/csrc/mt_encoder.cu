// mt_encoder.cu — whole-encoder driver (E0/E1 in SURVEY.md §8a): the reference's
//   encoder = nn.Sequential(*[Block(dim, num_heads, mlp_ratio=4., qkv_bias=True, norm_layer=nn.LayerNorm,
//                                   act_layer=nn.GELU) for i in range(depth)])          (README.md:125-147)
//   Block.forward: x = x + attn(norm1(x)); x = x + mlp(norm2(x))   (PointCloud/.../layers/attention.py:55-58)
// run as 5 launches per Block on one stream with caller-provided scratch, instead of ~15 eager kernels:
//   GEMM(LN1 folded, qkv, +bias) -> attention -> GEMM(proj,+bias,+x) -> GEMM(LN2 folded, fc1,+bias,GELU) -> GEMM(fc2,+bias,+x)
// The LayerNorm statistics come out of the epilogues of the residual GEMMs (one read-only statistics pass for layer 0),
// and the residual stream x is a split pair hi + lo of bf16 tensors (fp32-grade accumulation over the 2L residual adds
// at 8 B per element of traffic; `hi` alone feeds the tensor cores).
#include "mt_common.cuh"
#include <cuda_runtime.h>
#include <stdlib.h>
#include <vector>
#include <new>

struct mt_encoder {
  int dim, heads, hidden, depth;
  float eps;
  std::vector<mt_block_weights> blocks;
};

namespace {
struct WsLayout {
  size_t h, qkv, attn, mid, mid_lo, xo0, xo1, xo0_lo, xo1_lo, hid, stats, stats2, crow, mr, total;
};
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
WsLayout ws_layout(const mt_encoder* e, size_t M) {
  WsLayout w;
  size_t off = 0;
  auto take = [&](size_t elems) { size_t o = off; off = align_up(off + elems * 2, 1024); return o; };
  w.h = take(M * e->dim);
  w.qkv = take(M * 3 * e->dim);
  w.attn = take(M * e->dim);
  w.mid = take(M * e->dim);
  w.mid_lo = take(M * e->dim);
  w.xo0 = take(M * e->dim);
  w.xo1 = take(M * e->dim);
  w.xo0_lo = take(M * e->dim);
  w.xo1_lo = take(M * e->dim);
  w.hid = take(M * e->hidden);
  const size_t slots = (size_t(e->dim) + 63) / 64;
  w.stats = take(M * 8 * slots);   // float4 per (row, 64-column slice)
  w.stats2 = take(M * 8 * slots);
  w.crow = take(M * 2);            // one float per row
  w.mr = take(M * 4);              // (mean, rstd) per row
  w.total = off;
  return w;
}
}  // namespace

extern "C" int mt_encoder_create(mt_encoder** out, int dim, int heads, int hidden, int depth, float eps,
                                 const mt_block_weights* blocks) {
  using namespace mt;
  MT_REQUIRE(out && blocks, MT_ERR_BAD_SHAPE, "mt_encoder_create: null pointer");
  MT_REQUIRE(dim > 0 && heads > 0 && hidden > 0 && depth > 0, MT_ERR_BAD_SHAPE, "mt_encoder_create: bad sizes");
  MT_REQUIRE(dim % heads == 0 && dim / heads == 64, MT_ERR_UNSUPPORTED,
             "mt_encoder_create: head_dim must be 64 (dim=%d heads=%d)", dim, heads);
  MT_REQUIRE(dim % 32 == 0 && hidden % 32 == 0, MT_ERR_BAD_SHAPE, "mt_encoder_create: dim/hidden must be multiples of 32");
  for (int l = 0; l < depth; ++l) {
    const mt_block_weights& b = blocks[l];
    MT_REQUIRE(b.qkv_w && b.proj_w && b.fc1_w && b.fc2_w, MT_ERR_BAD_SHAPE,
               "mt_encoder_create: block %d has a null weight pointer", l);
    MT_REQUIRE((b.qkv_colsum != nullptr) == (b.fc1_colsum != nullptr), MT_ERR_BAD_SHAPE,
               "mt_encoder_create: block %d: fold both LayerNorms or none", l);
    MT_REQUIRE(b.qkv_colsum || (b.ln1_w && b.ln1_b && b.ln2_w && b.ln2_b), MT_ERR_BAD_SHAPE,
               "mt_encoder_create: block %d has neither folded weights nor LayerNorm parameters", l);
  }
  mt_encoder* e = new (std::nothrow) mt_encoder;
  MT_REQUIRE(e, MT_ERR_CUDA, "mt_encoder_create: out of host memory");
  e->dim = dim; e->heads = heads; e->hidden = hidden; e->depth = depth; e->eps = eps;
  e->blocks.assign(blocks, blocks + depth);
  *out = e;
  return MT_OK;
}

extern "C" int mt_encoder_destroy(mt_encoder* enc) {
  delete enc;
  return MT_OK;
}

extern "C" int mt_encoder_workspace_bytes(const mt_encoder* enc, int B, int N, size_t* bytes) {
  using namespace mt;
  MT_REQUIRE(enc && bytes, MT_ERR_BAD_SHAPE, "mt_encoder_workspace_bytes: null pointer");
  MT_REQUIRE(B > 0 && N > 0, MT_ERR_BAD_SHAPE, "mt_encoder_workspace_bytes: B and N must be positive");
  *bytes = ws_layout(enc, size_t(B) * N).total;
  return MT_OK;
}

// kernels per forward: 7 per Block either way — folded LayerNorms (default): 4 GEMMs + attention + 2 statistics
// finalisations (the first one is the entry pass for layer 0); un-folded: 4 GEMMs + attention + 2 LayerNorm kernels.
extern "C" int mt_encoder_launches_per_forward(const mt_encoder* enc) {
  if (!enc) return 0;
  return enc->blocks[0].qkv_colsum ? 7 * enc->depth : 7 * enc->depth;
}

extern "C" int mt_encoder_forward_ex(mt_encoder* enc, const void* x, int x_dtype, void* y, int y_dtype, int B, int N,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mt;
  MT_REQUIRE(enc && x && y && workspace, MT_ERR_BAD_SHAPE, "mt_encoder_forward: null pointer");
  MT_REQUIRE(B > 0 && N > 0, MT_ERR_BAD_SHAPE, "mt_encoder_forward: B and N must be positive");
  MT_REQUIRE((long long)B * N < (1ll << 31) / 4096, MT_ERR_BAD_SHAPE, "mt_encoder_forward: B*N too large");
  MT_REQUIRE((x_dtype == MT_DT_BF16 || x_dtype == MT_DT_F32) && (y_dtype == MT_DT_BF16 || y_dtype == MT_DT_F32),
             MT_ERR_BAD_DTYPE, "mt_encoder_forward: x / y must be bf16 or fp32");
  const int M = B * N;
  const WsLayout w = ws_layout(enc, size_t(M));
  MT_REQUIRE(workspace_bytes >= w.total, MT_ERR_BAD_SHAPE, "mt_encoder_forward: workspace too small (%zu < %zu)",
             workspace_bytes, w.total);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, MT_ERR_BAD_ALIGN,
             "mt_encoder_forward: workspace must be 256-byte aligned");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 31) == 0, MT_ERR_BAD_ALIGN,
             "mt_encoder_forward: x and y must be 32-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* h = ws + w.h;
  void* qkv = ws + w.qkv;
  void* attn = ws + w.attn;
  void* mid = ws + w.mid;
  void* mid_lo = ws + w.mid_lo;
  void* xo[2] = {ws + w.xo0, ws + w.xo1};
  void* xo_lo[2] = {ws + w.xo0_lo, ws + w.xo1_lo};
  void* hid = ws + w.hid;
  float* stats_a = reinterpret_cast<float*>(ws + w.stats);
  float* stats_b = reinterpret_cast<float*>(ws + w.stats2);
  const int d = enc->dim;
  const size_t n_elems = size_t(M) * d;
  const float scale = 0.125f;  // head_dim^-0.5 with head_dim = 64
  const bool folded = enc->blocks[0].qkv_colsum != nullptr;
  int rc;
  // The residual stream x is the split pair (cur, cur_lo) of bf16 tensors holding x - c_row (the input's row means): the
  // entry pass centres and splits the input into the buffers layer 0 does not write (xo[1]) and leaves the statistics of
  // layer 0's first LayerNorm; the last residual GEMM adds c_row back.
  float* c_row = reinterpret_cast<float*>(ws + w.crow);
  float* mean_rstd = reinterpret_cast<float*>(ws + w.mr);
  const void* cur = x;
  const void* cur_lo = nullptr;
  if (folded) {
    if ((rc = mt_center_split(x, x_dtype, xo[1], xo_lo[1], c_row, mean_rstd, M, d, enc->eps, stream))) return rc;
    cur = xo[1];
    cur_lo = xo_lo[1];
  } else if (x_dtype == MT_DT_F32) {
    if ((rc = mt_cast(x, MT_DT_F32, xo[1], MT_DT_BF16, n_elems, stream))) return rc;
    cur = xo[1];
  }
  for (int l = 0; l < enc->depth; ++l) {
    const mt_block_weights& bw = enc->blocks[l];
    const bool last = (l == enc->depth - 1);
    void* outp = (last && y_dtype == MT_DT_BF16) ? y : xo[l & 1];
    void* outp_lo = (last && y_dtype == MT_DT_BF16) ? nullptr : xo_lo[l & 1];
    if (folded) {
      // Folded LayerNorms.  Layer 0's statistics come from the entry pass; every other LayerNorm input is produced by a
      // residual GEMM (proj / fc2) whose epilogue emits per-row shifted partial sums.
      mt_gemm_opts o = {};
      // the statistics records of the previous layer's fc2 epilogue -> (mean, rstd) per row (layer 0: from the entry pass)
      if (l > 0 && (rc = mt_ln_stats_finalize(stats_a, mean_rstd, M, d, enc->eps, stream))) return rc;
      o.bias = bw.qkv_b; o.ln_colsum = bw.qkv_colsum; o.ln_row_stats = mean_rstd;
      if ((rc = mt_gemm_bf16_ex(cur, bw.qkv_w, qkv, M, 3 * d, d, &o, stream))) return rc;
      if ((rc = mt_attention_fwd(qkv, attn, B, N, enc->heads, 64, scale, stream))) return rc;
      o = mt_gemm_opts{};
      o.bias = bw.proj_b; o.residual = cur; o.residual_lo = cur_lo; o.y_lo = mid_lo; o.out_row_sums = stats_b;
      if ((rc = mt_gemm_bf16_ex(attn, bw.proj_w, mid, M, d, d, &o, stream))) return rc;
      o = mt_gemm_opts{};
      if ((rc = mt_ln_stats_finalize(stats_b, mean_rstd, M, d, enc->eps, stream))) return rc;
      o.bias = bw.fc1_b; o.act = 1; o.ln_colsum = bw.fc1_colsum; o.ln_row_stats = mean_rstd;
      if ((rc = mt_gemm_bf16_ex(mid, bw.fc1_w, hid, M, enc->hidden, d, &o, stream))) return rc;
      o = mt_gemm_opts{};
      o.bias = bw.fc2_b; o.residual = mid; o.residual_lo = mid_lo; o.y_lo = outp_lo;
      if (!last) o.out_row_sums = stats_a;
      else o.y_row_offset = c_row;
      if ((rc = mt_gemm_bf16_ex(hid, bw.fc2_w, outp, M, d, enc->hidden, &o, stream))) return rc;
    } else {
      // un-folded LayerNorm kernels (A/B option): bf16 residual stream
      outp_lo = nullptr;
      if (last && y_dtype != MT_DT_BF16) outp = xo[l & 1];
      if ((rc = mt_layernorm(cur, MT_DT_BF16, bw.ln1_w, bw.ln1_b, h, M, d, enc->eps, stream))) return rc;
      if ((rc = mt_gemm_bf16(h, bw.qkv_w, bw.qkv_b, nullptr, qkv, M, 3 * d, d, 0, 0, stream))) return rc;
      if ((rc = mt_attention_fwd(qkv, attn, B, N, enc->heads, 64, scale, stream))) return rc;
      if ((rc = mt_gemm_bf16(attn, bw.proj_w, bw.proj_b, cur, mid, M, d, d, 0, 0, stream))) return rc;
      if ((rc = mt_layernorm(mid, MT_DT_BF16, bw.ln2_w, bw.ln2_b, h, M, d, enc->eps, stream))) return rc;
      if ((rc = mt_gemm_bf16(h, bw.fc1_w, bw.fc1_b, nullptr, hid, M, enc->hidden, d, 1, 0, stream))) return rc;
      if ((rc = mt_gemm_bf16(hid, bw.fc2_w, bw.fc2_b, mid, outp, M, d, enc->hidden, 0, 0, stream))) return rc;
    }
    cur = outp;
    cur_lo = outp_lo;
  }
  if (y_dtype == MT_DT_F32) {
    if (cur_lo) rc = mt_merge_f32(cur, cur_lo, reinterpret_cast<float*>(y), n_elems, stream);
    else rc = mt_cast(cur, MT_DT_BF16, y, MT_DT_F32, n_elems, stream);
    if (rc) return rc;
  }
  return MT_OK;
}

extern "C" int mt_encoder_forward(mt_encoder* enc, const void* x, void* y, int B, int N, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return mt_encoder_forward_ex(enc, x, MT_DT_BF16, y, MT_DT_BF16, B, N, workspace, workspace_bytes, stream);
}
