/* mt_b200.h — C ABI of libmt_b200.so: the B200 (sm_100a) kernels behind Meta-Transformer's shared encoder path.
 *
 * The reference has NO native boundary on this path: it runs timm's Block / Attention / Mlp and the Data2Seq
 * tokenizers as eager PyTorch (README.md:116-149).  Each entry point below replaces the torch.nn.functional
 * calls of one reference span (cited per function; in-tree verbatim copies of the timm code are used for
 * line numbers because timm itself is a pip dependency that is not vendored).
 *
 * Conventions
 *  - plain C types only; device pointers are `void*` / typed pointers to DEVICE memory unless stated otherwise;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing here synchronises the host;
 *  - the caller owns every buffer (activations, outputs, workspaces); the library never allocates on the hot path;
 *  - every function returns 0 (MT_OK) or a negative mt_status; mt_last_error() gives a thread-local message;
 *  - bf16 tensors are row-major with the feature dimension contiguous; all pointers 16-byte aligned.
 */
#ifndef MT_B200_H_
#define MT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MT_OK = 0,
  MT_ERR_BAD_SHAPE = -1,
  MT_ERR_BAD_DTYPE = -2,
  MT_ERR_BAD_ALIGN = -3,
  MT_ERR_UNSUPPORTED = -4,
  MT_ERR_CUDA = -5
} mt_status;

typedef enum { MT_DT_BF16 = 0, MT_DT_F32 = 1, MT_DT_F16 = 2 } mt_dtype;

/* ---- library / device ---- */
int mt_version(void);                 /* MAJOR*10000 + MINOR*100 + PATCH */
const char* mt_last_error(void);      /* thread-local, never NULL */
int mt_device_ok(int device);         /* 0 if `device` is compute capability 10.x (B200), else MT_ERR_UNSUPPORTED */

/* ---- E3/E5/E6: Linear (+bias) (+GELU) (+residual)  --------------------------------------------------------
 * y[M,N] = act(x[M,K] @ w[N,K]^T + bias[N]) (+ residual[M,N]);  x,w,residual,y bf16; bias fp32 (may be NULL).
 * act: 0 = identity, 1 = exact-erf GELU (nn.GELU()), 2 = y = acc * gelu'(residual) (backward).  cta_group: 0 = auto, 1 = single-CTA tiles, 2 = CTA pairs.
 * Replaces F.linear at PointCloud/openpoints/models/layers/attention.py:28 (qkv), :36 (proj) and
 * mlp.py:30-33 (fc1, act, fc2), plus the residual adds of Block.forward (attention.py:56-57).           */
int mt_gemm_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y,
                 int M, int N, int K, int act, int cta_group, void* stream);

/* ---- E2+E3 / E2+E6 fused: Linear(LayerNorm(x)) with the LayerNorm folded into the GEMM ------------------------
 * LN(x) W^T + b = rstd_m * (x W'^T - mean_m * colsum_n) + b'_n  with  W' = W*gamma (bf16), colsum_n = sum_k W'[n,k],
 * b' = b + W beta  (prepared once by mt_fold_layernorm) and per-row (mean, rstd) from mt_row_stats.  The
 * normalised activations never exist in HBM.  Same reference spans as mt_layernorm + mt_gemm_bf16
 * (attention.py:56 `attn(norm1(x))` -> qkv, :57 `mlp(norm2(x))` -> fc1).                                      */
int mt_row_stats(const void* x, int x_dtype, float* stats /* [M][2] = mean, rstd */, int M, int d, float eps,
                 void* stream);
/* (mean, rstd) per row from the statistics records a residual GEMM emitted with `out_row_sums` (see mt_gemm_opts):
 * records [M][ceil(d/64)][4] -> stats [M][2].  One tiny pass instead of every CTA column of the consuming GEMM
 * re-reading (and re-combining) all the records of its rows.                                                      */
int mt_ln_stats_finalize(const float* records, float* stats, int M, int d, float eps, void* stream);
/* same for a split (hi + lo) bf16 row: statistics of hi + lo in fp32 */
int mt_row_stats_split(const void* x_hi, const void* x_lo, float* stats, int M, int d, float eps, void* stream);
/* Entry transform of the whole-encoder driver: c_row[m] = mean(x[m,:]); (hi, lo) = split(x - c_row) as bf16 [M,d] each;
 * stats[m] = (0, rstd) = LayerNorm statistics of the centred row.  LayerNorm is shift-invariant and residual adds
 * commute with a per-row constant, so the Block stack runs unchanged on the centred stream (bf16(x - mean) instead of
 * bf16(x) feeds the LN-folded GEMMs: no loss when |mean| >> std) and c_row comes back through y_row_offset below. */
int mt_center_split(const void* x, int x_dtype, void* hi, void* lo, float* c_row, float* stats, int M, int d, float eps,
                    void* stream);
int mt_fold_layernorm(const void* w /* [N,K] */, int w_dtype, const float* gamma, const float* beta,
                      const float* bias /* [N] or NULL */, void* w_folded /* bf16 [N,K] */, float* colsum /* [N] */,
                      float* bias_folded /* [N] */, int N, int K, void* stream);
int mt_gemm_bf16_ln(const void* x, const void* w_folded, const float* bias_folded, const float* colsum,
                    const float* row_stats, void* y, int M, int N, int K, int act, int cta_group, void* stream);

/* General form of the two calls above.
 *
 * Split residual stream.  Block.forward adds two branch outputs per layer to x (attention.py:56-57); with a bf16-only x
 * the 2L roundings dominate the error of the whole encoder (1.8e-2 of max|x| after 12 layers).  The residual GEMMs
 * (proj, fc2) therefore keep x as a PAIR of bf16 tensors hi + lo (16 mantissa bits): `residual` / `residual_lo` are
 * read, added to the fp32 accumulator, and the sum is written as `y` = bf16(v) and `y_lo` = bf16(v - y).  `y` alone is
 * the A operand of the next GEMM; either lo pointer may be NULL (a bf16 input has no lo; a bf16 result needs none).
 *
 * `out_row_sums` lets a GEMM whose output is the next LayerNorm's input emit that LayerNorm's statistics from its
 * epilogue: for every output row one record per 64-column slice, layout [M][S][4] fp32 with S = ceil(N/64):
 * (shift, sum(v - shift), sum((v - shift)^2), 0) with shift = the row's first value in the slice (shifted sums: no
 * cancellation when |mean| >> std; plain stores in a fixed layout => deterministic, no zero-initialisation).  The
 * consuming LN-folded GEMM takes them with ln_stats_are_sums = S = ceil(K/64) and combines them pairwise-exactly
 * (mean_i = shift_i + s1_i/n_i, M2 = sum M2_i + n_i (mean_i - mean)^2): no separate statistics pass over HBM.       */
typedef struct {
  const float* bias;            /* [N] or NULL */
  const void* residual;         /* bf16 [M,N] or NULL (act 2: the saved pre-activation); 32-byte aligned */
  int act;                      /* 0, 1 (GELU), 2 (times gelu'(residual)) */
  const float* ln_colsum;       /* LN folding of the INPUT (with ln_row_stats), or NULL */
  const float* ln_row_stats;    /* [M][2] (mean, rstd), or [M][S][4] records if ln_stats_are_sums = S > 0 */
  int ln_stats_are_sums;
  float ln_eps;                 /* used when ln_stats_are_sums */
  float* out_row_sums;          /* [M][ceil(N/64)][4] statistics records of the OUTPUT rows, or NULL */
  int cta_group;                /* 0 auto, 1, 2 */
  int residual_period;          /* > 0: `residual` is a [residual_period, N] table and output row m adds row
                                   m % residual_period (positional embedding broadcast over the batch:
                                   Audio/src/models/ast_models.py:159, Video/models/modeling_finetune.py:434-436) */
  const void* residual_lo;      /* bf16 [M,N] low half of the residual, or NULL */
  void* y_lo;                   /* bf16 [M,N] low half of the output, or NULL */
  const float* y_row_offset;    /* [M] or NULL: added to every element of output row m before it is rounded / split
                                   (undoes mt_center_split in the last residual GEMM); residual GEMMs only */
} mt_gemm_opts;
int mt_gemm_bf16_ex(const void* x, const void* w, void* y, int M, int N, int K, const mt_gemm_opts* opts, void* stream);

/* Split-K form for weight gradients: y_f32[M,N] = x[M,K] @ w[N,K]^T in fp32, the K range split over `splits` CTA pairs
 * per output tile (0 = auto) and reduced with fp32 atomics (y is zeroed inside).  dW = dY^T X with x = dY^T, w = X^T.  */
int mt_gemm_bf16_splitk(const void* x, const void* w, float* y_f32, int M, int N, int K, int splits, void* stream);

/* Weight gradient of a Linear straight from the row-major activations: dw_f32[N_out, K_in] = dy[T, N_out]^T @ x[T, K_in]
 * (bf16 operands consumed MN-major by tcgen05 — no transposed copies —, fp32 accumulation over the T tokens, split over
 * `splits` CTA pairs per output tile (0 = auto) and reduced with fp32 atomics; dw is zeroed inside).
 * Autograd of nn.Linear.weight in fine-tune mode (Audio/src/traintest.py:126-143) and of the conv tokenizers.       */
int mt_gemm_bf16_wgrad(const void* dy, const void* x, float* dw_f32, int T, int N_out, int K_in, int splits, void* stream);

/* ---- fp8 (e4m3) opt-in Linear --------------------------------------------------------------
 * mt_quantize_fp8: q[i] = sat_e4m3(x[i] / scale) (per-tensor scale; x bf16, n % 8 == 0).
 * mt_gemm_fp8:     y = act(scale * (xq @ wq^T) + bias) (+ residual); xq [M,K] / wq [N,K] e4m3 bytes, K % 16 == 0,
 *                  scale = scale_x * scale_w, bias fp32, residual / y bf16.  tcgen05.mma.kind::f8f6f4, fp32 accumulation.
 * An inference option for callers that own a calibration (the reference has none); the encoder driver does not use it. */
int mt_quantize_fp8(const void* src_bf16, void* dst_e4m3, size_t n, float scale, void* stream);
int mt_gemm_fp8(const void* xq, const void* wq, const float* bias, const void* residual, void* y, int M, int N, int K, int act,
                float scale, void* stream);

/* ---- E2: LayerNorm over the last dim, fp32 statistics, biased variance -------------------------------------
 * y[M,d] (bf16) = (x - mean) * rsqrt(var + eps) * gamma + beta;  x bf16 or fp32 (x_dtype), gamma/beta fp32.
 * Replaces norm1 / norm2 = nn.LayerNorm(d) of Block (attention.py:46,50; eps 1e-5 per README.md:131).       */
int mt_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* y,
                 int M, int d, float eps, void* stream);

/* ---- E8 (backward pieces) -----------------------------------------------------------------------------------
 * mt_layernorm_bwd: dx = add + rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma (gamma NULL = already folded),
 *                   xhat = (x - mean)*rstd with (mean, rstd) from mt_row_stats; dy, x, add, dx bf16 [M,d].
 * mt_gelu:          h = gelu(u), bf16, n % 8 == 0 (training-mode forward keeps u for mt_gemm_bf16 act = 2).
 * Autograd of Block.forward (attention.py:55-58) when only the tokenizer trains (Audio/src/models/ast_models.py:66-70). */
int mt_layernorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const void* add, void* dx,
                     int M, int d, void* stream);
int mt_gelu(const void* u, void* h, size_t n, void* stream);
/* weight-gradient helpers for the trainable tokenizers: dW = dY^T X is run as the TN GEMM on transposed copies
 * (mt_transpose_bf16: dst[C, ldd] = src[R, C]^T, zero-padded to ldd >= R), db = column sums of dY (fp32).  */
int mt_transpose_bf16(const void* src, void* dst, int R, int C, int ldd, void* stream);
int mt_colsum(const void* x, float* out, int M, int N, void* stream);
/* weight gradients of a LayerNorm-folded Linear (fine-tune mode): from dWf = dY^T xhat (fp32 [N,K]) and dyb = colsum(dY):
 * dW = dWf*gamma + dyb (x) beta, dgamma = sum_n dWf*W, dbeta = sum_n dyb*W (fp32 in / out; dw may alias dw_folded). */
int mt_ln_fold_bwd(const void* dw_folded /* fp32 [N,K] */, const void* w, int w_dtype, const float* gamma, const float* beta,
                   const float* dyb, float* dw, float* dgamma, float* dbeta, int N, int K, void* stream);

/* ---- E4: softmax attention core -----------------------------------------------------------------------------
 * qkv: bf16 [B, N, 3, H, dh] (the packed output of the qkv Linear, attention.py:28), out: bf16 [B, N, H*dh].
 * out[b,n,h,:] = softmax_j(scale * q[b,n,h]·k[b,j,h]) · v[b,j,h]   (attention.py:31-35, no mask, no dropout).
 * dh = 64: tcgen05 kernel (both published encoders); dh in {8,16,...,56}: generic SIMT kernel (e.g. Graph: 24). */
int mt_attention_fwd(const void* qkv, void* out, int B, int N, int H, int dh, float scale, void* stream);

/* Same, additionally writing lse2[B,H,N] = log2-sum-exp of the scaled scores (needed by mt_attention_bwd). */
int mt_attention_fwd_lse(const void* qkv, void* out, float* lse2, int B, int N, int H, int dh, float scale,
                         void* stream);
/* Backward of the attention core (E8): dqkv [B,N,3,H,64] bf16 from qkv, out, d(out) and the forward's lse2.
 * workspace: mt_attention_bwd_workspace_bytes (fp32 dQ accumulator + row-wise delta = sum(dout*out)).       */
int mt_attention_bwd_workspace_bytes(int B, int N, int H, size_t* bytes);
int mt_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                     void* workspace, size_t workspace_bytes, int B, int N, int H, int dh, float scale, void* stream);

/* Training forms with attention dropout (timm Attention.attn_drop, attention.py:33; the Graph fork trains with
 * attn_drop = 0.1, Graph/metatransformer/modules/tokengt_graph_encoder.py:200): attn = softmax(..); attn = keep * attn /
 * (1 - drop_p); out = attn @ v.  keep(b, h, q, k) is a hash of (seed, b*H + h, q*N + k) evaluated inside the kernels —
 * nothing is stored; the backward and mt_attention_dropout_mask (uint8 [B,H,N,N], test aid) replay it.  dh = 64 only
 * (smaller heads are zero-padded to 64 lanes by the host modules).  lse2 may be NULL in the forward.              */
int mt_attention_fwd_ex(const void* qkv, void* out, float* lse2, int B, int N, int H, int dh, float scale, float drop_p,
                        unsigned long long seed, void* stream);
int mt_attention_bwd_ex(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                        void* workspace, size_t workspace_bytes, int B, int N, int H, int dh, float scale, float drop_p,
                        unsigned long long seed, void* stream);
int mt_attention_dropout_mask(void* mask, int B, int N, int H, float drop_p, unsigned long long seed, void* stream);

/* Debugging hook (tools/attn_trace.py): in builds compiled with -DMT_ATTN_TRACE=1 the following attention forward
 * launches log (clock64 << 8 | event id) per warp of CTAs 0 and 1 into buf (device memory, 2 x 12 x 1024 int64);
 * nullptr switches it off.  Returns 1 (and does nothing) in the shipped build, which has no trace points. */
int mt_attention_set_trace(void* buf);
/* Same for the GEMM kernel (tools/gemm_trace.py, -DMT_GEMM_TRACE=1): cycles every warp of CTAs 0-3 spends in its barrier
 * waits, 4 x 20 x 8 int64.  Returns 1 in the shipped build. */
int mt_gemm_set_trace(void* buf);

/* ---- T1/T2/T3: patch / spectrogram / tubelet embedding -----------------------------------------------------
 * Non-overlapping or strided Conv{2,3}d as im2row + GEMM.
 * mt_im2row: x [B, C, T, H, W] (x_dtype fp32 or bf16; T = kt = 1 for 2-D) -> rows bf16 [B*Tp*Hp*Wp, C*kt*kh*kw]
 * in the K-order (c, kt, kh, kw) of conv.weight.view(d, -1); token order (t, h, w) row-major
 * (Data2Seq/Image.py:16,26; Video/models/modeling_finetune.py:285-297; Audio/src/models/ast_models.py:86).   */
int mt_im2row(const void* x, int x_dtype, void* rows, int B, int C, int T, int H, int W,
              int kt, int kh, int kw, int st, int sh, int sw, void* stream);

/* ---- T4: time-series / IMU token embedding --------------------------------------------------------------------
 * Circular Conv1d(c_in, d, k=3, padding=1) of TokenEmbedding (Data2Seq/Time_Series.py:29-42) as im2row + GEMM:
 * x [B,T,C] -> rows bf16 [B*T, Kp], rows[b*T+t, c*3+kk] = x[b, (t+kk-1) mod T, c], zero-padded to Kp (% 8 == 0).
 * The sinusoid positional embedding (Time_Series.py:8-26) is added by the GEMM's residual input.               */
int mt_im2row1d_circular(const void* x, int x_dtype, void* rows, int B, int T, int C, int Kp, void* stream);

/* ---- hyperspectral tokenizer helper: out[b,0,:] = first, out[b,1+i,:] = tok[b,i,:] (cls-token concat,
 * Hyper-spectrum/metatransformer.py:155-157); bf16, d % 8 == 0 ---- */
int mt_prepend_row(const void* tok, const void* first, void* out, int B, int n, int d, void* stream);

/* ---- steps either side of the encoder ---------------------------------------------------
 * mt_assemble_tokens: out[b, i, :] = (i < k ? prefix[i] : tok[b, i-k]) + pos[i]   (bf16; prefix [k,d] may be NULL when
 *   k == 0, pos [(n+k), d] may be NULL): cls / dist tokens and the positional embedding in one HBM pass
 *   (Audio/src/models/ast_models.py:157-159, Hyper-spectrum/metatransformer.py:155-158).
 * mt_pool_layernorm: the final LayerNorm fused with the pooling each fork applies before its head; x [B,N,d] bf16,
 *   out [B,d] bf16.  MT_POOL_FIRST: norm(x)[:,0] (Video/models/modeling_finetune.py:449, Hyper-spectrum/
 *   metatransformer.py:163-165); MT_POOL_MEAN: fc_norm(x.mean(1)) (modeling_finetune.py:446-447);
 *   MT_POOL_FIRST_TWO: (norm(x)[:,0] + norm(x)[:,1]) / 2 (ast_models.py:163-164).
 * mt_pool_layernorm_bwd: its autograd: dx [B,N,d] bf16 (zero for tokens that are not pooled), dgamma / dbeta fp32 [d]
 *   (either may be NULL; zeroed inside, accumulated with fp32 atomics over the batch).                              */
typedef enum { MT_POOL_FIRST = 0, MT_POOL_MEAN = 1, MT_POOL_FIRST_TWO = 2 } mt_pool_mode;
int mt_assemble_tokens(const void* tok, const void* prefix, const void* pos, void* out, int B, int n, int k, int d,
                       void* stream);
int mt_pool_layernorm(const void* x, const float* gamma, const float* beta, void* out, int B, int N, int d, int mode,
                      float eps, void* stream);
int mt_pool_layernorm_bwd(const void* dy, const void* x, const float* gamma, void* dx, float* dgamma, float* dbeta,
                          int B, int N, int d, int mode, float eps, void* stream);

/* ---- training-time dropout / stochastic depth ---------------------------------------------
 * out[r, c] = residual[r, c] + row_scale[r / rows_per_scale] * keep(r, c) / (1 - p) * x[r, c]      (bf16 [rows, d])
 * keep() is a Philox4x32-10 stream addressed by (seed, offset + element index / 8): the same (seed, offset) gives the
 * same mask, so the backward pass calls this again on the gradient instead of storing the mask.  residual and
 * row_scale may be NULL.  nn.Dropout in Mlp / Attention.proj_drop (PointCloud/openpoints/models/layers/mlp.py:31-34,
 * attention.py:37) = p > 0, no row_scale; timm DropPath (attention.py:48,56-57) = p = 0 with row_scale[b] =
 * bernoulli(keep)/keep per sample and rows_per_scale = tokens per sample.  A call advances the stream by rows*d/8. */
int mt_dropout(const void* x, const void* residual, const float* row_scale, int rows_per_scale, void* out,
               long long rows, int d, float p, unsigned long long seed, unsigned long long offset, void* stream);

/* ---- embedding-bag sum: out[r,:] = pos[r % pos_period,:] + sum_f table[idx[r,f] + offsets[f], :] -----------------------
 * table fp32 [V,d] (several nn.Embedding weights concatenated row-wise, `offsets` = first row of each, or NULL), idx
 * int32 [rows,F], pos fp32 [pos_period,d] or NULL, out bf16 / fp32 [rows,d].  Calendar embeddings of the time-series
 * tokenizer (Data2Seq/Time_Series.py:64-92,120-126) and atom / bond feature sums of the graph tokenizer
 * (Data2Seq/Graph.py:254-255).                                                                                    */
int mt_embed_sum(const float* table, const int* idx, const int* offsets, int F, const float* pos, int pos_period,
                 void* out, int out_dtype, long long rows, int d, void* stream);

/* ---- text features: zero padding to the encoder width (Data2Seq/Text.py:15-19); dtype fp32 or bf16 ---- */
int mt_pad_cols(const void* x, void* out, int dtype, long long rows, int c_in, int c_out, void* stream);

/* ---- dtype conversion helper (weights packing, fp32 activations at the boundary) ---- */
int mt_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t n, void* stream);

/* split representation helpers: hi = bf16(x), lo = bf16(x - hi) / dst = hi + lo (fp32); n % 4 == 0 */
int mt_split_f32(const float* src, void* hi, void* lo, size_t n, void* stream);
int mt_merge_f32(const void* hi, const void* lo, float* dst, size_t n, void* stream);

/* ---- E7: fp32-accuracy path (plain-fp32 callers: X-Ray/train.py:52-60; north_star gate 1e-4) ----------------------
 * Linear layers as "3xTF32" on the tensor cores: every fp32 operand is a PAIR big = x with the low 13 mantissa bits
 * cleared (exact in TF32) and small = x - big; x.w = small_x.big_w + big_x.small_w + big_x.big_w with fp32
 * accumulation (tcgen05.mma.kind::tf32).  Producers write pairs directly; the residual stream is plain fp32.
 *   mt_split_tf32:          src -> (big, small), n % 4 == 0 (weights, once)
 *   mt_layernorm_f32_split: LayerNorm (fp32, two-pass statistics) -> (big, small)
 *   mt_gemm_f32x3:          y = act(x.w^T + bias) (+ residual fp32); y_small == NULL: y is the fp32 result, else (y,
 *                           y_small) is its pair (feeds the next GEMM); act 0 | 1 (exact-erf GELU); K % 4, N % 16 == 0
 *   mt_attention_f32:       softmax(q k^T scale) v on fp32 qkv [B,N,3,H,dh] (dh 16 | 24 | 32 | 64), SIMT, output pair  */
int mt_split_tf32(const float* src, float* big, float* small, size_t n, void* stream);
int mt_layernorm_f32_split(const float* x, const float* gamma, const float* beta, float* big, float* small, int M, int d,
                           float eps, void* stream);
int mt_gemm_f32x3(const float* x_big, const float* x_small, const float* w_big, const float* w_small, const float* bias,
                  const float* residual, float* y, float* y_small, int M, int N, int K, int act, void* stream);
int mt_attention_f32(const float* qkv, float* out_big, float* out_small, int B, int N, int H, int dh, float scale,
                     void* stream);

/* ---- E8: gradient all-reduce for integrators that own an NCCL communicator ------------------------------------------
 * In place: buf = scale * sum over ranks of buf (scale = 1 / world_size for DDP averaging,
 * Video/run_class_finetuning.py:739-741).  `nccl_comm` is the caller's ncclComm_t; ncclAllReduce is resolved with dlopen
 * at first use, so the library itself does not link NCCL.  The Python package drives the same collective through
 * torch.distributed (metatransformer_b200/distributed.py: GradBucketReducer, overlapped with the backward).        */
int mt_allreduce(void* nccl_comm, void* buf, size_t count, int dtype, float scale, void* stream);

/* ---- device timeline (measurement aid, off by default) -----------------------------------------------------------
 * Between mt_timeline_begin and mt_timeline_end every instrumented launch (GEMM, attention forward, row statistics)
 * takes the next slot of `slots` (device memory, capacity x 2 x uint64): slot[0] = min over CTAs of %globaltimer at
 * kernel start, slot[1] = max at kernel end (atomicMin / atomicMax: initialise the buffer to {~0, 0} before running).
 * Launches captured into a CUDA graph keep their slot, so one replay of the graph fills the table with the start/end
 * times of every kernel inside the replay — the gaps between them are what the per-kernel durations do not show.
 * mt_timeline_end returns the number of slots handed out.  Process-wide, not thread-safe.                          */
int mt_timeline_begin(void* slots, int capacity);
int mt_timeline_end(void);

/* ---- E0/E1: whole-encoder driver ---------------------------------------------------------------------------
 * nn.Sequential(*[Block(dim, heads, mlp_ratio, qkv_bias=True) for _ in range(depth)]) (README.md:125-147).
 * Weights are borrowed device pointers (bf16 matrices, fp32 vectors) that must outlive the handle.          */
typedef struct mt_encoder mt_encoder;

typedef struct {
  const float* ln1_w; const float* ln1_b;      /* [d]            norm1.{weight,bias}      */
  const void* qkv_w;  const float* qkv_b;      /* [3d,d] bf16, [3d] fp32 (NULL = no bias) */
  const void* proj_w; const float* proj_b;     /* [d,d],  [d]    */
  const float* ln2_w; const float* ln2_b;      /* [d]            norm2                    */
  const void* fc1_w;  const float* fc1_b;      /* [hidden,d], [hidden] */
  const void* fc2_w;  const float* fc2_b;      /* [d,hidden], [d]      */
  /* Optional LN folding: when both are non-NULL, qkv_w/qkv_b and fc1_w/fc1_b hold the outputs of mt_fold_layernorm
   * (W*gamma, b + W beta) and these are the matching column sums; ln*_w / ln*_b are then unused.              */
  const float* qkv_colsum;                      /* [3d] or NULL */
  const float* fc1_colsum;                      /* [hidden] or NULL */
} mt_block_weights;

int mt_encoder_create(mt_encoder** out, int dim, int heads, int hidden, int depth, float eps,
                      const mt_block_weights* blocks /* [depth], host array of device pointers */);
int mt_encoder_destroy(mt_encoder* enc);
/* bytes of scratch needed for a (B, N) forward */
int mt_encoder_workspace_bytes(const mt_encoder* enc, int B, int N, size_t* bytes);
/* x, y: bf16 [B, N, dim] (y may alias x); workspace: device scratch of at least workspace_bytes */
int mt_encoder_forward(mt_encoder* enc, const void* x, void* y, int B, int N,
                       void* workspace, size_t workspace_bytes, void* stream);
/* General form: x / y are bf16 or fp32 [B, N, dim] (x_dtype / y_dtype = MT_DT_BF16 | MT_DT_F32).  Inside, the residual
 * stream is the split hi + lo bf16 pair in either case; an fp32 x is split on entry (no precision lost to bf16), an
 * fp32 y is hi + lo merged on exit.                                                                              */
int mt_encoder_forward_ex(mt_encoder* enc, const void* x, int x_dtype, void* y, int y_dtype, int B, int N,
                          void* workspace, size_t workspace_bytes, void* stream);
/* number of kernels one mt_encoder_forward launches (for accounting) */
int mt_encoder_launches_per_forward(const mt_encoder* enc);

#ifdef __cplusplus
}
#endif
#endif /* MT_B200_H_ */
