// mt_gemm.cu — bf16 "TN" GEMM with fused epilogue for the encoder's four Linear layers
//   Y[M,N] = epi( X[M,K] · W[N,K]^T + bias[N] ) (+ R[M,N])
// (E3 qkv, E5 proj+residual, E6 fc1+GELU / fc2+residual in SURVEY.md §8a; the reference runs these as
//  F.linear + separate GELU / add kernels: PointCloud/openpoints/models/layers/attention.py:28,36,56-57,
//  mlp.py:29-35.)
//
// Blackwell design (sm_100a only):
//   * persistent CTAs (one per SM), static round-robin tile schedule, N-tile index fastest so the CTAs that
//     run concurrently share the same activation rows and the whole weight matrix in L2;
//   * operands staged by TMA (128B swizzle) into a kStages-deep smem ring with full/empty mbarriers;
//   * tcgen05.mma (kind::f16, bf16 in / fp32 accumulate) issued by one thread, accumulators in TMEM,
//     two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * CG = 2: CTA pairs (cta_group::2): one 256x256 output tile per pair, each CTA loads its own 128 rows of X
//     and half (128 rows) of the W tile, halving per-SM operand traffic from L2;
//   * 8 epilogue warps: TMEM -> registers (tcgen05.ld) -> bias / exact-erf GELU / residual -> bf16 ->
//     swizzled smem -> TMA store (coalesced, clips the M/N tails in hardware).
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warp 3 idle, warps 4..11 = epilogue (warp%4 selects the TMEM lane quarter, (warp-4)/4 the column half).
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <stdlib.h>

namespace mt {

constexpr int kBM = 128;         // rows of X per CTA
constexpr int kBN = 256;         // UMMA N
constexpr int kBK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + 32 * kEpiWarps;
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBN;  // 512
constexpr int kEpiStageBytes = 32 * 128;     // per-warp staging: 32 rows x 64 bf16
constexpr int kBiasBytes = 2 * 128 * 4;      // per-warp slices: 128 fp32 bias + 128 fp32 LN-fold column sums

template <int CG>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;                       // 16 KB
  static constexpr int kBRows = (CG == 2) ? kBN / 2 : kBN;            // rows of W this CTA loads
  static constexpr int kBBytes = kBRows * kBK * 2;                    // 16 / 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (CG == 2) ? 5 : 3;
  static constexpr int kSmemEpiOff = kStages * kStageBytes;
  static constexpr int kSmemBiasOff = kSmemEpiOff + kEpiWarps * kEpiStageBytes;
  static constexpr int kSmemBarOff = kSmemBiasOff + kEpiWarps * kBiasBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
  static constexpr int kSmemTmemPtrOff = kSmemBarOff + kNumBars * 8;
  static constexpr int kSmemBytes = kSmemTmemPtrOff + 16 + 1024;  // +1024: manual 1 KB alignment slack
};

struct GemmParams {
  const float* bias;            // [N] fp32 or nullptr
  const __nv_bfloat16* resid;   // [M, ldr] bf16 (or [resid_period, ldr], see below) or nullptr
  int ldr;
  int resid_ld256;              // residual rows are 32-byte aligned: fetch them with 256-bit loads
  int resid_period;             // > 0: output row m reads residual row m % resid_period (a [n, d] positional table
                                // broadcast over the batch: tokenizer GEMMs)
  int M, N, K;
  int act;                      // 0 = identity, 1 = exact-erf GELU, 2 = multiply by gelu'(resid) (backward)
  int num_m_tiles, num_n_tiles;
  // LayerNorm folded into the GEMM (see mt_fold_layernorm): y = rstd_m * (acc - mean_m * colsum_n) + bias_n
  const float2* row_stats;      // [M] (mean, rstd) — or (sum, sum of squares) when ln_sums != 0 — or nullptr
  int ln_sums;                  // > 0: row_stats holds ln_sums partial (sum, sum of squares) pairs per row, written by
                                // the epilogue of the GEMM that produced this GEMM's input
  float ln_eps;
  // split-K accumulation mode (weight gradients: few output tiles, very long K): work item = (tile, K split); the
  // epilogue adds the fp32 partial tile to out_f32 [M,N] with red.global.add (zero-initialised by the host wrapper)
  float* out_f32;
  int split_k;                  // >= 1
  int kb_per_split;             // k-blocks per split (every split is non-empty)
  float* out_row_sums;          // [M][2*num_n_tiles][2]: per-(row, 128-column slice) partial (sum, sum of squares) of
                                // this GEMM's output, or nullptr (plain stores: deterministic, no zero-init)
  const float* colsum;          // [N] sum_k W'[n,k] (of the bf16-rounded folded weight) or nullptr
  int mn_major;                 // 1: both operands are stored reduction-dim-major (A = dY[k, m], B = X[k, n]: the weight
                                // gradient dW = dY^T X straight from the row-major activations, no transposed copies)
};

// gelu(x) = x * Phi(x) with the exact (erf) definition, nn.GELU() default.
// Evaluated as relu(x) - 0.5*|x|*erfc(|x|/sqrt2) with erfc(t/sqrt2) = 2^(t*p(t)); p = degree-4 minimax fit,
// max abs error of the whole expression 5.6e-7 over all x (fit script: oracle/fit_gelu_poly.py).
MT_DEVINL float gelu_erf(float x) {
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

// d/dx gelu(x) = Phi(x) + x*phi(x), same erfc fit as gelu_erf (backward of the MLP activation)
MT_DEVINL float gelu_erf_grad(float x) {
  const float t = fminf(fabsf(x), 12.0f);
  float p = -0.00048811757005751133f;
  p = fmaf(p, t, 0.007198806386440992f);
  p = fmaf(p, t, -0.052146803587675095f);
  p = fmaf(p, t, -0.4595957100391388f);
  p = fmaf(p, t, -1.1510006189346313f);
  float e, g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));                      // erfc(|x|/sqrt2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(t * t * -0.7213475204444817f)); // exp(-x^2/2)
  const float cdf = (x >= 0.f) ? fmaf(e, -0.5f, 1.0f) : 0.5f * e;
  return fmaf(x * 0.3989422804014327f, g, cdf);
}

// two elements at a time with packed fp32x2 FMAs (the epilogue is issue-bound: ~1/3 fewer instruction slots)
MT_DEVINL uint64_t gelu_erf_x2(uint64_t x) {
  float x0, x1;
  unpack_f32x2(x, x0, x1);
  const uint64_t t = pack_f32x2(fminf(fabsf(x0), 12.0f), fminf(fabsf(x1), 12.0f));
  uint64_t p = pack_f32x2(-0.00048811757005751133f, -0.00048811757005751133f);
  p = fma_f32x2(p, t, pack_f32x2(0.007198806386440992f, 0.007198806386440992f));
  p = fma_f32x2(p, t, pack_f32x2(-0.052146803587675095f, -0.052146803587675095f));
  p = fma_f32x2(p, t, pack_f32x2(-0.4595957100391388f, -0.4595957100391388f));
  p = fma_f32x2(p, t, pack_f32x2(-1.1510006189346313f, -1.1510006189346313f));
  float w0, w1;
  unpack_f32x2(mul_f32x2(p, t), w0, w1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
  const uint64_t te = mul_f32x2(t, pack_f32x2(e0, e1));
  return fma_f32x2(te, pack_f32x2(-0.5f, -0.5f), pack_f32x2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}

template <int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_y, const GemmParams p) {
  using Cfg = GemmCfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  const int cluster_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_clusters = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;

  const uint32_t bar_base = smem_base + Cfg::kSmemBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + kAccStages + s); };
  const uint32_t tmem_ptr_addr = smem_base + Cfg::kSmemTmemPtrOff;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kSmemTmemPtrOff);

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);    // one arrive.expect_tx (leader CTA) + the TMA bytes of both CTAs
      mbar_init(empty_bar(s), 1);   // one tcgen05.commit per use
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(tfull_bar(s), 1);                 // tcgen05.commit
      mbar_init(tempty_bar(s), CG * kEpiWarps);   // one arrive per epilogue warp (both CTAs)
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 2) tmem_alloc<CG>(tmem_ptr_addr, kTmemCols);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  pdl_launch_dependents();
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;   // work items: (tile, K split), split fastest
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader_base = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int tile = t / p.split_k, split = t - tile * p.split_k;
        const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
        const int row_a = m_tile * (kBM * CG) + int(rank) * kBM;
        const int row_b = n_tile * kBN + ((CG == 2) ? int(rank) * (kBN / 2) : 0);
        const int kb_begin = split * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (p.mn_major) {
            // MN-major operand tiles: [64 reduction rows x 64 columns] boxes (128 B swizzle rows), 8 KB each; the
            // 128-wide A / B slices of this CTA are two such boxes 8 KB apart (the descriptor's leading byte offset)
            if constexpr (CG == 2) {
              const uint32_t fb = full_leader_base + 8u * stage;
              if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * Cfg::kStageBytes);
#pragma unroll
              for (int blk = 0; blk < kBM / 64; ++blk)
                tma_load_2d_cg2(sa + blk * 8192, &tm_a, fb, row_a + 64 * blk, kb * kBK, kEvictNormal);
#pragma unroll
              for (int blk = 0; blk < Cfg::kBRows / 64; ++blk)
                tma_load_2d_cg2(sb + blk * 8192, &tm_b, fb, row_b + 64 * blk, kb * kBK, kEvictNormal);
            } else {
              mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
              for (int blk = 0; blk < kBM / 64; ++blk)
                tma_load_2d(sa + blk * 8192, &tm_a, full_bar(stage), row_a + 64 * blk, kb * kBK, kEvictNormal);
#pragma unroll
              for (int blk = 0; blk < Cfg::kBRows / 64; ++blk)
                tma_load_2d(sb + blk * 8192, &tm_b, full_bar(stage), row_b + 64 * blk, kb * kBK, kEvictNormal);
            }
          } else if constexpr (CG == 2) {
            const uint32_t fb = full_leader_base + 8u * stage;
            // Only the leader arms its barrier, for the bytes of BOTH CTAs.  The peer's complete_tx may land before
            // the leader's expect_tx of the same phase (tx-count goes transiently negative, which is legal); it can
            // never land in a later phase because the peer only re-uses a slot after the leader's MMA consumed it.
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * Cfg::kStageBytes);
            tma_load_2d_cg2(sa, &tm_a, fb, kb * kBK, row_a, kEvictNormal);
            tma_load_2d_cg2(sb, &tm_b, fb, kb * kBK, row_b, kEvictLast);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, &tm_a, full_bar(stage), kb * kBK, row_a, kEvictNormal);
            tma_load_2d(sb, &tm_b, full_bar(stage), kb * kBK, row_b, kEvictLast);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (leader && lane == 0) {
      const uint32_t idesc = p.mn_major ? make_idesc_bf16(kBM * CG, kBN, 1, 1) : make_idesc_bf16(kBM * CG, kBN);
      // K step of 16: K-major = 32 B inside the swizzle row; MN-major = 16 rows of 128 B
      const uint64_t kstep = p.mn_major ? uint64_t(2048 >> 4) : uint64_t(2);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * kBN);
        const int split = t % p.split_k;
        const int kb_begin = split * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = p.mn_major ? make_smem_desc_sw128(sa, 1024, 8192) : make_smem_desc_sw128(sa);
          const uint64_t bdesc = p.mn_major ? make_smem_desc_sw128(sa + Cfg::kABytes, 1024, 8192)
                                            : make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 along K: +2 (32 B inside the 128B swizzle row) or +128 (16 rows) in the (addr >> 4) field
            umma_f16_ss<CG>(d_tmem, adesc + kstep * uint64_t(k), bdesc + kstep * uint64_t(k), idesc,
                            (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2_mc(empty_bar(stage), 0x3);
          else umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (CG == 2) umma_commit_cg2_mc(tfull_bar(as), 0x3);
        else umma_commit(tfull_bar(as));
        if (++as == kAccStages) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===================================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int hf = (warp - 4) >> 2;  // column half of the 256-wide tile
    const int ew = warp - 4;
    const uint32_t stg = smem_base + Cfg::kSmemEpiOff + ew * kEpiStageBytes;
    float* bias_s = reinterpret_cast<float*>(smem_gen + Cfg::kSmemBiasOff + ew * kBiasBytes);
    int as = 0;
    uint32_t aphase = 0;
    bool store_pending = false;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int tile = t / p.split_k;
      const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
      const int row0 = m_tile * (kBM * CG) + int(rank) * kBM + 32 * q;  // first row of this warp
      const int col0 = n_tile * kBN + 128 * hf;                          // first column of this warp
      const int row = row0 + lane;
      // bias (and LN-fold column-sum) slices for this warp's 128 columns -> private smem (float4 per lane)
      float ln_rstd = 1.f, ln_nmr = 0.f;  // rstd and -mean*rstd of this thread's row
      {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = col0 + 4 * lane;
        if (p.bias != nullptr && c < p.N) b4 = *reinterpret_cast<const float4*>(p.bias + c);
        if (p.colsum != nullptr && c < p.N) c4 = *reinterpret_cast<const float4*>(p.colsum + c);
        __syncwarp();
        reinterpret_cast<float4*>(bias_s)[lane] = b4;
        reinterpret_cast<float4*>(bias_s)[32 + lane] = c4;
        __syncwarp();
        if (p.row_stats != nullptr && row < p.M) {
          float2 st;
          if (p.ln_sums > 0) {   // partial (sum, sum of squares) per 128-column slice, fixed summation order
            const float2* ps = p.row_stats + size_t(row) * p.ln_sums;
            float sx = 0.f, sy = 0.f;
            for (int i = 0; i < p.ln_sums; ++i) { const float2 t = ps[i]; sx += t.x; sy += t.y; }
            const float inv_k = 1.0f / float(p.K);
            const float mean = sx * inv_k;
            st = make_float2(mean, rsqrtf(fmaxf(fmaf(sy, inv_k, -mean * mean), 0.f) + p.ln_eps));
          } else {
            st = p.row_stats[row];
          }
          ln_rstd = st.y;
          ln_nmr = -st.x * st.y;
        }
      }
      const uint64_t ln_rstd2 = pack_f32x2(ln_rstd, ln_rstd), ln_nmr2 = pack_f32x2(ln_nmr, ln_nmr);
      uint64_t rs2 = pack_f32x2(0.f, 0.f), rss2 = rs2;   // running (sum, sum of squares) of this thread's output row
      // residual slices are fetched one 32-column slice ahead of their use (the first one before the accumulator
      // is even ready), so the L2/HBM latency hides behind the MMAs of this tile and the math of the previous slice
      const bool row_ok = (p.resid != nullptr) && (row < p.M);
      const __nv_bfloat16* rrow =
          row_ok ? p.resid + size_t(p.resid_period > 0 ? row % p.resid_period : row) * p.ldr : nullptr;
      uint4 rv_next[4];
      auto load_resid = [&](int sc_, uint4 (&dst)[4]) {
        const int c_ = col0 + 32 * sc_;
        if (row_ok && c_ < p.N) {
          if (p.resid_ld256) {
            // two 32-byte loads: whole sectors per thread (rows are scattered, 16-byte loads would touch each sector twice)
#pragma unroll
            for (int j = 0; j < 2; ++j)
              asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(dst[2 * j].x), "=r"(dst[2 * j].y), "=r"(dst[2 * j].z), "=r"(dst[2 * j].w),
                             "=r"(dst[2 * j + 1].x), "=r"(dst[2 * j + 1].y), "=r"(dst[2 * j + 1].z), "=r"(dst[2 * j + 1].w)
                           : "l"(rrow + c_ + 16 * j));
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(rrow + c_);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = __ldg(rp + j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(0u, 0u, 0u, 0u);
        }
      };
      load_resid(0, rv_next);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(32 * q) << 16) + uint32_t(as * kBN + 128 * hf);
#pragma unroll 1
      for (int sc = 0; sc < 4; ++sc) {  // four 32-column slices; two slices = one 64-column TMA store
        const int c_sc = col0 + 32 * sc;
        uint4 rv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rv[j] = rv_next[j];
        if (sc < 3) load_resid(sc + 1, rv_next);
        uint32_t acc[32];
        tmem_ld_32x32b_x32(t_row + uint32_t(32 * sc), acc);
        tmem_wait_ld();
        if (sc == 3) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_remote(tempty_bar(as), 0);
            else mbar_arrive(tempty_bar(as));
          }
        }
        if (p.out_f32 != nullptr) {
          // split-K accumulation: add this partial slice straight from registers (16-byte vector reductions)
          if (row < p.M) {
            float* dst = p.out_f32 + size_t(row) * p.N + c_sc;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              if (c_sc + e < p.N)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(acc[e])),
                             "f"(__uint_as_float(acc[e + 1])), "f"(__uint_as_float(acc[e + 2])),
                             "f"(__uint_as_float(acc[e + 3])) : "memory");
            }
          }
          continue;
        }
        if ((sc & 1) == 0 && store_pending) {
          // staging buffer is about to be overwritten: previous TMA store must have read it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          store_pending = false;
        }
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = reinterpret_cast<const float4*>(bias_s)[8 * sc + j];
          const float4 c4 = reinterpret_cast<const float4*>(bias_s)[32 + 8 * sc + j];
          // rstd*(acc - mean*colsum) + bias; without LN folding rstd = 1 and colsum = 0: plain acc + bias
          uint64_t v01 = fma_f32x2(pack_f32x2(__uint_as_float(acc[4 * j + 0]), __uint_as_float(acc[4 * j + 1])), ln_rstd2,
                                   fma_f32x2(ln_nmr2, pack_f32x2(c4.x, c4.y), pack_f32x2(b4.x, b4.y)));
          uint64_t v23 = fma_f32x2(pack_f32x2(__uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])), ln_rstd2,
                                   fma_f32x2(ln_nmr2, pack_f32x2(c4.z, c4.w), pack_f32x2(b4.z, b4.w)));
          if (p.act == 1) {
            v01 = gelu_erf_x2(v01);
            v23 = gelu_erf_x2(v23);
          }
          const uint32_t r01 = (j & 1) ? rv[j >> 1].z : rv[j >> 1].x;  // columns 4j, 4j+1
          const uint32_t r23 = (j & 1) ? rv[j >> 1].w : rv[j >> 1].y;  // columns 4j+2, 4j+3
          if (p.act == 2) {
            // backward of fc1's GELU: the "residual" operand holds the saved pre-activation u; y = acc * gelu'(u)
            v01 = mul_f32x2(v01, pack_f32x2(gelu_erf_grad(bf16_lo(r01)), gelu_erf_grad(bf16_hi(r01))));
            v23 = mul_f32x2(v23, pack_f32x2(gelu_erf_grad(bf16_lo(r23)), gelu_erf_grad(bf16_hi(r23))));
          } else {
            v01 = add_f32x2(v01, pack_f32x2(bf16_lo(r01), bf16_hi(r01)));
            v23 = add_f32x2(v23, pack_f32x2(bf16_lo(r23), bf16_hi(r23)));
          }
          if (p.out_row_sums != nullptr) {
            // LayerNorm statistics of the NEXT layer's input, gathered while the row is in registers (fp32 values
            // before the bf16 rounding: the difference is O(2^-9/sqrt(d)) on the mean)
            rs2 = add_f32x2(rs2, add_f32x2(v01, v23));
            rss2 = fma_f32x2(v01, v01, fma_f32x2(v23, v23, rss2));
          }
          float v0, v1, v2, v3;
          unpack_f32x2(v01, v0, v1);
          unpack_f32x2(v23, v2, v3);
          packed[2 * j + 0] = pack_bf16x2(v0, v1);
          packed[2 * j + 1] = pack_bf16x2(v2, v3);
        }
        // swizzled (128B pattern) store into the staging tile: row = lane, 16B chunk index ^ (row & 7)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = (sc & 1) * 4 + j;
          const uint32_t addr = stg + uint32_t(lane * 128) + uint32_t(((chunk ^ (lane & 7)) << 4));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * j + 0]),
                       "r"(packed[4 * j + 1]), "r"(packed[4 * j + 2]), "r"(packed[4 * j + 3])
                       : "memory");
        }
        if (sc & 1) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M && (c_sc - 32) < p.N) {
            tma_store_2d(&tm_y, stg, c_sc - 32, row0);
            tma_store_commit();
          }
          store_pending = true;
        }
      }
      if (p.out_row_sums != nullptr && row < p.M) {
        float a0, a1, q0, q1;
        unpack_f32x2(rs2, a0, a1);
        unpack_f32x2(rss2, q0, q1);
        float2* dst = reinterpret_cast<float2*>(p.out_row_sums) + size_t(row) * (2 * p.num_n_tiles) + (2 * n_tile + hf);
        *dst = make_float2(a0 + a1, q0 + q1);
      }
      if (++as == kAccStages) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  // ---- teardown ----
  __syncwarp();  // re-converge the single-lane role loops before the aligned barrier
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int CG>
static int launch_gemm(const void* x, const void* w, const float* bias, const void* resid, int ldr, void* y, int M,
                       int N, int K, int act, const float* row_stats, const float* colsum, int ln_sums, float ln_eps,
                       float* out_row_sums, cudaStream_t stream, float* out_f32 = nullptr, int split_k = 1,
                       int resid_period = 0, int mn_major = 0) {
  using Cfg = GemmCfg<CG>;
  CUtensorMap tm_a, tm_b, tm_y;
  if (mn_major) {
    // x = dY [K tokens, M], w = X [K tokens, N], both row-major: 64 x 64 boxes with the output dimension innermost
    uint32_t box[2] = {64u, uint32_t(kBK)};
    uint64_t dims_a[2] = {uint64_t(M), uint64_t(K)};
    uint64_t str_a[1] = {uint64_t(M) * 2};
    int rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims_a, str_a, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t dims_b[2] = {uint64_t(N), uint64_t(K)};
    uint64_t str_b[1] = {uint64_t(N) * 2};
    rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims_b, str_b, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    {
      uint64_t dims[2] = {uint64_t(K), uint64_t(M)};
      uint64_t strides[1] = {uint64_t(K) * 2};
      uint32_t box[2] = {uint32_t(kBK), uint32_t(kBM)};
      int rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    {
      uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
      uint64_t strides[1] = {uint64_t(K) * 2};
      uint32_t box[2] = {uint32_t(kBK), uint32_t(Cfg::kBRows)};
      int rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  {
    // (in split-K mode the bf16 output map is unused; it is encoded over the X operand only to keep the launch uniform)
    const uint64_t inner = out_f32 ? uint64_t(mn_major ? M : K) : uint64_t(N);
    const uint64_t outer = out_f32 ? uint64_t(mn_major ? K : M) : uint64_t(M);
    uint64_t dims[2] = {inner, outer};
    uint64_t strides[1] = {inner * 2};
    uint32_t box[2] = {64u, 32u};
    int rc = encode_tmap(&tm_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_f32 ? x : y, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  GemmParams p;
  p.bias = bias;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(resid);
  p.ldr = ldr;
  p.resid_period = resid_period;
  {
    static int ld256 = -1;   // MT_B200_RESID256=0 falls back to 16-byte loads (A/B runs)
    if (ld256 < 0) {
      const char* e = getenv("MT_B200_RESID256");
      ld256 = (e && e[0] == '0') ? 0 : 1;
    }
    p.resid_ld256 = (ld256 && resid != nullptr && (reinterpret_cast<uintptr_t>(resid) & 31) == 0 && (ldr % 16) == 0) ? 1 : 0;
  }
  p.mn_major = mn_major;
  p.M = M; p.N = N; p.K = K;
  p.act = act;
  p.row_stats = reinterpret_cast<const float2*>(row_stats);
  p.colsum = colsum;
  p.ln_sums = ln_sums;
  p.ln_eps = ln_eps;
  p.out_row_sums = out_row_sums;
  p.out_f32 = out_f32;
  {
    const int num_kb = ceil_div(K, kBK);
    int sk = split_k < 1 ? 1 : (split_k > num_kb ? num_kb : split_k);
    p.kb_per_split = ceil_div(num_kb, sk);
    p.split_k = ceil_div(num_kb, p.kb_per_split);   // every split non-empty
  }
  p.num_m_tiles = ceil_div(M, kBM * CG);
  p.num_n_tiles = ceil_div(N, kBN);

  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
  int clusters = sms / CG;
  if (clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;

  auto kern = gemm_bf16_kernel<CG>;
  {
    // the opt-in to > 48 KB of dynamic shared memory is per device (nn.DataParallel drives several from one process)
    static bool attr_set[64] = {false};
    int dev = 0;
    MT_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      MT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  MT_CHECK_CUDA(launch_pdl(kern, dim3(unsigned(clusters * CG), 1, 1), dim3(kGemmThreads, 1, 1), size_t(Cfg::kSmemBytes),
                           stream, unsigned(CG), tm_a, tm_b, tm_y, p));
  return MT_OK;
}

}  // namespace mt

static int gemm_entry(const char* who, const void* x, const void* w, const float* bias, const void* residual, void* y,
                      int M, int N, int K, int act, const float* row_stats, const float* colsum, int cta_group,
                      void* stream, int ln_sums = 0, float ln_eps = 0.f, float* out_row_sums = nullptr,
                      int resid_period = 0) {
  using namespace mt;
  MT_REQUIRE(x && w && y, MT_ERR_BAD_SHAPE, "%s: null pointer", who);
  MT_REQUIRE(M > 0 && N > 0 && K > 0, MT_ERR_BAD_SHAPE, "%s: M,N,K must be positive (got %d,%d,%d)", who, M, N, K);
  MT_REQUIRE(K % 8 == 0 && N % 8 == 0, MT_ERR_BAD_SHAPE,
             "%s: K and N must be multiples of 8 (16-byte rows for TMA); got K=%d N=%d", who, K, N);
  MT_REQUIRE(N % 32 == 0, MT_ERR_BAD_SHAPE, "%s: N must be a multiple of 32 (got %d)", who, N);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(colsum) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(row_stats) & 7) == 0,
             MT_ERR_BAD_ALIGN, "%s: all pointers must be 16-byte aligned", who);
  MT_REQUIRE(act == 0 || act == 1 || act == 2, MT_ERR_UNSUPPORTED,
             "%s: act must be 0 (none), 1 (gelu) or 2 (times gelu'(residual))", who);
  MT_REQUIRE(act != 2 || residual != nullptr, MT_ERR_BAD_SHAPE, "%s: act 2 needs the pre-activation as `residual`", who);
  MT_REQUIRE(resid_period >= 0 && (resid_period == 0 || (residual != nullptr && act != 2)), MT_ERR_BAD_SHAPE,
             "%s: residual_period needs a residual table and act != 2", who);
  MT_REQUIRE((row_stats == nullptr) == (colsum == nullptr), MT_ERR_BAD_SHAPE,
             "%s: row_stats and colsum must be given together", who);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cta_group == 0) {
    // auto: CTA pairs unless overridden (MT_B200_GEMM_CG=1|2, read once; for A/B experiments only)
    static int auto_cg = 0;
    if (auto_cg == 0) {
      const char* e = getenv("MT_B200_GEMM_CG");
      auto_cg = (e && e[0] == '1') ? 1 : 2;
    }
    cta_group = auto_cg;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(out_row_sums) & 7) == 0, MT_ERR_BAD_ALIGN, "%s: out_row_sums must be 8-byte aligned", who);
  if (cta_group == 1)
    return launch_gemm<1>(x, w, bias, residual, N, y, M, N, K, act, row_stats, colsum, ln_sums, ln_eps, out_row_sums, s,
                          nullptr, 1, resid_period);
  if (cta_group == 2)
    return launch_gemm<2>(x, w, bias, residual, N, y, M, N, K, act, row_stats, colsum, ln_sums, ln_eps, out_row_sums, s,
                          nullptr, 1, resid_period);
  set_error("%s: cta_group must be 0 (auto), 1 or 2", who);
  return MT_ERR_UNSUPPORTED;
}

extern "C" int mt_gemm_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y, int M,
                            int N, int K, int act, int cta_group, void* stream) {
  return gemm_entry("mt_gemm_bf16", x, w, bias, residual, y, M, N, K, act, nullptr, nullptr, cta_group, stream);
}

extern "C" int mt_gemm_bf16_ln(const void* x, const void* w_folded, const float* bias_folded, const float* colsum,
                               const float* row_stats, void* y, int M, int N, int K, int act, int cta_group,
                               void* stream) {
  if (!row_stats || !colsum) {
    mt::set_error("mt_gemm_bf16_ln: row_stats and colsum are required");
    return MT_ERR_BAD_SHAPE;
  }
  return gemm_entry("mt_gemm_bf16_ln", x, w_folded, bias_folded, nullptr, y, M, N, K, act, row_stats, colsum, cta_group,
                    stream);
}

extern "C" int mt_gemm_bf16_ex(const void* x, const void* w, void* y, int M, int N, int K, const mt_gemm_opts* o,
                               void* stream) {
  if (!o) {
    mt::set_error("mt_gemm_bf16_ex: opts is required");
    return MT_ERR_BAD_SHAPE;
  }
  return gemm_entry("mt_gemm_bf16_ex", x, w, o->bias, o->residual, y, M, N, K, o->act, o->ln_row_stats, o->ln_colsum,
                    o->cta_group, stream, o->ln_stats_are_sums, o->ln_eps, o->out_row_sums, o->residual_period);
}

// Number of K splits for the fp32-accumulating (weight-gradient) GEMM: (output tiles x splits) should fill whole waves of
// CTA pairs — e.g. 9 tiles (768 x 768) x 8 = 72 items on 74 pairs, 36 tiles x 4 = 144 items = 2 waves at 97 % — while
// keeping every split at least 8 k-blocks long and the number of atomic passes over the output low.
static int auto_splits(int tiles, int num_kb, int sms) {
  const int pairs = sms / 2 > 0 ? sms / 2 : 1;
  int best = 1;
  double best_score = -1.0;
  for (int sp = 1; sp <= 32; ++sp) {
    if (sp > 1 && num_kb / sp < 8) break;
    const int items = tiles * sp;
    const int waves = (items + pairs - 1) / pairs;
    double score = double(items) / double(waves * pairs);     // wave efficiency
    if (waves < 2) score *= 0.97;                               // a single wave cannot hide its own prologue
    score -= 0.004 * sp;                                        // every split is one more atomic pass over dW
    if (score > best_score) { best_score = score; best = sp; }
  }
  return best;
}

// y_f32[M,N] += x[M,K] @ w[N,K]^T with the K range split over `splits` CTAs per output tile (fp32 atomics).
// Meant for weight gradients dW = dY^T X (x = dY^T, w = X^T): a handful of output tiles and K = number of tokens.
extern "C" int mt_gemm_bf16_splitk(const void* x, const void* w, float* y_f32, int M, int N, int K, int splits,
                                   void* stream) {
  using namespace mt;
  MT_REQUIRE(x && w && y_f32, MT_ERR_BAD_SHAPE, "mt_gemm_bf16_splitk: null pointer");
  MT_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0 && N % 4 == 0, MT_ERR_BAD_SHAPE,
             "mt_gemm_bf16_splitk: need M,N,K > 0, K %% 8 == 0, N %% 4 == 0 (got %d,%d,%d)", M, N, K);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y_f32)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_gemm_bf16_splitk: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (splits <= 0) {
    // auto: enough (tile, split) work items for two waves of CTA pairs
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    splits = auto_splits(ceil_div(M, 2 * kBM) * ceil_div(N, kBN), ceil_div(K, kBK), sms);
  }
  MT_CHECK_CUDA(cudaMemsetAsync(y_f32, 0, size_t(M) * N * sizeof(float), s));
  return launch_gemm<2>(x, w, nullptr, nullptr, N, nullptr, M, N, K, 0, nullptr, nullptr, 0, 0.f, nullptr, s, y_f32, splits);
}

// dW_f32[N_out, K_in] = dY[T, N_out]^T @ X[T, K_in]: the weight gradient read straight from the two row-major activation
// matrices (MN-major tcgen05 operands: no transposed copies), reduction over the T tokens split over CTA pairs.
extern "C" int mt_gemm_bf16_wgrad(const void* dy, const void* x, float* dw_f32, int T, int N_out, int K_in, int splits,
                                  void* stream) {
  using namespace mt;
  MT_REQUIRE(dy && x && dw_f32, MT_ERR_BAD_SHAPE, "mt_gemm_bf16_wgrad: null pointer");
  MT_REQUIRE(T > 0 && N_out > 0 && K_in > 0 && N_out % 8 == 0 && K_in % 8 == 0, MT_ERR_BAD_SHAPE,
             "mt_gemm_bf16_wgrad: need T > 0 and N_out, K_in positive multiples of 8 (got %d,%d,%d)", T, N_out, K_in);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dw_f32)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_gemm_bf16_wgrad: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (splits <= 0) {
    int sms = 0;
    int rc = sm_count(&sms);
    if (rc) return rc;
    splits = auto_splits(ceil_div(N_out, 2 * kBM) * ceil_div(K_in, kBN), ceil_div(T, kBK), sms);
  }
  MT_CHECK_CUDA(cudaMemsetAsync(dw_f32, 0, size_t(N_out) * K_in * sizeof(float), s));
  return launch_gemm<2>(dy, x, nullptr, nullptr, K_in, nullptr, N_out, K_in, T, 0, nullptr, nullptr, 0, 0.f, nullptr, s,
                        dw_f32, splits, 0, 1);
}
