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
//   * 16 epilogue warps (four per scheduler: the epilogue is a latency-bound instruction stream that has to finish a
//     128 x 256 tile in less time than the 12 k-blocks of a K = 768 main loop): TMEM -> registers (tcgen05.ld x16) ->
//     bias / LayerNorm fold / exact-erf GELU / residual -> bf16 -> 64B-swizzled smem -> TMA store (coalesced, clips
//     the M/N tails in hardware); compile-time specialisations for the encoder's three hot epilogues;
//   * the residual stream is a split pair hi + lo of bf16 tensors (16 mantissa bits): the residual GEMMs (proj, fc2)
//     read both halves, add in fp32 and write both; `hi` alone is the A operand of the next GEMM.  A bf16-only
//     stream costs 1.8e-2 max-abs error after 12 layers, the split one 3e-3 (tools/precision_model.py).
//
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warp 3 idle,
// warps 4..19 = epilogue (warp%4 selects the TMEM lane quarter, (warp-4)/4 the 64-column quarter of the tile).
// a timed-out barrier wait traps without the device printf (the call sequence costs registers in every role)
#define MT_WAIT_QUIET 1
#ifndef MT_GEMM_TRACE
#define MT_GEMM_TRACE 0   // 1: debugging build that counts the cycles every warp of CTAs 0-3 spends in its barrier waits
#endif
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <stdlib.h>

namespace mt {

constexpr int kBM = 128;         // rows of X per CTA
constexpr int kBN = 256;         // UMMA N
constexpr int kBK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kRoleWarps = 4;    // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle
constexpr int kEpiWarps = 16;    // four per scheduler: the epilogue is a latency-bound instruction stream
constexpr int kGemmThreads = 32 * (kRoleWarps + kEpiWarps);   // 640 (<= 96 registers per thread)
constexpr int kEpiCols = kBN / (kEpiWarps / 4);               // 64 output columns per epilogue warp
constexpr int kPiece = 16;       // columns per tcgen05.ld (keeps the accumulator slice at 16 registers)
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBN;  // 512
constexpr int kBiasBytes = 2 * kEpiCols * 4; // per warp: 64 fp32 bias + 64 fp32 LN-fold column sums

// Epilogue specialisations (compile-time): the hot shapes of the encoder get straight-line code, everything else
// (backward act 2, split-K accumulation, periodic residual tables, odd combinations) runs the generic variant.
enum : int {
  kEpiGeneric = 0,   // run-time flags: bias, residual (bf16, optional period), act 0/1/2, LN fold, row sums, split-K
  kEpiLn = 1,        // LN fold + bias                          (qkv)
  kEpiLnGelu = 2,    // LN fold + bias + exact-erf GELU          (fc1)
  kEpiRes = 3,       // bias + residual (hi [+ lo]) -> hi [+ lo] output + LayerNorm row sums     (proj, fc2)
};

#ifndef MT_GEMM_STAGES_CG2
#define MT_GEMM_STAGES_CG2 5   // operand ring depth of the 2-CTA kernel (experiment knob: tools/build_variant.py)
#endif
template <int CG, bool HILO>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;                       // 16 KB
  static constexpr int kBRows = (CG == 2) ? kBN / 2 : kBN;            // rows of W this CTA loads
  static constexpr int kBBytes = kBRows * kBK * 2;                    // 16 / 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (CG == 2) ? MT_GEMM_STAGES_CG2 : 3;
  // per warp: 32 rows x 32 columns bf16 (one 2 KB TMA store per 32-column slice), or — split stream — 32 rows x 16
  // columns of hi + the same of lo (two 1 KB stores per 16-column piece)
  static constexpr int kEpiStageBytes = 2048;
  static constexpr int kSmemEpiOff = kStages * kStageBytes;
  static constexpr int kSmemBiasOff = kSmemEpiOff + kEpiWarps * kEpiStageBytes;
  static constexpr int kSmemBarOff = kSmemBiasOff + kEpiWarps * kBiasBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kAccStages;
  static constexpr int kSmemTmemPtrOff = kSmemBarOff + kNumBars * 8;
  static constexpr int kSmemBytes = kSmemTmemPtrOff + 16 + 1024;  // +1024: manual 1 KB alignment slack
};

struct GemmParams {
  const float* bias;            // [N] fp32 or nullptr
  const __nv_bfloat16* resid;   // [M, ldr] bf16 (or [resid_period, ldr], see below) or nullptr; rows 32-byte aligned
  const __nv_bfloat16* resid_lo;// low half of a split (hi + lo) residual stream, or nullptr (kEpiRes only)
  int ldr;
  int resid_period;             // > 0: output row m reads residual row m % resid_period (a [n, d] positional table
                                // broadcast over the batch: tokenizer GEMMs)
  int M, N, K;
  int act;                      // 0 = identity, 1 = exact-erf GELU, 2 = multiply by gelu'(resid) (backward)
  int num_m_tiles, num_n_tiles;
  // LayerNorm folded into the GEMM (see mt_fold_layernorm): y = rstd_m * (acc - mean_m * colsum_n) + bias_n
  const float* row_stats;       // [M][2] (mean, rstd) — or [M][ln_sums][4] partial statistics when ln_sums > 0
  int ln_sums;                  // > 0: row_stats holds ln_sums (shift, sum(v - shift), sum((v - shift)^2), -) records per
                                // row, one per 64-column slice, written by the epilogue of the GEMM that produced this
                                // GEMM's input (shifted sums: no cancellation when |mean| >> std)
  float ln_eps;
  // split-K accumulation mode (weight gradients: few output tiles, very long K): work item = (tile, K split); the
  // epilogue adds the fp32 partial tile to out_f32 [M,N] with red.global.add (zero-initialised by the host wrapper)
  float* out_f32;
  int split_k;                  // >= 1
  int kb_per_split;             // k-blocks per split (every split is non-empty)
  float* out_row_sums;          // [M][ceil(N/64)][4] statistics records of this GEMM's output rows, or nullptr (plain
                                // stores in a fixed layout: deterministic, no zero-init)
  int has_out_lo;               // kEpiRes: also store the low half of the output (tm_ylo)
  const float* row_offset;      // kEpiRes: [M] added to every element of output row m (undoes mt_center_split), or nullptr
  const float* colsum;          // [N] sum_k W'[n,k] (of the bf16-rounded folded weight) or nullptr
  int mn_major;                 // 1: both operands are stored reduction-dim-major (A = dY[k, m], B = X[k, n]: the weight
                                // gradient dW = dY^T X straight from the row-major activations, no transposed copies)
  unsigned long long* tl;       // optional timeline slot {first CTA start, last CTA end} in globaltimer ns, or nullptr
  int fp8;                      // operands are e4m3 bytes (kind::f8f6f4, 128 elements of K per swizzle row); generic epilogue
  float acc_scale;              // the accumulator is multiplied by this before the bias (fp8: scale_x * scale_w; else 1)
  long long* trace;             // MT_GEMM_TRACE builds only (tools/gemm_trace.py): per-warp wait-cycle counters
  __nv_bfloat16* y_direct;      // != nullptr: the epilogue writes its bf16 results straight from registers (one 32-byte
  __nv_bfloat16* ylo_direct;    // sector per thread and 16-column piece) instead of staging them in shared memory for a
                                // TMA store: the staging write + TMA read of every tile competes with the operand
                                // traffic of the main loop for the 128 B/clk of the shared-memory crossbar
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

// Two elements per call with packed fp32x2 FMAs, written in n = -min(|x|, 12) so that no sign flip or extra multiply
// is needed:  t p(t) = n q(n) with q(n) = -p(-n);  0.5 erfc(t/sqrt2) = 2^(n q(n) - 1);  gelu = relu(x) + n * that.
// Per element: 2 FMNMX + 3 packed FMA-class slots + 1 MUFU (the epilogue is issue-bound).
MT_DEVINL uint64_t gelu_erf_x2(uint64_t x) {
  float x0, x1;
  unpack_f32x2(x, x0, x1);
  const uint64_t n = pack_f32x2(fmaxf(-fabsf(x0), -12.0f), fmaxf(-fabsf(x1), -12.0f));
  uint64_t q = pack_f32x2(0.00048811757005751133f, 0.00048811757005751133f);
  q = fma_f32x2(q, n, pack_f32x2(0.007198806386440992f, 0.007198806386440992f));
  q = fma_f32x2(q, n, pack_f32x2(0.052146803587675095f, 0.052146803587675095f));
  q = fma_f32x2(q, n, pack_f32x2(-0.4595957100391388f, -0.4595957100391388f));
  q = fma_f32x2(q, n, pack_f32x2(1.1510006189346313f, 1.1510006189346313f));
  float w0, w1;
  unpack_f32x2(fma_f32x2(q, n, pack_f32x2(-1.0f, -1.0f)), w0, w1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
  return fma_f32x2(n, pack_f32x2(e0, e1), pack_f32x2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}

MT_DEVINL uint64_t bf16x2_to_f32x2(uint32_t v) { return pack_f32x2(bf16_lo(v), bf16_hi(v)); }

MT_DEVINL void ldg_nc_v8(const void* ptr, uint32_t (&d)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
               : "l"(ptr));
}

template <int CG, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_ylo,
                 const GemmParams p) {
  constexpr bool kHiLo = (MODE == kEpiRes);
  using Cfg = GemmCfg<CG, kHiLo>;
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
    if constexpr (kHiLo) tma_prefetch_desc(&tm_ylo);
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
  timeline_start(p.tl);

  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;   // work items: (tile, K split), split fastest
#if MT_GEMM_TRACE
  long long wt0 = 0, wt1 = 0, wt2 = 0, wt3 = 0;
  const long long wt_begin = clock64();
#define WT(slot, stmt)                                   \
  do {                                                   \
    const long long a_ = clock64();                      \
    stmt;                                                \
    wt##slot += clock64() - a_;                          \
  } while (0)
#else
#define WT(slot, stmt) do { stmt; } while (0)
#endif
  const int kbk = (MODE == kEpiGeneric && p.fp8) ? 2 * kBK : kBK;      // elements of K per 128-byte swizzle row
  const int num_kb = (p.K + kbk - 1) / kbk;

  // register re-distribution (compiled for 96 / thread): the four role warps need almost nothing, the residual epilogue
  // holds a whole tile's worth of prefetched residual (hi) per thread
  if (warp < kRoleWarps) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader_base = (CG == 2) ? mapa_u32(full_bar(0), 0) : full_bar(0);
      // tools/gemm_trace.py shows the MMA issuer waiting for operands 25-35 % of the time on every layer shape, the same
      // at 74 and at 37 clusters (a per-SM limit) and falling with ring depth (53 % / 41 % / 34 % at 3 / 4 / 5 stages).
      // tools/bench_tma.cu puts a number on it: one SM's TMA unit delivers 37 B/clk of 128-byte rows (870 cycles per
      // 32 KB stage, on 2 SMs and on 148 alike; 46-48 B/clk when half of the boxes arrive by cluster multicast), a
      // 128 x 256 tile per SM needs 64 B/clk.  Two attempts to go around it lost and were removed: prefetching A into
      // L2 16 / 32 k-blocks ahead (operand waits 42 %, +4-7 % time: the prefetches queue in the same TMA unit) and
      // copying A with cp.async from a loader warp (2.3x slower: at most ring-depth x 16 KB in flight per SM).
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int tile = t / p.split_k, split = t - tile * p.split_k;
        const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
        const int row_a = m_tile * (kBM * CG) + int(rank) * kBM;
        const int row_b = n_tile * kBN + ((CG == 2) ? int(rank) * (kBN / 2) : 0);
        const int kb_begin = split * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          WT(0, mbar_wait(empty_bar(stage), phase ^ 1u));
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          if (MODE == kEpiGeneric && p.mn_major) {
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
            tma_load_2d_cg2(sa, &tm_a, fb, kb * kbk, row_a, kEvictNormal);
            tma_load_2d_cg2(sb, &tm_b, fb, kb * kbk, row_b, kEvictLast);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(sa, &tm_a, full_bar(stage), kb * kbk, row_a, kEvictNormal);
            tma_load_2d(sb, &tm_b, full_bar(stage), kb * kbk, row_b, kEvictLast);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    if (leader && lane == 0) {
      const bool mnm = (MODE == kEpiGeneric) && p.mn_major;
      const bool f8 = (MODE == kEpiGeneric) && p.fp8;
      const uint32_t idesc = f8 ? make_idesc_e4m3(kBM * CG, kBN)
                                : (mnm ? make_idesc_bf16(kBM * CG, kBN, 1, 1) : make_idesc_bf16(kBM * CG, kBN));
      // K step of 16: K-major = 32 B inside the swizzle row; MN-major = 16 rows of 128 B
      const uint64_t kstep = mnm ? uint64_t(2048 >> 4) : uint64_t(2);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        WT(0, mbar_wait(tempty_bar(as), aphase ^ 1u));  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * kBN);
        const int split = t % p.split_k;
        const int kb_begin = split * p.kb_per_split, kb_end = min(num_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          WT(1, mbar_wait(full_bar(stage), phase));
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint64_t adesc = mnm ? make_smem_desc_sw128(sa, 1024, 8192) : make_smem_desc_sw128(sa);
          const uint64_t bdesc = mnm ? make_smem_desc_sw128(sa + Cfg::kABytes, 1024, 8192)
                                     : make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 along K: +2 (32 B inside the 128B swizzle row) or +128 (16 rows) in the (addr >> 4) field
            // (fp8: the same four 32-byte steps per 128-byte row, 32 elements each)
            if (f8) umma_f8_ss<CG>(d_tmem, adesc + kstep * uint64_t(k), bdesc + kstep * uint64_t(k), idesc,
                                   (kb > kb_begin || k > 0) ? 1u : 0u);
            else umma_f16_ss<CG>(d_tmem, adesc + kstep * uint64_t(k), bdesc + kstep * uint64_t(k), idesc,
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
  }
  } else {
    // =============================== epilogue ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // 16 warps: warp % 4 = TMEM lane quarter (32 output rows), (warp - 4) / 4 = 64-column quarter of the 256-wide tile.
    // Per tile and warp: four 16-column pieces (tcgen05.ld x16 -> math -> bf16 -> 64B-swizzled staging), two TMA stores
    // of 32 rows x 32 columns.  Everything a piece needs from global memory (residual) is fetched one piece ahead.
    constexpr bool kGen = (MODE == kEpiGeneric);
    const bool do_ln = kGen ? (p.row_stats != nullptr) : (MODE == kEpiLn || MODE == kEpiLnGelu);
    const bool do_gelu = kGen ? (p.act == 1) : (MODE == kEpiLnGelu);
    const bool do_act2 = kGen && (p.act == 2);
    const bool do_resid = kGen ? (p.resid != nullptr) : kHiLo;
    const bool do_sums = (kGen || kHiLo) && (p.out_row_sums != nullptr);
    const bool do_splitk = kGen && (p.out_f32 != nullptr);
    const int q = warp & 3;
    const int ew = warp - kRoleWarps;
    const int cq = ew >> 2;
    const uint32_t stg = smem_base + Cfg::kSmemEpiOff + ew * Cfg::kEpiStageBytes;
    const float4* bias_s4 = reinterpret_cast<const float4*>(smem_gen + Cfg::kSmemBiasOff + ew * kBiasBytes);
    // staging layouts: 64-byte rows with the 64B swizzle (chunk ^ ((row >> 1) & 3)); split stream: 32-byte rows with
    // the 32B swizzle (chunk ^ ((row >> 2) & 1)), hi at +0, lo at +1024
    const uint32_t srow = stg + uint32_t(lane * (kHiLo ? 32 : 64));
    const uint32_t swz = kHiLo ? (uint32_t(lane >> 2) & 1u) : (uint32_t(lane >> 1) & 3u);
    int as = 0;
    uint32_t aphase = 0;
    bool store_pending = false;
    auto load_bias = [&](int t_) {
      float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t_ < num_tiles) {
        const int n_tile_ = (t_ / p.split_k) % p.num_n_tiles;
        const float* src = (lane < 16) ? p.bias : p.colsum;
        const int c = n_tile_ * kBN + kEpiCols * cq + 4 * (lane & 15);
        if (src != nullptr && c < p.N) v4 = __ldg(reinterpret_cast<const float4*>(src + c));
      }
      return v4;
    };
    auto load_stat = [&](int t_) {
      float2 st = make_float2(0.f, 1.f);
      if (do_ln && p.ln_sums == 0 && t_ < num_tiles) {
        const int m_tile_ = (t_ / p.split_k) / p.num_n_tiles;
        const int row_ = m_tile_ * (kBM * CG) + int(rank) * kBM + 32 * q + lane;
        if (row_ < p.M) st = __ldg(reinterpret_cast<const float2*>(p.row_stats) + row_);
      }
      return st;
    };
    float4 bias_pf = load_bias(cluster_id);
    float2 stat_pf = load_stat(cluster_id);
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
#if MT_GEMM_TRACE
      const long long tile_top = clock64();
#endif
      const int tile = t / p.split_k;
      const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
      const int row0 = m_tile * (kBM * CG) + int(rank) * kBM + 32 * q;  // first row of this warp
      const int col0 = n_tile * kBN + kEpiCols * cq;                    // first column of this warp
      const int row = row0 + lane;
      // bias (lanes 0-15) and LN-fold column sums (lanes 16-31) of this warp's 64 columns -> private smem; the values
      // were requested one tile ahead (their L2 latency used to stall every tile's first instruction)
      __syncwarp();
      const_cast<float4*>(bias_s4)[lane] = bias_pf;
      __syncwarp();
      float ln_rstd = 1.f, ln_nmr = 0.f;  // rstd and -mean*rstd of this thread's row
      if (do_ln && row < p.M) {
        float mean, rstd;
        if (p.ln_sums > 0) {
          // Combine the per-slice shifted sums in one streaming pass: every slice is re-referenced to the first slice's
          // shift (d = shift_i - ref: S1 += s1_i + n_i d, S2 += s2_i + 2 d s1_i + n_i d^2), so the loads are independent
          // of each other (all in flight at once) and nothing cancels: ref is within a few sigma of the mean.
          // (mt_ln_stats_finalize does the same once per row; the whole-encoder driver uses that, because every CTA
          // column of a wide GEMM would otherwise re-read all the records of its rows.)
          const float4* ps = reinterpret_cast<const float4*>(p.row_stats) + size_t(row) * p.ln_sums;
          float4 rec[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) rec[i] = (i < p.ln_sums) ? __ldg(ps + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float ref = rec[0].x;
          float S1 = 0.f, S2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (i < p.ln_sums) {
              const float n_i = float(min(64, p.K - 64 * i));
              const float dd = rec[i].x - ref;
              S1 += fmaf(n_i, dd, rec[i].y);
              S2 += fmaf(dd, fmaf(n_i, dd, 2.0f * rec[i].y), rec[i].z);
            }
          }
          const float inv_k = 1.0f / float(p.K);
          const float m1 = S1 * inv_k;
          mean = ref + m1;
          rstd = rsqrtf(fmaxf(fmaf(S2, inv_k, -m1 * m1), 0.f) + p.ln_eps);
        } else {
          mean = stat_pf.x;
          rstd = stat_pf.y;
        }
        ln_rstd = rstd;
        ln_nmr = -mean * rstd;
      }
      // request the next tile's bias / column sums / row statistics now
      bias_pf = load_bias(t + num_clusters);
      stat_pf = load_stat(t + num_clusters);
      const uint64_t ln_rstd2 = pack_f32x2(ln_rstd, ln_rstd), ln_nmr2 = pack_f32x2(ln_nmr, ln_nmr);
      const float acc_scale = kGen ? p.acc_scale : 1.0f;
      const uint64_t acc_scale2 = pack_f32x2(acc_scale, acc_scale);
      uint64_t roff2 = pack_f32x2(0.f, 0.f);
      if constexpr (kHiLo) {
        if (p.row_offset != nullptr && row < p.M) {
          const float c_ = p.row_offset[row];
          roff2 = pack_f32x2(c_, c_);
        }
      }
      // running statistics of this thread's 64 output values (shifted by the first one)
      uint64_t s1 = pack_f32x2(0.f, 0.f), s2 = s1, nshift2 = s1;
      float shift = 0.f;
      // residual pieces (16 columns = 32 bytes of hi, 32 bytes of lo) are fetched one piece ahead of their use: the
      // first one before the accumulator is even ready, so the L2/HBM latency hides behind the MMAs of this tile
      const bool row_ok = do_resid && (row < p.M);
      const __nv_bfloat16* rrow = nullptr;
      const __nv_bfloat16* rrow_lo = nullptr;
      if (row_ok) {
        const size_t rr = size_t((kGen && p.resid_period > 0) ? row % p.resid_period : row) * p.ldr;
        rrow = p.resid + rr;
        if (kHiLo && p.resid_lo != nullptr) rrow_lo = p.resid_lo + rr;
      }
      // kEpiRes: the hi halves of all four pieces are requested up front (64 KB per SM in flight while the main loop of
      // this tile still runs), the lo halves one piece ahead; generic mode: one piece ahead
      constexpr int kRhBufs = kHiLo ? (kEpiCols / kPiece) : 2;
      uint32_t rh[kRhBufs][8], rl[2][8];
      auto load_hi = [&](int piece, uint32_t (&h)[8]) {
        const int c_ = col0 + kPiece * piece;
        if (rrow != nullptr && c_ < p.N) ldg_nc_v8(rrow + c_, h);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = 0u;
        }
      };
      auto load_lo = [&](int piece, uint32_t (&l)[8]) {
        const int c_ = col0 + kPiece * piece;
        if (rrow_lo != nullptr && c_ < p.N) ldg_nc_v8(rrow_lo + c_, l);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) l[j] = 0u;
        }
      };
      if (do_resid) {
        if constexpr (kHiLo) {
#pragma unroll
          for (int pc = 0; pc < kEpiCols / kPiece; ++pc) load_hi(pc, rh[pc]);
          load_lo(0, rl[0]);
          // When the epilogue is the slower side it never idles, so the loads above are issued exactly when their data
          // is needed and the whole HBM latency would be exposed once per tile: pull the NEXT tile's 128-byte row
          // segments (hi and lo) into L2 now, one tile ahead.
          const int t_n = t + num_clusters;
          if (t_n < num_tiles) {
            const int tile_n = t_n / p.split_k;
            const int row_n = (tile_n / p.num_n_tiles) * (kBM * CG) + int(rank) * kBM + 32 * q + lane;
            const int col_n = (tile_n % p.num_n_tiles) * kBN + kEpiCols * cq;
            if (row_n < p.M && col_n < p.N) {
              const size_t off = size_t(row_n) * p.ldr + col_n;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + off));
              if (p.resid_lo != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid_lo + off));
            }
          }
        } else {
          load_hi(0, rh[0]);
        }
      }
#if MT_GEMM_TRACE
      wt2 += clock64() - tile_top;
#endif
      WT(0, mbar_wait(tfull_bar(as), aphase));
#if MT_GEMM_TRACE
      const long long pieces_top = clock64();
#endif
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(32 * q) << 16) + uint32_t(as * kBN + kEpiCols * cq);
#pragma unroll
      for (int piece = 0; piece < kEpiCols / kPiece; ++piece) {
        const int c_p = col0 + kPiece * piece;
        uint32_t acc[16];
        tmem_ld_32x32b_x16(t_row + uint32_t(kPiece * piece), acc);
        if (do_resid && piece + 1 < kEpiCols / kPiece) {
          if constexpr (kHiLo) load_lo(piece + 1, rl[(piece + 1) & 1]);
          else load_hi(piece + 1, rh[(piece + 1) & 1]);
        }
        tmem_wait_ld();
        if (piece == kEpiCols / kPiece - 1) {
          // all TMEM reads of this accumulator stage are done: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_remote(tempty_bar(as), 0);
            else mbar_arrive(tempty_bar(as));
          }
        }
        if (do_splitk) {
          // split-K accumulation: add this partial piece straight from registers (16-byte vector reductions)
          if (row < p.M) {
            float* dst = p.out_f32 + size_t(row) * p.N + c_p;
#pragma unroll
            for (int e = 0; e < kPiece; e += 4) {
              if (c_p + e < p.N)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + e), "f"(__uint_as_float(acc[e])),
                             "f"(__uint_as_float(acc[e + 1])), "f"(__uint_as_float(acc[e + 2])),
                             "f"(__uint_as_float(acc[e + 3])) : "memory");
            }
          }
          continue;
        }
        const uint32_t (&rhc)[8] = rh[kHiLo ? piece : (piece & 1)];
        const uint32_t (&rlc)[8] = rl[piece & 1];
        uint32_t hp[8], lp[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {       // one 16-byte chunk (8 columns) per step
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {   // four columns per step
            const int i = 2 * j + ii;
            const float4 b4 = bias_s4[4 * piece + i];
            uint64_t v01 = pack_f32x2(__uint_as_float(acc[4 * i + 0]), __uint_as_float(acc[4 * i + 1]));
            uint64_t v23 = pack_f32x2(__uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
            if (do_ln) {
              // rstd*(acc - mean*colsum) + bias
              const float4 c4 = bias_s4[16 + 4 * piece + i];
              v01 = fma_f32x2(v01, ln_rstd2, fma_f32x2(ln_nmr2, pack_f32x2(c4.x, c4.y), pack_f32x2(b4.x, b4.y)));
              v23 = fma_f32x2(v23, ln_rstd2, fma_f32x2(ln_nmr2, pack_f32x2(c4.z, c4.w), pack_f32x2(b4.z, b4.w)));
            } else {
              v01 = fma_f32x2(v01, acc_scale2, pack_f32x2(b4.x, b4.y));
              v23 = fma_f32x2(v23, acc_scale2, pack_f32x2(b4.z, b4.w));
            }
            if (do_gelu) {
              v01 = gelu_erf_x2(v01);
              v23 = gelu_erf_x2(v23);
            }
            if (do_act2) {
              // backward of fc1's GELU: the "residual" operand holds the saved pre-activation u; y = acc * gelu'(u)
              const uint32_t r01 = rhc[2 * i], r23 = rhc[2 * i + 1];
              v01 = mul_f32x2(v01, pack_f32x2(gelu_erf_grad(bf16_lo(r01)), gelu_erf_grad(bf16_hi(r01))));
              v23 = mul_f32x2(v23, pack_f32x2(gelu_erf_grad(bf16_lo(r23)), gelu_erf_grad(bf16_hi(r23))));
            } else if (do_resid) {
              v01 = add_f32x2(v01, bf16x2_to_f32x2(rhc[2 * i]));
              v23 = add_f32x2(v23, bf16x2_to_f32x2(rhc[2 * i + 1]));
              if constexpr (kHiLo) {
                v01 = add_f32x2(add_f32x2(v01, bf16x2_to_f32x2(rlc[2 * i])), roff2);
                v23 = add_f32x2(add_f32x2(v23, bf16x2_to_f32x2(rlc[2 * i + 1])), roff2);
              }
            }
            if (do_sums) {
              // LayerNorm statistics of the NEXT layer's input, gathered while the row is in registers (fp32 values
              // before any rounding), as sums of (v - shift) with shift = the row's first value of this 64-column slice
              if (piece == 0 && i == 0) {
                float v0_, v1_;
                unpack_f32x2(v01, v0_, v1_);
                shift = v0_;
                nshift2 = pack_f32x2(-shift, -shift);
              }
              const uint64_t d01 = add_f32x2(v01, nshift2), d23 = add_f32x2(v23, nshift2);
              s1 = add_f32x2(s1, add_f32x2(d01, d23));
              s2 = fma_f32x2(d01, d01, fma_f32x2(d23, d23, s2));
            }
            float v0, v1, v2, v3;
            unpack_f32x2(v01, v0, v1);
            unpack_f32x2(v23, v2, v3);
            const uint32_t h01 = pack_bf16x2(v0, v1), h23 = pack_bf16x2(v2, v3);
            hp[2 * i + 0] = h01;
            hp[2 * i + 1] = h23;
            if constexpr (kHiLo) {
              // low half of the split representation: bf16(v - float(hi)); hi + lo carries 16 mantissa bits
              float l0, l1, l2, l3;
              unpack_f32x2(add_f32x2(v01, pack_f32x2(-bf16_lo(h01), -bf16_hi(h01))), l0, l1);
              unpack_f32x2(add_f32x2(v23, pack_f32x2(-bf16_lo(h23), -bf16_hi(h23))), l2, l3);
              lp[2 * i + 0] = pack_bf16x2(l0, l1);
              lp[2 * i + 1] = pack_bf16x2(l2, l3);
            }
          }
        }
        if (p.y_direct != nullptr) {
          // one full 32-byte sector per thread: columns c_p .. c_p + 15 of this thread's row
          if (row < p.M && c_p < p.N) {
            __nv_bfloat16* dst = p.y_direct + size_t(row) * p.N + c_p;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(hp[0]), "r"(hp[1]),
                         "r"(hp[2]), "r"(hp[3]), "r"(hp[4]), "r"(hp[5]), "r"(hp[6]), "r"(hp[7]) : "memory");
            if constexpr (kHiLo) {
              if (p.ylo_direct != nullptr) {
                __nv_bfloat16* dl = p.ylo_direct + size_t(row) * p.N + c_p;
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dl), "r"(lp[0]), "r"(lp[1]),
                             "r"(lp[2]), "r"(lp[3]), "r"(lp[4]), "r"(lp[5]), "r"(lp[6]), "r"(lp[7]) : "memory");
              }
            }
          }
          continue;
        }
        if constexpr (kHiLo) {
          // split stream: one 32-row x 16-column box of hi and one of lo per piece
          if (store_pending) {
            WT(1, if (lane == 0) tma_store_wait_read<0>(); __syncwarp());
            store_pending = false;
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t addr = srow + ((uint32_t(j) ^ swz) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[4 * j + 0]), "r"(hp[4 * j + 1]),
                         "r"(hp[4 * j + 2]), "r"(hp[4 * j + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 1024u), "r"(lp[4 * j + 0]),
                         "r"(lp[4 * j + 1]), "r"(lp[4 * j + 2]), "r"(lp[4 * j + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < p.M && c_p < p.N) {
            tma_store_2d(&tm_y, stg, c_p, row0);
            if (p.has_out_lo) tma_store_2d(&tm_ylo, stg + 1024u, c_p, row0);
            tma_store_commit();
          }
          store_pending = true;
        } else {
          if ((piece & 1) == 0 && store_pending) {
            // staging buffer is about to be overwritten: the previous TMA store must have read it
            WT(1, if (lane == 0) tma_store_wait_read<0>(); __syncwarp());
            store_pending = false;
          }
          // 64B-swizzled store into the staging tile: row = lane, 32 bytes of this piece = chunks 2*(piece&1), +1
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t chunk = uint32_t((piece & 1) * 2 + j);
            const uint32_t addr = srow + ((chunk ^ swz) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hp[4 * j + 0]), "r"(hp[4 * j + 1]),
                         "r"(hp[4 * j + 2]), "r"(hp[4 * j + 3]) : "memory");
          }
          if (piece & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && row0 < p.M && (c_p - kPiece) < p.N) {
              tma_store_2d(&tm_y, stg, c_p - kPiece, row0);
              tma_store_commit();
            }
            store_pending = true;
          }
        }
      }
      if (do_sums && row < p.M && col0 < p.N) {
        float a0, a1, q0, q1;
        unpack_f32x2(s1, a0, a1);
        unpack_f32x2(s2, q0, q1);
        float4* dst = reinterpret_cast<float4*>(p.out_row_sums) + size_t(row) * ((p.N + 63) >> 6) + (col0 >> 6);
        *dst = make_float4(shift, a0 + a1, q0 + q1, 0.f);
      }
#if MT_GEMM_TRACE
      wt3 += clock64() - pieces_top;
#endif
      if (++as == kAccStages) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

  // ---- teardown ----
#if MT_GEMM_TRACE
  if (p.trace != nullptr && blockIdx.x < 4 && lane == 0) {
    long long* o = p.trace + (size_t(blockIdx.x) * (kGemmThreads / 32) + warp) * 8;
    o[0] = wt0; o[1] = wt1; o[2] = wt2; o[3] = wt3; o[4] = clock64() - wt_begin;
  }
#endif
  __syncwarp();  // re-converge the single-lane role loops before the aligned barrier
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, kTmemCols);
  timeline_end(p.tl);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// MT_GEMM_TRACE builds: device buffer of 4 x 20 x 8 int64 that the next GEMM launch fills (nullptr = off)
static long long* g_gemm_trace = nullptr;
extern "C" int mt_gemm_set_trace(void* buf) {
  g_gemm_trace = reinterpret_cast<long long*>(buf);
  return MT_GEMM_TRACE ? -1 : 1;   // 1: this build has no counters
}

struct GemmCall {
  const void* x; const void* w; const float* bias;
  const void* resid; const void* resid_lo; int ldr;
  void* y; void* y_lo;
  int M, N, K, act;
  const float* row_stats; const float* colsum; int ln_sums; float ln_eps;
  float* out_row_sums;
  float* out_f32; int split_k; int resid_period; int mn_major;
  const float* row_offset;
  int fp8; float acc_scale;
};

template <int CG, int MODE>
static int launch_gemm(const GemmCall& c, cudaStream_t stream) {
  constexpr bool kHiLo = (MODE == kEpiRes);
  using Cfg = GemmCfg<CG, kHiLo>;
  const int M = c.M, N = c.N, K = c.K;
  CUtensorMap tm_a, tm_b, tm_y, tm_ylo;
  if (c.mn_major) {
    // x = dY [K tokens, M], w = X [K tokens, N], both row-major: 64 x 64 boxes with the output dimension innermost
    uint32_t box[2] = {64u, uint32_t(kBK)};
    uint64_t dims_a[2] = {uint64_t(M), uint64_t(K)};
    uint64_t str_a[1] = {uint64_t(M) * 2};
    int rc = encode_tmap(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.x, dims_a, str_a, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint64_t dims_b[2] = {uint64_t(N), uint64_t(K)};
    uint64_t str_b[1] = {uint64_t(N) * 2};
    rc = encode_tmap(&tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.w, dims_b, str_b, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    // K-major operands: 128-byte swizzle rows = 64 bf16 or 128 fp8 elements of K
    const CUtensorMapDataType dt = c.fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const uint64_t esz = c.fp8 ? 1 : 2;
    const uint32_t kbox = c.fp8 ? 2 * kBK : kBK;
    {
      uint64_t dims[2] = {uint64_t(K), uint64_t(M)};
      uint64_t strides[1] = {uint64_t(K) * esz};
      uint32_t box[2] = {kbox, uint32_t(kBM)};
      int rc = encode_tmap(&tm_a, dt, 2, c.x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    {
      uint64_t dims[2] = {uint64_t(K), uint64_t(N)};
      uint64_t strides[1] = {uint64_t(K) * esz};
      uint32_t box[2] = {kbox, uint32_t(Cfg::kBRows)};
      int rc = encode_tmap(&tm_b, dt, 2, c.w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  {
    // output tiles leave as 32-row x 32-column boxes (64-byte rows, 64B swizzle).  In split-K mode the bf16 output map is
    // unused; it is encoded over the X operand only to keep the launch uniform.
    const uint64_t inner = c.out_f32 ? uint64_t(c.mn_major ? M : K) : uint64_t(N);
    const uint64_t outer = c.out_f32 ? uint64_t(c.mn_major ? K : M) : uint64_t(M);
    uint64_t dims[2] = {inner, outer};
    uint64_t strides[1] = {inner * 2};
    uint32_t box[2] = {kHiLo ? 16u : 32u, 32u};
    const CUtensorMapSwizzle sw = kHiLo ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
    int rc = encode_tmap(&tm_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.out_f32 ? c.x : c.y, dims, strides, box, sw);
    if (rc) return rc;
    rc = encode_tmap(&tm_ylo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.out_f32 ? c.x : (c.y_lo ? c.y_lo : c.y), dims, strides,
                     box, sw);
    if (rc) return rc;
  }
  GemmParams p;
  p.bias = c.bias;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(c.resid);
  p.resid_lo = reinterpret_cast<const __nv_bfloat16*>(c.resid_lo);
  p.ldr = c.ldr;
  p.resid_period = c.resid_period;
  p.mn_major = c.mn_major;
  p.M = M; p.N = N; p.K = K;
  p.act = c.act;
  p.row_stats = c.row_stats;
  p.colsum = c.colsum;
  p.ln_sums = c.ln_sums;
  p.ln_eps = c.ln_eps;
  p.out_row_sums = c.out_row_sums;
  p.has_out_lo = c.y_lo != nullptr ? 1 : 0;
  p.row_offset = c.row_offset;
  p.out_f32 = c.out_f32;
  p.fp8 = c.fp8;
  p.acc_scale = c.fp8 ? c.acc_scale : 1.0f;
  p.tl = timeline_next_slot();
  p.trace = g_gemm_trace;
  {
    // Output path of the epilogue: staged in shared memory + TMA store, or 32-byte register stores straight to global
    // memory.  Direct stores lose wherever the epilogue is the slower side (qkv 196 vs 163 us, proj 115 vs 103, fc1 267
    // vs 242) and win on the split-stream residual GEMM with a long main loop, where staging writes + TMA-store reads
    // compete with the operand traffic for the shared-memory pipe (fc2, K = 3072: 212 vs 219 us, same-call A/B) — that
    // case is the default; MT_B200_EPI_DIRECT=1 / 0 forces direct / staged everywhere (A/B runs).
    static int direct = -2;
    if (direct == -2) {
      const char* e = getenv("MT_B200_EPI_DIRECT");
      direct = !e ? -1 : (e[0] == '1' ? 1 : 0);
    }
    const bool want = direct == 1 || (direct == -1 && kHiLo && CG == 2 && K >= 2048);
    const bool ok = want && c.out_f32 == nullptr && (reinterpret_cast<uintptr_t>(c.y) & 31) == 0 &&
                    (reinterpret_cast<uintptr_t>(c.y_lo) & 31) == 0 && (N % 16) == 0;
    p.y_direct = ok ? reinterpret_cast<__nv_bfloat16*>(c.y) : nullptr;
    p.ylo_direct = ok ? reinterpret_cast<__nv_bfloat16*>(c.y_lo) : nullptr;
  }
  {
    const int num_kb = ceil_div(K, c.fp8 ? 2 * kBK : kBK);
    int sk = c.split_k < 1 ? 1 : (c.split_k > num_kb ? num_kb : c.split_k);
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
  {
    // experiment knob (tools/bench_kernels.py): run on fewer CTA clusters, to tell per-SM limits from chip-wide ones
    static int cap = -1;
    if (cap < 0) {
      const char* e = getenv("MT_B200_GEMM_MAX_CLUSTERS");
      cap = (e && atoi(e) > 0) ? atoi(e) : 0;
    }
    if (cap > 0 && clusters > cap) clusters = cap;
  }
  if (clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;

  auto kern = gemm_bf16_kernel<CG, MODE>;
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
                           stream, unsigned(CG), tm_a, tm_b, tm_y, tm_ylo, p));
  return MT_OK;
}

// picks the epilogue specialisation for a call
template <int CG>
static int dispatch_gemm(const GemmCall& c, cudaStream_t s) {
  const bool special_ok = c.out_f32 == nullptr && c.act != 2 && c.resid_period == 0 && !c.mn_major && !c.fp8;
  if (special_ok && c.row_stats != nullptr && c.resid == nullptr && c.out_row_sums == nullptr && c.y_lo == nullptr) {
    if (c.act == 1) return launch_gemm<CG, kEpiLnGelu>(c, s);
    return launch_gemm<CG, kEpiLn>(c, s);
  }
  if (special_ok && c.resid != nullptr && c.act == 0 && c.row_stats == nullptr) return launch_gemm<CG, kEpiRes>(c, s);
  return launch_gemm<CG, kEpiGeneric>(c, s);
}

}  // namespace mt

static int gemm_entry(const char* who, const void* x, const void* w, void* y, int M, int N, int K, const mt_gemm_opts& o,
                      void* stream) {
  using namespace mt;
  MT_REQUIRE(x && w && y, MT_ERR_BAD_SHAPE, "%s: null pointer", who);
  MT_REQUIRE(M > 0 && N > 0 && K > 0, MT_ERR_BAD_SHAPE, "%s: M,N,K must be positive (got %d,%d,%d)", who, M, N, K);
  MT_REQUIRE(K % 8 == 0 && N % 8 == 0, MT_ERR_BAD_SHAPE,
             "%s: K and N must be multiples of 8 (16-byte rows for TMA); got K=%d N=%d", who, K, N);
  MT_REQUIRE(N % 32 == 0, MT_ERR_BAD_SHAPE, "%s: N must be a multiple of 32 (got %d)", who, N);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.y_lo) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(o.bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.ln_colsum) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(o.ln_row_stats) & 7) == 0,
             MT_ERR_BAD_ALIGN, "%s: all pointers must be 16-byte aligned", who);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(o.residual) & 31) == 0 && (reinterpret_cast<uintptr_t>(o.residual_lo) & 31) == 0,
             MT_ERR_BAD_ALIGN, "%s: residual pointers must be 32-byte aligned (rows are read as whole 32-byte sectors)", who);
  MT_REQUIRE(o.act == 0 || o.act == 1 || o.act == 2, MT_ERR_UNSUPPORTED,
             "%s: act must be 0 (none), 1 (gelu) or 2 (times gelu'(residual))", who);
  MT_REQUIRE(o.act != 2 || o.residual != nullptr, MT_ERR_BAD_SHAPE, "%s: act 2 needs the pre-activation as `residual`", who);
  MT_REQUIRE(o.residual_period >= 0 && (o.residual_period == 0 || (o.residual != nullptr && o.act != 2)), MT_ERR_BAD_SHAPE,
             "%s: residual_period needs a residual table and act != 2", who);
  MT_REQUIRE((o.ln_row_stats == nullptr) == (o.ln_colsum == nullptr), MT_ERR_BAD_SHAPE,
             "%s: row_stats and colsum must be given together", who);
  MT_REQUIRE(o.ln_stats_are_sums == 0 || o.ln_stats_are_sums == (K + 63) / 64, MT_ERR_BAD_SHAPE,
             "%s: ln_stats_are_sums must be 0 or ceil(K/64) = %d (got %d)", who, (K + 63) / 64, o.ln_stats_are_sums);
  MT_REQUIRE(o.ln_stats_are_sums == 0 || (reinterpret_cast<uintptr_t>(o.ln_row_stats) & 15) == 0, MT_ERR_BAD_ALIGN,
             "%s: partial-sum statistics must be 16-byte aligned", who);
  MT_REQUIRE((reinterpret_cast<uintptr_t>(o.out_row_sums) & 15) == 0, MT_ERR_BAD_ALIGN,
             "%s: out_row_sums must be 16-byte aligned", who);
  // the split (hi + lo) residual stream exists in the residual epilogue only
  const bool split_ok = o.residual != nullptr && o.act == 0 && o.ln_row_stats == nullptr && o.residual_period == 0;
  MT_REQUIRE((o.residual_lo == nullptr && o.y_lo == nullptr && o.y_row_offset == nullptr) || split_ok, MT_ERR_UNSUPPORTED,
             "%s: residual_lo / y_lo / y_row_offset need a plain residual GEMM (residual given, act 0, no LayerNorm fold, "
             "no period)", who);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int cta_group = o.cta_group;
  if (cta_group == 0) {
    // auto: CTA pairs unless overridden (MT_B200_GEMM_CG=1|2, read once; for A/B experiments only)
    static int auto_cg = 0;
    if (auto_cg == 0) {
      const char* e = getenv("MT_B200_GEMM_CG");
      auto_cg = (e && e[0] == '1') ? 1 : 2;
    }
    cta_group = auto_cg;
  }
  GemmCall c = {};
  c.x = x; c.w = w; c.bias = o.bias; c.resid = o.residual; c.resid_lo = o.residual_lo; c.ldr = N;
  c.y = y; c.y_lo = o.y_lo; c.M = M; c.N = N; c.K = K; c.act = o.act;
  c.row_stats = o.ln_row_stats; c.colsum = o.ln_colsum; c.ln_sums = o.ln_stats_are_sums; c.ln_eps = o.ln_eps;
  c.row_offset = o.y_row_offset;
  c.out_row_sums = o.out_row_sums; c.out_f32 = nullptr; c.split_k = 1; c.resid_period = o.residual_period; c.mn_major = 0;
  if (cta_group == 1) return dispatch_gemm<1>(c, s);
  if (cta_group == 2) return dispatch_gemm<2>(c, s);
  set_error("%s: cta_group must be 0 (auto), 1 or 2", who);
  return MT_ERR_UNSUPPORTED;
}

extern "C" int mt_gemm_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y, int M,
                            int N, int K, int act, int cta_group, void* stream) {
  mt_gemm_opts o = {};
  o.bias = bias; o.residual = residual; o.act = act; o.cta_group = cta_group;
  return gemm_entry("mt_gemm_bf16", x, w, y, M, N, K, o, stream);
}

extern "C" int mt_gemm_bf16_ln(const void* x, const void* w_folded, const float* bias_folded, const float* colsum,
                               const float* row_stats, void* y, int M, int N, int K, int act, int cta_group,
                               void* stream) {
  if (!row_stats || !colsum) {
    mt::set_error("mt_gemm_bf16_ln: row_stats and colsum are required");
    return MT_ERR_BAD_SHAPE;
  }
  mt_gemm_opts o = {};
  o.bias = bias_folded; o.act = act; o.ln_colsum = colsum; o.ln_row_stats = row_stats; o.cta_group = cta_group;
  return gemm_entry("mt_gemm_bf16_ln", x, w_folded, y, M, N, K, o, stream);
}

extern "C" int mt_gemm_bf16_ex(const void* x, const void* w, void* y, int M, int N, int K, const mt_gemm_opts* o,
                               void* stream) {
  if (!o) {
    mt::set_error("mt_gemm_bf16_ex: opts is required");
    return MT_ERR_BAD_SHAPE;
  }
  return gemm_entry("mt_gemm_bf16_ex", x, w, y, M, N, K, *o, stream);
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
  GemmCall c = {};
  c.x = x; c.w = w; c.ldr = N; c.M = M; c.N = N; c.K = K; c.out_f32 = y_f32; c.split_k = splits;
  return launch_gemm<2, kEpiGeneric>(c, s);
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
  GemmCall c = {};
  c.x = dy; c.w = x; c.ldr = K_in; c.M = N_out; c.N = K_in; c.K = T; c.out_f32 = dw_f32; c.split_k = splits; c.mn_major = 1;
  return launch_gemm<2, kEpiGeneric>(c, s);
}

// fp8 (e4m3) opt-in GEMM:  y = act(scale * (xq @ wq^T) + bias) (+ residual), xq / wq = e4m3 bytes
// [M,K] / [N,K] (K-major, K % 16 == 0), scale = scale_x * scale_w of the per-tensor quantisation x ~ scale_x * xq.
// Same pipeline as mt_gemm_bf16 (tcgen05.mma.kind::f8f6f4: twice the K per shared-memory byte, fp32 accumulation in TMEM),
// generic epilogue, bf16 output.
extern "C" int mt_gemm_fp8(const void* xq, const void* wq, const float* bias, const void* residual, void* y, int M, int N, int K,
                           int act, float scale, void* stream) {
  using namespace mt;
  MT_REQUIRE(xq && wq && y, MT_ERR_BAD_SHAPE, "mt_gemm_fp8: null pointer");
  MT_REQUIRE(M > 0 && N > 0 && K > 0 && K % 16 == 0 && N % 32 == 0, MT_ERR_BAD_SHAPE,
             "mt_gemm_fp8: need M, N, K > 0, K %% 16 == 0, N %% 32 == 0 (got %d,%d,%d)", M, N, K);
  MT_REQUIRE(act == 0 || act == 1, MT_ERR_UNSUPPORTED, "mt_gemm_fp8: act must be 0 or 1 (gelu)");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(wq) | reinterpret_cast<uintptr_t>(y) |
               reinterpret_cast<uintptr_t>(bias)) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 31) == 0,
             MT_ERR_BAD_ALIGN, "mt_gemm_fp8: pointers must be 16-byte (residual: 32-byte) aligned");
  GemmCall c = {};
  c.x = xq; c.w = wq; c.bias = bias; c.resid = residual; c.ldr = N; c.y = y; c.M = M; c.N = N; c.K = K; c.act = act;
  c.split_k = 1; c.fp8 = 1; c.acc_scale = scale;
  return launch_gemm<2, kEpiGeneric>(c, reinterpret_cast<cudaStream_t>(stream));
}
