// mt_attention.cu — softmax attention core of timm's Attention.forward (E4 in SURVEY.md §8a):
//   attn = (q @ k^T) * scale; attn = softmax(attn, -1); x = (attn @ v).transpose(1, 2).reshape(B, N, C)
//   (PointCloud/openpoints/models/layers/attention.py:31-35; Image/.../base/vit.py:108-112 is the same math).
// The reference materialises the B*H*N*N score matrix; here S and P never leave the SM.
//
// Blackwell design (sm_100a only), head_dim = 64, no mask, optional attention dropout:
//   * q, k, v are read straight out of the packed (B, N, 3, H, 64) qkv buffer through one 4-D TMA descriptor
//     (no permute / contiguous copies), output is written straight into (B, N, H*64) through a 4-D TMA store;
//   * persistent CTAs, TWO per SM (108 KB of shared memory, 256 TMEM columns, compiled for 80 registers per thread):
//     no unit is saturated at C2 (MUFU 40 %, issue slots 50 %, tensor pipe 23 %: profiles/r2_attention_trace.md) — the
//     kernel is a dependency cycle between the softmax warps and the two single-thread MMA issuers whose steps cost
//     latency, so what counts is how many independent softmax instruction streams an SM scheduler can pick from
//     (4 warps per scheduler) and that none of them waits for the tensor core longer than it must;
//   * a work item is (batch, head, 256 query rows) = two 128-row query tiles that ping-pong; warps whose 32 query rows
//     all lie beyond N only keep the barrier protocol going;
//   * S = Q·K^T and O += P·V are tcgen05.mma (bf16 in, fp32 accumulate) with S and O living in TMEM.  The unit of the
//     softmax pipeline is a HALF key tile of 32 keys: S_t has two 32-column halves (TMEM columns 64 t + 32 h) with their
//     own barriers (both halves of a key tile are produced by one N = 64 MMA group: a single-thread tcgen05.mma costs
//     60-105 cycles to issue whatever N is).  A softmax thread (one query row) pulls its 32 scores with ONE tcgen05.ld,
//     releases the half immediately and does max / exp2 / row sum / bf16 pack from registers — one pass, nothing is
//     read from TMEM twice;
//   * online softmax in fp32 registers with lazy rescaling, exp2 with the scale*log2(e) factor folded in; P goes back
//     as bf16 through the 128B-swizzled P_t tile (half h = 64 bytes of every row), P_t-half·V is issued per half;
//   * K/V tiles (64 keys = two halves) stream through 3- / 2-stage TMA rings shared by both query tiles; the last half
//     of a row of keys runs at N = 16 when it holds at most 16 valid keys (N = 196: 4 keys).
//
// Warps (384 threads): 0-3 softmax tile 0, 4-7 softmax tile 1 (warp%4 = TMEM lane quarter), 8 = TMA loader,
// 9 = S=QK^T issuer + TMEM allocator, 10 = O+=PV issuer, 11 idle (completes the third warpgroup for setmaxnreg).
// a timed-out barrier wait traps without the device printf: the call sequence costs the 40-register role warps spills
#define MT_WAIT_QUIET 1
#ifndef MT_ATTN_TRACE
#define MT_ATTN_TRACE 0   // 1: debugging build that logs (clock64 << 8 | event id) per warp for the first two CTAs
#endif
#include <type_traits>
#include "mt_ptx.cuh"
#include "mt_common.cuh"

namespace mt {

constexpr int kAtQT = 2;                    // query tiles per CTA
constexpr int kAtBM = 128;                  // query rows per tile
constexpr int kAtBN = 64;                   // keys per KV tile (one TMA box)
constexpr int kAtHN = 32;                   // keys per half tile (one softmax step)
constexpr int kAtD = 64;                    // head dim
constexpr int kAtKStages = 3;
constexpr int kAtVStages = 2;
constexpr int kAtThreads = 32 * (4 * kAtQT + 4);
constexpr int kAtQBytes = kAtBM * kAtD * 2;   // 16 KB: a 128 x 64 bf16 tile (Q_t, P_t)
constexpr int kAtKVBytes = kAtBN * kAtD * 2;  //  8 KB: a 64 x 64 bf16 tile (K_j, V_j)
constexpr int kAtSmemQ = 0;
constexpr int kAtSmemP = kAtSmemQ + kAtQT * kAtQBytes;
constexpr int kAtSmemK = kAtSmemP + kAtQT * kAtQBytes;
constexpr int kAtSmemV = kAtSmemK + kAtKStages * kAtKVBytes;
constexpr int kAtSmemBar = kAtSmemV + kAtVStages * kAtKVBytes;
// barriers: q_full[2] q_empty[2] k_full[3] k_empty[3] v_full[2] v_empty[2], then per (tile, half):
// s_full[4] s_empty[4] p_full[4] pv_done[4]
constexpr int kAtNumBars = 2 * kAtQT + 2 * kAtKStages + 2 * kAtVStages + 8 * kAtQT;
constexpr int kAtSmemTmemPtr = kAtSmemBar + kAtNumBars * 8;
constexpr int kAtSmemBytes = kAtSmemTmemPtr + 16 + 1024;   // 107.8 KB: two CTAs per SM
constexpr int kAtTmemCols = 256;   // two CTAs per SM share the 512 columns
constexpr int kAtTmemS = 0;        // S_t half h at columns 64*t + 32*h
constexpr int kAtTmemO = 128;      // O_t at columns 128 + 64*t
static_assert(2 * (kAtSmemBytes + 1024) <= 233472, "two attention CTAs must fit one SM's shared memory");

MT_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  int B, N, H;
  int num_qp;        // ceil(N / 256)
  int num_kv;        // ceil(N / 64)
  int num_half;      // ceil(N / 32)
  int num_items;     // B * H * num_qp
  float scale_log2;  // scale * log2(e)
  float* lse2;       // optional [B, H, N]: log2-sum-exp of the scaled scores (m*c + log2 l), saved for the backward
  unsigned long long* tl;   // optional timeline slot (mt_timeline_begin)
  long long* trace;         // MT_ATTN_TRACE builds only: per-warp event log (tools/attn_trace.py), else unused
  // attention dropout (training, Graph fork): P is multiplied by keep(q, k) before P.V and O by 1 / (1 - p); the row sum
  // l stays the sum of the undropped probabilities (softmax first, dropout second: attention.py:32-33)
  uint32_t drop_thresh;     // p * 2^32, 0 = no dropout
  float drop_scale;         // 1 / (1 - p)
  uint32_t seed_lo, seed_hi;
};

// kDrop: attention dropout compiled in (the keep test sits in the innermost loop: a run-time flag there costs a branch per
// score pair — 13 % of the softmax warps' stall samples were branch resolution in the round-2 capture)
template <bool kDrop>
__global__ void __launch_bounds__(kAtThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                     const __grid_constant__ CUtensorMap tm_out, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar0 = smem_base + kAtSmemBar;
  auto q_full = [&](int t) { return bar0 + 8u * (0 + t); };
  auto q_empty = [&](int t) { return bar0 + 8u * (2 + t); };
  auto k_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (4 + kAtKStages + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (4 + 2 * kAtKStages + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (4 + 2 * kAtKStages + kAtVStages + s); };
  constexpr int kB2 = 4 + 2 * kAtKStages + 2 * kAtVStages;
  // th = 2 * tile + half
  auto s_full = [&](int th) { return bar0 + 8u * (kB2 + th); };
  auto s_empty = [&](int th) { return bar0 + 8u * (kB2 + 4 + th); };
  auto p_full = [&](int th) { return bar0 + 8u * (kB2 + 8 + th); };
  auto pv_done = [&](int th) { return bar0 + 8u * (kB2 + 12 + th); };
  auto q_smem = [&](int t) { return smem_base + uint32_t(kAtSmemQ + t * kAtQBytes); };
  const uint32_t tmem_ptr_addr = smem_base + kAtSmemTmemPtr;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kAtSmemTmemPtr);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_out);
    for (int t = 0; t < kAtQT; ++t) {
      mbar_init(q_full(t), 1);
      mbar_init(q_empty(t), 1);
    }
    for (int th = 0; th < 2 * kAtQT; ++th) {
      mbar_init(s_full(th), 1);
      mbar_init(s_empty(th), 4);  // one arrive per softmax warp
      mbar_init(p_full(th), 4);
      mbar_init(pv_done(th), 1);
    }
    for (int s = 0; s < kAtKStages; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int s = 0; s < kAtVStages; ++s) {
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 9) tmem_alloc<1>(tmem_ptr_addr, kAtTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_launch_dependents();
  pdl_wait();
  timeline_start(p.tl);

  const int H = p.H;
  const int G = p.num_half;
#if MT_ATTN_TRACE == 1
  constexpr int kTrMax = 1024;
  int tr_n = 0;
#define TR(id)                                                                                              \
  do {                                                                                                      \
    if (p.trace != nullptr && blockIdx.x < 2 && lane == 0 && tr_n < kTrMax)                                 \
      p.trace[(size_t(blockIdx.x) * 12 + warp) * kTrMax + tr_n++] = (clock64() << 8) | (long long)(id);     \
  } while (0)
#else
#define TR(id) do { } while (0)
#endif
#if MT_ATTN_TRACE == 2
  // wait-time accounting: cycles each warp spends inside its barrier waits, by kind (slots 0..3), + kernel cycles
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

  // register re-distribution between warpgroups (the code is compiled for 80 registers per thread: two CTAs per SM; the
  // run-time split only returns the role warpgroup's share to the pool): the softmax warpgroups hold a 32-column score
  // row per thread plus the exp / pack temporaries, the loader / MMA warpgroup needs almost nothing
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 8) {
    // =============================== TMA loader =================================
    if (lane == 0) {
      int kst = 0; uint32_t kph = 0;
      int vst = 0; uint32_t vph = 0;
      uint32_t it = 0;
      // K / V of a (batch, head) are read once per 256-query item: keep them in L2 only when there are several items
      const uint64_t kv_hint = (p.num_qp > 1) ? kEvictLast : kEvictFirst;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int qp = item % p.num_qp;
        const int bh = item / p.num_qp;
        const int h = bh % H, b = bh / H;
        for (int t = 0; t < kAtQT; ++t) {
          WT(0, mbar_wait(q_empty(t), (it & 1u) ^ 1u));
          mbar_arrive_expect_tx(q_full(t), kAtQBytes);
          TR(40 + t);
          tma_load_4d(q_smem(t), &tm_q, q_full(t), 0, h, qp * 256 + t * 128, b, kEvictFirst);
        }
        for (int j = 0; j < p.num_kv; ++j) {
          WT(1, mbar_wait(k_empty(kst), kph ^ 1u));
          TR(42);
          mbar_arrive_expect_tx(k_full(kst), kAtKVBytes);
          tma_load_4d(smem_base + kAtSmemK + kst * kAtKVBytes, &tm_kv, k_full(kst), 0, H + h, j * kAtBN, b, kv_hint);
          if (++kst == kAtKStages) { kst = 0; kph ^= 1u; }
          WT(2, mbar_wait(v_empty(vst), vph ^ 1u));
          TR(43);
          mbar_arrive_expect_tx(v_full(vst), kAtKVBytes);
          tma_load_4d(smem_base + kAtSmemV + vst * kAtKVBytes, &tm_kv, v_full(vst), 0, 2 * H + h, j * kAtBN, b,
                      kv_hint);
          if (++vst == kAtVStages) { vst = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == 9) {
    // =============================== S = Q K^T issuer ============================
    // Two single-thread MMA issuers (this warp and warp 10) instead of one: the tcgen05.mma of a work item issued from one
    // thread (descriptor arithmetic + barrier polls are a serial dependency chain) were the critical path.
    // This thread runs ahead on its own, throttled only by s_empty (the softmax warps pulled a half of S_t into registers:
    // the half is then free for the scores two halves further on).
    if (lane == 0) {
      int kst = 0; uint32_t kph = 0;
      uint32_t n_h0 = 0, n_h1 = 0;   // uses of half 0 / half 1 so far (the same for both query tiles)
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        // One MMA group per 64-key tile and query tile (N = the valid keys rounded up to 16): every tcgen05.mma with
        // both operands in shared memory occupies the tensor pipe for at least ~60 cycles however small N is
        // (tools/bench_umma.cu), so the scores of both halves are produced by the same four instructions and only their
        // hand-over to the softmax warps is per half.
        for (int j = 0; j < p.num_kv; ++j) {
          WT(0, mbar_wait(k_full(kst), kph));
          TR(20);
          const uint64_t bdesc = make_smem_desc_sw128(smem_base + kAtSmemK + kst * kAtKVBytes);
          const int valid = min(kAtBN, p.N - j * kAtBN);
          const bool two = valid > kAtHN;                         // the tile has a second half
          const uint32_t idesc = make_idesc_bf16(kAtBM, (valid + 15) & ~15, 0, 0);   // both operands K-major
#pragma unroll
          for (int t = 0; t < kAtQT; ++t) {
            if (j == 0) { WT(1, mbar_wait(q_full(t), it & 1u)); TR(21); }
            WT(2, mbar_wait(s_empty(2 * t + 0), (n_h0 & 1u) ^ 1u));
            if (two) WT(3, mbar_wait(s_empty(2 * t + 1), (n_h1 & 1u) ^ 1u));
            TR(22 + t);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc_sw128(q_smem(t));
            const uint32_t d_tmem = tmem_base + uint32_t(kAtTmemS + t * kAtBN);
            umma_f16_ss<1>(d_tmem, adesc, bdesc, idesc, 0u);
            umma_f16_ss<1>(d_tmem, adesc + 2u, bdesc + 2u, idesc, 1u);
            umma_f16_ss<1>(d_tmem, adesc + 4u, bdesc + 4u, idesc, 1u);
            umma_f16_ss<1>(d_tmem, adesc + 6u, bdesc + 6u, idesc, 1u);
            umma_commit(s_full(2 * t + 0));
            if (two) umma_commit(s_full(2 * t + 1));
            if (j + 1 == p.num_kv) umma_commit(q_empty(t));   // last use of Q_t for this item
          }
          ++n_h0;
          if (two) ++n_h1;
          umma_commit(k_empty(kst));
          if (++kst == kAtKStages) { kst = 0; kph ^= 1u; }
        }
      }
    }
  } else if (warp == 10) {
    // =============================== O += P V issuer =============================
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(kAtBM, kAtD, 0, 1);    // V is consumed MN-major
      int vst = 0; uint32_t vph = 0;
      uint32_t n_h0 = 0, n_h1 = 0;   // P halves consumed so far
      uint64_t bdesc0 = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        for (int g = 0; g < G; ++g) {
          const int hf = g & 1;
          if (hf == 0) {
            WT(0, mbar_wait(v_full(vst), vph));
            bdesc0 = make_smem_desc_sw128(smem_base + kAtSmemV + vst * kAtKVBytes);
          }
          const int valid = min(kAtHN, p.N - g * kAtHN);
          const uint32_t u = hf ? n_h1 : n_h0;
#pragma unroll
          for (int t = 0; t < kAtQT; ++t) {
            WT(1, mbar_wait(p_full(2 * t + hf), u & 1u));
            TR(30 + t);
            tc_fence_after();
            // A: half hf = 64 bytes into P_t's 128-byte rows, 32 bytes per 16-key step;
            // B: 16 keys = 16 rows of 128 B = 2 KB per step (descriptor address field is in 16-byte units)
            const uint64_t adesc0 = make_smem_desc_sw128(smem_base + kAtSmemP + t * kAtQBytes) + uint64_t(4 * hf);
            const uint64_t bdesc = bdesc0 + uint64_t(2 * hf * (16 * 128 >> 4));
            const uint32_t d_tmem = tmem_base + kAtTmemO + t * kAtD;
            umma_f16_ss<1>(d_tmem, adesc0, bdesc, idesc_o, g > 0 ? 1u : 0u);
            if (valid > 16) umma_f16_ss<1>(d_tmem, adesc0 + 2u, bdesc + uint64_t(16 * 128 >> 4), idesc_o, 1u);
            umma_commit(pv_done(2 * t + hf));
          }
          if (hf) ++n_h1; else ++n_h0;
          if (hf == 1 || g + 1 == G) {
            umma_commit(v_empty(vst));
            if (++vst == kAtVStages) { vst = 0; vph ^= 1u; }
          }
        }
      }
    }
  }
  } else {
    // =============================== softmax warps ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const int t = warp >> 2;        // query tile
    const int q = warp & 3;         // TMEM lane quarter
    const int r_in_tile = 32 * q + lane;
    const uint32_t lane_off = uint32_t(32 * q) << 16;
    const uint32_t ts = tmem_base + lane_off + uint32_t(kAtTmemS + t * kAtBN);
    const uint32_t to = tmem_base + lane_off + uint32_t(kAtTmemO + t * kAtD);
    const uint32_t pbase = smem_base + kAtSmemP + t * kAtQBytes;
    const uint32_t prow = pbase + uint32_t(r_in_tile * 128);
    const uint32_t swz = uint32_t(r_in_tile & 7);
    uint32_t use0 = 0, use1 = 0;    // uses of half 0 / half 1 so far
    bool store_pending = false;
    const float c = p.scale_log2;
    const uint64_t c2 = pack_f32x2(c, c);

    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int qp = item % p.num_qp;
      const int bh = item / p.num_qp;
      const int h = bh % H, b = bh / H;
      const int row0 = qp * 256 + t * 128 + 32 * q;     // first query row of this warp
      if (row0 >= p.N) {
        // No valid query row in this warp (N = 196: the last warp of the second tile): keep the barrier protocol going,
        // skip all the arithmetic.  Its rows of P_t stay garbage, which only feeds rows of O_t that are never stored.
        for (int g = 0; g < G; ++g) {
          const int hf = g & 1;
          const uint32_t u = hf ? use1 : use0;
          mbar_wait(s_full(2 * t + hf), u & 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(2 * t + hf));
          mbar_wait(pv_done(2 * t + hf), (u & 1u) ^ 1u);
          if (store_pending) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            store_pending = false;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(2 * t + hf));
          if (hf) ++use1; else ++use0;
        }
        mbar_wait(pv_done(2 * t + 0), (use0 & 1u) ^ 1u);
        if (use1 > 0) mbar_wait(pv_done(2 * t + 1), (use1 & 1u) ^ 1u);
        continue;
      }
      float m_run = -INFINITY;
      float l_run = 0.f;
      const uint32_t bh_seed = attn_drop_seed(p.seed_lo, p.seed_hi, uint32_t(bh));
      const uint32_t drop_row = uint32_t(qp * 256 + t * 128 + r_in_tile) * uint32_t(p.N);   // q * N

      // one half tile: kC = 32 score columns (16 when the half holds at most 16 valid keys)
      auto half_step = [&](auto cols_tag, const int g, const int hf, const int valid, const uint32_t u) {
        constexpr int kC = decltype(cols_tag)::value;
        uint32_t s[kC];
        if constexpr (kC == 32) tmem_ld_32x32b_x32(ts + uint32_t(kAtHN * hf), s);
        else tmem_ld_32x32b_x16(ts + uint32_t(kAtHN * hf), s);
        tmem_wait_ld();
        TR(3);
        // the only read of this half of S_t: the tensor core may overwrite it with the scores two halves further on
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(2 * t + hf));
        float m_tile = -INFINITY;
        if (valid < kC) {
#pragma unroll
          for (int i = 0; i < kC; ++i) {
            if (i >= valid) s[i] = 0xff800000u;   // padded keys -> -inf: exp2 gives 0
            m_tile = fmaxf(m_tile, __uint_as_float(s[i]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < kC; i += 2)
            m_tile = fmaxf(m_tile, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
        }
        // Lazy rescaling: the running maximum is only advanced (and O_t rescaled) when some row of this warp saw its
        // maximum grow by more than 2^kLazyLog2; otherwise the stale maximum keeps being the reference — P may then
        // exceed 1 (by at most 2^kLazyLog2, harmless in fp32 / bf16) and O / l at the end is unchanged mathematically.
        constexpr float kLazyLog2 = 8.0f;
        const bool grow = (g > 0) && __any_sync(0xffffffffu, (m_tile - m_run) * c > kLazyLog2);
        const float m_new = (g == 0) ? m_tile : (grow ? fmaxf(m_run, m_tile) : m_run);
        const float alpha = grow ? ex2_approx((m_run - m_new) * c) : 1.0f;
        const float mc = m_new * c;
        m_run = m_new;

        // this half of P_t is free once the P·V of its previous use has finished (issued two halves ago: never blocks)
        TR(4);
        WT(1, mbar_wait(pv_done(2 * t + hf), (u & 1u) ^ 1u));
        TR(grow ? 6 : 5);
        if (grow) {
          // rescale the running output by alpha: no P·V of this tile may be in flight, also not the other half's
          const uint32_t uo = hf ? use0 : use1;   // >= 1 because g > 0
          mbar_wait(pv_done(2 * t + (hf ^ 1)), (uo & 1u) ^ 1u);
          tc_fence_after();
          const uint64_t alpha2 = pack_f32x2(alpha, alpha);
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(to + uint32_t(16 * hh), o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float a0, a1;
              unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), a0, a1);
              o[i] = __float_as_uint(a0);
              o[i + 1] = __float_as_uint(a1);
            }
            tmem_st_32x32b_x16(to + uint32_t(16 * hh), o);
          }
          tmem_wait_st();
          tc_fence_before();
        }
        if (store_pending) {
          // the previous item's output tile was staged in this warp's rows of P_t
          WT(3, if (lane == 0) tma_store_wait_read<0>(); __syncwarp());
          store_pending = false;
        }
        TR(7);
        // P = exp2(S*c - m*c) -> bf16 -> 128B-swizzled smem (A operand of P·V), 8 keys = one 16-byte chunk
        uint64_t l2 = pack_f32x2(0.f, 0.f);   // two partial row sums (even / odd keys), packed fp32x2 adds
        const uint64_t nmc2 = pack_f32x2(-mc, -mc);
#pragma unroll
        for (int ch = 0; ch < kC / 8; ++ch) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a0, a1;
            unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[8 * ch + 2 * i]), __uint_as_float(s[8 * ch + 2 * i + 1])),
                                   c2, nmc2), a0, a1);
            float e0 = ex2_approx(a0);
            float e1 = ex2_approx(a1);
            l2 = add_f32x2(l2, pack_f32x2(e0, e1));
            if constexpr (kDrop) {
              const uint32_t kidx = drop_row + uint32_t(g * kAtHN + 8 * ch + 2 * i);
              if (!attn_keep(bh_seed, kidx, p.drop_thresh)) e0 = 0.f;
              if (!attn_keep(bh_seed, kidx + 1u, p.drop_thresh)) e1 = 0.f;
            }
            pk[i] = pack_bf16x2(e0, e1);
          }
          const uint32_t chunk16 = uint32_t(4 * hf + ch);
          const uint32_t addr = prow + ((chunk16 ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                       "r"(pk[3])
                       : "memory");
        }
        float l_lo, l_hi;
        unpack_f32x2(l2, l_lo, l_hi);
        l_run = fmaf(l_run, alpha, l_lo + l_hi);
        TR(8);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(2 * t + hf));
        TR(9);
      };

#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        const int hf = g & 1;
        const int valid = min(kAtHN, p.N - g * kAtHN);
        const uint32_t u = hf ? use1 : use0;
        TR(1);
        WT(0, mbar_wait(s_full(2 * t + hf), u & 1u));
        TR(2);
        tc_fence_after();
        if (valid > 16) half_step(std::integral_constant<int, 32>{}, g, hf, valid, u);
        else half_step(std::integral_constant<int, 16>{}, g, hf, valid, u);
        if (hf) ++use1; else ++use0;
      }

      // ---- epilogue: O_t / l -> bf16 -> staged in this warp's rows of P_t -> TMA store ----
      TR(10);
      WT(2, mbar_wait(pv_done(2 * t + 0), (use0 & 1u) ^ 1u));
      if (use1 > 0) WT(2, mbar_wait(pv_done(2 * t + 1), (use1 & 1u) ^ 1u));
      TR(11);
      tc_fence_after();
      const float inv_l = p.drop_scale / l_run;
      const uint64_t inv_l2 = pack_f32x2(inv_l, inv_l);
      {
        const int qrow = qp * 256 + t * 128 + r_in_tile;
        if (p.lse2 != nullptr && qrow < p.N) p.lse2[(size_t(b) * H + h) * p.N + qrow] = fmaf(m_run, c, log2f(l_run));
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(to + uint32_t(32 * hh), o);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t ob[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a0, a1;
            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[8 * g + 2 * i]), __uint_as_float(o[8 * g + 2 * i + 1])), inv_l2),
                         a0, a1);
            ob[i] = pack_bf16x2(a0, a1);
          }
          const uint32_t addr = prow + ((uint32_t(4 * hh + g) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ob[0]), "r"(ob[1]), "r"(ob[2]), "r"(ob[3])
                       : "memory");
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_4d(&tm_out, pbase + uint32_t(32 * q * 128), 0, h, row0, b);
        tma_store_commit();
      }
      TR(12);
      store_pending = true;
    }
    if (lane == 0) tma_store_wait_all<0>();
    __syncwarp();
  }

#if MT_ATTN_TRACE == 2
  if (p.trace != nullptr && blockIdx.x < 4 && lane == 0) {
    long long* o = p.trace + (size_t(blockIdx.x) * 12 + warp) * 8;
    o[0] = wt0; o[1] = wt1; o[2] = wt2; o[3] = wt3; o[4] = clock64() - wt_begin;
  }
#endif
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, kAtTmemCols);
  timeline_end(p.tl);
}

}  // namespace mt

namespace mt {
int attention_generic_fwd(const void* qkv, void* out, int B, int N, int H, int dh, float scale, cudaStream_t s);
}

// MT_ATTN_TRACE builds: device buffer of 2 x 12 x 1024 int64 that the next launches log into (nullptr = off)
static long long* g_attn_trace = nullptr;
extern "C" int mt_attention_set_trace(void* buf) {
  g_attn_trace = reinterpret_cast<long long*>(buf);
  return MT_ATTN_TRACE ? -MT_ATTN_TRACE : 1;   // 1: this build has no trace points; -1 event log, -2 wait accounting
}

static int attention_fwd_impl(const void* qkv, void* out, float* lse2, int B, int N, int H, int dh, float scale,
                              void* stream, float drop_p = 0.f, unsigned long long seed = 0ull) {
  using namespace mt;
  MT_REQUIRE(qkv && out, MT_ERR_BAD_SHAPE, "mt_attention_fwd: null pointer");
  MT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MT_ERR_BAD_SHAPE, "mt_attention_fwd: drop_p must be in [0, 1)");
  MT_REQUIRE(drop_p == 0.f || (long long)N * N < (1ll << 32), MT_ERR_UNSUPPORTED, "mt_attention_fwd: dropout needs N < 65536");
  if (dh != 64 && lse2 == nullptr && drop_p == 0.f && B > 0 && N > 0 && H > 0 && scale > 0.f &&
      ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    return attention_generic_fwd(qkv, out, B, N, H, dh, scale, reinterpret_cast<cudaStream_t>(stream));
  MT_REQUIRE(B > 0 && N > 0 && H > 0, MT_ERR_BAD_SHAPE, "mt_attention_fwd: B, N, H must be positive");
  MT_REQUIRE(dh == 64, MT_ERR_UNSUPPORTED, "mt_attention_fwd: head_dim %d not supported (only 64)", dh);
  MT_REQUIRE(scale > 0.f, MT_ERR_UNSUPPORTED, "mt_attention_fwd: scale must be positive");
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, MT_ERR_BAD_ALIGN,
             "mt_attention_fwd: pointers must be 16-byte aligned");
  CUtensorMap tm_q, tm_kv, tm_out;
  {
    // (B, N, 3, H, 64) viewed as dims [64, 3H, N, B]; 128-row boxes for Q tiles, 64-row boxes for K / V tiles
    uint64_t dims[4] = {64, uint64_t(3 * H), uint64_t(N), uint64_t(B)};
    uint64_t strides[3] = {64 * 2, uint64_t(3 * H) * 64 * 2, uint64_t(N) * 3 * H * 64 * 2};
    uint32_t box_q[4] = {64, 1, uint32_t(kAtBM), 1};
    int rc = encode_tmap(&tm_q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, qkv, dims, strides, box_q,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    uint32_t box_kv[4] = {64, 1, uint32_t(kAtBN), 1};
    rc = encode_tmap(&tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, qkv, dims, strides, box_kv, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    // (B, N, H, 64) viewed as dims [64, H, N, B]; one store = 32 query rows of one head
    uint64_t dims[4] = {64, uint64_t(H), uint64_t(N), uint64_t(B)};
    uint64_t strides[3] = {64 * 2, uint64_t(H) * 64 * 2, uint64_t(N) * H * 64 * 2};
    uint32_t box[4] = {64, 1, 32, 1};
    int rc = encode_tmap(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.B = B; p.N = N; p.H = H;
  p.num_qp = ceil_div(N, 256);
  p.num_kv = ceil_div(N, kAtBN);
  p.num_half = ceil_div(N, kAtHN);
  p.num_items = B * H * p.num_qp;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.tl = timeline_next_slot();
  p.trace = g_attn_trace;
  p.drop_thresh = drop_p > 0.f ? uint32_t(double(drop_p) * 4294967296.0) : 0u;
  p.drop_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.seed_lo = uint32_t(seed);
  p.seed_hi = uint32_t(seed >> 32);
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  int grid = p.num_items < 2 * sms ? p.num_items : 2 * sms;    // two CTAs per SM
  {
    // the opt-in to > 48 KB of dynamic shared memory is per device (nn.DataParallel drives several from one process)
    static bool attr_set[64] = {false};
    int dev = 0;
    MT_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmemBytes));
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  if (p.drop_thresh != 0u)
    MT_CHECK_CUDA(launch_pdl(attention_fwd_kernel<true>, dim3(unsigned(grid)), dim3(kAtThreads), size_t(kAtSmemBytes),
                             reinterpret_cast<cudaStream_t>(stream), 1u, tm_q, tm_kv, tm_out, p));
  else
    MT_CHECK_CUDA(launch_pdl(attention_fwd_kernel<false>, dim3(unsigned(grid)), dim3(kAtThreads), size_t(kAtSmemBytes),
                             reinterpret_cast<cudaStream_t>(stream), 1u, tm_q, tm_kv, tm_out, p));
  return MT_OK;
}

extern "C" int mt_attention_fwd(const void* qkv, void* out, int B, int N, int H, int dh, float scale, void* stream) {
  return attention_fwd_impl(qkv, out, nullptr, B, N, H, dh, scale, stream);
}

extern "C" int mt_attention_fwd_lse(const void* qkv, void* out, float* lse2, int B, int N, int H, int dh, float scale,
                                    void* stream) {
  if (!lse2) {
    mt::set_error("mt_attention_fwd_lse: lse2 is required");
    return MT_ERR_BAD_SHAPE;
  }
  return attention_fwd_impl(qkv, out, lse2, B, N, H, dh, scale, stream);
}

// training forward with attention dropout (attn_drop > 0: Graph/metatransformer/modules/tokengt_graph_encoder.py:200):
// softmax, then Bernoulli(1 - drop_p) keep mask scaled by 1 / (1 - drop_p), then .V; the mask is a pure function of
// (seed, batch, head, query, key) that mt_attention_bwd_ex and mt_attention_dropout_mask replay.
extern "C" int mt_attention_fwd_ex(const void* qkv, void* out, float* lse2, int B, int N, int H, int dh, float scale,
                                   float drop_p, unsigned long long seed, void* stream) {
  return attention_fwd_impl(qkv, out, lse2, B, N, H, dh, scale, stream, drop_p, seed);
}

namespace mt {
__global__ void attn_dropout_mask_kernel(uint8_t* __restrict__ mask, int BH, int N, uint32_t thresh, uint32_t seed_lo,
                                         uint32_t seed_hi) {
  const long long total = (long long)BH * N * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t bh = uint32_t(i / ((long long)N * N));
    const uint32_t idx = uint32_t(i % ((long long)N * N));
    mask[i] = attn_keep(attn_drop_seed(seed_lo, seed_hi, bh), idx, thresh) ? 1 : 0;
  }
}
}  // namespace mt

// the keep mask [B, H, N, N] (uint8) the attention kernels apply for (drop_p, seed): test / debugging aid
extern "C" int mt_attention_dropout_mask(void* mask, int B, int N, int H, float drop_p, unsigned long long seed, void* stream) {
  using namespace mt;
  MT_REQUIRE(mask && B > 0 && N > 0 && H > 0 && drop_p > 0.f && drop_p < 1.f, MT_ERR_BAD_SHAPE,
             "mt_attention_dropout_mask: bad arguments");
  attn_dropout_mask_kernel<<<1184, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint8_t*>(mask), B * H, N, uint32_t(double(drop_p) * 4294967296.0), uint32_t(seed), uint32_t(seed >> 32));
  MT_CHECK_CUDA(cudaGetLastError());
  return MT_OK;
}
