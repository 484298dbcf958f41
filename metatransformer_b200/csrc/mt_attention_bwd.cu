// mt_attention_bwd.cu — backward of the softmax attention core (SURVEY.md §8a rows E4 + E8): gradients of
//   O = softmax(Q K^T * scale) V        (PointCloud/openpoints/models/layers/attention.py:31-35)
// with respect to q, k, v, written in the packed (B, N, 3, H, 64) layout of the qkv Linear's output so that the
// dgrad GEMM of `attn.qkv` can consume it directly.  Needed because the tokenizers train under a frozen encoder
// (Audio/src/models/ast_models.py:66-70): dX must flow through every Block.
//
// FlashAttention-2-style recomputation, Blackwell edition (sm_100a only, head_dim 64):
//   work item = (batch, head); loop over key tiles j of 128 keys, inside it over query tiles i of 128 rows:
//     S^T  = K_j Q_i^T            tcgen05.mma 128x128x16   (TMEM cols   0..127)
//     dP^T = V_j dO_i^T           tcgen05.mma 128x128x16   (TMEM cols 128..255)
//     P^T  = exp2(S^T c - lse2_i) ;  dS^T = P^T (dP^T - delta_i) scale      (one key row per thread, fp32 registers)
//     dV_j += P^T  dO_i           A = P^T  (smem, K-major),  B = dO_i (MN-major)     (TMEM cols 256..319)
//     dK_j += dS^T Q_i            A = dS^T (smem, K-major),  B = Q_i  (MN-major)     (TMEM cols 320..383)
//     dQ_i  = dS   K_j            A = dS^T read MN-major,     B = K_j  (MN-major)     (TMEM cols 384..447)
//   Work decomposition (picked on the host by the number of (batch, head) pairs per SM):
//    * owner: a CTA owns one (batch, head) at a time and walks its key tiles j; with <= 2 key tiles (N <= 256) dQ_i
//      stays RESIDENT in TMEM across the key tiles (one accumulator per query tile, accumulate flag of tcgen05.mma)
//      and is written once, as bf16 — no atomics, no scratch, no zero-fill, no conversion pass, deterministic;
//    * split: work item = (batch, head, key tile) for few, long sequences.
//   Whenever dQ is not resident, each warp stages its 32 x 32 fp32 partial in shared memory and the TMA engine adds
//   it to a zero-filled fp32 accumulator (cp.reduce.async.bulk.tensor ... add); a small kernel converts to bf16.
//   dK_j / dV_j leave once per key tile.
//   lse2 = log2-sum-exp saved by the forward, delta_i = rowsum(dO_i * O_i) from a small pre-pass kernel.
// Warps: 0-7 compute (warp & 3 = TMEM lane quarter = 32 key rows, warp >> 2 = which 64-query half of the tile),
//        8 = TMA loader, 9 = MMA issuer + TMEM allocator.
#include "mt_ptx.cuh"
#include "mt_common.cuh"
#include <stdlib.h>

namespace mt {

constexpr int kAbT = 128 * 64 * 2;          // bytes of one 128 x 64 bf16 tile
constexpr int kAbSmemK = 0;                 // 2 slots (key tiles of successive work items)
constexpr int kAbSmemV = kAbSmemK + 2 * kAbT;
constexpr int kAbSmemQ = kAbSmemV + 2 * kAbT;   // 2 stages
constexpr int kAbSmemDO = kAbSmemQ + 2 * kAbT;
constexpr int kAbSmemP = kAbSmemDO + 2 * kAbT;   // P^T: 2 chunks of 64 queries
constexpr int kAbSmemDS = kAbSmemP + 2 * kAbT;   // dS^T
constexpr int kAbSmemStage = kAbSmemDS + 2 * kAbT;   // dQ partial staging: 8 warps x (32 rows x 32 fp32), 128B-swizzled
constexpr int kAbSmemStat = kAbSmemStage + 8 * 4096;  // 2 x (128 lse2 + 128 delta) floats
constexpr int kAbSmemBar = kAbSmemStat + 2 * 256 * 4;
// barriers: kv_full[2] kv_empty[2] qd_full[2] qd_empty[2] s_full s_empty p_full pv_done
constexpr int kAbNumBars = 12;
constexpr int kAbSmemTmemPtr = kAbSmemBar + kAbNumBars * 8;
// No alignment slack: the dynamic shared memory window of a kernel without static __shared__ data starts 1 KB-aligned
// (checked at kernel entry), and 6 x 32 KB of tiles + 32 KB of staging leave < 1 KB of the 227 KB.
constexpr int kAbSmemBytes = kAbSmemTmemPtr + 16;
static_assert(kAbSmemBytes <= 227 * 1024, "attention_bwd_kernel: shared memory budget");
constexpr int kAbThreads = 320;
constexpr int kAbComputeWarps = 8;
constexpr uint32_t kAbTmS = 0, kAbTmDP = 128, kAbTmDV = 256, kAbTmDK = 320, kAbTmDQ = 384;   // dQ: 384 (+64 resident #2)

struct AttnBwdParams {
  int B, N, H;
  int num_t;         // ceil(N / 128): key tiles == query tiles
  int num_bh;        // B * H
  int owner;         // 1: work item = (batch, head), key tiles walked inside; 0: work item = (batch, head, key tile)
  int resident;      // owner && num_t <= 2: dQ_i stays in TMEM across the key tiles (no scratch at all)
  int num_work;      // owner ? num_bh : num_bh * num_t
  float scale, scale_log2;
  const float* lse2;     // [B, H, N]
  const float* delta;    // [B, H, N]
  float* dq_acc;         // [B, N, H, 64] fp32 zero-filled accumulator (unused in resident mode)
  __nv_bfloat16* dqkv;   // [B, N, 3, H, 64]
  // attention dropout of the forward (mt_attention_fwd_ex), replayed from the same hash: A = P * keep / (1 - p) feeds dV,
  // dS = P * (keep / (1 - p) * dP - delta) * scale (delta = rowsum(dO * O) already contains the dropped probabilities)
  uint32_t drop_thresh;
  float drop_scale;
  uint32_t seed_lo, seed_hi;
};

// 32-byte store of 16 bf16 values held as fp32 bit patterns in o[0..15] (full-sector writes from scattered rows)
MT_DEVINL void st_global_bf16x16(__nv_bfloat16* dst, const uint32_t* o) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst),
               "r"(pack_bf16x2(__uint_as_float(o[0]), __uint_as_float(o[1]))),
               "r"(pack_bf16x2(__uint_as_float(o[2]), __uint_as_float(o[3]))),
               "r"(pack_bf16x2(__uint_as_float(o[4]), __uint_as_float(o[5]))),
               "r"(pack_bf16x2(__uint_as_float(o[6]), __uint_as_float(o[7]))),
               "r"(pack_bf16x2(__uint_as_float(o[8]), __uint_as_float(o[9]))),
               "r"(pack_bf16x2(__uint_as_float(o[10]), __uint_as_float(o[11]))),
               "r"(pack_bf16x2(__uint_as_float(o[12]), __uint_as_float(o[13]))),
               "r"(pack_bf16x2(__uint_as_float(o[14]), __uint_as_float(o[15])))
               : "memory");
}

MT_DEVINL float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The sequence of (work item w, key tile j, query tile i) pairs one CTA walks; all three warp roles step through it
// with the same function, so their barrier phases stay in lock step.
struct AbPair {
  int w, j, i;
  bool valid;
};
MT_DEVINL AbPair ab_first(const AttnBwdParams& p, int NT) {
  AbPair s;
  s.w = int(blockIdx.x);
  s.valid = s.w < p.num_work;
  s.j = p.owner ? 0 : s.w % NT;
  s.i = 0;
  return s;
}
MT_DEVINL AbPair ab_next(AbPair s, const AttnBwdParams& p, int NT) {
  if (++s.i < NT) return s;
  s.i = 0;
  const int j1 = p.owner ? NT : (s.w % NT) + 1;
  if (++s.j < j1) return s;
  s.w += int(gridDim.x);
  s.valid = s.w < p.num_work;
  s.j = p.owner ? 0 : (s.valid ? s.w % NT : 0);
  return s;
}
MT_DEVINL int ab_bh(const AbPair& s, const AttnBwdParams& p, int NT) { return p.owner ? s.w : s.w / NT; }

// delta[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]   (one warp per (b,n) row; lane pairs cover one head)
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                              const __nv_bfloat16* __restrict__ dout,
                                                              float* __restrict__ delta, int B, int N, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;   // b*N + n
  if (row >= (long long)B * N) return;
  const int b = int(row / N), n = int(row % N);
  const size_t base = size_t(row) * H * 64;
  for (int c0 = 0; c0 < H * 8; c0 += 32) {   // 8-element chunks; 8 chunks per head; all lanes stay in the loop
    const int c = c0 + lane;
    uint4 a = make_uint4(0u, 0u, 0u, 0u), g = a;
    if (c < H * 8) {
      a = *reinterpret_cast<const uint4*>(o + base + c * 8);
      g = *reinterpret_cast<const uint4*>(dout + base + c * 8);
    }
    float s = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
              bf16_hi(a.y) * bf16_hi(g.y) + bf16_lo(a.z) * bf16_lo(g.z) + bf16_hi(a.z) * bf16_hi(g.z) +
              bf16_lo(a.w) * bf16_lo(g.w) + bf16_hi(a.w) * bf16_hi(g.w);
    // the 8 chunks of one head sit in 8 consecutive lanes
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && c < H * 8) delta[(size_t(b) * H + (c >> 3)) * N + n] = s;
  }
}

// dq_acc fp32 [B,N,H,64] -> bf16 q-slot of dqkv [B,N,3,H,64]
__global__ void __launch_bounds__(256) attn_bwd_dq_store_kernel(const float* __restrict__ dq_acc,
                                                                 __nv_bfloat16* __restrict__ dqkv, long long rows,
                                                                 int H) {
  const long long total = rows * H * 8;   // 8-element chunks
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / (H * 8);
    const int c = int(idx % (H * 8));
    const float4 a = *reinterpret_cast<const float4*>(dq_acc + size_t(idx) * 8);
    const float4 b = *reinterpret_cast<const float4*>(dq_acc + size_t(idx) * 8 + 4);
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y); w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y); w.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dqkv + size_t(row) * 3 * H * 64 + size_t(c) * 8) = w;
  }
}

template <bool kDrop>
__global__ void __launch_bounds__(kAbThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_dq, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  uint8_t* smem_gen = smem_raw;
  if ((smem_base & 1023u) != 0u) {   // the 128B-swizzled tiles need 1 KB alignment
    if (threadIdx.x == 0) printf("[mt_b200] attention_bwd_kernel: dynamic smem base 0x%x is not 1 KB aligned\n", smem_base);
    __trap();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_base + kAbSmemBar;
  auto kv_full = [&](int s) { return bar0 + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto qd_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto qd_empty = [&](int s) { return bar0 + 8u * (6 + s); };
  const uint32_t s_full = bar0 + 64, s_empty = bar0 + 72, p_full = bar0 + 80, pv_done = bar0 + 88;
  const uint32_t tmem_ptr_addr = smem_base + kAbSmemTmemPtr;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
      mbar_init(qd_full(s), 1); mbar_init(qd_empty(s), 1);
    }
    mbar_init(s_full, 1); mbar_init(s_empty, kAbComputeWarps); mbar_init(p_full, kAbComputeWarps); mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 9) tmem_alloc<1>(tmem_ptr_addr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kAbSmemTmemPtr);
  const int H = p.H, NT = p.num_t;

  if (warp == 8) {
    // =============================== TMA loader =================================
    // K_j / V_j: two slots, so the tiles of the next work item arrive while the current one is being consumed;
    // Q_i / dO_i: two stages.
    if (lane == 0) {
      uint32_t n_item = 0, n_pair = 0;
      for (AbPair cur = ab_first(p, NT); cur.valid; cur = ab_next(cur, p, NT), ++n_pair) {
        const int bh = ab_bh(cur, p, NT), h = bh % H, b = bh / H;
        if (cur.i == 0) {
          const int sl = int(n_item & 1u);
          mbar_wait(kv_empty(sl), ((n_item >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(kv_full(sl), 2 * kAbT);
          tma_load_4d(smem_base + kAbSmemK + sl * kAbT, &tm_qkv, kv_full(sl), 0, H + h, cur.j * 128, b, kEvictLast);
          tma_load_4d(smem_base + kAbSmemV + sl * kAbT, &tm_qkv, kv_full(sl), 0, 2 * H + h, cur.j * 128, b, kEvictLast);
          ++n_item;
        }
        const int st = int(n_pair & 1u);
        mbar_wait(qd_empty(st), ((n_pair >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(qd_full(st), 2 * kAbT);
        tma_load_4d(smem_base + kAbSmemQ + st * kAbT, &tm_qkv, qd_full(st), 0, h, cur.i * 128, b, kEvictLast);
        tma_load_4d(smem_base + kAbSmemDO + st * kAbT, &tm_do, qd_full(st), 0, h, cur.i * 128, b, kEvictLast);
      }
    }
  } else if (warp == 9) {
    // =============================== MMA issuer =================================
    // Software pipeline: S^T / dP^T of pair n+1 are issued as soon as the compute warps have pulled pair n's scores
    // into registers (s_empty), i.e. BEFORE the dV / dK / dQ MMAs of pair n (which need P^T / dS^T of pair n):
    // the tensor core computes the next scores while the compute warps are busy with the current ones.
    if (lane == 0) {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);    // K-major x K-major
      constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);    // A K-major (P^T / dS^T), B MN-major (dO / Q)
      constexpr uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);    // A = dS read MN-major, B = K MN-major
      const uint64_t p_desc = make_smem_desc_sw128(smem_base + kAbSmemP);
      const uint64_t ds_desc = make_smem_desc_sw128(smem_base + kAbSmemDS);
      // dS^T tile read as an MN-major A operand: the two 64-query column blocks are 16 KB apart (leading offset)
      const uint64_t ds_mn_desc = make_smem_desc_sw128(smem_base + kAbSmemDS, 1024, 2 * 8192);
      uint32_t n_pair = 0, n_item = 0;   // n_item: items whose K/V have been waited for
      int sl_cur = 0;
      auto issue_scores = [&](const AbPair& pr, uint32_t np, int sl) {
        const int st = int(np & 1u);
        mbar_wait(qd_full(st), (np >> 1) & 1u);
        tc_fence_after();
        const uint64_t k_desc = make_smem_desc_sw128(smem_base + kAbSmemK + sl * kAbT);
        const uint64_t v_desc = make_smem_desc_sw128(smem_base + kAbSmemV + sl * kAbT);
        const uint64_t q_desc = make_smem_desc_sw128(smem_base + kAbSmemQ + st * kAbT);
        const uint64_t do_desc = make_smem_desc_sw128(smem_base + kAbSmemDO + st * kAbT);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // S^T = K Q^T
          umma_f16_ss<1>(tmem_base + kAbTmS, k_desc + uint64_t(2 * k), q_desc + uint64_t(2 * k), id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // dP^T = V dO^T
          umma_f16_ss<1>(tmem_base + kAbTmDP, v_desc + uint64_t(2 * k), do_desc + uint64_t(2 * k), id_s, k > 0);
        umma_commit(s_full);
        (void)pr;
      };
      AbPair cur = ab_first(p, NT);
      if (cur.valid) {
        mbar_wait(kv_full(0), 0u);
        n_item = 1;
        issue_scores(cur, 0u, 0);
      }
      while (cur.valid) {
        const AbPair nxt = ab_next(cur, p, NT);
        int sl_nxt = sl_cur;
        if (nxt.valid) {
          mbar_wait(s_empty, n_pair & 1u);          // pair n's scores are in registers: S / dP TMEM is free
          if (nxt.i == 0) {                         // next pair opens a new (work item, key tile): its K / V slot
            sl_nxt = int(n_item & 1u);
            mbar_wait(kv_full(sl_nxt), (n_item >> 1) & 1u);
            ++n_item;
          }
          issue_scores(nxt, n_pair + 1, sl_nxt);
        }
        mbar_wait(p_full, n_pair & 1u);             // P^T / dS^T of pair n are in smem; older dQ / dK / dV drains done
        tc_fence_after();
        {
          const int st = int(n_pair & 1u);
          const uint64_t k_desc = make_smem_desc_sw128(smem_base + kAbSmemK + sl_cur * kAbT);
          const uint64_t q_desc = make_smem_desc_sw128(smem_base + kAbSmemQ + st * kAbT);
          const uint64_t do_desc = make_smem_desc_sw128(smem_base + kAbSmemDO + st * kAbT);
          const uint32_t acc_kv = cur.i > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // 16 queries per step: A chunk (k/4) 16 KB further, 32 B per step; B 2 KB per step
            const uint64_t a_off = uint64_t((k >> 2) * (kAbT >> 4) + 2 * (k & 3));
            const uint64_t b_off = uint64_t(k * (2048 >> 4));
            umma_f16_ss<1>(tmem_base + kAbTmDV, p_desc + a_off, do_desc + b_off, id_kv, (acc_kv || k > 0) ? 1u : 0u);
            umma_f16_ss<1>(tmem_base + kAbTmDK, ds_desc + a_off, q_desc + b_off, id_kv, (acc_kv || k > 0) ? 1u : 0u);
          }
          // dQ_i: resident mode keeps one accumulator per query tile across the key tiles of this (batch, head)
          const uint32_t dq_col = kAbTmDQ + (p.resident ? uint32_t(64 * cur.i) : 0u);
          const uint32_t acc_q = (p.resident && cur.j > 0) ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // dQ = dS K: 16 keys (= 16 smem rows = 2 KB) per step on both operands
            const uint64_t off = uint64_t(k * (2048 >> 4));
            umma_f16_ss<1>(tmem_base + dq_col, ds_mn_desc + off, k_desc + off, id_dq, (acc_q || k > 0) ? 1u : 0u);
          }
          umma_commit(pv_done);
          umma_commit(qd_empty(st));
          if (cur.i == NT - 1) umma_commit(kv_empty(sl_cur));
        }
        cur = nxt;
        sl_cur = sl_nxt;
        ++n_pair;
      }
    }
  } else {
    // =============================== compute warps ==============================
    const int q4 = warp & 3;                 // TMEM lane quarter
    const int hf = warp >> 2;                // which 64-query half of S^T / dP^T (and which half of the drains)
    const int r = 32 * q4 + lane;            // row inside the tile (key row for S^T/dP^T/dK/dV, query row for dQ)
    const uint32_t lane_off = uint32_t(32 * q4) << 16;
    const uint32_t row_smem = uint32_t(r * 128);
    const uint32_t swz = uint32_t(r & 7);
    float* stat = reinterpret_cast<float*>(smem_gen + kAbSmemStat);
    uint32_t n_pair = 0;
    const float c = p.scale_log2, sc = p.scale;
    // lse2 / delta of a query tile are fetched one pair ahead by warps 0-3 (one query row per thread), so their
    // global-load latency hides behind the previous pair (padded queries: +inf / 0 -> P = 0)
    float pf_l = INFINITY, pf_d = 0.f;
    auto prefetch_stats = [&](const AbPair& pr) {
      pf_l = INFINITY;
      pf_d = 0.f;
      if (hf == 0 && pr.valid) {
        const int bh_ = ab_bh(pr, p, NT), h_ = bh_ % H, b_ = bh_ / H;
        const int qi = pr.i * 128 + r;
        if (qi < p.N) {
          const size_t off = (size_t(b_) * H + h_) * p.N + qi;
          pf_l = p.lse2[off];
          pf_d = p.delta[off];
        }
      }
    };
    // dQ of one pair leaves TMEM (this warp: 32 query rows x 32 of the 64 columns)
    bool stage_pending = false;   // this warp's staging buffer is being read by an earlier TMA reduce
    auto drain_dq = [&](const AbPair& pr) {
      const int bh_ = ab_bh(pr, p, NT), h_ = bh_ % H, b_ = bh_ / H;
      const int qi = pr.i * 128 + r;
      uint32_t o[32];
      tmem_ld_32x32b_x32(tmem_base + lane_off + kAbTmDQ + (p.resident ? uint32_t(64 * pr.i) : 0u) + uint32_t(32 * hf), o);
      tmem_wait_ld();
      tc_fence_before();
      if (!p.resident) {
        // partial of one key tile: staged in this warp's 4 KB of shared memory (128B swizzle) and added to the fp32
        // accumulator in HBM by the TMA engine (one bulk reduce per warp instead of 8 scattered 16-byte atomics per
        // thread; rows beyond N are clipped by the tensor map)
        const uint32_t stg = smem_base + kAbSmemStage + uint32_t(warp) * 4096u;
        if (stage_pending) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t addr = stg + uint32_t(lane * 128) + ((uint32_t(e) ^ uint32_t(lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[4 * e]), "r"(o[4 * e + 1]),
                       "r"(o[4 * e + 2]), "r"(o[4 * e + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_4d(&tm_dq, stg, 32 * hf, h_, pr.i * 128 + 32 * q4, b_);
          tma_store_commit();
        }
        stage_pending = true;
        return;
      }
      if (qi >= p.N) return;
      {
        __nv_bfloat16* dst = p.dqkv + ((size_t(b_) * p.N + qi) * 3) * H * 64 + size_t(h_) * 64 + 32 * hf;   // q slot
        st_global_bf16x16(dst, o);
        st_global_bf16x16(dst + 16, o + 16);
      }
    };
    // dK_j / dV_j of a finished key tile leave TMEM: warps 0-3 store dK (slot 1), warps 4-7 dV (slot 2)
    auto drain_kv = [&](const AbPair& pr) {
      const int bh_ = ab_bh(pr, p, NT), h_ = bh_ % H, b_ = bh_ / H;
      const int kv = pr.j * 128 + r;
      __nv_bfloat16* dst = p.dqkv + ((size_t(b_) * p.N + kv) * 3 + size_t(hf + 1)) * H * 64 + size_t(h_) * 64;
      const uint32_t tcol = hf == 0 ? kAbTmDK : kAbTmDV;
      uint32_t o[64];
      tmem_ld_32x32b_x32_at<0>(tmem_base + lane_off + tcol, o);
      tmem_ld_32x32b_x32_at<32>(tmem_base + lane_off + tcol + 32u, o);
      tmem_wait_ld();
      tc_fence_before();
      if (kv < p.N) {
#pragma unroll
        for (int e = 0; e < 64; e += 16) st_global_bf16x16(dst + e, o + e);
      }
    };
    auto needs_drain = [&](const AbPair& pr) { return !p.resident || pr.j == NT - 1; };

    AbPair cur = ab_first(p, NT);
    prefetch_stats(cur);
    // Results of pair n (its dQ; dK_j / dV_j when it closes a key tile) stay in TMEM until the middle of pair n+1: by
    // then their MMAs have retired (no idle wait), and the MMAs that overwrite those accumulators are only issued
    // after pair n+1's p_full, which this warp signals after the drain.
    AbPair pending = cur;
    bool pend_dq = false, pend_kv = false;
    for (; cur.valid; ++n_pair) {
      const AbPair nxt = ab_next(cur, p, NT);
      const bool key_ok = (cur.j * 128 + r) < p.N;
      const uint32_t bh_seed = attn_drop_seed(p.seed_lo, p.seed_hi, uint32_t(ab_bh(cur, p, NT)));
      const uint32_t key_idx = uint32_t(cur.j * 128 + r);
      float* st_l = stat + (n_pair & 1u) * 256;
      if (hf == 0) {
        st_l[r] = pf_l;
        st_l[128 + r] = pf_d;
      }
      named_bar_sync(1, 32 * kAbComputeWarps);
      prefetch_stats(nxt);
      mbar_wait(s_full, n_pair & 1u);
      tc_fence_after();
      uint32_t pp[2][16], dd[2][16];     // bf16x2-packed P^T / dS^T of the two 32-query slices of this warp's half
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + kAbTmS + uint32_t(64 * hf + 32 * sub), s);
        tmem_ld_32x32b_x32(tmem_base + lane_off + kAbTmDP + uint32_t(64 * hf + 32 * sub), dp);
        tmem_wait_ld();
        if (sub == 1) {
          // both slices are in registers: the tensor core may start the next pair's scores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int c0 = 2 * e;
          const float2 l2 = *reinterpret_cast<const float2*>(st_l + 64 * hf + 32 * sub + c0);
          const float2 dl = *reinterpret_cast<const float2*>(st_l + 128 + 64 * hf + 32 * sub + c0);
          float p0 = ex2f(fmaf(__uint_as_float(s[c0]), c, -l2.x));
          float p1 = ex2f(fmaf(__uint_as_float(s[c0 + 1]), c, -l2.y));
          if (!key_ok) { p0 = 0.f; p1 = 0.f; }
          float m0 = 1.f, m1 = 1.f;
          if constexpr (kDrop) {
            const uint32_t qidx = uint32_t(cur.i * 128 + 64 * hf + 32 * sub + c0) * uint32_t(p.N) + key_idx;
            m0 = attn_keep(bh_seed, qidx, p.drop_thresh) ? p.drop_scale : 0.f;
            m1 = attn_keep(bh_seed, qidx + uint32_t(p.N), p.drop_thresh) ? p.drop_scale : 0.f;
          }
          const float d0 = p0 * fmaf(__uint_as_float(dp[c0]), m0, -dl.x) * sc;
          const float d1 = p1 * fmaf(__uint_as_float(dp[c0 + 1]), m1, -dl.y) * sc;
          p0 *= m0;
          p1 *= m1;
          pp[sub][e] = pack_bf16x2(p0, p1);
          dd[sub][e] = pack_bf16x2(d0, d1);
        }
      }
      // P^T / dS^T smem still feeds the previous pair's dV / dK / dQ MMAs: wait for them, and take its results out
      if (n_pair > 0) {
        mbar_wait(pv_done, (n_pair - 1) & 1u);
        tc_fence_after();
        if (pend_dq) drain_dq(pending);
        if (pend_kv) drain_kv(pending);
      }
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // four 16-byte groups = 8 queries each
          const uint32_t a = row_smem + uint32_t(hf * kAbT) + ((uint32_t(4 * sub + g) ^ swz) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + kAbSmemP + a), "r"(pp[sub][4 * g]),
                       "r"(pp[sub][4 * g + 1]), "r"(pp[sub][4 * g + 2]), "r"(pp[sub][4 * g + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_base + kAbSmemDS + a), "r"(dd[sub][4 * g]),
                       "r"(dd[sub][4 * g + 1]), "r"(dd[sub][4 * g + 2]), "r"(dd[sub][4 * g + 3]) : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      pending = cur;
      pend_dq = needs_drain(cur);
      pend_kv = (cur.i == NT - 1);
      cur = nxt;
    }
    if (n_pair > 0) {   // results of the last pair
      mbar_wait(pv_done, (n_pair - 1) & 1u);
      tc_fence_after();
      if (pend_dq) drain_dq(pending);
      if (pend_kv) drain_kv(pending);
    }
    if (stage_pending) {
      if (lane == 0) tma_store_wait_all<0>();
      __syncwarp();
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace mt

extern "C" int mt_attention_bwd_workspace_bytes(int B, int N, int H, size_t* bytes) {
  using namespace mt;
  MT_REQUIRE(bytes && B > 0 && N > 0 && H > 0, MT_ERR_BAD_SHAPE, "mt_attention_bwd_workspace_bytes: bad arguments");
  *bytes = size_t(B) * N * H * 64 * 4 + size_t(B) * H * N * 4 + 256;
  return MT_OK;
}

static int attention_bwd_impl(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                              void* workspace, size_t workspace_bytes, int B, int N, int H, int dh, float scale,
                              float drop_p, unsigned long long seed, void* stream) {
  using namespace mt;
  MT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MT_ERR_BAD_SHAPE, "mt_attention_bwd: drop_p must be in [0, 1)");
  MT_REQUIRE(drop_p == 0.f || (long long)N * N < (1ll << 32), MT_ERR_UNSUPPORTED, "mt_attention_bwd: dropout needs N < 65536");
  MT_REQUIRE(qkv && out && dout && lse2 && dqkv && workspace, MT_ERR_BAD_SHAPE, "mt_attention_bwd: null pointer");
  MT_REQUIRE(B > 0 && N > 0 && H > 0, MT_ERR_BAD_SHAPE, "mt_attention_bwd: B, N, H must be positive");
  MT_REQUIRE(dh == 64, MT_ERR_UNSUPPORTED, "mt_attention_bwd: head_dim %d not supported (only 64)", dh);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
               reinterpret_cast<uintptr_t>(dqkv) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
             MT_ERR_BAD_ALIGN, "mt_attention_bwd: pointers must be 16-byte aligned");
  MT_REQUIRE((reinterpret_cast<uintptr_t>(dqkv) & 31) == 0, MT_ERR_BAD_ALIGN, "mt_attention_bwd: dqkv must be 32-byte aligned");
  const size_t dq_bytes = size_t(B) * N * H * 64 * 4;
  const size_t need = dq_bytes + size_t(B) * H * N * 4 + 256;
  MT_REQUIRE(workspace_bytes >= need, MT_ERR_BAD_SHAPE, "mt_attention_bwd: workspace too small (%zu < %zu)",
             workspace_bytes, need);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* dq_acc = reinterpret_cast<float*>(workspace);
  float* delta = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((dq_bytes + 255) / 256) * 256);
  {
    const long long rows = (long long)B * N;
    attn_bwd_delta_kernel<<<unsigned((rows + 7) / 8), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(out),
                                                                  reinterpret_cast<const __nv_bfloat16*>(dout), delta,
                                                                  B, N, H);
    MT_CHECK_CUDA(cudaGetLastError());
  }
  CUtensorMap tm_qkv, tm_do;
  {
    uint64_t dims[4] = {64, uint64_t(3 * H), uint64_t(N), uint64_t(B)};
    uint64_t strides[3] = {64 * 2, uint64_t(3 * H) * 64 * 2, uint64_t(N) * 3 * H * 64 * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = encode_tmap(&tm_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, qkv, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {64, uint64_t(H), uint64_t(N), uint64_t(B)};
    uint64_t strides[3] = {64 * 2, uint64_t(H) * 64 * 2, uint64_t(N) * H * 64 * 2};
    uint32_t box[4] = {64, 1, 128, 1};
    int rc = encode_tmap(&tm_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dout, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnBwdParams p;
  p.B = B; p.N = N; p.H = H;
  p.num_t = ceil_div(N, 128);
  p.num_bh = B * H;
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  {
    // owner mode needs enough (batch, head) pairs to balance the SMs (C2: 20.8 per SM, C3: 13.8); long sequences with
    // few heads (C5: 3.5 per SM) split the key tiles over CTAs instead.  MT_B200_ATTN_BWD_OWNER=0|1 forces one (A/B runs).
    static int forced = -2;
    if (forced == -2) {
      const char* e = getenv("MT_B200_ATTN_BWD_OWNER");
      forced = (e && (e[0] == '0' || e[0] == '1')) ? (e[0] - '0') : -1;
    }
    p.owner = forced >= 0 ? forced : (p.num_bh >= 8 * sms || p.num_t == 1) ? 1 : 0;
  }
  p.resident = (p.owner && p.num_t <= 2) ? 1 : 0;

  p.num_work = p.owner ? p.num_bh : p.num_bh * p.num_t;
  if (!p.resident) MT_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, dq_bytes, s));
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.delta = delta;
  p.dq_acc = dq_acc;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.drop_thresh = drop_p > 0.f ? uint32_t(double(drop_p) * 4294967296.0) : 0u;
  p.drop_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.seed_lo = uint32_t(seed);
  p.seed_hi = uint32_t(seed >> 32);
  const int grid = p.num_work < sms ? p.num_work : sms;
  {
    // the opt-in to > 48 KB of dynamic shared memory is per device (nn.DataParallel drives several from one process)
    static bool attr_set[64] = {false};
    int dev = 0;
    MT_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes));
      MT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  CUtensorMap tm_dq;
  {
    // fp32 dQ accumulator [B, N, H, 64]: one warp's partial = 32 query rows x 32 of the 64 columns (128 B rows)
    uint64_t dims[4] = {64, uint64_t(H), uint64_t(N), uint64_t(B)};
    uint64_t strides[3] = {64 * 4, uint64_t(H) * 64 * 4, uint64_t(N) * H * 64 * 4};
    uint32_t box[4] = {32, 1, 32, 1};
    int rc2 = encode_tmap(&tm_dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc2) return rc2;
  }
  if (p.drop_thresh != 0u) attention_bwd_kernel<true><<<unsigned(grid), kAbThreads, kAbSmemBytes, s>>>(tm_qkv, tm_do, tm_dq, p);
  else attention_bwd_kernel<false><<<unsigned(grid), kAbThreads, kAbSmemBytes, s>>>(tm_qkv, tm_do, tm_dq, p);
  MT_CHECK_CUDA(cudaGetLastError());
  if (!p.resident) {
    const long long rows = (long long)B * N;
    long long blocks = (rows * H * 8 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    attn_bwd_dq_store_kernel<<<unsigned(blocks), 256, 0, s>>>(dq_acc, reinterpret_cast<__nv_bfloat16*>(dqkv), rows, H);
    MT_CHECK_CUDA(cudaGetLastError());
  }
  return MT_OK;
}

extern "C" int mt_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                                void* workspace, size_t workspace_bytes, int B, int N, int H, int dh, float scale,
                                void* stream) {
  return attention_bwd_impl(qkv, out, dout, lse2, dqkv, workspace, workspace_bytes, B, N, H, dh, scale, 0.f, 0ull, stream);
}

// backward of mt_attention_fwd_ex: same (drop_p, seed) => same keep mask
extern "C" int mt_attention_bwd_ex(const void* qkv, const void* out, const void* dout, const float* lse2, void* dqkv,
                                   void* workspace, size_t workspace_bytes, int B, int N, int H, int dh, float scale,
                                   float drop_p, unsigned long long seed, void* stream) {
  return attention_bwd_impl(qkv, out, dout, lse2, dqkv, workspace, workspace_bytes, B, N, H, dh, scale, drop_p, seed, stream);
}
