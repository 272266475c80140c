// tcgen05 / TMEM / TMA flash-attention forward (head_dim 128 or 80) over the packed token axis of attention.cu.
//
// Same masking model and call sites as t1_attn_fwd (LLM prefill / packed GRPO scoring, HF Q25:740-755; and the ViT's
// windowed / per-slot full attention, head_dim 80, non-causal [lo, hi) row ranges, HF Q25:244-261,498-509 / Q2:771-790 —
// both reached from src/time_r1/rl/timer1_trainer.py:452-457,569-573) — the Blackwell-native replacement of the HMMA
// kernel.  head_dim 80 = 5 UMMA K-steps of 16 for S = Q K^T and an N = 80 accumulator for O += P V; the Q / K / V tiles keep
// the [128 x 128] two-atom shared-memory layout (the second 64-column TMA box of a head overlaps its neighbour or is
// zero-filled past the last head; only the first 80 columns are ever addressed by an MMA).
// One CTA = 128 query rows x one q head, looping over 128-key tiles:
//   warp 0       TMA producer : Q tile once, K/V tiles (2 x [128 x 64] SWIZZLE_128B boxes each) into a 2-stage ring
//   warp 1       MMA issuer   : S = Q K^T   (tcgen05.mma M128 N128 K16 x 8, both operands K-major) -> TMEM S[it & 1]
//                               O += P V    (A = P read from TMEM — the bf16 probabilities are packed back into the S
//                               columns they came from; B = V tile consumed in place as an MN-major operand) -> TMEM O
//   warps 2..17  softmax      : 4 threads per query row (32 key columns each, so every SM sub-partition has four
//                               resident softmax warps); tcgen05.ld S -> mask, scale, lazy-rescaled online softmax (the
//                               running max is only refreshed when it grows by > 8 in log2 units, so the O accumulator
//                               in TMEM is rarely rescaled) -> P (bf16) packed into the first half of each thread's
//                               own 32 S columns (no shared-memory round trip, no cross-thread hazard)
// S(it+1) is issued before softmax(it) finishes and the exponentials of tile it do not wait for PV(it-1): only a
// (rare) rescale of O does, so QK^T, the exponentials and PV of neighbouring tiles overlap.
#include "t1_common.cuh"
#ifdef T1_SOFTMAX_POLY_EXP2
#include "fast_exp2.cuh"
#endif
#include <limits.h>
#include <stdlib.h>

namespace {

constexpr int TQ = 128;   // query rows per CTA
constexpr int TK = 128;   // keys per tile
constexpr int ATOM_BYTES = 128 * 128;            // [128 rows x 64 bf16] box
constexpr int TILE_BYTES = 2 * ATOM_BYTES;       // [128 x 128] bf16
constexpr int NS = 4;                  // softmax threads per query row
constexpr int FA_TC_THREADS = 64 + 128 * NS;   // TMA warp, MMA warp, 4*NS softmax warps
constexpr int FA_TC_SMEM = TILE_BYTES /*Q*/ + 2 * TILE_BYTES /*K*/ + 2 * TILE_BYTES /*V*/ + 1024 + 256 + 3 * NS * 128 * 4 /*row exchange*/;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Experiment switch (T1_EXTRA_NVCC=-DT1_SOFTMAX_POLY_EXP2): every second exponential of a score pair on the FMA pipe
// (fast_exp2.cuh, relative error 7.5e-5, far inside bf16) instead of the MUFU, which the softmax saturates.  Default: MUFU.
#ifdef T1_SOFTMAX_POLY_EXP2
#define EX2_B(x) fast_exp2::exp2<3>(x)
#else
#define EX2_B(x) ex2_approx(x)
#endif

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

struct Range {
  int tA_end, tB_start, tB_end;
  __device__ __forceinline__ int count() const { return tA_end + max(0, tB_end - tB_start); }
  __device__ __forceinline__ int tile(int idx) const { return idx < tA_end ? idx : tB_start + (idx - tA_end); }
};

template <int HD>
__global__ void __launch_bounds__(FA_TC_THREADS, 1)   // 576 threads: 96 registers is the most that still launches
fa_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, bf16* __restrict__ O, long long ldo, float* __restrict__ LSE2, long long lse_ld,
                 const int* __restrict__ row_lo, const int* __restrict__ row_hi, int shared_len, int T, int q_row0,
                 int group, float scale_log2, uint32_t v_lbo, uint32_t v_sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;            // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;        // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * TILE_BYTES);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // 2
  uint64_t* k_empty = bars + 3;             // 2  (released as soon as S = Q K^T of the tile retires)
  uint64_t* v_full = bars + 5;              // 2
  uint64_t* v_empty = bars + 7;             // 2  (released when O += P V retires)
  uint64_t* s_full = bars + 9;              // 2
  // p_full is double-buffered by tile parity: a softmax warp may run one tile ahead of the slowest one (S(it+1) is issued
  // before p_full(it) completes), and its early arrival must not count towards the previous tile's phase
  uint64_t* p_full = bars + 11;             // 2 (one arrival per softmax warp)
  uint64_t* pv_done = bars + 13;            // 2 (PV(it) arrives on pv_done[it & 1])
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);
  int* s_red = reinterpret_cast<int*>(bars + 16);  // 4 ints
  float* s_xch = reinterpret_cast<float*>(bars + 18);   // [3*NS][128] floats: per-row partial max (double buffered) and sum

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Q / O / the tensor map of Q are LOCAL to the processed query rows [q_row0, T); row_lo / row_hi / LSE2 and the key
  // axis use absolute packed-row indices (lets the scoring pass skip the prompt rows whose K/V come from the cache).
  // CTAs are dispatched x-fastest: walk (query tile, head) with the HEAD fastest and the LONGEST rows (last tiles of the packed
  // causal layout) first, so the heavy tiles of all heads start early and the light ones fill the tail, and the q heads of a KV
  // group stream the same K/V tiles side by side (was tile-fastest per head: ncu elapsed / active IPC 0.79)
  const int lin = (int)(blockIdx.x + gridDim.x * blockIdx.y), Hq_ = (int)gridDim.y;
  const int q0l = ((int)gridDim.x - 1 - lin / Hq_) * TQ;           // local row of this CTA's first query
  const int q0 = q_row0 + q0l;               // absolute
  const int h = lin % Hq_, hk = h / group;

  // ---- tile range from the rows' (lo, hi): threads 64..191 (softmax warps) own one row each ----
  if (threadIdx.x == 0) {
    s_red[0] = INT_MAX; s_red[1] = INT_MIN; s_red[2] = INT_MAX; s_red[3] = INT_MIN;
  }
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(&p_full[0], 4 * NS);         // one arrival per softmax warp
    mbar_init(&p_full[1], 4 * NS);
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  __syncthreads();
  int my_row = -1, my_lo = 0, my_hi = 0;
  if (warp_idx >= 2) {
    const int quad = warp_idx & 3;
    my_row = q0 + quad * 32 + lane;
    if (my_row < T) {
      my_lo = row_lo[my_row];
      my_hi = row_hi[my_row];
      atomicMin(&s_red[0], my_lo); atomicMax(&s_red[1], my_lo);
      atomicMin(&s_red[2], my_hi); atomicMax(&s_red[3], my_hi);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int min_lo = s_red[0], max_lo = s_red[1], min_hi = s_red[2], max_hi = s_red[3];
  Range tr;
  if (max_hi <= 0 || min_lo == INT_MAX) {
    tr.tA_end = tr.tB_start = tr.tB_end = 0;
  } else {
    const int a_end = min(min(shared_len, min_lo), max_hi);
    tr.tA_end = (a_end + TK - 1) / TK;
    tr.tB_start = max(min_lo / TK, tr.tA_end);
    tr.tB_end = (max_hi + TK - 1) / TK;
  }
  const int ntiles = tr.count();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && ntiles > 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(sQ, &tm_q, q_full, h * HD, q0l);
      tma_load_2d(sQ + ATOM_BYTES, &tm_q, q_full, h * HD + 64, q0l);
      for (int it = 0; it < ntiles; ++it) {
        const int st = it & 1;
        const int j0 = tr.tile(it) * TK;
        mbar_wait(&k_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(sK + st * TILE_BYTES, &tm_k, &k_full[st], hk * HD, j0);
        tma_load_2d(sK + st * TILE_BYTES + ATOM_BYTES, &tm_k, &k_full[st], hk * HD + 64, j0);
        mbar_wait(&v_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_2d(sV + st * TILE_BYTES, &tm_v, &v_full[st], hk * HD, j0);
        tma_load_2d(sV + st * TILE_BYTES + ATOM_BYTES, &tm_v, &v_full[st], hk * HD + 64, j0);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK);                   // both K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(TQ, HD) | (1u << 16);     // B (= V) MN-major
      const uint32_t tmem_o = tmem_base + 256;
      auto issue_s = [&](int it) {
        const int st = it & 1;
        mbar_wait(&k_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_s = tmem_base + (it & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
          umma_bf16_ss(tmem_s, umma_desc_sw128_kmajor(smem_u32(sQ + off)),
                       umma_desc_sw128_kmajor(smem_u32(sK + st * TILE_BYTES + off)), idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[it & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) issue_s(it + 1);
        const int st = it & 1;
        mbar_wait(&v_full[st], (it >> 1) & 1);
        mbar_wait(&p_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tmem_p = tmem_base + (it & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          // thread `part` packed its 32 probabilities into the first 16 of its own S columns: K-step kk -> part kk / 2
          const uint64_t b_desc = umma_desc_sw128_mnmajor(smem_u32(sV + st * TILE_BYTES + kk * 16 * 128), v_lbo, v_sbo);
          umma_bf16_ts(tmem_o, tmem_p + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), b_desc, idesc_pv, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&pv_done[it & 1]);
      }
    }
  } else {
    // ===================== softmax warps: NS threads per query row, 128/NS key columns each =====================
    constexpr int CW = TK / NS;                         // key columns per thread; O columns [part * 32, part * 32 + 32) n [0, HD)
    constexpr int NCH = CW / 32;                        // 32-column TMEM chunks per thread
    const int quad = warp_idx & 3;                      // TMEM lane quadrant this warp may touch
    const int part = (warp_idx - 2) >> 2;               // which column slice
    const int r = quad * 32 + lane;                     // row inside the tile == TMEM lane
    const bool valid = my_row < T;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tmem_o = t_lane + 256 + part * CW;
    const bool o_on = part * CW < HD;                   // warp-uniform: head_dim 80 leaves the last column slice empty
    const int o_w = HD - part * CW < CW ? HD - part * CW : CW;   // valid output columns of this slice (80: 32, 32, 16, -)
    float m_ref = -INFINITY, l = 0.f;
    for (int it = 0; it < ntiles; ++it) {
      const int j0 = tr.tile(it) * TK, j1 = j0 + TK;
      mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      uint32_t sv[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(t_lane + (it & 1) * 128 + part * CW + c * 32, sv[c]);
      tmem_ld_wait();
      const bool need_mask = !(j1 <= min_hi && (j0 >= max_lo || j1 <= shared_len));
      float tmax = -INFINITY;   // raw (unscaled) scores; scale_log2 > 0 so max commutes with the scaling
      if (need_mask) {          // warp-uniform: the common unmasked tile runs a branch-free max (ncu: the per-element
                                // branch of the fused form was 42% of all executed instructions)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int key = j0 + part * CW + c * 32 + j;
            const bool vis = valid && key < my_hi && (key >= my_lo || key < shared_len);
            const float v = vis ? __uint_as_float(sv[c][j]) : -INFINITY;
            sv[c][j] = __float_as_uint(v);
            tmax = fmaxf(tmax, v);
          }
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(sv[c][j]));
      }
      // the NS threads of a row exchange their partial maxima (named barrier per TMEM quadrant, 32*NS threads)
      s_xch[((it & 1) * NS + part) * TQ + r] = tmax;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
#pragma unroll
      for (int o = 0; o < NS; ++o) tmax = fmaxf(tmax, s_xch[((it & 1) * NS + o) * TQ + r]);
      tmax *= scale_log2;
      // lazy rescale: refresh the reference max only when it grows by more than 8 (factor 256)
      const bool grow = tmax > m_ref + 8.0f;
      float alpha = 1.0f;
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - tmax);
        m_ref = tmax;
        l *= alpha;
      }
      if (it > 0 && __any_sync(0xffffffffu, grow) && o_on) {
        // rare after the first tiles: O is rescaled in place once PV(it-1) has retired (columns past head_dim inside the
        // slice hold whatever TMEM held: scaled and written back, never read)
        mbar_wait(&pv_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tmem_o + c * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
          tmem_st_32x32b_x32(tmem_o + c * 32, ov);
        }
        tmem_st_wait();
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float rs = 0.f;
      static_assert(NCH == 1, "one 32-column chunk per softmax thread");
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[0][2 * j]), scale_log2, neg_m));
        const float p1 = EX2_B(fmaf(__uint_as_float(sv[0][2 * j + 1]), scale_log2, neg_m));
        rs += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      l += rs;
      // P(it) goes into the S columns P(it-2) was read from: PV(it-2) must have retired (it has, two tiles ago)
      if (it >= 2) mbar_wait(&pv_done[it & 1], ((it >> 1) - 1) & 1);
      tmem_st_32x32b_x16(t_lane + (it & 1) * 128 + part * CW, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[it & 1]);     // 16 arrivals per tile instead of 512 same-address shared-memory atomics
    }
    // ---- epilogue: combine the partial row sums, normalise, store this thread's output columns ----
    s_xch[(2 * NS + part) * TQ + r] = l;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
    l = 0.f;
#pragma unroll
    for (int o = 0; o < NS; ++o) l += s_xch[(2 * NS + o) * TQ + r];
    if (ntiles > 0) {
      mbar_wait(&pv_done[(ntiles - 1) & 1], ((ntiles - 1) >> 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    if (valid && LSE2 && part == 0) LSE2[(long long)h * lse_ld + my_row] = l > 0.f ? m_ref + log2f(l) : -INFINITY;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      uint32_t ov[32];
      if (ntiles > 0 && o_on) {
        tmem_ld_32x32b_x32(tmem_o + c * 32, ov);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = 0;
      }
      if (valid && o_on) {
        bf16* op = O + (long long)(my_row - q_row0) * ldo + (long long)h * HD + part * CW + c * 32;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          if (j8 * 8 >= o_w) break;
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 0]) * inv, __uint_as_float(ov[j8 * 8 + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 2]) * inv, __uint_as_float(ov[j8 * 8 + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 4]) * inv, __uint_as_float(ov[j8 * 8 + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 6]) * inv, __uint_as_float(ov[j8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(op + j8 * 8) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =====================================================================================================================
// Persistent variant for BLOCK-LOCAL masks: every 128-row query tile attends only to keys of the SAME 128-row tile (the ViT's
// 64-token windows, HF Q25:244-261 with cu_window_seqlens; checked on the host from the window plan).  The general kernel
// above spends one CTA — barrier init, TMEM allocation, three TMA loads, one half-masked tile, epilogue, teardown: ~7.7 us of
// mostly latency — per (tile, head).  Here a CTA is resident for the whole launch and walks (tile, head) items: Q/K/V of item
// i+1 and i+2 are in flight (2-stage ring) while S(i+1) = QK^T runs under the softmax of item i, and the normalise-and-store
// of item i-1 is deferred behind the softmax of item i, so no thread ever waits for the tensor pipe.
// =====================================================================================================================
constexpr int WIN_STAGE_BYTES = 3 * TILE_BYTES;                      // Q, K, V of one item
constexpr int FA_WIN_SMEM = 2 * WIN_STAGE_BYTES + 1024 + 256 + 4 * NS * 128 * 4;

template <int HD>
__global__ void __launch_bounds__(FA_TC_THREADS, 1)
fa_fwd_win_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, bf16* __restrict__ O, long long ldo, const int* __restrict__ row_lo,
                     const int* __restrict__ row_hi, int T, int H, float scale_log2, uint32_t v_lbo, uint32_t v_sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem;                                               // [2][Q | K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 2 * WIN_STAGE_BYTES);
  uint64_t* full = bars;           // 2  TMA -> MMA: Q, K, V of the stage landed
  uint64_t* empty = bars + 2;      // 2  PV of the item retired: the stage may be refilled
  uint64_t* s_full = bars + 4;     // 2
  uint64_t* p_full = bars + 6;     // 2  (one arrival per softmax warp)
  uint64_t* o_full = bars + 8;     // 2  PV retired: O of the item is complete
  uint64_t* o_free = bars + 10;    // 2  (one arrival per softmax warp) the epilogue has read O
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
  float* s_xch = reinterpret_cast<float*>(bars + 16);   // [2][NS][128] partial row maxima
  float* s_lx = s_xch + 2 * NS * 128;                   // [2][NS][128] partial row sums

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (T + TQ - 1) / TQ;
  const int n_items = n_qt * H;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = first < n_items ? (n_items - first + stride - 1) / stride : 0;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4 * NS);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 4 * NS);
    }
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;      // S[2] at columns 0 / 128, O[2] at 256 / 384

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < n_mine; ++it) {
        const int item = first + it * stride, qt = item / H, h = item - qt * H;
        const int st = it & 1, u = it >> 1;
        uint8_t* sQ = stg + st * WIN_STAGE_BYTES;
        mbar_wait(&empty[st], (u & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], WIN_STAGE_BYTES);
        tma_load_2d(sQ, &tm_q, &full[st], h * HD, qt * TQ);
        tma_load_2d(sQ + ATOM_BYTES, &tm_q, &full[st], h * HD + 64, qt * TQ);
        tma_load_2d(sQ + TILE_BYTES, &tm_k, &full[st], h * HD, qt * TQ);
        tma_load_2d(sQ + TILE_BYTES + ATOM_BYTES, &tm_k, &full[st], h * HD + 64, qt * TQ);
        tma_load_2d(sQ + 2 * TILE_BYTES, &tm_v, &full[st], h * HD, qt * TQ);
        tma_load_2d(sQ + 2 * TILE_BYTES + ATOM_BYTES, &tm_v, &full[st], h * HD + 64, qt * TQ);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_mine > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(TQ, TK);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(TQ, HD) | (1u << 16);     // B (= V) MN-major
      auto issue_s = [&](int it) {
        const int st = it & 1, u = it >> 1;
        const uint8_t* sQ = stg + st * WIN_STAGE_BYTES;
        mbar_wait(&full[st], u & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
          umma_bf16_ss(tmem_base + st * 128, umma_desc_sw128_kmajor(smem_u32(sQ + off)),
                       umma_desc_sw128_kmajor(smem_u32(sQ + TILE_BYTES + off)), idesc_s, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
      };
      issue_s(0);
      for (int it = 0; it < n_mine; ++it) {
        if (it + 1 < n_mine) issue_s(it + 1);
        const int st = it & 1, u = it >> 1;
        mbar_wait(&p_full[st], u & 1);
        mbar_wait(&o_free[st], (u & 1) ^ 1);        // the epilogue of item it-2 has read this O buffer
        tc_fence_after();
        const uint8_t* sV = stg + st * WIN_STAGE_BYTES + 2 * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint64_t b_desc = umma_desc_sw128_mnmajor(smem_u32(sV + kk * 16 * 128), v_lbo, v_sbo);
          umma_bf16_ts(tmem_base + 256 + st * 128, tmem_base + st * 128 + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), b_desc, idesc_pv,
                       kk != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[st]);
        umma_commit(&empty[st]);
      }
    }
  } else {
    // ===================== softmax warps (NS threads per row) + deferred epilogue =====================
    const int quad = warp_idx & 3, part = (warp_idx - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool o_on = part * 32 < HD;
    const int o_w = HD - part * 32 < 32 ? HD - part * 32 : 32;
    auto epilogue = [&](int j) {
      const int item = first + j * stride, qt = item / H, h = item - qt * H;
      const int st = j & 1, u = j >> 1;
      const int row = qt * TQ + r;
      float l = 0.f;
#pragma unroll
      for (int o = 0; o < NS; ++o) l += s_lx[(st * NS + o) * TQ + r];
      const float inv = l > 0.f ? 1.f / l : 0.f;
      mbar_wait(&o_full[st], u & 1);
      tc_fence_after();
      if (o_on) {
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + 256 + st * 128 + part * 32, ov);
        tmem_ld_wait();
        if (row < T) {
          bf16* op = O + (long long)row * ldo + (long long)h * HD + part * 32;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (j8 * 8 >= o_w) break;
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 0]) * inv, __uint_as_float(ov[j8 * 8 + 1]) * inv);
            pk.y = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 2]) * inv, __uint_as_float(ov[j8 * 8 + 3]) * inv);
            pk.z = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 4]) * inv, __uint_as_float(ov[j8 * 8 + 5]) * inv);
            pk.w = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 6]) * inv, __uint_as_float(ov[j8 * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(op + j8 * 8) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[st]);
    };
    for (int it = 0; it < n_mine; ++it) {
      const int item = first + it * stride, qt = item / H;
      const int st = it & 1, u = it >> 1;
      const int q0 = qt * TQ, row = q0 + r;
      const bool valid = row < T;
      const int my_lo = valid ? row_lo[row] : 0, my_hi = valid ? row_hi[row] : 0;
      mbar_wait(&s_full[st], u & 1);
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld_32x32b_x32(t_lane + st * 128 + part * 32, sv);
      tmem_ld_wait();
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = q0 + part * 32 + j;
        const float v = (key >= my_lo && key < my_hi) ? __uint_as_float(sv[j]) : -INFINITY;
        sv[j] = __float_as_uint(v);
        tmax = fmaxf(tmax, v);
      }
      s_xch[(st * NS + part) * TQ + r] = tmax;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
#pragma unroll
      for (int o = 0; o < NS; ++o) tmax = fmaxf(tmax, s_xch[(st * NS + o) * TQ + r]);
      const float neg_m = (tmax == -INFINITY) ? 0.f : -tmax * scale_log2;
      float rs = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * j]), scale_log2, neg_m));
        const float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), scale_log2, neg_m));
        rs += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      tmem_st_32x32b_x16(t_lane + st * 128 + part * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
      s_lx[(st * NS + part) * TQ + r] = rs;        // read by epilogue(it), one iteration (and at least one bar.sync) later
      if (it > 0) epilogue(it - 1);
    }
    if (n_mine > 0) {
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
      epilogue(n_mine - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// Same contract as t1_attn_fwd, head_dim 128 (LLM) or 80 (ViT); additionally only query rows [q_row0, T) are processed and Q / O
// point at row q_row0 (row_lo / row_hi / LSE2 / K / V keep absolute packed-row indexing; LSE2 is [Hq, lse_ld], lse_ld <= 0 -> T); Q/K/V token strides must be multiples of 8 and the head
// column offsets multiples of 64 (true for the fused qkv layout).
extern "C" int t1_attn_fwd_tc(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                              void* O, long long ldo, float* LSE2, long long lse_ld, const int* row_lo, const int* row_hi,
                              int shared_len, int T, int Tk, int Hq, int Hkv, int head_dim, float scale, int q_row0,
                              cudaStream_t stream) {
  if (T - q_row0 <= 0) return T1_OK;
  T1_CHECK_ARG(q_row0 >= 0, "t1_attn_fwd_tc: q_row0 must be >= 0");
  T1_CHECK_ARG(head_dim == 128 || head_dim == 80, "t1_attn_fwd_tc: head_dim must be 128 or 80 (got %d)", head_dim);
  const int HD = head_dim;
  T1_CHECK_ARG(Q && K && V && O && row_lo && row_hi, "t1_attn_fwd_tc: null pointer");
  T1_CHECK_ARG(Hkv > 0 && Hq % Hkv == 0, "t1_attn_fwd_tc: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  T1_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "t1_attn_fwd_tc: strides must be multiples of 8");
  T1_CHECK_ARG((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) & 15) == 0, "t1_attn_fwd_tc: 16-byte alignment required");
  CUtensorMap tq, tk, tv;
  int rc = t1_make_tmap_bf16(&tq, Q, T - q_row0, (long long)Hq * HD, ldq, TQ, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tk, K, Tk, (long long)Hkv * HD, ldk, TK, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tv, V, Tk, (long long)Hkv * HD, ldv, TK, 64);
  if (rc) return rc;
  static bool set = false;
  static uint32_t lbo = ATOM_BYTES, sbo = 1024;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_TC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fa_fwd_tc_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_TC_SMEM);
    if (e != cudaSuccess) {
      t1_set_error("t1_attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    if (getenv("T1_FA_LBO")) lbo = (uint32_t)atoi(getenv("T1_FA_LBO"));   // descriptor bring-up hooks
    if (getenv("T1_FA_SBO")) sbo = (uint32_t)atoi(getenv("T1_FA_SBO"));
    set = true;
  }
  dim3 grid((T - q_row0 + TQ - 1) / TQ, Hq);
  if (HD == 128)
    fa_fwd_tc_kernel<128><<<grid, FA_TC_THREADS, FA_TC_SMEM, stream>>>(tq, tk, tv, (bf16*)O, ldo, LSE2, lse_ld > 0 ? lse_ld : T, row_lo, row_hi,
                                                                      shared_len, T, q_row0, Hq / Hkv, scale * 1.4426950408889634f, lbo, sbo);
  else
    fa_fwd_tc_kernel<80><<<grid, FA_TC_THREADS, FA_TC_SMEM, stream>>>(tq, tk, tv, (bf16*)O, ldo, LSE2, lse_ld > 0 ? lse_ld : T, row_lo, row_hi,
                                                                     shared_len, T, q_row0, Hq / Hkv, scale * 1.4426950408889634f, lbo, sbo);
  T1_CUDA_LAUNCH_CHECK("t1_attn_fwd_tc");
  return T1_OK;
}

// Block-local variant (see fa_fwd_win_tc_kernel): the caller guarantees that every row's [lo, hi) lies inside the row's own
// 128-row tile (the ViT's 64-token windows; time_r1_b200.indexing.vision_plan checks it).  Q, K, V: [T, >= H * head_dim] views
// (same head count: no GQA in the ViT), no LSE.  head_dim 80 or 128.
extern "C" int t1_attn_fwd_tc_win(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                                  long long ldo, const int* row_lo, const int* row_hi, int T, int H, int head_dim, float scale,
                                  cudaStream_t stream) {
  if (T <= 0) return T1_OK;
  T1_CHECK_ARG(head_dim == 128 || head_dim == 80, "t1_attn_fwd_tc_win: head_dim must be 128 or 80 (got %d)", head_dim);
  T1_CHECK_ARG(Q && K && V && O && row_lo && row_hi && H > 0, "t1_attn_fwd_tc_win: null pointer");
  T1_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "t1_attn_fwd_tc_win: strides must be multiples of 8");
  T1_CHECK_ARG((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O) & 15) == 0, "t1_attn_fwd_tc_win: 16-byte alignment required");
  CUtensorMap tq, tk, tv;
  int rc = t1_make_tmap_bf16(&tq, Q, T, (long long)H * head_dim, ldq, TQ, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tk, K, T, (long long)H * head_dim, ldk, TK, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tv, V, T, (long long)H * head_dim, ldv, TK, 64);
  if (rc) return rc;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_win_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_WIN_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fa_fwd_win_tc_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_WIN_SMEM);
    if (e != cudaSuccess) {
      t1_set_error("t1_attn_fwd_tc_win: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    set = true;
  }
  const int items = ((T + TQ - 1) / TQ) * H;
  const int grid = items < t1_num_sms() ? items : t1_num_sms();
  if (head_dim == 128)
    fa_fwd_win_tc_kernel<128><<<grid, FA_TC_THREADS, FA_WIN_SMEM, stream>>>(tq, tk, tv, (bf16*)O, ldo, row_lo, row_hi, T, H,
                                                                          scale * 1.4426950408889634f, ATOM_BYTES, 1024);
  else
    fa_fwd_win_tc_kernel<80><<<grid, FA_TC_THREADS, FA_WIN_SMEM, stream>>>(tq, tk, tv, (bf16*)O, ldo, row_lo, row_hi, T, H,
                                                                         scale * 1.4426950408889634f, ATOM_BYTES, 1024);
  T1_CUDA_LAUNCH_CHECK("t1_attn_fwd_tc_win");
  return T1_OK;
}
