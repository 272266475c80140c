// tcgen05 / TMEM / TMA flash-attention BACKWARD (head_dim 128) over the packed token axis of attention.cu.
//
// Same masking model, layouts and call site as t1_attn_bwd (autograd through the flash-attn varlen call the reference
// reaches from src/time_r1/rl/timer1_trainer.py:452-457 via HF Q25:740-755, run by HF trainer.py:1927-1934) — this
// is the Blackwell-native replacement of the HMMA backward for the LLM shapes.  Two kernels, no atomics:
//
//   dKV kernel : one CTA per (128-key tile, q head).  TMEM lanes = keys.  Per visible 128-query tile:
//                  S^T  = K Q^T          (SS, both K-major)            -> TMEM [  0,128)
//                  dP^T = V dO^T         (SS)                          -> TMEM [128,256)
//                  P^T  = exp2(S^T*c - LSE2[q])   -> bf16 packed back into the S^T columns (TMEM A operand)
//                  dV  += P^T dO         (TS: A from TMEM, B = dO tile consumed in place as MN-major)   -> TMEM [256,384)
//                  dS^T = P^T o (dP^T - delta[q]) * scale -> bf16 packed into the dP^T columns
//                  dK  += dS^T Q         (TS, B = Q tile MN-major)      -> TMEM [384,512)
//                fp32 per-q-head partials go to a workspace; a small reduce kernel sums the heads of each KV group.
//   dQ kernel  : one CTA per (128-query tile, q head).  TMEM lanes = queries.  Per visible 128-key tile:
//                  S = Q K^T -> TMEM S[it&1];  dP = dO V^T -> TMEM;  dS (bf16) packed into the S columns;
//                  dQ += dS K            (TS, B = K tile MN-major)
// Roles per CTA: warp 0 = TMA producer (+ per-tile LSE/delta/lo/hi staging), warp 1 = single-thread MMA issuer,
// warps 2..17 = 512 element-wise threads, 4 per TMEM lane (32 columns each).  The bf16 P / dS operands never touch
// shared memory: each thread overwrites the first half of the 32 fp32 columns it just read, so there is no
// cross-thread hazard and the in-order tensor pipe guarantees the next S / dP MMA lands after the TS MMA that reads them.
#define T1_MBAR_NO_HINT   // both kernels sit at the 96-register cap of 576-thread CTAs: the hinted wait spills
#include "t1_common.cuh"
#include <limits.h>
#include <type_traits>

namespace {

constexpr int BT = 128;                          // tile edge (queries and keys)
constexpr int HD = 128;
constexpr int ATOM_BYTES = 128 * 128;            // [128 rows x 64 bf16] TMA box
constexpr int TILE_BYTES = 2 * ATOM_BYTES;       // [128 x 128] bf16
constexpr int NS = 4;                            // element-wise threads per TMEM lane
constexpr int BWD_THREADS = 64 + 128 * NS;
constexpr uint32_t MN_LBO = ATOM_BYTES, MN_SBO = 1024;   // MN-major SW128 tile: 64-column atoms 16 KB apart, 8-row groups 1 KB

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[tmem] (+)= A[tmem] * B[smem]; A = 128 lanes x K, bf16 pairs packed in 32-bit columns (8 columns per K=16 step)
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
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
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

// TMEM column (relative to the fp32 region it overlays) of the packed bf16 A operand for K-step kk (16 elements):
// element-wise thread `part` owns elements [32 part, 32 part + 32) and stores them in its first 16 columns.
__device__ __forceinline__ uint32_t packed_col(int kk) { return (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8); }

// ---------------------------------------------------------------------------------------------
// delta[h, i] = sum_d dO[i,h,d] * O[i,h,d];  summ[tile] = (min_lo, max_lo, min_hi, max_hi) over the tile's rows
// ---------------------------------------------------------------------------------------------
__global__ void bwd_delta_kernel(const bf16* __restrict__ O, long long ldo, const bf16* __restrict__ dO, long long lddo,
                                 float* __restrict__ delta, int T, int Hq) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (long long)T * Hq) return;
  const int i = (int)(warp / Hq), h = (int)(warp - (long long)i * Hq);
  const uint2 a = *reinterpret_cast<const uint2*>(O + (long long)i * ldo + h * HD + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(dO + (long long)i * lddo + h * HD + lane * 4);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y);
  float acc = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
  acc = warp_sum(acc);
  if (lane == 0) delta[(long long)h * T + i] = acc;
}

__global__ void bwd_tile_summary_kernel(const int* __restrict__ row_lo, const int* __restrict__ row_hi, int T,
                                        int* __restrict__ summ) {
  __shared__ int s[4];
  if (threadIdx.x == 0) { s[0] = INT_MAX; s[1] = INT_MIN; s[2] = INT_MAX; s[3] = INT_MIN; }
  __syncthreads();
  const int r = blockIdx.x * BT + threadIdx.x;
  if (r < T) {
    const int lo = row_lo[r], hi = row_hi[r];
    atomicMin(&s[0], lo); atomicMax(&s[1], lo); atomicMin(&s[2], hi); atomicMax(&s[3], hi);
  }
  __syncthreads();
  if (threadIdx.x < 4) summ[blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x];
}

// dK / dV = sum over the q heads of each KV group of the fp32 partials, rounded once to bf16
__global__ void bwd_reduce_heads_kernel(const float* __restrict__ part_dk, const float* __restrict__ part_dv,
                                        bf16* __restrict__ dK, long long lddk, bf16* __restrict__ dV, long long lddv,
                                        int Tk, int Hkv, int group) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread = 4 columns
  const long long total = (long long)Tk * Hkv * (HD / 4);
  if (idx >= total) return;
  const int c4 = (int)(idx % (HD / 4));
  const long long t = idx / (HD / 4);
  const int hk = (int)(t % Hkv);
  const int j = (int)(t / Hkv);
  float4 ak = make_float4(0.f, 0.f, 0.f, 0.f), av = ak;
  for (int g = 0; g < group; ++g) {
    const long long off = (((long long)(hk * group + g) * Tk + j) * HD) + c4 * 4;
    const float4 x = *reinterpret_cast<const float4*>(part_dk + off);
    const float4 y = *reinterpret_cast<const float4*>(part_dv + off);
    ak.x += x.x; ak.y += x.y; ak.z += x.z; ak.w += x.w;
    av.x += y.x; av.y += y.y; av.z += y.z; av.w += y.w;
  }
  uint2 pk, pv;
  pk.x = pack_bf16x2(ak.x, ak.y); pk.y = pack_bf16x2(ak.z, ak.w);
  pv.x = pack_bf16x2(av.x, av.y); pv.y = pack_bf16x2(av.z, av.w);
  *reinterpret_cast<uint2*>(dK + (long long)j * lddk + hk * HD + c4 * 4) = pk;
  *reinterpret_cast<uint2*>(dV + (long long)j * lddv + hk * HD + c4 * 4) = pv;
}

// ---------------------------------------------------------------------------------------------
// dKV kernel
// ---------------------------------------------------------------------------------------------
struct AuxStage {
  float neg_lse[BT];
  float delta[BT];
  int lo[BT];
  int hi[BT];
};
constexpr int DKV_SMEM = 2 * TILE_BYTES /*K,V*/ + 4 * TILE_BYTES /*Q,dO x 2 stages*/ + 2 * (int)sizeof(AuxStage) + 256 + 1024;

__device__ __forceinline__ bool q_tile_sees(const int* __restrict__ summ, int qt, int j0, int j1, int shared_len) {
  const int4 s = *reinterpret_cast<const int4*>(summ + qt * 4);     // min_lo, max_lo, min_hi, max_hi
  return s.x != INT_MAX && s.w > j0 && (s.x < j1 || j0 < shared_len);
}

__global__ void __launch_bounds__(BWD_THREADS, 1)   // 576 threads: 96 registers is the most that still launches
fa_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                     const float* __restrict__ LSE2, const float* __restrict__ delta, float* __restrict__ part_dk,
                     float* __restrict__ part_dv, const int* __restrict__ row_lo, const int* __restrict__ row_hi,
                     const int* __restrict__ summ, int shared_len, int T, int Tk, int group, float scale,
                     float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sQ = sV + TILE_BYTES;             // 2 stages
  uint8_t* sdO = sQ + 2 * TILE_BYTES;        // 2 stages
  AuxStage* aux = reinterpret_cast<AuxStage*>(sdO + 2 * TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2);
  uint64_t* kv_full = bars;                  // 1
  uint64_t* qdo_full = bars + 1;             // 2
  uint64_t* qdo_empty = bars + 3;            // 2
  uint64_t* aux_full = bars + 5;             // 2 (32 arrivals)
  uint64_t* s_full = bars + 7;               // 1
  uint64_t* dp_full = bars + 8;              // 1
  uint64_t* p_ready = bars + 9;              // 1 (one arrival per element-wise warp)
  uint64_t* ds_ready = bars + 10;            // 1 (one arrival per element-wise warp)
  uint64_t* acc_done = bars + 11;            // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTAs are dispatched x-fastest: walk (key tile, head) with the HEAD fastest so that the heavy tiles (low key index: seen by
  // every later row) of ALL heads start first and the light ones fill the tail (was: per head, so the last heads' heavy
  // tiles started in the last wave: ncu elapsed / active IPC 0.81); the q heads of a KV group also run side by side
  const int lin = (int)(blockIdx.x + gridDim.x * blockIdx.y), Hq_ = (int)gridDim.y;
  const int jt = lin / Hq_, j0 = jt * BT, j1 = j0 + BT;
  const int h = lin - jt * Hq_, hk = h / group;
  const int nqt = (T + BT - 1) / BT;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1);
      mbar_init(&qdo_empty[i], 1);
      mbar_init(&aux_full[i], 32);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 128 * NS);
    mbar_init(ds_ready, 128 * NS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 384;
  int n = 0;     // query tiles that see this key tile (every role walks the same list)
  for (int qt = 0; qt < nqt; ++qt) n += q_tile_sees(summ, qt, j0, j1, shared_len) ? 1 : 0;

  if (warp_idx == 0) {
    // ===================== producer: TMA tiles + per-tile row statistics =====================
    if (lane == 0 && n > 0) {
      mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_2d(sK, &tm_k, kv_full, hk * HD, j0);
      tma_load_2d(sK + ATOM_BYTES, &tm_k, kv_full, hk * HD + 64, j0);
      tma_load_2d(sV, &tm_v, kv_full, hk * HD, j0);
      tma_load_2d(sV + ATOM_BYTES, &tm_v, kv_full, hk * HD + 64, j0);
    }
    int it = 0;
    for (int qt = 0; qt < nqt; ++qt) {
      if (!q_tile_sees(summ, qt, j0, j1, shared_len)) continue;
      const int st = it & 1;
      mbar_wait(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(&qdo_full[st], 2 * TILE_BYTES);
        tma_load_2d(sQ + st * TILE_BYTES, &tm_q, &qdo_full[st], h * HD, qt * BT);
        tma_load_2d(sQ + st * TILE_BYTES + ATOM_BYTES, &tm_q, &qdo_full[st], h * HD + 64, qt * BT);
        tma_load_2d(sdO + st * TILE_BYTES, &tm_do, &qdo_full[st], h * HD, qt * BT);
        tma_load_2d(sdO + st * TILE_BYTES + ATOM_BYTES, &tm_do, &qdo_full[st], h * HD + 64, qt * BT);
      }
#pragma unroll
      for (int rr = 0; rr < BT / 32; ++rr) {
        const int r = rr * 32 + lane, row = qt * BT + r;
        float nl = -INFINITY, dl = 0.f;
        int lo = 0, hi = 0;
        if (row < T) {
          const float l2 = LSE2[(long long)h * T + row];
          nl = (l2 == -INFINITY) ? -INFINITY : -l2;
          dl = delta[(long long)h * T + row];
          lo = row_lo[row];
          hi = row_hi[row];
        }
        aux[st].neg_lse[r] = nl; aux[st].delta[r] = dl; aux[st].lo[r] = lo; aux[st].hi[r] = hi;
      }
      mbar_arrive(&aux_full[st]);
      ++it;
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (n > 0) {
        constexpr uint32_t idesc_ss = umma_idesc_bf16(BT, BT);                    // A, B both K-major
        constexpr uint32_t idesc_ts = umma_idesc_bf16(BT, HD) | (1u << 16);       // B MN-major
        auto issue_s = [&](int it) {
          const int st = it & 1;
          mbar_wait(&qdo_full[st], (it >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
            umma_bf16_ss(tmem_base + T_S, umma_desc_sw128_kmajor(smem_u32(sK + off)),
                         umma_desc_sw128_kmajor(smem_u32(sQ + st * TILE_BYTES + off)), idesc_ss, ks != 0 ? 1u : 0u);
          }
          umma_commit(s_full);
        };
        auto issue_dp = [&](int it) {
          const int st = it & 1;
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks) {
            const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
            umma_bf16_ss(tmem_base + T_DP, umma_desc_sw128_kmajor(smem_u32(sV + off)),
                         umma_desc_sw128_kmajor(smem_u32(sdO + st * TILE_BYTES + off)), idesc_ss, ks != 0 ? 1u : 0u);
          }
          umma_commit(dp_full);
        };
        mbar_wait(kv_full, 0);
        issue_s(0);
        issue_dp(0);
        for (int it = 0; it < n; ++it) {
          const int st = it & 1;
          mbar_wait(p_ready, it & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk)
            umma_bf16_ts(tmem_base + T_DV, tmem_base + T_S + packed_col(kk),
                         umma_desc_sw128_mnmajor(smem_u32(sdO + st * TILE_BYTES + kk * 16 * 128), MN_LBO, MN_SBO), idesc_ts,
                         (it | kk) != 0 ? 1u : 0u);
          if (it + 1 < n) issue_s(it + 1);
          mbar_wait(ds_ready, it & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk)
            umma_bf16_ts(tmem_base + T_DK, tmem_base + T_DP + packed_col(kk),
                         umma_desc_sw128_mnmajor(smem_u32(sQ + st * TILE_BYTES + kk * 16 * 128), MN_LBO, MN_SBO), idesc_ts,
                         (it | kk) != 0 ? 1u : 0u);
          umma_commit(&qdo_empty[st]);
          if (it + 1 < n) issue_dp(it + 1);
        }
        umma_commit(acc_done);
      }
    }
  } else {
    // ===================== element-wise threads: lane = key, 32 query columns each =====================
    const int quad = warp_idx & 3;
    const int part = (warp_idx - 2) >> 2;
    const int key = j0 + quad * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    int it = 0;
    for (int qt = 0; qt < nqt; ++qt) {
      const int4 sm = *reinterpret_cast<const int4*>(summ + qt * 4);
      if (!(sm.x != INT_MAX && sm.w > j0 && (sm.x < j1 || j0 < shared_len))) continue;
      const int st = it & 1;
      const bool need_mask = !(j1 <= sm.z && (j0 >= sm.y || j1 <= shared_len)) || (qt + 1) * BT > T;
      mbar_wait(&aux_full[st], (it >> 1) & 1);
      const AuxStage& ax = aux[st];
      // ---- P^T ----
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t pv[32];
      tmem_ld_32x32b_x32(t_lane + T_S + part * 32, pv);
      tmem_ld_wait();
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 nl = *reinterpret_cast<const float4*>(&ax.neg_lse[part * 32 + c4 * 4]);
        float p0 = ex2_approx(fmaf(__uint_as_float(pv[c4 * 4 + 0]), scale_log2, nl.x));
        float p1 = ex2_approx(fmaf(__uint_as_float(pv[c4 * 4 + 1]), scale_log2, nl.y));
        float p2 = ex2_approx(fmaf(__uint_as_float(pv[c4 * 4 + 2]), scale_log2, nl.z));
        float p3 = ex2_approx(fmaf(__uint_as_float(pv[c4 * 4 + 3]), scale_log2, nl.w));
        if (need_mask) {      // warp-uniform, one branch per 4 elements
          const int4 lo = *reinterpret_cast<const int4*>(&ax.lo[part * 32 + c4 * 4]);
          const int4 hi = *reinterpret_cast<const int4*>(&ax.hi[part * 32 + c4 * 4]);
          const bool sh = key < shared_len;
          p0 = (key < hi.x && (key >= lo.x || sh)) ? p0 : 0.f;
          p1 = (key < hi.y && (key >= lo.y || sh)) ? p1 : 0.f;
          p2 = (key < hi.z && (key >= lo.z || sh)) ? p2 : 0.f;
          p3 = (key < hi.w && (key >= lo.w || sh)) ? p3 : 0.f;
        }
        pv[c4 * 4 + 0] = __float_as_uint(p0); pv[c4 * 4 + 1] = __float_as_uint(p1);
        pv[c4 * 4 + 2] = __float_as_uint(p2); pv[c4 * 4 + 3] = __float_as_uint(p3);
      }
      {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) pk[c] = pack_bf16x2(__uint_as_float(pv[2 * c]), __uint_as_float(pv[2 * c + 1]));
        tmem_st_32x32b_x16(t_lane + T_S + part * 32, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- dS^T ----
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      uint32_t dk[16];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t dp[16];
        tmem_ld_32x32b_x16(t_lane + T_DP + part * 32 + half * 16, dp);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const float4 dl = *reinterpret_cast<const float4*>(&ax.delta[part * 32 + half * 16 + c4 * 4]);
          const int b = half * 16 + c4 * 4;
          const float d0 = __uint_as_float(pv[b + 0]) * (__uint_as_float(dp[c4 * 4 + 0]) - dl.x) * scale;
          const float d1 = __uint_as_float(pv[b + 1]) * (__uint_as_float(dp[c4 * 4 + 1]) - dl.y) * scale;
          const float d2 = __uint_as_float(pv[b + 2]) * (__uint_as_float(dp[c4 * 4 + 2]) - dl.z) * scale;
          const float d3 = __uint_as_float(pv[b + 3]) * (__uint_as_float(dp[c4 * 4 + 3]) - dl.w) * scale;
          dk[half * 8 + c4 * 2 + 0] = pack_bf16x2(d0, d1);
          dk[half * 8 + c4 * 2 + 1] = pack_bf16x2(d2, d3);
        }
      }
      tmem_st_32x32b_x16(t_lane + T_DP + part * 32, dk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_ready);
      ++it;
    }
    // ---- epilogue: fp32 partials of this q head ----
    if (it > 0) {
      mbar_wait(acc_done, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t ov[32];
      if (it > 0) {
        tmem_ld_32x32b_x32(t_lane + (which == 0 ? T_DV : T_DK) + part * 32, ov);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = 0u;
      }
      if (key < Tk) {
        float* dst = (which == 0 ? part_dv : part_dk) + ((long long)h * Tk + key) * HD + part * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<uint4*>(dst + j4 * 4) = make_uint4(ov[j4 * 4], ov[j4 * 4 + 1], ov[j4 * 4 + 2], ov[j4 * 4 + 3]);
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

// ---------------------------------------------------------------------------------------------
// dQ kernel
// ---------------------------------------------------------------------------------------------
constexpr int KST = 3;   // K stages (a K tile lives from S(it) to dQ(it)); V has one stage (only dP(it) reads it)
constexpr int DQ_SMEM = 2 * TILE_BYTES /*Q,dO*/ + KST * TILE_BYTES + TILE_BYTES /*V*/ + 256 + 1024;

struct Range {
  int tA_end, tB_start, tB_end;
  __device__ __forceinline__ int count() const { return tA_end + max(0, tB_end - tB_start); }
  __device__ __forceinline__ int tile(int idx) const { return idx < tA_end ? idx : tB_start + (idx - tA_end); }
};

__global__ void __launch_bounds__(BWD_THREADS, 1)   // 576 threads: 96 registers is the most that still launches
fa_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const float* __restrict__ LSE2, const float* __restrict__ delta, bf16* __restrict__ dQ, long long lddq,
                    const int* __restrict__ row_lo, const int* __restrict__ row_hi, int shared_len, int T, int group,
                    float scale, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + TILE_BYTES;
  uint8_t* sK = sdO + TILE_BYTES;            // KST stages
  uint8_t* sV = sK + KST * TILE_BYTES;       // 1 stage
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + TILE_BYTES);
  uint64_t* q_full = bars;                   // 1
  uint64_t* k_full = bars + 1;               // KST
  uint64_t* k_empty = bars + 1 + KST;        // KST
  uint64_t* v_full = bars + 1 + 2 * KST;     // 1
  uint64_t* v_empty = v_full + 1;            // 1
  uint64_t* s_full = v_full + 2;             // 2
  uint64_t* dp_full = v_full + 4;            // 1
  uint64_t* dp_taken = v_full + 5;           // 1 (one arrival per element-wise warp)
  // ds_ready is double-buffered by tile parity: dP(it+1) is issued before ds_ready(it) completes, so a fast warp can
  // arrive for tile it+1 while a slow one has not arrived for tile it
  uint64_t* ds_ready = v_full + 6;           // 2 (one arrival per element-wise warp)
  uint64_t* dq_done = v_full + 8;            // 1
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(v_full + 9);
  int* s_red = reinterpret_cast<int*>(v_full + 10);   // 4 ints

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // long rows (most visible keys) first, ACROSS heads: (query tile, head) with the head fastest (see fa_bwd_dkv_tc_kernel)
  const int lin = (int)(blockIdx.x + gridDim.x * blockIdx.y), Hq_ = (int)gridDim.y;
  const int qt = (int)gridDim.x - 1 - lin / Hq_;
  const int q0 = qt * BT;
  const int h = lin % Hq_, hk = h / group;

  if (threadIdx.x == 0) { s_red[0] = INT_MAX; s_red[1] = INT_MIN; s_red[2] = INT_MAX; s_red[3] = INT_MIN; }
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_do);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(dp_full, 1);
    mbar_init(dp_taken, 4 * NS);           // one arrival per element-wise warp
    mbar_init(&ds_ready[0], 4 * NS);
    mbar_init(&ds_ready[1], 4 * NS);
    mbar_init(dq_done, 1);
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
      if (warp_idx < 6) {     // one of the NS threads of each row feeds the tile statistics
        atomicMin(&s_red[0], my_lo); atomicMax(&s_red[1], my_lo);
        atomicMin(&s_red[2], my_hi); atomicMax(&s_red[3], my_hi);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t T_S0 = 0, T_DP = 256, T_DQ = 384;
  const int min_lo = s_red[0], max_lo = s_red[1], min_hi = s_red[2], max_hi = s_red[3];
  Range tr;
  if (max_hi <= 0 || min_lo == INT_MAX) {
    tr.tA_end = tr.tB_start = tr.tB_end = 0;
  } else {
    const int a_end = min(min(shared_len, min_lo), max_hi);
    tr.tA_end = (a_end + BT - 1) / BT;
    tr.tB_start = max(min_lo / BT, tr.tA_end);
    tr.tB_end = (max_hi + BT - 1) / BT;
  }
  const int n = tr.count();

  if (warp_idx == 0) {
    if (lane == 0 && n > 0) {
      mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
      tma_load_2d(sQ, &tm_q, q_full, h * HD, q0);
      tma_load_2d(sQ + ATOM_BYTES, &tm_q, q_full, h * HD + 64, q0);
      tma_load_2d(sdO, &tm_do, q_full, h * HD, q0);
      tma_load_2d(sdO + ATOM_BYTES, &tm_do, q_full, h * HD + 64, q0);
      // K runs up to KST tiles ahead of V (the MMA thread issues S two tiles ahead of dP)
      int itk = 0;
      auto load_k = [&](int i) {
        const int st = i % KST, j0 = tr.tile(i) * BT;
        mbar_wait(&k_empty[st], ((i / KST) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(sK + st * TILE_BYTES, &tm_k, &k_full[st], hk * HD, j0);
        tma_load_2d(sK + st * TILE_BYTES + ATOM_BYTES, &tm_k, &k_full[st], hk * HD + 64, j0);
      };
      for (; itk < min(n, KST - 1); ++itk) load_k(itk);
      for (int it = 0; it < n; ++it) {
        const int j0 = tr.tile(it) * BT;
        mbar_wait(v_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(v_full, TILE_BYTES);
        tma_load_2d(sV, &tm_v, v_full, hk * HD, j0);
        tma_load_2d(sV + ATOM_BYTES, &tm_v, v_full, hk * HD + 64, j0);
        if (itk < n) { load_k(itk); ++itk; }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_ss = umma_idesc_bf16(BT, BT);
      constexpr uint32_t idesc_ts = umma_idesc_bf16(BT, HD) | (1u << 16);
      auto issue_s = [&](int it) {
        const int st = it % KST;
        mbar_wait(&k_full[st], (it / KST) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + T_S0 + (it & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
          umma_bf16_ss(tS, umma_desc_sw128_kmajor(smem_u32(sQ + off)),
                       umma_desc_sw128_kmajor(smem_u32(sK + st * TILE_BYTES + off)), idesc_ss, ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[it & 1]);
      };
      auto issue_dp = [&](int it) {
        mbar_wait(v_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
          umma_bf16_ss(tmem_base + T_DP, umma_desc_sw128_kmajor(smem_u32(sdO + off)),
                       umma_desc_sw128_kmajor(smem_u32(sV + off)), idesc_ss, ks != 0 ? 1u : 0u);
        }
        umma_commit(v_empty);
        umma_commit(dp_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      issue_dp(0);
      if (n > 1) issue_s(1);
      for (int it = 0; it < n; ++it) {
        const int st = it % KST;
        mbar_wait(dp_taken, it & 1);             // every thread holds dP(it) in registers: the dP columns are free
        tc_fence_after();
        if (it + 1 < n) issue_dp(it + 1);
        mbar_wait(&ds_ready[it & 1], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t tS = tmem_base + T_S0 + (it & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < BT / 16; ++kk)
          umma_bf16_ts(tmem_base + T_DQ, tS + packed_col(kk),
                       umma_desc_sw128_mnmajor(smem_u32(sK + st * TILE_BYTES + kk * 16 * 128), MN_LBO, MN_SBO), idesc_ts,
                       (it | kk) != 0 ? 1u : 0u);
        umma_commit(&k_empty[st]);
        if (it + 2 < n) issue_s(it + 2);
      }
      umma_commit(dq_done);
    }
  } else {
    // ===================== element-wise threads: lane = query row, 32 key columns each =====================
    const int quad = warp_idx & 3;
    const int part = (warp_idx - 2) >> 2;
    const bool valid = my_row < T;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    float nl = -INFINITY, dl = 0.f;
    if (valid) {
      const float l2 = LSE2[(long long)h * T + my_row];
      nl = (l2 == -INFINITY) ? -INFINITY : -l2;
      dl = delta[(long long)h * T + my_row];
    }
    for (int it = 0; it < n; ++it) {
      const int j0 = tr.tile(it) * BT, j1 = j0 + BT;
      const bool need_mask = !(j1 <= min_hi && (j0 >= max_lo || j1 <= shared_len));
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      uint32_t dp[32];
      tmem_ld_32x32b_x32(t_lane + T_DP + part * 32, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dp_taken);
      mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = t_lane + T_S0 + (it & 1) * 128 + part * 32;
      uint32_t pk[16];
      auto ds_phase = [&](auto masked) {      // need_mask is warp-uniform
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sv[16];
          tmem_ld_32x32b_x16(tS + half * 16, sv);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(sv[c]), scale_log2, nl));
            float p1 = ex2_approx(fmaf(__uint_as_float(sv[c + 1]), scale_log2, nl));
            if constexpr (decltype(masked)::value) {
              const int k0 = j0 + part * 32 + half * 16 + c;
              p0 = (valid && k0 < my_hi && (k0 >= my_lo || k0 < shared_len)) ? p0 : 0.f;
              p1 = (valid && k0 + 1 < my_hi && (k0 + 1 >= my_lo || k0 + 1 < shared_len)) ? p1 : 0.f;
            }
            const float d0 = p0 * (__uint_as_float(dp[half * 16 + c]) - dl) * scale;
            const float d1 = p1 * (__uint_as_float(dp[half * 16 + c + 1]) - dl) * scale;
            pk[half * 8 + (c >> 1)] = pack_bf16x2(d0, d1);
          }
        }
      };
      if (need_mask) ds_phase(std::true_type{}); else ds_phase(std::false_type{});
      tmem_st_32x32b_x16(tS, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_ready[it & 1]);
    }
    if (n > 0) {
      mbar_wait(dq_done, 0);
      tc_fence_after();
    }
    uint32_t ov[32];
    if (n > 0) {
      tmem_ld_32x32b_x32(t_lane + T_DQ + part * 32, ov);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ov[j] = 0u;
    }
    if (valid) {
      bf16* op = dQ + (long long)my_row * lddq + (long long)h * HD + part * 32;
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 0]), __uint_as_float(ov[j8 * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 2]), __uint_as_float(ov[j8 * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 4]), __uint_as_float(ov[j8 * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(ov[j8 * 8 + 6]), __uint_as_float(ov[j8 * 8 + 7]));
        *reinterpret_cast<uint4*>(op + j8 * 8) = o;
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

}  // namespace

// Workspace (fp32 words): delta [Hq*T] | tile summaries [4*ceil(T/128)] (int32) | dK partials [Hq*Tk*128] | dV partials.
extern "C" long long t1_attn_bwd_tc_workspace_bytes(int T, int Tk, int Hq) {
  const long long nqt = (T + BT - 1) / BT;
  const long long words = ((long long)Hq * T + 3) / 4 * 4 + 4 * nqt + 2LL * Hq * Tk * HD;
  return words * 4;
}

// Same contract as t1_attn_bwd for head_dim 128 (Q/K/V/O/dO/dQ/dK/dV token strides multiples of 8 elements, head
// column offsets multiples of 64, 16-byte aligned bases); `ws` from t1_attn_bwd_tc_workspace_bytes.
extern "C" int t1_attn_bwd_tc(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                              const void* O, long long ldo, const void* dO, long long lddo, const float* LSE2, void* ws,
                              long long ws_bytes, void* dQ, long long lddq, void* dK, long long lddk, void* dV,
                              long long lddv, const int* row_lo, const int* row_hi, int shared_len, int T, int Tk, int Hq,
                              int Hkv, int head_dim, float scale, cudaStream_t stream) {
  if (T <= 0) return T1_OK;
  T1_CHECK_ARG(head_dim == HD, "t1_attn_bwd_tc: head_dim must be 128 (got %d)", head_dim);
  T1_CHECK_ARG(Q && K && V && O && dO && LSE2 && ws && dQ && dK && dV && row_lo && row_hi, "t1_attn_bwd_tc: null pointer");
  T1_CHECK_ARG(Hkv > 0 && Hq % Hkv == 0, "t1_attn_bwd_tc: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  T1_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 &&
                   lddk % 4 == 0 && lddv % 4 == 0,
               "t1_attn_bwd_tc: strides must be multiples of 8");
  T1_CHECK_ARG((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)V | (uintptr_t)O | (uintptr_t)dO | (uintptr_t)dQ |
                 (uintptr_t)dK | (uintptr_t)dV | (uintptr_t)ws) & 15) == 0,
               "t1_attn_bwd_tc: 16-byte alignment required");
  T1_CHECK_ARG(ws_bytes >= t1_attn_bwd_tc_workspace_bytes(T, Tk, Hq), "t1_attn_bwd_tc: workspace too small");
  const int nqt = (T + BT - 1) / BT, nkt = (Tk + BT - 1) / BT;
  float* delta = reinterpret_cast<float*>(ws);
  const long long delta_words = ((long long)Hq * T + 3) / 4 * 4;
  int* summ = reinterpret_cast<int*>(delta + delta_words);
  float* part_dk = delta + delta_words + 4LL * nqt;
  float* part_dv = part_dk + (long long)Hq * Tk * HD;

  CUtensorMap tq, tk, tv, tdo;
  int rc = t1_make_tmap_bf16(&tq, Q, T, (long long)Hq * HD, ldq, BT, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tk, K, Tk, (long long)Hkv * HD, ldk, BT, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tv, V, Tk, (long long)Hkv * HD, ldv, BT, 64);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tdo, dO, T, (long long)Hq * HD, lddo, BT, 64);
  if (rc) return rc;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fa_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fa_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM);
    if (e != cudaSuccess) {
      t1_set_error("t1_attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    set = true;
  }
  const float sl2 = scale * 1.4426950408889634f;
  {
    const long long warps = (long long)T * Hq;
    bwd_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>((const bf16*)O, ldo, (const bf16*)dO, lddo,
                                                                               delta, T, Hq);
    T1_CUDA_LAUNCH_CHECK("t1_attn_bwd_tc.delta");
    bwd_tile_summary_kernel<<<nqt, BT, 0, stream>>>(row_lo, row_hi, T, summ);
    T1_CUDA_LAUNCH_CHECK("t1_attn_bwd_tc.summary");
  }
  fa_bwd_dkv_tc_kernel<<<dim3(nkt, Hq), BWD_THREADS, DKV_SMEM, stream>>>(tq, tk, tv, tdo, LSE2, delta, part_dk, part_dv,
                                                                         row_lo, row_hi, summ, shared_len, T, Tk,
                                                                         Hq / Hkv, scale, sl2);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd_tc.dkv");
  fa_bwd_dq_tc_kernel<<<dim3(nqt, Hq), BWD_THREADS, DQ_SMEM, stream>>>(tq, tk, tv, tdo, LSE2, delta, (bf16*)dQ, lddq, row_lo,
                                                                       row_hi, shared_len, T, Hq / Hkv, scale, sl2);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd_tc.dq");
  {
    const long long total = (long long)Tk * Hkv * (HD / 4);
    bwd_reduce_heads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(part_dk, part_dv, (bf16*)dK, lddk, (bf16*)dV,
                                                                                 lddv, Tk, Hkv, Hq / Hkv);
    T1_CUDA_LAUNCH_CHECK("t1_attn_bwd_tc.reduce");
  }
  return T1_OK;
}
