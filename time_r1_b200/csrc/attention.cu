// Flash attention (forward + backward) over a packed token axis with per-row key ranges.
//
// Replaces the flash-attn-2 varlen calls the reference reaches through HF (ViT windowed / full attention:
// modeling_qwen2_5_vl.py:244-261,498-509; LLM causal attention :740-755; invoked from
// src/time_r1/rl/timer1_trainer.py:452-457 and :569-573).
//
// Masking model (covers every variant on the path with one kernel):
//   key j is visible to query row i  <=>  j < row_hi[i]  &&  (j >= row_lo[i]  ||  j < shared_len)
//     * causal prefill           : lo = 0,            hi = i + 1,      shared = 0
//     * packed GRPO scoring      : prefix rows as above; completion rows lo = start of own completion,
//                                  hi = i + 1, shared = P  (all G completions share ONE copy of the prompt)
//     * ViT window / full blocks : lo = window start, hi = window end, shared = 0 (non-causal)
// Layout: Q [T, Hq, HD] / K,V [Tk, Hkv, HD] with arbitrary token strides (so the fused QKV GEMM output is consumed
// in place), O [T, Hq*HD] contiguous heads, LSE2 [Hq, T] fp32 in log2 domain (m2 + log2(l)).
// Tensor path: mma.sync m16n8k16 bf16 (legacy HMMA path; the tcgen05 port of this kernel is the next step),
// cp.async double-buffered K/V tiles, ldmatrix fragments, padded smem rows (HD + 8) to avoid bank conflicts.
#include "t1_common.cuh"
#include <limits.h>

namespace {

constexpr int FA_BM = 64;   // query rows per CTA (4 warps x 16)
constexpr int FA_BN = 64;   // keys per tile
constexpr int FA_THREADS = 128;

template <int HD>
__device__ __forceinline__ void load_tile_async(bf16* smem, const bf16* g, long long ld, int row0, int nrows_valid,
                                                int tid) {
  // rows x HD bf16 -> smem rows of (HD + 8); 16-byte chunks; rows >= nrows_valid are zero-filled
  constexpr int CH = HD / 8;
  constexpr int LDS = HD + 8;
  for (int e = tid; e < 64 * CH; e += FA_THREADS) {
    const int r = e / CH, c = e - r * CH;
    const bool ok = (row0 + r) < nrows_valid;
    const bf16* src = g + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async_16(smem_u32(smem + r * LDS + c * 8), src, ok);
  }
}

struct TileRange {
  int tA_end, tB_start, tB_end;
  __device__ __forceinline__ int count() const { return tA_end + max(0, tB_end - tB_start); }
  __device__ __forceinline__ int tile(int idx) const { return idx < tA_end ? idx : tB_start + (idx - tA_end); }
};

// =============================================================================================
// forward
// =============================================================================================
template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
fa_fwd_kernel(const bf16* __restrict__ Q, long long ldq, const bf16* __restrict__ K, long long ldk,
              const bf16* __restrict__ V, long long ldv, bf16* __restrict__ O, long long ldo,
              float* __restrict__ LSE2, const int* __restrict__ row_lo, const int* __restrict__ row_hi,
              int shared_len, int T, int Tk, int group, float scale_log2) {
  constexpr int LDS = HD + 8;
  constexpr int KS = HD / 16;       // k-steps for Q K^T
  constexpr int DT = HD / 8;        // n-tiles of the output
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(fa_smem);
  bf16* sK = sQ + 64 * LDS;          // 2 stages
  bf16* sV = sK + 2 * 64 * LDS;      // 2 stages
  __shared__ int sred[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * FA_BM;
  const int h = blockIdx.y;
  const int hk = h / group;
  const bf16* Qh = Q + (long long)h * HD;
  const bf16* Kh = K + (long long)hk * HD;
  const bf16* Vh = V + (long long)hk * HD;

  // the two rows this thread owns
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const bool v0 = r0 < T, v1 = r1 < T;
  const int lo0 = v0 ? row_lo[r0] : 0, hi0 = v0 ? row_hi[r0] : 0;
  const int lo1 = v1 ? row_lo[r1] : 0, hi1 = v1 ? row_hi[r1] : 0;
  int min_lo, max_lo, min_hi, max_hi;
  // reduce over both rows: fold row1 into the per-thread values first
  int mlo = INT_MAX, xlo = INT_MIN, mhi = INT_MAX, xhi = INT_MIN;
  if (v0) { mlo = lo0; xlo = lo0; mhi = hi0; xhi = hi0; }
  if (v1) { mlo = min(mlo, lo1); xlo = max(xlo, lo1); mhi = min(mhi, hi1); xhi = max(xhi, hi1); }
  TileRange tr;
  {
    // make_range reduces a single (lo,hi) per thread; feed it the folded min via two calls' worth of data
    int a = mlo, b = xlo, c = mhi, d = xhi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
      c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
      d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    }
    if (lane == 0) { sred[warp * 4] = a; sred[warp * 4 + 1] = b; sred[warp * 4 + 2] = c; sred[warp * 4 + 3] = d; }
    __syncthreads();
    min_lo = INT_MAX; max_lo = INT_MIN; min_hi = INT_MAX; max_hi = INT_MIN;
    for (int i = 0; i < 4; ++i) {
      min_lo = min(min_lo, sred[i * 4]); max_lo = max(max_lo, sred[i * 4 + 1]);
      min_hi = min(min_hi, sred[i * 4 + 2]); max_hi = max(max_hi, sred[i * 4 + 3]);
    }
    if (max_hi <= 0 || min_lo == INT_MAX) {
      tr.tA_end = tr.tB_start = tr.tB_end = 0;
    } else {
      const int a_end = min(min(shared_len, min_lo), max_hi);
      tr.tA_end = (a_end + FA_BN - 1) / FA_BN;
      tr.tB_start = max(min_lo / FA_BN, tr.tA_end);
      tr.tB_end = (max_hi + FA_BN - 1) / FA_BN;
    }
  }
  const int ntiles = tr.count();

  // ---- prologue: Q tile + first K/V tile ----
  load_tile_async<HD>(sQ, Qh, ldq, q0, T, tid);
  if (ntiles > 0) {
    const int j0 = tr.tile(0) * FA_BN;
    load_tile_async<HD>(sK, Kh, ldk, j0, Tk, tid);
    load_tile_async<HD>(sV, Vh, ldv, j0, Tk, tid);
  }
  cp_async_commit();

  float o_acc[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  uint32_t qf[KS][4];
  bool q_loaded = false;

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) {
      const int jn = tr.tile(it + 1) * FA_BN;
      load_tile_async<HD>(sK + (buf ^ 1) * 64 * LDS, Kh, ldk, jn, Tk, tid);
      load_tile_async<HD>(sV + (buf ^ 1) * 64 * LDS, Vh, ldv, jn, Tk, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (!q_loaded) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const bf16* p = sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qf[ks], smem_u32(p));
      }
      q_loaded = true;
    }
    const bf16* kb = sK + buf * 64 * LDS;
    const bf16* vb = sV + buf * 64 * LDS;
    const int j0 = tr.tile(it) * FA_BN;

    // ---- S = Q K^T ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4];
        const bf16* p = kb + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(kf, smem_u32(p));
        mma_bf16_16816(s[np * 2], qf[ks], kf);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], kf + 2);
      }
    }
    // ---- scale + mask ----
    const int j1 = j0 + FA_BN;
    const bool need_mask = !(j1 <= min_hi && (j0 >= max_lo || j1 <= shared_len));
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float val = s[nt][e] * scale_log2;
        if (need_mask) {
          const int j = j0 + nt * 8 + t4 * 2 + (e & 1);
          const bool vis = (e < 2) ? (j < hi0 && (j >= lo0 || j < shared_len)) : (j < hi1 && (j >= lo1 || j < shared_len));
          if (!vis) val = -INFINITY;
        }
        s[nt][e] = val;
      }
    }
    // ---- online softmax ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float mu0 = mn0 == -INFINITY ? 0.f : mn0, mu1 = mn1 == -INFINITY ? 0.f : mn1;
    const float al0 = exp2f(m0 - mu0), al1 = exp2f(m1 - mu1);  // m == -inf -> 0
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mu0), p1 = exp2f(s[nt][1] - mu0);
      const float p2 = exp2f(s[nt][2] - mu1), p3 = exp2f(s[nt][3] - mu1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      // accumulator (16x8) pair -> A fragment (16x16): tiles 2k and 2k+1 form k-step k
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      o_acc[i][0] *= al0;
      o_acc[i][1] *= al0;
      o_acc[i][2] *= al1;
      o_acc[i][3] *= al1;
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        uint32_t vf[4];
        const bf16* p = vb + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vf, smem_u32(p));
        mma_bf16_16816(o_acc[dp * 2], pf[kk], vf);
        mma_bf16_16816(o_acc[dp * 2 + 1], pf[kk], vf + 2);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- epilogue: normalise, stage through sQ, coalesced 16-byte stores ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g) * LDS + i * 8 + t4 * 2) = pack_bf16x2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g + 8) * LDS + i * 8 + t4 * 2) = pack_bf16x2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
  }
  if (LSE2 && t4 == 0) {
    if (v0) LSE2[(long long)h * T + r0] = l0 > 0.f ? m0 + log2f(l0) : -INFINITY;
    if (v1) LSE2[(long long)h * T + r1] = l1 > 0.f ? m1 + log2f(l1) : -INFINITY;
  }
  __syncthreads();
  constexpr int CH = HD / 8;
  for (int e = tid; e < 64 * CH; e += FA_THREADS) {
    const int r = e / CH, c = e - r * CH;
    if (q0 + r < T)
      *reinterpret_cast<uint4*>(O + (long long)(q0 + r) * ldo + (long long)h * HD + c * 8) =
          *reinterpret_cast<const uint4*>(sQ + r * LDS + c * 8);
  }
}

// =============================================================================================
// backward
//   delta[h, i] = sum_d dO[i,h,d] * O[i,h,d]
//   dKV kernel : one CTA per (key tile, kv head); loops over the q heads of the group and the query tiles that
//                can see the key tile;  S^T = K Q^T, P^T, dP^T = V dO^T, dS^T;  dV += P^T dO, dK += dS^T Q
//   dQ  kernel : one CTA per (query tile, q head); loops over visible key tiles;  dQ += dS K
// =============================================================================================
template <int HD>
__global__ void fa_delta_kernel(const bf16* __restrict__ O, long long ldo, const bf16* __restrict__ dO, long long lddo,
                                float* __restrict__ delta, int T, int Hq) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * Hq) return;
  const int i = warp / Hq, h = warp - i * Hq;
  const bf16* o = O + (long long)i * ldo + h * HD;
  const bf16* d = dO + (long long)i * lddo + h * HD;
  float acc = 0.f;
  for (int c = lane * 2; c < HD; c += 64) {
    const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(o + c));
    const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d + c));
    acc += a.x * b.x + a.y * b.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[(long long)h * T + i] = acc;
}

// query-tile ranges that can see a given key tile: rows with hi > j0 and (lo < j1 or shared covers): conservative
// host-free approach — iterate all query tiles whose [min_lo, max_hi) / shared range intersects, decided per tile
// from per-tile summaries computed by fa_tile_summary_kernel.
__global__ void fa_tile_summary_kernel(const int* __restrict__ row_lo, const int* __restrict__ row_hi, int T,
                                       int* __restrict__ summ /*[ntiles,4] min_lo,max_lo,min_hi,max_hi*/) {
  const int tile = blockIdx.x;
  const int lane = threadIdx.x;  // 64 threads
  const int r = tile * FA_BM + lane;
  int a = INT_MAX, b = INT_MIN, c = INT_MAX, d = INT_MIN;
  if (r < T) { a = b = row_lo[r]; c = d = row_hi[r]; }
  __shared__ int s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
    d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
  }
  if ((lane & 31) == 0) { s[(lane >> 5) * 4] = a; s[(lane >> 5) * 4 + 1] = b; s[(lane >> 5) * 4 + 2] = c; s[(lane >> 5) * 4 + 3] = d; }
  __syncthreads();
  if (lane == 0) {
    summ[tile * 4 + 0] = min(s[0], s[4]);
    summ[tile * 4 + 1] = max(s[1], s[5]);
    summ[tile * 4 + 2] = min(s[2], s[6]);
    summ[tile * 4 + 3] = max(s[3], s[7]);
  }
}

template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
fa_bwd_dkv_kernel(const bf16* __restrict__ Q, long long ldq, const bf16* __restrict__ K, long long ldk,
                  const bf16* __restrict__ V, long long ldv, const bf16* __restrict__ dO, long long lddo,
                  const float* __restrict__ LSE2, const float* __restrict__ delta, bf16* __restrict__ dK,
                  long long lddk, bf16* __restrict__ dV, long long lddv, const int* __restrict__ row_lo,
                  const int* __restrict__ row_hi, const int* __restrict__ summ, int shared_len, int T, int Tk,
                  int group, float scale, float scale_log2) {
  constexpr int LDS = HD + 8;
  constexpr int KS = HD / 16;
  constexpr int DT = HD / 8;
  constexpr int QB = 32;  // queries per inner step
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16* sK = reinterpret_cast<bf16*>(fa_smem);   // 64 x LDS
  bf16* sV = sK + 64 * LDS;                      // 64 x LDS
  bf16* sQ = sV + 64 * LDS;                      // 2 stages x 32 x LDS
  bf16* sdO = sQ + 2 * QB * LDS;                 // 2 stages x 32 x LDS
  __shared__ float sLse[2][QB], sDelta[2][QB];
  __shared__ int sLo[2][QB], sHi[2][QB];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int j0 = blockIdx.x * FA_BN, j1 = j0 + FA_BN;
  const int hk = blockIdx.y;
  const int nq_tiles64 = (T + FA_BM - 1) / FA_BM;

  // K, V tiles (whole CTA lifetime)
  load_tile_async<HD>(sK, K + (long long)hk * HD, ldk, j0, Tk, tid);
  load_tile_async<HD>(sV, V + (long long)hk * HD, ldv, j0, Tk, tid);
  cp_async_commit();

  float dk_acc[DT][4], dv_acc[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
    dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
  }
  // keys owned by this thread in the S^T accumulator: rows (warp*16 + g) and (+8)
  const int key0 = j0 + warp * 16 + g, key1 = key0 + 8;

  auto issue_q = [&](int stage, int qh, int qrow0) {
    constexpr int CH = HD / 8;
    for (int e = tid; e < QB * CH; e += FA_THREADS) {
      const int r = e / CH, c = e - r * CH;
      const bool ok = (qrow0 + r) < T;
      const long long row = ok ? qrow0 + r : 0;
      cp_async_16(smem_u32(sQ + (stage * QB + r) * LDS + c * 8), Q + row * ldq + (long long)qh * HD + c * 8, ok);
      cp_async_16(smem_u32(sdO + (stage * QB + r) * LDS + c * 8), dO + row * lddo + (long long)qh * HD + c * 8, ok);
    }
    if (tid < QB) {
      const int r = qrow0 + tid;
      const bool ok = r < T;
      sLse[stage][tid] = ok ? LSE2[(long long)qh * T + r] : INFINITY;
      sDelta[stage][tid] = ok ? delta[(long long)qh * T + r] : 0.f;
      sLo[stage][tid] = ok ? row_lo[r] : 0;
      sHi[stage][tid] = ok ? row_hi[r] : 0;
    }
  };

  // enumerate work items: (q head in group) x (32-row query blocks whose 64-row tile can see this key tile)
  // a 64-row tile qt can see key tile [j0,j1) iff max_hi > j0 and (min_lo < j1 or shared_len > j0)
  auto tile_visible = [&](int qt) {
    const int mlo = summ[qt * 4 + 0], xhi = summ[qt * 4 + 3];
    return xhi > j0 && (mlo < j1 || shared_len > j0);
  };
  // iterate linear index over (qh_local, qt, half) lazily
  int cur_h = 0, cur_qt = 0, cur_half = 0;
  auto advance = [&]() -> bool {  // moves (cur_h,cur_qt,cur_half) to the next visible item; false when exhausted
    while (cur_h < group) {
      while (cur_qt < nq_tiles64) {
        if (tile_visible(cur_qt)) {
          if (cur_half < 2) return true;
        }
        cur_qt++;
        cur_half = 0;
      }
      cur_h++;
      cur_qt = 0;
      cur_half = 0;
    }
    return false;
  };
  bool have = advance();
  int stage = 0;
  int item_h = 0, item_row0 = 0;
  if (have) {
    item_h = hk * group + cur_h;
    item_row0 = cur_qt * FA_BM + cur_half * QB;
    issue_q(0, item_h, item_row0);
    cur_half++;
  }
  cp_async_commit();

  while (have) {
    // prefetch next item
    const bool have_next = advance();
    int nh = 0, nrow0 = 0;
    if (have_next) {
      nh = hk * group + cur_h;
      nrow0 = cur_qt * FA_BM + cur_half * QB;
      issue_q(stage ^ 1, nh, nrow0);
      cur_half++;
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    const bf16* q_s = sQ + stage * QB * LDS;
    const bf16* do_s = sdO + stage * QB * LDS;
    // ---- S^T = K Q^T  (16 keys x 32 queries per warp) and dP^T = V dO^T ----
    float st[4][4], dpt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t kf[4], vf[4];
      const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, acol = ks * 16 + (lane >> 4) * 8;
      ldmatrix_x4(kf, smem_u32(sK + arow * LDS + acol));
      ldmatrix_x4(vf, smem_u32(sV + arow * LDS + acol));
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t qf[4], df[4];
        const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(qf, smem_u32(q_s + brow * LDS + bcol));
        ldmatrix_x4(df, smem_u32(do_s + brow * LDS + bcol));
        mma_bf16_16816(st[np * 2], kf, qf);
        mma_bf16_16816(st[np * 2 + 1], kf, qf + 2);
        mma_bf16_16816(dpt[np * 2], vf, df);
        mma_bf16_16816(dpt[np * 2 + 1], vf, df + 2);
      }
    }
    // ---- P^T, dS^T (rows = keys key0/key1; cols = queries nt*8 + t4*2 + {0,1}) ----
    uint32_t pf[2][4], dsf[2][4];  // A fragments: 2 k-steps of 16 queries
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qi = nt * 8 + t4 * 2 + (e & 1);
        const int key = (e < 2) ? key0 : key1;
        const int lo = sLo[stage][qi], hi = sHi[stage][qi];
        const bool vis = key < hi && (key >= lo || key < shared_len);
        const float pv = vis ? exp2f(st[nt][e] * scale_log2 - sLse[stage][qi]) : 0.f;
        p[e] = pv;
        ds[e] = pv * (dpt[nt][e] - sDelta[stage][qi]);
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p[0], p[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
      dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // ---- dV += P^T dO ; dK += dS^T Q   (contraction over the 32 queries) ----
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        uint32_t bf_do[4], bf_q[4];
        const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, bcol = dp * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(bf_do, smem_u32(do_s + brow * LDS + bcol));
        ldmatrix_x4_trans(bf_q, smem_u32(q_s + brow * LDS + bcol));
        mma_bf16_16816(dv_acc[dp * 2], pf[kk], bf_do);
        mma_bf16_16816(dv_acc[dp * 2 + 1], pf[kk], bf_do + 2);
        mma_bf16_16816(dk_acc[dp * 2], dsf[kk], bf_q);
        mma_bf16_16816(dk_acc[dp * 2 + 1], dsf[kk], bf_q + 2);
      }
    }
    __syncthreads();
    have = have_next;
    stage ^= 1;
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- write dK (x scale) and dV: stage through sK / sV for 16-byte stores ----
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + g) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dk_acc[i][0] * scale, dk_acc[i][1] * scale);
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + g + 8) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dk_acc[i][2] * scale, dk_acc[i][3] * scale);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + g) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dv_acc[i][0], dv_acc[i][1]);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + g + 8) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dv_acc[i][2], dv_acc[i][3]);
  }
  __syncthreads();
  constexpr int CH = HD / 8;
  for (int e = tid; e < 64 * CH; e += FA_THREADS) {
    const int r = e / CH, c = e - r * CH;
    if (j0 + r < Tk) {
      *reinterpret_cast<uint4*>(dK + (long long)(j0 + r) * lddk + (long long)hk * HD + c * 8) = *reinterpret_cast<const uint4*>(sK + r * LDS + c * 8);
      *reinterpret_cast<uint4*>(dV + (long long)(j0 + r) * lddv + (long long)hk * HD + c * 8) = *reinterpret_cast<const uint4*>(sV + r * LDS + c * 8);
    }
  }
  (void)key1;
}

template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
fa_bwd_dq_kernel(const bf16* __restrict__ Q, long long ldq, const bf16* __restrict__ K, long long ldk,
                 const bf16* __restrict__ V, long long ldv, const bf16* __restrict__ dO, long long lddo,
                 const float* __restrict__ LSE2, const float* __restrict__ delta, bf16* __restrict__ dQ,
                 long long lddq, const int* __restrict__ row_lo, const int* __restrict__ row_hi, int shared_len, int T,
                 int Tk, int group, float scale, float scale_log2) {
  constexpr int LDS = HD + 8;
  constexpr int KS = HD / 16;
  constexpr int DT = HD / 8;
  extern __shared__ __align__(16) uint8_t fa_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(fa_smem);
  bf16* sdO = sQ + 64 * LDS;
  bf16* sK = sdO + 64 * LDS;     // 2 stages
  bf16* sV = sK + 2 * 64 * LDS;  // 2 stages
  __shared__ int sred[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * FA_BM;
  const int h = blockIdx.y, hk = h / group;
  const bf16* Kh = K + (long long)hk * HD;
  const bf16* Vh = V + (long long)hk * HD;

  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const bool v0 = r0 < T, v1 = r1 < T;
  const int lo0 = v0 ? row_lo[r0] : 0, hi0 = v0 ? row_hi[r0] : 0;
  const int lo1 = v1 ? row_lo[r1] : 0, hi1 = v1 ? row_hi[r1] : 0;
  const float lse0 = v0 ? LSE2[(long long)h * T + r0] : INFINITY, lse1 = v1 ? LSE2[(long long)h * T + r1] : INFINITY;
  const float dl0 = v0 ? delta[(long long)h * T + r0] : 0.f, dl1 = v1 ? delta[(long long)h * T + r1] : 0.f;
  int a = INT_MAX, b = INT_MIN, c = INT_MAX, d = INT_MIN;
  if (v0) { a = lo0; b = lo0; c = hi0; d = hi0; }
  if (v1) { a = min(a, lo1); b = max(b, lo1); c = min(c, hi1); d = max(d, hi1); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
    d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
  }
  if (lane == 0) { sred[warp * 4] = a; sred[warp * 4 + 1] = b; sred[warp * 4 + 2] = c; sred[warp * 4 + 3] = d; }
  __syncthreads();
  int min_lo = INT_MAX, max_lo = INT_MIN, min_hi = INT_MAX, max_hi = INT_MIN;
  for (int i = 0; i < 4; ++i) {
    min_lo = min(min_lo, sred[i * 4]); max_lo = max(max_lo, sred[i * 4 + 1]);
    min_hi = min(min_hi, sred[i * 4 + 2]); max_hi = max(max_hi, sred[i * 4 + 3]);
  }
  TileRange tr;
  if (max_hi <= 0 || min_lo == INT_MAX) {
    tr.tA_end = tr.tB_start = tr.tB_end = 0;
  } else {
    const int a_end = min(min(shared_len, min_lo), max_hi);
    tr.tA_end = (a_end + FA_BN - 1) / FA_BN;
    tr.tB_start = max(min_lo / FA_BN, tr.tA_end);
    tr.tB_end = (max_hi + FA_BN - 1) / FA_BN;
  }
  const int ntiles = tr.count();

  load_tile_async<HD>(sQ, Q + (long long)h * HD, ldq, q0, T, tid);
  load_tile_async<HD>(sdO, dO + (long long)h * HD, lddo, q0, T, tid);
  if (ntiles > 0) {
    const int j0 = tr.tile(0) * FA_BN;
    load_tile_async<HD>(sK, Kh, ldk, j0, Tk, tid);
    load_tile_async<HD>(sV, Vh, ldv, j0, Tk, tid);
  }
  cp_async_commit();

  float dq_acc[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) {
      const int jn = tr.tile(it + 1) * FA_BN;
      load_tile_async<HD>(sK + (buf ^ 1) * 64 * LDS, Kh, ldk, jn, Tk, tid);
      load_tile_async<HD>(sV + (buf ^ 1) * 64 * LDS, Vh, ldv, jn, Tk, tid);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const bf16* kb = sK + buf * 64 * LDS;
    const bf16* vb = sV + buf * 64 * LDS;
    const int j0 = tr.tile(it) * FA_BN;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qf[4], df[4];
      const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, acol = ks * 16 + (lane >> 4) * 8;
      ldmatrix_x4(qf, smem_u32(sQ + arow * LDS + acol));
      ldmatrix_x4(df, smem_u32(sdO + arow * LDS + acol));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4], vf[4];
        const int brow = np * 16 + (lane & 7) + (lane >> 4) * 8, bcol = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(kf, smem_u32(kb + brow * LDS + bcol));
        ldmatrix_x4(vf, smem_u32(vb + brow * LDS + bcol));
        mma_bf16_16816(s[np * 2], qf, kf);
        mma_bf16_16816(s[np * 2 + 1], qf, kf + 2);
        mma_bf16_16816(dp[np * 2], df, vf);
        mma_bf16_16816(dp[np * 2 + 1], df, vf + 2);
      }
    }
    uint32_t dsf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + nt * 8 + t4 * 2 + (e & 1);
        const bool vis = (e < 2) ? (j < hi0 && (j >= lo0 || j < shared_len)) : (j < hi1 && (j >= lo1 || j < shared_len));
        const float p = vis ? exp2f(s[nt][e] * scale_log2 - ((e < 2) ? lse0 : lse1)) : 0.f;
        ds[e] = p * (dp[nt][e] - ((e < 2) ? dl0 : dl1));
      }
      dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
    // dQ += dS K   (B[k=key][n=d] = K[key][d] -> transposed ldmatrix)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dpi = 0; dpi < DT / 2; ++dpi) {
        uint32_t kf[4];
        const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, bcol = dpi * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(kf, smem_u32(kb + brow * LDS + bcol));
        mma_bf16_16816(dq_acc[dpi * 2], dsf[kk], kf);
        mma_bf16_16816(dq_acc[dpi * 2 + 1], dsf[kk], kf + 2);
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dq_acc[i][0] * scale, dq_acc[i][1] * scale);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + g + 8) * LDS + i * 8 + t4 * 2) = pack_bf16x2(dq_acc[i][2] * scale, dq_acc[i][3] * scale);
  }
  __syncthreads();
  constexpr int CH = HD / 8;
  for (int e = tid; e < 64 * CH; e += FA_THREADS) {
    const int r = e / CH, cc = e - r * CH;
    if (q0 + r < T)
      *reinterpret_cast<uint4*>(dQ + (long long)(q0 + r) * lddq + (long long)h * HD + cc * 8) =
          *reinterpret_cast<const uint4*>(sQ + r * LDS + cc * 8);
  }
  (void)max_lo; (void)min_hi;
}

template <int HD>
int launch_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
               long long ldo, float* LSE2, const int* row_lo, const int* row_hi, int shared_len, int T, int Tk, int Hq,
               int Hkv, float scale, cudaStream_t stream) {
  constexpr int SMEM = (64 + 4 * 64) * (HD + 8) * 2;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(fa_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  dim3 grid((T + FA_BM - 1) / FA_BM, Hq);
  fa_fwd_kernel<HD><<<grid, FA_THREADS, SMEM, stream>>>((const bf16*)Q, ldq, (const bf16*)K, ldk, (const bf16*)V, ldv,
                                                        (bf16*)O, ldo, LSE2, row_lo, row_hi, shared_len, T, Tk,
                                                        Hq / Hkv, scale * 1.4426950408889634f);
  T1_CUDA_LAUNCH_CHECK("t1_attn_fwd");
  return T1_OK;
}

template <int HD>
int launch_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, const void* O,
               long long ldo, const void* dO, long long lddo, const float* LSE2, float* delta, int* summ, void* dQ,
               long long lddq, void* dK, long long lddk, void* dV, long long lddv, const int* row_lo,
               const int* row_hi, int shared_len, int T, int Tk, int Hq, int Hkv, float scale, cudaStream_t stream) {
  const float sl2 = scale * 1.4426950408889634f;
  fa_delta_kernel<HD><<<(T * Hq * 32 + 255) / 256, 256, 0, stream>>>((const bf16*)O, ldo, (const bf16*)dO, lddo, delta, T, Hq);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd.delta");
  const int nqt = (T + FA_BM - 1) / FA_BM;
  fa_tile_summary_kernel<<<nqt, 64, 0, stream>>>(row_lo, row_hi, T, summ);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd.summary");
  constexpr int SMEM_KV = (2 * 64 + 4 * 32) * (HD + 8) * 2;
  constexpr int SMEM_Q = (2 * 64 + 4 * 64) * (HD + 8) * 2;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(fa_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV);
    cudaFuncSetAttribute(fa_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_Q);
    set = true;
  }
  dim3 gkv((Tk + FA_BN - 1) / FA_BN, Hkv);
  fa_bwd_dkv_kernel<HD><<<gkv, FA_THREADS, SMEM_KV, stream>>>(
      (const bf16*)Q, ldq, (const bf16*)K, ldk, (const bf16*)V, ldv, (const bf16*)dO, lddo, LSE2, delta, (bf16*)dK,
      lddk, (bf16*)dV, lddv, row_lo, row_hi, summ, shared_len, T, Tk, Hq / Hkv, scale, sl2);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd.dkv");
  dim3 gq(nqt, Hq);
  fa_bwd_dq_kernel<HD><<<gq, FA_THREADS, SMEM_Q, stream>>>((const bf16*)Q, ldq, (const bf16*)K, ldk, (const bf16*)V,
                                                           ldv, (const bf16*)dO, lddo, LSE2, delta, (bf16*)dQ, lddq,
                                                           row_lo, row_hi, shared_len, T, Tk, Hq / Hkv, scale, sl2);
  T1_CUDA_LAUNCH_CHECK("t1_attn_bwd.dq");
  return T1_OK;
}

}  // namespace

extern "C" int t1_attn_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                           void* O, long long ldo, float* LSE2, const int* row_lo, const int* row_hi, int shared_len,
                           int T, int Tk, int Hq, int Hkv, int head_dim, float scale, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(Q && K && V && O && row_lo && row_hi, "t1_attn_fwd: null pointer");
  T1_CHECK_ARG(Hkv > 0 && Hq % Hkv == 0, "t1_attn_fwd: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  T1_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "t1_attn_fwd: strides must be multiples of 8");
  switch (head_dim) {
    case 128: return launch_fwd<128>(Q, ldq, K, ldk, V, ldv, O, ldo, LSE2, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    case 80: return launch_fwd<80>(Q, ldq, K, ldk, V, ldv, O, ldo, LSE2, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    case 64: return launch_fwd<64>(Q, ldq, K, ldk, V, ldv, O, ldo, LSE2, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    default:
      t1_set_error("t1_attn_fwd: unsupported head_dim %d (80, 64 and 128 are built)", head_dim);
      return T1_ERR_UNSUPPORTED;
  }
}

// workspace: delta [Hq*T] floats, summ [ceil(T/64)*4] ints
extern "C" int t1_attn_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                           const void* O, long long ldo, const void* dO, long long lddo, const float* LSE2,
                           float* delta_ws, int* summ_ws, void* dQ, long long lddq, void* dK, long long lddk, void* dV,
                           long long lddv, const int* row_lo, const int* row_hi, int shared_len, int T, int Tk, int Hq,
                           int Hkv, int head_dim, float scale, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(Q && K && V && O && dO && LSE2 && delta_ws && summ_ws && dQ && dK && dV, "t1_attn_bwd: null pointer");
  T1_CHECK_ARG(Hkv > 0 && Hq % Hkv == 0, "t1_attn_bwd: Hq=%d must be a multiple of Hkv=%d", Hq, Hkv);
  switch (head_dim) {
    case 128: return launch_bwd<128>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, LSE2, delta_ws, summ_ws, dQ, lddq, dK, lddk, dV, lddv, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    case 80: return launch_bwd<80>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, LSE2, delta_ws, summ_ws, dQ, lddq, dK, lddk, dV, lddv, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    case 64: return launch_bwd<64>(Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, LSE2, delta_ws, summ_ws, dQ, lddq, dK, lddk, dV, lddv, row_lo, row_hi, shared_len, T, Tk, Hq, Hkv, scale, stream);
    default:
      t1_set_error("t1_attn_bwd: unsupported head_dim %d (64, 80 and 128 are built)", head_dim);
      return T1_ERR_UNSUPPORTED;
  }
}
