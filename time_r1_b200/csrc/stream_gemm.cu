// Weight-streaming decode GEMM for sm_100a:  y[M<=64, N] = x[M,K] * W[N,K]^T (+ bias[N])
//
// Replaces the per-token nn.Linear calls of HF GenerationMixin._sample (driven from
// src/time_r1/rl/timer1_trainer.py:568-573).  With M = (prompt-groups x G) <= 64 rows the op is a pure stream of the
// weight matrix: HBM-bound, 2*N*K bytes per launch.  Design:
//   * operands swapped for the tensor core: the 128 weight rows of a slab are the UMMA M dimension, the (padded) batch
//     rows MP in {16, 32, 64} are the UMMA N dimension; accumulator = MP fp32 TMEM columns (double buffered);
//   * stream-K work split: the (slab, 64-wide k-block) units are laid out slab-major and cut into equal contiguous
//     ranges, one per CTA, so every SM streams the same number of bytes whatever N and K are; a slab shared by
//     several CTAs is finished by the last one to arrive (partials in an L2-resident workspace, summed in a fixed
//     order => deterministic);
//   * warp 0 = TMA producer (128x64 W box + MPx64 x box per stage, SWIZZLE_128B, mbarrier ring), warp 1 = single
//     thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> transposed global stores);
//   * programmatic dependent launch: the producer fills the whole ring with WEIGHT tiles before
//     griddepcontrol.wait -- weights do not depend on the previous kernel -- so the HBM stream of kernel k+1 runs
//     under the tail of kernel k and under the latency-bound kernels (norm, rope, attention) in between.
#include "t1_common.cuh"
#include <stdlib.h>

namespace {

constexpr int SG_BM = 128;              // weight rows per slab
constexpr int SG_BK = 64;               // 64 bf16 = one 128-byte swizzle row
constexpr int SG_W_BYTES = SG_BM * SG_BK * 2;
constexpr int SG_MAX_STAGES = 24;
constexpr int SG_HEADER = 2048;         // barriers + tmem pointer + flags (0..511), folded-norm row factors and scratch (512..1535)
constexpr int SG_THREADS = 192;
constexpr int SG_COUNTER_BYTES = 65536; // slab counters at the head of the workspace (N <= 2M rows)

__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct SgParams {
  void* y;
  const bf16* bias;
  float* ws;          // [grid][2][MP][128] fp32 partial tiles
  int* counters;      // [slabs], zero on entry, reset by the finishing CTA
  long long ldy;
  int M, N, K;
  int kblocks;        // ceil(K / 64)
  int total_units;    // slabs * kblocks
  int per_cta;        // units per CTA
  int stages;
  int out_f32;
  int swiglu;         // 1: W = [gate rows | up rows] (N = 2 * half_n); a slab pairs 64 gate rows with their 64 up rows and
  int half_n;         //    the epilogue emits y[M, half_n] = silu(gate) * up (bf16), as t1_swiglu does on the bf16 GEMM output
  // fused residual + RMSNorm tail (t1_stream_gemm_resnorm): the LAST CTA to finish adds y to the residual stream and
  // normalises it, which removes one kernel (and one kernel boundary) per norm from the decode chain
  bf16* tail_x;       // residual stream [M, N], updated in place: x += y   (nullptr: no tail)
  const bf16* tail_w; // norm weight [N]
  bf16* tail_h;       // rmsnorm(x) * w  [M, N]
  float tail_eps;
  int* tail_counter;  // [2]: arrivals of all CTAs, completions of the row CTAs; zero on entry, re-armed by the last row CTA
  // dynamic scheduler (stream_gemm_dyn_kernel): work items = (slab, k-segment), handed out through an atomic counter
  int segs;           // k-segments per slab
  int seg_k;          // k-blocks per segment
  int total_items;    // slabs * segs
  int* sched_counter; // zero on entry, reset by the CTA that performs the last fetch
  unsigned long long* dbg;   // optional [4] globaltimer stamps of CTA 0: start, weights issued, dependency resolved, done
  // RMSNorm folded into the GEMMs on either side of it (no norm kernel, no grid-wide wait):
  //   producer (t1_stream_gemm_fold): a finished slab adds its y columns to the residual stream h, writes
  //   z = bf16(norm_w * h) -- the operand the NEXT GEMM streams against -- and the slab's per-row sum of h^2;
  //   consumer (t1_stream_gemm_scaled): x = z; the per-row factor r[m] = rsqrt(sum_slabs(parts) / K + eps), which is
  //   constant along k, multiplies the fp32 accumulator in the epilogue:  r[m] * sum_k W[n,k] z[m,k] = W . rmsnorm(h).
  bf16* fold_x;         // residual stream [M, N], updated in place (nullptr: plain epilogue)
  const bf16* fold_w;   // norm weight [N]
  bf16* fold_z;         // [M, N]
  float* fold_parts;    // [slabs][MP] per-row partial sums of h^2 (each written by exactly one CTA: deterministic)
  const float* in_parts;  // consumer: [in_nparts][MP] (nullptr: x is already normalised)
  int in_nparts;
  float in_eps;
  // L2 prefetch of the NEXT weight stream (t1_stream_gemm_prefetch_next): while this launch waits for its dependency the
  // idle epilogue warps pull the units the next GEMM's CTAs read right after their own ring into L2, so the HBM works
  // through the latency-bound part of the chain (qkv -> attention -> o) instead of idling
  const bf16* pf_w;       // next W (nullptr: off)
  long long pf_ldw;
  int pf_n;               // rows of next W
  int pf_kblocks, pf_per_cta, pf_total, pf_grid;   // the next launch's unit split
  int pf_skip, pf_units;  // per CTA of the next launch: units [skip, skip + units) of its range
  int pf_swiglu, pf_half_n;
  int pf_mode;            // 1: per-line prefetch.global.L2, 2: cp.async.bulk.prefetch.L2 per row segment
  int early_resid;        // folded norm: fetch the residual column before the accumulator wait (T1_STREAM_EARLY_RESID=0: after it)
  int fast_sync;          // split-slab hand-over: bar.sync + ONE acq_rel atomic instead of 128 x membar.gl + atomic + membar.gl
};
__device__ __forceinline__ int atom_add_acq_rel_gpu(int* p, int v) {
  int old;
  asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The prefetch warp walks the units [skip, skip + units) of every next-launch CTA it is responsible for.
//   mode 1: one prefetch.global.L2 per 128-byte line (lane l takes rows l, l + 32, l + 64, l + 96 of the slab);
//   mode 2: the row segments of consecutive units of one slab are contiguous along k, so lane l issues ONE
//           cp.async.bulk.prefetch.L2 per (row, run of units) -- asynchronous, nothing for the warp to wait on.
__device__ __forceinline__ void sg_prefetch_next(const SgParams& p, int lane) {
  for (int c = blockIdx.x; c < p.pf_grid; c += gridDim.x) {
    const int c0 = c * p.pf_per_cta;
    const int v0 = c0 + p.pf_skip;
    const int v1 = min(min(v0 + p.pf_units, c0 + p.pf_per_cta), p.pf_total);
    int u = v0;
    while (u < v1) {
      const int slab = u / p.pf_kblocks, kb = u - slab * p.pf_kblocks;
      const int run = min(v1 - u, p.pf_kblocks - kb);            // units of this slab in the range
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = lane + 32 * j;
        const int row = p.pf_swiglu ? (rr < 64 ? slab * 64 + rr : p.pf_half_n + slab * 64 + rr - 64) : slab * SG_BM + rr;
        if (row >= p.pf_n) continue;
        const bf16* src = p.pf_w + (long long)row * p.pf_ldw + kb * SG_BK;
        if (p.pf_mode == 2) {
          const unsigned bytes = (unsigned)run * SG_BK * 2;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        } else {
          for (int i = 0; i < run; ++i) prefetch_l2(src + i * SG_BK);
        }
      }
      u += run;
    }
  }
}
#ifdef T1_SG_DBG
__device__ unsigned long long g_cta_t[3][512];     // per-CTA start / first-data / end stamps of the last launch
#endif
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Epilogue of the units [u0, u1) of one GEMM (warps 2..5 = TMEM lane quadrants 2,3,0,1): accumulator out of TMEM, the
// stream-K fix-up of split slabs, then the plain / SwiGLU / folded-norm store.  it0 = accumulator-buffer parity counter on
// entry (the paired kernel runs two GEMMs back to back through the same TMEM double buffer); returns the counter on exit.
template <int MP>
__device__ __forceinline__ int sg_epilogue(const SgParams& p, uint8_t* smem, int warp_idx, int lane, uint32_t tmem_base,
                                           uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar, int* s_flag, float* xs, int u0,
                                           int u1, int it0) {
  const int et = (warp_idx - 2) * 32 + lane;   // 0..127
  const int KB = p.kblocks;
  const int q = warp_idx & 3;
  const int r = q * 32 + lane;                 // weight row inside the slab == TMEM lane
  float* r_s = reinterpret_cast<float*>(smem + 512);   // [MP] row factors of a folded RMSNorm (header bytes 512..767)
  if (p.in_parts) {
    // all 128 threads fetch (one L2 round trip for <= 4 independent loads each), then MP threads finish in a fixed order
    float* r_tmp = reinterpret_cast<float*>(smem + 1024);   // [128] (header bytes 1024..1535)
    constexpr int GROUPS = 128 / MP;
    const int m = et % MP, g0 = et / MP;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    const int np = p.in_nparts;
    int i = g0;
    for (; i + 3 * GROUPS < np; i += 4 * GROUPS) {
      const float a = __ldcg(p.in_parts + i * MP + m), b = __ldcg(p.in_parts + (i + GROUPS) * MP + m);
      const float c = __ldcg(p.in_parts + (i + 2 * GROUPS) * MP + m), d = __ldcg(p.in_parts + (i + 3 * GROUPS) * MP + m);
      v0 += a; v1 += b; v2 += c; v3 += d;
    }
    for (; i < np; i += GROUPS) v0 += __ldcg(p.in_parts + i * MP + m);
    r_tmp[et] = (v0 + v1) + (v2 + v3);
    epi_bar_sync();
    if (et < MP) {
      float ssq = 0.f;
#pragma unroll
      for (int g = 0; g < GROUPS; ++g) ssq += r_tmp[g * MP + et];
      r_s[et] = rsqrtf(ssq / (float)p.K + p.in_eps);
    }
    epi_bar_sync();
  }
  int it = it0;
  for (int u = u0; u < u1; ++it) {
    const int slab = u / KB, kb0 = u - slab * KB;
    const int kb1 = min(KB, kb0 + (u1 - u));
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    // folded norm: this thread's residual column (and norm weight) is fetched BEFORE the accumulator wait -- nothing else writes
    // column n during the launch (only its own slab's finisher does, below) -- so its L2 round trip leaves the chain's critical path
    constexpr bool kEarlyResid = MP <= 16;
    bf16 xold[kEarlyResid ? MP : 1];
    float gw_early = 0.f;
    float bias_early = 0.f;          // plain epilogue: this thread's bias element, fetched ahead of the wait for the same reason
    if (p.bias && !p.swiglu && p.early_resid) {
      const int n_e = slab * SG_BM + r;
      if (n_e < p.N) bias_early = __bfloat162float(p.bias[n_e]);
    }
    if (kEarlyResid && p.fold_x && p.early_resid) {
      const int n_e = slab * SG_BM + r;
      gw_early = __bfloat162float(p.fold_w[n_e]);
#pragma unroll
      for (int m = 0; m < (kEarlyResid ? MP : 1); ++m) xold[m] = m < p.M ? __ldcg(p.fold_x + (long long)m * p.N + n_e) : __float2bfloat16(0.f);
    }
    mbar_wait(&tmem_full_bar[acc], acc_phase);
    tc_fence_after();
    uint32_t v[MP];
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * MP;
    if constexpr (MP == 16) {
      tmem_ld_32x32b_x16(taddr, v);
    } else {
#pragma unroll
      for (int c = 0; c < MP / 32; ++c) tmem_ld_32x32b_x32(taddr + c * 32, v + c * 32);
    }
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&tmem_empty_bar[acc]);          // accumulator is in registers: the MMA warp may reuse it
    const int n = slab * SG_BM + r;
    const bool whole = (kb0 == 0 && kb1 == KB);
    bool finish = whole;
    int c_first = 0, c_last = 0;
    if (!whole) {
      float* mine = p.ws + ((size_t)(blockIdx.x * 2 + (it == it0 ? 0 : 1)) * MP) * SG_BM;
#pragma unroll
      for (int m = 0; m < MP; ++m) mine[m * SG_BM + r] = __uint_as_float(v[m]);
      // release: the partial tile is visible device-wide before the arrival is.  fast_sync: the CTA barrier orders the 128
      // threads' stores before thread 0's acq_rel atomic (cumulativity), which is also the acquire for the finisher's reads
      if (!p.fast_sync) __threadfence();
      epi_bar_sync();
      c_first = (slab * KB) / p.per_cta;
      c_last = ((slab + 1) * KB - 1) / p.per_cta;
      if (et == 0) {
        const int old = p.fast_sync ? atom_add_acq_rel_gpu(&p.counters[slab], 1) : atomicAdd(&p.counters[slab], 1);
        const int last = (old == c_last - c_first) ? 1 : 0;
        if (last) p.counters[slab] = 0;         // self-resetting: the next launch finds zeros
        *s_flag = last;
      }
      epi_bar_sync();
      finish = *s_flag != 0;
      epi_bar_sync();                            // s_flag may be rewritten by the next segment
      if (finish) {
        if (!p.fast_sync) __threadfence();
#pragma unroll
        for (int m = 0; m < MP; ++m) v[m] = 0u;
        // the partial tiles of FB contributors are fetched together (one L2 round trip per batch instead of one per
        // contributor: this sum sits on the chain's critical path at every split slab) and added in the fixed order c_first..c_last
        constexpr int FB = MP == 16 ? 7 : (MP == 32 ? 2 : 1);      // down_proj at M <= 16: up to 7 contributors per slab -> one batch
        for (int c = c_first; c <= c_last; c += FB) {
          float t[FB][MP];
#pragma unroll
          for (int j = 0; j < FB; ++j) {
            const int cc = min(c + j, c_last);
            const int slot = (slab == (cc * p.per_cta) / KB) ? 0 : 1;
            const float* part = p.ws + ((size_t)(cc * 2 + slot) * MP) * SG_BM;
#pragma unroll
            for (int m = 0; m < MP; ++m) t[j][m] = __ldcg(part + m * SG_BM + r);
          }
#pragma unroll
          for (int j = 0; j < FB; ++j) {
            if (c + j <= c_last) {
#pragma unroll
              for (int m = 0; m < MP; ++m) v[m] = __float_as_uint(__uint_as_float(v[m]) + t[j][m]);
            }
          }
        }
      }
    }
    if (finish && p.in_parts) {
#pragma unroll
      for (int m = 0; m < MP; ++m) v[m] = __float_as_uint(__uint_as_float(v[m]) * r_s[m]);
    }
    if (p.fold_x) {
      // residual add + norm-weight scaling + per-row sum of squares of this slab's 128 columns (N % 128 == 0)
      if (finish) {
        float gw;
        bf16 xo[MP];
        if (kEarlyResid && p.early_resid) {
          gw = gw_early;
#pragma unroll
          for (int m = 0; m < MP; ++m) xo[m] = xold[kEarlyResid ? m : 0];
        } else {
          gw = __bfloat162float(p.fold_w[n]);
#pragma unroll
          for (int m = 0; m < MP; ++m)           // all residual loads in flight together (one L2 round trip, not MP)
            xo[m] = m < p.M ? p.fold_x[(long long)m * p.N + n] : __float2bfloat16(0.f);
        }
        float sq[MP];
#pragma unroll
        for (int m = 0; m < MP; ++m) {
          sq[m] = 0.f;
          if (m < p.M) {
            const float y = __bfloat162float(__float2bfloat16(__uint_as_float(v[m])));
            const bf16 hn = __float2bfloat16(__bfloat162float(xo[m]) + y);  // hidden = residual + hidden in bf16 (HF Q25:805-827)
            p.fold_x[(long long)m * p.N + n] = hn;
            const float hf = __bfloat162float(hn);
            p.fold_z[(long long)m * p.N + n] = __float2bfloat16(gw * hf);
            sq[m] = hf * hf;
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)   // MP independent butterflies: the shuffles of different rows overlap
#pragma unroll
          for (int m = 0; m < MP; ++m) sq[m] += __shfl_xor_sync(0xffffffffu, sq[m], off);
        if (lane == 0) {
#pragma unroll
          for (int m = 0; m < MP; ++m) xs[(warp_idx - 2) * MP + m] = sq[m];
        }
        epi_bar_sync();
        if (et < MP) p.fold_parts[slab * MP + et] = xs[et] + xs[MP + et] + xs[2 * MP + et] + xs[3 * MP + et];
        epi_bar_sync();                          // xs is rewritten by the next finished slab
      }
    } else if (p.swiglu) {
      // lanes 0..63 hold the gate rows, lanes 64..127 the matching up rows: hand the up values over through shared
      // memory, then y = silu(gate) * up on the bf16-rounded GEMM outputs (bit-identical to gemm -> t1_swiglu)
      if (finish) {
        const int col = slab * 64 + (r & 63);
        if (r >= 64) {
          const float b = (p.bias && col < p.half_n) ? __bfloat162float(p.bias[p.half_n + col]) : 0.f;
#pragma unroll
          for (int m = 0; m < MP; ++m) xs[m * 64 + (r - 64)] = __bfloat162float(__float2bfloat16(__uint_as_float(v[m]) + b));
        }
        epi_bar_sync();
        if (r < 64 && col < p.half_n) {
          const float b = p.bias ? __bfloat162float(p.bias[col]) : 0.f;
          bf16* yp = reinterpret_cast<bf16*>(p.y) + col;
#pragma unroll
          for (int m = 0; m < MP; ++m) {
            if (m < p.M) {
              const float g = __bfloat162float(__float2bfloat16(__uint_as_float(v[m]) + b));
              const float u = xs[m * 64 + r];
              yp[(long long)m * p.ldy] = __float2bfloat16(g * (1.0f / (1.0f + __expf(-g))) * u);
            }
          }
        }
        epi_bar_sync();                          // xs is rewritten by the next finished slab
      }
    } else if (finish && n < p.N) {
      const float b = p.bias ? (p.early_resid ? bias_early : __bfloat162float(p.bias[n])) : 0.f;
      if (p.out_f32) {
        float* yp = reinterpret_cast<float*>(p.y) + n;
#pragma unroll
        for (int m = 0; m < MP; ++m)
          if (m < p.M) yp[(long long)m * p.ldy] = __uint_as_float(v[m]) + b;
      } else {
        bf16* yp = reinterpret_cast<bf16*>(p.y) + n;
#pragma unroll
        for (int m = 0; m < MP; ++m)
          if (m < p.M) yp[(long long)m * p.ldy] = __float2bfloat16(__uint_as_float(v[m]) + b);
      }
    }
    u += kb1 - kb0;
  }
  return it;
}

template <int MP>
__global__ void __launch_bounds__(SG_THREADS + 32, 1)
stream_gemm_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_x, SgParams p) {
  constexpr int X_BYTES = MP * SG_BK * 2;
  constexpr int STAGE_BYTES = SG_W_BYTES + X_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * MP < 32) ? 32 : 2 * MP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + SG_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SG_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr_smem + 1);
  uint8_t* tiles = smem + SG_HEADER;
  float* xs = reinterpret_cast<float*>(tiles + p.stages * STAGE_BYTES);   // swiglu: [MP][64] up values

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int KB = p.kblocks;
  const int u0 = blockIdx.x * p.per_cta;
  const int u1 = min(u0 + p.per_cta, p.total_units);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_x);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();   // the next kernel of the chain may become resident (it blocks in its own griddepcontrol.wait)
#ifdef T1_SG_DBG
  if (threadIdx.x == 0) {
    g_cta_t[0][blockIdx.x] = gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_cta_t[1][blockIdx.x] = smid;
  }
#endif

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && u1 > u0) {
      const bool dbg = p.dbg && blockIdx.x == 0;
      if (dbg) p.dbg[0] = gtimer();
      uint64_t pol_stream, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      const int n = u1 - u0;
      const int pre = min(n, stages);
      // This one thread issues every load of the CTA (one 18 KB unit every ~700 cycles at the HBM rate), so the steady state
      // carries no division: (slab, k-block) advance incrementally.  Measured with the paired kernel (same loop): five run-time
      // divisions per unit made the ISSUE the bottleneck (gate_up 67 us instead of 40).
      const int swg = p.swiglu, half_n = p.half_n;
      const int slab0 = u0 / KB, kb0 = u0 - slab0 * KB;
      int slab = slab0, kb = kb0;
      // weights first: they do not depend on the previous kernel
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], STAGE_BYTES);
        if (swg) {
          tma_load_2d_hint(tiles + i * STAGE_BYTES, &tma_w, &full_bar[i], kb * SG_BK, slab * 64, pol_stream);
          tma_load_2d_hint(tiles + i * STAGE_BYTES + 64 * 128, &tma_w, &full_bar[i], kb * SG_BK, half_n + slab * 64, pol_stream);
        } else {
          tma_load_2d_hint(tiles + i * STAGE_BYTES, &tma_w, &full_bar[i], kb * SG_BK, slab * SG_BM, pol_stream);
        }
        if (++kb == KB) { kb = 0; ++slab; }
      }
      if (dbg) p.dbg[1] = gtimer();
      pdl_wait();
      if (dbg) p.dbg[2] = gtimer();
      for (int i = 0, kx = kb0; i < pre; ++i) {
        tma_load_2d_hint(tiles + i * STAGE_BYTES + SG_W_BYTES, &tma_x, &full_bar[i], kx * SG_BK, 0, pol_keep);
        if (++kx == KB) kx = 0;
      }
      int stage = pre == stages ? 0 : pre;
      uint32_t phase = pre == stages ? 1 : 0;   // parity of the NEXT use of `stage`
      for (int i = pre; i < n; ++i) {
        uint8_t* dst = tiles + stage * STAGE_BYTES;
        uint64_t* bar = &full_bar[stage];
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
        if (swg) {
          tma_load_2d_hint(dst, &tma_w, bar, kb * SG_BK, slab * 64, pol_stream);
          tma_load_2d_hint(dst + 64 * 128, &tma_w, bar, kb * SG_BK, half_n + slab * 64, pol_stream);
        } else {
          tma_load_2d_hint(dst, &tma_w, bar, kb * SG_BK, slab * SG_BM, pol_stream);
        }
        tma_load_2d_hint(dst + SG_W_BYTES, &tma_x, bar, kb * SG_BK, 0, pol_keep);
        if (++kb == KB) { kb = 0; ++slab; }
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SG_BM, MP);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = u0; u < u1; ++it) {
        const int slab = u / KB, kb0 = u - slab * KB;
        const int kb1 = min(KB, kb0 + (u1 - u));
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MP;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES));
          const uint64_t b_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES + SG_W_BYTES));
#pragma unroll
          for (int k = 0; k < SG_BK / 16; ++k)
            umma_bf16_ss(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        u += kb1 - kb0;
      }
    }
  } else if (warp_idx == 6) {
    // ===================== L2 prefetch warp (launched only with a prefetch hint) =====================
    // off the critical path: the producer / MMA / epilogue warps never wait for it, and it leaves without the final barrier
    if (p.pf_w) sg_prefetch_next(p, lane);
    return;
  } else {
    // ===================== epilogue (warps 2..5 = TMEM lane quadrants 2,3,0,1) =====================
    const int et = (warp_idx - 2) * 32 + lane;   // 0..127
    pdl_wait();
    sg_epilogue<MP>(p, smem, warp_idx, lane, tmem_base, tmem_full_bar, tmem_empty_bar, s_flag, xs, u0, u1, 0);
    if (p.tail_x) {
      // ---- residual + RMSNorm: a grid-wide arrive counter, then CTA b (b < M) normalises row b ----
      // Every CTA of the launch is resident at the same time (grid <= SM count, one CTA per SM), so the M row-CTAs may
      // spin on the counter; a single-CTA tail would be latency-bound (16 rows x 3 arrays through one SM).
      __threadfence();                             // this CTA's y rows are visible device-wide
      epi_bar_sync();
      if (et == 0) atomicAdd(p.tail_counter, 1);
      const int row_ctas = min(p.M, (int)gridDim.x);
      if ((int)blockIdx.x < row_ctas) {
        if (et == 0) {
          while (atomicAdd(p.tail_counter, 0) < (int)gridDim.x) {}
          __threadfence();
        }
        epi_bar_sync();
        float* s_red = reinterpret_cast<float*>(s_flag + 1);
        const int nv = p.N >> 3;
        const bf16* yb = reinterpret_cast<const bf16*>(p.y);
        constexpr int TB = 4;                                    // 16-byte chunks per thread kept in registers per batch
        for (int m = blockIdx.x; m < p.M; m += gridDim.x) {
          bf16* xr = p.tail_x + (long long)m * p.N;
          bf16* hr = p.tail_h + (long long)m * p.N;
          float sq = 0.f;
          for (int c0 = et; c0 < nv; c0 += 128 * TB) {
            uint4 uy[TB], ux[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) {
              const int c = c0 + t * 128;
              if (c < nv) {
                uy[t] = __ldcg(reinterpret_cast<const uint4*>(yb + (long long)m * p.ldy + c * 8));
                ux[t] = *reinterpret_cast<const uint4*>(xr + c * 8);
              }
            }
#pragma unroll
            for (int t = 0; t < TB; ++t) {
              const int c = c0 + t * 128;
              if (c < nv) {
                const uint32_t* py = reinterpret_cast<const uint32_t*>(&uy[t]);
                const uint32_t* px = reinterpret_cast<const uint32_t*>(&ux[t]);
                uint32_t o[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float2 a = unpack_bf16x2(px[jj]), b = unpack_bf16x2(py[jj]);
                  o[jj] = pack_bf16x2(a.x + b.x, a.y + b.y);      // hidden = residual + hidden in bf16 (HF Q25:805-827)
                  const float2 r2 = unpack_bf16x2(o[jj]);         // the norm sees what later consumers of x read
                  sq += r2.x * r2.x + r2.y * r2.y;
                }
                *reinterpret_cast<uint4*>(xr + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          sq = warp_sum(sq);
          if (lane == 0) s_red[warp_idx - 2] = sq;
          epi_bar_sync();
          const float rstd = rsqrtf((s_red[0] + s_red[1] + s_red[2] + s_red[3]) / (float)p.N + p.tail_eps);
          epi_bar_sync();                                        // s_red is rewritten by the next row
          for (int c0 = et; c0 < nv; c0 += 128 * TB) {
            uint4 ux[TB], uw[TB];
#pragma unroll
            for (int t = 0; t < TB; ++t) {
              const int c = c0 + t * 128;
              if (c < nv) {
                ux[t] = *reinterpret_cast<const uint4*>(xr + c * 8);          // this thread's own stores
                uw[t] = *reinterpret_cast<const uint4*>(p.tail_w + c * 8);
              }
            }
#pragma unroll
            for (int t = 0; t < TB; ++t) {
              const int c = c0 + t * 128;
              if (c < nv) {
                const uint32_t* px = reinterpret_cast<const uint32_t*>(&ux[t]);
                const uint32_t* pw = reinterpret_cast<const uint32_t*>(&uw[t]);
                uint32_t o[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float2 a = unpack_bf16x2(px[jj]), w2 = unpack_bf16x2(pw[jj]);
                  o[jj] = pack_bf16x2(a.x * rstd * w2.x, a.y * rstd * w2.y);
                }
                *reinterpret_cast<uint4*>(hr + c * 8) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
        }
        epi_bar_sync();
        if (et == 0) {                              // the last row-CTA re-arms both counters for the next launch
          const int old = atomicAdd(p.tail_counter + 1, 1);
          if (old == row_ctas - 1) {
            p.tail_counter[1] = 0;
            p.tail_counter[0] = 0;
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && et == 0) p.dbg[3] = gtimer();
  }

#ifdef T1_SG_DBG
  if (threadIdx.x == 64 + 0) g_cta_t[2][blockIdx.x] = gtimer();
#endif
  tc_fence_before();
  asm volatile("bar.sync 2, %0;" ::"n"(SG_THREADS) : "memory");   // warps 0..5 (the prefetch warp has left)
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------------------
// TWO dependent GEMMs of the decode chain in ONE launch (t1_stream_gemm_fold_pair): phase A is the producer half of a
// folded RMSNorm (o_proj or down_proj: residual add, z = norm_w * resid, per-slab row sums of squares), phase B the
// consumer half streaming against z (gate_up with SwiGLU, or the next layer's qkv).  Between two kernels that boundary
// costs the grid teardown, the memory flush and the dependent launch on top of the slowest CTA's fix-up; here it is a
// device-side arrival counter every CTA of the (fully resident) grid increments after its phase-A stores:
//   * the TMA producer never stops: phase-B WEIGHT tiles enter the ring as soon as phase-A stages retire; only the small
//     z tiles of phase B wait for the counter (acquire + fence.proxy.async before the first TMA read of z);
//   * the MMA warp runs straight through both unit ranges (same ring, same TMEM double buffer);
//   * the epilogue warps arrive after phase A, read the row factors once the counter is complete, run phase B, and the
//     last CTA to leave re-arms the counters.
// Same unit split, same arithmetic and the same summation order as t1_stream_gemm_fold followed by
// t1_stream_gemm_scaled: bit-identical results.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_shared(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

template <int MP>
__global__ void __launch_bounds__(SG_THREADS, 1)
stream_gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_w0, const __grid_constant__ CUtensorMap tma_x0,
                        const __grid_constant__ CUtensorMap tma_w1, const __grid_constant__ CUtensorMap tma_x1,
                        const __grid_constant__ SgParams p0, const __grid_constant__ SgParams p1, int* pair_counter) {
  constexpr int X_BYTES = MP * SG_BK * 2;
  constexpr int STAGE_BYTES = SG_W_BYTES + X_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * MP < 32) ? 32 : 2 * MP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + SG_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SG_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr_smem + 1);
  uint8_t* tiles = smem + SG_HEADER;
  const int stages = p0.stages;
  float* xs = reinterpret_cast<float*>(tiles + stages * STAGE_BYTES);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * p0.per_cta;
  const int u1 = max(u0, min(u0 + p0.per_cta, p0.total_units));
  const int v0 = blockIdx.x * p1.per_cta;
  const int v1 = max(v0, min(v0 + p1.per_cta, p1.total_units));
  const int nA = u1 - u0, nB = v1 - v0;
  // ONE thread per CTA (epilogue thread 0) polls the global arrival counter, with back-off: 2 x 148 threads spinning on one L2
  // line delay the very arrivals they wait for (measured: 4.41 instead of 3.20 ms per decode step).  It hands the result to the
  // TMA producer through this shared-memory flag.
  int* pair_flag = s_flag + 2;

  if (warp_idx == 0 && lane == 0) {
    *pair_flag = 0;
    tma_prefetch_desc(&tma_w0);
    tma_prefetch_desc(&tma_x0);
    tma_prefetch_desc(&tma_w1);
    tma_prefetch_desc(&tma_x1);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && nA + nB > 0) {
      uint64_t pol_stream, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      const int total = nA + nB;
      // This one thread issues every load of the CTA (a unit every ~700 cycles at the HBM rate): no division in the steady
      // state -- ring slot, wrap parity and (slab, k-block) of the W and of the x cursor advance incrementally.  (First version:
      // five run-time divisions per unit made the ISSUE the bottleneck, gate_up 67 instead of 40 us.)  Each TMA instruction
      // names its tensor map directly.
      const int KBa = p0.kblocks, KBb = p1.kblocks, swb = p1.swiglu, halfb = p1.half_n;
      const int slabB0 = v0 / KBb, kbB0 = v0 - slabB0 * KBb;
      struct Cur { int g, st, wrap, slab, kb; };
      Cur cw, cx;
      cw.g = 0; cw.st = 0; cw.wrap = 0;
      if (nA > 0) { cw.slab = u0 / KBa; cw.kb = u0 - cw.slab * KBa; } else { cw.slab = slabB0; cw.kb = kbB0; }
      cx = cw;
      auto advance = [&](Cur& c) {
        const bool was_a = c.g < nA;
        ++c.g;
        if (++c.st == stages) { c.st = 0; ++c.wrap; }
        if (was_a && c.g == nA) { c.slab = slabB0; c.kb = kbB0; }
        else if (++c.kb == (was_a ? KBa : KBb)) { c.kb = 0; ++c.slab; }
      };
      auto do_w = [&]() {
        uint8_t* dst = tiles + cw.st * STAGE_BYTES;
        uint64_t* bar = &full_bar[cw.st];
        mbar_arrive_expect_tx(bar, STAGE_BYTES);
        if (cw.g < nA) {                                 // phase A: fold producer, never SwiGLU
          tma_load_2d_hint(dst, &tma_w0, bar, cw.kb * SG_BK, cw.slab * SG_BM, pol_stream);
        } else if (swb) {
          tma_load_2d_hint(dst, &tma_w1, bar, cw.kb * SG_BK, cw.slab * 64, pol_stream);
          tma_load_2d_hint(dst + 64 * 128, &tma_w1, bar, cw.kb * SG_BK, halfb + cw.slab * 64, pol_stream);
        } else {
          tma_load_2d_hint(dst, &tma_w1, bar, cw.kb * SG_BK, cw.slab * SG_BM, pol_stream);
        }
        advance(cw);
      };
      auto do_x = [&]() {
        uint8_t* dst = tiles + cx.st * STAGE_BYTES + SG_W_BYTES;
        if (cx.g < nA) tma_load_2d_hint(dst, &tma_x0, &full_bar[cx.st], cx.kb * SG_BK, 0, pol_keep);
        else tma_load_2d_hint(dst, &tma_x1, &full_bar[cx.st], cx.kb * SG_BK, 0, pol_keep);
        advance(cx);
      };
      const bool dbg = p0.dbg && blockIdx.x == 0;
      if (dbg) p0.dbg[0] = gtimer();
      const int pre = min(total, stages);
      for (int i = 0; i < pre; ++i) do_w();              // weights first: they depend on nothing
      if (dbg) p0.dbg[1] = gtimer();
      pdl_wait();
      if (dbg) p0.dbg[2] = gtimer();
      const int preA = min(pre, nA);
      while (cx.g < preA) do_x();
      bool bar_ok = false;
      auto pass_barrier = [&](bool spin) {
        if (!spin && ld_acquire_cta_shared(pair_flag) == 0) return;
        while (ld_acquire_cta_shared(pair_flag) == 0) {}
        asm volatile("fence.proxy.async;" ::: "memory");   // z was written through the generic proxy, TMA reads it through the async proxy
        bar_ok = true;
        if (dbg && p1.dbg) p1.dbg[1] = gtimer();
      };
      // cx.g = first unit whose x tile has not been requested: == cw.g through phase A, stuck at the first phase-B unit until
      // the arrival counter is complete
      while (cw.g < total) {
        const int g = cw.g;
        if (!bar_ok && g >= nA) {
          // the slot's previous occupant must be able to retire without this thread: a phase-B unit only can once its z tile is requested
          if (g - stages >= cx.g) pass_barrier(true);
          else if (cx.g < g) pass_barrier(false);
        }
        if (bar_ok) {
          while (cx.g < g) do_x();
        }
        mbar_wait(&empty_bar[cw.st], (cw.wrap & 1) ^ 1);
        do_w();
        if (g < nA || bar_ok) do_x();
      }
      if (cx.g < total) {
        if (!bar_ok) pass_barrier(true);
        while (cx.g < total) do_x();
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SG_BM, MP);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int ph = 0; ph < 2; ++ph) {
        const int KB = ph ? p1.kblocks : p0.kblocks;
        const int a = ph ? v0 : u0, b = ph ? v1 : u1;
        for (int u = a; u < b; ++it) {
          const int slab = u / KB, kb0 = u - slab * KB;
          const int kb1 = min(KB, kb0 + (b - u));
          const int acc = it & 1;
          const uint32_t acc_phase = (it >> 1) & 1;
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * MP;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES));
            const uint64_t b_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES + SG_W_BYTES));
#pragma unroll
            for (int k = 0; k < SG_BK / 16; ++k)
              umma_bf16_ss(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(&tmem_full_bar[acc]);
          u += kb1 - kb0;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int et = (warp_idx - 2) * 32 + lane;
    pdl_wait();
    const bool dbg = p0.dbg && p1.dbg && blockIdx.x == 0 && et == 0;
    int it = sg_epilogue<MP>(p0, smem, warp_idx, lane, tmem_base, tmem_full_bar, tmem_empty_bar, s_flag, xs, u0, u1, 0);
    if (dbg) p0.dbg[3] = p1.dbg[0] = gtimer();
    __threadfence();                               // this CTA's residual / z / parts stores are visible device-wide
    epi_bar_sync();
    if (et == 0) {
      atomicAdd(pair_counter, 1);
      while (ld_acquire_gpu(pair_counter) < (int)gridDim.x) __nanosleep(64);
      __threadfence();
      st_release_cta_shared(pair_flag, 1);         // the TMA producer may request the z tiles
    }
    epi_bar_sync();
    if (dbg) p1.dbg[2] = gtimer();
    sg_epilogue<MP>(p1, smem, warp_idx, lane, tmem_base, tmem_full_bar, tmem_empty_bar, s_flag, xs, v0, v1, it);
    epi_bar_sync();
    if (dbg) p1.dbg[3] = gtimer();
    if (et == 0) {                                 // the last CTA to leave re-arms both counters for the next launch
      const int old = atomicAdd(pair_counter + 1, 1);
      if (old == (int)gridDim.x - 1) {
        pair_counter[1] = 0;
        pair_counter[0] = 0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Dynamically scheduled variant.  The 148 CTAs of a static equal split do NOT finish together: per-SM streaming rates
// differ by ~15% (measured finishing-time spread: 7 us of 49 for gate_up, 41 us of 183 for lm_head), and every later
// kernel waits for the slowest one.  Here the work is cut into items = (slab, k-segment) several times smaller than a
// CTA's share; CTA b starts with item b (so its weights can still be prefetched before griddepcontrol.wait) and then
// pulls items from an atomic counter.  The producer warp fetches one item ahead and publishes the ids to the MMA and
// epilogue warps through a 4-slot shared-memory queue.  Partials of a slab are reduced by whoever finishes its last
// segment, always in segment order (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int SG_QN = 4;

template <int MP>
__global__ void __launch_bounds__(SG_THREADS, 1)
stream_gemm_dyn_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_x, SgParams p) {
  constexpr int X_BYTES = MP * SG_BK * 2;
  constexpr int STAGE_BYTES = SG_W_BYTES + X_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * MP < 32) ? 32 : 2 * MP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + SG_MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SG_MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* q_full = tmem_empty_bar + 2;        // SG_QN
  uint64_t* q_empty = q_full + SG_QN;           // SG_QN (1 MMA arrival + 128 epilogue arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_empty + SG_QN);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr_smem + 1);
  int* q_item = s_flag + 8;                     // SG_QN ints (s_flag + 1 .. + 4 hold the norm-tail reduction scratch)
  uint8_t* tiles = smem + SG_HEADER;
  float* xs = reinterpret_cast<float*>(tiles + p.stages * STAGE_BYTES);   // swiglu: [MP][64] up values

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  const int KB = p.kblocks;
  const int first = (int)blockIdx.x < p.total_items ? (int)blockIdx.x : -1;
  auto item_range = [&](int id, int& slab, int& kb0, int& kb1) {
    slab = id / p.segs;
    const int sg = id - slab * p.segs;
    kb0 = sg * p.seg_k;
    kb1 = min(KB, kb0 + p.seg_k);
  };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_w);
    tma_prefetch_desc(&tma_x);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    for (int i = 0; i < SG_QN; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 129);
    }
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_trigger();
#ifdef T1_SG_DBG
  if (threadIdx.x == 0) {
    g_cta_t[0][blockIdx.x] = gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_cta_t[1][blockIdx.x] = smid;
  }
#endif

  if (warp_idx == 0) {
    // ===================== TMA producer + scheduler =====================
    if (lane == 0 && first >= 0) {
      const bool dbg = p.dbg && blockIdx.x == 0;
      if (dbg) p.dbg[0] = gtimer();
      uint64_t pol_stream, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      auto load_w = [&](int st, int slab, int kb) {
        if (p.swiglu) {
          tma_load_2d_hint(tiles + st * STAGE_BYTES, &tma_w, &full_bar[st], kb * SG_BK, slab * 64, pol_stream);
          tma_load_2d_hint(tiles + st * STAGE_BYTES + 64 * 128, &tma_w, &full_bar[st], kb * SG_BK, p.half_n + slab * 64, pol_stream);
        } else {
          tma_load_2d_hint(tiles + st * STAGE_BYTES, &tma_w, &full_bar[st], kb * SG_BK, slab * SG_BM, pol_stream);
        }
      };
      int stage = 0;
      uint32_t phase = 0;          // parity the consumer will wait for on `stage`
      int slab, kb0, kb1;
      item_range(first, slab, kb0, kb1);
      // the first item's weights do not depend on the previous kernel: start streaming them before the dependency resolves
      const int pre = min(kb1 - kb0, stages);
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], STAGE_BYTES);
        load_w(i, slab, kb0 + i);
      }
      if (dbg) p.dbg[1] = gtimer();
      pdl_wait();
      if (dbg) p.dbg[2] = gtimer();
      for (int i = 0; i < pre; ++i)
        tma_load_2d_hint(tiles + i * STAGE_BYTES + SG_W_BYTES, &tma_x, &full_bar[i], (kb0 + i) * SG_BK, 0, pol_keep);
      stage = pre == stages ? 0 : pre;
      phase = pre == stages ? 1 : 0;
      const int F = min((int)gridDim.x, p.total_items);
      auto fetch = [&]() {
        const int old = atomicAdd(p.sched_counter, 1);
        if (old == p.total_items - 1) *p.sched_counter = 0;       // the last fetch of the launch re-arms the counter
        const int id = F + old;
        return id < p.total_items ? id : -1;
      };
      int nxt = fetch();            // one item ahead: the round trip of the atomic hides under this item's loads
      int qn = 0;
      int kb_done = pre;
      while (true) {
        for (int kb = kb0 + kb_done; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          load_w(stage, slab, kb);
          tma_load_2d_hint(tiles + stage * STAGE_BYTES + SG_W_BYTES, &tma_x, &full_bar[stage], kb * SG_BK, 0, pol_keep);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // publish the next item (or -1 = no more work) to the MMA and epilogue warps
        const int slot = qn % SG_QN;
        mbar_wait(&q_empty[slot], ((qn / SG_QN) & 1) ^ 1);
        q_item[slot] = nxt;
        mbar_arrive(&q_full[slot]);
        ++qn;
        if (nxt < 0) break;
        item_range(nxt, slab, kb0, kb1);
        kb_done = 0;
        nxt = fetch();
      }
    } else if (lane == 0) {
      pdl_wait();
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0 && first >= 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(SG_BM, MP);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0, qn = 0;
      int item = first;
      while (item >= 0) {
        int slab, kb0, kb1;
        item_range(item, slab, kb0, kb1);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MP;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES));
          const uint64_t b_desc = umma_desc_sw128_kmajor(smem_u32(tiles + stage * STAGE_BYTES + SG_W_BYTES));
#pragma unroll
          for (int k = 0; k < SG_BK / 16; ++k)
            umma_bf16_ss(tmem_d, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        ++it;
        const int slot = qn % SG_QN;
        mbar_wait(&q_full[slot], (qn / SG_QN) & 1);
        item = q_item[slot];
        mbar_arrive(&q_empty[slot]);
        ++qn;
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 = TMEM lane quadrants 2,3,0,1) =====================
    pdl_wait();
    const int q = warp_idx & 3;
    const int et = (warp_idx - 2) * 32 + lane;   // 0..127
    const int r = q * 32 + lane;                 // weight row inside the slab == TMEM lane
    int it = 0, qn = 0;
    int item = first;
    while (item >= 0) {
      int slab, kb0, kb1;
      item_range(item, slab, kb0, kb1);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      uint32_t v[MP];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * MP;
      if constexpr (MP == 16) {
        tmem_ld_32x32b_x16(taddr, v);
      } else {
#pragma unroll
        for (int c = 0; c < MP / 32; ++c) tmem_ld_32x32b_x32(taddr + c * 32, v + c * 32);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);          // accumulator is in registers: the MMA warp may reuse it
      const int n = slab * SG_BM + r;
      bool finish = p.segs == 1;
      if (!finish) {
        float* mine = p.ws + (size_t)item * MP * SG_BM;
#pragma unroll
        for (int m = 0; m < MP; ++m) mine[m * SG_BM + r] = __uint_as_float(v[m]);
        __threadfence();
        epi_bar_sync();
        if (et == 0) {
          const int old = atomicAdd(&p.counters[slab], 1);
          const int last = (old == p.segs - 1) ? 1 : 0;
          if (last) p.counters[slab] = 0;         // self-resetting: the next launch finds zeros
          *s_flag = last;
        }
        epi_bar_sync();
        finish = *s_flag != 0;
        epi_bar_sync();                            // s_flag may be rewritten by the next item
        if (finish) {
          __threadfence();
#pragma unroll
          for (int m = 0; m < MP; ++m) v[m] = 0u;
          for (int sg = 0; sg < p.segs; ++sg) {    // fixed order: deterministic
            const float* part = p.ws + (size_t)(slab * p.segs + sg) * MP * SG_BM;
#pragma unroll
            for (int m = 0; m < MP; ++m) v[m] = __float_as_uint(__uint_as_float(v[m]) + __ldcg(part + m * SG_BM + r));
          }
        }
      }
      if (p.swiglu) {
        if (finish) {
          const int col = slab * 64 + (r & 63);
          if (r >= 64) {
            const float b = (p.bias && col < p.half_n) ? __bfloat162float(p.bias[p.half_n + col]) : 0.f;
#pragma unroll
            for (int m = 0; m < MP; ++m) xs[m * 64 + (r - 64)] = __bfloat162float(__float2bfloat16(__uint_as_float(v[m]) + b));
          }
          epi_bar_sync();
          if (r < 64 && col < p.half_n) {
            const float b = p.bias ? __bfloat162float(p.bias[col]) : 0.f;
            bf16* yp = reinterpret_cast<bf16*>(p.y) + col;
#pragma unroll
            for (int m = 0; m < MP; ++m) {
              if (m < p.M) {
                const float g = __bfloat162float(__float2bfloat16(__uint_as_float(v[m]) + b));
                const float u = xs[m * 64 + r];
                yp[(long long)m * p.ldy] = __float2bfloat16(g * (1.0f / (1.0f + __expf(-g))) * u);
              }
            }
          }
          epi_bar_sync();                          // xs is rewritten by the next finished slab
        }
      } else if (finish && n < p.N) {
        const float b = p.bias ? __bfloat162float(p.bias[n]) : 0.f;
        if (p.out_f32) {
          float* yp = reinterpret_cast<float*>(p.y) + n;
#pragma unroll
          for (int m = 0; m < MP; ++m)
            if (m < p.M) yp[(long long)m * p.ldy] = __uint_as_float(v[m]) + b;
        } else {
          bf16* yp = reinterpret_cast<bf16*>(p.y) + n;
#pragma unroll
          for (int m = 0; m < MP; ++m)
            if (m < p.M) yp[(long long)m * p.ldy] = __float2bfloat16(__uint_as_float(v[m]) + b);
        }
      }
      ++it;
      const int slot = qn % SG_QN;
      mbar_wait(&q_full[slot], (qn / SG_QN) & 1);
      item = q_item[slot];
      mbar_arrive(&q_empty[slot]);
      ++qn;
    }
    if (p.dbg && blockIdx.x == 0 && et == 0) p.dbg[3] = gtimer();
  }
#ifdef T1_SG_DBG
  if (threadIdx.x == 64 + 0) g_cta_t[2][blockIdx.x] = gtimer();
#endif
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
unsigned long long* g_dbg = nullptr;   // device ring of 4-stamp records (t1_stream_gemm_debug)
int g_dbg_slots = 0, g_dbg_next = 0;
int g_max_ctas = 0;                      // t1_stream_gemm_max_ctas(): cap on the grid (0 = one CTA per SM)
int g_dynamic = -1;                      // T1_STREAM_DYNAMIC (default 1): dynamically scheduled kernel
int g_stages = -1, g_ctas_per_sm = -1;   // tuning knobs: T1_STREAM_STAGES / T1_STREAM_CTAS_PER_SM or t1_stream_gemm_tune()

template <int MP>
int launch_stream_dyn(const CUtensorMap& tw, const CUtensorMap& tx, SgParams p, int grid, int smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stream_gemm_dyn_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      t1_set_error("t1_stream_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t e = t1_launch(stream_gemm_dyn_kernel<MP>, dim3(grid), dim3(SG_THREADS), (size_t)smem, stream, tw, tx, p);
  if (e != cudaSuccess) {
    t1_set_error("t1_stream_gemm: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  return T1_OK;
}

template <int MP>
int launch_stream(const CUtensorMap& tw, const CUtensorMap& tx, SgParams p, int grid, int smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stream_gemm_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      t1_set_error("t1_stream_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t e = t1_launch(stream_gemm_kernel<MP>, dim3(grid), dim3(SG_THREADS + (p.pf_w ? 32 : 0)), (size_t)smem, stream, tw, tx, p);
  if (e != cudaSuccess) {
    t1_set_error("t1_stream_gemm: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  return T1_OK;
}

template <int MP>
int launch_stream_pair(const CUtensorMap& tw0, const CUtensorMap& tx0, const CUtensorMap& tw1, const CUtensorMap& tx1, SgParams p0,
                       SgParams p1, int* pair_counter, int grid, int smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stream_gemm_pair_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      t1_set_error("t1_stream_gemm_fold_pair: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaError_t e = t1_launch(stream_gemm_pair_kernel<MP>, dim3(grid), dim3(SG_THREADS), (size_t)smem, stream, tw0, tx0, tw1, tx1, p0, p1,
                            pair_counter);
  if (e != cudaSuccess) {
    t1_set_error("t1_stream_gemm_fold_pair: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  return T1_OK;
}

}  // namespace

// Workspace bytes a caller must provide for t1_stream_gemm with up to max_m rows: slab counters (the first 64 KB, which
// must be ZERO before the first launch; every launch leaves them zero) followed by the fp32 partial tiles.
extern "C" long long t1_stream_gemm_workspace_bytes(int max_m) {
  const long long mp = max_m <= 16 ? 16 : (max_m <= 32 ? 32 : 64);
  // static split: two partial tiles per CTA (<= 2 CTAs per SM); dynamic scheduler: one per (slab, k-segment) item, at
  // most 8 items per SM when slabs are split
  return SG_COUNTER_BYTES + 8LL * t1_num_sms() * mp * SG_BM * 4;
}

// Diagnostics: when buf != NULL, launch i (counted from this call) makes CTA 0 write 4 %globaltimer stamps
// {kernel start, weight prefetch issued, dependency resolved, last slab written} to buf[4*(i % slots)].
extern "C" void t1_stream_gemm_debug(void* buf, int slots) {
  g_dbg = reinterpret_cast<unsigned long long*>(buf);
  g_dbg_slots = slots;
  g_dbg_next = 0;
}

// Ring depth (0 = default: ~108 KB per CTA) and CTAs per SM (1 or 2) of subsequent launches.
extern "C" void t1_stream_gemm_max_ctas(int n) { g_max_ctas = n > 0 ? n : 0; }

extern "C" void t1_stream_gemm_dynamic(int on) { g_dynamic = on < 0 ? 0 : on; }

extern "C" void t1_stream_gemm_tune(int stages, int ctas_per_sm) {
  g_stages = stages > 0 ? stages : 0;
  g_ctas_per_sm = ctas_per_sm > 0 ? ctas_per_sm : 1;
}

// The unit split of a launch (shared by the launch itself and by the L2 prefetch hint that describes the NEXT launch).
struct SgPlan {
  int MP, stage_bytes, xs_bytes, stages, smem, slabs, kblocks, per, grid, half_n;
  long long total;
};
static SgPlan sg_plan(int M, int N, int K, int swiglu, bool fold_producer) {
  SgPlan pl;
  pl.MP = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  pl.stage_bytes = SG_W_BYTES + pl.MP * SG_BK * 2;
  pl.half_n = N / 2;
  pl.xs_bytes = swiglu ? pl.MP * 64 * 4 : (fold_producer ? 4 * pl.MP * 4 : 0);
  if (g_stages < 0) {
    g_stages = env_int("T1_STREAM_STAGES", 0);
    g_ctas_per_sm = env_int("T1_STREAM_CTAS_PER_SM", 1);
  }
  const int env_stages = g_stages, env_ctas = g_ctas_per_sm;
  // default: ~180 KB of ring per CTA (10 stages at MP = 16).  Measured on B200 inside the captured 7B decode step
  // (profiles/r01_decode_timeline_v2.txt): 10 stages 3.65 ms/step, 6 stages (two kernels co-resident) 3.71 ms.
  int stages = env_stages > 0 ? env_stages : (180 * 1024) / pl.stage_bytes;
  if (stages > SG_MAX_STAGES) stages = SG_MAX_STAGES;
  while (stages > 2 && SG_HEADER + 1024 + stages * pl.stage_bytes + pl.xs_bytes > 227 * 1024) --stages;
  pl.stages = stages;
  pl.smem = SG_HEADER + 1024 + stages * pl.stage_bytes + pl.xs_bytes;
  pl.slabs = swiglu ? pl.half_n / 64 : (N + SG_BM - 1) / SG_BM;
  pl.kblocks = (K + SG_BK - 1) / SG_BK;
  pl.total = (long long)pl.slabs * pl.kblocks;
  int max_ctas = t1_num_sms() * (env_ctas > 0 ? env_ctas : 1);
  if (max_ctas > 2 * t1_num_sms()) max_ctas = 2 * t1_num_sms();
  if (g_max_ctas > 0 && g_max_ctas < max_ctas) max_ctas = g_max_ctas;     // leave SMs to a concurrent stream (experiments)
  pl.per = (int)((pl.total + max_ctas - 1) / max_ctas);
  // Small GEMMs -- the weights about fit the rings of the grid (o_proj, qkv of the decode step): nothing streams for long, the
  // launch is latency-bound and its tail is the split-slab fix-up.  Cut the unit ranges at slab-aligned points (per divides the
  // slab's k-blocks, or is a multiple of them) when that costs <= 40 % more units per CTA: every CTA then works on ONE slab, writes
  // one partial tile, and a finisher fetches all its contributors in one batch.  Measured on B200 (7B decode step, o_proj
  // 1568 units: 11 per CTA on 143 CTAs -> 14 per CTA on 112 CTAs, exactly 4 CTAs per slab): 3.21 -> 3.17 ms per decode step
  // (profiles/r02_decode_o_ctas_sweep.txt); the streaming-bound GEMMs keep the equal-bytes split (down_proj aligned: slower).
  const int align = env_int("T1_STREAM_ALIGN", 1);      // (read per call: A/B runs toggle it inside one process)
  if (align && pl.kblocks > 1 && pl.total <= 3LL * stages * max_ctas / 2) {
    for (int cand = pl.per; cand <= pl.per + (2 * pl.per + 4) / 5; ++cand) {
      if (pl.kblocks % cand == 0 || cand % pl.kblocks == 0) {
        pl.per = cand;
        break;
      }
    }
  }
  pl.grid = (int)((pl.total + pl.per - 1) / pl.per);
  return pl;
}

// One-shot hint consumed by the next statically split t1_stream_gemm* launch (see SgParams::pf_w).
struct SgPrefetchHint {
  const void* W = nullptr;
  long long ldw = 0, bytes = 0;
  int M = 0, N = 0, K = 0, swiglu = 0, fold_producer = 0;
};
static SgPrefetchHint g_pf_hint;

// While the NEXT t1_stream_gemm* launch (any variant on the static split) waits for its dependency, its idle warps
// prefetch into L2 up to max_bytes of W_next [N, K] -- exactly the units the launch AFTER it (described by M, N, K,
// swiglu = t1_stream_gemm_swiglu / _scaled(swiglu) layout, fold_producer = t1_stream_gemm_fold) reads first after
// filling its own shared-memory ring.  Purely a performance hint: results do not depend on it.  max_bytes <= 0 clears.
extern "C" int t1_stream_gemm_prefetch_next(const void* W_next, long long ldw, int M, int N, int K, int swiglu,
                                            int fold_producer, long long max_bytes) {
  g_pf_hint = SgPrefetchHint();
  if (!W_next || max_bytes <= 0) return T1_OK;
  T1_CHECK_ARG(M > 0 && M <= 64 && N > 0 && K > 0 && K % 8 == 0 && ldw >= K, "t1_stream_gemm_prefetch_next: bad shape");
  T1_CHECK_ARG(!swiglu || N % 128 == 0, "t1_stream_gemm_prefetch_next: swiglu needs N = 2*I with I %% 64 == 0");
  g_pf_hint.W = W_next;
  g_pf_hint.ldw = ldw;
  g_pf_hint.bytes = max_bytes;
  g_pf_hint.M = M;
  g_pf_hint.N = N;
  g_pf_hint.K = K;
  g_pf_hint.swiglu = swiglu ? 1 : 0;
  g_pf_hint.fold_producer = fold_producer ? 1 : 0;
  return T1_OK;
}

struct SgTail {
  void* resid = nullptr;
  const void* norm_w = nullptr;
  void* h_out = nullptr;
  float eps = 0.f;
  // folded RMSNorm (see SgParams): producer side / consumer side
  void* fold_resid = nullptr;
  void* fold_z = nullptr;
  float* fold_parts = nullptr;
  const float* in_parts = nullptr;
  int in_nparts = 0;
};

static int stream_gemm_impl(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                            long long ldy, int M, int N, int K, int out_f32, int swiglu, void* workspace,
                            long long workspace_bytes, cudaStream_t stream, SgTail tail = SgTail()) {
  T1_CHECK_ARG(x && W && y && workspace, "t1_stream_gemm: null operand");
  T1_CHECK_ARG(M > 0 && M <= 64, "t1_stream_gemm: M=%d must be in [1, 64]", M);
  T1_CHECK_ARG(N > 0 && K > 0 && K % 8 == 0, "t1_stream_gemm: K %% 8 required (N=%d K=%d)", N, K);
  T1_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K, "t1_stream_gemm: ldx/ldw must be multiples of 8 and >= K");
  T1_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)W & 15) == 0, "t1_stream_gemm: operands must be 16-byte aligned");
  if (swiglu) T1_CHECK_ARG(N % 128 == 0 && !out_f32, "t1_stream_gemm_swiglu: N = 2*I with I %% 64 == 0 and bf16 output required (N=%d)", N);
  const SgPlan pl = sg_plan(M, N, K, swiglu, tail.fold_resid != nullptr);
  const int MP = pl.MP, half_n = pl.half_n, stages = pl.stages, smem = pl.smem, slabs = pl.slabs, kblocks = pl.kblocks;
  const int per = pl.per, grid = pl.grid;
  const long long total = pl.total;
  T1_CHECK_ARG(total < (1LL << 30), "t1_stream_gemm: problem too large");
  const long long need = SG_COUNTER_BYTES + (long long)grid * 2 * MP * SG_BM * 4;
  T1_CHECK_ARG(slabs * 4 <= SG_COUNTER_BYTES, "t1_stream_gemm: N=%d exceeds the counter region", N);
  T1_CHECK_ARG(workspace_bytes >= need, "t1_stream_gemm: workspace too small (%lld < %lld)", workspace_bytes, need);
  T1_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "t1_stream_gemm: workspace must be 16-byte aligned");
  CUtensorMap tw, tx;
  int rc = t1_make_tmap_bf16(&tw, W, N, K, ldw, swiglu ? 64 : SG_BM, SG_BK);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tx, x, M, K, ldx, MP, SG_BK);
  if (rc) return rc;
  SgParams p;
  p.y = y;
  p.bias = (const bf16*)bias;
  p.counters = reinterpret_cast<int*>(workspace);
  p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + SG_COUNTER_BYTES);
  p.ldy = ldy;
  p.M = M;
  p.N = N;
  p.K = K;
  p.kblocks = kblocks;
  p.total_units = (int)total;
  p.per_cta = per;
  p.stages = stages;
  p.out_f32 = out_f32;
  p.swiglu = swiglu;
  p.half_n = half_n;
  p.tail_x = reinterpret_cast<bf16*>(tail.resid);
  p.tail_w = reinterpret_cast<const bf16*>(tail.norm_w);
  p.tail_h = reinterpret_cast<bf16*>(tail.h_out);
  p.tail_eps = tail.eps;
  p.fold_x = reinterpret_cast<bf16*>(tail.fold_resid);
  p.fold_w = reinterpret_cast<const bf16*>(tail.fold_resid ? tail.norm_w : nullptr);
  p.fold_z = reinterpret_cast<bf16*>(tail.fold_z);
  p.fold_parts = tail.fold_parts;
  p.in_parts = tail.in_parts;
  p.in_nparts = tail.in_nparts;
  p.in_eps = tail.eps;
  p.pf_w = nullptr;
  p.early_resid = env_int("T1_STREAM_EARLY_RESID", 1);
  p.fast_sync = env_int("T1_STREAM_FAST_SYNC", 1);
  const SgPrefetchHint hint = g_pf_hint;
  g_pf_hint = SgPrefetchHint();          // one-shot
  const bool folded = tail.fold_resid || tail.in_parts;       // only the statically split kernel carries the folded norm
  p.segs = 1;
  p.seg_k = kblocks;
  p.total_items = slabs;
  p.sched_counter = nullptr;
  p.tail_counter = reinterpret_cast<int*>(workspace) + SG_COUNTER_BYTES / 4 - 2;     // last two words of the counter region
  T1_CHECK_ARG(slabs * 4 <= SG_COUNTER_BYTES - 8, "t1_stream_gemm: N=%d exceeds the counter region", N);
  T1_CHECK_ARG(!tail.resid || grid <= t1_num_sms(), "t1_stream_gemm_resnorm: needs every CTA resident (grid %d > %d SMs)", grid, t1_num_sms());
  p.dbg = nullptr;
  if (g_dbg && g_dbg_slots > 0) {
    p.dbg = g_dbg + 4 * (size_t)(g_dbg_next % g_dbg_slots);
    ++g_dbg_next;
  }
  if (g_dynamic < 0) g_dynamic = env_int("T1_STREAM_DYNAMIC", 1);
  // Measured on B200 (tools/stream_gemm_spread.py, profiles/r01_stream_gemm_spread.txt): whole-slab items (no k-split)
  // pay off when there are many slabs per SM (lm_head: last CTA done at 177 us vs 187 us static); splitting slabs into
  // k-segments does not (gate_up 57 vs 50 us, down 68 vs 30 us: per-item partial tiles + the segment reduction cost more
  // than the balance gains), so mode 1 only takes GEMMs with >= 4 slabs per SM.  Mode 2 forces the k-split (experiments).
  if (g_dynamic && !tail.resid && !folded && (g_dynamic >= 2 || slabs >= 4 * t1_num_sms())) {
    // items = (slab, k-segment): at most 8 per SM, segments of >= 8 k-blocks (128 KB of weights)
    const int max_items = 8 * t1_num_sms();
    int segs = 1;
    if (g_dynamic >= 2 && slabs < max_items) {
      segs = kblocks / 8;
      if (segs > max_items / slabs) segs = max_items / slabs;
      if (segs < 1) segs = 1;
    }
    const int seg_k = (kblocks + segs - 1) / segs;
    segs = (kblocks + seg_k - 1) / seg_k;
    p.segs = segs;
    p.seg_k = seg_k;
    p.total_items = slabs * segs;
    p.sched_counter = reinterpret_cast<int*>(workspace) + SG_COUNTER_BYTES / 4 - 3;
    T1_CHECK_ARG(slabs * 4 <= SG_COUNTER_BYTES - 12, "t1_stream_gemm: N=%d exceeds the counter region", N);
    const long long need_dyn = SG_COUNTER_BYTES + (segs > 1 ? (long long)p.total_items * MP * SG_BM * 4 : 0);
    T1_CHECK_ARG(workspace_bytes >= need_dyn, "t1_stream_gemm: workspace too small (%lld < %lld)", workspace_bytes, need_dyn);
    int grid_dyn = p.total_items < t1_num_sms() ? p.total_items : t1_num_sms();
    if (g_max_ctas > 0 && g_max_ctas < grid_dyn) grid_dyn = g_max_ctas;
    if (MP == 16) rc = launch_stream_dyn<16>(tw, tx, p, grid_dyn, smem, stream);
    else if (MP == 32) rc = launch_stream_dyn<32>(tw, tx, p, grid_dyn, smem, stream);
    else rc = launch_stream_dyn<64>(tw, tx, p, grid_dyn, smem, stream);
  } else {
    if (hint.W) {        // describe the launch after this one
      const SgPlan nx = sg_plan(hint.M, hint.N, hint.K, hint.swiglu, hint.fold_producer != 0);
      if (nx.total < (1LL << 30)) {
        long long units = hint.bytes / ((long long)SG_W_BYTES * nx.grid);
        if (units > nx.per - nx.stages) units = nx.per - nx.stages;
        if (units > 0) {
          p.pf_w = reinterpret_cast<const bf16*>(hint.W);
          p.pf_ldw = hint.ldw;
          p.pf_n = hint.N;
          p.pf_kblocks = nx.kblocks;
          p.pf_per_cta = nx.per;
          p.pf_total = (int)nx.total;
          p.pf_grid = nx.grid;
          p.pf_skip = nx.stages;
          p.pf_units = (int)units;
          p.pf_swiglu = hint.swiglu;
          p.pf_half_n = nx.half_n;
          p.pf_mode = env_int("T1_STREAM_PF_MODE", 2);
        }
      }
    }
    if (MP == 16) rc = launch_stream<16>(tw, tx, p, grid, smem, stream);
    else if (MP == 32) rc = launch_stream<32>(tw, tx, p, grid, smem, stream);
    else rc = launch_stream<64>(tw, tx, p, grid, smem, stream);
  }
  if (rc) return rc;
  t1_count_launch();
  return T1_OK;
}

extern "C" int t1_stream_gemm(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                              long long ldy, int M, int N, int K, int out_f32, void* workspace, long long workspace_bytes,
                              cudaStream_t stream) {
  return stream_gemm_impl(x, ldx, W, ldw, bias, y, ldy, M, N, K, out_f32, 0, workspace, workspace_bytes, stream);
}

// y[M, I] = silu(x W_gate^T) * (x W_up^T) with W = [W_gate; W_up] of shape [2I, K] (the fused gate_up weight of the
// Qwen2 MLP, HF Q25:611-624): the gate_up projection and the SwiGLU of the decode step in one weight stream.
// Same arithmetic per element as t1_stream_gemm (bf16 output) followed by t1_swiglu; bit-identical when the stream-K
// cut points of the two launches coincide (whole slabs per CTA, e.g. the 7B decode shapes).  I % 64 == 0.
extern "C" int t1_stream_gemm_swiglu(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                                     long long ldy, int M, int I, int K, void* workspace, long long workspace_bytes,
                                     cudaStream_t stream) {
  T1_CHECK_ARG(I > 0 && I % 64 == 0, "t1_stream_gemm_swiglu: I=%d must be a positive multiple of 64", I);
  return stream_gemm_impl(x, ldx, W, ldw, bias, y, ldy, M, 2 * I, K, 0, 1, workspace, workspace_bytes, stream);
}

// y[M, N] = x W^T (bf16), then  resid[M, N] += y  (in place, bf16 add as HF Q25:805-827) and
// h_out[M, N] = rmsnorm(resid, eps) * norm_w — the projection, the residual add and the NEXT RMSNorm of a decode layer
// in one launch (after a grid-wide arrive counter, CTA b normalises row b; replaces t1_stream_gemm + t1_rmsnorm(delta = y)).  N % 8 == 0,
// resid / h_out contiguous [M, N].
extern "C" int t1_stream_gemm_resnorm(const void* x, long long ldx, const void* W, long long ldw, void* y, long long ldy,
                                      int M, int N, int K, void* resid, const void* norm_w, float eps, void* h_out,
                                      void* workspace, long long workspace_bytes, cudaStream_t stream) {
  T1_CHECK_ARG(resid && norm_w && h_out, "t1_stream_gemm_resnorm: null pointer");
  T1_CHECK_ARG(N % 8 == 0 && ldy % 8 == 0, "t1_stream_gemm_resnorm: N and ldy must be multiples of 8");
  T1_CHECK_ARG((((uintptr_t)resid | (uintptr_t)norm_w | (uintptr_t)h_out | (uintptr_t)y) & 15) == 0,
               "t1_stream_gemm_resnorm: 16-byte alignment required");
  SgTail t;
  t.resid = resid;
  t.norm_w = norm_w;
  t.h_out = h_out;
  t.eps = eps;
  return stream_gemm_impl(x, ldx, W, ldw, nullptr, y, ldy, M, N, K, 0, 0, workspace, workspace_bytes, stream, t);
}

// RMSNorm folded into the two GEMMs around it (decode chain: o_proj -> [residual, post-attention norm] -> gate_up and
// down_proj -> [residual, next input norm] -> qkv; HF Q25:805-827).  Producer half:  y = x W^T is NOT stored; instead
//   resid[M, N] += bf16(y)  (in place),  z_out[M, N] = bf16(norm_w[n] * resid[m, n]),
//   parts_out[slab][MP] = sum over the slab's 128 columns of resid[m, n]^2   (MP = 16 / 32 / 64 for M <= 16 / 32 / 64),
// so the consumer (t1_stream_gemm_scaled with x = z_out, in_parts = parts_out, in_nparts = N / 128) needs no norm kernel
// in between.  N % 128 == 0; resid / z_out contiguous [M, N].
extern "C" int t1_stream_gemm_fold(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K, void* resid,
                                   const void* norm_w, void* z_out, float* parts_out, void* workspace,
                                   long long workspace_bytes, cudaStream_t stream) {
  T1_CHECK_ARG(resid && norm_w && z_out && parts_out, "t1_stream_gemm_fold: null pointer");
  T1_CHECK_ARG(N % 128 == 0, "t1_stream_gemm_fold: N=%d must be a multiple of 128", N);
  SgTail t;
  t.fold_resid = resid;
  t.norm_w = norm_w;
  t.fold_z = z_out;
  t.fold_parts = parts_out;
  return stream_gemm_impl(x, ldx, W, ldw, nullptr, z_out /*unused*/, N, M, N, K, 0, 0, workspace, workspace_bytes, stream, t);
}

// Consumer half:  y[m, n] = r[m] * sum_k W[n, k] x[m, k] (+ bias[n]),  r[m] = rsqrt(sum_i in_parts[i][m] / K + eps),
// i.e. W . rmsnorm(h) when x = norm_w * h came from t1_stream_gemm_fold.  swiglu != 0: N = 2 * I, output [M, I] as
// t1_stream_gemm_swiglu.
extern "C" int t1_stream_gemm_scaled(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                                     long long ldy, int M, int N, int K, int swiglu, const float* in_parts, int in_nparts,
                                     float eps, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  T1_CHECK_ARG(in_parts && in_nparts > 0, "t1_stream_gemm_scaled: in_parts required");
  T1_CHECK_ARG(!swiglu || (N % 128 == 0), "t1_stream_gemm_scaled: swiglu needs N = 2*I with I %% 64 == 0");
  SgTail t;
  t.in_parts = in_parts;
  t.in_nparts = in_nparts;
  t.eps = eps;
  return stream_gemm_impl(x, ldx, W, ldw, bias, y, ldy, M, N, K, 0, swiglu ? 1 : 0, workspace, workspace_bytes, stream, t);
}

// t1_stream_gemm_fold(x, W_a -> resid, z, parts) immediately followed by t1_stream_gemm_scaled(z, W_b, parts -> y) as ONE
// launch (stream_gemm_pair_kernel): the decode chain's o_proj -> [residual, post-attention norm] -> gate_up (swiglu != 0:
// N_b = 2 * I, y [M, I]) and down_proj -> [residual, next input norm] -> qkv (bias) pairs, HF Q25:805-827.  The boundary between
// the two GEMMs is a device-side arrival counter instead of a kernel boundary; the weight stream does not stop at it.
// Same unit split and summation order as the two separate launches: bit-identical y, resid, z and parts.
// N_a % 128 == 0; every CTA of the launch must be resident at once (grid <= SM count: true for the default one CTA per SM).
extern "C" int t1_stream_gemm_fold_pair(const void* x, long long ldx, const void* W_a, long long ldw_a, int M, int N_a, int K_a,
                                        void* resid, const void* norm_w, void* z, float* parts, const void* W_b, long long ldw_b,
                                        const void* bias_b, void* y, long long ldy, int N_b, int swiglu, float eps, void* workspace,
                                        long long workspace_bytes, cudaStream_t stream) {
  T1_CHECK_ARG(x && W_a && W_b && y && workspace && resid && norm_w && z && parts, "t1_stream_gemm_fold_pair: null pointer");
  T1_CHECK_ARG(M > 0 && M <= 64, "t1_stream_gemm_fold_pair: M=%d must be in [1, 64]", M);
  T1_CHECK_ARG(N_a > 0 && N_a % 128 == 0 && K_a > 0 && K_a % 8 == 0 && N_b > 0, "t1_stream_gemm_fold_pair: N_a %% 128, K_a %% 8 required");
  T1_CHECK_ARG(ldx % 8 == 0 && ldw_a % 8 == 0 && ldw_b % 8 == 0 && ldx >= K_a && ldw_a >= K_a && ldw_b >= N_a,
               "t1_stream_gemm_fold_pair: leading dimensions must be multiples of 8 and cover K");
  T1_CHECK_ARG((((uintptr_t)x | (uintptr_t)W_a | (uintptr_t)W_b | (uintptr_t)z | (uintptr_t)workspace) & 15) == 0,
               "t1_stream_gemm_fold_pair: operands must be 16-byte aligned");
  if (swiglu) T1_CHECK_ARG(N_b % 128 == 0, "t1_stream_gemm_fold_pair: swiglu needs N_b = 2*I with I %% 64 == 0 (N_b=%d)", N_b);
  const SgPlan pa = sg_plan(M, N_a, K_a, 0, true);
  const SgPlan pb = sg_plan(M, N_b, N_a, swiglu, false);
  const int MP = pa.MP;
  const int stages = pa.stages < pb.stages ? pa.stages : pb.stages;
  const int xs_bytes = pa.xs_bytes > pb.xs_bytes ? pa.xs_bytes : pb.xs_bytes;
  const int smem = SG_HEADER + 1024 + stages * pa.stage_bytes + xs_bytes;
  const int grid = pa.grid > pb.grid ? pa.grid : pb.grid;
  T1_CHECK_ARG(pa.total < (1LL << 30) && pb.total < (1LL << 30), "t1_stream_gemm_fold_pair: problem too large");
  T1_CHECK_ARG(grid <= t1_num_sms(), "t1_stream_gemm_fold_pair: needs every CTA resident (grid %d > %d SMs)", grid, t1_num_sms());
  constexpr int HALF = SG_COUNTER_BYTES / 8;      // ints: phase A counters [0, HALF), phase B counters [HALF, 2 * HALF - 8)
  T1_CHECK_ARG(pa.slabs <= HALF && pb.slabs <= HALF - 8, "t1_stream_gemm_fold_pair: too many slabs for the counter region");
  const long long need = SG_COUNTER_BYTES + (long long)grid * 2 * MP * SG_BM * 4;
  T1_CHECK_ARG(workspace_bytes >= need, "t1_stream_gemm_fold_pair: workspace too small (%lld < %lld)", workspace_bytes, need);
  CUtensorMap tw0, tx0, tw1, tx1;
  int rc = t1_make_tmap_bf16(&tw0, W_a, N_a, K_a, ldw_a, SG_BM, SG_BK);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tx0, x, M, K_a, ldx, MP, SG_BK);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tw1, W_b, N_b, N_a, ldw_b, swiglu ? 64 : SG_BM, SG_BK);
  if (rc) return rc;
  rc = t1_make_tmap_bf16(&tx1, z, M, N_a, N_a, MP, SG_BK);
  if (rc) return rc;
  int* counters = reinterpret_cast<int*>(workspace);
  float* ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + SG_COUNTER_BYTES);
  SgParams p[2];
  for (int i = 0; i < 2; ++i) {
    const SgPlan& pl = i ? pb : pa;
    SgParams& q = p[i];
    q.y = i ? y : z;
    q.bias = i ? (const bf16*)bias_b : nullptr;
    q.ws = ws;
    q.counters = counters + (i ? HALF : 0);
    q.ldy = i ? ldy : N_a;
    q.M = M;
    q.N = i ? N_b : N_a;
    q.K = i ? N_a : K_a;
    q.kblocks = pl.kblocks;
    q.total_units = (int)pl.total;
    q.per_cta = pl.per;
    q.stages = stages;
    q.out_f32 = 0;
    q.swiglu = i ? (swiglu ? 1 : 0) : 0;
    q.half_n = pl.half_n;
    q.tail_x = nullptr;
    q.tail_w = nullptr;
    q.tail_h = nullptr;
    q.tail_eps = 0.f;
    q.tail_counter = nullptr;
    q.segs = 1;
    q.seg_k = pl.kblocks;
    q.total_items = pl.slabs;
    q.sched_counter = nullptr;
    q.dbg = nullptr;
    if (g_dbg && g_dbg_slots > 0) {      // diagnostics: one record per GEMM of the pair, as the separate launches would write
      q.dbg = g_dbg + 4 * (size_t)(g_dbg_next % g_dbg_slots);
      ++g_dbg_next;
    }
    q.fold_x = i ? nullptr : reinterpret_cast<bf16*>(resid);
    q.fold_w = i ? nullptr : reinterpret_cast<const bf16*>(norm_w);
    q.fold_z = i ? nullptr : reinterpret_cast<bf16*>(z);
    q.fold_parts = i ? nullptr : parts;
    q.in_parts = i ? parts : nullptr;
    q.in_nparts = i ? N_a / 128 : 0;
    q.in_eps = eps;
    q.pf_w = nullptr;
    q.early_resid = env_int("T1_STREAM_EARLY_RESID", 1);
    q.fast_sync = env_int("T1_STREAM_FAST_SYNC", 1);
  }
  g_pf_hint = SgPrefetchHint();
  int* pair_counter = counters + SG_COUNTER_BYTES / 4 - 5;      // words -5, -4 of the counter region (zero on entry, re-armed on exit)
  if (MP == 16) rc = launch_stream_pair<16>(tw0, tx0, tw1, tx1, p[0], p[1], pair_counter, grid, smem, stream);
  else if (MP == 32) rc = launch_stream_pair<32>(tw0, tx0, tw1, tx1, p[0], p[1], pair_counter, grid, smem, stream);
  else rc = launch_stream_pair<64>(tw0, tx0, tw1, tx1, p[0], p[1], pair_counter, grid, smem, stream);
  if (rc) return rc;
  t1_count_launch();
  return T1_OK;
}

#ifdef T1_SG_DBG
extern "C" int t1_stream_gemm_cta_times(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_cta_t, sizeof(unsigned long long) * 3 * 512) == cudaSuccess ? 0 : -2;
}
#endif
