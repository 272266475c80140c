// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N])
//
// Every dense contraction of the Qwen2-VL / Qwen2.5-VL rollout, scoring and backward passes goes through this
// kernel (reference call sites: the nn.Linear / Conv3d modules inside HF modeling_qwen2_5_vl.py that
// timer1_trainer.py:452-457 and :569-573 execute through cuBLAS).  Operands are bf16; each may be K-major (row-major
// [rows, K]) or MN-major (row-major [K, rows], flags GEMM_A_MN / GEMM_B_MN) — the backward pass consumes dY, X and W
// in place as MN-major UMMA operands (dgrad: B = W; wgrad: A = dY, B = X), no explicit transposes.
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D loads of 128x64 (A) and 256x64 (B) bf16 boxes, SWIZZLE_128B
//   warp 1      MMA issuer     : one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16)
//   warp 2      TMEM allocator : 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4..7  epilogue       : tcgen05.ld 32x32b -> registers -> bias / SwiGLU / convert -> bf16 rows staged in shared
//                                memory (SWIZZLE_128B, one [32 x 64] box per warp, double buffered) -> TMA tensor store
//                                (cp.async.bulk.tensor .. global.shared::cta, SASS UTMASTG): full 128-byte rows leave the SM
//                                as whole lines instead of 32 x 16-byte fragments of 32 different rows per warp store.
//                                fp32 outputs and accumulating (read-modify-write) outputs keep the direct register path.
// Pipelines: smem ring (full/empty mbarriers, 4 stages of 48 KB) and TMEM ring (tmem_full/tmem_empty).
#include "t1_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 128;
constexpr int BN_MAX = 256;   // output-tile width is a template parameter BN in {256, 224, 192}; 256 is the default
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
constexpr int B_BYTES_MAX = BN_MAX * BK * 2;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int EPI_BOX_BYTES = 32 * 128;                  // [32 rows x 64 bf16] staging box of one epilogue warp
constexpr int EPI_STAGE_BYTES = 4 * 2 * EPI_BOX_BYTES;   // 4 warps x 2 buffers = 32 KB
constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES_MAX) + EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum : int {
  GEMM_OUT_F32 = 1,     // D is fp32 (else bf16)
  GEMM_ACCUMULATE = 2,  // D += result (else D = result)
  GEMM_BIAS_F32 = 4,    // bias is fp32 (else bf16)
  GEMM_A_MN = 8,        // A is given as row-major [K, M] (M contiguous)
  GEMM_B_MN = 16,       // B is given as row-major [K, N] (N contiguous)
  GEMM_SWIGLU = 32,     // B = [W_gate; W_up] ([2N, K], K-major): D[M, N] = silu(A Wg^T + bg) * (A Wu^T + bu), bf16; each
                        // 256-wide tile pairs 128 gate rows with their 128 up rows (t1_gemm_swiglu)
  GEMM_TMA_STORE = 64,  // (internal) the epilogue stores through the tensor maps of D (and of the gate / up halves of D2)
  GEMM_LSE_PARTS = 128, // (t1_gemm_lse_parts) no D: per (256-column tile, row) softmax statistics of the accumulator row
  GEMM_DLOGITS = 256,   // (t1_gemm_dlogits) D = bf16(dlogp[row] * (1[col == target[row]] - exp(acc - lse[row])))
};
constexpr int MN_ATOM_BYTES = BK * 128;   // MN-major smem atom: [64 k-rows x 64 elements], SWIZZLE_128B

struct EpiParams {
  void* D;
  const void* bias;
  long long ldd;
  int flags;
  void* D2;         // SWIGLU: optional [M, 2N] bf16 copy of the gate_up projection (kept for the backward pass)
  long long ldd2;
  // fused lm_head epilogues (timer1_trainer.py:452-481: logits -> log_softmax -> gather -> entropy, never materialised here)
  const int* tgt;        // [M] target column per row
  const float* row_lse;  // [M] log-sum-exp per row   (DLOGITS)
  const float* row_g;    // [M] d loss / d logp       (DLOGITS)
};

// Per-warp state of the TMA-store epilogue: this warp's two [32 x 64] bf16 staging boxes; lane 0 owns the bulk async-groups.
struct EpiWarp {
  uint8_t* stage;
  uint32_t n_store;
};

// Epilogue of ONE accumulator tile for one epilogue warp (32 rows starting at row_base, TMEM lanes of this warp's quadrant;
// taddr_row = TMEM address of the tile's column 0 for that quadrant).  Shared by the 1-CTA and the 2-CTA (cta_group::2) kernels.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const EpiParams& epi, const CUtensorMap* tma_d_p, const CUtensorMap* tma_d2g_p,
                                                   const CUtensorMap* tma_d2u_p, uint32_t taddr_row, int row_base, int n_blk, int M,
                                                   int N, int lane, EpiWarp& ew) {
  const bool swiglu = epi.flags & GEMM_SWIGLU;
  const int BNO = swiglu ? BN / 2 : BN;
  const bool out_f32 = epi.flags & GEMM_OUT_F32;
  const bool accum = epi.flags & GEMM_ACCUMULATE;
  const bool bias_f32 = epi.flags & GEMM_BIAS_F32;
  const bool tma_out = (epi.flags & GEMM_TMA_STORE) && BN % 64 == 0;
    // ---- TMA-store path: this warp's two [32 x 64] bf16 staging boxes; lane 0 owns the bulk async-groups ----
      // pk = this lane's 64 bf16 of the chunk (32 packed words): row `lane` of the box, 16-byte units XOR-swizzled by row & 7
  auto store_box = [&](const CUtensorMap* tm, int col, int row0, const uint32_t* pk) {
    uint8_t* buf = ew.stage + (ew.n_store & 1) * EPI_BOX_BYTES;
    if (lane == 0) bulk_wait_group_read<1>();          // the group that last read this buffer (two stores ago) is done
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tm, buf, col, row0);                // rows >= M and columns >= the map's extent are clipped by the TMA unit
      bulk_commit_group();
    }
    ++ew.n_store;
  };
  const int row = row_base + lane;
  const bool dlog = epi.flags & GEMM_DLOGITS;
  int d_tgt = -1;
  float d_lse = 0.f, d_g = 0.f;
  if (dlog && row < M) {
    d_tgt = epi.tgt[row];
    d_lse = epi.row_lse[row];
    d_g = epi.row_g[row];
  }
  // dlogits: v[j] (accumulator of column col0 + j) -> dlogp * (onehot - softmax), the formula of t1_logits_backward
  auto dlog_x32 = [&](uint32_t* v, int col0) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      v[j] = __float_as_uint(d_g * (((col0 + j) == d_tgt ? 1.f : 0.f) - expf(__uint_as_float(v[j]) - d_lse)));
  };
      if (epi.flags & GEMM_LSE_PARTS) {
        // one TMEM lane == one logits row: running max, sum exp(x - max), sum exp(x - max) * x and the target's logit over this
        // tile's columns, all in this thread's registers (no cross-thread traffic); t1_lse_merge folds the tiles of a row
        float m = -INFINITY, ssum = 0.f, tsum = 0.f, tl = -INFINITY;
        const int tg = row < M ? epi.tgt[row] : -1;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr_row + c * 32, v);
          tmem_ld_wait();
          const int nv = min(32, N - col0);
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) cm = fmaxf(cm, __uint_as_float(v[j]));
          if (cm > m) {
            const float sc = expf(m - cm);
            ssum *= sc;
            tsum *= sc;
            m = cm;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nv) {
              const float x = __uint_as_float(v[j]);
              const float e = expf(x - m);
              ssum += e;
              tsum += e * x;
              if (col0 + j == tg) tl = x;
            }
          }
        }
        if (row < M) reinterpret_cast<float4*>(epi.D)[(long long)n_blk * M + row] = make_float4(m, ssum, tsum, tl);
      } else if (swiglu && tma_out) {
        // as below, 64 output columns at a time: y, and (when kept for the backward) g and u, leave through TMA stores
        const int row0 = row_base;
#pragma unroll 1
        for (int c = 0; c < BNO / 64; ++c) {
          const int col0 = n_blk * BNO + c * 64;
          if (col0 >= N) break;  // warp-uniform
          uint32_t og[32], ou[32];
          {
            uint32_t vg[32], vu[32];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              tmem_ld_32x32b_x32(taddr_row + c * 64 + hlf * 32, vg);
              tmem_ld_32x32b_x32(taddr_row + BNO + c * 64 + hlf * 32, vu);
              tmem_ld_wait();
              if (epi.bias) {      // bf16 bias [2N]: gate part then up part (columns past N hold clipped garbage: skip the loads)
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                  const int col = col0 + hlf * 32 + g8 * 8;
                  if (col < N) {
                    const uint4 bg = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + col);
                    const uint4 bu = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + N + col);
                    const uint32_t* pg = reinterpret_cast<const uint32_t*>(&bg);
                    const uint32_t* pu = reinterpret_cast<const uint32_t*>(&bu);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float2 a = unpack_bf16x2(pg[j]), b = unpack_bf16x2(pu[j]);
                      vg[g8 * 8 + 2 * j] = __float_as_uint(__uint_as_float(vg[g8 * 8 + 2 * j]) + a.x);
                      vg[g8 * 8 + 2 * j + 1] = __float_as_uint(__uint_as_float(vg[g8 * 8 + 2 * j + 1]) + a.y);
                      vu[g8 * 8 + 2 * j] = __float_as_uint(__uint_as_float(vu[g8 * 8 + 2 * j]) + b.x);
                      vu[g8 * 8 + 2 * j + 1] = __float_as_uint(__uint_as_float(vu[g8 * 8 + 2 * j + 1]) + b.y);
                    }
                  }
                }
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                og[hlf * 16 + j] = pack_bf16x2(__uint_as_float(vg[2 * j]), __uint_as_float(vg[2 * j + 1]));
                ou[hlf * 16 + j] = pack_bf16x2(__uint_as_float(vu[2 * j]), __uint_as_float(vu[2 * j + 1]));
              }
            }
          }
          if (epi.D2) {
            store_box(tma_d2g_p, col0, row0, og);
            store_box(tma_d2u_p, col0, row0, ou);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {      // y = silu(g) * u on the bf16-rounded projections, in place over og
            const float2 gg = unpack_bf16x2(og[j]), uu = unpack_bf16x2(ou[j]);
            og[j] = pack_bf16x2(gg.x * (1.0f / (1.0f + __expf(-gg.x))) * uu.x, gg.y * (1.0f / (1.0f + __expf(-gg.y))) * uu.y);
          }
          store_box(tma_d_p, col0, row0, og);
        }
      } else if (swiglu) {
        // TMEM columns [0, BN/2) hold the gate projection of output columns n_blk*BNO + j, columns [BN/2, BN) the up
        // projection of the same columns: y = silu(g) * u on the bf16-rounded projections (== gemm -> t1_swiglu)
#pragma unroll 1
        for (int c = 0; c < BNO / 32; ++c) {
          const int col0 = n_blk * BNO + c * 32;
          if (col0 >= N) break;  // warp-uniform
          uint32_t vg[32], vu[32];
          tmem_ld_32x32b_x32(taddr_row + c * 32, vg);
          tmem_ld_32x32b_x32(taddr_row + BNO + c * 32, vu);
          tmem_ld_wait();
          if (row < M) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = col0 + g8 * 8;
              if (col < N) {
                float fg[8], fu[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { fg[j] = __uint_as_float(vg[g8 * 8 + j]); fu[j] = __uint_as_float(vu[g8 * 8 + j]); }
                if (epi.bias) {      // bf16 bias [2N]: gate part then up part
                  const uint4 bg = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + col);
                  const uint4 bu = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + N + col);
                  const uint32_t* pg = reinterpret_cast<const uint32_t*>(&bg);
                  const uint32_t* pu = reinterpret_cast<const uint32_t*>(&bu);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float2 a = unpack_bf16x2(pg[j]), b = unpack_bf16x2(pu[j]);
                    fg[2 * j] += a.x; fg[2 * j + 1] += a.y; fu[2 * j] += b.x; fu[2 * j + 1] += b.y;
                  }
                }
                uint32_t og[4], ou[4], oy[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  og[j] = pack_bf16x2(fg[2 * j], fg[2 * j + 1]);
                  ou[j] = pack_bf16x2(fu[2 * j], fu[2 * j + 1]);
                  const float2 gg = unpack_bf16x2(og[j]), uu = unpack_bf16x2(ou[j]);
                  oy[j] = pack_bf16x2(gg.x * (1.0f / (1.0f + __expf(-gg.x))) * uu.x, gg.y * (1.0f / (1.0f + __expf(-gg.y))) * uu.y);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(epi.D) + (long long)row * epi.ldd + col) = make_uint4(oy[0], oy[1], oy[2], oy[3]);
                if (epi.D2) {
                  bf16* d2 = reinterpret_cast<bf16*>(epi.D2) + (long long)row * epi.ldd2;
                  *reinterpret_cast<uint4*>(d2 + col) = make_uint4(og[0], og[1], og[2], og[3]);
                  *reinterpret_cast<uint4*>(d2 + N + col) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
                }
              }
            }
          }
        }
      } else if (tma_out) {
        const int row0 = row_base;
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          const int col0 = n_blk * BN + c * 64;
          if (col0 >= N) break;  // warp-uniform
          uint32_t pk[32];
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr_row + c * 64 + hlf * 32, v);
            tmem_ld_wait();
            if (epi.bias) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = col0 + hlf * 32 + g * 8;
                if (col < N) {
                  float b8[8];
                  if (bias_f32) {
                    const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.bias) + col);
                    const float4 b0 = bp[0], b1 = bp[1];
                    b8[0] = b0.x; b8[1] = b0.y; b8[2] = b0.z; b8[3] = b0.w; b8[4] = b1.x; b8[5] = b1.y; b8[6] = b1.z; b8[7] = b1.w;
                  } else {
                    const uint4 b = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + col);
                    float2 t;
                    t = unpack_bf16x2(b.x); b8[0] = t.x; b8[1] = t.y;
                    t = unpack_bf16x2(b.y); b8[2] = t.x; b8[3] = t.y;
                    t = unpack_bf16x2(b.z); b8[4] = t.x; b8[5] = t.y;
                    t = unpack_bf16x2(b.w); b8[6] = t.x; b8[7] = t.y;
                  }
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[g * 8 + j] = __float_as_uint(__uint_as_float(v[g * 8 + j]) + b8[j]);
                }
              }
            }
            if (dlog) dlog_x32(v, col0 + hlf * 32);
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[hlf * 16 + j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          }
          store_box(tma_d_p, col0, row0, pk);
        }
      } else
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {      // BN is a multiple of 32 for every instantiated width
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr_row + c * 32, v);
        tmem_ld_wait();
        if (dlog) dlog_x32(v, col0);
        if (row < M) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < N) {  // N % 8 == 0 is enforced on the host
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
              if (epi.bias) {
                if (bias_f32) {
                  const float4* bp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.bias) + col);
                  float4 b0 = bp[0], b1 = bp[1];
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                } else {
                  uint4 b = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(epi.bias) + col);
                  float2 t;
                  t = unpack_bf16x2(b.x); f[0] += t.x; f[1] += t.y;
                  t = unpack_bf16x2(b.y); f[2] += t.x; f[3] += t.y;
                  t = unpack_bf16x2(b.z); f[4] += t.x; f[5] += t.y;
                  t = unpack_bf16x2(b.w); f[6] += t.x; f[7] += t.y;
                }
              }
              if (out_f32) {
                float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(epi.D) + (long long)row * epi.ldd + col);
                if (accum) {
                  float4 o0 = dp[0], o1 = dp[1];
                  f[0] += o0.x; f[1] += o0.y; f[2] += o0.z; f[3] += o0.w;
                  f[4] += o1.x; f[5] += o1.y; f[6] += o1.z; f[7] += o1.w;
                }
                dp[0] = make_float4(f[0], f[1], f[2], f[3]);
                dp[1] = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(epi.D) + (long long)row * epi.ldd + col);
                if (accum) {
                  uint4 o = *dp;
                  float2 t;
                  t = unpack_bf16x2(o.x); f[0] += t.x; f[1] += t.y;
                  t = unpack_bf16x2(o.y); f[2] += t.x; f[3] += t.y;
                  t = unpack_bf16x2(o.z); f[4] += t.x; f[5] += t.y;
                  t = unpack_bf16x2(o.w); f[6] += t.x; f[7] += t.y;
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]);
                o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]);
                o.w = pack_bf16x2(f[6], f[7]);
                *dp = o;
              }
            }
          }
        }
      }
}

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_d2g,
                    const __grid_constant__ CUtensorMap tma_d2u, int M, int N, int K, int group_m, EpiParams epi) {
  static_assert(BN % 16 == 0 && BN <= BN_MAX && (!B_MN || BN % 64 == 0), "tile width");
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tiles = smem;
  uint8_t* b_tiles = smem + STAGES * A_BYTES;
  uint8_t* epi_stage = smem + STAGES * (A_BYTES + B_BYTES_MAX);          // 1024-byte aligned (48 KB stages)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const bool swiglu = epi.flags & GEMM_SWIGLU;
  const int BNO = swiglu ? BN / 2 : BN;                    // output columns per tile
  const int num_n = (N + BNO - 1) / BNO;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;
  // Tile rasterisation: m-blocks are walked in groups of group_m; inside a group the n-blocks advance slowest, so the
  // 148 tiles in flight share group_m A row-blocks and ~148/group_m B row-blocks.  group_m = num_m is plain m-fastest.
  const int tiles_per_group = group_m * num_n;
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int g = tile / tiles_per_group, rem = tile - g * tiles_per_group;
    const int m0 = g * group_m;
    const int gm = min(group_m, num_m - m0);
    n_blk = rem / gm;
    m_blk = m0 + (rem - n_blk * gm);
  };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    mbar_fence_init();
  }
  if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);      // two BN-column accumulators (2 * BN <= 512)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if constexpr (A_MN) {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a)
              tma_load_2d(a_tiles + stage * A_BYTES + a * MN_ATOM_BYTES, &tma_a, &full_bar[stage], m_blk * BM + a * 64, kb * BK);
          } else {
            tma_load_2d(a_tiles + stage * A_BYTES, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int b = 0; b < BN / 64; ++b)
              tma_load_2d(b_tiles + stage * B_BYTES + b * MN_ATOM_BYTES, &tma_b, &full_bar[stage], n_blk * BN + b * 64, kb * BK);
          } else if (swiglu) {
            tma_load_2d(b_tiles + stage * B_BYTES, &tma_b, &full_bar[stage], kb * BK, n_blk * BNO);                          // gate rows
            tma_load_2d(b_tiles + stage * B_BYTES + BNO * BK * 2, &tma_b, &full_bar[stage], kb * BK, N + n_blk * BNO);       // up rows
          } else {
            tma_load_2d(b_tiles + stage * B_BYTES, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_tiles + stage * A_BYTES), b_addr = smem_u32(b_tiles + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 32 B (= 16 bf16) along K inside the 128 B swizzle row; MN-major: advance 16 k-rows
            // (2 KB) inside the [64 x 64] atoms, 64-element M/N blocks one atom (8 KB) apart
            const uint64_t a_desc = A_MN ? umma_desc_sw128_mnmajor(a_addr + k * UMMA_K * 128, MN_ATOM_BYTES, 1024)
                                         : umma_desc_sw128_kmajor(a_addr + k * 32);
            const uint64_t b_desc = B_MN ? umma_desc_sw128_mnmajor(b_addr + k * UMMA_K * 128, MN_ATOM_BYTES, 1024)
                                         : umma_desc_sw128_kmajor(b_addr + k * 32);
            umma_bf16_ss(tmem_d, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    const int q = warp_idx & 3;  // TMEM lane quadrant this warp may access
    EpiWarp ew{epi_stage + q * 2 * EPI_BOX_BYTES, 0u};
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN>(epi, &tma_d, &tma_d2g, &tma_d2u, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, m_blk * BM + q * 32, n_blk,
                             M, N, lane, ew);
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
    if ((epi.flags & GEMM_TMA_STORE) && lane == 0) bulk_wait_group<0>();      // every store of this warp has landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =====================================================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a thread-block CLUSTER of two CTAs (an SM pair) computes one 256 x 256 output tile.
// UMMA M = 256: each CTA supplies its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), the MMA (issued by
// the leader CTA only) reads both CTAs' shared memory and writes 128 accumulator rows x 256 columns into EACH CTA's TMEM.
// Per SM and per k-block that is 32 KB of operands for 2 x 128 x 256 x 64 MACs instead of 48 KB: the SS MMA is fed from
// shared memory, and the 1-CTA form is shared-memory-bandwidth bound (its tensor pipe peaks at ~82 %).
//   both CTAs, warp 0 : TMA producer for its halves; every load completes on the LEADER's full barrier
//                       (cp.async.bulk.tensor ... .cta_group::2, barrier address with the peer bit cleared)
//   leader,   warp 1  : tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster releases the smem stage / publishes
//                       the accumulator in BOTH CTAs
//   both CTAs, warp 2 : tcgen05.alloc.cta_group::2 (512 columns: two 256-column accumulators)
//   both CTAs, w 4..7 : epilogue of the CTA's own 128 rows (same code as the 1-CTA kernel); "accumulator free" arrives on the
//                       leader's barrier (mbarrier.arrive.shared::cluster)
// =====================================================================================================================
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = 128 * BK * 2;                       // this CTA's half of the widest (256-column) B tile
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;             // 32 KB
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + EPI_STAGE_BYTES + 1024 + 256;
static_assert(SMEM2_BYTES <= 227 * 1024, "shared memory budget (2-CTA kernel)");
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;              // shared::cluster address of the same offset in the pair's CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* desc, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm_multicast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// BN = output-tile width in {256, 192, 128}: narrower tiles quantise better when N / 256 leaves a thin last wave (N = 3584 is
// 14 tiles of 256 but 18.7 of 192); each CTA then loads BN / 2 rows of B per k-block.  SwiGLU pairs 128 + 128 columns: BN = 256.
template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_d2g,
                     const __grid_constant__ CUtensorMap tma_d2u, int M, int N, int K, int group_m, EpiParams epi) {
  static_assert(BN % 64 == 0 && BN <= 256 && (!B_MN || (BN / 2) % 64 == 0), "2-CTA tile width");
  constexpr int HB = BN / 2;                                  // B rows this CTA loads per k-block
  constexpr int STAGE_TX = A_BYTES + HB * BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tiles = smem;
  uint8_t* b_tiles = smem + STAGES2 * A_BYTES;
  uint8_t* epi_stage = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + EPI_STAGE_BYTES);   // used in the leader CTA
  uint64_t* empty_bar = full_bar + STAGES2;                                          // each CTA: released by the multicast commit
  uint64_t* tmem_full_bar = empty_bar + STAGES2;                                     // each CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                      // leader: 2 x 128 epilogue threads
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const bool swiglu = epi.flags & GEMM_SWIGLU;
  const int BNO = swiglu ? BN / 2 : BN;
  const int num_m = (M + 2 * BM - 1) / (2 * BM);           // 256-row tiles
  const int num_n = (N + BNO - 1) / BNO;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;
  const int tiles_per_group = group_m * num_n;
  auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
    const int g = tile / tiles_per_group, rem = tile - g * tiles_per_group;
    const int m0 = g * group_m;
    const int gm = min(group_m, num_m - m0);
    n_blk = rem / gm;
    m_blk = m0 + (rem - n_blk * gm);
  };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 256);
    }
    mbar_fence_init();
  }
  if (warp_idx == 2) tmem_alloc_2sm<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer's barriers are initialised and its TMEM is allocated before anything remote happens
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, m_blk, n_blk);
        const int m0 = m_blk * 2 * BM + (int)rank * BM;                              // this CTA's 128 A rows
        // this CTA's half of the B tile: plain = rows n_blk*256 + rank*128; SwiGLU = rank 0 the gate rows, rank 1 the up rows
        const int n0 = swiglu ? (int)rank * N + n_blk * BNO : n_blk * BN + (int)rank * HB;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_TX);    // both CTAs' bytes land on this barrier
          if constexpr (A_MN) {
#pragma unroll
            for (int a = 0; a < BM / 64; ++a)
              tma_load_2d_2sm(a_tiles + stage * A_BYTES + a * MN_ATOM_BYTES, &tma_a, &full_bar[stage], m0 + a * 64, kb * BK);
          } else {
            tma_load_2d_2sm(a_tiles + stage * A_BYTES, &tma_a, &full_bar[stage], kb * BK, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int b = 0; b < HB / 64; ++b)
              tma_load_2d_2sm(b_tiles + stage * B2_BYTES + b * MN_ATOM_BYTES, &tma_b, &full_bar[stage], n0 + b * 64, kb * BK);
          } else {
            tma_load_2d_2sm(b_tiles + stage * B2_BYTES, &tma_b, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA, single thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_tiles + stage * A_BYTES), b_addr = smem_u32(b_tiles + stage * B2_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_desc = A_MN ? umma_desc_sw128_mnmajor(a_addr + k * UMMA_K * 128, MN_ATOM_BYTES, 1024)
                                         : umma_desc_sw128_kmajor(a_addr + k * 32);
            const uint64_t b_desc = B_MN ? umma_desc_sw128_mnmajor(b_addr + k * UMMA_K * 128, MN_ATOM_BYTES, 1024)
                                         : umma_desc_sw128_kmajor(b_addr + k * 32);
            umma_bf16_ss_2sm(tmem_d, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm_multicast(&empty_bar[stage]);      // frees this stage in both CTAs once the MMAs retire
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm_multicast(&tmem_full_bar[acc]);      // accumulator complete -> both epilogues
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (both CTAs: own 128 rows of the 256-row tile) =====================
    const int q = warp_idx & 3;
    EpiWarp ew{epi_stage + q * 2 * EPI_BOX_BYTES, 0u};
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      int m_blk, n_blk;
      tile_coords(tile, m_blk, n_blk);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      gemm_epilogue_tile<BN>(epi, &tma_d, &tma_d2g, &tma_d2u, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN,
                             m_blk * 2 * BM + (int)rank * BM + q * 32, n_blk, M, N, lane, ew);
      tc_fence_before();
      mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
    if ((epi.flags & GEMM_TMA_STORE) && lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // the peer has read its accumulators and issued its last remote arrive
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace

static int g_gemm_max_ctas = 0;   // t1_gemm_max_ctas(): cap on the persistent grid (0 = one CTA per SM)
extern "C" void t1_gemm_max_ctas(int n) { g_gemm_max_ctas = n > 0 ? n : 0; }

static int gemm_impl(const void* A, long long lda, const void* B, long long ldb, void* D, long long ldd, int M, int N, int K,
                     const void* bias, int flags, void* D2, long long ldd2, cudaStream_t stream, const int* tgt = nullptr,
                     const float* row_lse = nullptr, const float* row_g = nullptr) {
  T1_CHECK_ARG(A && B && D, "t1_gemm_bf16_tn: null operand");
  T1_CHECK_ARG(M > 0 && N > 0 && K > 0, "t1_gemm_bf16_tn: empty problem M=%d N=%d K=%d", M, N, K);
  T1_CHECK_ARG(N % 8 == 0 && ldd % 8 == 0, "t1_gemm_bf16_tn: N and ldd must be multiples of 8 (N=%d ldd=%lld)", N, ldd);
  T1_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "t1_gemm_bf16_tn: lda/ldb must be multiples of 8 (16-byte TMA strides)");
  T1_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)D & 15) == 0,
               "t1_gemm_bf16_tn: operands must be 16-byte aligned");
  const bool a_mn = flags & GEMM_A_MN, b_mn = flags & GEMM_B_MN;
  T1_CHECK_ARG(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K) && ldd >= N,
               "t1_gemm_bf16_tn: leading dimension smaller than row length");
  T1_CHECK_ARG(!a_mn || M % 8 == 0, "t1_gemm_bf16_tn: MN-major A needs M %% 8 == 0 (M=%d)", M);
  // ---- output-tile width ----
  // 224 / 192-wide tiles fill the last wave better for N = 3584 (4.35 -> 4.97 waves), but measured on B200 the whole step's
  // GEMM time went 626.9 -> 637.3 ms with the fullest-last-wave choice: the narrower SS MMAs are shared-memory bound
  // (more operand bytes per FLOP).  256 stays the default; T1_GEMM_BN = 224 | 192 selects the others for experiments.
  const int num_m = (M + BM - 1) / BM;
  const int sms = t1_num_sms();
  int BN = 256;
  if (!(flags & GEMM_LSE_PARTS)) {
    static const char* env = getenv("T1_GEMM_BN");
    if (env && (atoi(env) == 192 || (atoi(env) == 224 && !b_mn))) BN = atoi(env);
  }
  CUtensorMap ta, tb;
  int rc = a_mn ? t1_make_tmap_bf16(&ta, A, K, M, lda, BK, 64) : t1_make_tmap_bf16(&ta, A, M, K, lda, BM, BK);
  if (rc) return rc;
  const bool swiglu = flags & GEMM_SWIGLU;
  if (swiglu) {
    T1_CHECK_ARG(!a_mn && !b_mn && !(flags & (GEMM_OUT_F32 | GEMM_ACCUMULATE | GEMM_BIAS_F32)) && BN == 256,
                 "t1_gemm_swiglu: K-major operands, bf16 output, bf16 bias, no accumulate");
    T1_CHECK_ARG(!D2 || (ldd2 % 8 == 0 && ldd2 >= 2LL * N && ((uintptr_t)D2 & 15) == 0), "t1_gemm_swiglu: bad gate_up output");
  }
  rc = b_mn ? t1_make_tmap_bf16(&tb, B, K, N, ldb, BK, 64)
            : (swiglu ? t1_make_tmap_bf16(&tb, B, 2LL * N, K, ldb, BN / 2, BK) : t1_make_tmap_bf16(&tb, B, N, K, ldb, BN, BK));
  if (rc) return rc;
  const int BNO = swiglu ? BN / 2 : BN;                    // output columns per tile
  const int num_n = (N + BNO - 1) / BNO;
  const int num_tiles = num_m * num_n;
  int grid = num_tiles < sms ? num_tiles : sms;
  if (g_gemm_max_ctas > 0 && g_gemm_max_ctas < grid) grid = g_gemm_max_ctas;     // leave the other SMs to a concurrent stream
  // TMA-store epilogue for bf16, non-accumulating outputs (T1_GEMM_EPI=direct restores the register path for A/B runs)
  CUtensorMap td = ta, tdg = ta, tdu = ta;
  {
    static const char* env = getenv("T1_GEMM_EPI");
    const bool want = !(env && env[0] == 'd') && !(flags & (GEMM_OUT_F32 | GEMM_ACCUMULATE | GEMM_LSE_PARTS)) && BN % 64 == 0;
    if (want) {
      rc = t1_make_tmap_bf16(&td, D, M, N, ldd, 32, 64);
      if (rc) return rc;
      if (swiglu && D2) {
        rc = t1_make_tmap_bf16(&tdg, D2, M, N, ldd2, 32, 64);
        if (rc) return rc;
        rc = t1_make_tmap_bf16(&tdu, reinterpret_cast<const bf16*>(D2) + N, M, N, ldd2, 32, 64);
        if (rc) return rc;
      }
      flags |= GEMM_TMA_STORE;
    }
  }
  EpiParams epi{D, bias, ldd, flags, swiglu ? D2 : nullptr, ldd2, tgt, row_lse, row_g};
  // DRAM-traffic model for the raster order (L2 = 126 MB; measured with ncu on the gate_up wgrad, M=37888 N=3584 K=5888:
  // m-fastest re-read dY 14x = 6.3 GB per launch).  An operand that fits in ~48 MB of L2 is read once whatever the order;
  // otherwise keep a group of A row-blocks resident (<= 40 MB) while B streams past it.
  int group_m = num_m;
  {
    const double a_bytes = 2.0 * M * K, b_bytes = 2.0 * N * K * (swiglu ? 2 : 1), cap = 48e6;
    if (a_bytes > cap) {
      int gm = (int)(40e6 / (2.0 * BM * K));
      if (gm < 1) gm = 1;
      if (gm > num_m) gm = num_m;
      const double grouped = a_bytes + (b_bytes <= cap ? b_bytes : b_bytes * ((num_m + gm - 1) / gm));
      const double waves_n = (double)num_n * num_m / sms / (num_m < sms ? 1.0 : (double)num_m / sms);
      const double m_fast = a_bytes * (num_m >= sms ? num_n : (waves_n < 1.0 ? 1.0 : waves_n)) + b_bytes;
      if (grouped < m_fast) group_m = gm;
    }
    static const char* env = getenv("T1_GEMM_GROUP_M");
    if (env && atoi(env) > 0) group_m = atoi(env) < num_m ? atoi(env) : num_m;
  }
  cudaError_t le = cudaSuccess;
  // ---- 2-CTA (cta_group::2) kernel: 256 x 256 tiles on SM pairs.  T1_GEMM_2CTA=0 keeps everything on the 1-CTA kernel ----
  {
    static const char* env2 = getenv("T1_GEMM_2CTA");
    // a cap at or below half the machine means "keep to a few SMs beside another stream's kernels" (the reference-model prefill
    // beside the decode, RunConfig.side_ctas): which SMs are free is the other stream's choice and they need not form pairs, so
    // those launches stay on the 1-CTA kernel (a cluster would wait for a whole pair)
    const bool few = g_gemm_max_ctas > 0 && g_gemm_max_ctas <= sms / 2;
    const bool want2 = !(env2 && env2[0] == '0') && !few && BN == 256 && M > BM && (!a_mn || M % 64 == 0) && (!b_mn || N % 64 == 0);
    if (want2) {
      const int num_m2 = (M + 2 * BM - 1) / (2 * BM);
      int max_clusters = sms / 2;
      if (g_gemm_max_ctas > 0) {
        // SMs are reserved for a concurrent stream (NCCL): each of them may break a different SM pair, and a cluster that finds
        // no free pair would only start when another one retires — i.e. after the whole persistent loop
        int free_pairs = sms / 2 - (sms - g_gemm_max_ctas);
        if (free_pairs < 1) free_pairs = 1;
        if (free_pairs < max_clusters) max_clusters = free_pairs;
      }
      // tile width: the persistent loop runs ceil(tiles / clusters) waves of ~(BN + overhead) each; take the cheapest — but only
      // for short problems (few rows: the completion-row scoring pass; or short K: the ViT).  Measured on B200
      // (profiles/r02_gemm2_tile_width.txt): 1672 x 3584 x 3584 0.060 -> 0.048 ms, 16384 x 3840 x 1280 0.132 -> 0.110 ms, but
      // 4216 x 3584 x 18944 0.536 -> 0.611 ms and the whole step 1407 -> 1441 ms when every shape may narrow (more A re-reads)
      int bn2 = 256;
      if (!swiglu && !(flags & GEMM_LSE_PARTS) && (M <= 2048 || K <= 2048)) {      // (the statistics tiles are 256 columns by contract)
        static const char* envbn = getenv("T1_GEMM2_BN");
        double best = 1e30;
        for (int cand : {256, 192, 128}) {
          if (b_mn && (cand / 2) % 64) continue;
          if (envbn && atoi(envbn) != cand) continue;
          const long long tiles = (long long)num_m2 * ((N + cand - 1) / cand);
          const double cost = (double)((tiles + max_clusters - 1) / max_clusters) * (cand + 24);
          if (cost < best - 1e-9) best = cost, bn2 = cand;
        }
      }
      const int BNO2 = swiglu ? 128 : bn2;
      const int num_n2 = (N + BNO2 - 1) / BNO2;
      CUtensorMap tb2;
      rc = b_mn ? t1_make_tmap_bf16(&tb2, B, K, N, ldb, BK, 64)
                : (swiglu ? t1_make_tmap_bf16(&tb2, B, 2LL * N, K, ldb, 128, BK) : t1_make_tmap_bf16(&tb2, B, N, K, ldb, bn2 / 2, BK));
      if (rc) return rc;
      const long long tiles2 = (long long)num_m2 * num_n2;
      const int clusters = tiles2 < max_clusters ? (int)tiles2 : max_clusters;
      int group_m2 = (group_m + 1) / 2;
      if (group_m2 > num_m2) group_m2 = num_m2;
      if (group_m2 < 1) group_m2 = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * clusters);
      cfg.blockDim = dim3(NUM_THREADS);
      cfg.dynamicSmemBytes = SMEM2_BYTES;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
#define T1_GEMM2_LAUNCH(AMN, BMN, BNV)                                                                                          \
  do {                                                                                                                          \
    static bool attr_set2 = false;                                                                                              \
    if (!attr_set2) {                                                                                                           \
      le = cudaFuncSetAttribute(gemm2_bf16_tn_kernel<AMN, BMN, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES); \
      attr_set2 = (le == cudaSuccess);                                                                                          \
    }                                                                                                                           \
    if (le == cudaSuccess)                                                                                                      \
      le = cudaLaunchKernelEx(&cfg, gemm2_bf16_tn_kernel<AMN, BMN, BNV>, ta, tb2, td, tdg, tdu, M, N, K, group_m2, epi);          \
  } while (0)
#define T1_GEMM2_BY_BN(AMN, BMN)                         \
  do {                                                   \
    if (bn2 == 256) T1_GEMM2_LAUNCH(AMN, BMN, 256);      \
    else if (bn2 == 128) T1_GEMM2_LAUNCH(AMN, BMN, 128); \
    else T1_GEMM2_LAUNCH(AMN, false, 192);               \
  } while (0)
      if (a_mn && b_mn) T1_GEMM2_BY_BN(true, true);
      else if (a_mn) T1_GEMM2_BY_BN(true, false);
      else if (b_mn) T1_GEMM2_BY_BN(false, true);
      else T1_GEMM2_BY_BN(false, false);
#undef T1_GEMM2_BY_BN
#undef T1_GEMM2_LAUNCH
      if (le != cudaSuccess) {
        t1_set_error("t1_gemm_bf16_tn (2-CTA): launch failed: %s", cudaGetErrorString(le));
        return T1_ERR_CUDA;
      }
      t1_count_launch();
      return T1_OK;
    }
  }
#define T1_GEMM_LAUNCH(AMN, BMN, BNV)                                                                                         \
  do {                                                                                                                        \
    static bool attr_set = false;                                                                                             \
    if (!attr_set) {                                                                                                          \
      le = cudaFuncSetAttribute(gemm_bf16_tn_kernel<AMN, BMN, BNV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); \
      attr_set = (le == cudaSuccess);                                                                                         \
    }                                                                                                                         \
    if (le == cudaSuccess)                                                                                                    \
      gemm_bf16_tn_kernel<AMN, BMN, BNV><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, td, tdg, tdu, M, N, K, group_m, epi); \
  } while (0)
#define T1_GEMM_BY_MAJOR(BNV)                                   \
  do {                                                          \
    if (a_mn && b_mn) T1_GEMM_LAUNCH(true, true, BNV);          \
    else if (a_mn) T1_GEMM_LAUNCH(true, false, BNV);            \
    else if (b_mn) T1_GEMM_LAUNCH(false, true, BNV);            \
    else T1_GEMM_LAUNCH(false, false, BNV);                     \
  } while (0)
  if (BN == 256) T1_GEMM_BY_MAJOR(256);
  else if (BN == 192) T1_GEMM_BY_MAJOR(192);
  else {
    if (a_mn) T1_GEMM_LAUNCH(true, false, 224);
    else T1_GEMM_LAUNCH(false, false, 224);
  }
#undef T1_GEMM_BY_MAJOR
#undef T1_GEMM_LAUNCH
  if (le != cudaSuccess) {
    t1_set_error("t1_gemm_bf16_tn: cudaFuncSetAttribute: %s", cudaGetErrorString(le));
    return T1_ERR_CUDA;
  }
  T1_CUDA_LAUNCH_CHECK("t1_gemm_bf16_tn");
  return T1_OK;
}

// C-ABI: see include/timer1.h
extern "C" int t1_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* D, long long ldd,
                               int M, int N, int K, const void* bias, int flags, cudaStream_t stream) {
  T1_CHECK_ARG(!(flags & GEMM_SWIGLU), "t1_gemm_bf16_tn: use t1_gemm_swiglu");
  return gemm_impl(A, lda, B, ldb, D, ldd, M, N, K, bias, flags, nullptr, 0, stream);
}

// act[M, I] = silu(A Wg^T + bg) * (A Wu^T + bu) with W = [Wg; Wu] ([2I, K], K-major) and bias [2I] bf16 (nullable): the
// gate_up projection of the Qwen2 MLP (HF Q25:611-624; the ViT MLP Q25:82-88 has biases) with the SwiGLU in the GEMM
// epilogue; gate_up (nullable, [M, 2I] bf16) receives the projection itself when the backward pass needs it.
// Bit-identical to t1_gemm_bf16_tn (bf16) followed by t1_swiglu.  I % 8 == 0.
extern "C" int t1_gemm_swiglu(const void* A, long long lda, const void* W, long long ldw, const void* bias, void* act,
                              long long ld_act, void* gate_up, long long ld_gate_up, int M, int I, int K, cudaStream_t stream) {
  return gemm_impl(A, lda, W, ldw, act, ld_act, M, I, K, bias, GEMM_SWIGLU, gate_up, ld_gate_up, stream);
}

// ---- lm_head fused with the log-softmax statistics (timer1_trainer.py:452-481): logits never reach HBM ----
// parts[ceil(N / 256)][M][4] fp32 = per 256-column tile and row: {max, sum exp(x - max), sum exp(x - max) * x, x[target[row]] if the
// target falls into the tile else -inf} of x = A B^T (fp32 accumulators).  t1_lse_merge folds the tiles into logp / entropy / lse.
extern "C" int t1_gemm_lse_parts(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const int* target,
                                 float* parts, cudaStream_t stream) {
  T1_CHECK_ARG(target && parts && ((uintptr_t)parts & 15) == 0, "t1_gemm_lse_parts: target / parts (16-byte aligned) required");
  // D is only used as the parts pointer; ldd is irrelevant (pass N to satisfy the generic checks)
  return gemm_impl(A, lda, B, ldb, parts, (N + 7) / 8 * 8, M, N, K, nullptr, GEMM_OUT_F32 | GEMM_LSE_PARTS, nullptr, 0, stream, target);
}

// dlogits[M, N] (bf16) = dlogp[row] * (1[col == target[row]] - exp(x - lse[row])) with x = A B^T straight from the accumulator:
// the recomputed lm_head GEMM of the backward with t1_logits_backward in its epilogue (no fp32 logits in HBM).
extern "C" int t1_gemm_dlogits(const void* A, long long lda, const void* B, long long ldb, void* dlogits, long long ldd, int M, int N,
                               int K, const int* target, const float* lse, const float* dlogp, cudaStream_t stream) {
  T1_CHECK_ARG(target && lse && dlogp, "t1_gemm_dlogits: target / lse / dlogp required");
  return gemm_impl(A, lda, B, ldb, dlogits, ldd, M, N, K, nullptr, GEMM_DLOGITS, nullptr, 0, stream, target, lse, dlogp);
}
