// Single-launch tcgen05 decode attention for the captured decode step (replaces HF GenerationMixin._sample's per-token
// attention, driven from src/time_r1/rl/timer1_trainer.py:568-573): one thread-block CLUSTER per (prompt-group, KV head).
//
// The G sequences of a group share ONE copy of the prompt K/V, so the (G x q-heads-per-kv-head) <= 64 query rows form
// one query tile against every prefix key.  The cached keys [prompt prefix | per-sequence completion slots < step] are
// cut into 128-key units and dealt evenly to the S CTAs of the cluster (equal bytes per SM).  Per CTA:
//   warp 0       TMA producer : K and V unit tiles ([128 x 64] SWIZZLE_128B boxes from the per-layer cache tensor
//                               maps, which travel as kernel parameters) through a 2-stage mbarrier ring
//   warp 1       MMA issuer   : S = Q K^T (tcgen05.mma M128 N128 K16 x 8, SS) -> TMEM S[it & 1];
//                               O += P V (TS: P read from TMEM, V consumed in place as an MN-major operand) -> TMEM O;
//                               ONE tcgen05.commit per MMA batch (a commit costs the issuing thread ~300 clk)
//   warps 2..17  softmax      : 4 threads per query row (32 key columns each); rows are spread over the four TMEM lane
//                               quadrants; lazy-rescaled online softmax; P (bf16) packed back into the S columns
// This step's token (whose k / v are still in the qkv buffer) is one extra split-KV partial (m = q.k, l = 1, O = v)
// computed on the CUDA cores by the CTA that also rotates k and appends k, v to the cache, so no tile is patched.
// The (m, l, O) partials are merged ACROSS the cluster through distributed shared memory: every CTA PUSHES its rows to
// the CTA that owns them (remote stores pipeline; remote loads measured ~190 clk each, serialised), one cluster barrier,
// then a local merge.  No global workspace, no combine kernel, no atomics; deterministic.  q is rotated (decode-position
// rope) while the Q tile is built.
// HBM-bound: algorithmic bytes per launch = 2 * Hkv * 128 * 2 * sum_groups(P + G * step) (+ q, out).
// The cache must hold finite values everywhere (it is allocated zeroed): masked keys are multiplied by P = 0.
#include "t1_common.cuh"
#ifdef T1_SOFTMAX_POLY_EXP2
#include "fast_exp2.cuh"
#endif
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace {

struct DecodeGroup {          // == t1_decode_group (include/timer1.h)
  bf16* K;
  bf16* V;
  float* ws_o;
  float* ws_ml;
  const float* cos_t;
  const float* sin_t;
  int prefix_len;
  int reserved;
};
static_assert(sizeof(DecodeGroup) == 56, "t1_decode_group layout");

constexpr int HD = 128;
constexpr int UNIT = 128;                        // keys per unit
constexpr int ATOM = 128 * 128;                  // [128 rows x 64 cols] bf16 SW128 atom
constexpr int TILE = 2 * ATOM;                   // [128 x 128] bf16
constexpr int STAGE_BYTES = 2 * TILE;            // K then V
constexpr int NSTG = 2;                          // == the two S buffers: s_full doubles as "K tile free"
constexpr int NS = 4;                            // softmax threads per query row
constexpr int NSOFT = 128 * NS;                  // 16 warps
static_assert(NSOFT >= 64 * 8, "the Q-tile build takes one (row, 8-channel chunk) per softmax thread");
constexpr int TC_THREADS = 64 + NSOFT;
constexpr int MAX_S = 16;
constexpr int MAX_OWN = 8;                       // sequences whose current token one CTA may own (ceil(G / S))
constexpr int MAX_SPLIT = 48;                    // wide form: CTAs per (group, KV head); sources per row = MAX_SPLIT + 1 <= 64
constexpr int CNT_STRIDE = 8;                    // wide form: one 32-byte sector of counters per (group, KV head)
constexpr int MAX_GROUPS = 8;                    // prompt-groups per launch (tensor maps travel as kernel parameters)
struct MapTable {
  CUtensorMap m[2 * MAX_GROUPS];                 // [group][K, V]
};
constexpr int MISC_BYTES = 512 /*barriers*/ + 2 * NS * 128 * 4 /*max xch*/ + NS * 128 * 4 /*l xch*/ +
                           2 * MAX_OWN * HD * 4 /*k_new, v_new*/;
constexpr int TC_SMEM = TILE /*Q*/ + NSTG * STAGE_BYTES + 1024 /*align*/ + MISC_BYTES;
// receive buffers of the cluster merge alias the K/V ring: [64/S rows][2S sources][128] fp32 + [..][2] (m, l)
static_assert(64 * 2 * HD * 4 + 64 * 2 * 2 * 4 <= NSTG * STAGE_BYTES, "merge buffers must fit in the K/V ring");

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
// byte offset of element (row, col) inside a K-major SWIZZLE_128B tile made of [128 x 64] atoms
__device__ __forceinline__ int sw128_off(int row, int col) {
  return (col >> 6) * ATOM + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
}

#ifdef T1_DECODE_DBG
__device__ long long g_dbg[1024];
#define DBG(cond, idx) do { if ((cond) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_dbg[(idx)] = clock64(); } while (0)
#else
#define DBG(cond, idx) do { } while (0)
#endif

struct UnitDesc {
  int slot0;     // first cache row
  int nvalid;    // real keys in the unit
  int seq;       // -1: prompt prefix (visible to every row); else the owning sequence
};

__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// WIDE = false: one cluster of S CTAs per (group, KV head), partials merged through distributed shared memory.
// WIDE = true : S = gridDim.x plain CTAs per (group, KV head) -- S is chosen so that the launch fills the machine (18 at
//   cfg2: 144 of 148 SMs instead of 64) -- and the partials travel through an L2-resident workspace: every CTA stores
//   its (m, l, O) rows, arrives on a per-(group, head) counter (release), spins until all S have arrived (acquire) and
//   merges the rows it owns.  All CTAs of the launch are co-resident (host: grid <= SM count, one CTA per SM), so the
//   spin cannot starve a peer; the last CTA to leave resets the counters for the next launch.
template <bool WIDE>
__global__ void __launch_bounds__(TC_THREADS, 1)
decode_attn_tc_kernel(const bf16* __restrict__ q, long long ldq, int Hq, int Hkv, int G, int cpad,
                      const int* __restrict__ step_ptr, bf16* __restrict__ out, long long ldo, float scale_log2,
                      const DecodeGroup* __restrict__ groups, const __grid_constant__ MapTable maps,
                      float* __restrict__ ws_o, float* __restrict__ ws_ml, unsigned* __restrict__ ws_cnt) {
  cg::cluster_group cluster = cg::this_cluster();
  const int S = WIDE ? (int)gridDim.x : (int)cluster.num_blocks();
  const int rank = WIDE ? (int)blockIdx.x : (int)cluster.block_rank();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NSTG * STAGE_BYTES);
  uint64_t* full = bars;                       // NSTG   TMA -> MMA: K tile of the stage landed
  uint64_t* done = bars + NSTG;                // NSTG   PV(it) retired: V tile free (producer), P / O safe (softmax)
  uint64_t* s_full = bars + 2 * NSTG;          // 2
  // p_ready is double-buffered by unit parity: a softmax warp may run one unit ahead of the slowest one (S(it+1) is
  // issued before p_ready(it) completes) and its early arrival must not count towards the previous unit's phase
  uint64_t* p_ready = bars + 2 * NSTG + 2;     // 2 (one arrival per softmax warp)
  uint64_t* q_ready = bars + 2 * NSTG + 4;     // 1 (one arrival per softmax warp)
  // K and V tiles of a stage are tracked separately: the K tile is free as soon as S(it) has retired (s_full, which the
  // producer also waits on), long before PV(it) frees the V tile, so K(it+2) is in flight while softmax(it) still runs
  uint64_t* v_full = bars + 2 * NSTG + 5;      // NSTG   TMA -> MMA: V tile of the stage landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 3 * NSTG + 5);
  float* s_xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [2 it][NS parts][128 lanes]
  float* s_lx = s_xch + 2 * NS * 128;                                               // [NS parts][128 lanes]
  float* s_knew = s_lx + NS * 128;                                                  // [MAX_OWN][128]
  float* s_vnew = s_knew + MAX_OWN * HD;                                            // [MAX_OWN][128]

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gi = blockIdx.z, hk = blockIdx.y;
  // descriptors live in kernel-parameter space (a tensor map in global memory is fetched per TMA instruction)
  const CUtensorMap* tm_k = &maps.m[gi * 2];
  const CUtensorMap* tm_v = &maps.m[gi * 2 + 1];

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(tm_k);
    tma_prefetch_desc(tm_v);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < NSTG; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&done[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(&p_ready[0], NSOFT / 32);
    mbar_init(&p_ready[1], NSOFT / 32);
    mbar_init(q_ready, NSOFT / 32);
    mbar_fence_init();
  }
  if (warp_idx == 1) tmem_alloc<512>(tmem_ptr_smem);
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Only q was written by the previous kernel, and only the softmax warps read it -- they execute griddepcontrol.wait.
  // The TMA producer does not: the cached K/V rows and *step_ptr were written several kernels up the chain (every
  // kernel of the chain waits, so completion is transitive), and its first NSTG units travel under the tail of the
  // qkv GEMM and the Q-tile build.  The MMA warp only touches shared memory / TMEM behind mbarriers.
  // The softmax warps wait inside their branch, after the part of the Q-tile build that does not need q.
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t T_S = 0, T_O = 256;           // S[2] x 128 columns, O 128 columns

  DBG(threadIdx.x == 64, 900);
  const DecodeGroup grp = groups[gi];
  const int ratio = Hq / Hkv;
  const int n_rows = G * ratio;
  const int rpq = (n_rows + 3) >> 2;           // query rows per TMEM lane quadrant
  const int P = grp.prefix_len;
  const int step = *step_ptr;                  // cache slots [0, step) of every sequence hold earlier tokens
  q += (long long)gi * G * ldq;
  out += (long long)gi * G * ldo;

  // ---- this CTA's units: an even share of the prefix units plus an even share of the completion units ----
  const int UP = (P + UNIT - 1) / UNIT;
  const int UPS = (step + UNIT - 1) / UNIT;    // units per sequence holding slots 0..step-1
  const int UC = G * UPS;
  int p_lo, p_hi, c_lo, c_hi;
  if constexpr (WIDE) {      // one even deal of the concatenated list [prefix units | completion units]
    const int U = UP + UC;
    const int u_lo = (int)((long long)rank * U / S), u_hi = (int)((long long)(rank + 1) * U / S);
    p_lo = min(u_lo, UP), p_hi = min(u_hi, UP);
    c_lo = max(u_lo, UP) - UP, c_hi = max(u_hi, UP) - UP;
  } else {
    p_lo = (int)((long long)rank * UP / S), p_hi = (int)((long long)(rank + 1) * UP / S);
    c_lo = (int)((long long)rank * UC / S), c_hi = (int)((long long)(rank + 1) * UC / S);
  }
  const int n_units = (p_hi - p_lo) + (c_hi - c_lo);
  auto unit_desc = [&](int i) {
    UnitDesc d;
    if (i < p_hi - p_lo) {
      const int u = p_lo + i;
      d.slot0 = u * UNIT;
      d.nvalid = min(UNIT, P - u * UNIT);
      d.seq = -1;
    } else {
      const int cu = c_lo + (i - (p_hi - p_lo));
      d.seq = cu / UPS;
      const int c0 = (cu - d.seq * UPS) * UNIT;
      d.slot0 = P + d.seq * cpad + c0;
      d.nvalid = min(UNIT, step - c0);
    }
    return d;
  };

  // per-thread state of the softmax warps (used again after the cluster barrier)
  const int quad = warp_idx & 3;               // TMEM lane quadrant this warp may touch
  const int part = (warp_idx - 2) >> 2;        // which 32-column slice of S (and of O)
  const int st_tid = threadIdx.x - 64;         // 0..NSOFT-1 for softmax threads
  const int L = quad * 32 + lane;              // TMEM lane == Q tile row
  const int r = (lane < rpq) ? quad * rpq + lane : -1;      // logical query row
  const bool valid = warp_idx >= 2 && r >= 0 && r < n_rows;
  const int my_seq = valid ? r / ratio : -1;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  float m_ref = -INFINITY, l = 0.f;

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < n_units; ++it) {
        const int st = it % NSTG;
        const UnitDesc d = unit_desc(it);
        uint8_t* sK = sKV + st * STAGE_BYTES;
        uint8_t* sV = sK + TILE;
        // S(it - NSTG) read this K tile; it cannot complete a later phase before K(it) is loaded, so the parity is exact
        if (it >= NSTG) mbar_wait(&s_full[(it - NSTG) & 1], ((it - NSTG) >> 1) & 1);
        mbar_arrive_expect_tx(&full[st], TILE);
        tma_load_2d(sK, tm_k, &full[st], hk * HD, d.slot0);
        tma_load_2d(sK + ATOM, tm_k, &full[st], hk * HD + 64, d.slot0);
        mbar_wait(&done[st], ((it / NSTG) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[st], TILE);
        tma_load_2d(sV, tm_v, &v_full[st], hk * HD, d.slot0);
        tma_load_2d(sV + ATOM, tm_v, &v_full[st], hk * HD + 64, d.slot0);
        DBG(it < 100, it * 8 + 0);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_units > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, UNIT);                   // A = Q, B = K, both K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD) | (1u << 16);       // B = V MN-major
      auto issue_s = [&](int it) {
        const int st = it % NSTG;
        mbar_wait(&full[st], (it / NSTG) & 1);
        tc_fence_after();
        DBG(it < 100, it * 8 + 1);
        const uint32_t tS = tmem_base + T_S + (it & 1) * UNIT;
        const uint8_t* sK = sKV + st * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * ATOM + (ks & 3) * 32;
          umma_bf16_ss(tS, umma_desc_sw128_kmajor(smem_u32(sQ + off)), umma_desc_sw128_kmajor(smem_u32(sK + off)), idesc_s,
                       ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[it & 1]);
      };
      mbar_wait(q_ready, 0);
      tc_fence_after();
      issue_s(0);
      for (int it = 0; it < n_units; ++it) {
        if (it + 1 < n_units) issue_s(it + 1);
        const int st = it % NSTG;
        mbar_wait(&v_full[st], (it / NSTG) & 1);
        mbar_wait(&p_ready[it & 1], (it >> 1) & 1);
        tc_fence_after();
        DBG(it < 100, it * 8 + 2);
        const uint8_t* sV = sKV + st * STAGE_BYTES + TILE;
        const uint32_t tP = tmem_base + T_S + (it & 1) * UNIT;
#pragma unroll
        for (int kk = 0; kk < UNIT / 16; ++kk)
          umma_bf16_ts(tmem_base + T_O, tP + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8),
                       umma_desc_sw128_mnmajor(smem_u32(sV + kk * 16 * 128), ATOM, 1024), idesc_pv, (it | kk) != 0 ? 1u : 0u);
        umma_commit(&done[st]);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax warps: NS threads per query row, 32 key columns each =====================
    // ---- Q tile: zero, then the rotated query rows (row r at lane 32*(r/rpq) + r%rpq) in the UMMA K-major swizzle ----
    for (int e = st_tid; e < TILE / 16; e += NSOFT) reinterpret_cast<uint4*>(sQ)[e] = make_uint4(0u, 0u, 0u, 0u);
    const float* cs = grp.cos_t + (long long)step * (HD / 2);
    const float* sn = grp.sin_t + (long long)step * (HD / 2);
    // one (row, 8-channel chunk) per thread; chunk c pairs with chunk c + 8.  The rope factors do not depend on q.
    const int q_row = st_tid >> 3, q_c = st_tid & 7;
    const bool q_on = q_row < n_rows;
    float c8[8], s8[8];
    if (q_on) {
      *reinterpret_cast<float4*>(c8) = *reinterpret_cast<const float4*>(cs + q_c * 8);
      *reinterpret_cast<float4*>(c8 + 4) = *reinterpret_cast<const float4*>(cs + q_c * 8 + 4);
      *reinterpret_cast<float4*>(s8) = *reinterpret_cast<const float4*>(sn + q_c * 8);
      *reinterpret_cast<float4*>(s8 + 4) = *reinterpret_cast<const float4*>(sn + q_c * 8 + 4);
    }
    // rope factors of this step's k (they do not depend on q either)
    float kc = 0.f, ks = 0.f;
    if (st_tid < HD / 2) {
      kc = cs[st_tid];
      ks = sn[st_tid];
    }
    pdl_wait();                                  // q (the fused qkv rows of this step) is complete and visible
    // every global load that had to wait for the dependency is issued here, together: the query chunk AND this thread's element
    // of the new k / v rows of the owned sequences (one L2 round trip on the chain's critical path instead of two)
    uint4 ua = make_uint4(0u, 0u, 0u, 0u), ub = ua;
    if (q_on) {
      const int sq = q_row / ratio, hq = hk * ratio + q_row % ratio;
      const bf16* src = q + (long long)sq * ldq + (long long)hq * HD + q_c * 8;
      ua = *reinterpret_cast<const uint4*>(src);
      ub = *reinterpret_cast<const uint4*>(src + HD / 2);
    }
    // (the FIRST owned sequence -- the only one when G <= cluster size; further ones load inside the loop below)
    bf16 kv0 = __float2bfloat16(0.f), kv1 = kv0;     // k pair (threads < HD/2) or v element in kv0 (threads HD/2 .. HD/2 + HD)
    if (rank < G) {
      const bf16* src = q + (long long)rank * ldq;
      if (st_tid < HD / 2) {
        const bf16* pk = src + (long long)(Hq + hk) * HD + st_tid;
        kv0 = pk[0];
        kv1 = pk[HD / 2];
      } else if (st_tid < HD / 2 + HD) {
        kv0 = src[(long long)(Hq + Hkv + hk) * HD + (st_tid - HD / 2)];
      }
    }
    asm volatile("bar.sync 9, %0;" ::"n"(NSOFT) : "memory");
    if (q_on) {
      const int row = q_row, c = q_c;
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&ua);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(&ub);
      uint32_t oa[4], ob[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16x2(pa[j]), b = unpack_bf16x2(pb[j]);
        float ra0, rb0, ra1, rb1;
        rope_rotate(a.x, b.x, c8[2 * j], s8[2 * j], ra0, rb0);
        rope_rotate(a.y, b.y, c8[2 * j + 1], s8[2 * j + 1], ra1, rb1);
        oa[j] = pack_bf16x2(ra0, ra1);
        ob[j] = pack_bf16x2(rb0, rb1);
      }
      const int Lr = (row / rpq) * 32 + row % rpq;
      *reinterpret_cast<uint4*>(sQ + sw128_off(Lr, c * 8)) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
      *reinterpret_cast<uint4*>(sQ + sw128_off(Lr, c * 8 + 64)) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
    }
    // ---- this step's k / v of the sequences this CTA owns: rotate k, append both to the cache, keep fp32 copies ----
    for (int b = rank, o = 0; b < G && o < MAX_OWN; b += S, ++o) {
      const bf16* src = q + (long long)b * ldq;
      const long long slot = (long long)P + (long long)b * cpad + step;
      if (o > 0) {
        if (st_tid < HD / 2) {
          const bf16* pk = src + (long long)(Hq + hk) * HD + st_tid;
          kv0 = pk[0];
          kv1 = pk[HD / 2];
        } else if (st_tid < HD / 2 + HD) {
          kv0 = src[(long long)(Hq + Hkv + hk) * HD + (st_tid - HD / 2)];
        }
      }
      if (st_tid < HD / 2) {
        float ra, rb;
        rope_rotate(__bfloat162float(kv0), __bfloat162float(kv1), kc, ks, ra, rb);
        const bf16 ka = __float2bfloat16(ra), kb = __float2bfloat16(rb);
        bf16* kd = grp.K + (slot * Hkv + hk) * HD + st_tid;
        kd[0] = ka;
        kd[HD / 2] = kb;
        s_knew[o * HD + st_tid] = __bfloat162float(ka);
        s_knew[o * HD + st_tid + HD / 2] = __bfloat162float(kb);
      } else if (st_tid < HD / 2 + HD) {
        const int j = st_tid - HD / 2;
        const bf16 vv = kv0;
        grp.V[(slot * Hkv + hk) * HD + j] = vv;
        s_vnew[o * HD + j] = __bfloat162float(vv);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(q_ready);
    asm volatile("bar.sync 9, %0;" ::"n"(NSOFT) : "memory");     // Q tile, k_new, v_new visible to all softmax threads
    DBG(st_tid == 0, 901);

    // ---- split-KV online softmax over the units ----
    for (int it = 0; it < n_units; ++it) {
      const UnitDesc d = unit_desc(it);
      mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      DBG(st_tid == 0 && it < 100, it * 8 + 3);
      uint32_t sv[32];
      tmem_ld_32x32b_x32(t_lane + T_S + (it & 1) * UNIT + part * 32, sv);
      tmem_ld_wait();
      const bool need_mask = d.seq >= 0 || d.nvalid < UNIT;
      const bool row_sees = valid && (d.seq < 0 || d.seq == my_seq);
      float tmax = -INFINITY;
      if (need_mask) {                           // warp-uniform; keeps the unmasked prefix units branch-free
        const int lim = row_sees ? d.nvalid - part * 32 : 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = j < lim ? __uint_as_float(sv[j]) : -INFINITY;
          sv[j] = __float_as_uint(v);
          tmax = fmaxf(tmax, v);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(sv[j]));
      }
      s_xch[((it & 1) * NS + part) * 128 + L] = tmax;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
#pragma unroll
      for (int o = 0; o < NS; ++o) tmax = fmaxf(tmax, s_xch[((it & 1) * NS + o) * 128 + L]);
      tmax *= scale_log2;
      DBG(st_tid == 0 && it < 100, it * 8 + 4);
      const bool grow = tmax > m_ref + 8.0f;     // lazy rescale (the NS threads of a row take the same decision)
      float alpha = 1.0f;
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - tmax);
        m_ref = tmax;
        l *= alpha;
      }
      // P(it) overwrites the S columns P(it-2) was read from: PV(it-2) must have retired (it has, long ago)
      if (it >= 2) mbar_wait(&done[it % NSTG], ((it / NSTG) - 1) & 1);
      if (it > 0 && __any_sync(0xffffffffu, grow)) {
        // rare after the first units: the O accumulator is rescaled in place once PV(it-1) has retired
        mbar_wait(&done[(it - 1) % NSTG], ((it - 1) / NSTG) & 1);
        tc_fence_after();
        uint32_t ov[32];
        tmem_ld_32x32b_x32(t_lane + T_O + part * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * alpha);
        tmem_st_32x32b_x32(t_lane + T_O + part * 32, ov);
        tmem_st_wait();
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float rs = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * j]), scale_log2, neg_m));
        const float p1 = EX2_B(fmaf(__uint_as_float(sv[2 * j + 1]), scale_log2, neg_m));
        rs += p0 + p1;
        pk[j] = pack_bf16x2(p0, p1);
      }
      l += rs;
      tmem_st_32x32b_x16(t_lane + T_S + (it & 1) * UNIT + part * 32, pk);
      tmem_st_wait();                              // warp-collective: every lane's P columns are in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[it & 1]);
      DBG(st_tid == 0 && it < 100, it * 8 + 5);
    }
    DBG(st_tid == 0, 902);
    // ---- row sum over the NS parts; wait for the last PV ----
    s_lx[part * 128 + L] = l;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NS) : "memory");
    l = 0.f;
#pragma unroll
    for (int o = 0; o < NS; ++o) l += s_lx[o * 128 + L];
    if (n_units > 0) {
      mbar_wait(&done[(n_units - 1) % NSTG], ((n_units - 1) / NSTG) & 1);
      tc_fence_after();
    }
  }
  DBG(threadIdx.x == 64, 903);
  // Split-phase cluster barrier: arrive = "this CTA is done with its K/V ring" (which becomes a merge receive buffer of
  // the cluster); the work that needs no remote memory (O out of TMEM, the score of this step's token) runs before the wait.
  if constexpr (!WIDE) cluster.barrier_arrive();
  float* rO = reinterpret_cast<float*>(sKV);            // [64/S local rows][2S sources][128]
  float* rML = rO + 64 * 2 * HD;                        // [64/S local rows][2S sources][2]
  uint32_t ov[32];
  float s_new = -INFINITY, l_new = 0.f;
  const bool own = valid && (my_seq % S) == rank && (my_seq / S) < MAX_OWN;
  const int own_o = own ? my_seq / S : 0;
  if (warp_idx >= 2) {
    if (n_units > 0) {                            // warp-uniform: tcgen05.ld is .sync.aligned
      tmem_ld_32x32b_x32(t_lane + T_O + part * 32, ov);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) ov[j] = 0u;
    }
    if (own && part == 0) {
      float acc = 0.f;
#pragma unroll 4
      for (int c = 0; c < 16; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(sQ + sw128_off(L, c * 8));
        const uint32_t* pu = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 qq = unpack_bf16x2(pu[j]);
          acc = fmaf(qq.x, s_knew[own_o * HD + c * 8 + 2 * j], acc);
          acc = fmaf(qq.y, s_knew[own_o * HD + c * 8 + 2 * j + 1], acc);
        }
      }
      s_new = acc * scale_log2;
      l_new = 1.f;
    }
  }
  // number of merge sources per query row and where the merge reads them from
  const int NSRC = WIDE ? S + 1 : 2 * S;
  const int n_local = rank < n_rows ? (n_rows - rank + S - 1) / S : 0;      // rows rank, rank + S, ... are merged here
  if constexpr (WIDE) {
    // ---- store: source `rank` = cached-key partial, source S = this step's token (owner of the sequence only) ----
    const int gh = gi * Hkv + hk;
    float* wO = ws_o + (size_t)gh * 64 * NSRC * HD;     // [row][source][128]
    float* wML = ws_ml + (size_t)gh * 64 * NSRC * 2;    // [row][source][m, l]
    if (valid) {
      float* dO = wO + ((size_t)r * NSRC + rank) * HD + part * 32;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        __stcg(reinterpret_cast<uint4*>(dO) + j4, make_uint4(ov[j4 * 4], ov[j4 * 4 + 1], ov[j4 * 4 + 2], ov[j4 * 4 + 3]));
      if (part == 0) __stcg(reinterpret_cast<float2*>(wML + ((size_t)r * NSRC + rank) * 2), make_float2(m_ref, l));
      if (own) {
        float* dN = wO + ((size_t)r * NSRC + S) * HD + part * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          __stcg(reinterpret_cast<float4*>(dN) + j4, *reinterpret_cast<const float4*>(s_vnew + own_o * HD + part * 32 + j4 * 4));
        if (part == 0) __stcg(reinterpret_cast<float2*>(wML + ((size_t)r * NSRC + S) * 2), make_float2(s_new, l_new));
      }
      __threadfence();
    }
    DBG(threadIdx.x == 64, 904);
    __syncthreads();           // every partial row of this CTA is stored and fenced
    if (threadIdx.x == 64) {   // a softmax thread: it executed griddepcontrol.wait, so the counters are the reset ones
      unsigned* cnt = ws_cnt + CNT_STRIDE * gh;
      red_release_add(cnt, 1u);
      // ~1 us per poll: a peer that never arrives (a co-residency violation) traps after seconds instead of hanging the device
      unsigned polls = 0;
      while (ld_acquire(cnt) < (unsigned)S) {
        if (++polls > (1u << 22)) __trap();
      }
      if (atomicAdd(cnt + 1, 1u) == (unsigned)(S - 1)) {      // last one out: nobody spins on cnt any more
        atomicExch(cnt, 0u);
        atomicExch(cnt + 1, 0u);
      }
    }
    DBG(threadIdx.x == 64, 905);
    __syncthreads();           // all S CTAs of this (group, head) have stored their partials
    DBG(threadIdx.x == 64, 906);
    if (warp_idx >= 2) {
      float* s_w = s_xch;                           // [local rows][NSRC]
      float* s_L = s_lx;                            // [local rows]
      for (int li = warp_idx - 2; li < n_local; li += NSOFT / 32) {
        const float* ml = wML + (size_t)(rank + li * S) * NSRC * 2;
        float mv[2], lv[2];                         // NSRC <= 64: two sources per lane
        float M = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int src = lane + 32 * h;
          mv[h] = -INFINITY;
          lv[h] = 0.f;
          if (src < NSRC) {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(ml + src * 2));
            mv[h] = v.x;
            lv[h] = v.y;
          }
          if (lv[h] > 0.f) M = fmaxf(M, mv[h]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        float Ls = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int src = lane + 32 * h;
          const float w = lv[h] > 0.f ? ex2_approx(mv[h] - M) : 0.f;
          Ls += w * lv[h];
          if (src < NSRC) s_w[li * NSRC + src] = w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Ls += __shfl_xor_sync(0xffffffffu, Ls, o);
        if (lane == 0) s_L[li] = Ls;
      }
      asm volatile("bar.sync 9, %0;" ::"n"(NSOFT) : "memory");
      for (int e = st_tid; e < n_local * HD; e += NSOFT) {
        const int li = e >> 7, col = e & (HD - 1);
        const int row = rank + li * S;
        const float* po = wO + (size_t)row * NSRC * HD + col;
        const float* w = s_w + li * NSRC;
        float acc = 0.f;
        // every source's partial of this element is requested before the first one is used (L2: written by other SMs during this
        // launch): one round trip per element instead of one per 8 sources; same summation order
        for (int s0 = 0; s0 < NSRC; s0 += 32) {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = (s0 + j < NSRC) ? __ldcg(po + (s0 + j) * HD) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (s0 + j < NSRC) {
              const float ws = w[s0 + j];
              acc += ws > 0.f ? ws * o[j] : 0.f;    // never-written rows of empty sources are dropped
            }
          }
        }
        const float Ls = s_L[li];
        const int sq = row / ratio, hq = hk * ratio + row % ratio;
        out[(long long)sq * ldo + (long long)hq * HD + col] = __float2bfloat16(Ls > 0.f ? acc / Ls : 0.f);
      }
    }
  } else {
  cluster.barrier_wait();    // every CTA of the cluster is done with its K/V ring
  DBG(threadIdx.x == 64, 904);

  if (valid) {
    // ---- push: source 2*rank = cached-key partial, source 2*rank+1 = this step's token (owner of the sequence only) ----
    const int owner = r % S, li = r / S;
    float* dO = cluster.map_shared_rank(rO, owner) + (long long)(li * 2 * S + 2 * rank) * HD + part * 32;
    float* dML = cluster.map_shared_rank(rML, owner) + (li * 2 * S + 2 * rank) * 2;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
      *reinterpret_cast<uint4*>(dO + j4 * 4) = make_uint4(ov[j4 * 4], ov[j4 * 4 + 1], ov[j4 * 4 + 2], ov[j4 * 4 + 3]);
    if (own) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(dO + HD + j4 * 4) = *reinterpret_cast<const float4*>(s_vnew + own_o * HD + part * 32 + j4 * 4);
    }
    if (part == 0) *reinterpret_cast<float4*>(dML) = make_float4(m_ref, l, s_new, l_new);
  }
  DBG(threadIdx.x == 64, 905);
  cluster.sync();            // all pushes have landed (release / acquire at cluster scope)
  DBG(threadIdx.x == 64, 906);

  // ---- local merge of the rows this CTA owns (rank, rank + S, ...) ----
  if (warp_idx >= 2) {
    // pass 1, one warp per row: lane = source; w = 2^(m - M) (0 for empty sources), L = sum w * l
    float* s_w = s_xch;                           // [local rows][2S]  (the exchange buffers are free now)
    float* s_L = s_lx;                            // [local rows]
    for (int li = warp_idx - 2; li < n_local; li += NSOFT / 32) {
      const float* ml = rML + (li * 2 * S) * 2;
      float mv = -INFINITY, lv = 0.f;
      if (lane < 2 * S) {
        mv = ml[lane * 2];
        lv = ml[lane * 2 + 1];
      }
      const bool on = lv > 0.f;
      float M = on ? mv : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
      const float w = on ? ex2_approx(mv - M) : 0.f;
      float Ls = w * lv;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Ls += __shfl_xor_sync(0xffffffffu, Ls, o);
      if (lane < 2 * S) s_w[li * 2 * S + lane] = w;
      if (lane == 0) s_L[li] = Ls;
    }
    asm volatile("bar.sync 9, %0;" ::"n"(NSOFT) : "memory");
    // pass 2, one thread per (row, head-dim column)
    for (int e = st_tid; e < n_local * HD; e += NSOFT) {
      const int li = e >> 7, col = e & (HD - 1);
      const float* po = rO + (long long)(li * 2 * S) * HD + col;
      const float* w = s_w + li * 2 * S;
      float acc = 0.f;
#pragma unroll 4
      for (int src = 0; src < 2 * S; ++src) {
        const float ws = w[src];
        const float o = po[src * HD];             // local shared memory; garbage of empty sources is dropped
        acc += ws > 0.f ? ws * o : 0.f;
      }
      const float Ls = s_L[li];
      const int row = rank + li * S;
      const int sq = row / ratio, hq = hk * ratio + row % ratio;
      out[(long long)sq * ldo + (long long)hq * HD + col] = __float2bfloat16(Ls > 0.f ? acc / Ls : 0.f);
    }
  }
  }
  DBG(threadIdx.x == 64, 907);
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

#ifdef T1_DECODE_DBG
extern "C" int t1_decode_attn_dbg_read(long long* out) {
  return cudaMemcpyFromSymbol(out, g_dbg, sizeof(long long) * 1024) == cudaSuccess ? 0 : -2;
}
#endif

// Encode a 2-D SWIZZLE_128B bf16 tensor map (row-major [rows, cols], leading dimension ld elements, box
// box_rows x box_cols) into the 128 bytes at `out` (host memory).  Used to build the device-resident per-layer K/V
// cache map table of t1_decode_attn_cluster.
extern "C" int t1_encode_tmap_bf16(void* out, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                                   int box_cols) {
  T1_CHECK_ARG(out && ptr, "t1_encode_tmap_bf16: null pointer");
  CUtensorMap m;
  const int rc = t1_make_tmap_bf16(&m, ptr, rows, cols, ld, box_rows, box_cols);
  if (rc) return rc;
  memcpy(out, &m, sizeof(CUtensorMap));
  return T1_OK;
}

// Decode attention of one layer for all prompt-groups in ONE launch (see the header of this file).
//   q [n_groups*G, ldq] = the UN-rotated fused qkv rows of this step; groups: device array of t1_decode_group;
//   maps: HOST array [n_groups][2] of 128-byte tensor maps over each group's K and V cache ([rows, Hkv*128], box
//   64 x 64, from t1_encode_tmap_bf16), copied into the kernel parameters (n_groups <= 8); *step_ptr = number of tokens already cached per sequence (this step's slot index).
//   cluster_size in {1, 2, 4, 8} (0 = 8) CTAs per (group, KV head).
extern "C" int t1_decode_attn_cluster(const void* q, long long ldq, const void* groups, const void* maps, int n_groups,
                                      int G, int Hq, int Hkv, int head_dim, int cpad, const int* step_ptr, void* out,
                                      long long ldo, float scale, int cluster_size, cudaStream_t stream) {
  T1_CHECK_ARG(head_dim == HD, "t1_decode_attn_cluster: head_dim must be 128 (got %d)", head_dim);
  T1_CHECK_ARG(cpad > 0, "t1_decode_attn_cluster: cpad must be positive");
  T1_CHECK_ARG(n_groups > 0 && G > 0 && Hkv > 0 && Hq % Hkv == 0 && G * (Hq / Hkv) <= 64,
               "t1_decode_attn_cluster: G * Hq/Hkv must be <= 64");
  T1_CHECK_ARG(q && groups && maps && step_ptr && out, "t1_decode_attn_cluster: null pointer");
  T1_CHECK_ARG(ldq % 8 == 0 && ((uintptr_t)q & 15) == 0, "t1_decode_attn_cluster: q must be 16-byte aligned with ldq %% 8 == 0");
  T1_CHECK_ARG(n_groups <= MAX_GROUPS, "t1_decode_attn_cluster: at most %d prompt-groups per launch (got %d)", MAX_GROUPS, n_groups);
  MapTable tbl;
  memset(&tbl, 0, sizeof(tbl));
  memcpy(&tbl, maps, (size_t)n_groups * 2 * sizeof(CUtensorMap));
  int S = cluster_size > 0 ? cluster_size : 8;
  if (getenv("T1_DECODE_CLUSTER")) S = atoi(getenv("T1_DECODE_CLUSTER"));
  T1_CHECK_ARG(S == 1 || S == 2 || S == 4 || S == 8 || S == 16, "t1_decode_attn_cluster: cluster size %d not in {1,2,4,8,16}", S);
  T1_CHECK_ARG((G + S - 1) / S <= MAX_OWN, "t1_decode_attn_cluster: G=%d too large for cluster size %d", G, S);
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(decode_attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(decode_attn_tc_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
      t1_set_error("t1_decode_attn_cluster: cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, Hkv, n_groups);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t1_pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_attn_tc_kernel<false>, (const bf16*)q, ldq, Hq, Hkv, G, cpad, step_ptr, (bf16*)out,
                                     ldo, scale * 1.4426950408889634f, (const DecodeGroup*)groups, tbl, (float*)nullptr,
                                     (float*)nullptr, (unsigned*)nullptr);
  if (e != cudaSuccess) {
    t1_set_error("t1_decode_attn_cluster: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  return T1_OK;
}

// ---- wide form: the same CTA program on (almost) every SM, partials merged through an L2-resident workspace ----
static int wide_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}
static size_t wide_ws_bytes(int n_groups, int Hkv, int S) {
  const size_t gh = (size_t)n_groups * Hkv;
  return 256 + gh * CNT_STRIDE * 4 + gh * 64 * (S + 1) * (HD + 2) * sizeof(float);
}

// Plan of t1_decode_attn_wide for n_groups prompt-groups of Hkv KV heads on the current device: *splits = CTAs per
// (group, KV head) (the largest count for which the whole launch is co-resident, one CTA per SM; `requested` > 0 overrides
// it and is validated), *ws_bytes = size of the workspace the caller allocates AND ZEROES ONCE (the launch leaves its
// counters zeroed again).
extern "C" int t1_decode_attn_wide_plan(int n_groups, int Hkv, int requested, int* splits, long long* ws_bytes) {
  T1_CHECK_ARG(n_groups > 0 && n_groups <= MAX_GROUPS && Hkv > 0 && splits && ws_bytes, "t1_decode_attn_wide_plan: bad arguments");
  const int sms = wide_sm_count();
  T1_CHECK_ARG(sms > 0, "t1_decode_attn_wide_plan: no CUDA device");
  const int fit = sms / (n_groups * Hkv);
  T1_CHECK_ARG(fit >= 1, "t1_decode_attn_wide_plan: %d (group, KV head) pairs do not fit on %d SMs", n_groups * Hkv, sms);
  int S = requested > 0 ? requested : (fit < MAX_SPLIT ? fit : MAX_SPLIT);
  if (requested <= 0 && getenv("T1_DECODE_SPLITS")) S = atoi(getenv("T1_DECODE_SPLITS"));
  T1_CHECK_ARG(S >= 1 && S <= MAX_SPLIT && S <= fit,
               "t1_decode_attn_wide_plan: %d CTAs per (group, KV head) not in [1, min(%d, %d)] (all CTAs must be co-resident)", S,
               MAX_SPLIT, fit);
  *splits = S;
  *ws_bytes = (long long)wide_ws_bytes(n_groups, Hkv, S);
  return T1_OK;
}

// Decode attention of one layer for all prompt-groups in ONE launch of `splits` CTAs per (group, KV head) without
// clusters (arguments as t1_decode_attn_cluster).  ws / ws_bytes / splits: from t1_decode_attn_wide_plan; the first
// 256 + n_groups * Hkv * 32 bytes of ws are arrival counters that must be zero before the first launch.  The launch
// spins on them: it must not share the device with a kernel that waits for it.
extern "C" int t1_decode_attn_wide(const void* q, long long ldq, const void* groups, const void* maps, int n_groups, int G,
                                   int Hq, int Hkv, int head_dim, int cpad, const int* step_ptr, void* out, long long ldo,
                                   float scale, int splits, void* ws, long long ws_bytes, cudaStream_t stream) {
  T1_CHECK_ARG(head_dim == HD, "t1_decode_attn_wide: head_dim must be 128 (got %d)", head_dim);
  T1_CHECK_ARG(cpad > 0, "t1_decode_attn_wide: cpad must be positive");
  T1_CHECK_ARG(n_groups > 0 && G > 0 && Hkv > 0 && Hq % Hkv == 0 && G * (Hq / Hkv) <= 64,
               "t1_decode_attn_wide: G * Hq/Hkv must be <= 64");
  T1_CHECK_ARG(q && groups && maps && step_ptr && out && ws, "t1_decode_attn_wide: null pointer");
  T1_CHECK_ARG(ldq % 8 == 0 && ((uintptr_t)q & 15) == 0, "t1_decode_attn_wide: q must be 16-byte aligned with ldq %% 8 == 0");
  T1_CHECK_ARG(((uintptr_t)ws & 255) == 0, "t1_decode_attn_wide: ws must be 256-byte aligned");
  T1_CHECK_ARG(n_groups <= MAX_GROUPS, "t1_decode_attn_wide: at most %d prompt-groups per launch (got %d)", MAX_GROUPS, n_groups);
  const int S = splits;
  const int sms = wide_sm_count();
  T1_CHECK_ARG(S >= 1 && S <= MAX_SPLIT && S * Hkv * n_groups <= sms,
               "t1_decode_attn_wide: splits=%d: need 1 <= splits <= %d and splits * Hkv * n_groups <= %d SMs", S, MAX_SPLIT, sms);
  T1_CHECK_ARG((G + S - 1) / S <= MAX_OWN, "t1_decode_attn_wide: G=%d too large for %d splits", G, S);
  T1_CHECK_ARG(ws_bytes >= (long long)wide_ws_bytes(n_groups, Hkv, S), "t1_decode_attn_wide: workspace too small (%lld < %lld bytes)",
               ws_bytes, (long long)wide_ws_bytes(n_groups, Hkv, S));
  MapTable tbl;
  memset(&tbl, 0, sizeof(tbl));
  memcpy(&tbl, maps, (size_t)n_groups * 2 * sizeof(CUtensorMap));
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(decode_attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) {
      t1_set_error("t1_decode_attn_wide: cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e));
      return T1_ERR_CUDA;
    }
    set = true;
  }
  const size_t gh = (size_t)n_groups * Hkv;
  unsigned* ws_cnt = reinterpret_cast<unsigned*>(ws);
  float* ws_ml = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 256 + gh * CNT_STRIDE * 4);
  float* ws_o = ws_ml + gh * 64 * (S + 1) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, Hkv, n_groups);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t1_pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_attn_tc_kernel<true>, (const bf16*)q, ldq, Hq, Hkv, G, cpad, step_ptr, (bf16*)out,
                                     ldo, scale * 1.4426950408889634f, (const DecodeGroup*)groups, tbl, ws_o, ws_ml, ws_cnt);
  if (e != cudaSuccess) {
    t1_set_error("t1_decode_attn_wide: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  return T1_OK;
}
