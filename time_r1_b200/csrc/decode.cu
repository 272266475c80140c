// Decode-phase kernels of the GRPO rollout (replaces HF GenerationMixin._sample driven from
// src/time_r1/rl/timer1_trainer.py:568-573): all HBM-bandwidth-bound.
//   * skinny GEMM      y[M<=64, N] = x[M,K] W[N,K]^T   — streams W once with 16-byte loads, HMMA m16n8k16 with the
//                      batch rows on the N=8 side and a k-permutation so that each thread's 16-byte load IS its
//                      fragment (no shared-memory staging of weights)
//   * decode attention — split-KV over a unified key space [prompt prefix | per-sequence completion slots]; the G
//                      sequences of a group share ONE copy of the prompt K/V, so (G x q-heads-per-kv-head) query rows
//                      form a 64-row MMA tile against each prefix key chunk; partial (m, l, O) + a combine kernel
//   * rope + KV append, sampler (temperature -> top-k -> softmax -> multinomial, Philox4x32-10), step counter
#include "t1_common.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace {

// =============================================================================================
// skinny GEMM
// =============================================================================================
template <int NB>
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ W, long long ldw,
                   const bf16* __restrict__ bias, void* __restrict__ y, long long ldy, int M, int N, int K,
                   int out_f32, int accumulate, int ksplit) {
  // grid = (N/16, ksplit); the ksplit CTAs of one 16-row output slab form a thread-block cluster and reduce their
  // partial sums through distributed shared memory.
  namespace cg = cooperative_groups;
  __shared__ float red[8][16][8 * NB + 1];
  __shared__ float part[16 * 8 * NB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int n0 = blockIdx.x * 16;
  const int ks = blockIdx.y;
  // K range of this CTA in units of 32-element blocks
  const int kblocks = K >> 5;
  const int kb_lo = (int)((long long)kblocks * ks / ksplit) << 5, kb_hi = (int)((long long)kblocks * (ks + 1) / ksplit) << 5;
  float acc[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const bf16* w0 = W + (long long)(n0 + g) * ldw + t4 * 8;
  const bf16* w1 = W + (long long)(n0 + g + 8) * ldw + t4 * 8;
  // each warp owns a CONTIGUOUS k range so that its unrolled loads walk adjacent 64-byte segments of the same rows
  const int wblocks = (kb_hi - kb_lo) >> 5;
  const int wk_lo = kb_lo + ((wblocks * warp / 8) << 5), wk_hi = kb_lo + ((wblocks * (warp + 1) / 8) << 5);
#pragma unroll 8
  for (int kb = wk_lo; kb < wk_hi; kb += 32) {
    const uint4 wa = __ldg(reinterpret_cast<const uint4*>(w0 + kb));
    const uint4 wb = __ldg(reinterpret_cast<const uint4*>(w1 + kb));
    const uint32_t a_lo[4] = {wa.x, wb.x, wa.y, wb.y};
    const uint32_t a_hi[4] = {wa.z, wb.z, wa.w, wb.w};
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int m = nb * 8 + g;
      uint4 xb = make_uint4(0, 0, 0, 0);
      if (m < M) xb = *reinterpret_cast<const uint4*>(x + (long long)m * ldx + kb + t4 * 8);
      const uint32_t b_lo[2] = {xb.x, xb.y};
      const uint32_t b_hi[2] = {xb.z, xb.w};
      mma_bf16_16816(acc[nb], a_lo, b_lo);
      mma_bf16_16816(acc[nb], a_hi, b_hi);
    }
  }
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    red[warp][g][nb * 8 + t4 * 2] = acc[nb][0];
    red[warp][g][nb * 8 + t4 * 2 + 1] = acc[nb][1];
    red[warp][g + 8][nb * 8 + t4 * 2] = acc[nb][2];
    red[warp][g + 8][nb * 8 + t4 * 2 + 1] = acc[nb][3];
  }
  __syncthreads();
  constexpr int NOUT = 16 * 8 * NB;
  for (int e = tid; e < NOUT; e += 256) {
    const int f = e & 15, m = e >> 4;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][f][m];
    part[e] = sum;
  }
  if (ksplit > 1) cg::this_cluster().sync();
  else __syncthreads();
  // distributed final reduction: cluster rank r finishes outputs e == r (mod ksplit)
  for (int e = tid * ksplit + ks; e < NOUT; e += 256 * ksplit) {
    const int f = e & 15, m = e >> 4;
    if (m < M) {
      float sum = 0.f;
      if (ksplit > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        for (int r = 0; r < ksplit; ++r) sum += cluster.map_shared_rank(part, r)[e];
      } else {
        sum = part[e];
      }
      if (bias) sum += __bfloat162float(bias[n0 + f]);
      if (out_f32) {
        float* yp = reinterpret_cast<float*>(y) + (long long)m * ldy + n0 + f;
        *yp = accumulate ? *yp + sum : sum;
      } else {
        bf16* yp = reinterpret_cast<bf16*>(y) + (long long)m * ldy + n0 + f;
        *yp = __float2bfloat16(accumulate ? __bfloat162float(*yp) + sum : sum);
      }
    }
  }
  if (ksplit > 1) cg::this_cluster().sync();   // keep `part` alive until every peer has read it
}

template <int NB>
int launch_skinny(const bf16* x, long long ldx, const bf16* W, long long ldw, const bf16* bias, void* y, long long ldy,
                  int M, int N, int K, int out_f32, int accumulate, cudaStream_t stream) {
  // split K across a cluster until there are enough CTAs in flight to cover HBM latency
  int ksplit = 1;
  const int slabs = N / 16;
  // measured on B200 (profiles/r01_skinny_ksplit_sweep.txt): 2-CTA clusters win when one slab per CTA cannot fill
  // the machine (N <= ~19k); beyond that the DSMEM reduction only costs.
  if (slabs < 1184 && (K / 32) >= 32) ksplit = 2;
  static int forced = -1;   // tuning hook: T1_SKINNY_KSPLIT=1|2|4|8
  if (forced < 0) {
    const char* e = getenv("T1_SKINNY_KSPLIT");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0 && slabs * ksplit < 4096) ksplit = forced;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(slabs, ksplit, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = ksplit;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<NB>, x, ldx, W, ldw, bias, y, ldy, M, N, K, out_f32, accumulate, ksplit);
  if (e != cudaSuccess) {
    t1_set_error("t1_skinny_gemm: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  return T1_OK;
}

// =============================================================================================
// decode rope + KV append. One CTA per sequence. q,k rotated with table row *step_ptr; k,v appended at
// slot = prefix_len + b * cpad + step of the layer's cache.
// =============================================================================================
__global__ void __launch_bounds__(256)
decode_rope_append_kernel(bf16* __restrict__ qkv, long long ld, int Hq, int Hkv, int HD, const float* __restrict__ cos_t,
                          const float* __restrict__ sin_t, const int* __restrict__ step_ptr, bf16* __restrict__ Kc,
                          bf16* __restrict__ Vc, int prefix_len, int cpad) {
  const int b = blockIdx.x;
  const int step = *step_ptr;
  const int half = HD >> 1;
  const float* c = cos_t + (long long)step * half;
  const float* s = sin_t + (long long)step * half;
  bf16* row = qkv + (long long)b * ld;
  const long long slot = (long long)prefix_len + (long long)b * cpad + step;
  const int n_rot = (Hq + Hkv) * half;
  for (int e = threadIdx.x; e < n_rot; e += blockDim.x) {
    const int h = e / half, i = e - h * half;
    bf16* p = row + h * HD + i;
    float ra, rb;
    rope_rotate(__bfloat162float(p[0]), __bfloat162float(p[half]), c[i], s[i], ra, rb);
    const bf16 oa = __float2bfloat16(ra), ob = __float2bfloat16(rb);
    p[0] = oa;
    p[half] = ob;
    if (h >= Hq) {
      bf16* kd = Kc + slot * (Hkv * HD) + (h - Hq) * HD + i;
      kd[0] = oa;
      kd[half] = ob;
    }
  }
  const bf16* vsrc = row + (Hq + Hkv) * HD;
  bf16* vd = Vc + slot * (Hkv * HD);
  for (int e = threadIdx.x; e < Hkv * HD; e += blockDim.x) vd[e] = vsrc[e];
}

// Grouped variant for the PDL decode chain: one launch covers every sequence of every prompt-group; the per-group
// pointers live in a device-resident table (struct t1_decode_group, include/timer1.h).
struct DecodeGroup {
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

__global__ void __launch_bounds__(256)
decode_rope_append_groups_kernel(bf16* __restrict__ qkv, long long ld, int G, int Hq, int Hkv, int HD,
                                 const DecodeGroup* __restrict__ groups, const int* __restrict__ step_ptr, int cpad) {
  pdl_trigger();
  pdl_wait();
  const int bg = blockIdx.x;
  const int gi = bg / G, b = bg - gi * G;
  const DecodeGroup grp = groups[gi];
  const int step = *step_ptr;
  const int half = HD >> 1;
  const float* c = grp.cos_t + (long long)step * half;
  const float* s = grp.sin_t + (long long)step * half;
  bf16* row = qkv + (long long)bg * ld;
  const long long slot = (long long)grp.prefix_len + (long long)b * cpad + step;
  const int n_rot = (Hq + Hkv) * half;
  for (int e = threadIdx.x; e < n_rot; e += blockDim.x) {
    const int h = e / half, i = e - h * half;
    bf16* p = row + h * HD + i;
    float ra, rb;
    rope_rotate(__bfloat162float(p[0]), __bfloat162float(p[half]), c[i], s[i], ra, rb);
    const bf16 oa = __float2bfloat16(ra), ob = __float2bfloat16(rb);
    p[0] = oa;
    p[half] = ob;
    if (h >= Hq) {
      bf16* kd = grp.K + slot * (Hkv * HD) + (h - Hq) * HD + i;
      kd[0] = oa;
      kd[half] = ob;
    }
  }
  const bf16* vsrc = row + (Hq + Hkv) * HD;
  bf16* vd = grp.V + slot * (Hkv * HD);
  for (int e = threadIdx.x; e < Hkv * HD; e += blockDim.x) vd[e] = vsrc[e];
}

// =============================================================================================
// decode attention (HD = 128): grid (n_chunks, Hkv, row_tiles); 128 threads; 128 keys per CTA.
// Key slot j: j < prefix_len -> prompt key (visible to every row); otherwise seq = (j-P)/cpad, c = (j-P)%cpad,
// visible iff seq == row's sequence and c <= step.
// =============================================================================================
constexpr int DA_KEYS = 128;
constexpr int DA_LDS = 128 + 8;

// With `groups` != nullptr the launch covers several prompt-groups: blockIdx.z = group (one 64-row tile per group),
// and q / Kc / Vc / prefix_len / workspaces come from the group table.
__global__ void __launch_bounds__(128)
decode_attn_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ Kc, const bf16* __restrict__ Vc,
                   int Hq, int Hkv, int B, int prefix_len, int cpad, const int* __restrict__ step_ptr,
                   float* __restrict__ ws_o, float* __restrict__ ws_ml, int n_prefix_chunks, int rows_pad,
                   float scale_log2, const DecodeGroup* __restrict__ groups, int fuse_rope) {
  constexpr int HD = 128, LDS = DA_LDS;
  pdl_trigger();
  pdl_wait();
  int rt_override = -1;
  const float* rope_cos = nullptr;
  const float* rope_sin = nullptr;
  bf16* Kw = nullptr;
  bf16* Vw = nullptr;
  if (groups) {
    const DecodeGroup grp = groups[blockIdx.z];
    rope_cos = grp.cos_t;
    rope_sin = grp.sin_t;
    Kw = grp.K;
    Vw = grp.V;
    q += (long long)blockIdx.z * B * ldq;
    Kc = grp.K;
    Vc = grp.V;
    ws_o = grp.ws_o;
    ws_ml = grp.ws_ml;
    prefix_len = grp.prefix_len;
    n_prefix_chunks = (prefix_len + DA_KEYS - 1) / DA_KEYS;
    if ((int)blockIdx.x >= n_prefix_chunks + B * (cpad / DA_KEYS)) return;
    rt_override = 0;
  }
  extern __shared__ __align__(16) uint8_t da_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(da_smem);   // 64 x LDS
  bf16* sK = sQ + 64 * LDS;                      // 128 x LDS
  bf16* sV = sK + DA_KEYS * LDS;                 // 128 x LDS
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int chunk = blockIdx.x, hk = blockIdx.y, rt = rt_override >= 0 ? rt_override : (int)blockIdx.z;
  const int ratio = Hq / Hkv;
  const int n_rows = B * ratio;
  const int step = *step_ptr;
  const int chunks_per_seq = cpad / DA_KEYS;

  // key slot range of this chunk
  int j0, nvalid;  // first slot, number of real keys
  int chunk_seq = -1;
  if (chunk < n_prefix_chunks) {
    j0 = chunk * DA_KEYS;
    nvalid = min(DA_KEYS, prefix_len - j0);
  } else {
    const int sc = chunk - n_prefix_chunks;
    chunk_seq = sc / chunks_per_seq;
    const int c0 = (sc - chunk_seq * chunks_per_seq) * DA_KEYS;
    j0 = prefix_len + chunk_seq * cpad + c0;
    nvalid = min(DA_KEYS, step + 1 - c0);
  }
  const long long ws_row0 = ((long long)chunk * Hkv + hk) * rows_pad + rt * 64;
  const int row_a = rt * 64 + warp * 16 + g, row_b = row_a + 8;
  const int seq_a = row_a / ratio, seq_b = row_b / ratio;
  // rows of other sequences see nothing in a completion chunk
  const bool tile_has_seq = chunk_seq < 0 || (chunk_seq >= (rt * 64) / ratio && chunk_seq <= min(n_rows - 1, rt * 64 + 63) / ratio);
  if (nvalid <= 0 || !tile_has_seq) {
    if (t4 == 0) {
      ws_ml[(ws_row0 + warp * 16 + g) * 2] = -INFINITY;
      ws_ml[(ws_row0 + warp * 16 + g) * 2 + 1] = 0.f;
      ws_ml[(ws_row0 + warp * 16 + g + 8) * 2] = -INFINITY;
      ws_ml[(ws_row0 + warp * 16 + g + 8) * 2 + 1] = 0.f;
    }
    return;
  }
  // ---- loads: Q rows (gathered per (seq, head)), K/V chunk ----
  for (int e = tid; e < 64 * 16; e += 128) {
    const int r = e >> 4, c = e & 15;
    const int row = rt * 64 + r;
    const bool ok = row < n_rows;
    const int sq = ok ? row / ratio : 0, hq = hk * ratio + (ok ? row % ratio : 0);
    cp_async_16(smem_u32(sQ + r * LDS + c * 8), q + (long long)sq * ldq + (long long)hq * HD + c * 8, ok);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    for (int e = tid; e < 64 * 16; e += 128) {
      const int r = half * 64 + (e >> 4), c = e & 15;
      const bool ok = r < nvalid;
      const long long slot = j0 + (ok ? r : 0);
      cp_async_16(smem_u32(sK + r * LDS + c * 8), Kc + (slot * Hkv + hk) * HD + c * 8, ok);
      cp_async_16(smem_u32(sV + r * LDS + c * 8), Vc + (slot * Hkv + hk) * HD + c * 8, ok);
    }
    cp_async_commit();
  }
  if (fuse_rope) {
    // q arrives un-rotated (q points at the qkv rows of this group); this step's k/v row of the chunk's sequence is
    // still in the qkv buffer: rotate q in shared memory, and let the CTA that owns slot `step` rotate k, patch its
    // shared-memory tiles and append the row to the global cache (what t1_decode_rope_append does as a separate launch).
    cp_async_wait<0>();
    __syncthreads();
    const float* cs = rope_cos + (long long)step * (HD / 2);
    const float* sn = rope_sin + (long long)step * (HD / 2);
    for (int e = tid; e < 64 * (HD / 2); e += 128) {
      const int r = e / (HD / 2), i = e - r * (HD / 2);
      if (rt * 64 + r < n_rows) {
        bf16* pq = sQ + r * LDS + i;
        float ra, rb;
        rope_rotate(__bfloat162float(pq[0]), __bfloat162float(pq[HD / 2]), cs[i], sn[i], ra, rb);
        pq[0] = __float2bfloat16(ra);
        pq[HD / 2] = __float2bfloat16(rb);
      }
    }
    const int lr = step - (j0 - prefix_len - chunk_seq * cpad);   // row of slot `step` inside this chunk
    if (chunk_seq >= 0 && lr >= 0 && lr < DA_KEYS) {
      const bf16* src = q + (long long)chunk_seq * ldq;
      const long long slot = (long long)j0 + lr;
      if (tid < HD / 2) {
        const bf16* pk = src + (long long)(Hq + hk) * HD + tid;
        float ra, rb;
        rope_rotate(__bfloat162float(pk[0]), __bfloat162float(pk[HD / 2]), cs[tid], sn[tid], ra, rb);
        const bf16 oa = __float2bfloat16(ra), ob = __float2bfloat16(rb);
        sK[lr * LDS + tid] = oa;
        sK[lr * LDS + tid + HD / 2] = ob;
        bf16* kd = Kw + (slot * Hkv + hk) * HD + tid;
        kd[0] = oa;
        kd[HD / 2] = ob;
      } else {
        const int j = (tid - HD / 2) * 2;
        const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(src + (long long)(Hq + Hkv + hk) * HD + j);
        *reinterpret_cast<__nv_bfloat162*>(sV + lr * LDS + j) = vv;
        *reinterpret_cast<__nv_bfloat162*>(Vw + (slot * Hkv + hk) * HD + j) = vv;
      }
    }
    __syncthreads();
  } else {
    cp_async_wait<1>();
    __syncthreads();
  }

  uint32_t qf[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks)
    ldmatrix_x4(qf[ks], smem_u32(sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8));

  float o_acc[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // a completion chunk belongs to ONE sequence: warps whose 16 query rows hold none of its heads have nothing to do
  const bool warp_active = chunk_seq < 0 || (chunk_seq * ratio < rt * 64 + warp * 16 + 16 && (chunk_seq + 1) * ratio > rt * 64 + warp * 16);
#pragma unroll 1
  for (int kt = 0; kt < DA_KEYS / 64; ++kt) {
    if (kt * 64 >= nvalid) break;
    if (kt == 1) {
      cp_async_wait<0>();
      __syncthreads();
    }
    if (!warp_active) continue;
    const bf16* kb = sK + kt * 64 * LDS;
    const bf16* vb = sV + kt * 64 * LDS;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t kf[4];
        ldmatrix_x4(kf, smem_u32(kb + (np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8));
        mma_bf16_16816(s[np * 2], qf[ks], kf);
        mma_bf16_16816(s[np * 2 + 1], qf[ks], kf + 2);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kidx = kt * 64 + nt * 8 + t4 * 2 + (e & 1);
        const int rseq = (e < 2) ? seq_a : seq_b;
        const bool vis = kidx < nvalid && (chunk_seq < 0 || chunk_seq == rseq);
        const float val = vis ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = val;
        if (e < 2) mx0 = fmaxf(mx0, val); else mx1 = fmaxf(mx1, val);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float mu0 = mn0 == -INFINITY ? 0.f : mn0, mu1 = mn1 == -INFINITY ? 0.f : mn1;
    const float al0 = exp2f(m0 - mu0), al1 = exp2f(m1 - mu1);
    m0 = mn0;
    m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - mu0), p1 = exp2f(s[nt][1] - mu0);
      const float p2 = exp2f(s[nt][2] - mu1), p3 = exp2f(s[nt][3] - mu1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    l0 = l0 * al0 + rs0;
    l1 = l1 * al1 + rs1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o_acc[i][0] *= al0; o_acc[i][1] *= al0; o_acc[i][2] *= al1; o_acc[i][3] *= al1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 8; ++dp) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, smem_u32(vb + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8));
        mma_bf16_16816(o_acc[dp * 2], pf[kk], vf);
        mma_bf16_16816(o_acc[dp * 2 + 1], pf[kk], vf + 2);
      }
    }
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // un-normalised partials (rows with l == 0 are never read by the combine kernel)
  float* oa = ws_o + (ws_row0 + warp * 16 + g) * HD;
  float* ob = ws_o + (ws_row0 + warp * 16 + g + 8) * HD;
  if (warp_active) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      *reinterpret_cast<float2*>(oa + i * 8 + t4 * 2) = make_float2(o_acc[i][0], o_acc[i][1]);
      *reinterpret_cast<float2*>(ob + i * 8 + t4 * 2) = make_float2(o_acc[i][2], o_acc[i][3]);
    }
  }
  if (t4 == 0) {
    ws_ml[(ws_row0 + warp * 16 + g) * 2] = m0;
    ws_ml[(ws_row0 + warp * 16 + g) * 2 + 1] = l0;
    ws_ml[(ws_row0 + warp * 16 + g + 8) * 2] = m1;
    ws_ml[(ws_row0 + warp * 16 + g + 8) * 2 + 1] = l1;
  }
}

// combine partials: one CTA of 8 warps per (seq, q head).  Warp w owns chunks w, w+8, ... (prefix chunks + the
// sequence's own completion chunks only); lane l owns head-dim columns 4l..4l+3 (float4).  All loads of a warp are
// independent, so the kernel costs ~one memory round trip; the 8 warp partials are merged through shared memory.
__global__ void __launch_bounds__(256)
decode_attn_combine_kernel(const float* __restrict__ ws_o, const float* __restrict__ ws_ml, int n_chunks,
                           int n_prefix_chunks, int chunks_per_seq, int Hq, int Hkv, int rows_pad,
                           bf16* __restrict__ out, long long ldo, const DecodeGroup* __restrict__ groups, int G) {
  __shared__ float s_m[8], s_l[8];
  __shared__ float4 s_acc[8][32];
  pdl_trigger();
  pdl_wait();
  int b = blockIdx.x / Hq;
  const int hq = blockIdx.x % Hq;
  if (groups) {
    const int gi = b / G;
    const DecodeGroup grp = groups[gi];
    out += (long long)gi * G * ldo;
    b -= gi * G;
    ws_o = grp.ws_o;
    ws_ml = grp.ws_ml;
    n_prefix_chunks = (grp.prefix_len + DA_KEYS - 1) / DA_KEYS;
  }
  const int ratio = Hq / Hkv, hk = hq / ratio;
  const int row = b * ratio + (hq - hk * ratio);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nc = n_prefix_chunks + chunks_per_seq;
  const long long stride = (long long)Hkv * rows_pad;
  const long long base = (long long)hk * rows_pad + row;
  constexpr int MAXC = 8;  // chunks per warp handled in registers (nc <= 64); larger nc loops
  float m_run = -INFINITY, l_run = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = warp; i0 < nc; i0 += 8 * MAXC) {
    float2 ml[MAXC];
    float4 o[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      const int i = i0 + j * 8;
      ml[j] = make_float2(-INFINITY, 0.f);
      o[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < nc) {
        const int c = i < n_prefix_chunks ? i : n_prefix_chunks + b * chunks_per_seq + (i - n_prefix_chunks);
        const long long r = (long long)c * stride + base;
        ml[j] = *reinterpret_cast<const float2*>(ws_ml + r * 2);
        if (ml[j].y > 0.f) o[j] = *reinterpret_cast<const float4*>(ws_o + r * 128 + lane * 4);
      }
    }
    float m_new = m_run;
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
      if (ml[j].y > 0.f) m_new = fmaxf(m_new, ml[j].x);
    if (m_new > -INFINITY) {
      const float sc = exp2f(m_run - m_new);  // m_run == -inf -> 0
      l_run *= sc;
      acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        if (ml[j].y > 0.f) {
          const float w = exp2f(ml[j].x - m_new);
          l_run += w * ml[j].y;
          acc.x += w * o[j].x; acc.y += w * o[j].y; acc.z += w * o[j].z; acc.w += w * o[j].w;
        }
      }
      m_run = m_new;
    }
  }
  if (lane == 0) {
    s_m[warp] = m_run;
    s_l[warp] = l_run;
  }
  s_acc[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
    float L = 0.f;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (s_m[w] > -INFINITY) {
        const float sc = exp2f(s_m[w] - M);
        const float4 a = s_acc[w][lane];
        L += sc * s_l[w];
        t.x += sc * a.x; t.y += sc * a.y; t.z += sc * a.z; t.w += sc * a.w;
      }
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
    uint2 pk;
    pk.x = pack_bf16x2(t.x * inv, t.y * inv);
    pk.y = pack_bf16x2(t.z * inv, t.w * inv);
    *reinterpret_cast<uint2*>(out + (long long)b * ldo + hq * 128 + lane * 4) = pk;
  }
  (void)n_chunks;
}

// =============================================================================================
// sampler: one CTA per sequence.
// =============================================================================================
__global__ void __launch_bounds__(1024, 1)
sample_kernel(const float* __restrict__ logits, long long ld, int V, float temperature, int top_k,
              unsigned long long seed_in, const unsigned long long* __restrict__ seed_ptr,
              const int* __restrict__ step_ptr, int step_offset, const int* __restrict__ eos_ids,
              int n_eos, int pad_id, int* __restrict__ finished, int* __restrict__ out_ids, int out_ld,
              int* __restrict__ next_ids, float* __restrict__ out_u) {
  __shared__ unsigned int hist[2048];
  __shared__ float red[33];
  __shared__ float wsum[32];
  __shared__ unsigned int s_prefix, s_krem;
  __shared__ unsigned long long s_best;
  __shared__ int s_tok;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int step = (step_ptr ? *step_ptr : 0) + step_offset;
  const unsigned long long seed = seed_ptr ? *seed_ptr : seed_in;  // device-resident seed keeps CUDA graphs reusable
  const float* x = logits + (long long)b * ld;
  const bool greedy = !(temperature > 0.f);
  const float inv_t = greedy ? 1.f : 1.f / temperature;
  int token = 0;

  if (greedy) {
    // argmax, lowest index wins ties
    unsigned long long best = 0;
    for (int i = tid; i < V; i += blockDim.x) {
      const unsigned long long cand = ((unsigned long long)float_key(x[i]) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)i);
      best = cand > best ? cand : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (tid == 0) s_best = 0;
    __syncthreads();
    if ((tid & 31) == 0) atomicMax(&s_best, best);
    __syncthreads();
    token = (int)(0xFFFFFFFFu - (uint32_t)(s_best & 0xFFFFFFFFu));
  } else {
    // ---- radix select of the k-th largest scaled logit (11 + 11 + 10 bits) ----
    uint32_t thr_key = 0;  // keep everything when top_k is disabled
    if (top_k > 0 && top_k < V) {
      if (tid == 0) { s_prefix = 0; s_krem = (unsigned)top_k; }
      const int shifts[3] = {21, 10, 0};
      const int bits[3] = {11, 11, 10};
      uint32_t mask = 0;
      for (int pass = 0; pass < 3; ++pass) {
        for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const int sh = shifts[pass];
        const uint32_t dmask = (1u << bits[pass]) - 1u;
        for (int i = tid; i < V; i += blockDim.x) {
          const uint32_t key = float_key(x[i] * inv_t);
          if ((key & mask) == prefix) atomicAdd(&hist[(key >> sh) & dmask], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          unsigned krem = s_krem;
          int d = (int)dmask;
          for (; d > 0; --d) {
            if (hist[d] >= krem) break;
            krem -= hist[d];
          }
          s_prefix = prefix | ((uint32_t)d << sh);
          s_krem = krem;
        }
        mask |= dmask << sh;
        __syncthreads();
      }
      thr_key = s_prefix;
    }
    // ---- softmax over the kept set {key >= thr_key}: contiguous index range per thread (deterministic) ----
    const int per = (V + blockDim.x - 1) / blockDim.x;
    const int i0 = tid * per, i1 = min(V, i0 + per);
    float mx = -INFINITY;
    for (int i = i0; i < i1; ++i) {
      const float v = x[i] * inv_t;
      if (float_key(v) >= thr_key) mx = fmaxf(mx, v);
    }
    const float gmax = block_max(mx, red);
    float part = 0.f;
    for (int i = i0; i < i1; ++i) {
      const float v = x[i] * inv_t;
      if (float_key(v) >= thr_key) part += expf(v - gmax);
    }
    // block exclusive scan of `part`
    const int lane = tid & 31, wid = tid >> 5;
    float inc = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      float w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      wsum[lane] = w;  // inclusive
    }
    __syncthreads();
    const float total = wsum[31];
    const float excl = inc - part + (wid > 0 ? wsum[wid - 1] : 0.f);
    uint32_t ctr[4] = {(uint32_t)b, (uint32_t)step, 0u, 0u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u = ((float)(ctr[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    if (out_u && tid == 0) out_u[b] = u;
    const float target = u * total;
    if (tid == 0) s_tok = 0x7fffffff;
    __syncthreads();
    if (part > 0.f && target >= excl && target < excl + part) {
      float run = excl;
      int pick = -1;
      for (int i = i0; i < i1; ++i) {
        const float v = x[i] * inv_t;
        if (float_key(v) >= thr_key) {
          run += expf(v - gmax);
          pick = i;
          if (target < run) break;
        }
      }
      if (pick >= 0) atomicMin(&s_tok, pick);
    }
    __syncthreads();
    if (s_tok == 0x7fffffff) {
      // rounding at the very end of the CDF: take the highest kept index
      if (tid == 0) s_best = 0;
      __syncthreads();
      unsigned long long best = 0;
      for (int i = i0; i < i1; ++i)
        if (float_key(x[i] * inv_t) >= thr_key) best = (unsigned long long)(i + 1);
      if (best) atomicMax(&s_best, best);
      __syncthreads();
      token = (int)s_best - 1;
    } else {
      token = s_tok;
    }
  }
  if (tid == 0) {
    int fin = finished ? finished[b] : 0;
    if (fin) token = pad_id;
    else
      for (int e = 0; e < n_eos; ++e)
        if (token == eos_ids[e]) fin = 1;
    if (finished) finished[b] = fin;
    if (out_ids) out_ids[(long long)b * out_ld + step] = token;
    if (next_ids) next_ids[b] = token;
  }
}

__global__ void incr_kernel(int* p, int v) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *p += v;
}

}  // namespace

extern "C" int t1_skinny_gemm(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                              long long ldy, int M, int N, int K, int out_f32, int accumulate, cudaStream_t stream) {
  T1_CHECK_ARG(M > 0 && M <= 64, "t1_skinny_gemm: M=%d must be in [1, 64]", M);
  T1_CHECK_ARG(N % 16 == 0 && K % 32 == 0, "t1_skinny_gemm: N %% 16 and K %% 32 required (N=%d K=%d)", N, K);
  T1_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "t1_skinny_gemm: ldx/ldw must be multiples of 8");
  const bf16* xb = (const bf16*)x;
  const bf16* wb = (const bf16*)W;
  const bf16* bb = (const bf16*)bias;
  int rc;
  if (M <= 8) rc = launch_skinny<1>(xb, ldx, wb, ldw, bb, y, ldy, M, N, K, out_f32, accumulate, stream);
  else if (M <= 16) rc = launch_skinny<2>(xb, ldx, wb, ldw, bb, y, ldy, M, N, K, out_f32, accumulate, stream);
  else if (M <= 32) rc = launch_skinny<4>(xb, ldx, wb, ldw, bb, y, ldy, M, N, K, out_f32, accumulate, stream);
  else rc = launch_skinny<8>(xb, ldx, wb, ldw, bb, y, ldy, M, N, K, out_f32, accumulate, stream);
  if (rc) return rc;
  T1_CUDA_LAUNCH_CHECK("t1_skinny_gemm");
  return T1_OK;
}

extern "C" int t1_decode_rope_append(void* qkv, long long ld, int B, int Hq, int Hkv, int head_dim, const float* cos_t,
                                     const float* sin_t, const int* step_ptr, void* Kc, void* Vc, int prefix_len,
                                     int cpad, cudaStream_t stream) {
  T1_CHECK_ARG(B > 0 && head_dim % 2 == 0, "t1_decode_rope_append: bad arguments");
  decode_rope_append_kernel<<<B, 256, 0, stream>>>((bf16*)qkv, ld, Hq, Hkv, head_dim, cos_t, sin_t, step_ptr, (bf16*)Kc,
                                                   (bf16*)Vc, prefix_len, cpad);
  T1_CUDA_LAUNCH_CHECK("t1_decode_rope_append");
  return T1_OK;
}

// workspace sizes (floats): ws_o = n_chunks*Hkv*rows_pad*128, ws_ml = n_chunks*Hkv*rows_pad*2,
// with n_chunks = ceil(prefix_len/128) + B*cpad/128 and rows_pad = ceil(B*Hq/Hkv / 64) * 64
extern "C" int t1_decode_attn(const void* q, long long ldq, const void* Kc, const void* Vc, int B, int Hq, int Hkv,
                              int head_dim, int prefix_len, int cpad, const int* step_ptr, float* ws_o, float* ws_ml,
                              void* out, long long ldo, float scale, cudaStream_t stream) {
  T1_CHECK_ARG(head_dim == 128, "t1_decode_attn: head_dim must be 128 (got %d)", head_dim);
  T1_CHECK_ARG(cpad % DA_KEYS == 0 && cpad > 0, "t1_decode_attn: cpad must be a positive multiple of %d", DA_KEYS);
  T1_CHECK_ARG(B > 0 && Hkv > 0 && Hq % Hkv == 0, "t1_decode_attn: bad head configuration");
  const int n_prefix_chunks = (prefix_len + DA_KEYS - 1) / DA_KEYS;
  const int chunks_per_seq = cpad / DA_KEYS;
  const int n_chunks = n_prefix_chunks + B * chunks_per_seq;
  const int ratio = Hq / Hkv;
  const int row_tiles = (B * ratio + 63) / 64;
  const int rows_pad = row_tiles * 64;
  constexpr int SMEM = (64 + 2 * DA_KEYS) * DA_LDS * 2;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  dim3 grid(n_chunks, Hkv, row_tiles);
  decode_attn_kernel<<<grid, 128, SMEM, stream>>>((const bf16*)q, ldq, (const bf16*)Kc, (const bf16*)Vc, Hq, Hkv, B,
                                                  prefix_len, cpad, step_ptr, ws_o, ws_ml, n_prefix_chunks, rows_pad,
                                                  scale * 1.4426950408889634f, nullptr, 0);
  T1_CUDA_LAUNCH_CHECK("t1_decode_attn");
  decode_attn_combine_kernel<<<B * Hq, 256, 0, stream>>>(ws_o, ws_ml, n_chunks, n_prefix_chunks, chunks_per_seq, Hq, Hkv,
                                                        rows_pad, (bf16*)out, ldo, nullptr, 0);
  T1_CUDA_LAUNCH_CHECK("t1_decode_attn.combine");
  return T1_OK;
}

// ---- grouped decode entry points (one launch for all prompt-groups; PDL-aware) ----
extern "C" int t1_decode_rope_append_groups(void* qkv, long long ld, int n_groups, int G, int Hq, int Hkv, int head_dim,
                                            const void* groups, const int* step_ptr, int cpad, cudaStream_t stream) {
  T1_CHECK_ARG(n_groups > 0 && G > 0 && head_dim % 2 == 0 && groups, "t1_decode_rope_append_groups: bad arguments");
  cudaError_t e = t1_launch(decode_rope_append_groups_kernel, dim3(n_groups * G), dim3(256), 0, stream, (bf16*)qkv, ld, G,
                            Hq, Hkv, head_dim, (const DecodeGroup*)groups, step_ptr, cpad);
  if (e != cudaSuccess) {
    t1_set_error("t1_decode_rope_append_groups: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  return T1_OK;
}

// Each group's workspaces: ws_o = n_chunks*Hkv*64*128 floats, ws_ml = n_chunks*Hkv*64*2 floats with
// n_chunks = ceil(prefix_len/128) + G*cpad/128; G*Hq/Hkv <= 64 query rows per KV head.
extern "C" int t1_decode_attn_groups(const void* q, long long ldq, const void* groups, int n_groups, int G, int Hq,
                                     int Hkv, int head_dim, int max_prefix_len, int cpad, const int* step_ptr, void* out,
                                     long long ldo, float scale, int fuse_rope, cudaStream_t stream) {
  T1_CHECK_ARG(head_dim == 128, "t1_decode_attn_groups: head_dim must be 128 (got %d)", head_dim);
  T1_CHECK_ARG(cpad % DA_KEYS == 0 && cpad > 0, "t1_decode_attn_groups: cpad must be a positive multiple of %d", DA_KEYS);
  T1_CHECK_ARG(n_groups > 0 && G > 0 && Hkv > 0 && Hq % Hkv == 0 && G * (Hq / Hkv) <= 64,
               "t1_decode_attn_groups: G * Hq/Hkv must be <= 64");
  const int chunks_per_seq = cpad / DA_KEYS;
  const int max_chunks = (max_prefix_len + DA_KEYS - 1) / DA_KEYS + G * chunks_per_seq;
  constexpr int SMEM = (64 + 2 * DA_KEYS) * DA_LDS * 2;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  cudaError_t e = t1_launch(decode_attn_kernel, dim3(max_chunks, Hkv, n_groups), dim3(128), (size_t)SMEM, stream,
                            (const bf16*)q, ldq, (const bf16*)nullptr, (const bf16*)nullptr, Hq, Hkv, G, 0, cpad, step_ptr,
                            (float*)nullptr, (float*)nullptr, 0, 64, scale * 1.4426950408889634f, (const DecodeGroup*)groups,
                            fuse_rope ? 1 : 0);
  if (e == cudaSuccess)
    e = t1_launch(decode_attn_combine_kernel, dim3(n_groups * G * Hq), dim3(256), 0, stream, (const float*)nullptr,
                  (const float*)nullptr, 0, 0, chunks_per_seq, Hq, Hkv, 64, (bf16*)out, ldo, (const DecodeGroup*)groups, G);
  if (e != cudaSuccess) {
    t1_set_error("t1_decode_attn_groups: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  t1_count_launch();
  return T1_OK;
}

extern "C" int t1_sample(const float* logits, long long ld, int B, int V, float temperature, int top_k,
                         unsigned long long seed, const unsigned long long* seed_ptr, const int* step_ptr,
                         int step_offset, const int* eos_ids, int n_eos, int pad_id, int* finished, int* out_ids,
                         int out_ld, int* next_ids, float* out_u, cudaStream_t stream) {
  T1_CHECK_ARG(B > 0 && V > 0 && logits, "t1_sample: bad arguments");
  cudaError_t le = t1_launch(sample_kernel, dim3(B), dim3(1024), 0, stream, logits, ld, V, temperature, top_k, seed, seed_ptr,
                             step_ptr, step_offset, eos_ids, n_eos, pad_id, finished, out_ids, out_ld, next_ids, out_u);
  if (le != cudaSuccess) {
    t1_set_error("t1_sample: launch failed: %s", cudaGetErrorString(le));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  return T1_OK;
}

extern "C" int t1_incr(int* p, int v, cudaStream_t stream) {
  cudaError_t le = t1_launch(incr_kernel, dim3(1), dim3(32), 0, stream, p, v);
  if (le != cudaSuccess) {
    t1_set_error("t1_incr: launch failed: %s", cudaGetErrorString(le));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  return T1_OK;
}
