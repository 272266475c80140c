// Two-kernel top-k sampler for the decode chain (replaces HF's TopKLogitsWarper + softmax + torch.multinomial inside
// GenerationMixin._sample, driven from src/time_r1/rl/timer1_trainer.py:568-573; defaults temperature 1, top_k 50).
//
// The one-CTA-per-row sampler in decode.cu walks the 152k-entry row six times with 16 CTAs on the whole GPU
// (330 us per decode step at 7B).  Here the row is cut into 4096-entry chunks:
//   scan kernel  grid (chunks, rows): each CTA keeps its 16 values per thread in registers, radix-selects the chunk's
//                k-th largest key (4 x 8-bit digits, parallel suffix scan) and emits, in index order, every entry >= it
//                (<= 64 per chunk).  The global top-k set is a subset of the union of the per-chunk top-k sets.
//   pick kernel  grid (rows): compacts the candidates (still in index order), radix-selects the global k-th largest key,
//                soft-maxes the kept set and walks its CDF in index order with a Philox4x32-10 uniform keyed by
//                (seed, row, step) -- the same draw rule as t1_sample, so results are deterministic.
// temperature <= 0 is greedy arg-max (k = 1, lowest index wins ties).
#include "t1_common.cuh"

namespace {

constexpr int SC_CHUNK = 4096;
constexpr int SC_PER = 16;        // values per thread (256 threads)
constexpr int SC_CAND = 64;       // candidates kept per chunk (>= top_k)
constexpr int SC_MAX_CHUNKS = 64; // V <= 262144

// k-th largest of the keys selected by `n_local` / `get(j)` across the 256-thread block. All threads return the key.
template <typename GetKey>
__device__ __forceinline__ uint32_t block_radix_select(int n_local, GetKey get, unsigned k, unsigned* hist, unsigned* s_pk) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ unsigned s_wsum[8];
  uint32_t prefix = 0, mask = 0;
  unsigned krem = k;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    for (int j = 0; j < n_local; ++j) {
      const uint32_t key = get(j);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    // inclusive suffix sum S(t) = sum_{d >= t} hist[d]
    const unsigned c = hist[tid];
    unsigned s = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned n = __shfl_down_sync(0xffffffffu, s, o);
      if (lane + o < 32) s += n;
    }
    if (lane == 0) s_wsum[wid] = s;
    __syncthreads();
    unsigned above = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w > wid) above += s_wsum[w];
    s += above;
    if (s >= krem && s - c < krem) {   // exactly one digit holds the krem-th largest (when the total count >= krem)
      s_pk[0] = prefix | ((uint32_t)tid << shift);
      s_pk[1] = krem - (s - c);
    }
    if (pass == 0 && tid == 0 && s < krem) {   // fewer than k entries in total: keep everything
      s_pk[0] = 0;
      s_pk[1] = 0xffffffffu;
    }
    __syncthreads();
    prefix = s_pk[0];
    krem = s_pk[1];
    mask |= 255u << shift;
    if (krem == 0xffffffffu) return 0u;
  }
  return prefix;
}

__global__ void __launch_bounds__(256)
topk_scan_kernel(const float* __restrict__ logits, long long ld, int V, float inv_t, int k, int nch,
                 int* __restrict__ counts, float* __restrict__ cval, int* __restrict__ cidx) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_pk[2];
  __shared__ unsigned s_scan[8];
  pdl_trigger();
  pdl_wait();
  const int ch = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* x = logits + (long long)b * ld;
  const int base = ch * SC_CHUNK + tid * SC_PER;
  float val[SC_PER];
  uint32_t key[SC_PER];
  if (base + SC_PER <= V && ((reinterpret_cast<uintptr_t>(x + base) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < SC_PER / 4; ++j) {
      const float4 f = *reinterpret_cast<const float4*>(x + base + j * 4);
      val[j * 4 + 0] = f.x * inv_t; val[j * 4 + 1] = f.y * inv_t; val[j * 4 + 2] = f.z * inv_t; val[j * 4 + 3] = f.w * inv_t;
    }
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) key[j] = float_key(val[j]);
  } else {
#pragma unroll
    for (int j = 0; j < SC_PER; ++j) {
      const bool ok = base + j < V;
      val[j] = ok ? x[base + j] * inv_t : 0.f;
      key[j] = ok ? float_key(val[j]) : 0u;   // below every real value
    }
  }
  const uint32_t thr = block_radix_select(SC_PER, [&](int j) { return key[j]; }, (unsigned)k, hist, s_pk);
  // emit entries >= thr in index order (thread t owns 16 consecutive indices)
  unsigned mine = 0;
#pragma unroll
  for (int j = 0; j < SC_PER; ++j) mine += (key[j] >= thr && key[j] != 0u && base + j < V) ? 1u : 0u;
  unsigned inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_scan[wid] = inc;
  __syncthreads();
  unsigned off = inc - mine;
#pragma unroll
  for (int w = 0; w < 8; ++w)
    if (w < wid) off += s_scan[w];
  const long long slot0 = ((long long)b * nch + ch) * SC_CAND;
#pragma unroll
  for (int j = 0; j < SC_PER; ++j) {
    if (key[j] >= thr && key[j] != 0u && base + j < V) {
      if (off < SC_CAND) {
        cval[slot0 + off] = val[j];
        cidx[slot0 + off] = base + j;
      }
      ++off;
    }
  }
  if (tid == 255) counts[b * nch + ch] = (int)min(off, (unsigned)SC_CAND);   // thread 255's running offset = total
}

__global__ void __launch_bounds__(256)
topk_pick_kernel(const int* __restrict__ counts, const float* __restrict__ cval, const int* __restrict__ cidx, int nch,
                 int k, int greedy, unsigned long long seed_in, const unsigned long long* __restrict__ seed_ptr,
                 const int* __restrict__ step_ptr, int step_offset, const int* __restrict__ eos_ids, int n_eos, int pad_id,
                 int* __restrict__ finished, int* __restrict__ out_ids, int out_ld, int* __restrict__ next_ids,
                 float* __restrict__ out_u) {
  extern __shared__ __align__(16) uint8_t pk_smem[];
  float* sv = reinterpret_cast<float*>(pk_smem);                  // [nch * 64] candidate values (scaled logits)
  int* si = reinterpret_cast<int*>(sv + nch * SC_CAND);           // [nch * 64] token ids
  __shared__ unsigned hist[256];
  __shared__ unsigned s_pk[2];
  __shared__ int s_off[SC_MAX_CHUNKS + 1];
  __shared__ float red[33];
  __shared__ float wsum[8];
  __shared__ int s_tok;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int step = (step_ptr ? *step_ptr : 0) + step_offset;
  const unsigned long long seed = seed_ptr ? *seed_ptr : seed_in;
  // compact the per-chunk candidate lists (chunk order == index order)
  if (tid == 0) {
    int run = 0;
    for (int c = 0; c < nch; ++c) {
      s_off[c] = run;
      run += counts[b * nch + c];
    }
    s_off[nch] = run;
  }
  __syncthreads();
  const int n = s_off[nch];
  for (int e = tid; e < nch * SC_CAND; e += 256) {
    const int c = e / SC_CAND, r = e - c * SC_CAND;
    if (r < s_off[c + 1] - s_off[c]) {
      sv[s_off[c] + r] = cval[((long long)b * nch + c) * SC_CAND + r];
      si[s_off[c] + r] = cidx[((long long)b * nch + c) * SC_CAND + r];
    }
  }
  __syncthreads();
  const int per = (n + 255) / 256;
  const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
  const uint32_t thr = block_radix_select(i1 - i0, [&](int j) { return float_key(sv[i0 + j]); }, (unsigned)k, hist, s_pk);
  int token = 0;
  float mx = -INFINITY;
  for (int i = i0; i < i1; ++i)
    if (float_key(sv[i]) >= thr) mx = fmaxf(mx, sv[i]);
  const float gmax = block_max(mx, red);
  if (tid == 0) s_tok = 0x7fffffff;
  __syncthreads();
  if (greedy) {
    for (int i = i0; i < i1; ++i)
      if (sv[i] == gmax) {
        atomicMin(&s_tok, si[i]);
        break;
      }
    __syncthreads();
    token = s_tok;
  } else {
    float part = 0.f;
    for (int i = i0; i < i1; ++i)
      if (float_key(sv[i]) >= thr) part += expf(sv[i] - gmax);
    float inc = part;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float nb = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nb;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    float before = 0.f, total = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < wid) before += wsum[w];
      total += wsum[w];
    }
    const float excl = before + inc - part;
    uint32_t ctr[4] = {(uint32_t)b, (uint32_t)step, 0u, 0u};
    philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float u = ((float)(ctr[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    if (out_u && tid == 0) out_u[b] = u;
    const float target = u * total;
    if (part > 0.f && target >= excl && target < excl + part) {
      float run = excl;
      int pick = -1;
      for (int i = i0; i < i1; ++i) {
        if (float_key(sv[i]) >= thr) {
          run += expf(sv[i] - gmax);
          pick = si[i];
          if (target < run) break;
        }
      }
      if (pick >= 0) atomicMin(&s_tok, pick);
    }
    __syncthreads();
    if (s_tok == 0x7fffffff) {
      // rounding at the very end of the CDF: take the highest kept index
      __syncthreads();
      if (tid == 0) s_tok = -1;
      __syncthreads();
      int best = -1;
      for (int i = i0; i < i1; ++i)
        if (float_key(sv[i]) >= thr) best = si[i];
      if (best >= 0) atomicMax(&s_tok, best);
      __syncthreads();
    }
    token = s_tok;
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

}  // namespace

extern "C" long long t1_sample_topk_workspace_bytes(int B, int V) {
  const long long nch = (V + SC_CHUNK - 1) / SC_CHUNK;
  return (long long)B * nch * (4 + SC_CAND * 8) + 256;
}

// Same contract as t1_sample for greedy (temperature <= 0) or 1 <= top_k <= 64, V <= 262144, with a caller-provided
// workspace of t1_sample_topk_workspace_bytes(B, V) bytes.
extern "C" int t1_sample_topk(const float* logits, long long ld, int B, int V, float temperature, int top_k,
                              unsigned long long seed, const unsigned long long* seed_ptr, const int* step_ptr,
                              int step_offset, const int* eos_ids, int n_eos, int pad_id, int* finished, int* out_ids,
                              int out_ld, int* next_ids, float* out_u, void* workspace, long long workspace_bytes,
                              cudaStream_t stream) {
  T1_CHECK_ARG(B > 0 && V > 0 && logits && workspace, "t1_sample_topk: bad arguments");
  const bool greedy = !(temperature > 0.f);
  T1_CHECK_ARG(greedy || (top_k >= 1 && top_k <= SC_CAND), "t1_sample_topk: top_k=%d must be in [1, %d] (use t1_sample)", top_k, SC_CAND);
  const int nch = (V + SC_CHUNK - 1) / SC_CHUNK;
  T1_CHECK_ARG(nch <= SC_MAX_CHUNKS, "t1_sample_topk: V=%d too large (use t1_sample)", V);
  T1_CHECK_ARG(workspace_bytes >= t1_sample_topk_workspace_bytes(B, V), "t1_sample_topk: workspace too small");
  T1_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "t1_sample_topk: workspace must be 16-byte aligned");
  int* counts = reinterpret_cast<int*>(workspace);
  const long long n_slots = (long long)B * nch * SC_CAND;
  float* cval = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + (((long long)B * nch * 4 + 255) & ~255LL));
  int* cidx = reinterpret_cast<int*>(cval + n_slots);
  const int k = greedy ? 1 : top_k;
  const float inv_t = greedy ? 1.f : 1.f / temperature;
  cudaError_t e = t1_launch(topk_scan_kernel, dim3(nch, B), dim3(256), 0, stream, logits, ld, V, inv_t, k, nch, counts, cval, cidx);
  if (e == cudaSuccess) {
    const size_t smem = (size_t)nch * SC_CAND * 8;
    e = t1_launch(topk_pick_kernel, dim3(B), dim3(256), smem, stream, (const int*)counts, (const float*)cval, (const int*)cidx,
                  nch, k, greedy ? 1 : 0, seed, seed_ptr, step_ptr, step_offset, eos_ids, n_eos, pad_id, finished, out_ids,
                  out_ld, next_ids, out_u);
  }
  if (e != cudaSuccess) {
    t1_set_error("t1_sample_topk: launch failed: %s", cudaGetErrorString(e));
    return T1_ERR_CUDA;
  }
  t1_count_launch();
  t1_count_launch();
  return T1_OK;
}
