// Device detokeniser + answer scanner: completion token ids -> UTF-8 bytes -> parsed [start, end],
// a validity flag and the format flag, without the completions ever leaving the GPU.  Together with t1_iou_reward this
// removes the device->host copy, tokenizer.batch_decode, Python regexes and host->device copy the reference puts between
// the scoring passes and the loss (src/time_r1/rl/timer1_trainer.py:642-698 -> main.py:122-239).
//
//   t1_detokenize    one CTA per completion row: token byte lengths -> block prefix sum -> byte copy.  Same result as
//                    tokenizers' ByteLevel decoder before its from_utf8_lossy (special tokens dropped when the table
//                    marks them, as batch_decode(skip_special_tokens=True) does).
//   t1_parse_answers one warp per row; the byte-level regex restatements live in answer_scan.cuh.
//
// Both are latency-bound byte work on a few KB per row (G rows per prompt-group): no tensor cores, no staging beyond L1.
#include "answer_scan.cuh"
#include "t1_common.cuh"

namespace {

constexpr int kDetokThreads = 256;

__global__ void __launch_bounds__(kDetokThreads)
detokenize_kernel(const int* __restrict__ ids, int C, const int* __restrict__ row_tokens, const int* __restrict__ tok_off,
                  const unsigned char* __restrict__ tok_bytes, const unsigned char* __restrict__ tok_skip, int V,
                  unsigned char* __restrict__ out, int cap, int* __restrict__ out_len) {
  __shared__ int warp_sum[kDetokThreads / 32];
  __shared__ int base_s;
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int ntok = row_tokens ? row_tokens[row] : C;
  ntok = ntok < 0 ? 0 : (ntok > C ? C : ntok);
  unsigned char* dst = out + (size_t)row * cap;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ntok; c0 += kDetokThreads) {
    const int c = c0 + tid;
    int id = -1, n = 0, src = 0;
    if (c < ntok) {
      id = ids[(size_t)row * C + c];
      if (id >= 0 && id < V && !(tok_skip && tok_skip[id])) {
        src = tok_off[id];
        n = tok_off[id + 1] - src;
      }
    }
    int incl = n;  // inclusive scan over the warp, then over the warps
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += y;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    int before = base_s;
    for (int k = 0; k < wid; ++k) before += warp_sum[k];
    const int start = before + incl - n;
    for (int k = 0; k < n; ++k)
      if (start + k < cap) dst[start + k] = tok_bytes[src + k];
    __syncthreads();
    if (tid == kDetokThreads - 1) base_s = before + incl;
    __syncthreads();
  }
  if (tid == 0) out_len[row] = base_s;  // > cap tells the caller the row was truncated (the C-ABI wrapper rules it out)
}

struct WarpPolicy {
  static constexpr int kLanes = 32;
  int lane_;
  __device__ __forceinline__ int lane() const { return lane_; }
  __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ static __forceinline__ int first_set(uint32_t m) { return __ffs((int)m) - 1; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

constexpr int kScanWarps = 4;

__global__ void __launch_bounds__(kScanWarps * 32)
parse_answers_kernel(const unsigned char* __restrict__ text, const int* __restrict__ len, int n, int cap, int mode,
                     double* __restrict__ out_s, double* __restrict__ out_e, unsigned char* __restrict__ out_valid,
                     unsigned char* __restrict__ out_format, float* __restrict__ out_format_f32) {
  const int row = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
  if (row >= n) return;  // whole warps leave together
  WarpPolicy w{(int)(threadIdx.x & 31)};
  int L = len[row];
  L = L < 0 ? 0 : (L > cap ? cap : L);
  const ascan::Parsed r = ascan::scan_row(text + (size_t)row * cap, L, mode, w);
  if (w.lane() == 0) {
    out_s[row] = r.s;
    out_e[row] = r.e;
    out_valid[row] = r.valid;
    if (out_format) out_format[row] = r.format;
    if (out_format_f32) out_format_f32[row] = r.format ? 1.0f : 0.0f;
  }
}

}  // namespace

// ids [n, C] int32 (device); row_tokens [n] int32 nullable = tokens of each row to decode (null: all C);
// tok_off [V + 1] int32 / tok_bytes [tok_off[V]] = the vocabulary's byte strings; tok_skip [V] uint8 nullable = tokens
// batch_decode(skip_special_tokens=True) drops; max_tok_bytes = longest vocabulary entry (host knowledge of the table).
// out_bytes [n, cap], out_len [n].  cap >= C * max_tok_bytes is required so that no row can be truncated.
extern "C" int t1_detokenize(const int* ids, int n, int C, const int* row_tokens, const int* tok_off,
                             const unsigned char* tok_bytes, const unsigned char* tok_skip, int V, int max_tok_bytes,
                             unsigned char* out_bytes, int cap, int* out_len, cudaStream_t stream) {
  T1_CHECK_ARG(ids && tok_off && tok_bytes && out_bytes && out_len, "t1_detokenize: null pointer");
  T1_CHECK_ARG(n >= 0 && C > 0 && V > 0 && max_tok_bytes >= 0, "t1_detokenize: bad sizes n=%d C=%d V=%d", n, C, V);
  T1_CHECK_ARG((long long)cap >= (long long)C * max_tok_bytes && cap > 0,
               "t1_detokenize: cap=%d is smaller than C * max_tok_bytes = %d * %d", cap, C, max_tok_bytes);
  if (n == 0) return T1_OK;
  detokenize_kernel<<<n, kDetokThreads, 0, stream>>>(ids, C, row_tokens, tok_off, tok_bytes, tok_skip, V, out_bytes, cap,
                                                     out_len);
  T1_CUDA_LAUNCH_CHECK("t1_detokenize");
  return T1_OK;
}

// text [n, cap] uint8 + len [n]: UTF-8 bytes of each completion.  mode 0: parse_timestamp_output + format_reward
// (main.py:122-142, 234-239); 1: extract_answer (evaluate.py:125-151); 2: extract_answer with the extract_answer_force
// fallback (calc_difficulty.py:40-60).  out_s / out_e [n] fp64 = float() of the two numbers (0 when nothing parsed),
// out_valid [n] uint8; out_format [n] uint8 and/or out_format_f32 [n] (1.0 / 0.0, the format_reward column) — nullable,
// written as 0 for modes 1 and 2.
extern "C" int t1_parse_answers(const unsigned char* text, const int* len, int n, int cap, int mode, double* out_s,
                                double* out_e, unsigned char* out_valid, unsigned char* out_format, float* out_format_f32,
                                cudaStream_t stream) {
  T1_CHECK_ARG(text && len && out_s && out_e && out_valid, "t1_parse_answers: null pointer");
  T1_CHECK_ARG(n >= 0 && cap > 0 && mode >= 0 && mode <= 2, "t1_parse_answers: bad arguments n=%d cap=%d mode=%d", n, cap, mode);
  if (n == 0) return T1_OK;
  parse_answers_kernel<<<(n + kScanWarps - 1) / kScanWarps, kScanWarps * 32, 0, stream>>>(text, len, n, cap, mode, out_s, out_e,
                                                                                        out_valid, out_format, out_format_f32);
  T1_CUDA_LAUNCH_CHECK("t1_parse_answers");
  return T1_OK;
}
