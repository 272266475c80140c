/* timer1.h — C ABI of libtimer1.so, the B200-native (sm_100a) kernel library behind Time-R1's GRPO hot path.
 *
 * The reference (xiaomi-research/time-r1) has no FFI: its "operator API" is the Python surface main.py uses
 * (src/time_r1/__init__.py:1-6, TimeR1_Trainer) and every FLOP runs inside third-party Python packages.  This
 * header is therefore the boundary a maintainer binds with ctypes; each entry point cites the
 * reference call site whose arithmetic it replaces.  Paths are relative to /root/reference; "Q25:" is
 * transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py (the module timer1_trainer.py:244 instantiates).
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error (T1_ERR_*); t1_last_error() returns a thread-local message
 *     owned by the library, valid until the next failing call on that thread;
 *   - all tensor arguments are DEVICE pointers owned by the caller, row-major; dtype is stated per argument
 *     (bf16 = __nv_bfloat16); the library never allocates or frees caller memory — workspaces are caller-provided;
 *   - leading dimensions (ld*) are in ELEMENTS; unless stated they must be multiples of 8 (16-byte rows);
 *   - every call is asynchronous on the given cudaStream_t and capturable into a CUDA graph;
 *   - no torch types, no host callbacks; one host thread per device.
 */
#ifndef TIMER1_H
#define TIMER1_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T1_OK 0
#define T1_ERR_ARG (-1)
#define T1_ERR_CUDA (-2)
#define T1_ERR_UNSUPPORTED (-3)

/* t1_gemm_bf16_tn flags */
#define T1_GEMM_OUT_F32 1
#define T1_GEMM_ACCUMULATE 2
#define T1_GEMM_BIAS_F32 4

/* ---- core --------------------------------------------------------------------------------------------- */
const char* t1_last_error(void);
int t1_version(void);
int t1_device_info(int* out4 /* host: {sm_count, cc_major, cc_minor, l2_bytes} */);
void t1_count_launch(void);
unsigned long long t1_launch_count(int reset); /* kernels launched through this library (bench "gpu_launches") */

/* ---- dense contractions -------------------------------------------------------------------------------- */
/* D[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N]).  tcgen05/TMEM/TMA kernel.  Replaces every nn.Linear / Conv3d the
 * reference executes through cuBLAS/cuDNN: Q25:106-114 (patch embed), :214-215 (ViT qkv/proj), :77-88 (ViT MLP),
 * :133-146 (merger), :704-707 (LLM q/k/v/o), :611-624 (LLM MLP), lm_head (timer1_trainer.py:452-457), and their
 * autograd backward (HF trainer.py:1934).  A, B bf16; D bf16 or fp32 (T1_GEMM_OUT_F32); N % 8 == 0. */
int t1_gemm_bf16_tn(const void* A, long long lda, const void* B, long long ldb, void* D, long long ldd, int M, int N,
                    int K, const void* bias, int flags, cudaStream_t stream);
/* act[M, I] = silu(A Wg^T + bg) * (A Wu^T + bu), W = [Wg; Wu] ([2I, K] K-major), bias [2I] bf16 or NULL: gate_up projection
 * with the SwiGLU (HF Q25:611-624, ViT Q25:82-88) in the GEMM epilogue; gate_up ([M, 2I] bf16 or NULL) also receives the
 * projection for the backward pass.  Bit-identical to t1_gemm_bf16_tn + t1_swiglu. */
int t1_gemm_swiglu(const void* A, long long lda, const void* W, long long ldw, const void* bias, void* act, long long ld_act,
                   void* gate_up, long long ld_gate_up, int M, int I, int K, cudaStream_t stream);
/* Cap on the persistent grid of subsequent t1_gemm_* launches (0 = one CTA per SM): a side stream that overlaps a
 * bandwidth-bound phase (the decode) keeps to a few SMs. */
void t1_gemm_max_ctas(int n);

/* y[M<=64, N] (+)= x[M,K] * W[N,K]^T (+ bias bf16).  Weight-streaming decode GEMM (HBM-bound; thread-block-cluster
 * split-K).  Replaces the per-token linears inside HF generate (timer1_trainer.py:568-573).  N % 16, K % 32. */
int t1_skinny_gemm(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                   long long ldy, int M, int N, int K, int out_f32, int accumulate, cudaStream_t stream);

/* Same contract on the tensor core: y[M<=64, N] = x[M,K] * W[N,K]^T (+ bias bf16), y bf16 or fp32.  Stream-K over
 * (128-row slab, 64-wide k-block) units, one contiguous range per SM; TMA -> shared-memory ring -> tcgen05.mma with
 * the weight rows as the UMMA M dimension and the batch rows as N; deterministic last-arriver fix-up for slabs shared
 * by several CTAs.  Under t1_set_pdl(1) the weight tiles are fetched BEFORE griddepcontrol.wait, so the stream of
 * launch k+1 overlaps the tail of launch k and the latency-bound kernels between them.  K % 8 == 0.
 * workspace: t1_stream_gemm_workspace_bytes(max M) bytes, zero-filled once by the caller, never shared by launches
 * that may run concurrently. */
long long t1_stream_gemm_workspace_bytes(int max_m);
int t1_stream_gemm(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y, long long ldy,
                   int M, int N, int K, int out_f32, void* workspace, long long workspace_bytes, cudaStream_t stream);
/* Scheduling of the weight stream.  0: static stream-K split (equal bytes per CTA; leaves a 15-20% finishing-time spread
 * because per-SM streaming rates differ).  1 (default): GEMMs with >= 4 slabs per SM (lm_head) hand whole slabs out
 * through an atomic counter, so faster SMs take more; the others stay static.  2: also split slabs into k-segments
 * (measured slower: per-item partial tiles). */
void t1_stream_gemm_dynamic(int on);
/* Cap on the number of CTAs of subsequent t1_stream_gemm* launches (0 = one per SM): leaves SMs to a concurrent stream. */
void t1_stream_gemm_max_ctas(int n);
/* y = x W^T (bf16); resid[M, N] += y in place; h_out = rmsnorm(resid, eps) * norm_w: the o / down projection of a decode
 * layer fused with the residual add and the NEXT RMSNorm (HF Q25:805-827): after a grid-wide arrive counter CTA b
 * normalises row b (every CTA of the launch is resident: grid <= SM count).
 * Replaces t1_stream_gemm + t1_rmsnorm(delta = y).  resid / h_out contiguous [M, N], N % 8 == 0. */
int t1_stream_gemm_resnorm(const void* x, long long ldx, const void* W, long long ldw, void* y, long long ldy, int M, int N,
                           int K, void* resid, const void* norm_w, float eps, void* h_out, void* workspace,
                           long long workspace_bytes, cudaStream_t stream);
/* RMSNorm FOLDED into the two GEMMs around it, no norm kernel and no grid-wide wait (decode chain o_proj -> [residual add,
 * post-attention norm] -> gate_up and down_proj -> [residual add, next input norm] -> qkv; HF Q25:805-827).
 * Producer half: y = x W^T is not stored; resid[M, N] += bf16(y) in place, z_out[M, N] = bf16(norm_w[n] * resid[m, n]) and
 * parts_out[N / 128][MP] (MP = 16 / 32 / 64 for M <= 16 / 32 / 64) = per-row sums of resid^2 over each 128-column slab.
 * N % 128 == 0; resid / z_out contiguous [M, N]. */
int t1_stream_gemm_fold(const void* x, long long ldx, const void* W, long long ldw, int M, int N, int K, void* resid,
                        const void* norm_w, void* z_out, float* parts_out, void* workspace, long long workspace_bytes,
                        cudaStream_t stream);
/* Consumer half: y[m, n] = r[m] * sum_k W[n, k] x[m, k] (+ bias[n]) with r[m] = rsqrt(sum_i in_parts[i][m] / K + eps): the
 * row factor of the RMSNorm is constant along k, so it multiplies the fp32 accumulator in the epilogue; with x = z_out and
 * in_parts = parts_out of t1_stream_gemm_fold this is W . rmsnorm(resid).  swiglu != 0: N = 2 * I and y is [M, I] as in
 * t1_stream_gemm_swiglu. */
int t1_stream_gemm_scaled(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y, long long ldy,
                          int M, int N, int K, int swiglu, const float* in_parts, int in_nparts, float eps, void* workspace,
                          long long workspace_bytes, cudaStream_t stream);
/* t1_stream_gemm_fold(x, W_a -> resid, z, parts) immediately followed by t1_stream_gemm_scaled(z, W_b, parts -> y) as ONE
 * launch: the o_proj -> [residual, post-attention norm] -> gate_up (swiglu != 0: N_b = 2 * I, y [M, I]) and down_proj ->
 * [residual, next layer's input norm] -> qkv (+ bias) pairs of one decode step (HF Q25:805-827; the nn.Linear calls of
 * GenerationMixin._sample driven from timer1_trainer.py:568-573).  The boundary between the two GEMMs is a device-side arrival
 * counter instead of a kernel boundary and the weight stream runs through it; same unit split and summation order as the two
 * separate launches, so y, resid, z and parts are bit-identical to them.  N_a % 128 == 0; the grid (<= one CTA per SM) must be
 * resident at once. */
int t1_stream_gemm_fold_pair(const void* x, long long ldx, const void* W_a, long long ldw_a, int M, int N_a, int K_a, void* resid,
                             const void* norm_w, void* z, float* parts, const void* W_b, long long ldw_b, const void* bias_b,
                             void* y, long long ldy, int N_b, int swiglu, float eps, void* workspace, long long workspace_bytes,
                             cudaStream_t stream);
/* y[M, I] = silu(x Wg^T) * (x Wu^T), W = [Wg; Wu] ([2I, K]): the decode step's gate_up projection fused with the SwiGLU
 * (HF Q25:611-624) in one weight stream; same arithmetic per element as t1_stream_gemm (bf16) + t1_swiglu (bit-identical when
 * the stream-K cut points coincide, as for the 7B decode shapes).  I % 64 == 0. */
int t1_stream_gemm_swiglu(const void* x, long long ldx, const void* W, long long ldw, const void* bias, void* y,
                          long long ldy, int M, int I, int K, void* workspace, long long workspace_bytes,
                          cudaStream_t stream);
/* Tuning: shared-memory ring depth (0 = default, ~108 KB per CTA) and CTAs per SM (1 or 2) of later launches. */
void t1_stream_gemm_tune(int stages, int ctas_per_sm);
/* One-shot performance hint for the NEXT t1_stream_gemm* launch (static split): while it waits for its dependency its idle
 * warps prefetch into L2 up to max_bytes of W_next [N, K] -- the units the launch AFTER it (described by M, N, K, swiglu
 * layout, fold_producer = t1_stream_gemm_fold) reads first after filling its shared-memory ring -- so the HBM keeps
 * streaming weights through the latency-bound qkv -> attention -> o part of a decode layer (HF GenerationMixin._sample's
 * per-token nn.Linear chain, timer1_trainer.py:568-573).  Results never depend on it; max_bytes <= 0 clears the hint. */
int t1_stream_gemm_prefetch_next(const void* W_next, long long ldw, int M, int N, int K, int swiglu, int fold_producer,
                                 long long max_bytes);
/* Diagnostics: CTA 0 of launch i writes 4 %globaltimer stamps {start, weight prefetch issued, dependency resolved,
 * done} to buf[4*(i % slots)] (device memory, uint64); buf = NULL switches it off. */
void t1_stream_gemm_debug(void* buf, int slots);
/* Programmatic dependent launch switch for the decode chain (gather_rows, rmsnorm, stream_gemm, decode_*_groups,
 * swiglu, sample, incr): while on, those entry points launch with
 * cudaLaunchAttributeProgrammaticStreamSerialization; each of their kernels executes griddepcontrol.wait before
 * touching memory a predecessor may own.  Process-wide, not thread-safe; default off. */
void t1_set_pdl(int on);

/* ---- attention ------------------------------------------------------------------------------------------ */
/* Flash attention over a packed token axis.  Key j is visible to query row i iff
 *     j < row_hi[i] && (j >= row_lo[i] || j < shared_len).
 * Replaces flash-attn-2 varlen: ViT window/full attention Q25:244-261,498-509 (non-causal: lo/hi = window bounds),
 * LLM causal attention Q25:740-755 (lo = 0, hi = i+1) and — new — the prefix-sharing packed scoring layout
 * (prompt rows causal; completion rows lo = own start, hi = i+1, shared_len = P).
 * Q [T, Hq, hd] / K,V [Tk, Hkv, hd] bf16 with token strides ldq/ldk/ldv; O [T, Hq*hd] bf16; LSE2 [Hq, T] fp32 in
 * log2 domain (nullable).  head_dim in {64, 80, 128}. */
int t1_attn_fwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                long long ldo, float* LSE2, const int* row_lo, const int* row_hi, int shared_len, int T, int Tk, int Hq,
                int Hkv, int head_dim, float scale, cudaStream_t stream);
/* Same contract, head_dim == 128: tcgen05/TMEM/TMA kernel (S and O accumulators in TMEM, V consumed as an MN-major
 * UMMA operand, lazy-rescaled online softmax).  Used for the tensor-bound LLM prefill / scoring shapes.  Only query
 * rows [q_row0, T) are processed; Q and O point at row q_row0 (lets policy scoring reuse the prompt K/V of the
 * rollout's prefill instead of re-scoring the prompt). */
int t1_attn_fwd_tc(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                   long long ldo, float* LSE2, long long lse_ld, const int* row_lo, const int* row_hi, int shared_len, int T,
                   int Tk, int Hq, int Hkv, int head_dim, float scale, int q_row0, cudaStream_t stream);
/* Persistent tcgen05 attention for BLOCK-LOCAL masks: every row's [row_lo, row_hi) lies inside the row's own 128-row tile — the
 * ViT's 64-token windows (HF modeling_qwen2_5_vl.py:244-261 with cu_window_seqlens; the reference runs them through
 * flash_attn_varlen_func).  Q, K, V: [T, >= H * head_dim] bf16 views with the same head count, head_dim 80 or 128, no LSE. */
int t1_attn_fwd_tc_win(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
                       long long ldo, const int* row_lo, const int* row_hi, int T, int H, int head_dim, float scale,
                       cudaStream_t stream);
/* Backward of the above (autograd through flash-attn in the reference).  Workspaces: delta_ws [Hq*T] fp32,
 * summ_ws [4*ceil(T/64)] int32.  head_dim in {64, 128}. */
int t1_attn_bwd(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, const void* O,
                long long ldo, const void* dO, long long lddo, const float* LSE2, float* delta_ws, int* summ_ws, void* dQ,
                long long lddq, void* dK, long long lddk, void* dV, long long lddv, const int* row_lo, const int* row_hi,
                int shared_len, int T, int Tk, int Hq, int Hkv, int head_dim, float scale, cudaStream_t stream);

/* tcgen05/TMEM backward for head_dim 128 (same contract as t1_attn_bwd; replaces the autograd pass through the
 * flash-attn call at timer1_trainer.py:452-457 / HF trainer.py:1927-1934).  ws: t1_attn_bwd_tc_workspace_bytes(). */
long long t1_attn_bwd_tc_workspace_bytes(int T, int Tk, int Hq);
int t1_attn_bwd_tc(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, const void* O,
                   long long ldo, const void* dO, long long lddo, const float* LSE2, void* ws, long long ws_bytes, void* dQ,
                   long long lddq, void* dK, long long lddk, void* dV, long long lddv, const int* row_lo, const int* row_hi,
                   int shared_len, int T, int Tk, int Hq, int Hkv, int head_dim, float scale, cudaStream_t stream);

/* ---- decode (HF GenerationMixin._sample, timer1_trainer.py:568-573) -------------------------------------- */
/* Rotate q,k of qkv [B, (Hq+2Hkv)*hd] with table row *step_ptr and append k,v at cache slot
 * prefix_len + b*cpad + *step_ptr.  Kc/Vc [prefix_len + B*cpad, Hkv*hd] bf16. */
int t1_decode_rope_append(void* qkv, long long ld, int B, int Hq, int Hkv, int head_dim, const float* cos_t,
                          const float* sin_t, const int* step_ptr, void* Kc, void* Vc, int prefix_len, int cpad,
                          cudaStream_t stream);
/* Split-KV decode attention over [shared prompt prefix | per-sequence completion slots] + combine.
 * Workspaces (fp32): ws_o = n_chunks*Hkv*rows_pad*128, ws_ml = n_chunks*Hkv*rows_pad*2 with
 * n_chunks = ceil(prefix_len/128) + B*cpad/128, rows_pad = ceil(B*Hq/Hkv/64)*64.  head_dim = 128, cpad % 128 == 0. */
int t1_decode_attn(const void* q, long long ldq, const void* Kc, const void* Vc, int B, int Hq, int Hkv, int head_dim,
                   int prefix_len, int cpad, const int* step_ptr, float* ws_o, float* ws_ml, void* out, long long ldo,
                   float scale, cudaStream_t stream);
/* Multi-group forms used by the captured decode step: ONE launch covers the n_groups x G sequences of all
 * prompt-groups decoded together.  `groups` is a device array of n_groups t1_decode_group records (this layer's K/V
 * caches, split-KV workspaces, decode-position rope tables and prompt length of each group); qkv / q / out rows are
 * ordered group-major.  G * Hq/Hkv <= 64. */
typedef struct t1_decode_group {
  void* K;              /* [prefix_len + G*cpad, Hkv*hd] bf16 */
  void* V;
  float* ws_o;          /* n_chunks*Hkv*64*128 floats, n_chunks = ceil(prefix_len/128) + G*cpad/128 */
  float* ws_ml;         /* n_chunks*Hkv*64*2 floats */
  const float* cos_t;   /* [C, hd/2] */
  const float* sin_t;
  int prefix_len;
  int reserved;
} t1_decode_group;
int t1_decode_rope_append_groups(void* qkv, long long ld, int n_groups, int G, int Hq, int Hkv, int head_dim,
                                 const void* groups, const int* step_ptr, int cpad, cudaStream_t stream);
/* fuse_rope != 0: q is the UN-rotated qkv buffer [n_groups*G, (Hq+2Hkv)*hd]; the kernel rotates q on load and the CTA
 * that owns cache slot *step_ptr rotates k and appends k,v itself (no t1_decode_rope_append_groups launch needed). */
int t1_decode_attn_groups(const void* q, long long ldq, const void* groups, int n_groups, int G, int Hq, int Hkv,
                          int head_dim, int max_prefix_len, int cpad, const int* step_ptr, void* out, long long ldo,
                          float scale, int fuse_rope, cudaStream_t stream);
/* The captured decode step's attention in ONE launch (tcgen05): a thread-block cluster of cluster_size CTAs (1, 2, 4
 * or 8; 0 = 8) per (group, KV head) streams an even share of the cached 128-key units by TMA, treats this step's token
 * (k, v still in the un-rotated qkv rows `q`) as one more split-KV partial while rotating q / k and appending k, v to
 * the cache, and merges the partials through distributed shared memory (no workspace, no combine launch; ws_o / ws_ml
 * of the group records are unused).  maps: HOST array [n_groups][2] of 128-byte tensor maps over each group's K and
 * V cache ([rows, Hkv*128] bf16, box 128 rows x 64 columns) produced by t1_encode_tmap_bf16; they travel as kernel parameters
 * (n_groups <= 8).  The cache must hold finite values. */
int t1_encode_tmap_bf16(void* out_128_bytes_host, const void* ptr, long long rows, long long cols, long long ld,
                        int box_rows, int box_cols);
int t1_decode_attn_cluster(const void* q, long long ldq, const void* groups, const void* maps, int n_groups, int G,
                           int Hq, int Hkv, int head_dim, int cpad, const int* step_ptr, void* out, long long ldo,
                           float scale, int cluster_size, cudaStream_t stream);
/* The same CTA program WITHOUT clusters, sized to fill the machine: `splits` plain CTAs per (group, KV head) (the plan
 * picks floor(SMs / (n_groups * Hkv)) <= 48: 18 at cfg2 = 144 of 148 SMs, where the 8-CTA clusters use 64), one even
 * deal of [prefix units | completion units]; the (m, l, O) partials go through an L2-resident workspace, each CTA
 * arrives on a per-(group, head) counter, spins until its peers have arrived and merges the rows it owns -- still one
 * launch, no combine kernel, deterministic.  Replaces the same reference call site as t1_decode_attn_cluster (HF
 * GenerationMixin._sample's per-token attention, src/time_r1/rl/timer1_trainer.py:568-573).
 * t1_decode_attn_wide_plan: requested = 0 picks splits for the current device; *ws_bytes = workspace size.  The caller
 * allocates ws (256-byte aligned) and ZEROES it once; every launch leaves the counters zero again.  All CTAs of the
 * launch must be co-resident (checked: splits * Hkv * n_groups <= SM count). */
int t1_decode_attn_wide_plan(int n_groups, int Hkv, int requested, int* splits, long long* ws_bytes);
int t1_decode_attn_wide(const void* q, long long ldq, const void* groups, const void* maps, int n_groups, int G, int Hq,
                        int Hkv, int head_dim, int cpad, const int* step_ptr, void* out, long long ldo, float scale,
                        int splits, void* ws, long long ws_bytes, cudaStream_t stream);
/* temperature -> top-k -> softmax -> multinomial (Philox4x32-10 keyed by seed, counter (row, step)); temperature <= 0
 * is greedy arg-max.  HF 4.51.1 generate defaults.  logits [B, V] fp32 (ld may be 0: shared row);
 * seed_ptr (device, nullable) overrides seed so captured graphs stay seed-independent; writes
 * out_ids[b, *step_ptr + step_offset], next_ids[b]; finished[b] latches on any of eos_ids, then emits pad_id. */
int t1_sample(const float* logits, long long ld, int B, int V, float temperature, int top_k, unsigned long long seed,
              const unsigned long long* seed_ptr, const int* step_ptr, int step_offset, const int* eos_ids, int n_eos,
              int pad_id, int* finished, int* out_ids, int out_ld, int* next_ids, float* out_u, cudaStream_t stream);
/* Same contract, for greedy or 1 <= top_k <= 64 and V <= 262144: two short kernels instead of one CTA per row (the
 * row is scanned in 4096-entry chunks by chunks x B CTAs that emit per-chunk top-k candidates; one CTA per row then
 * selects the global top-k among them and draws).  Same Philox draw and index-ordered CDF walk as t1_sample.
 * workspace: t1_sample_topk_workspace_bytes(B, V) bytes, 16-byte aligned. */
long long t1_sample_topk_workspace_bytes(int B, int V);
int t1_sample_topk(const float* logits, long long ld, int B, int V, float temperature, int top_k, unsigned long long seed,
                   const unsigned long long* seed_ptr, const int* step_ptr, int step_offset, const int* eos_ids, int n_eos,
                   int pad_id, int* finished, int* out_ids, int out_ld, int* next_ids, float* out_u, void* workspace,
                   long long workspace_bytes, cudaStream_t stream);
int t1_incr(int* p, int v, cudaStream_t stream);

/* ---- token-wise kernels (all bf16 storage, fp32 math) ----------------------------------------------------- */
/* y = w * rmsnorm(x + delta); with delta the rounded sum is written to x_out (may alias x).  Q25:66-71 + the
 * residual adds of Q25:313-321,805-827 fused in. */
int t1_rmsnorm(const void* x, const void* delta, void* x_out, const void* w, void* y, int T, int D, float eps,
               cudaStream_t stream);
/* LayerNorm variant for the Qwen2-VL ViT (modeling_qwen2_vl.py:461-464). */
int t1_layernorm(const void* x, const void* delta, void* x_out, const void* w, const void* b, void* y, int T, int D,
                 float eps, cudaStream_t stream);
/* dx = dres + d(rmsnorm)/dx; dw (fp32, caller-zeroed) += sum_t dy * xhat. */
int t1_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, float* dw, int T, int D,
                   float eps, cudaStream_t stream);
/* LayerNorm backward for the Qwen2-VL PatchMerger ln_q (modeling_qwen2_vl.py:313-326; trainable under --fix_vit,
 * timer1_trainer.py:272-280): dx (may be NULL) = d(layernorm)/dx; dw / db (fp32, caller-zeroed) += sum_t dy * xhat / dy. */
int t1_layernorm_bwd(const void* dy, const void* x, const void* w, void* dx, float* dw, float* db, int T, int D, float eps,
                     cudaStream_t stream);
/* cos/sin[t, i] of pos[axis_id[i], t] * inv_freq[i]: M-RoPE (Q25:595-608,659-665) and ViT 2-D RoPE (Q25:382-409). */
int t1_rope_table(const int* pos, int T, const int* axis_id, const float* inv_freq, int half, float* cos_t,
                  float* sin_t, cudaStream_t stream);
/* rotate_half rotation of n_heads heads at column col0 of x [T, ld] in place (inverse != 0: transpose rotation, used
 * for gradients); optional copy of the rotated heads to out (KV cache).  Q25:149-167, :627-669. */
int t1_rope_apply(void* x, long long ld, int col0, int n_heads, int head_dim, const float* cos_t, const float* sin_t,
                  int inverse, int T, void* out, long long ldo, cudaStream_t stream);
int t1_copy_cols(const void* src, long long lds, void* dst, long long ldd, int T, int width, cudaStream_t stream);
/* act = silu(gate) * up for gate_up [T, 2I] (Q25:611-624, :77-88) and its backward. */
int t1_swiglu(const void* gate_up, void* act, long long T, int I, cudaStream_t stream);
int t1_swiglu_bwd(const void* dact, const void* gate_up, void* dgate_up, long long T, int I, cudaStream_t stream);
/* mode 0: exact GELU (merger, Q25:139); mode 1: QuickGELU (Qwen2-VL ViT MLP). */
int t1_act(const void* x, void* y, long long n, int mode, cudaStream_t stream);
int t1_act_bwd(const void* dy, const void* x, void* dx, long long n, int mode, cudaStream_t stream);
/* dst[i] = src[idx[i]] / dst[idx[i]] (+)= src[i] over bf16 rows (idx < 0 skips): embedding lookup, masked_scatter of
 * video embeddings (Q25:1309-1315), window permutation (Q25:478-484,512-513), embedding-gradient scatter-add. */
int t1_gather_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n, int width,
                   cudaStream_t stream);
int t1_gather_rows_f32(const float* src, long long lds, const int* idx, void* dst, long long ldd, int n, int width,
                       cudaStream_t stream); /* fp32 pixel rows -> bf16, permuted (Q25:1149 cast fused) */
int t1_scatter_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n, int width,
                    int accumulate, cudaStream_t stream);
int t1_transpose_bf16(const void* in, long long ld, void* out, long long ldo, int R, int C, cudaStream_t stream);
int t1_colsum(const void* x, long long ld, int R, int C, float* out /* fp32, += */, cudaStream_t stream);
int t1_f32_to_bf16(const float* src, void* dst, long long n, int accumulate, cudaStream_t stream);
int t1_add_bf16(void* y, const void* x, long long n, cudaStream_t stream);
int t1_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream);

/* ---- video pre-processing (HF Qwen2-VL video processor, timer1_trainer.py:547-556) --------------------------- */
/* Rescale + normalise + patchify of resized frames [T, 3, H, W] (uint8 when in_u8 != 0, else fp32; values 0..255):
 * out_f32 [N, 1176] = pixel_values_videos bit-identical to HF models/qwen2_vl/video_processing_qwen2_vl.py:236-272
 * ((x - mean*255) / (std*255) in fp32, permute (0,1,4,7,5,8,3,2,6,9), odd T padded with the last frame); and / or
 * out_bf16 [N, 1176] with rows gathered through perm_rows (device int32 [N], nullable) — the ViT window order the
 * patch-embed GEMM reads.  mean255 / std255: HOST arrays of 3 floats = fp32(mean) * 255, fp32(std) * 255. */
int t1_video_patchify(const void* frames, int in_u8, int T, int H, int W, int patch, int tpatch, int merge,
                      const float* mean255, const float* std255, const int* perm_rows, float* out_f32, void* out_bf16,
                      cudaStream_t stream);

/* ---- GRPO: rewards, advantage, loss (warp/block-reduction tier) ------------------------------------------- */
/* IoU rewards in fp64, bit-identical to CPython: mode 0 = iou_timestamp_reward (main.py:145-181), mode 1 =
 * iou_timestamp_reward_v2 (main.py:184-231).  valid[i] == 0 (unparsable completion) -> 0.  union <= 0 -> 0 (the
 * reference leaves `iou` unbound there, main.py:165-168).  mode 2 = difficulty of the per-epoch sample filter,
 * 100 * compute_IoU with numpy float64 semantics incl. NaN / inf for an empty union
 * (src/vllm_inference/calc_difficulty.py:10-37). */
int t1_iou_reward(int mode, const double* pred_s, const double* pred_e, const unsigned char* valid, const double* gt_s,
                  const double* gt_e, const double* dur, double* out64, float* out32, int n, cudaStream_t stream);
/* Completion text on the device: replaces the device->host copy + tokenizer.batch_decode +
 * Python regexes between the scoring passes and the loss (timer1_trainer.py:642-698 -> main.py:122-239).
 * t1_detokenize: ids [n, C] int32 -> UTF-8 bytes [n, cap] + byte length [n], i.e. the byte string tokenizers' ByteLevel
 * decoder builds for batch_decode (timer1_trainer.py:642-644).  row_tokens [n] int32 (nullable: all C) = tokens of each
 * row to decode; tok_off [V + 1] int32 + tok_bytes = the vocabulary's byte strings; tok_skip [V] uint8 (nullable) marks
 * the special tokens skip_special_tokens=True drops; ids outside [0, V) produce nothing.  max_tok_bytes = longest entry of
 * the table; cap >= C * max_tok_bytes is required (no row can be truncated).  All pointers are device pointers. */
int t1_detokenize(const int* ids, int n, int C, const int* row_tokens, const int* tok_off, const unsigned char* tok_bytes,
                  const unsigned char* tok_skip, int V, int max_tok_bytes, unsigned char* out_bytes, int cap, int* out_len,
                  cudaStream_t stream);
/* t1_parse_answers: byte-exact restatement of the reference's regexes + float() on text [n, cap] / len [n]
 * (invalid UTF-8 behaves as the U+FFFD of decode(errors="replace"); `\d`, `\s`, strip() are Unicode-aware as in CPython;
 * float() is correctly rounded for any digit count).  mode 0: parse_timestamp_output (main.py:122-142: last
 * `<num> to|and <num>`, case-insensitive, inside the last <answer>..</answer>) and format_reward (main.py:234-239);
 * mode 1: extract_answer (evaluate.py:125-151); mode 2: mode 1, then the first two numbers when it finds nothing
 * (src/vllm_inference/calc_difficulty.py:40-60).  out_s / out_e [n] fp64, out_valid [n] uint8 feed t1_iou_reward;
 * out_format [n] uint8 and out_format_f32 [n] (the format_reward column) are nullable. */
int t1_parse_answers(const unsigned char* text, const int* len, int n, int cap, int mode, double* out_s, double* out_e,
                     unsigned char* out_valid, unsigned char* out_format, float* out_format_f32, cudaStream_t stream);
/* rewards = sum over functions; per-group mean, UNBIASED std; adv = (r - mean) / (std + eps)
 * (timer1_trainer.py:700-712).  rewards_per_func [n, nf] fp32; n % G == 0. */
int t1_group_advantage(const float* rewards_per_func, int n, int nf, int G, float eps, float* rewards, float* adv,
                       float* mean, float* std_, cudaStream_t stream);
/* k3 KL, PPO-clip or per-sequence GRPO loss, dL/dlogp and the logged metrics (timer1_trainer.py:635-639,713-777).
 * old_logp NULL = on-policy (ratio 1, as in the reference).  out_stats[8] = {loss, mean_kl, mean_entropy,
 * mean_completion_length, sum(mask), clip_low_frac, clip_high_frac, mean_ratio}. */
int t1_grpo_loss(const float* logp, const float* old_logp, const float* ref_logp, const float* entropy,
                 const float* adv, const int* mask, int rows, int C, float beta, float eps_lo, float eps_hi,
                 int use_grpo, float loss_scale, float* out_loss, float* out_dlogp, float* out_stats,
                 cudaStream_t stream);
/* log_softmax gather + entropy per row of fp32 logits (timer1_trainer.py:458-481). */
int t1_logprob_entropy(const float* logits, long long ld, int rows, int V, const int* target, float* out_logp,
                       float* out_entropy, float* out_lse, cudaStream_t stream);
/* lm_head fused with the log-softmax statistics -- timer1_trainer.py:452-481 (logits -> log_softmax -> gather -> exp/sum entropy,
 * per row in a Python loop) without the logits tensor: t1_gemm_lse_parts runs x = A B^T (B = lm_head [V, K]) on the tcgen05 GEMM and
 * keeps, per 256-column tile and row, {max, sum exp(x - max), sum exp(x - max) * x, x[target[row]] or -inf} in
 * parts[ceil(N / 256)][M][4] (fp32); t1_lse_merge folds the tiles into log p(target), entropy and log-sum-exp per row (any output
 * may be NULL).  Same quantities as t1_gemm_bf16_tn (fp32 out) + t1_logprob_entropy. */
int t1_gemm_lse_parts(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const int* target,
                      float* parts, cudaStream_t stream);
int t1_lse_merge(const float* parts, int n_tiles, int M, float* logp, float* entropy, float* lse, cudaStream_t stream);
/* The backward's recomputed lm_head GEMM with t1_logits_backward in its epilogue: dlogits[M, N] (bf16) = dlogp[row] *
 * (1[col == target[row]] - exp(x - lse[row])), x = A B^T read from the accumulator (autograd of timer1_trainer.py:458-481). */
int t1_gemm_dlogits(const void* A, long long lda, const void* B, long long ldb, void* dlogits, long long ldd, int M, int N, int K,
                    const int* target, const float* lse, const float* dlogp, cudaStream_t stream);
/* dlogits = dlogp * (onehot(target) - softmax) in bf16 (autograd of the above). */
int t1_logits_backward(const float* logits, long long ld, int rows, int V, const int* target, const float* lse,
                       const float* dlogp, void* dlogits, long long ldd, cudaStream_t stream);
/* completion mask up to and including the first EOS (timer1_trainer.py:581-590). */
int t1_eos_mask(const int* ids, int rows, int C, int eos_id, int* mask, int* lengths, cudaStream_t stream);

/* ---- optimizer (DeepSpeed AdamW + HF clip_grad_norm_, HF trainer.py:2493,1760) ----------------------------- */
int t1_sumsq_bf16(const void* g, long long n, float* out_accum, cudaStream_t stream);
/* Decoupled AdamW on a flat shard: fp32 master/m/v, bf16 grad in, bf16 param out; the clip coefficient
 * min(1, max_norm / (sqrt(*gnorm_sq) * grad_scale + 1e-6)) is applied in-kernel. n % 4 == 0. */
int t1_adamw(float* master, float* m, float* v, const void* grad, void* param, long long n, float lr, float beta1,
             float beta2, float eps, float wd, int step, const float* gnorm_sq, float max_norm, float grad_scale,
             cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TIMER1_H */
