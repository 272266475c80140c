// GRPO reward / advantage / loss / log-prob kernels (warp- and block-reduction tier).
//
// Reference semantics restated on the device:
//   IoU rewards           main.py:145-231  (iou_timestamp_reward, iou_timestamp_reward_v2) — fp64 like CPython
//   group advantage       src/time_r1/rl/timer1_trainer.py:700-712 (sum over reward fns, mean, unbiased std, +1e-4)
//   KL / clip / loss      timer1_trainer.py:635-639, 713-737
//   metrics               timer1_trainer.py:740-777
//   log-prob + entropy    timer1_trainer.py:458-481 (log_softmax, gather, -sum p log p)
#include "t1_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// IoU reward: one thread per completion; fp64 with explicitly rounded ops (no FMA contraction) so that the
// result is bit-identical to the CPython float arithmetic of the reference.
// ---------------------------------------------------------------------------------------------
__global__ void iou_reward_kernel(int mode, const double* __restrict__ ps, const double* __restrict__ pe,
                                  const unsigned char* __restrict__ valid, const double* __restrict__ gs,
                                  const double* __restrict__ ge, const double* __restrict__ dur,
                                  double* __restrict__ out64, float* __restrict__ out32, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double reward = 0.0;
  if (valid[i] && mode == 2) {
    // difficulty of the per-epoch sample filter: 100 * compute_IoU (src/vllm_inference/calc_difficulty.py:10-37), numpy
    // float64 semantics: the union is clamped at 0 and 0/0 = NaN, x/0 = inf propagate (dropped later by
    // get_difficulty_safe, src/utils/process_data.py:11-24)
    const double s = gs[i], e = ge[i], from = ps[i], to = pe[i];
    const double inter = fmax(0.0, __dsub_rn(fmin(to, e), fmax(from, s)));
    const double uni = fmax(0.0, __dsub_rn(fmax(to, e), fmin(from, s)));
    reward = __dmul_rn(__ddiv_rn(inter, uni), 100.0);
  } else if (valid[i]) {
    const double s = gs[i], e = ge[i], from = ps[i], to = pe[i];
    const double inter = fmax(0.0, __dsub_rn(fmin(to, e), fmax(from, s)));
    const double uni = __dsub_rn(fmax(to, e), fmin(from, s));
    // reference leaves `iou` unbound when union <= 0 (main.py:165-168); defined as 0 here
    const double iou = uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
    if (mode == 0) {
      reward = iou;
    } else {
      const double d = dur[i];
      const double gsn = __ddiv_rn(s, d), gen = __ddiv_rn(e, d);
      const double psn = __ddiv_rn(from, d), pen = __ddiv_rn(to, d);
      const double f1 = __dsub_rn(1.0, fabs(__dsub_rn(gsn, psn)));
      const double f2 = __dsub_rn(1.0, fabs(__dsub_rn(gen, pen)));
      reward = __dmul_rn(__dmul_rn(iou, f1), f2);
    }
  }
  if (out64) out64[i] = reward;
  if (out32) out32[i] = (float)reward;
}

// ---------------------------------------------------------------------------------------------
// Group advantage: one warp per prompt-group (G members), shuffle reductions.
// ---------------------------------------------------------------------------------------------
__global__ void group_advantage_kernel(const float* __restrict__ rewards_per_func, int nf, int G, int n_groups,
                                       float eps, float* __restrict__ rewards, float* __restrict__ adv,
                                       float* __restrict__ mean_out, float* __restrict__ std_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_groups) return;
  const int base = warp * G;
  float local_sum = 0.f;
  for (int j = lane; j < G; j += 32) {
    float r = 0.f;
    for (int f = 0; f < nf; ++f) r += rewards_per_func[(size_t)(base + j) * nf + f];
    if (rewards) rewards[base + j] = r;
    local_sum += r;
  }
  const float mean = warp_sum(local_sum) / (float)G;
  float local_sq = 0.f;
  for (int j = lane; j < G; j += 32) {
    float r = 0.f;
    for (int f = 0; f < nf; ++f) r += rewards_per_func[(size_t)(base + j) * nf + f];
    const float d = r - mean;
    local_sq += d * d;
  }
  // torch.std default: unbiased (ddof = 1); G == 1 gives NaN in the reference, mirrored here.
  const float var = warp_sum(local_sq) / (float)(G - 1);
  const float sd = sqrtf(var);
  for (int j = lane; j < G; j += 32) {
    float r = 0.f;
    for (int f = 0; f < nf; ++f) r += rewards_per_func[(size_t)(base + j) * nf + f];
    adv[base + j] = (r - mean) / (sd + eps);
  }
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (std_out) std_out[warp] = sd;
  }
}

// ---------------------------------------------------------------------------------------------
// GRPO loss + dL/dlogp + metrics. One CTA (rows*C is a few thousand elements).
// stats[0]=loss  [1]=mean_kl (per-seq masked mean, then mean)  [2]=mean entropy (same reduction)
// stats[3]=mean completion length  [4]=sum(mask)  [5]=clip-low fraction  [6]=clip-high fraction  [7]=mean ratio
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
grpo_loss_kernel(const float* __restrict__ logp, const float* __restrict__ old_logp,
                 const float* __restrict__ ref_logp, const float* __restrict__ entropy,
                 const float* __restrict__ adv, const int* __restrict__ mask, int rows, int C, float beta,
                 float eps_lo, float eps_hi, int use_grpo, float loss_scale, float* __restrict__ out_loss,
                 float* __restrict__ dlogp, float* __restrict__ stats) {
  __shared__ float red[33];
  extern __shared__ float row_cnt[];  // [rows]
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  // pass 1: per-row mask counts (warp per row)
  float tot_cnt_local = 0.f;
  for (int r = wid; r < rows; r += nw) {
    float c = 0.f;
    for (int t = lane; t < C; t += 32) c += (float)mask[(size_t)r * C + t];
    c = warp_sum(c);
    if (lane == 0) row_cnt[r] = c;
    tot_cnt_local += (lane == 0) ? c : 0.f;
  }
  const float total_cnt = block_sum(tot_cnt_local, red);
  // pass 2: per-token loss, gradients, metrics
  float loss_acc = 0.f, kl_acc = 0.f, ent_acc = 0.f, lo_acc = 0.f, hi_acc = 0.f, ratio_acc = 0.f;
  const int n = rows * C;
  for (int i = tid; i < n; i += blockDim.x) {
    const int r = i / C;
    const float m = (float)mask[i];
    const float lp = logp[i];
    const float A = adv[r];
    const float ratio = old_logp ? expf(lp - old_logp[i]) : 1.0f;  // exp(lp - lp.detach()) == 1 on-policy
    float kl = 0.f, dkl = 0.f;
    if (ref_logp && beta != 0.f) {
      const float d = ref_logp[i] - lp;
      const float ed = expf(d);
      kl = ed - d - 1.0f;
      dkl = 1.0f - ed;  // d kl / d lp
    }
    float tok_loss, dtok;
    if (use_grpo) {
      tok_loss = -(ratio * A - beta * kl);
      dtok = -(ratio * A) + beta * dkl;
    } else {
      const float clipped = fminf(fmaxf(ratio, 1.0f - eps_lo), 1.0f + eps_hi);
      const float l1 = ratio * A, l2 = clipped * A;
      const bool use_unclipped = l1 <= l2;  // torch.min picks l1 when equal (gradient flows through l1)
      tok_loss = -(use_unclipped ? l1 : l2) + beta * kl;
      const bool clip_active = (ratio < 1.0f - eps_lo) || (ratio > 1.0f + eps_hi);
      const float dmin = use_unclipped ? ratio * A : (clip_active ? 0.f : ratio * A);
      dtok = -dmin + beta * dkl;
      lo_acc += m * ((ratio < 1.0f - eps_lo && A < 0.f) ? 1.f : 0.f);
      hi_acc += m * ((ratio > 1.0f + eps_hi && A > 0.f) ? 1.f : 0.f);
    }
    const float rc = row_cnt[r];
    const float w = use_grpo ? (m / rc) / (float)rows : m / total_cnt;  // reference divides by 0 -> NaN if empty
    loss_acc += tok_loss * w;
    if (dlogp) dlogp[i] = dtok * w * loss_scale;
    const float wseq = (m / rc) / (float)rows;
    kl_acc += kl * wseq;
    if (entropy) ent_acc += entropy[i] * (m / fmaxf(rc, 1.0f)) / (float)rows;
    ratio_acc += ratio * m;
  }
  const float loss = block_sum(loss_acc, red);
  const float mkl = block_sum(kl_acc, red);
  const float ment = block_sum(ent_acc, red);
  const float lo = block_sum(lo_acc, red);
  const float hi = block_sum(hi_acc, red);
  const float rs = block_sum(ratio_acc, red);
  if (tid == 0) {
    if (out_loss) out_loss[0] = loss * loss_scale;
    if (stats) {
      stats[0] = loss;
      stats[1] = mkl;
      stats[2] = ment;
      stats[3] = total_cnt / (float)rows;
      stats[4] = total_cnt;
      stats[5] = lo / total_cnt;
      stats[6] = hi / total_cnt;
      stats[7] = rs / total_cnt;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Per-row log-softmax gather + entropy from fp32 logits: one CTA per row, single online pass.
//   lse = m + log(sum exp(x-m));  logp = x[target] - lse;  H = lse - sum(softmax * x)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
logprob_entropy_kernel(const float* __restrict__ logits, long long ld, int V, const int* __restrict__ target,
                       float* __restrict__ out_logp, float* __restrict__ out_entropy, float* __restrict__ out_lse) {
  __shared__ float red[33];
  const int r = blockIdx.x;
  const float* x = logits + (size_t)r * ld;
  float m = -INFINITY, s = 0.f, t = 0.f;  // running max, sum exp(x-m), sum exp(x-m)*x
  const int V4 = V >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int i = threadIdx.x; i < V4; i += blockDim.x) {
    const float4 v = x4[i];
    const float vm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (vm > m) {
      const float sc = expf(m - vm);
      s *= sc;
      t *= sc;
      m = vm;
    }
    const float e0 = expf(v.x - m), e1 = expf(v.y - m), e2 = expf(v.z - m), e3 = expf(v.w - m);
    s += (e0 + e1) + (e2 + e3);
    t += (e0 * v.x + e1 * v.y) + (e2 * v.z + e3 * v.w);
  }
  for (int i = (V4 << 2) + threadIdx.x; i < V; i += blockDim.x) {
    const float v = x[i];
    if (v > m) {
      const float sc = expf(m - v);
      s *= sc;
      t *= sc;
      m = v;
    }
    const float e = expf(v - m);
    s += e;
    t += e * v;
  }
  const float gm = block_max(m, red);
  const float sc = (m == -INFINITY) ? 0.f : expf(m - gm);
  const float gs = block_sum(s * sc, red);
  const float gt = block_sum(t * sc, red);
  if (threadIdx.x == 0) {
    const float lse = gm + logf(gs);
    if (out_lse) out_lse[r] = lse;
    if (out_logp) out_logp[r] = x[target[r]] - lse;
    if (out_entropy) out_entropy[r] = lse - gt / gs;
  }
}

// Fold the per-tile softmax statistics of t1_gemm_lse_parts (parts[n_tiles][M][4] = max, sum exp(x - max), sum exp(x - max) * x,
// target logit or -inf) into log p(target), entropy and log-sum-exp per row: one warp per row, same formulas as above.
__global__ void lse_merge_kernel(const float4* __restrict__ parts, int n_tiles, int M, float* __restrict__ out_logp,
                                 float* __restrict__ out_entropy, float* __restrict__ out_lse) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  float m = -INFINITY, s = 0.f, t = 0.f, tl = -INFINITY;
  for (int i = lane; i < n_tiles; i += 32) {
    const float4 p = parts[(size_t)i * M + r];
    if (p.x > m) {
      const float sc = expf(m - p.x);
      s *= sc;
      t *= sc;
      m = p.x;
    }
    const float e = (p.x == -INFINITY) ? 0.f : expf(p.x - m);
    s += p.y * e;
    t += p.z * e;
    tl = fmaxf(tl, p.w);
  }
  float gm = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
  const float sc = (m == -INFINITY) ? 0.f : expf(m - gm);
  s *= sc;
  t *= sc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    t += __shfl_xor_sync(0xffffffffu, t, o);
    tl = fmaxf(tl, __shfl_xor_sync(0xffffffffu, tl, o));
  }
  if (lane == 0) {
    const float lse = gm + logf(s);
    if (out_lse) out_lse[r] = lse;
    if (out_logp) out_logp[r] = tl - lse;
    if (out_entropy) out_entropy[r] = lse - t / s;
  }
}

// dlogits[r, v] = dlogp[r] * (1[v == target[r]] - exp(x - lse[r]))   (bf16 out, fp32 in)
__global__ void __launch_bounds__(1024, 1)
logits_backward_kernel(const float* __restrict__ logits, long long ld, int V, const int* __restrict__ target,
                       const float* __restrict__ lse, const float* __restrict__ dlogp, bf16* __restrict__ dlogits,
                       long long ldd) {
  const int r = blockIdx.x;
  const float* x = logits + (size_t)r * ld;
  bf16* d = dlogits + (size_t)r * ldd;
  const float g = dlogp[r], l = lse[r];
  const int tgt = target[r];
  const int V8 = V >> 3;
  for (int i = threadIdx.x; i < V8; i += blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = g * (((i * 8 + j) == tgt ? 1.f : 0.f) - expf(v[j] - l));
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(d)[i] = o;
  }
  for (int i = (V8 << 3) + threadIdx.x; i < V; i += blockDim.x)
    d[i] = __float2bfloat16(g * ((i == tgt ? 1.f : 0.f) - expf(x[i] - l)));
}

// First-EOS completion mask (timer1_trainer.py:581-590): mask[r,t] = t <= first index with ids == eos (inclusive).
__global__ void eos_mask_kernel(const int* __restrict__ ids, int rows, int C, int eos, int* __restrict__ mask,
                                int* __restrict__ lengths) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int first = C;
  for (int t = lane; t < C; t += 32)
    if (ids[(size_t)warp * C + t] == eos) {
      first = t;
      break;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  for (int t = lane; t < C; t += 32) mask[(size_t)warp * C + t] = (t <= first) ? 1 : 0;
  if (lane == 0 && lengths) lengths[warp] = min(first + 1, C);
}

}  // namespace

extern "C" int t1_iou_reward(int mode, const double* pred_s, const double* pred_e, const unsigned char* valid,
                             const double* gt_s, const double* gt_e, const double* dur, double* out64, float* out32,
                             int n, cudaStream_t stream) {

  T1_CHECK_ARG(mode >= 0 && mode <= 2, "t1_iou_reward: mode %d not in {0: iou, 1: iou_v2, 2: difficulty}", mode);
  T1_CHECK_ARG(mode != 1 || dur != nullptr, "t1_iou_reward: iou_v2 needs durations");
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n > 0 && pred_s && pred_e && valid && gt_s && gt_e, "t1_iou_reward: bad arguments");
  iou_reward_kernel<<<(n + 255) / 256, 256, 0, stream>>>(mode, pred_s, pred_e, valid, gt_s, gt_e, dur, out64, out32, n);
  T1_CUDA_LAUNCH_CHECK("t1_iou_reward");
  return T1_OK;
}

extern "C" int t1_group_advantage(const float* rewards_per_func, int n, int nf, int G, float eps, float* rewards,
                                  float* adv, float* mean, float* std_, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n > 0 && nf > 0 && G > 0 && n % G == 0, "t1_group_advantage: n=%d must be a positive multiple of G=%d", n, G);
  T1_CHECK_ARG(rewards_per_func && adv, "t1_group_advantage: null pointer");
  const int groups = n / G;
  const int threads = 128;
  group_advantage_kernel<<<(groups * 32 + threads - 1) / threads, threads, 0, stream>>>(rewards_per_func, nf, G, groups,
                                                                                      eps, rewards, adv, mean, std_);
  T1_CUDA_LAUNCH_CHECK("t1_group_advantage");
  return T1_OK;
}

extern "C" int t1_grpo_loss(const float* logp, const float* old_logp, const float* ref_logp, const float* entropy,
                            const float* adv, const int* mask, int rows, int C, float beta, float eps_lo, float eps_hi,
                            int use_grpo, float loss_scale, float* out_loss, float* out_dlogp, float* out_stats,
                            cudaStream_t stream) {
  T1_CHECK_ARG(rows > 0 && C > 0 && logp && adv && mask, "t1_grpo_loss: bad arguments");
  T1_CHECK_ARG(rows <= 8192, "t1_grpo_loss: rows=%d exceeds the single-CTA limit of 8192", rows);
  grpo_loss_kernel<<<1, 1024, rows * sizeof(float), stream>>>(logp, old_logp, ref_logp, entropy, adv, mask, rows, C,
                                                              beta, eps_lo, eps_hi, use_grpo, loss_scale, out_loss,
                                                              out_dlogp, out_stats);
  T1_CUDA_LAUNCH_CHECK("t1_grpo_loss");
  return T1_OK;
}

extern "C" int t1_logprob_entropy(const float* logits, long long ld, int rows, int V, const int* target, float* out_logp,
                                  float* out_entropy, float* out_lse, cudaStream_t stream) {
  if (rows == 0) return T1_OK;
  T1_CHECK_ARG(rows > 0 && V > 0 && logits && target, "t1_logprob_entropy: bad arguments");
  T1_CHECK_ARG(ld % 4 == 0 && ((uintptr_t)logits & 15) == 0, "t1_logprob_entropy: logits rows must be 16-byte aligned");
  logprob_entropy_kernel<<<rows, 1024, 0, stream>>>(logits, ld, V, target, out_logp, out_entropy, out_lse);
  T1_CUDA_LAUNCH_CHECK("t1_logprob_entropy");
  return T1_OK;
}

extern "C" int t1_lse_merge(const float* parts, int n_tiles, int M, float* logp, float* entropy, float* lse, cudaStream_t stream) {
  if (M == 0) return T1_OK;
  T1_CHECK_ARG(parts && n_tiles > 0 && M > 0 && ((uintptr_t)parts & 15) == 0, "t1_lse_merge: bad arguments");
  lse_merge_kernel<<<(M * 32 + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float4*>(parts), n_tiles, M, logp, entropy, lse);
  T1_CUDA_LAUNCH_CHECK("t1_lse_merge");
  return T1_OK;
}

extern "C" int t1_logits_backward(const float* logits, long long ld, int rows, int V, const int* target,
                                  const float* lse, const float* dlogp, void* dlogits, long long ldd,
                                  cudaStream_t stream) {
  if (rows == 0) return T1_OK;
  T1_CHECK_ARG(rows > 0 && V > 0 && logits && target && lse && dlogp && dlogits, "t1_logits_backward: bad arguments");
  T1_CHECK_ARG(ld % 4 == 0 && ldd % 8 == 0, "t1_logits_backward: unaligned leading dimensions");
  logits_backward_kernel<<<rows, 1024, 0, stream>>>(logits, ld, V, target, lse, dlogp, (bf16*)dlogits, ldd);
  T1_CUDA_LAUNCH_CHECK("t1_logits_backward");
  return T1_OK;
}

extern "C" int t1_eos_mask(const int* ids, int rows, int C, int eos_id, int* mask, int* lengths, cudaStream_t stream) {
  if (rows == 0) return T1_OK;
  T1_CHECK_ARG(rows > 0 && C > 0 && ids && mask, "t1_eos_mask: bad arguments");
  eos_mask_kernel<<<(rows * 32 + 127) / 128, 128, 0, stream>>>(ids, rows, C, eos_id, mask, lengths);
  T1_CUDA_LAUNCH_CHECK("t1_eos_mask");
  return T1_OK;
}
