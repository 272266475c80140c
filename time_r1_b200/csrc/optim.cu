// Optimizer kernels: global grad-norm (sum of squares) and sharded AdamW with fp32 master weights.
// Replaces DeepSpeed's (CPU-offloaded) AdamW + HF Trainer's clip_grad_norm_ that the reference runs after
// accelerator.backward (HF trainer.py:2493,1760; scripts/zero3_offload.json:24-31).  HBM-bound: 16 B read + 14 B
// written per parameter (fp32 master/m/v + bf16 grad in; fp32 master/m/v + bf16 param out).
#include "t1_common.cuh"
#include <stdlib.h>

namespace {

__global__ void __launch_bounds__(512)
sumsq_bf16_kernel(const bf16* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[33];
  float acc = 0.f;
  const long long n8 = n >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(g) + i);
    float2 t;
    t = unpack_bf16x2(u.x); acc += t.x * t.x + t.y * t.y;
    t = unpack_bf16x2(u.y); acc += t.x * t.x + t.y * t.y;
    t = unpack_bf16x2(u.z); acc += t.x * t.x + t.y * t.y;
    t = unpack_bf16x2(u.w); acc += t.x * t.x + t.y * t.y;
  }
  if (blockIdx.x == 0)
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = __bfloat162float(g[i]);
      acc += v * v;
    }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

// torch.optim.AdamW semantics (decoupled weight decay, bias correction), gradient pre-scaled by
// grad_scale * clip where clip = min(1, max_norm / (sqrt(*gnorm_sq * grad_scale^2) + 1e-6)) as in clip_grad_norm_.
// 128-thread CTAs, 92 registers, four independent chunks per thread per trip with every load issued before the first use.
// Built to test whether AdamW can run UNDER the next step's frozen ViT blocks (side stream): one such CTA fits in the registers
// a persistent tcgen05 GEMM CTA leaves free on an SM.  Measured (profiles/r02_adamw_coresident.txt): a single CTA per SM only
// sustains 2.0 TB/s (116 ms for the 7B shard; 2 per SM 3.7 TB/s, 4 per SM 5.9 TB/s = the HBM floor, 39 ms), so capping the
// grid at one CTA per SM made the step SLOWER (1407 -> 1419..1430 ms).  Default: 4 CTAs per SM, same speed as the previous
// 256-thread kernel; T1_ADAMW_CTAS_PER_SM overrides.
constexpr int ADAMW_UNROLL = 4;
__global__ void __launch_bounds__(128, 5)
adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v, const bf16* __restrict__ grad,
             bf16* __restrict__ param, long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1,
             float bc2, const float* __restrict__ gnorm_sq, float max_norm, float grad_scale) {
  float gs = grad_scale;
  if (gnorm_sq && max_norm > 0.f) {
    const float total = sqrtf(*gnorm_sq) * grad_scale;
    gs *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.0f - lr * wd;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * ADAMW_UNROLL) {
    float4 p[ADAMW_UNROLL], mm[ADAMW_UNROLL], vv[ADAMW_UNROLL];
    uint2 gu[ADAMW_UNROLL];
#pragma unroll
    for (int u = 0; u < ADAMW_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        p[u] = reinterpret_cast<const float4*>(master)[i];
        mm[u] = reinterpret_cast<const float4*>(m)[i];
        vv[u] = reinterpret_cast<const float4*>(v)[i];
        gu[u] = __ldg(reinterpret_cast<const uint2*>(grad) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < ADAMW_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) {
        const float2 g01 = unpack_bf16x2(gu[u].x), g23 = unpack_bf16x2(gu[u].y);
        float pa[4] = {p[u].x, p[u].y, p[u].z, p[u].w}, ma[4] = {mm[u].x, mm[u].y, mm[u].z, mm[u].w};
        float va[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
        const float ga[4] = {g01.x * gs, g01.y * gs, g23.x * gs, g23.y * gs};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pa[j] *= decay;
          ma[j] = beta1 * ma[j] + (1.0f - beta1) * ga[j];
          va[j] = beta2 * va[j] + (1.0f - beta2) * ga[j] * ga[j];
          const float denom = sqrtf(va[j]) * inv_sqrt_bc2 + eps;
          pa[j] -= step_size * (ma[j] / denom);
        }
        reinterpret_cast<float4*>(master)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
        uint2 o;
        o.x = pack_bf16x2(pa[0], pa[1]);
        o.y = pack_bf16x2(pa[2], pa[3]);
        reinterpret_cast<uint2*>(param)[i] = o;
      }
    }
  }
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}

}  // namespace

extern "C" int t1_sumsq_bf16(const void* g, long long n, float* out_accum, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(g && out_accum && ((uintptr_t)g & 15) == 0, "t1_sumsq_bf16: bad arguments");
  const long long want = (n / 8 + 511) / 512;
  const int grid = (int)(want < 8LL * t1_num_sms() ? (want > 0 ? want : 1) : 8LL * t1_num_sms());
  sumsq_bf16_kernel<<<grid, 512, 0, stream>>>((const bf16*)g, n, out_accum);
  T1_CUDA_LAUNCH_CHECK("t1_sumsq_bf16");
  return T1_OK;
}

extern "C" int t1_adamw(float* master, float* m, float* v, const void* grad, void* param, long long n, float lr,
                        float beta1, float beta2, float eps, float wd, int step, const float* gnorm_sq, float max_norm,
                        float grad_scale, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n % 4 == 0, "t1_adamw: n must be a multiple of 4 (pad the flat shard)");
  T1_CHECK_ARG(step >= 1 && master && m && v && grad && param, "t1_adamw: bad arguments");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  const long long want = (n / 4 + 128 * ADAMW_UNROLL - 1) / (128 * ADAMW_UNROLL);
  static const char* env = getenv("T1_ADAMW_CTAS_PER_SM");
  const long long per_sm = env && atoi(env) > 0 ? atoi(env) : 4;
  const int grid = (int)(want < per_sm * t1_num_sms() ? (want > 0 ? want : 1) : per_sm * t1_num_sms());
  adamw_kernel<<<grid, 128, 0, stream>>>(master, m, v, (const bf16*)grad, (bf16*)param, n, lr, beta1, beta2, eps, wd, bc1,
                                         bc2, gnorm_sq, max_norm, grad_scale);
  T1_CUDA_LAUNCH_CHECK("t1_adamw");
  return T1_OK;
}

extern "C" int t1_bf16_to_f32(const void* src, float* dst, long long n, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  const long long want = (n + 255) / 256;
  const int grid = (int)(want < 16LL * t1_num_sms() ? want : 16LL * t1_num_sms());
  bf16_to_f32_kernel<<<grid, 256, 0, stream>>>((const bf16*)src, dst, n);
  T1_CUDA_LAUNCH_CHECK("t1_bf16_to_f32");
  return T1_OK;
}
