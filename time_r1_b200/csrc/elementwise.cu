// HBM-bound token-wise kernels of the Qwen2-VL / Qwen2.5-VL stack: RMSNorm / LayerNorm (+ fused residual),
// RoPE tables and rotation (M-RoPE for the LLM, 2-D RoPE for the ViT), SwiGLU / GELU / QuickGELU, embedding
// gather / scatter, row permutations, transposes and column sums — forward and backward.
// Reference semantics: HF modeling_qwen2_5_vl.py (RMSNorm :66-71, VisionMlp :77-88, apply_rotary_pos_emb_vision
// :149-167, PatchMerger :133-146, apply_multimodal_rotary_pos_emb :627-669, Qwen2MLP :611-624) which
// src/time_r1/rl/timer1_trainer.py:452-457,569-573 executes.  All math in fp32, bf16 storage, 16-byte vector
// accesses, one rounding at the end.
#include "t1_common.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ void load8(const bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------
// RMSNorm forward (optionally x += delta first, written back): y = w * (x * rsqrt(mean(x^2) + eps))
// One CTA per token row; D % 8 == 0.
// ---------------------------------------------------------------------------------------------
template <bool kLayerNorm>
__global__ void __launch_bounds__(256)
norm_fwd_kernel(const bf16* x, const bf16* __restrict__ delta, bf16* x_out, const bf16* __restrict__ w,
                const bf16* __restrict__ b, bf16* __restrict__ y, int D, float eps) {
  __shared__ float red[33];
  pdl_trigger();   // no-ops unless launched with the programmatic-serialization attribute (decode chain)
  pdl_wait();
  const size_t row = blockIdx.x;
  const bf16* xr = x + row * D;
  const int nv = D >> 3;
  float sum = 0.f, sq = 0.f;
  // D <= 8 * 256 * 4 handled with up to 4 vectors per thread kept in registers
  float v[4][8];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < nv) {
      load8(xr + i * 8, v[it]);
      if (delta) {
        float d[8];
        load8(delta + row * D + i * 8, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[it][j] += d[j];
        store8(x_out + row * D + i * 8, v[it]);
        // round in registers too, so the norm sees exactly what later consumers of x will read
#pragma unroll
        for (int j = 0; j < 8; ++j) v[it][j] = __bfloat162float(__float2bfloat16(v[it][j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum += v[it][j];
        sq += v[it][j] * v[it][j];
      }
    }
  }
  float mean = 0.f, rstd;
  if (kLayerNorm) {
    mean = block_sum(sum, red) / (float)D;
    float var = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < nv)
#pragma unroll
        for (int j = 0; j < 8; ++j) var += (v[it][j] - mean) * (v[it][j] - mean);
    }
    rstd = rsqrtf(block_sum(var, red) / (float)D + eps);
  } else {
    // one barrier instead of block_sum's three: the CTA owns one row, so `red` is written once
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];       // blockDim.x == 256
    rstd = rsqrtf(t / (float)D + eps);
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < nv) {
      float wv[8], o[8];
      load8(w + i * 8, wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mean) * rstd * wv[j];
      if (kLayerNorm && b) {
        float bv[8];
        load8(b + i * 8, bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += bv[j];
      }
      store8(y + row * D + i * 8, o);
    }
  }
}

// RMSNorm backward. Each CTA walks rows r = blockIdx.x, +gridDim.x, ...; keeps fp32 partial dw in registers and
// flushes with one atomicAdd per column at the end.  dx_out = (dres ? dres : 0) + dx.
// HBM-bound (3 reads + 1 write of [T, D] bf16).  A row is a serial chain load -> block reduction -> store, so the rows in
// flight per SM set the bandwidth: operands stay packed (bf16) in registers, NIT is the compile-time number of 8-wide
// vectors per thread (D = 3584 needs 2, not the 4 of the D <= 8192 worst case), two CTAs share an SM, each with the next
// row's x / dy already loading while the current row reduces, and the two row sums (x^2, g*w*x) share ONE barrier.
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// ---------------------------------------------------------------------------------------------
// Warp-per-row RMSNorm / LayerNorm forward for the big token batches (prefill, scoring, backward recompute, ViT).
// The one-CTA-per-row kernel above launches T tiny CTAs (a [16384, 1280] ViT norm = 16384 CTAs of 160 busy threads and a
// block barrier each): measured ~97 us per launch in the step against a 26 us byte floor.  Here a WARP owns a row — the row
// (and its delta) sits in registers as packed 16-byte vectors, the reductions are shuffles, no shared memory, no barrier —
// and a persistent grid of 8-warp CTAs strides over the rows, so every SM keeps dozens of independent 16-byte loads in flight.
// NV = 16-byte vectors per lane (D <= 256 * NV).  Same arithmetic and roundings as norm_fwd_kernel.
// ---------------------------------------------------------------------------------------------
template <bool kLayerNorm, int NV>
__global__ void __launch_bounds__(256)
norm_fwd_rows_kernel(const bf16* x, const bf16* __restrict__ delta, bf16* x_out, const bf16* __restrict__ w,
                     const bf16* __restrict__ b, bf16* __restrict__ y, int T, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 3;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < T; row += warps) {
    const size_t base = (size_t)row * D;
    uint4 xq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      xq[i] = v < nv ? *reinterpret_cast<const uint4*>(x + base + v * 8) : zero4;
    }
    if (delta) {
      uint4 dq[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        dq[i] = v < nv ? *reinterpret_cast<const uint4*>(delta + base + v * 8) : zero4;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        float a[8], d[8];
        unpack8(xq[i], a);
        unpack8(dq[i], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += d[j];
        xq[i].x = pack_bf16x2(a[0], a[1]);        // the norm sees exactly the bf16 values later consumers of x will read
        xq[i].y = pack_bf16x2(a[2], a[3]);
        xq[i].z = pack_bf16x2(a[4], a[5]);
        xq[i].w = pack_bf16x2(a[6], a[7]);
        if (v < nv) *reinterpret_cast<uint4*>(x_out + base + v * 8) = xq[i];
      }
    }
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a[8];
      unpack8(xq[i], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sum += a[j];
        sq += a[j] * a[j];
      }
    }
    float mean = 0.f, rstd;
    if (kLayerNorm) {
      mean = warp_sum(sum) / (float)D;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (lane + i * 32 < nv) {
          float a[8];
          unpack8(xq[i], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) var += (a[j] - mean) * (a[j] - mean);
        }
      }
      rstd = rsqrtf(warp_sum(var) / (float)D + eps);
    } else {
      rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        float a[8], wv[8], o[8];
        unpack8(xq[i], a);
        load8(w + v * 8, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (a[j] - mean) * rstd * wv[j];
        if (kLayerNorm && b) {
          float bv[8];
          load8(b + v * 8, bv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += bv[j];
        }
        store8(y + base + v * 8, o);
      }
    }
  }
}

template <bool kLayerNorm>
static bool launch_norm_rows(const void* x, const void* delta, void* x_out, const void* w, const void* b, void* y, int T, int D,
                             float eps, cudaStream_t stream) {
  const int nvl = (D / 8 + 31) / 32;            // vectors per lane
  static const char* env = getenv("T1_NORM_ROWS");
  if (env && env[0] == '0') return false;       // A/B switch: the one-CTA-per-row kernel
  // measured on B200 (profiles/r02_norm_ab.txt): [16384, 1280] + delta 34.8 us here vs 55.3 us for the CTA-per-row kernel,
  // but [5888, 3584] + delta 55.3 vs 43.0 us (14 vectors per lane cost 200 registers: one CTA per SM) -> narrow rows only
  if (T <= 64 || nvl > 5) return false;         // decode-sized batches keep the PDL-aware kernel
  const int want = (T + 7) / 8;
  const int cap = 8 * t1_num_sms();
  const int grid = want < cap ? want : cap;
#define T1_NORM_ROWS(NVV)                                                                                                    \
  norm_fwd_rows_kernel<kLayerNorm, NVV><<<grid, 256, 0, stream>>>((const bf16*)x, (const bf16*)delta, (bf16*)x_out, (const bf16*)w, \
                                                                  (const bf16*)b, (bf16*)y, T, D, eps)
  if (nvl <= 2) T1_NORM_ROWS(2);
  else T1_NORM_ROWS(5);
#undef T1_NORM_ROWS
  return true;
}

template <int NIT>
__global__ void __launch_bounds__(256, NIT <= 2 ? 2 : 1)
rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                   const bf16* __restrict__ dres, bf16* __restrict__ dx, float* __restrict__ dw, int T, int D,
                   float eps) {
  __shared__ float red[2][2][8];               // [row parity][sum kind][warp]
  const int nv = D >> 3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  float dwacc[NIT][8];
  uint4 wq[NIT], xq[NIT], gq[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int i = threadIdx.x + it * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[it][j] = 0.f;
    wq[it] = i < nv ? *reinterpret_cast<const uint4*>(w + i * 8) : zero4;
  }
  size_t row = blockIdx.x;
  if (row < (size_t)T) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      xq[it] = i < nv ? *reinterpret_cast<const uint4*>(x + row * D + i * 8) : zero4;
      gq[it] = i < nv ? *reinterpret_cast<const uint4*>(dy + row * D + i * 8) : zero4;
    }
  }
  for (int n = 0; row < (size_t)T; row += gridDim.x, ++n) {
    // next row's operands (and this row's residual gradient) start loading before the reduction below
    const size_t nrow = row + gridDim.x;
    uint4 xn[NIT], gn[NIT], rq[NIT];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      const bool on = i < nv;
      rq[it] = (dres && on) ? *reinterpret_cast<const uint4*>(dres + row * D + i * 8) : zero4;
      xn[it] = (on && nrow < (size_t)T) ? *reinterpret_cast<const uint4*>(x + nrow * D + i * 8) : zero4;
      gn[it] = (on && nrow < (size_t)T) ? *reinterpret_cast<const uint4*>(dy + nrow * D + i * 8) : zero4;
    }
    float sq = 0.f, dot = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      float xv[8], gv[8], wv[8];
      unpack8(xq[it], xv);
      unpack8(gq[it], gv);
      unpack8(wq[it], wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {             // padding vectors are zero: they add nothing
        sq += xv[j] * xv[j];
        dot += gv[j] * wv[j] * xv[j];
      }
    }
    sq = warp_sum(sq);
    dot = warp_sum(dot);
    float(*rb)[8] = red[n & 1];                  // parity buffers: the barrier of row n+1 orders reuse at row n+2
    if (lane == 0) {
      rb[0][wid] = sq;
      rb[1][wid] = dot;
    }
    __syncthreads();
    sq = 0.f;
    dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sq += rb[0][k];
      dot += rb[1][k];
    }
    const float rstd = rsqrtf(sq / (float)D + eps);
    const float c = dot * rstd * rstd / (float)D;  // mean(g*w*xhat) * rstd, with xhat = x*rstd
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      float xv[8], gv[8], wv[8], o[8];
      unpack8(xq[it], xv);
      unpack8(gq[it], gv);
      unpack8(wq[it], wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xhat = xv[j] * rstd;
        dwacc[it][j] += gv[j] * xhat;
        o[j] = rstd * (gv[j] * wv[j]) - xv[j] * rstd * c;
      }
      if (dres) {
        float r[8];
        unpack8(rq[it], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      if (i < nv) store8(dx + row * D + i * 8, o);
      xq[it] = xn[it];
      gq[it] = gn[it];
    }
  }
  if (dw) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < nv)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + i * 8 + j, dwacc[it][j]);
    }
  }
}

// LayerNorm backward (Qwen2-VL PatchMerger ln_q, HF modeling_qwen2_vl.py:313-326): with xhat = (x - mean) * rstd and
// gw = dy * w,  dx = rstd * (gw - mean(gw) - xhat * mean(gw * xhat)),  dw += dy * xhat,  db += dy.
// Same walk as the RMSNorm backward (CTA strides over rows, fp32 partial dw / db in registers, one atomicAdd per column at
// the end); the merger's [N, 1280] is a small fraction of a step, so rows are processed one at a time.
template <int NIT>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                     bf16* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int T, int D, float eps) {
  __shared__ float red[33];
  const int nv = D >> 3;
  float dwacc[NIT][8], dbacc[NIT][8], wv[NIT][8];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int i = threadIdx.x + it * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j) dwacc[it][j] = dbacc[it][j] = wv[it][j] = 0.f;
    if (i < nv) load8(w + i * 8, wv[it]);
  }
  for (size_t row = blockIdx.x; row < (size_t)T; row += gridDim.x) {
    float xv[NIT][8], gv[NIT][8];
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[it][j] = gv[it][j] = 0.f;
      if (i < nv) {
        load8(x + row * D + i * 8, xv[it]);
        load8(dy + row * D + i * 8, gv[it]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += xv[it][j];
    }
    const float mean = block_sum(sum, red) / (float)D;
    float var = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < nv)
#pragma unroll
        for (int j = 0; j < 8; ++j) var += (xv[it][j] - mean) * (xv[it][j] - mean);
    }
    const float rstd = rsqrtf(block_sum(var, red) / (float)D + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < nv)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (xv[it][j] - mean) * rstd, gw = gv[it][j] * wv[it][j];
          s1 += gw;
          s2 += gw * xhat;
        }
    }
    const float m1 = block_sum(s1, red) / (float)D;
    const float m2 = block_sum(s2, red) / (float)D;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * 256;
      if (i < nv) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (xv[it][j] - mean) * rstd;
          dwacc[it][j] += gv[it][j] * xhat;
          dbacc[it][j] += gv[it][j];
          o[j] = rstd * (gv[it][j] * wv[it][j] - m1 - xhat * m2);
        }
        if (dx) store8(dx + row * D + i * 8, o);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int i = threadIdx.x + it * 256;
    if (i < nv)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (dw) atomicAdd(dw + i * 8 + j, dwacc[it][j]);
        if (db) atomicAdd(db + i * 8 + j, dbacc[it][j]);
      }
  }
}

// ---------------------------------------------------------------------------------------------
// RoPE: table build + rotation.  Channel i of the half-dim uses position axis axis_id[i] and inv_freq[i].
//   table[t, i] = cos/sin(pos[axis_id[i], t] * inv_freq[i])
// Rotation (rotate_half convention): x'[i] = x[i] c - x[i+h] s ; x'[i+h] = x[i+h] c + x[i] s   (sign=-1: inverse)
// ---------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const int* __restrict__ pos, int T, const int* __restrict__ axis_id,
                                  const float* __restrict__ inv_freq, int half, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * half) return;
  const int t = idx / half, i = idx - t * half;
  const float ang = (float)pos[(size_t)axis_id[i] * T + t] * inv_freq[i];
  float s, c;
  sincosf(ang, &s, &c);
  cos_t[idx] = c;
  sin_t[idx] = s;
}

// x: [T rows, ld] bf16; rotates n_heads heads of head_dim starting at column col0. One thread per 8 channels of
// the half dimension. Optional copy-out of the rotated heads to out (row stride ldo) for the KV cache.
__global__ void rope_apply_kernel(bf16* __restrict__ x, long long ld, int col0, int n_heads, int head_dim,
                                  const float* __restrict__ cos_t, const float* __restrict__ sin_t, float sign, int T,
                                  bf16* __restrict__ out, long long ldo) {
  const int half = head_dim >> 1;
  const int vec_per_head = half >> 3;
  const long long total = (long long)T * n_heads * vec_per_head;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx % vec_per_head);
  const int h = (int)((idx / vec_per_head) % n_heads);
  const long long t = idx / ((long long)vec_per_head * n_heads);
  bf16* p = x + t * ld + col0 + h * head_dim + v * 8;
  float a[8], b[8], c[8], s[8];
  load8(p, a);
  load8(p + half, b);
  const float* cp = cos_t + t * half + v * 8;
  const float* sp = sin_t + t * half + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    c[j] = cp[j];
    s[j] = sp[j] * sign;
  }
  float ra[8], rb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ra[j] = a[j] * c[j] - b[j] * s[j];
    rb[j] = b[j] * c[j] + a[j] * s[j];
  }
  store8(p, ra);
  store8(p + half, rb);
  if (out) {
    bf16* q = out + t * ldo + h * head_dim + v * 8;
    store8(q, ra);
    store8(q + half, rb);
  }
}

// copy a [T, width] column block of a strided matrix (used to append V to the KV cache)
__global__ void copy_cols_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                                 int T, int width) {
  const int vpr = width >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * vpr) return;
  const long long t = idx / vpr;
  const int v = (int)(idx - t * vpr);
  *reinterpret_cast<uint4*>(dst + t * ldd + v * 8) = *reinterpret_cast<const uint4*>(src + t * lds + v * 8);
}

// ---------------------------------------------------------------------------------------------
// gated / plain activations.  gate_up: [T, 2*I] (gate | up).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, long long T, int I) {
  pdl_trigger();
  pdl_wait();
  const int vpr = I >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * vpr) return;
  const long long t = idx / vpr;
  const int v = (int)(idx - t * vpr);
  float g[8], u[8], o[8];
  load8(gu + t * 2 * I + v * 8, g);
  load8(gu + t * 2 * I + I + v * 8, u);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = g[j] * sigmoidf_(g[j]) * u[j];
  store8(act + t * I + v * 8, o);
}

__global__ void swiglu_bwd_kernel(const bf16* __restrict__ dact, const bf16* __restrict__ gu, bf16* __restrict__ dgu,
                                  long long T, int I) {
  const int vpr = I >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * vpr) return;
  const long long t = idx / vpr;
  const int v = (int)(idx - t * vpr);
  float g[8], u[8], d[8], dg[8], du[8];
  load8(gu + t * 2 * I + v * 8, g);
  load8(gu + t * 2 * I + I + v * 8, u);
  load8(dact + t * I + v * 8, d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = sigmoidf_(g[j]);
    du[j] = d[j] * g[j] * s;
    dg[j] = d[j] * u[j] * s * (1.0f + g[j] * (1.0f - s));
  }
  store8(dgu + t * 2 * I + v * 8, dg);
  store8(dgu + t * 2 * I + I + v * 8, du);
}

// mode 0: exact-erf GELU (nn.GELU, PatchMerger :139); mode 1: QuickGELU x*sigmoid(1.702x) (Qwen2-VL ViT MLP)
__global__ void act_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long n8, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float a[8], o[8];
  load8(x + idx * 8, a);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = mode == 0 ? 0.5f * a[j] * (1.0f + erff(a[j] * 0.70710678118654752f)) : a[j] * sigmoidf_(1.702f * a[j]);
  store8(y + idx * 8, o);
}
__global__ void act_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, bf16* __restrict__ dx,
                               long long n8, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float a[8], g[8], o[8];
  load8(x + idx * 8, a);
  load8(dy + idx * 8, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float d;
    if (mode == 0) {
      const float cdf = 0.5f * (1.0f + erff(a[j] * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * a[j] * a[j]);
      d = cdf + a[j] * pdf;
    } else {
      const float s = sigmoidf_(1.702f * a[j]);
      d = s + 1.702f * a[j] * s * (1.0f - s);
    }
    o[j] = g[j] * d;
  }
  store8(dx + idx * 8, o);
}

// ---------------------------------------------------------------------------------------------
// row gather / scatter (embedding lookup, vision-token scatter, window permutation, lm_head row selection)
//   gather : dst[i, :] = src[idx[i], :]        (idx < 0 -> row skipped / left untouched)
//   scatter: dst[idx[i], :] = src[i, :]   or  += (atomic bf16x2) when accumulate
// ---------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                   bf16* __restrict__ dst, long long ldd, int n, int width) {
  pdl_trigger();
  pdl_wait();
  const int vpr = width >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * vpr) return;
  const long long i = gid / vpr;
  const int v = (int)(gid - i * vpr);
  const int s = idx ? idx[i] : (int)i;
  if (s < 0) return;
  *reinterpret_cast<uint4*>(dst + i * ldd + v * 8) = *reinterpret_cast<const uint4*>(src + (long long)s * lds + v * 8);
}
__global__ void scatter_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ idx,
                                    bf16* __restrict__ dst, long long ldd, int n, int width, int accumulate) {
  const int vpr = width >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * vpr) return;
  const long long i = gid / vpr;
  const int v = (int)(gid - i * vpr);
  const int d = idx ? idx[i] : (int)i;
  if (d < 0) return;
  const uint4 val = *reinterpret_cast<const uint4*>(src + i * lds + v * 8);
  bf16* dp = dst + (long long)d * ldd + v * 8;
  if (!accumulate) {
    *reinterpret_cast<uint4*>(dp) = val;
  } else {
    const __nv_bfloat162* vv = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(reinterpret_cast<__nv_bfloat162*>(dp) + j, vv[j]);
  }
}

// pixel rows: dst_bf16[i, :] = (bf16) src_f32[idx[i], :]  (window-order permutation fused with the dtype cast)
__global__ void gather_rows_f32_kernel(const float* __restrict__ src, long long lds, const int* __restrict__ idx,
                                       bf16* __restrict__ dst, long long ldd, int n, int width) {
  const int vpr = width >> 3;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * vpr) return;
  const long long i = gid / vpr;
  const int v = (int)(gid - i * vpr);
  const long long s = idx ? idx[i] : i;
  const float4 a = *reinterpret_cast<const float4*>(src + s * lds + v * 8);
  const float4 b = *reinterpret_cast<const float4*>(src + s * lds + v * 8 + 4);
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y);
  o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(b.x, b.y);
  o.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(dst + i * ldd + v * 8) = o;
}

// bf16 transpose: in [R, C] (ld) -> out [C, R] (ldo). 64x64 tiles through padded shared memory.
__global__ void __launch_bounds__(256)
transpose_kernel(const bf16* __restrict__ in, long long ld, bf16* __restrict__ out, long long ldo, int R, int C) {
  __shared__ bf16 tile[64][64 + 8];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int e = tid + it * 256;  // 512 vectors of 8
    const int r = e >> 3, cv = (e & 7) * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < R && c0 + cv < C) val = *reinterpret_cast<const uint4*>(in + (long long)(r0 + r) * ld + c0 + cv);
    *reinterpret_cast<uint4*>(&tile[r][cv]) = val;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int e = tid + it * 256;
    const int c = e >> 3, rv = (e & 7) * 8;  // output row c (input column), 8 consecutive input rows
    if (c0 + c < C && r0 + rv < R) {
      bf16 tmp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tmp[j] = tile[rv + j][c];
      *reinterpret_cast<uint4*>(out + (long long)(c0 + c) * ldo + r0 + rv) = *reinterpret_cast<uint4*>(tmp);
    }
  }
}

// column sums: out[c] += sum_r x[r, c]  (fp32 atomics; grid = (col chunks of 256, row slabs))
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, long long ld, int R, int C, float* __restrict__ out, int rows_per_block) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float a = 0.f, b = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + (long long)r * ld + c));
    a += v.x;
    b += v.y;
  }
  atomicAdd(out + c, a);
  atomicAdd(out + c + 1, b);
}

// dst_bf16 (+)= src_f32
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src[i];
  if (accumulate) v += __bfloat162float(dst[i]);
  dst[i] = __float2bfloat16(v);
}

// y (+)= x elementwise on bf16 (n % 8 == 0)
__global__ void add_bf16_kernel(bf16* __restrict__ y, const bf16* __restrict__ x, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float a[8], b[8];
  load8(y + i * 8, a);
  load8(x + i * 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  store8(y + i * 8, a);
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// y = norm(x + delta) * w; when delta is given the (bf16-rounded) sum is also written to x_out (may alias x).
extern "C" int t1_rmsnorm(const void* x, const void* delta, void* x_out, const void* w, void* y, int T, int D,
                          float eps, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 8192, "t1_rmsnorm: D=%d must be a multiple of 8 and <= 8192", D);
  T1_CHECK_ARG(!delta || x_out, "t1_rmsnorm: x_out is required when delta is given");
  if (launch_norm_rows<false>(x, delta, x_out, w, nullptr, y, T, D, eps, stream)) {
    T1_CUDA_LAUNCH_CHECK("t1_rmsnorm");
    return T1_OK;
  }
  if (t1_launch(norm_fwd_kernel<false>, dim3(T), dim3(256), 0, stream, (const bf16*)x, (const bf16*)delta, (bf16*)x_out,
                (const bf16*)w, (const bf16*)nullptr, (bf16*)y, D, eps) != cudaSuccess) {
    t1_set_error("t1_rmsnorm: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return T1_ERR_CUDA;
  }
  T1_CUDA_LAUNCH_CHECK("t1_rmsnorm");
  return T1_OK;
}

extern "C" int t1_layernorm(const void* x, const void* delta, void* x_out, const void* w, const void* b, void* y, int T,
                            int D, float eps, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 8192, "t1_layernorm: D=%d must be a multiple of 8 and <= 8192", D);
  T1_CHECK_ARG(!delta || x_out, "t1_layernorm: x_out is required when delta is given");
  if (launch_norm_rows<true>(x, delta, x_out, w, b, y, T, D, eps, stream)) {
    T1_CUDA_LAUNCH_CHECK("t1_layernorm");
    return T1_OK;
  }
  norm_fwd_kernel<true><<<T, 256, 0, stream>>>((const bf16*)x, (const bf16*)delta, (bf16*)x_out, (const bf16*)w, (const bf16*)b, (bf16*)y, D, eps);
  T1_CUDA_LAUNCH_CHECK("t1_layernorm");
  return T1_OK;
}

extern "C" int t1_rmsnorm_bwd(const void* dy, const void* x, const void* w, const void* dres, void* dx, float* dw,
                              int T, int D, float eps, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 8192, "t1_rmsnorm_bwd: D=%d must be a multiple of 8 and <= 8192", D);
  const int nit = (D / 8 + 255) / 256;
  const int per_sm = nit <= 2 ? 2 : 1;           // resident CTAs per SM (register budget, see the kernel)
  const int grid = T < per_sm * t1_num_sms() ? T : per_sm * t1_num_sms();
  if (nit <= 1)
    rmsnorm_bwd_kernel<1><<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, (const bf16*)dres,
                                                    (bf16*)dx, dw, T, D, eps);
  else if (nit == 2)
    rmsnorm_bwd_kernel<2><<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, (const bf16*)dres,
                                                    (bf16*)dx, dw, T, D, eps);
  else
    rmsnorm_bwd_kernel<4><<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, (const bf16*)dres,
                                                    (bf16*)dx, dw, T, D, eps);
  T1_CUDA_LAUNCH_CHECK("t1_rmsnorm_bwd");
  return T1_OK;
}

// dx may be NULL (the merger's input is the frozen ViT output: only dw / db are needed).  dw / db accumulate (fp32).
extern "C" int t1_layernorm_bwd(const void* dy, const void* x, const void* w, void* dx, float* dw, float* db, int T, int D,
                                float eps, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(T > 0 && D > 0 && D % 8 == 0 && D <= 4096, "t1_layernorm_bwd: D=%d must be a multiple of 8 and <= 4096", D);
  const int nit = (D / 8 + 255) / 256;
  const int grid = T < 2 * t1_num_sms() ? T : 2 * t1_num_sms();
  if (nit <= 1)
    layernorm_bwd_kernel<1><<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, (bf16*)dx, dw, db, T, D, eps);
  else
    layernorm_bwd_kernel<2><<<grid, 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (const bf16*)w, (bf16*)dx, dw, db, T, D, eps);
  T1_CUDA_LAUNCH_CHECK("t1_layernorm_bwd");
  return T1_OK;
}

extern "C" int t1_rope_table(const int* pos, int T, const int* axis_id, const float* inv_freq, int half, float* cos_t,
                             float* sin_t, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(T > 0 && half > 0 && pos && axis_id && inv_freq && cos_t && sin_t, "t1_rope_table: bad arguments");
  rope_table_kernel<<<blocks_for((long long)T * half, 256), 256, 0, stream>>>(pos, T, axis_id, inv_freq, half, cos_t, sin_t);
  T1_CUDA_LAUNCH_CHECK("t1_rope_table");
  return T1_OK;
}

extern "C" int t1_rope_apply(void* x, long long ld, int col0, int n_heads, int head_dim, const float* cos_t,
                             const float* sin_t, int inverse, int T, void* out, long long ldo, cudaStream_t stream) {
  if (T == 0 || n_heads == 0) return T1_OK;
  T1_CHECK_ARG(head_dim % 16 == 0 && col0 % 8 == 0 && ld % 8 == 0, "t1_rope_apply: head_dim %% 16, col0 %% 8, ld %% 8 required");
  const long long total = (long long)T * n_heads * (head_dim / 16);
  rope_apply_kernel<<<blocks_for(total, 256), 256, 0, stream>>>((bf16*)x, ld, col0, n_heads, head_dim, cos_t, sin_t,
                                                               inverse ? -1.f : 1.f, T, (bf16*)out, ldo);
  T1_CUDA_LAUNCH_CHECK("t1_rope_apply");
  return T1_OK;
}

extern "C" int t1_copy_cols(const void* src, long long lds, void* dst, long long ldd, int T, int width,
                            cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(width % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "t1_copy_cols: widths must be multiples of 8");
  copy_cols_kernel<<<blocks_for((long long)T * (width / 8), 256), 256, 0, stream>>>((const bf16*)src, lds, (bf16*)dst, ldd, T, width);
  T1_CUDA_LAUNCH_CHECK("t1_copy_cols");
  return T1_OK;
}

extern "C" int t1_swiglu(const void* gate_up, void* act, long long T, int I, cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(I % 8 == 0, "t1_swiglu: I must be a multiple of 8");
  if (t1_launch(swiglu_fwd_kernel, dim3(blocks_for(T * (I / 8), 256)), dim3(256), 0, stream, (const bf16*)gate_up, (bf16*)act,
                T, I) != cudaSuccess) {
    t1_set_error("t1_swiglu: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return T1_ERR_CUDA;
  }
  T1_CUDA_LAUNCH_CHECK("t1_swiglu");
  return T1_OK;
}

extern "C" int t1_swiglu_bwd(const void* dact, const void* gate_up, void* dgate_up, long long T, int I,
                             cudaStream_t stream) {
  if (T == 0) return T1_OK;
  T1_CHECK_ARG(I % 8 == 0, "t1_swiglu_bwd: I must be a multiple of 8");
  swiglu_bwd_kernel<<<blocks_for(T * (I / 8), 256), 256, 0, stream>>>((const bf16*)dact, (const bf16*)gate_up, (bf16*)dgate_up, T, I);
  T1_CUDA_LAUNCH_CHECK("t1_swiglu_bwd");
  return T1_OK;
}

extern "C" int t1_act(const void* x, void* y, long long n, int mode, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n % 8 == 0 && (mode == 0 || mode == 1), "t1_act: n %% 8 and mode in {0,1} required");
  act_fwd_kernel<<<blocks_for(n / 8, 256), 256, 0, stream>>>((const bf16*)x, (bf16*)y, n / 8, mode);
  T1_CUDA_LAUNCH_CHECK("t1_act");
  return T1_OK;
}

extern "C" int t1_act_bwd(const void* dy, const void* x, void* dx, long long n, int mode, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n % 8 == 0 && (mode == 0 || mode == 1), "t1_act_bwd: n %% 8 and mode in {0,1} required");
  act_bwd_kernel<<<blocks_for(n / 8, 256), 256, 0, stream>>>((const bf16*)dy, (const bf16*)x, (bf16*)dx, n / 8, mode);
  T1_CUDA_LAUNCH_CHECK("t1_act_bwd");
  return T1_OK;
}

extern "C" int t1_gather_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n, int width,
                              cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(width % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "t1_gather_rows: widths must be multiples of 8");
  if (t1_launch(gather_rows_kernel, dim3(blocks_for((long long)n * (width / 8), 256)), dim3(256), 0, stream, (const bf16*)src,
                lds, idx, (bf16*)dst, ldd, n, width) != cudaSuccess) {
    t1_set_error("t1_gather_rows: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return T1_ERR_CUDA;
  }
  T1_CUDA_LAUNCH_CHECK("t1_gather_rows");
  return T1_OK;
}

extern "C" int t1_gather_rows_f32(const float* src, long long lds, const int* idx, void* dst, long long ldd, int n,
                                  int width, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(width % 8 == 0 && lds % 4 == 0 && ldd % 8 == 0, "t1_gather_rows_f32: width %% 8, lds %% 4, ldd %% 8 required");
  gather_rows_f32_kernel<<<blocks_for((long long)n * (width / 8), 256), 256, 0, stream>>>(src, lds, idx, (bf16*)dst, ldd, n, width);
  T1_CUDA_LAUNCH_CHECK("t1_gather_rows_f32");
  return T1_OK;
}

extern "C" int t1_scatter_rows(const void* src, long long lds, const int* idx, void* dst, long long ldd, int n,
                               int width, int accumulate, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(width % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "t1_scatter_rows: widths must be multiples of 8");
  scatter_rows_kernel<<<blocks_for((long long)n * (width / 8), 256), 256, 0, stream>>>((const bf16*)src, lds, idx, (bf16*)dst, ldd, n, width, accumulate);
  T1_CUDA_LAUNCH_CHECK("t1_scatter_rows");
  return T1_OK;
}

extern "C" int t1_transpose_bf16(const void* in, long long ld, void* out, long long ldo, int R, int C,
                                 cudaStream_t stream) {
  if (R == 0 || C == 0) return T1_OK;
  T1_CHECK_ARG(ld % 8 == 0 && ldo % 8 == 0 && R % 8 == 0 && C % 8 == 0, "t1_transpose_bf16: R, C, ld, ldo must be multiples of 8 (R=%d C=%d)", R, C);
  dim3 grid((C + 63) / 64, (R + 63) / 64);
  transpose_kernel<<<grid, 256, 0, stream>>>((const bf16*)in, ld, (bf16*)out, ldo, R, C);
  T1_CUDA_LAUNCH_CHECK("t1_transpose_bf16");
  return T1_OK;
}

extern "C" int t1_colsum(const void* x, long long ld, int R, int C, float* out, cudaStream_t stream) {
  if (R == 0 || C == 0) return T1_OK;
  T1_CHECK_ARG(C % 2 == 0 && ld % 2 == 0, "t1_colsum: C and ld must be even");
  const int rows_per_block = 64;
  dim3 grid((C / 2 + 255) / 256, (R + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, 256, 0, stream>>>((const bf16*)x, ld, R, C, out, rows_per_block);
  T1_CUDA_LAUNCH_CHECK("t1_colsum");
  return T1_OK;
}

extern "C" int t1_f32_to_bf16(const float* src, void* dst, long long n, int accumulate, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  f32_to_bf16_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(src, (bf16*)dst, n, accumulate);
  T1_CUDA_LAUNCH_CHECK("t1_f32_to_bf16");
  return T1_OK;
}

extern "C" int t1_add_bf16(void* y, const void* x, long long n, cudaStream_t stream) {
  if (n == 0) return T1_OK;
  T1_CHECK_ARG(n % 8 == 0, "t1_add_bf16: n must be a multiple of 8");
  add_bf16_kernel<<<blocks_for(n / 8, 256), 256, 0, stream>>>((bf16*)y, (const bf16*)x, n / 8);
  T1_CUDA_LAUNCH_CHECK("t1_add_bf16");
  return T1_OK;
}
