// 2^x on the FMA / ALU pipes instead of the MUFU (groundwork for taking the softmax of the attention kernels off the
// 4-lane special-function unit: `MUFU.EX2` issues one warp instruction per 8 clocks per scheduler, an FFMA one per clock,
// so a polynomial of <= 8 FMA-pipe instructions evaluated for half of the scores doubles the exponential throughput —
// DESIGN.md §7 "next").  Not used by any kernel yet; accuracy is pinned on the CPU (tests/test_fast_exp2_cpu.py).
//
//   t  = x + 1.5 * 2^23          rounds x to the nearest integer n, which lands in the low mantissa bits of t
//   f  = x - (t - 1.5 * 2^23)    in [-0.5, 0.5]
//   2^x = p(f) * 2^n             p = degree-DEG minimax fit of 2^f (relative error 7.5e-5 / 2.7e-6 / 2.4e-7 for DEG 3 / 4 / 5;
//                                bf16 probabilities need 3.9e-3), 2^n applied by adding n << 23 to the exponent field
// Inputs <= -126 return 0 (what ex2.approx.ftz does with a subnormal result: masked scores stay exact zeros); inputs
// above 127 are not supported (softmax arguments are <= 0, or <= 8 with the lazy rescale).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
#endif

namespace fast_exp2 {

FE_HD int32_t f2i(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  int32_t i;
  memcpy(&i, &x, 4);
  return i;
#endif
}
FE_HD float i2f(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float x;
  memcpy(&x, &i, 4);
  return x;
#endif
}
FE_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}

template <int DEG>
FE_HD float poly(float f);
template <>
FE_HD float poly<3>(float f) {
  return fma_(fma_(fma_(0.055171653628349304f, f, 0.2426111251115799f), f, 0.6932609677314758f), f, 0.9999280571937561f);
}
template <>
FE_HD float poly<4>(float f) {
  return fma_(fma_(fma_(fma_(0.009570100344717503f, f, 0.05591785907745361f), f, 0.240247443318367f), f, 0.6931217908859253f), f,
              0.9999992847442627f);
}
template <>
FE_HD float poly<5>(float f) {
  return fma_(fma_(fma_(fma_(fma_(0.0013276470126584172f, f, 0.009675540961325169f), f, 0.05550713464617729f), f,
                        0.24022120237350464f), f, 0.6931469440460205f), f, 1.0000001192092896f);
}

template <int DEG>
FE_HD float exp2(float x) {
  const float magic = 12582912.0f;                 // 1.5 * 2^23
  const float xc = x < -126.0f ? -126.0f : x;      // also catches -inf
  const float t = xc + magic;
  const float f = xc - (t - magic);
  const float p = poly<DEG>(f);
  const float r = i2f(f2i(p) + (int32_t)((uint32_t)f2i(t) << 23));
  return x <= -126.0f ? 0.0f : r;
}

}  // namespace fast_exp2
