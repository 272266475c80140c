// Test harness only: compiles time_r1_b200/csrc/fast_exp2.cuh for the host (tests/test_fast_exp2_cpu.py).
#include "fast_exp2.cuh"

extern "C" void fe_exp2_batch(const float* x, int n, int deg, float* out) {
  for (int i = 0; i < n; ++i)
    out[i] = deg == 3 ? fast_exp2::exp2<3>(x[i]) : (deg == 4 ? fast_exp2::exp2<4>(x[i]) : fast_exp2::exp2<5>(x[i]));
}
