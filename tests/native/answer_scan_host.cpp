// Test harness only: compiles the device scanner's header (time_r1_b200/csrc/answer_scan.cuh) for the host with g++ so
// that tests/test_answer_scan_cpu.py can check the exact code the kernels run against CPython's re / float() without a
// GPU.  Not part of the package, never loaded by the product.
#include "answer_scan.cuh"

namespace {
struct OneLane {
  static constexpr int kLanes = 1;
  int lane() const { return 0; }
  uint32_t ballot(bool p) const { return p ? 1u : 0u; }
  static int first_set(uint32_t) { return 0; }
  void sync() const {}
};
}  // namespace

extern "C" {

void as_scan_batch(const uint8_t* bytes, const long long* off, int n, int mode, double* s, double* e, uint8_t* valid,
                   uint8_t* fmt) {
  OneLane w;
  for (int i = 0; i < n; ++i) {
    const ascan::Parsed r = ascan::scan_row(bytes + off[i], (int)(off[i + 1] - off[i]), mode, w);
    s[i] = r.s; e[i] = r.e; valid[i] = r.valid; fmt[i] = r.format;
  }
}

void as_float_batch(const uint8_t* bytes, const long long* off, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = ascan::to_double(bytes + off[i], 0, (int)(off[i + 1] - off[i]));
}

void as_float_slow_batch(const uint8_t* bytes, const long long* off, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = ascan::to_double_slow(bytes + off[i], 0, (int)(off[i + 1] - off[i]));
}

int as_digit_value(uint32_t cp) { return ascan::digit_value(cp); }
int as_is_space(uint32_t cp) { return ascan::is_space(cp) ? 1 : 0; }
}
