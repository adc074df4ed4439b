// Compile-time helpers: unrolled loops with constant indices, and twiddle factors evaluated by
// the compiler (so the register-resident butterflies take them as FFMA immediates).
#pragma once
#include <type_traits>

namespace sg {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

constexpr int bitrev(int x, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

namespace ct {
constexpr double kPi = 3.141592653589793238462643383279502884;
// Taylor series on |x| <= pi/4 (error < 1e-17 there)
constexpr double sin_small(double x) {
  const double x2 = x * x;
  double term = x, sum = x;
  for (int k = 1; k <= 11; ++k) { term *= -x2 / ((2 * k) * (2 * k + 1)); sum += term; }
  return sum;
}
constexpr double cos_small(double x) {
  const double x2 = x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k <= 11; ++k) { term *= -x2 / ((2 * k - 1) * (2 * k)); sum += term; }
  return sum;
}
// cos / sin of 2*pi*p/n by exact octant reduction on the integers
struct CS { double c, s; };
constexpr CS cossin_frac(long long p, long long n) {
  long long t = ((8 * p) % (8 * n) + 8 * n) % (8 * n);
  const int oct = (int)(t / n);
  const double a = (kPi / 4) * (double)(t % n) / (double)n, b = kPi / 4 - a;
  switch (oct) {
    case 0: return {cos_small(a), sin_small(a)};
    case 1: return {sin_small(b), cos_small(b)};
    case 2: return {-sin_small(a), cos_small(a)};
    case 3: return {-cos_small(b), sin_small(b)};
    case 4: return {-cos_small(a), -sin_small(a)};
    case 5: return {-sin_small(b), -cos_small(b)};
    case 6: return {sin_small(a), -cos_small(a)};
    default: return {cos_small(b), -sin_small(b)};
  }
}
}  // namespace ct

// W_N^P = exp(-2 pi i P / N) as float immediates
template <int P, int N>
struct Twiddle {
  static constexpr float re = (float)ct::cossin_frac(P, N).c;
  static constexpr float im = (float)(-ct::cossin_frac(P, N).s);
};

}  // namespace sg
