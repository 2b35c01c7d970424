// Forward-mode dual numbers (value + N tangents) so that the backward kernels can differentiate the *same* templated
// math the forward kernels run (csrc/sg_math.cuh) instead of a second, hand-derived formula set.
#pragma once
#include "sg_math.cuh"

namespace nefii {

template <typename S, int N> struct Dual {
  S v;
  S d[N];
  NEFII_HD constexpr Dual() : v(S(0)), d{} {}
  NEFII_HD constexpr Dual(S x) : v(x), d{} {}
  NEFII_HD constexpr Dual(double x, int) : v(S(x)), d{} {}
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value && !std::is_same<U, S>::value>::type>
  NEFII_HD constexpr Dual(U x) : v(S(x)), d{} {}
  NEFII_HD static Dual variable(S x, int i) { Dual r(x); r.d[i] = S(1); return r; }
};

#define NEFII_DUAL_T template <typename S, int N> NEFII_HD
NEFII_DUAL_T Dual<S, N> operator+(const Dual<S, N>& a, const Dual<S, N>& b) { Dual<S, N> r(a.v + b.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
NEFII_DUAL_T Dual<S, N> operator-(const Dual<S, N>& a, const Dual<S, N>& b) { Dual<S, N> r(a.v - b.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
NEFII_DUAL_T Dual<S, N> operator-(const Dual<S, N>& a) { Dual<S, N> r(-a.v); for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
NEFII_DUAL_T Dual<S, N> operator*(const Dual<S, N>& a, const Dual<S, N>& b) { Dual<S, N> r(a.v * b.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
NEFII_DUAL_T Dual<S, N> operator/(const Dual<S, N>& a, const Dual<S, N>& b) {
  const S inv = S(1) / b.v;
  Dual<S, N> r(a.v * inv);
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
NEFII_DUAL_T Dual<S, N>& operator+=(Dual<S, N>& a, const Dual<S, N>& b) { a = a + b; return a; }
NEFII_DUAL_T bool operator<(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v < b.v; }
NEFII_DUAL_T bool operator>(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v > b.v; }
NEFII_DUAL_T bool operator>=(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v >= b.v; }
NEFII_DUAL_T bool operator<=(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v <= b.v; }
NEFII_DUAL_T bool operator!=(const Dual<S, N>& a, const Dual<S, N>& b) { return a.v != b.v; }
#undef NEFII_DUAL_T

// math overloads live next to Dual so that argument-dependent lookup finds them from inside sgm's templates
template <typename S, int N> NEFII_HD Dual<S, N> m_exp(const Dual<S, N>& x) {
  Dual<S, N> r(sgm::m_exp(x.v));
  for (int i = 0; i < N; ++i) r.d[i] = r.v * x.d[i];
  return r;
}
template <typename S, int N> NEFII_HD Dual<S, N> m_sqrt(const Dual<S, N>& x) {
  Dual<S, N> r(sgm::m_sqrt(x.v));
  const S h = S(0.5) / r.v;
  for (int i = 0; i < N; ++i) r.d[i] = h * x.d[i];
  return r;
}
template <typename S, int N> NEFII_HD Dual<S, N> m_abs(const Dual<S, N>& x) { return x.v < S(0) ? -x : x; }
template <typename S, int N> NEFII_HD Dual<S, N> m_pow(const Dual<S, N>& a, const Dual<S, N>& b) {
  Dual<S, N> r(sgm::m_pow(a.v, b.v));   // only used with constant base / exponents whose tangents are zero on this path
  return r;
}

}  // namespace nefii
