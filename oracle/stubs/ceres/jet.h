// Minimal stand-in for <ceres/jet.h> (see oracle/stubs/README.md). Test infrastructure only.
// Dual number a + sum_i v[i] eps_i with the operator definitions published in ceres/jet.h.
#ifndef MESHODE_STUB_CERES_JET_
#define MESHODE_STUB_CERES_JET_
#include <cmath>

#include <Eigen/Core>   // as the real ceres/jet.h does
namespace ceres {
template <typename T, int N>
struct Jet {
  T a;      // scalar part first: the reference type-puns &jet to T* (uniformgrid.cc:20, :87)
  T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  explicit Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1); }
  // Jet(a, Eigen::DenseBase<Derived>): scalar part and the N partials given as a vector (src/interface/distance_layer.cc:62-66)
  template <class S> Jet(const T& value, const Eigen::Matrix<S, N, 1>& vec) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(vec[i]); }
  // compound operators as in ceres/jet.h: x op= y  is  x = x op y
  Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
  Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};
#define MESHODE_JET_LOOP for (int i = 0; i < N; ++i)
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = -f.a; MESHODE_JET_LOOP h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a + g.a; MESHODE_JET_LOOP h.v[i] = f.v[i] + g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a - g.a; MESHODE_JET_LOOP h.v[i] = f.v[i] - g.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a * g.a; MESHODE_JET_LOOP h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; const T g_a_inverse = T(1.0) / g.a; const T f_a_by_g_a = f.a * g_a_inverse; h.a = f_a_by_g_a;
  MESHODE_JET_LOOP h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) {
  Jet<T, N> h; h.a = s - f.a; MESHODE_JET_LOOP h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) {
  Jet<T, N> h; h.a = f.a * s; MESHODE_JET_LOOP h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) {
  Jet<T, N> h; h.a = f.a * s; MESHODE_JET_LOOP h.v[i] = f.v[i] * s; return h; }
#define MESHODE_JET_CMP(op)                                                                              \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  template <typename T, int N> inline bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; }           \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
MESHODE_JET_CMP(<) MESHODE_JET_CMP(<=) MESHODE_JET_CMP(>) MESHODE_JET_CMP(>=) MESHODE_JET_CMP(==) MESHODE_JET_CMP(!=)
#undef MESHODE_JET_CMP
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::sqrt(f.a); const T two_a_inverse = T(1.0) / (T(2.0) * h.a);
  MESHODE_JET_LOOP h.v[i] = f.v[i] * two_a_inverse; return h; }
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::cos(f.a); const T m = -std::sin(f.a); MESHODE_JET_LOOP h.v[i] = m * f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); MESHODE_JET_LOOP h.v[i] = c * f.v[i]; return h; }
#undef MESHODE_JET_LOOP
using std::cos;
using std::sin;
using std::sqrt;
}  // namespace ceres
#endif
