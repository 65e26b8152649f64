// Dual-number arithmetic (ceres::Jet restated) and the reference's trilinear sampler,
// shared by sampler.cu and the fused loss kernel in edges.cu.
#pragma once
#include "common.cuh"

namespace mo {

template <class S> struct Ops;
template <> struct Ops<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
};
template <> struct Ops<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
};

// ceres::Jet<S,3>: f*g = (f.a*g.a, f.a*g.v + f.v*g.a), f+-g componentwise.
template <class S>
struct Jet3 {
  S a, v0, v1, v2;
  __device__ __forceinline__ Jet3() {}
  __device__ __forceinline__ explicit Jet3(S s) : a(s), v0(S(0)), v1(S(0)), v2(S(0)) {}
  __device__ __forceinline__ Jet3(S s, S x, S y, S z) : a(s), v0(x), v1(y), v2(z) {}
};
template <class S> __device__ __forceinline__ Jet3<S> operator*(const Jet3<S>& f, const Jet3<S>& g) {
  typedef Ops<S> O;
  return Jet3<S>(O::mul(f.a, g.a), O::add(O::mul(f.a, g.v0), O::mul(f.v0, g.a)), O::add(O::mul(f.a, g.v1), O::mul(f.v1, g.a)),
                 O::add(O::mul(f.a, g.v2), O::mul(f.v2, g.a)));
}
template <class S> __device__ __forceinline__ Jet3<S> operator+(const Jet3<S>& f, const Jet3<S>& g) {
  typedef Ops<S> O;
  return Jet3<S>(O::add(f.a, g.a), O::add(f.v0, g.v0), O::add(f.v1, g.v1), O::add(f.v2, g.v2));
}
template <class S> __device__ __forceinline__ Jet3<S> operator-(const Jet3<S>& f, const Jet3<S>& g) {
  typedef Ops<S> O;
  return Jet3<S>(O::sub(f.a, g.a), O::sub(f.v0, g.v0), O::sub(f.v1, g.v1), O::sub(f.v2, g.v2));
}
template <class S> __device__ __forceinline__ Jet3<S> operator-(const Jet3<S>& f) { return Jet3<S>(-f.a, -f.v0, -f.v1, -f.v2); }

// scalar "T" wrappers so the sampler below is written once
template <class S>
struct Num {
  S a;
  __device__ __forceinline__ Num() {}
  __device__ __forceinline__ explicit Num(S s) : a(s) {}
};
template <class S> __device__ __forceinline__ Num<S> operator*(const Num<S>& f, const Num<S>& g) { return Num<S>(Ops<S>::mul(f.a, g.a)); }
template <class S> __device__ __forceinline__ Num<S> operator+(const Num<S>& f, const Num<S>& g) { return Num<S>(Ops<S>::add(f.a, g.a)); }
template <class S> __device__ __forceinline__ Num<S> operator-(const Num<S>& f, const Num<S>& g) { return Num<S>(Ops<S>::sub(f.a, g.a)); }
template <class S> __device__ __forceinline__ Num<S> operator-(const Num<S>& f) { return Num<S>(-f.a); }

// UniformGrid::distance<T> / DistanceFloat<T> with T = Num<S> or Jet3<S>; grid holds S.
template <class T, class S>
__device__ __forceinline__ T sample(const S* __restrict__ grid, const int n, const T p0, const T p1, const T p2) {
  const int px = (int)Ops<S>::mul(p0.a, (S)n);   // uniformgrid.cc:20-22 / :87-89 (C-cast truncation)
  const int py = (int)Ops<S>::mul(p1.a, (S)n);
  const int pz = (int)Ops<S>::mul(p2.a, (S)n);
  const T tn((S)n);
  if (px < 0 || py < 0 || pz < 0 || px >= n - 1 || py >= n - 1 || pz >= n - 1) {   // :23-26
    const T edge((S)(n - 1 - 1e-3));
    T l((S)0);
    if (px < 0) l = l + (-p0) * tn;                  // :29-30
    else if (px >= n) l = l + (p0 * tn - edge);      // :31-33
    if (py < 0) l = l + (-p1) * tn;
    else if (py >= n) l = l + (p1 * tn - edge);
    if (pz < 0) l = l + (-p2) * tn;
    else if (pz >= n) l = l + (p2 * tn - edge);
    return l;
  }
  const T wx = p0 * tn - T((S)px), wy = p1 * tn - T((S)py), wz = p2 * tn - T((S)pz);   // :50-52
  const T one((S)1);
  const T ux = one - wx, uy = one - wy, uz = one - wz;
  const size_t nn = (size_t)n;
  const S* g0 = grid + ((size_t)pz * nn + (size_t)py) * nn + (size_t)px;
  const S* g1 = g0 + nn * nn;
  const S c000 = __ldg(g0), c001 = __ldg(g0 + 1), c010 = __ldg(g0 + nn), c011 = __ldg(g0 + nn + 1);
  const S c100 = __ldg(g1), c101 = __ldg(g1 + 1), c110 = __ldg(g1 + nn), c111 = __ldg(g1 + nn + 1);
  const T w0 = ux * uy * uz * T(c000);   // :54-76
  const T w1 = wx * uy * uz * T(c001);
  const T w2 = ux * wy * uz * T(c010);
  const T w3 = wx * wy * uz * T(c011);
  const T w4 = ux * uy * wz * T(c100);
  const T w5 = wx * uy * wz * T(c101);
  const T w6 = ux * wy * wz * T(c110);
  const T w7 = wx * wy * wz * T(c111);
  const T res = w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7;   // :78
  if (res.a > (S)0.2) return T((S)0);                    // :80-81
  return res;
}

}  // namespace mo
