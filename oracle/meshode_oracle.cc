// meshode_oracle.cc -- CPU restatement of MeshODE's data-parallel hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the "port" CPU
// baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it.  The product (meshode_b200/) never does.
//
// WHAT PINS IT.  The reference ships no tests, golden vectors or expected outputs
//  and cannot be built as a whole here (Eigen, libigl, Ceres,
// CGAL absent, 3rd_party/* empty).  So:
//   * PINNED against the reference's own compiled code: both trilinear samplers
//     (scalar and Jet), DistanceLoss, EdgeLoss, AdaptiveEdgeLoss, EdgeLossWithRot, and
//     the interface loops of the pyDeform module -- DistanceFieldLoss_forward/backward,
//     Store{Rigidity,Graph,Cad}Information, {Rigid,Graph,Cad}EdgeLoss_forward/backward,
//     Normalize/DenormalizeByTemplate.  oracle/Makefile compiles src/lib/uniformgrid.cc +
//     distanceloss.h + edgeloss.h and src/interface/{distance,rigid,graph,cad}_layer.cc +
//     normalize.cc where they lie into oracle/_ref/ (against stand-in Eigen / Ceres /
//     torch::Tensor headers, oracle/stubs/); tests/test_oracle_ref.py compares bit for
//     bit, tests/golden/golden_ref.npz and golden_iface.npz carry the same vectors to
//     machines without /root/reference.
//   * PINNED against torch: the float32 Adam restatement (tests/test_oracle_kat.py).
//   * PARITY UNPINNED: the libigl nearest-triangle query
//     (igl::point_mesh_squared_distance @ 7100764c, call site src/lib/mesh.cc:140),
//     restated as an exact FP64 closest-point search (Ericson, "Real-Time Collision
//     Detection" 5.1.5, which is what libigl's point_simplex_squared_distance
//     implements) -- the distance is a unique mathematical quantity, only the index at
//     exact ties is implementation defined; Mesh::Normalize / CopyTensorToMesh (mesh.cc
//     pulls libigl and CGAL into its translation unit), restated line by line; and, in
//     oracle/lm.py, Ceres' Levenberg-Marquardt loop (ceres-solver @ d93fac4b).
//
// Build: g++ -O2 -ffp-contract=off (the reference's Release flags are -O2,
// CMakeLists.txt:14; contraction is off so results do not depend on -march).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// ceres::Jet<T,N> restated (ceres/jet.h): value a + N partials v.
// ---------------------------------------------------------------------------
template <class T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  template <class S> explicit Jet(const S& s) : a(T(s)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
};
template <class T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> r; r.a = -f.a; for (int i = 0; i < N; ++i) r.v[i] = -f.v[i]; return r; }
template <class T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r; r.a = f.a + g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i]; return r; }
template <class T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r; r.a = f.a - g.a; for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i]; return r; }
template <class T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r; r.a = f.a * g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
template <class T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // ceres: g_a_inverse = 1/g.a; f_a_by_g_a = f.a * g_a_inverse;
  //        v = (f.v - f_a_by_g_a * g.v) * g_a_inverse
  Jet<T, N> r; const T gi = T(1) / g.a; const T fg = f.a * gi; r.a = fg;
  for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - fg * g.v[i]) * gi; return r; }
template <class T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <class T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> r; r.a = std::sqrt(f.a); const T two_a_inv = T(1) / (T(2) * r.a);
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * two_a_inv; return r; }
template <class T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> r; r.a = std::cos(f.a); const T ms = -std::sin(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = ms * f.v[i]; return r; }
template <class T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> r; r.a = std::sin(f.a); const T c = std::cos(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = c * f.v[i]; return r; }
using std::cos; using std::sin; using std::sqrt;

template <class T> struct Scalar { typedef T type; static T get(const T& x) { return x; } };
template <class T, int N> struct Scalar<Jet<T, N> > { typedef T type; static T get(const Jet<T, N>& x) { return x.a; } };

// ---------------------------------------------------------------------------
// UniformGrid::distance<T> / DistanceFloat<T>  (src/lib/uniformgrid.cc:18-83,
// :85-150).  grid is the FP64 voxel array in [z][y][x] order
// (uniformgrid.h:25-31, mesh.cc:143-147).  S = scalar type used for the
// C-cast index (double in distance<>, float in DistanceFloat<>).
// ---------------------------------------------------------------------------
template <class T>
T sample_grid(const double* grid, int n, const T* const p) {
  typedef typename Scalar<T>::type S;
  int px = Scalar<T>::get(p[0]) * n;   // uniformgrid.cc:20 / :87 -- S*int -> S, C-cast truncation
  int py = Scalar<T>::get(p[1]) * n;
  int pz = Scalar<T>::get(p[2]) * n;
  if (px < 0 || py < 0 || pz < 0 || px >= n - 1 || py >= n - 1 || pz >= n - 1) {   // :23-26
    T l = (T)0;
    if (px < 0) l = l + -p[0] * (T)n;                                     // :29-30
    else if (px >= n) l = l + (p[0] * (T)n - (T)(n - 1 - 1e-3));          // :31-33
    if (py < 0) l = l + -p[1] * (T)n;
    else if (py >= n) l = l + (p[1] * (T)n - (T)(n - 1 - 1e-3));
    if (pz < 0) l = l + -p[2] * (T)n;
    else if (pz >= n) l = l + (p[2] * (T)n - (T)(n - 1 - 1e-3));
    return l;
  }
  T wx = p[0] * (T)n - (T)px;   // :50-52
  T wy = p[1] * (T)n - (T)py;
  T wz = p[2] * (T)n - (T)pz;
  const size_t nn = (size_t)n;
#define G(z, y, x) (T)((S)grid[((size_t)(z) * nn + (size_t)(y)) * nn + (size_t)(x)])
  T w0 = ((T)1 - wx) * ((T)1 - wy) * ((T)1 - wz) * G(pz, py, px);           // :54-76
  T w1 = wx * ((T)1 - wy) * ((T)1 - wz) * G(pz, py, px + 1);
  T w2 = ((T)1 - wx) * wy * ((T)1 - wz) * G(pz, py + 1, px);
  T w3 = wx * wy * ((T)1 - wz) * G(pz, py + 1, px + 1);
  T w4 = ((T)1 - wx) * ((T)1 - wy) * wz * G(pz + 1, py, px);
  T w5 = wx * ((T)1 - wy) * wz * G(pz + 1, py, px + 1);
  T w6 = ((T)1 - wx) * wy * wz * G(pz + 1, py + 1, px);
  T w7 = wx * wy * wz * G(pz + 1, py + 1, px + 1);
#undef G
  T res = w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7;   // :78
  if (res > (T)0.2) return T(0);                   // :80-81
  return res;
}

// ---------------------------------------------------------------------------
// Closest point on triangle (Ericson 5.1.5) -- the arithmetic behind
// igl::point_mesh_squared_distance at src/lib/mesh.cc:140.  Returns |p-c|^2.
// Deviation (documented): 0/0 in the edge parametrisations of degenerate
// triangles is guarded to 0 instead of producing NaN.
// ---------------------------------------------------------------------------
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double safe_div(double num, double den) { return den != 0.0 ? num / den : 0.0; }

inline double point_triangle_sqr(const double* p, const double* a, const double* b, const double* c,
                                 double* closest) {
  double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  double ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
  double q[3];
  const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  bool done = false;
  if (d1 <= 0.0 && d2 <= 0.0) { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; done = true; }
  double d3 = 0, d4 = 0, d5 = 0, d6 = 0;
  if (!done) {
    double bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
    d3 = dot3(ab, bp); d4 = dot3(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; done = true; }
  }
  double vc = 0;
  if (!done) {
    vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
      const double v = safe_div(d1, d1 - d3);
      q[0] = a[0] + v * ab[0]; q[1] = a[1] + v * ab[1]; q[2] = a[2] + v * ab[2]; done = true;
    }
  }
  if (!done) {
    double cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    d5 = dot3(ab, cp); d6 = dot3(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; done = true; }
  }
  double vb = 0;
  if (!done) {
    vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
      const double w = safe_div(d2, d2 - d6);
      q[0] = a[0] + w * ac[0]; q[1] = a[1] + w * ac[1]; q[2] = a[2] + w * ac[2]; done = true;
    }
  }
  if (!done) {
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
      const double w = safe_div(d4 - d3, (d4 - d3) + (d5 - d6));
      q[0] = b[0] + w * (c[0] - b[0]); q[1] = b[1] + w * (c[1] - b[1]); q[2] = b[2] + w * (c[2] - b[2]);
      done = true;
    } else {
      const double sum = va + vb + vc;
      if (sum != 0.0) {
        const double denom = 1.0 / sum;
        const double v = vb * denom, w = vc * denom;
        q[0] = a[0] + ab[0] * v + ac[0] * w; q[1] = a[1] + ab[1] * v + ac[1] * w; q[2] = a[2] + ab[2] * v + ac[2] * w;
      } else {   // fully degenerate triangle that escaped every region test
        q[0] = a[0]; q[1] = a[1]; q[2] = a[2];
      }
    }
  }
  if (closest) { closest[0] = q[0]; closest[1] = q[1]; closest[2] = q[2]; }
  const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
  return dx * dx + dy * dy + dz * dz;
}

template <class Fn>
void parallel_slices(int z0, int z1, int nthreads, Fn fn) {
  if (nthreads <= 1 || z1 - z0 <= 1) { for (int z = z0; z < z1; ++z) fn(z); return; }
  std::atomic<int> next(z0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&]() { for (;;) { int z = next.fetch_add(1); if (z >= z1) break; fn(z); } });
  for (auto& th : pool) th.join();
}

// AngleAxisRotatePoint restated (ceres/rotation.h), used by EdgeLossWithRot.
template <class T>
void angle_axis_rotate_point(const T aa[3], const T pt[3], T result[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// EdgeLossWithRot::operator() (src/lib/edgeloss.h:69-87)
template <class T>
void edge_rot_functor(const T* p1, const T* p2, const T* rot1, const T* rot2, const double* v, double lambda,
                      T* residuals) {
  T p[3], q[3];
  p[0] = p1[0] - p2[0]; p[1] = p1[1] - p2[1]; p[2] = p1[2] - p2[2];
  angle_axis_rotate_point(rot1, p, q);
  residuals[0] = (q[0] - (T)v[0]) * (T)lambda;
  residuals[1] = (q[1] - (T)v[1]) * (T)lambda;
  residuals[2] = (q[2] - (T)v[2]) * (T)lambda;
  residuals[3] = (rot1[0] - rot2[0]) * (T)1;
  residuals[4] = (rot1[1] - rot2[1]) * (T)1;
  residuals[5] = (rot1[2] - rot2[2]) * (T)1;
}

}  // namespace

extern "C" {

// ---- Mesh::Normalize after CopyTensorToMesh(normalize=1) ------------------
// src/interface/mesh_tensor.cc:62-84 (float32 -> FT=double), src/lib/mesh.cc:66-85.
void orc_normalize_target(const float* V, int nV, double* Vn, double* scale_out, double* pos_out) {
  double min_p[3], max_p[3];
  for (int j = 0; j < 3; ++j) {
    min_p[j] = 1e30; max_p[j] = -1e30;
    for (int i = 0; i < nV; ++i) {
      const double x = (double)V[i * 3 + j];
      if (x < min_p[j]) min_p[j] = x;
      if (x > max_p[j]) max_p[j] = x;
    }
  }
  const double scale = std::max(max_p[0] - min_p[0], std::max(max_p[1] - min_p[1], max_p[2] - min_p[2])) * 1.1;
  double pos[3];
  for (int j = 0; j < 3; ++j) pos[j] = min_p[j] - 0.05 * scale;
  for (int i = 0; i < nV; ++i)
    for (int j = 0; j < 3; ++j) Vn[i * 3 + j] = ((double)V[i * 3 + j] - pos[j]) / scale;
  *scale_out = scale;
  pos_out[0] = pos[0]; pos_out[1] = pos[1]; pos_out[2] = pos[2];
}

// Mesh::ApplyTransform (src/lib/mesh.cc:98-105) on double vertices.
void orc_apply_transform(const double* V, int nV, double scale, const double* pos, double* Vn) {
  for (int i = 0; i < nV; ++i)
    for (int j = 0; j < 3; ++j) Vn[i * 3 + j] = (V[i * 3 + j] - pos[j]) / scale;
}

double orc_point_triangle_sqr(const double* p, const double* a, const double* b, const double* c, double* closest) {
  return point_triangle_sqr(p, a, b, c, closest);
}

// ---- Mesh::ConstructDistanceField (src/lib/mesh.cc:106-152) ---------------
// Query point of voxel (i,j,k) is (k/N, j/N, i/N) (:112-120); stored value is
// sqrt(min_t d^2) (:146).  grid / idx are full N^3 arrays in [z][y][x] order;
// only z in [z0,z1) is written.  Ties: lowest triangle index wins.
void orc_build_grid_brute(const double* Vn, const int* F, int nF, int N, int z0, int z1, double* grid, int* idx,
                          int nthreads) {
  parallel_slices(z0, z1, nthreads, [&](int i) {
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k) {
        const double p[3] = {double(k) / N, double(j) / N, double(i) / N};
        double best = std::numeric_limits<double>::infinity();
        int bi = -1;
        for (int t = 0; t < nF; ++t) {
          const double d = point_triangle_sqr(p, Vn + 3 * F[3 * t], Vn + 3 * F[3 * t + 1], Vn + 3 * F[3 * t + 2], nullptr);
          if (d < best) { best = d; bi = t; }
        }
        const size_t o = ((size_t)i * N + j) * N + k;
        grid[o] = std::sqrt(best);
        if (idx) idx[o] = bi;
      }
  });
}

// Same result as orc_build_grid_brute, found with a uniform cell index and a
// ring search that stops only when no unvisited cell can hold a closer
// triangle (exact).  Kept as a second, independent accelerated builder for the tests; the CPU
// baseline uses the bounding-box tree below (orc_build_grid_bvh).
void orc_build_grid_fast(const double* Vn, int nV, const int* F, int nF, int N, int z0, int z1, double* grid, int* idx,
                         int nthreads) {
  (void)nV;
  // cell grid over the bounding box of the triangles and the unit cube
  double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  for (int t = 0; t < nF; ++t)
    for (int c = 0; c < 3; ++c)
      for (int j = 0; j < 3; ++j) {
        const double x = Vn[3 * F[3 * t + c] + j];
        lo[j] = std::min(lo[j], x); hi[j] = std::max(hi[j], x);
      }
  int C = (int)std::lround(std::cbrt((double)std::max(nF, 1) / 2.0));
  C = std::max(1, std::min(C, 128));
  double cs = 0;
  for (int j = 0; j < 3; ++j) cs = std::max(cs, (hi[j] - lo[j]) / C);
  cs *= 1.0000001;
  auto cell_of = [&](double x, int j) { int c = (int)std::floor((x - lo[j]) / cs); return std::max(0, std::min(C - 1, c)); };
  std::vector<int> start((size_t)C * C * C + 1, 0);
  std::vector<int> tri_lo(6 * (size_t)nF);
  for (int t = 0; t < nF; ++t) {
    for (int j = 0; j < 3; ++j) {
      double mn = 1e300, mx = -1e300;
      for (int c = 0; c < 3; ++c) { const double x = Vn[3 * F[3 * t + c] + j]; mn = std::min(mn, x); mx = std::max(mx, x); }
      tri_lo[6 * t + j] = cell_of(mn, j); tri_lo[6 * t + 3 + j] = cell_of(mx, j);
    }
    for (int z = tri_lo[6 * t + 2]; z <= tri_lo[6 * t + 5]; ++z)
      for (int y = tri_lo[6 * t + 1]; y <= tri_lo[6 * t + 4]; ++y)
        for (int x = tri_lo[6 * t]; x <= tri_lo[6 * t + 3]; ++x) start[((size_t)z * C + y) * C + x + 1]++;
  }
  for (size_t i = 0; i < (size_t)C * C * C; ++i) start[i + 1] += start[i];
  std::vector<int> items(start.back());
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int t = 0; t < nF; ++t)
      for (int z = tri_lo[6 * t + 2]; z <= tri_lo[6 * t + 5]; ++z)
        for (int y = tri_lo[6 * t + 1]; y <= tri_lo[6 * t + 4]; ++y)
          for (int x = tri_lo[6 * t]; x <= tri_lo[6 * t + 3]; ++x) items[fill[((size_t)z * C + y) * C + x]++] = t;
  }
  parallel_slices(z0, z1, nthreads, [&](int i) {
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k) {
        const double p[3] = {double(k) / N, double(j) / N, double(i) / N};
        const int cx = cell_of(p[0], 0), cy = cell_of(p[1], 1), cz = cell_of(p[2], 2);
        double best = std::numeric_limits<double>::infinity();
        int bi = -1;
        for (int r = 0; r <= C; ++r) {
          // all cells at Chebyshev distance exactly r
          for (int z = cz - r; z <= cz + r; ++z) {
            if (z < 0 || z >= C) continue;
            for (int y = cy - r; y <= cy + r; ++y) {
              if (y < 0 || y >= C) continue;
              const bool shell_zy = (z == cz - r || z == cz + r || y == cy - r || y == cy + r);
              const int step = shell_zy ? 1 : 2 * r;
              for (int x = cx - r; x <= cx + r; x += (step > 0 ? step : 1)) {
                if (x < 0 || x >= C) continue;
                const size_t c = ((size_t)z * C + y) * C + x;
                for (int s = start[c]; s < start[c + 1]; ++s) {
                  const int t = items[s];
                  const double d = point_triangle_sqr(p, Vn + 3 * F[3 * t], Vn + 3 * F[3 * t + 1], Vn + 3 * F[3 * t + 2], nullptr);
                  if (d < best || (d == best && t < bi)) { best = d; bi = t; }
                }
              }
            }
          }
          // every unvisited triangle has its closest point in a cell at
          // Chebyshev distance >= r+1, i.e. at Euclidean distance >= r*cs
          const double lb = r * cs;
          if (bi >= 0 && best <= lb * lb) break;
        }
        const size_t o = ((size_t)i * N + j) * N + k;
        grid[o] = std::sqrt(best);
        if (idx) idx[o] = bi;
      }
  });
}

// Same result again, found the way libigl finds it (igl::AABB<...,3>::squared_distance behind
// igl::point_mesh_squared_distance, src/lib/mesh.cc:140; libigl is un-vendored, restated from its
// published algorithm): a bounding-box tree over the triangles, split at the median centroid of the
// longest axis, queried depth first with the nearer child first and a subtree skipped when the
// squared distance from the query point to its box exceeds the best distance found so far.  This is
// the CPU baseline's builder: O(log M)-ish work per voxel wherever the voxel lies, where the ring
// search above degenerates to most of the mesh for voxels far from the surface.
// Exactness: a subtree is skipped only if its box bound exceeds the running best by more than the
// rounding of either number can explain (relative 1e-9), so every triangle that could win or tie
// is evaluated with point_triangle_sqr; ties go to the lowest index as in the brute-force loop.
namespace {
struct BvhNode { double lo[3], hi[3]; int left, right, start, count; };
struct Bvh {
  std::vector<BvhNode> nodes;
  std::vector<int> order;   // triangle indices, leaf ranges contiguous
};
int bvh_build(Bvh& T, const double* cen, const double* tlo, const double* thi, int start, int count) {
  const int id = (int)T.nodes.size();
  T.nodes.emplace_back();
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
  for (int i = start; i < start + count; ++i) {
    const int t = T.order[i];
    for (int j = 0; j < 3; ++j) {
      lo[j] = std::min(lo[j], tlo[3 * t + j]); hi[j] = std::max(hi[j], thi[3 * t + j]);
      clo[j] = std::min(clo[j], cen[3 * t + j]); chi[j] = std::max(chi[j], cen[3 * t + j]);
    }
  }
  int axis = 0;
  for (int j = 1; j < 3; ++j) if (chi[j] - clo[j] > chi[axis] - clo[axis]) axis = j;
  int left = -1, right = -1;
  if (count > 4 && chi[axis] > clo[axis]) {
    const int mid = start + count / 2;
    std::nth_element(T.order.begin() + start, T.order.begin() + mid, T.order.begin() + start + count,
                     [&](int a, int b) { return cen[3 * a + axis] < cen[3 * b + axis] || (cen[3 * a + axis] == cen[3 * b + axis] && a < b); });
    left = bvh_build(T, cen, tlo, thi, start, mid - start);
    right = bvh_build(T, cen, tlo, thi, mid, start + count - mid);
  }
  BvhNode& n = T.nodes[id];
  for (int j = 0; j < 3; ++j) { n.lo[j] = lo[j]; n.hi[j] = hi[j]; }
  n.left = left; n.right = right; n.start = start; n.count = count;
  return id;
}
inline double box_dist2(const BvhNode& n, const double* p) {
  double d2 = 0;
  for (int j = 0; j < 3; ++j) {
    const double g = std::max(0.0, std::max(n.lo[j] - p[j], p[j] - n.hi[j]));
    d2 += g * g;
  }
  return d2;
}
}  // namespace

void orc_build_grid_bvh(const double* Vn, int nV, const int* F, int nF, int N, int z0, int z1, double* grid, int* idx,
                        int nthreads) {
  (void)nV;
  Bvh T;
  std::vector<double> cen(3 * (size_t)std::max(nF, 1)), tlo(3 * (size_t)std::max(nF, 1)), thi(3 * (size_t)std::max(nF, 1));
  T.order.resize((size_t)nF);
  for (int t = 0; t < nF; ++t) {
    T.order[t] = t;
    for (int j = 0; j < 3; ++j) {
      const double a = Vn[3 * F[3 * t] + j], b = Vn[3 * F[3 * t + 1] + j], c = Vn[3 * F[3 * t + 2] + j];
      tlo[3 * t + j] = std::min(a, std::min(b, c)); thi[3 * t + j] = std::max(a, std::max(b, c));
      cen[3 * t + j] = (a + b + c) / 3.0;
    }
  }
  if (nF > 0) { T.nodes.reserve(2 * (size_t)nF); bvh_build(T, cen.data(), tlo.data(), thi.data(), 0, nF); }
  parallel_slices(z0, z1, nthreads, [&](int i) {
    std::vector<int> stack;
    stack.reserve(128);
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k) {
        const double p[3] = {double(k) / N, double(j) / N, double(i) / N};
        double best = std::numeric_limits<double>::infinity();
        int bi = -1;
        stack.clear();
        if (nF > 0) stack.push_back(0);
        while (!stack.empty()) {
          const BvhNode& n = T.nodes[stack.back()];
          stack.pop_back();
          if (box_dist2(n, p) > best * (1.0 + 1e-9) + 1e-300) continue;
          if (n.left < 0) {
            for (int s2 = n.start; s2 < n.start + n.count; ++s2) {
              const int t = T.order[s2];
              const double d = point_triangle_sqr(p, Vn + 3 * F[3 * t], Vn + 3 * F[3 * t + 1], Vn + 3 * F[3 * t + 2], nullptr);
              if (d < best || (d == best && t < bi)) { best = d; bi = t; }
            }
          } else {
            const double dl = box_dist2(T.nodes[n.left], p), dr = box_dist2(T.nodes[n.right], p);
            if (dl <= dr) { stack.push_back(n.right); stack.push_back(n.left); }   // nearer child on top
            else { stack.push_back(n.left); stack.push_back(n.right); }
          }
        }
        const size_t o = ((size_t)i * N + j) * N + k;
        grid[o] = std::sqrt(best);
        if (idx) idx[o] = bi;
      }
  });
}

// ---- samplers --------------------------------------------------------------
void orc_distance_float(const double* grid, int N, const float* P, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = sample_grid<float>(grid, N, P + 3 * i);
}
void orc_distance_double(const double* grid, int N, const double* P, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = sample_grid<double>(grid, N, P + 3 * i);
}
// value + 3 partials, the way ceres evaluates DistanceLoss (src/lib/distanceloss.h:11-16)
void orc_distance_double_jet(const double* grid, int N, const double* P, int n, double* val, double* grad) {
  typedef Jet<double, 3> J;
  for (int i = 0; i < n; ++i) {
    J p[3] = {J(P[3 * i], 0), J(P[3 * i + 1], 1), J(P[3 * i + 2], 2)};
    const J r = sample_grid<J>(grid, N, p);
    val[i] = r.a; grad[3 * i] = r.v[0]; grad[3 * i + 1] = r.v[1]; grad[3 * i + 2] = r.v[2];
  }
}
void orc_distance_float_jet(const double* grid, int N, const float* P, int n, float* val, float* grad) {
  typedef Jet<float, 3> J;
  for (int i = 0; i < n; ++i) {
    J p[3] = {J(P[3 * i], 0), J(P[3 * i + 1], 1), J(P[3 * i + 2], 2)};
    const J r = sample_grid<J>(grid, N, p);
    val[i] = r.a; grad[3 * i] = r.v[0]; grad[3 * i + 1] = r.v[1]; grad[3 * i + 2] = r.v[2];
  }
}

// DistanceFieldLoss_forward (src/interface/distance_layer.cc:27-34)
void orc_distfield_forward(const double* grid, int N, const float* V, int n, float* out) {
  for (int i = 0; i < n; ++i) {
    out[i] = sample_grid<float>(grid, N, V + 3 * i);
    out[i] *= out[i];
  }
}
// DistanceFieldLoss_backward (src/interface/distance_layer.cc:58-78)
void orc_distfield_backward(const double* grid, int N, const float* V, int n, float* out) {
  typedef Jet<float, 3> J;
  for (int i = 0; i < n; ++i) {
    const float* v = V + 3 * i;
    J p[3] = {J(v[0], 0), J(v[1], 1), J(v[2], 2)};
    J vd = sample_grid<J>(grid, N, p);
    vd = vd * vd;
    out[3 * i] = vd.v[0] * 0.5;
    out[3 * i + 1] = vd.v[1] * 0.5;
    out[3 * i + 2] = vd.v[2] * 0.5;
  }
}

// ---- rigid edges (src/interface/rigid_layer.cc) ----------------------------
void orc_store_rigid(const float* V, const int* F, int nF, float* rest) {   // :32-45
  int offset = 0;
  for (int i = 0; i < nF; ++i)
    for (int j = 0; j < 3; ++j) {
      const int v0 = F[i * 3 + j], v1 = F[i * 3 + (j + 1) % 3];
      for (int k = 0; k < 3; ++k) rest[offset * 3 + k] = V[v1 * 3 + k] - V[v0 * 3 + k];
      offset += 1;
    }
}
void orc_rigid_forward(const float* V, const int* F, int nF, const float* rest, float* out) {   // :72-86
  for (int i = 0; i < nF; ++i)
    for (int j = 0; j < 3; ++j) {
      const int v0 = F[i * 3 + j], v1 = F[i * 3 + (j + 1) % 3];
      float* l = out + (i * 3 + j) * 3;
      for (int k = 0; k < 3; ++k) {
        l[k] = (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[(i * 3 + j) * 3 + k]);
        l[k] *= l[k];
      }
    }
}
void orc_rigid_backward(const float* V, int nV, const int* F, int nF, const float* rest, float* out) {   // :113-130
  std::memset(out, 0, sizeof(float) * 3 * (size_t)nV);
  for (int i = 0; i < nF; ++i)
    for (int j = 0; j < 3; ++j) {
      const int v0 = F[i * 3 + j], v1 = F[i * 3 + (j + 1) % 3];
      for (int k = 0; k < 3; ++k) {
        out[v0 * 3 + k] -= (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[(i * 3 + j) * 3 + k]);
        out[v1 * 3 + k] += (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[(i * 3 + j) * 3 + k]);
      }
    }
}

// ---- graph edges (src/interface/graph_layer.cc) ----------------------------
void orc_store_graph(const float* V, const int* E, int nE, float* rest) {   // :32-44
  for (int i = 0; i < nE; ++i) {
    const int v0 = E[i * 2], v1 = E[i * 2 + 1];
    for (int k = 0; k < 3; ++k) rest[i * 3 + k] = V[v1 * 3 + k] - V[v0 * 3 + k];
  }
}
void orc_graph_forward(const float* V, const int* E, int nE, const float* rest, float* out) {   // :74-88
  for (int i = 0; i < nE; ++i) {
    const int v0 = E[i * 2], v1 = E[i * 2 + 1];
    for (int k = 0; k < 3; ++k) {
      float l = (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[i * 3 + k]);
      out[i * 3 + k] = l * l;
    }
  }
}
void orc_graph_backward(const float* V, int nV, const int* E, int nE, const float* rest, float* out) {   // :116-132
  std::memset(out, 0, sizeof(float) * 3 * (size_t)nV);
  for (int i = 0; i < nE; ++i) {
    const int v0 = E[i * 2], v1 = E[i * 2 + 1];
    for (int k = 0; k < 3; ++k) {
      out[v0 * 3 + k] -= (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[i * 3 + k]);
      out[v1 * 3 + k] += (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[i * 3 + k]);
    }
  }
}

// ---- CAD edges (src/interface/cad_layer.cc): E rows then 3F face edges -----
static inline void cad_edge(int idx, int nE, const int* F, const int* E, int* v0, int* v1) {
  if (idx < nE) { *v0 = E[idx * 2]; *v1 = E[idx * 2 + 1]; }
  else { const int f = (idx - nE) / 3, j = (idx - nE) % 3; *v0 = F[f * 3 + j]; *v1 = F[f * 3 + (j + 1) % 3]; }
}
void orc_store_cad(const float* V, const int* F, int nF, const int* E, int nE, float* rest, float* lambda) {   // :37-76
  for (int o = 0; o < nE + 3 * nF; ++o) {
    int v0, v1; cad_edge(o, nE, F, E, &v0, &v1);
    float r[3];
    for (int k = 0; k < 3; ++k) { r[k] = V[v1 * 3 + k] - V[v0 * 3 + k]; rest[o * 3 + k] = r[k]; }
    const float norm = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);   // Eigen Vector3f::norm()
    lambda[o] = (2e-2 / (norm + 1e-8));                                     // :48-49 (double arithmetic, stored float)
  }
}
void orc_cad_forward(const float* V, const int* F, int nF, const int* E, int nE, const float* rest, const float* lambda,
                     float* out) {   // :109-147
  for (int o = 0; o < nE + 3 * nF; ++o) {
    int v0, v1; cad_edge(o, nE, F, E, &v0, &v1);
    for (int k = 0; k < 3; ++k) {
      float l = (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[o * 3 + k]) * lambda[o];
      out[o * 3 + k] = l * l;
    }
  }
}
void orc_cad_backward(const float* V, int nV, const int* F, int nF, const int* E, int nE, const float* rest,
                      const float* lambda, float* out) {   // :178-220
  std::memset(out, 0, sizeof(float) * 3 * (size_t)nV);
  for (int o = 0; o < nE + 3 * nF; ++o) {
    int v0, v1; cad_edge(o, nE, F, E, &v0, &v1);
    float lam = lambda[o];
    lam *= lam;
    for (int k = 0; k < 3; ++k) {
      out[v0 * 3 + k] -= (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[o * 3 + k]) * lam;
      out[v1 * 3 + k] += (V[v1 * 3 + k] - V[v0 * 3 + k] - rest[o * 3 + k]) * lam;
    }
  }
}

// ---- Normalize/DenormalizeByTemplate (src/interface/normalize.cc:19-23, :40-44)
void orc_normalize_by_template(float* V, int n, double scale, const double* trans) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) V[i * 3 + j] = (V[i * 3 + j] - trans[j]) / scale;
}
void orc_denormalize_by_template(float* V, int n, double scale, const double* trans) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) V[i * 3 + j] = V[i * 3 + j] * scale + trans[j];
}

// ---- Ceres functors (src/lib/edgeloss.h, src/lib/distanceloss.h) -----------
// EdgeLoss (:8-33): r = (p1 - p2 - v) * lambda ; d r/d p1 = lambda I, d r/d p2 = -lambda I.
// AdaptiveEdgeLoss (:35-62): lambda <- lambda * (2e-2 / (|v| + 1e-8)) at construction.
void orc_edge_loss(const double* p1, const double* p2, const double* v, double lambda, int adaptive, double* residual,
                   double* lambda_eff) {
  if (adaptive) lambda = lambda * (2e-2 / (std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-8));
  for (int k = 0; k < 3; ++k) residual[k] = ((p1[k] - p2[k]) - v[k]) * lambda;
  if (lambda_eff) *lambda_eff = lambda;
}
// EdgeLossWithRot (:64-98): 6 residuals and the 6x12 Jacobian w.r.t.
// (p1, p2, rot1, rot2) obtained exactly as ceres::AutoDiffCostFunction does (Jet<double,12>).
void orc_edge_rot(const double* p1, const double* p2, const double* rot1, const double* rot2, const double* v,
                  double lambda, double* residual, double* jac /* [6][12] row-major */) {
  typedef Jet<double, 12> J;
  J a[3], b[3], r1[3], r2[3], res[6];
  for (int k = 0; k < 3; ++k) { a[k] = J(p1[k], k); b[k] = J(p2[k], 3 + k); r1[k] = J(rot1[k], 6 + k); r2[k] = J(rot2[k], 9 + k); }
  edge_rot_functor<J>(a, b, r1, r2, v, lambda, res);
  for (int i = 0; i < 6; ++i) {
    residual[i] = res[i].a;
    if (jac) for (int j = 0; j < 12; ++j) jac[i * 12 + j] = res[i].v[j];
  }
}

// Cost 0.5*sum r^2 and its gradient for the DeformWithRot problem
// (src/lib/deformer.cc:94-171: V DistanceLoss blocks + 3F EdgeLossWithRot
// blocks with v = V0[F[i][j]] - V0[F[i][(j+1)%3]], rot initialised to 0).
// Accumulated in FP64 in block order.
double orc_rot_problem_cost_grad(const double* grid, int N, const double* V, const double* R, int nV, const int* F,
                                 int nF, const double* rest /* [3F,3] */, double lambda, double* gV, double* gR,
                                 double* cost_dist, double* cost_edge) {
  std::vector<double> tmpV(3 * (size_t)nV, 0.0), tmpR(3 * (size_t)nV, 0.0);
  double cd = 0, ce = 0;
  typedef Jet<double, 3> J3;
  for (int i = 0; i < nV; ++i) {
    J3 p[3] = {J3(V[3 * i], 0), J3(V[3 * i + 1], 1), J3(V[3 * i + 2], 2)};
    const J3 r = sample_grid<J3>(grid, N, p);
    cd += 0.5 * r.a * r.a;
    for (int k = 0; k < 3; ++k) tmpV[3 * i + k] += r.a * r.v[k];
  }
  for (int i = 0; i < nF; ++i)
    for (int j = 0; j < 3; ++j) {
      const int va = F[i * 3 + j], vb = F[i * 3 + (j + 1) % 3];
      double res[6], jac[72];
      orc_edge_rot(V + 3 * va, V + 3 * vb, R + 3 * va, R + 3 * vb, rest + 3 * (i * 3 + j), lambda, res, jac);
      for (int m = 0; m < 6; ++m) {
        ce += 0.5 * res[m] * res[m];
        for (int k = 0; k < 3; ++k) {
          tmpV[3 * va + k] += res[m] * jac[m * 12 + k];
          tmpV[3 * vb + k] += res[m] * jac[m * 12 + 3 + k];
          tmpR[3 * va + k] += res[m] * jac[m * 12 + 6 + k];
          tmpR[3 * vb + k] += res[m] * jac[m * 12 + 9 + k];
        }
      }
    }
  if (gV) std::memcpy(gV, tmpV.data(), sizeof(double) * tmpV.size());
  if (gR) std::memcpy(gR, tmpR.data(), sizeof(double) * tmpR.size());
  if (cost_dist) *cost_dist = cd;
  if (cost_edge) *cost_edge = ce;
  return cd + ce;
}

// Cost/gradient of the Deformer::Deform problem (src/lib/deformer.cc:32-53):
// V DistanceLoss blocks + 3F EdgeLoss blocks, v = V0[F[i][j]] - V0[F[i][(j+1)%3]],
// residual (p1 - p2 - v) * lambda with p1 = V[F[i][j]], p2 = V[F[i][(j+1)%3]].
double orc_deform_problem_cost_grad(const double* grid, int N, const double* V, int nV, const int* F, int nF,
                                    const double* rest, double lambda, int adaptive, double* gV, double* cost_dist,
                                    double* cost_edge) {
  std::vector<double> tmpV(3 * (size_t)nV, 0.0);
  double cd = 0, ce = 0;
  typedef Jet<double, 3> J3;
  for (int i = 0; i < nV; ++i) {
    J3 p[3] = {J3(V[3 * i], 0), J3(V[3 * i + 1], 1), J3(V[3 * i + 2], 2)};
    const J3 r = sample_grid<J3>(grid, N, p);
    cd += 0.5 * r.a * r.a;
    for (int k = 0; k < 3; ++k) tmpV[3 * i + k] += r.a * r.v[k];
  }
  for (int i = 0; i < nF; ++i)
    for (int j = 0; j < 3; ++j) {
      const int va = F[i * 3 + j], vb = F[i * 3 + (j + 1) % 3];
      double res[3], lam;
      orc_edge_loss(V + 3 * va, V + 3 * vb, rest + 3 * (i * 3 + j), lambda, adaptive, res, &lam);
      for (int k = 0; k < 3; ++k) {
        ce += 0.5 * res[k] * res[k];
        tmpV[3 * va + k] += res[k] * lam;
        tmpV[3 * vb + k] -= res[k] * lam;
      }
    }
  if (gV) std::memcpy(gV, tmpV.data(), sizeof(double) * tmpV.size());
  if (cost_dist) *cost_dist = cd;
  if (cost_edge) *cost_edge = ce;
  return cd + ce;
}

// ---- the Python rigid_deform loop (src/python/rigid_deform.py:32-41) -------
// loss = 0.5*sum(dist_fwd) + 0.5*sum(rigid_fwd)   (rigid_loss_layer.py:11-17)
// grad = dist_bwd + rigid_bwd                      (rigid_loss_layer.py:24-27)
// followed by torch.optim.Adam(lr) with default betas (0.9, 0.999), eps 1e-8,
// restated from torch/optim/adam.py::_single_tensor_adam in float32.
// V is the *normalised* source (RigidLossLayer.__init__ normalises in place).
// loss_log (optional) receives the loss of every log_every-th iteration.
void orc_rigid_adam(const double* grid, int N, float* V, int nV, const int* F, int nF, const float* rest, int iters,
                    double lr, double* loss_log, int log_every) {
  const size_t n3 = 3 * (size_t)nV;
  std::vector<float> gD(n3), gR(n3), m(n3, 0.f), v(n3, 0.f), fD, fR;
  const double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
  if (loss_log) { fD.resize(nV); fR.resize(9 * (size_t)nF); }
  for (int it = 0; it < iters; ++it) {
    if (loss_log && log_every > 0 && it % log_every == 0) {
      orc_distfield_forward(grid, N, V, nV, fD.data());
      orc_rigid_forward(V, F, nF, rest, fR.data());
      double s = 0;
      for (float x : fD) s += 0.5 * (double)x;
      for (float x : fR) s += 0.5 * (double)x;
      loss_log[it / log_every] = s;
    }
    orc_distfield_backward(grid, N, V, nV, gD.data());
    orc_rigid_backward(V, nV, F, nF, rest, gR.data());
    const int step = it + 1;
    const double bc1 = 1.0 - std::pow(beta1, step);
    const double bc2 = 1.0 - std::pow(beta2, step);
    const float step_size = (float)(lr / bc1);
    const float bc2_sqrt = (float)std::sqrt(bc2);
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), epsf = (float)eps;
    for (size_t i = 0; i < n3; ++i) {
      const float g = (gD[i] + gR[i]) * 1.0f;
      // Operation order and FMA use follow what torch's own kernels compute (checked against
      // torch 2.11 CPU in tests/test_oracle_kat.py): lerp is fma(weight, end - start, start),
      // addcmul is fma(value * t1, t2, self), addcdiv is self + (value * t1) / t2.
      m[i] = std::fmaf(w1, g - m[i], m[i]);              // exp_avg.lerp_(grad, 1-beta1)
      v[i] = std::fmaf(w2 * g, g, v[i] * b2);            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
      const float denom = std::sqrt(v[i]) / bc2_sqrt + epsf;
      V[i] = V[i] + ((-step_size) * m[i]) / denom;       // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
  }
}

// One whole pair, the way rigid_deform.py runs it: template from the target
// (InitializeDeformTemplate: normalise + grid build), source normalised by the
// template, rest edges stored, Adam iterations, denormalise.  srcV in/out.
void orc_deform_pair(const float* tarV, int nTv, const int* tarF, int nTf, float* srcV, int nSv, const int* srcF,
                     int nSf, int N, int iters, double lr, int build_threads) {
  std::vector<double> Vn(3 * (size_t)nTv), grid((size_t)N * N * N);
  double scale, pos[3];
  orc_normalize_target(tarV, nTv, Vn.data(), &scale, pos);
  orc_build_grid_bvh(Vn.data(), nTv, tarF, nTf, N, 0, N, grid.data(), nullptr, build_threads);
  orc_normalize_by_template(srcV, nSv, scale, pos);
  std::vector<float> rest(9 * (size_t)nSf);
  orc_store_rigid(srcV, srcF, nSf, rest.data());
  orc_rigid_adam(grid.data(), N, srcV, nSv, srcF, nSf, rest.data(), iters, lr, nullptr, 0);
  orc_denormalize_by_template(srcV, nSv, scale, pos);
}

int orc_abi_version() { return 1; }

}  // extern "C"
