// ref_shim.cc -- C entry points around the REFERENCE's own sources, compiled where they lie
// under /root/reference (never copied): src/lib/uniformgrid.{h,cc}, src/lib/distanceloss.h,
// src/lib/edgeloss.h, against the stand-in third-party headers in oracle/stubs/.
//
// TEST INFRASTRUCTURE ONLY (output: oracle/_ref/libmeshode_ref.so, git-ignored).  It is used
// to pin oracle/meshode_oracle.cc (tests/test_oracle_ref.py) and to generate
// tests/golden/golden_ref.npz (tests/golden/make_golden_ref.py).  Nothing in meshode_b200/
// loads it.
#include <cstddef>

#include <ceres/ceres.h>   // before the loss headers, as src/lib/deformer.cc:3-7 orders them

#include "distanceloss.h"   // reference: DistanceLoss (src/lib/distanceloss.h:6-25)
#include "edgeloss.h"       // reference: EdgeLoss / AdaptiveEdgeLoss / EdgeLossWithRot (src/lib/edgeloss.h)
#include "uniformgrid.h"    // reference: UniformGrid (src/lib/uniformgrid.h:8-34)

extern "C" {

// UniformGrid(N) + SetDistance(i=z, j=y, k=x) exactly as Mesh::ConstructDistanceField fills it
// (src/lib/mesh.cc:142-150); grid is [z][y][x] FP64.
void* ref_grid_create(int N, const double* grid) {
  UniformGrid* g = new UniformGrid(N);
  if (grid) {
    size_t o = 0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < N; ++k) g->SetDistance(i, j, k, grid[o++]);
  }
  return g;
}
void ref_grid_destroy(void* g) { delete static_cast<UniformGrid*>(g); }
int ref_grid_dimension(void* g) { return static_cast<UniformGrid*>(g)->Dimension(); }
double ref_grid_get(void* g, int i, int j, int k) { return static_cast<UniformGrid*>(g)->GetDistance(i, j, k); }

void ref_distance_double(void* g, const double* P, int n, double* out) {
  const UniformGrid* G = static_cast<UniformGrid*>(g);
  for (int i = 0; i < n; ++i) out[i] = G->distance<double>(P + 3 * (size_t)i);
}
void ref_distance_float(void* g, const float* P, int n, float* out) {
  const UniformGrid* G = static_cast<UniformGrid*>(g);
  for (int i = 0; i < n; ++i) out[i] = G->DistanceFloat<float>(P + 3 * (size_t)i);
}
void ref_distance_double_jet(void* g, const double* P, int n, double* val, double* grad) {
  typedef ceres::Jet<double, 3> J;
  const UniformGrid* G = static_cast<UniformGrid*>(g);
  for (int i = 0; i < n; ++i) {
    J p[3] = {J(P[3 * (size_t)i], 0), J(P[3 * (size_t)i + 1], 1), J(P[3 * (size_t)i + 2], 2)};
    const J r = G->distance<J>(p);
    val[i] = r.a;
    for (int k = 0; k < 3; ++k) grad[3 * (size_t)i + k] = r.v[k];
  }
}
void ref_distance_float_jet(void* g, const float* P, int n, float* val, float* grad) {
  typedef ceres::Jet<float, 3> J;
  const UniformGrid* G = static_cast<UniformGrid*>(g);
  for (int i = 0; i < n; ++i) {
    J p[3] = {J(P[3 * (size_t)i], 0), J(P[3 * (size_t)i + 1], 1), J(P[3 * (size_t)i + 2], 2)};
    const J r = G->DistanceFloat<J>(p);
    val[i] = r.a;
    for (int k = 0; k < 3; ++k) grad[3 * (size_t)i + k] = r.v[k];
  }
}

// DistanceLoss functor evaluated the way AutoDiffCostFunction<DistanceLoss,3,3> does:
// residuals[3] and the 3x3 Jacobian (row-major).
void ref_distance_loss(void* g, const double* p, double* residuals, double* jac) {
  typedef ceres::Jet<double, 3> J;
  DistanceLoss f(static_cast<UniformGrid*>(g));
  J x[3] = {J(p[0], 0), J(p[1], 1), J(p[2], 2)}, r[3];
  f(x, r);
  for (int i = 0; i < 3; ++i) {
    residuals[i] = r[i].a;
    for (int k = 0; k < 3; ++k) jac[3 * i + k] = r[i].v[k];
  }
}

// EdgeLoss / AdaptiveEdgeLoss in double; lambda_eff returns the functor's stored weight.
void ref_edge_loss(const double* p1, const double* p2, const double* v, double lambda, int adaptive, double* residuals,
                   double* lambda_eff) {
  const Vector3 vv(v[0], v[1], v[2]);
  if (adaptive) {
    AdaptiveEdgeLoss f(vv, lambda);
    f(p1, p2, residuals);
    *lambda_eff = f.lambda;
  } else {
    EdgeLoss f(vv, lambda);
    f(p1, p2, residuals);
    *lambda_eff = f.lambda;
  }
}

// EdgeLossWithRot: 6 residuals + 6x12 Jacobian w.r.t. (p1, p2, rot1, rot2), Jet<double,12>.
void ref_edge_rot(const double* p1, const double* p2, const double* rot1, const double* rot2, const double* v,
                  double lambda, double* residuals, double* jac) {
  typedef ceres::Jet<double, 12> J;
  EdgeLossWithRot f(Vector3(v[0], v[1], v[2]), lambda);
  J a[3], b[3], c[3], d[3], r[6];
  for (int k = 0; k < 3; ++k) { a[k] = J(p1[k], k); b[k] = J(p2[k], 3 + k); c[k] = J(rot1[k], 6 + k); d[k] = J(rot2[k], 9 + k); }
  f(a, b, c, d, r);
  for (int i = 0; i < 6; ++i) {
    residuals[i] = r[i].a;
    for (int k = 0; k < 12; ++k) jac[12 * i + k] = r[i].v[k];
  }
}

}  // extern "C"
