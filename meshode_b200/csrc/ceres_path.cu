// The loss terms of the C++ drivers' Ceres problems (reference src/lib/deformer.cc), FP64:
//   DistanceLoss            src/lib/distanceloss.h:6-25   residual [UniformGrid::distance<double>(p), 0, 0]
//   EdgeLoss                src/lib/edgeloss.h:8-33       (p1 - p2 - v) * lambda
//   AdaptiveEdgeLoss        src/lib/edgeloss.h:35-62      lambda <- lambda * 2e-2 / (|v| + 1e-8)
//   EdgeLossWithRot         src/lib/edgeloss.h:64-98      (AngleAxisRotatePoint(rot1, p1 - p2) - v) * lambda, rot1 - rot2
// evaluated per residual block (residuals + the Jacobian ceres::AutoDiffCostFunction would produce)
// and as whole problems (cost = 0.5 * sum r^2 and gradient J^T r, what Problem::Evaluate returns).
//
// Derivatives are forward-mode dual numbers with the operator definitions of ceres::Jet on
// contraction-free FP64, so every partial equals the reference's autodiff value up to the last-bit
// differences of sin/cos between CUDA's and the host's libm.
#include <cfloat>

#include "common.cuh"
#include "sampler.cuh"

namespace mo {
namespace {

constexpr int kBlock = 128;

template <int N>
struct JetD {
  double a;
  double v[N];
  __device__ __forceinline__ JetD() {}
  __device__ __forceinline__ explicit JetD(double s) : a(s) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  __device__ __forceinline__ JetD(double s, int k) : a(s) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = (i == k) ? 1.0 : 0.0;
  }
};
#define MO_JET_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)
template <int N> __device__ __forceinline__ JetD<N> operator+(const JetD<N>& f, const JetD<N>& g) {
  JetD<N> h; h.a = dadd(f.a, g.a); MO_JET_LOOP h.v[i] = dadd(f.v[i], g.v[i]); return h; }
template <int N> __device__ __forceinline__ JetD<N> operator-(const JetD<N>& f, const JetD<N>& g) {
  JetD<N> h; h.a = dsub(f.a, g.a); MO_JET_LOOP h.v[i] = dsub(f.v[i], g.v[i]); return h; }
template <int N> __device__ __forceinline__ JetD<N> operator*(const JetD<N>& f, const JetD<N>& g) {
  JetD<N> h; h.a = dmul(f.a, g.a); MO_JET_LOOP h.v[i] = dadd(dmul(f.a, g.v[i]), dmul(f.v[i], g.a)); return h; }
template <int N> __device__ __forceinline__ JetD<N> operator/(const JetD<N>& f, const JetD<N>& g) {
  JetD<N> h; const double gi = __ddiv_rn(1.0, g.a); const double fg = dmul(f.a, gi); h.a = fg;
  MO_JET_LOOP h.v[i] = dmul(dsub(f.v[i], dmul(fg, g.v[i])), gi); return h; }
template <int N> __device__ __forceinline__ JetD<N> jsqrt(const JetD<N>& f) {
  JetD<N> h; h.a = __dsqrt_rn(f.a); const double t = __ddiv_rn(1.0, dmul(2.0, h.a));
  MO_JET_LOOP h.v[i] = dmul(f.v[i], t); return h; }
template <int N> __device__ __forceinline__ JetD<N> jcos(const JetD<N>& f) {
  JetD<N> h; h.a = cos(f.a); const double m = -sin(f.a); MO_JET_LOOP h.v[i] = dmul(m, f.v[i]); return h; }
template <int N> __device__ __forceinline__ JetD<N> jsin(const JetD<N>& f) {
  JetD<N> h; h.a = sin(f.a); const double c = cos(f.a); MO_JET_LOOP h.v[i] = dmul(c, f.v[i]); return h; }
#undef MO_JET_LOOP

// ceres::AngleAxisRotatePoint (ceres/rotation.h) on dual numbers
template <int N>
__device__ __forceinline__ void angle_axis_rotate(const JetD<N> aa[3], const JetD<N> pt[3], JetD<N> out[3]) {
  typedef JetD<N> T;
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2.a > DBL_EPSILON) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    out[0] = pt[0] * costheta + wxp[0] * sintheta + w[0] * tmp;
    out[1] = pt[1] * costheta + wxp[1] * sintheta + w[1] * tmp;
    out[2] = pt[2] * costheta + wxp[2] * sintheta + w[2] * tmp;
  } else {   // first-order expansion near zero rotation (where DeformWithRot starts: rots = 0, deformer.cc:118)
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + wxp[0]; out[1] = pt[1] + wxp[1]; out[2] = pt[2] + wxp[2];
  }
}

// EdgeLossWithRot::operator() with partials w.r.t. (d = p1 - p2 : 0..2, rot1 : 3..5).
// d r/d p1 = d r/d d, d r/d p2 = -d r/d d (exact: the Jet of p1 - p2 has partials +1 / -1).
__device__ __forceinline__ void edge_rot_block(const double* p1, const double* p2, const double* rot1, const double* rot2,
                                               const double* v, const double lambda, double res[6], double Jd[3][3],
                                               double Jr[3][3]) {
  typedef JetD<6> T;
  T d[3], aa[3], q[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { d[k] = T(dsub(p1[k], p2[k]), k); aa[k] = T(rot1[k], 3 + k); }
  angle_axis_rotate<6>(aa, d, q);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    res[k] = dmul(dsub(q[k].a, v[k]), lambda);
    res[3 + k] = dmul(dsub(rot1[k], rot2[k]), 1.0);
#pragma unroll
    for (int c = 0; c < 3; ++c) { Jd[k][c] = dmul(q[k].v[c], lambda); Jr[k][c] = dmul(q[k].v[3 + c], lambda); }
  }
}

__device__ __forceinline__ double eff_lambda(const int kind, const double lambda, const double* v) {
  if (kind != MO_CERES_ADAPTIVE_EDGE) return lambda;
  const double n = __dsqrt_rn(dadd(dadd(dmul(v[0], v[0]), dmul(v[1], v[1])), dmul(v[2], v[2])));   // Vector3::norm()
  return dmul(lambda, __ddiv_rn(2e-2, dadd(n, 1e-8)));                                             // edgeloss.h:38
}

__global__ void __launch_bounds__(kBlock) k_ceres_edges(const int kind, const double* __restrict__ V,
                                                        const double* __restrict__ R, const int nV,
                                                        const int* __restrict__ I, const double* __restrict__ rest,
                                                        const int nE, const double lambda, double* __restrict__ res,
                                                        double* __restrict__ jac, double* __restrict__ cost,
                                                        double* __restrict__ gV, double* __restrict__ gR) {
  __shared__ double s_part[kBlock / 32];
  const int e = blockIdx.x * kBlock + threadIdx.x;
  double my = 0.0;
  if (e < nE) {
    const int ia = I[2 * (size_t)e], ib = I[2 * (size_t)e + 1];
    if ((unsigned)ia < (unsigned)nV && (unsigned)ib < (unsigned)nV) {
      const double* p1 = V + 3 * (size_t)ia;
      const double* p2 = V + 3 * (size_t)ib;
      const double v[3] = {rest[3 * (size_t)e], rest[3 * (size_t)e + 1], rest[3 * (size_t)e + 2]};
      if (kind == MO_CERES_ROT_EDGE) {
        double r[6], Jd[3][3], Jr[3][3];
        edge_rot_block(p1, p2, R + 3 * (size_t)ia, R + 3 * (size_t)ib, v, lambda, r, Jd, Jr);
        if (res) {
#pragma unroll
          for (int m = 0; m < 6; ++m) res[6 * (size_t)e + m] = r[m];
        }
        if (jac) {   // [6][12] row-major over (p1, p2, rot1, rot2), as AutoDiffCostFunction<.,6,3,3,3,3>
          double* J = jac + 72 * (size_t)e;
#pragma unroll
          for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              J[12 * m + c] = Jd[m][c]; J[12 * m + 3 + c] = -Jd[m][c]; J[12 * m + 6 + c] = Jr[m][c]; J[12 * m + 9 + c] = 0.0;
              J[12 * (3 + m) + c] = 0.0; J[12 * (3 + m) + 3 + c] = 0.0;
              J[12 * (3 + m) + 6 + c] = (m == c) ? 1.0 : 0.0; J[12 * (3 + m) + 9 + c] = (m == c) ? -1.0 : 0.0;
            }
          }
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) my += 0.5 * r[m] * r[m];
        if (gV) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double g = r[0] * Jd[0][c] + r[1] * Jd[1][c] + r[2] * Jd[2][c];
            atomicAdd(gV + 3 * (size_t)ia + c, g);
            atomicAdd(gV + 3 * (size_t)ib + c, -g);
          }
        }
        if (gR) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const double g = r[0] * Jr[0][c] + r[1] * Jr[1][c] + r[2] * Jr[2][c];
            atomicAdd(gR + 3 * (size_t)ia + c, g + r[3 + c]);
            atomicAdd(gR + 3 * (size_t)ib + c, -r[3 + c]);
          }
        }
      } else {
        const double lam = eff_lambda(kind, lambda, v);
        double r[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = dmul(dsub(dsub(p1[k], p2[k]), v[k]), lam);   // edgeloss.h:17-22
        if (res) {
#pragma unroll
          for (int k = 0; k < 3; ++k) res[3 * (size_t)e + k] = r[k];
        }
        if (jac) {   // [3][6]: lambda*I, -lambda*I
          double* J = jac + 18 * (size_t)e;
#pragma unroll
          for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { J[6 * m + c] = (m == c) ? lam : 0.0; J[6 * m + 3 + c] = (m == c) ? -lam : 0.0; }
          }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) my += 0.5 * r[k] * r[k];
        if (gV) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            atomicAdd(gV + 3 * (size_t)ia + c, r[c] * lam);
            atomicAdd(gV + 3 * (size_t)ib + c, -r[c] * lam);
          }
        }
      }
    }
  }
  if (cost) {
    for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < kBlock / 32; ++i) t += s_part[i];
      atomicAdd(cost, t);
    }
  }
}

// DistanceLoss blocks: cost += 0.5 d^2, gV[i] = d * grad d   (gV is overwritten: launched first)
__global__ void __launch_bounds__(kBlock) k_ceres_distance(const double* __restrict__ grid, const int n,
                                                           const double* __restrict__ V, const int nV,
                                                           double* __restrict__ cost, double* __restrict__ gV) {
  __shared__ double s_part[kBlock / 32];
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double my = 0.0;
  if (i < nV) {
    typedef Jet3<double> J;
    const J r = sample<J, double>(grid, n, J(V[3 * (size_t)i], 1.0, 0.0, 0.0), J(V[3 * (size_t)i + 1], 0.0, 1.0, 0.0),
                                  J(V[3 * (size_t)i + 2], 0.0, 0.0, 1.0));
    my = 0.5 * r.a * r.a;
    if (gV) { gV[3 * (size_t)i] = r.a * r.v0; gV[3 * (size_t)i + 1] = r.a * r.v1; gV[3 * (size_t)i + 2] = r.a * r.v2; }
  }
  if (cost) {
    for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int k = 0; k < kBlock / 32; ++k) t += s_part[k];
      atomicAdd(cost, t);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Levenberg-Marquardt for the problems above (what ceres::Solve does for Deformer::Deform /
// DeformWithRot / DeformSubdivision, src/lib/deformer.cc:55-74: trust-region LM, at most 100 iterations).
// The normal equations (J^T J + D^T D) delta = -g are never formed: J^T J v is applied block by block
// (distance block: gd (gd . v); edge block: lambda^2 (v_a - v_b), or J_d / J_r of EdgeLossWithRot) and solved
// by Jacobi-preconditioned conjugate gradients whose scalars stay on the device (one host read-back
// every kCgCheck iterations).  Ceres factorises the same matrix with a sparse Cholesky instead.
// ---------------------------------------------------------------------------------------------
constexpr int kCgCheck = 16;

// linearisation at x: distance blocks.  d[i], gd[i] = grad d, cost, g = d*gd (overwrites), diag = gd^2 (overwrites)
__global__ void __launch_bounds__(kBlock) k_lm_lin_distance(const double* __restrict__ grid, const int n,
                                                            const double* __restrict__ V, const int nV,
                                                            double* __restrict__ gd, double* __restrict__ cost,
                                                            double* __restrict__ g, double* __restrict__ diag) {
  __shared__ double s_part[kBlock / 32];
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double my = 0.0;
  if (i < nV) {
    double a = 0.0, v[3] = {0.0, 0.0, 0.0};
    if (grid) {
      typedef Jet3<double> J;
      const J r = sample<J, double>(grid, n, J(V[3 * (size_t)i], 1.0, 0.0, 0.0), J(V[3 * (size_t)i + 1], 0.0, 1.0, 0.0),
                                    J(V[3 * (size_t)i + 2], 0.0, 0.0, 1.0));
      a = r.a; v[0] = r.v0; v[1] = r.v1; v[2] = r.v2;
    }
    my = 0.5 * a * a;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (gd) gd[3 * (size_t)i + c] = v[c];
      if (g) g[3 * (size_t)i + c] = a * v[c];
      if (diag) diag[3 * (size_t)i + c] = v[c] * v[c];
    }
  }
  for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = my;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kBlock / 32; ++k) t += s_part[k];
    atomicAdd(cost, t);
  }
}

// linearisation at x: edge blocks.  ejac: per edge lambda_e (1 double) or J_d, J_r (18 doubles, ROT)
__global__ void __launch_bounds__(kBlock) k_lm_lin_edges(const int kind, const double* __restrict__ X, const int nV,
                                                         const int* __restrict__ I, const double* __restrict__ rest,
                                                         const int nE, const double lambda, double* __restrict__ ejac,
                                                         double* __restrict__ cost, double* __restrict__ g,
                                                         double* __restrict__ diag) {
  __shared__ double s_part[kBlock / 32];
  const int e = blockIdx.x * kBlock + threadIdx.x;
  const double* V = X;
  const double* R = X + 3 * (size_t)nV;
  double my = 0.0;
  if (e < nE) {
    const int ia = I[2 * (size_t)e], ib = I[2 * (size_t)e + 1];
    const double* p1 = V + 3 * (size_t)ia;
    const double* p2 = V + 3 * (size_t)ib;
    const double v[3] = {rest[3 * (size_t)e], rest[3 * (size_t)e + 1], rest[3 * (size_t)e + 2]};
    if (kind == MO_CERES_ROT_EDGE) {
      double r[6], Jd[3][3], Jr[3][3];
      edge_rot_block(p1, p2, R + 3 * (size_t)ia, R + 3 * (size_t)ib, v, lambda, r, Jd, Jr);
#pragma unroll
      for (int m = 0; m < 6; ++m) my += 0.5 * r[m] * r[m];
      if (ejac) {
#pragma unroll
        for (int m = 0; m < 3; ++m) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { ejac[18 * (size_t)e + 3 * m + c] = Jd[m][c]; ejac[18 * (size_t)e + 9 + 3 * m + c] = Jr[m][c]; }
        }
      }
      if (g) {
        double* gR = g + 3 * (size_t)nV;
        double* dR = diag + 3 * (size_t)nV;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double gp = r[0] * Jd[0][c] + r[1] * Jd[1][c] + r[2] * Jd[2][c];
          const double gr = r[0] * Jr[0][c] + r[1] * Jr[1][c] + r[2] * Jr[2][c];
          const double dp = Jd[0][c] * Jd[0][c] + Jd[1][c] * Jd[1][c] + Jd[2][c] * Jd[2][c];
          const double dr = Jr[0][c] * Jr[0][c] + Jr[1][c] * Jr[1][c] + Jr[2][c] * Jr[2][c];
          atomicAdd(g + 3 * (size_t)ia + c, gp); atomicAdd(g + 3 * (size_t)ib + c, -gp);
          atomicAdd(gR + 3 * (size_t)ia + c, gr + r[3 + c]); atomicAdd(gR + 3 * (size_t)ib + c, -r[3 + c]);
          atomicAdd(diag + 3 * (size_t)ia + c, dp); atomicAdd(diag + 3 * (size_t)ib + c, dp);
          atomicAdd(dR + 3 * (size_t)ia + c, dr + 1.0); atomicAdd(dR + 3 * (size_t)ib + c, 1.0);
        }
      }
    } else {
      const double lam = eff_lambda(kind, lambda, v);
      double r[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { r[k] = dmul(dsub(dsub(p1[k], p2[k]), v[k]), lam); my += 0.5 * r[k] * r[k]; }
      if (ejac) ejac[e] = lam;
      if (g) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          atomicAdd(g + 3 * (size_t)ia + c, r[c] * lam); atomicAdd(g + 3 * (size_t)ib + c, -r[c] * lam);
          atomicAdd(diag + 3 * (size_t)ia + c, lam * lam); atomicAdd(diag + 3 * (size_t)ib + c, lam * lam);
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = my;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kBlock / 32; ++k) t += s_part[k];
    atomicAdd(cost, t);
  }
}

// out = (distance blocks + damping) applied to v; the edge blocks are added by k_lm_hv_edges
__global__ void k_lm_hv_vertex(const double* __restrict__ gd, const double* __restrict__ damp, const double* __restrict__ v,
                               const int nV, const int n, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double o = damp ? damp[j] * v[j] : 0.0;
  if (j < 3 * nV) {
    const int i = j / 3;
    const double dot = gd[3 * (size_t)i] * v[3 * (size_t)i] + gd[3 * (size_t)i + 1] * v[3 * (size_t)i + 1] +
                       gd[3 * (size_t)i + 2] * v[3 * (size_t)i + 2];
    o += gd[j] * dot;
  }
  out[j] = o;
}

__global__ void __launch_bounds__(kBlock) k_lm_hv_edges(const int kind, const double* __restrict__ ejac,
                                                        const int* __restrict__ I, const int nE, const int nV,
                                                        const double* __restrict__ v, double* __restrict__ out) {
  const int e = blockIdx.x * kBlock + threadIdx.x;
  if (e >= nE) return;
  const int ia = I[2 * (size_t)e], ib = I[2 * (size_t)e + 1];
  if (kind == MO_CERES_ROT_EDGE) {
    const double* J = ejac + 18 * (size_t)e;
    const double* vr = v + 3 * (size_t)nV;
    double* outr = out + 3 * (size_t)nV;
    double dv[3], ra[3], y[3], y2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      dv[c] = v[3 * (size_t)ia + c] - v[3 * (size_t)ib + c];
      ra[c] = vr[3 * (size_t)ia + c];
      y2[c] = ra[c] - vr[3 * (size_t)ib + c];
    }
#pragma unroll
    for (int m = 0; m < 3; ++m)
      y[m] = J[3 * m] * dv[0] + J[3 * m + 1] * dv[1] + J[3 * m + 2] * dv[2] + J[9 + 3 * m] * ra[0] + J[9 + 3 * m + 1] * ra[1] +
             J[9 + 3 * m + 2] * ra[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double tp = J[c] * y[0] + J[3 + c] * y[1] + J[6 + c] * y[2];
      const double tr = J[9 + c] * y[0] + J[12 + c] * y[1] + J[15 + c] * y[2];
      atomicAdd(out + 3 * (size_t)ia + c, tp); atomicAdd(out + 3 * (size_t)ib + c, -tp);
      atomicAdd(outr + 3 * (size_t)ia + c, tr + y2[c]); atomicAdd(outr + 3 * (size_t)ib + c, -y2[c]);
    }
  } else {
    const double l2 = ejac[e] * ejac[e];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double t = l2 * (v[3 * (size_t)ia + c] - v[3 * (size_t)ib + c]);
      atomicAdd(out + 3 * (size_t)ia + c, t); atomicAdd(out + 3 * (size_t)ib + c, -t);
    }
  }
}

// sum_j a[j]*b[j] (and optionally max |a[j]|) accumulated into out[0] (out[1])
__global__ void __launch_bounds__(256) k_lm_dot(const double* __restrict__ a, const double* __restrict__ b, const int n,
                                                double* __restrict__ out, double* __restrict__ out_max) {
  __shared__ double s_part[8], s_max[8];
  double my = 0.0, mx = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    my += a[j] * b[j];
    mx = fmax(mx, fabs(a[j]));
  }
  for (int o = 16; o > 0; o >>= 1) { my += __shfl_xor_sync(0xffffffffu, my, o); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = my; s_max[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0, m = 0.0;
    for (int k = 0; k < 8; ++k) { t += s_part[k]; m = fmax(m, s_max[k]); }
    atomicAdd(out, t);
    if (out_max) atomicMax(reinterpret_cast<unsigned long long*>(out_max), (unsigned long long)__double_as_longlong(m));   // m >= 0
  }
}

// Jacobi scaling (once) and LM damping: scale = 1/(1+sqrt(diag)); damp = clamp(diag*scale^2, 1e-6, 1e32)/radius/scale^2
__global__ void k_lm_scale(const double* __restrict__ diag, const int n, double* __restrict__ scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) scale[j] = 1.0 / (1.0 + sqrt(diag[j]));
}
__global__ void k_lm_damp(const double* __restrict__ diag, const double* __restrict__ scale, const int n, const double radius,
                          double* __restrict__ damp, double* __restrict__ minv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double s2 = scale[j] * scale[j];
  const double ds = fmin(fmax(diag[j] * s2, 1e-6), 1e32);
  const double dd = ds / radius / s2;
  damp[j] = dd;
  minv[j] = 1.0 / (diag[j] + dd);
}

// CG start: delta = 0, r = -g, z = M^-1 r, p = z, sc[0] = r.z, rr[0] = r.r
__global__ void __launch_bounds__(256) k_cg_init(const double* __restrict__ g, const double* __restrict__ minv, const int n,
                                                 double* __restrict__ delta, double* __restrict__ r, double* __restrict__ p,
                                                 double* __restrict__ rz, double* __restrict__ rr) {
  __shared__ double s_a[8], s_b[8];
  double a = 0.0, b = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double rj = -g[j], zj = minv[j] * rj;
    delta[j] = 0.0; r[j] = rj; p[j] = zj;
    a += rj * zj; b += rj * rj;
  }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < 8; ++k) { ta += s_a[k]; tb += s_b[k]; }
    atomicAdd(rz, ta); atomicAdd(rr, tb);
  }
}
// delta += alpha p; r -= alpha Ap; accumulates r.z and r.r of the new residual into slot k+1
__global__ void __launch_bounds__(256) k_cg_update(const double* __restrict__ p, const double* __restrict__ Ap,
                                                   const double* __restrict__ minv, const int n, const double* __restrict__ rz,
                                                   const double* __restrict__ pAp, double* __restrict__ delta,
                                                   double* __restrict__ r, double* __restrict__ rz_next, double* __restrict__ rr_next) {
  __shared__ double s_a[8], s_b[8];
  const double den = *pAp;
  const double alpha = den > 0.0 ? *rz / den : 0.0;
  double a = 0.0, b = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    delta[j] += alpha * p[j];
    const double rj = r[j] - alpha * Ap[j];
    r[j] = rj;
    a += rj * rj * minv[j]; b += rj * rj;
  }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < 8; ++k) { ta += s_a[k]; tb += s_b[k]; }
    atomicAdd(rz_next, ta); atomicAdd(rr_next, tb);
  }
}
// p = M^-1 r + beta p
__global__ void k_cg_dir(const double* __restrict__ r, const double* __restrict__ minv, const int n, const double* __restrict__ rz,
                         const double* __restrict__ rz_next, double* __restrict__ p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double beta = *rz > 0.0 ? *rz_next / *rz : 0.0;
  p[j] = minv[j] * r[j] + beta * p[j];
}
__global__ void k_lm_axpy(const double* __restrict__ x, const double* __restrict__ d, const int n, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = x[j] + d[j];
}

}  // namespace

int ceres_edges(int kind, const double* d_V, const double* d_R, int nV, const int* d_I, const double* d_rest, int nE,
                double lambda, double* d_res, double* d_jac, cudaStream_t s) {
  if (nE == 0) return MO_OK;
  k_ceres_edges<<<div_up(nE, kBlock), kBlock, 0, s>>>(kind, d_V, d_R, nV, d_I, d_rest, nE, lambda, d_res, d_jac, nullptr,
                                                      nullptr, nullptr);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

int ceres_problem(const Template* TD, int kind, const double* d_V, const double* d_R, int nV, const int* d_I,
                  const double* d_rest, int nE, double lambda, double* d_cost2, double* d_gV, double* d_gR, cudaStream_t s) {
  if (d_cost2) MO_CUDA(cudaMemsetAsync(d_cost2, 0, 2 * sizeof(double), s));
  if (d_gR) MO_CUDA(cudaMemsetAsync(d_gR, 0, sizeof(double) * 3 * (size_t)nV, s));
  if (d_gV && !TD) MO_CUDA(cudaMemsetAsync(d_gV, 0, sizeof(double) * 3 * (size_t)nV, s));
  if (TD && nV > 0) {
    k_ceres_distance<<<div_up(nV, kBlock), kBlock, 0, s>>>(TD->d_grid64, TD->N, d_V, nV, d_cost2, d_gV);
    MO_LAUNCH_CHECK();
  }
  if (nE > 0) {
    k_ceres_edges<<<div_up(nE, kBlock), kBlock, 0, s>>>(kind, d_V, d_R, nV, d_I, d_rest, nE, lambda, nullptr, nullptr,
                                                        d_cost2 ? d_cost2 + 1 : nullptr, d_gV, d_gR);
    MO_LAUNCH_CHECK();
  }
  return MO_OK;
}


// ---------------------------------------------------------------------------------------------
// ceres::Solve restated for these problems: trust-region Levenberg-Marquardt with Ceres' default
// options (trust_region_minimizer.cc / levenberg_marquardt_strategy.cc; un-vendored, recalled):
// initial radius 1e4, max 1e16, min 1e-32, min/max LM diagonal 1e-6 / 1e32, Jacobi scaling taken once
// at the start, step accepted when the cost decrease exceeds 1e-3 of the model's, radius /=
// max(1/3, 1 - (2 rho - 1)^3) after an accepted step and /= 2, 4, 8 ... after rejected ones,
// function / gradient / parameter tolerances 1e-6 / 1e-10 / 1e-8.
// ---------------------------------------------------------------------------------------------
int ceres_solve(const Template* TD, int kind, double* d_V, double* d_R, int nV, const int* d_I, const double* d_rest, int nE,
                double lambda, int max_iters, int max_cg, double cg_tol, int verbose, double* h_summary, cudaStream_t s) {
  const bool rot = kind == MO_CERES_ROT_EDGE;
  const int n = (rot ? 6 : 3) * nV;
  if (h_summary) for (int i = 0; i < 10; ++i) h_summary[i] = 0.0;
  if (n == 0) return MO_OK;
  if (max_cg <= 0) max_cg = 4000;
  if (cg_tol <= 0.0) cg_tol = 1e-10;
  const size_t ejac_n = rot ? 18 * (size_t)nE : (size_t)nE;
  const int nsc = 16 + 3 * (max_cg + 2);
  const size_t total = 12 * (size_t)n + 3 * (size_t)nV + ejac_n + nsc;
  ScratchBuf<double> scratch;   // released on every exit path
  MO_CUDA(scratch.alloc(total, s));
  double* buf = scratch.p;
  double* x = buf; double* xn = x + n; double* g = xn + n; double* diag = g + n; double* scale = diag + n;
  double* damp = scale + n; double* minv = damp + n; double* delta = minv + n; double* r = delta + n;
  double* p = r + n; double* Ap = p + n; double* q = Ap + n; double* gd = q + n; double* ejac = gd + 3 * (size_t)nV;
  double* sc = ejac + ejac_n;
  double* rz = sc + 16; double* rr = rz + (max_cg + 2); double* pAp = rr + (max_cg + 2);
  MO_CUDA(cudaMemcpyAsync(x, d_V, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s));
  if (rot) MO_CUDA(cudaMemcpyAsync(x + 3 * (size_t)nV, d_R, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s));
  const double* grid = TD ? TD->d_grid64 : nullptr;
  const int N = TD ? TD->N : 0;
  const int gb = div_up(n, 256), rb = std::min(gb, 592);

  auto evaluate = [&](double* at, double* cost2, bool lin) -> int {
    MO_CUDA(cudaMemsetAsync(cost2, 0, 2 * sizeof(double), s));
    if (lin) {
      if (rot) {
        MO_CUDA(cudaMemsetAsync(g + 3 * (size_t)nV, 0, sizeof(double) * 3 * (size_t)nV, s));
        MO_CUDA(cudaMemsetAsync(diag + 3 * (size_t)nV, 0, sizeof(double) * 3 * (size_t)nV, s));
      }
      MO_CUDA(cudaMemsetAsync(sc + 4, 0, 2 * sizeof(double), s));
    }
    k_lm_lin_distance<<<div_up(nV, kBlock), kBlock, 0, s>>>(grid, N, at, nV, lin ? gd : nullptr, cost2, lin ? g : nullptr,
                                                            lin ? diag : nullptr);
    MO_LAUNCH_CHECK();
    if (nE > 0) {
      k_lm_lin_edges<<<div_up(nE, kBlock), kBlock, 0, s>>>(kind, at, nV, d_I, d_rest, nE, lambda, lin ? ejac : nullptr,
                                                            cost2 + 1, lin ? g : nullptr, lin ? diag : nullptr);
      MO_LAUNCH_CHECK();
    }
    if (lin) {
      k_lm_dot<<<rb, 256, 0, s>>>(g, g, n, sc + 5, sc + 4);   // sc[4] = max |g|
      MO_LAUNCH_CHECK();
    }
    return MO_OK;
  };
  auto apply_h = [&](const double* dmp, const double* v, double* out) -> int {
    k_lm_hv_vertex<<<gb, 256, 0, s>>>(gd, dmp, v, nV, n, out);
    MO_LAUNCH_CHECK();
    if (nE > 0) {
      k_lm_hv_edges<<<div_up(nE, kBlock), kBlock, 0, s>>>(kind, ejac, d_I, nE, nV, v, out);
      MO_LAUNCH_CHECK();
    }
    return MO_OK;
  };
  double h[16];
  auto fetch = [&](const double* src, int cnt) -> int {
    MO_CUDA(cudaMemcpyAsync(h, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost, s));
    MO_CUDA(cudaStreamSynchronize(s));
    return MO_OK;
  };
#define MO_TRY(expr) do { int rc__ = (expr); if (rc__ != MO_OK) return rc__; } while (0)

  MO_TRY(evaluate(x, sc, true));
  k_lm_scale<<<gb, 256, 0, s>>>(diag, n, scale);
  MO_LAUNCH_CHECK();
  MO_TRY(fetch(sc, 5));
  double cost = h[0] + h[1], gmax = h[4];
  const double initial_cost = cost;
  double radius = 1e4, decrease = 2.0;
  const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8;
  int iter = 0, accepted = 0, invalid = 0, cg_total = 0, term = 3;
  if (verbose) printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius  cg_iter\n%4d  %.6e    0.00e+00    %.2e   0.00e+00   0.00e+00  %.2e        0\n", 0, cost, gmax, radius);
  if (gmax <= gtol) term = 1;
  while (term == 3 && iter < max_iters) {
    ++iter;
    k_lm_damp<<<gb, 256, 0, s>>>(diag, scale, n, radius, damp, minv);
    MO_LAUNCH_CHECK();
    // ---- PCG on (J^T J + D^T D) delta = -g ---------------------------------------------------------------
    MO_CUDA(cudaMemsetAsync(rz, 0, sizeof(double) * 3 * (max_cg + 2), s));
    k_cg_init<<<rb, 256, 0, s>>>(g, minv, n, delta, r, p, rz, rr);
    MO_LAUNCH_CHECK();
    int k = 0;
    double rr0 = -1.0;
    while (k < max_cg) {
      const int stop = std::min(max_cg, k + kCgCheck);
      for (; k < stop; ++k) {
        MO_TRY(apply_h(damp, p, Ap));
        k_lm_dot<<<rb, 256, 0, s>>>(p, Ap, n, pAp + k, nullptr);
        MO_LAUNCH_CHECK();
        k_cg_update<<<rb, 256, 0, s>>>(p, Ap, minv, n, rz + k, pAp + k, delta, r, rz + k + 1, rr + k + 1);
        MO_LAUNCH_CHECK();
        k_cg_dir<<<gb, 256, 0, s>>>(r, minv, n, rz + k, rz + k + 1, p);
        MO_LAUNCH_CHECK();
      }
      if (rr0 < 0.0) { MO_TRY(fetch(rr, 1)); rr0 = h[0]; }
      MO_TRY(fetch(rr + k, 1));
      if (!(h[0] > cg_tol * cg_tol * rr0)) break;
    }
    cg_total += k;
    // ---- model decrease, step norm -------------------------------------------------------------------------
    MO_TRY(apply_h(nullptr, delta, q));
    MO_CUDA(cudaMemsetAsync(sc + 6, 0, 4 * sizeof(double), s));
    k_lm_dot<<<rb, 256, 0, s>>>(g, delta, n, sc + 6, nullptr);
    k_lm_dot<<<rb, 256, 0, s>>>(delta, q, n, sc + 7, nullptr);
    k_lm_dot<<<rb, 256, 0, s>>>(delta, delta, n, sc + 8, nullptr);
    k_lm_dot<<<rb, 256, 0, s>>>(x, x, n, sc + 9, nullptr);
    MO_LAUNCH_CHECK();
    MO_TRY(fetch(sc + 6, 4));
    const double model_change = -(h[0] + 0.5 * h[1]);
    const double step_norm = std::sqrt(h[2]), x_norm = std::sqrt(h[3]);
    if (!(model_change > 0.0)) {   // invalid step
      if (++invalid >= 5) { term = 4; break; }
      radius /= decrease; decrease *= 2.0;
      if (verbose) printf("%4d  %.6e   invalid step (model change %.2e)            %.2e  %7d\n", iter, cost, model_change, radius, k);
      continue;
    }
    invalid = 0;
    if (step_norm <= ptol * (x_norm + ptol)) { term = 2; break; }
    k_lm_axpy<<<gb, 256, 0, s>>>(x, delta, n, xn);
    MO_LAUNCH_CHECK();
    MO_TRY(evaluate(xn, sc + 2, false));
    MO_TRY(fetch(sc + 2, 2));
    const double new_cost = h[0] + h[1];
    const double rho = (cost - new_cost) / model_change;
    // Ceres' order (TrustRegionMinimizer::Minimize): ParameterToleranceReached (above), then FunctionToleranceReached on
    // the CANDIDATE whether or not it would be accepted -- the solver returns without applying the step -- then
    // IsStepSuccessful
    if (std::fabs(cost - new_cost) <= ftol * cost) { term = 0; break; }
    if (rho > 1e-3) {
      std::swap(x, xn);
      const double t = 2.0 * rho - 1.0;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
      decrease = 2.0;
      const double change = cost - new_cost;
      MO_TRY(evaluate(x, sc, true));
      MO_TRY(fetch(sc, 5));
      cost = h[0] + h[1]; gmax = h[4];
      ++accepted;
      if (verbose) printf("%4d  %.6e   %9.2e    %.2e   %.2e  %9.2e  %.2e  %7d\n", iter, cost, change, gmax, step_norm, rho, radius, k);
      if (gmax <= gtol) { term = 1; break; }
    } else {
      radius /= decrease; decrease *= 2.0;
      if (verbose) printf("%4d  %.6e   %9.2e    %.2e   %.2e  %9.2e  %.2e  %7d\n", iter, cost, 0.0, gmax, step_norm, rho, radius, k);
      if (radius < 1e-32) { term = 5; break; }
    }
  }
  MO_CUDA(cudaMemcpyAsync(d_V, x, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s));
  if (rot) MO_CUDA(cudaMemcpyAsync(d_R, x + 3 * (size_t)nV, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s));
  MO_TRY(fetch(sc, 2));
  if (h_summary) {
    h_summary[0] = initial_cost; h_summary[1] = h[0] + h[1]; h_summary[2] = h[0]; h_summary[3] = h[1]; h_summary[4] = iter;
    h_summary[5] = accepted; h_summary[6] = cg_total; h_summary[7] = term; h_summary[8] = radius; h_summary[9] = gmax;
  }
#undef MO_TRY
  return MO_OK;
}

}  // namespace mo
