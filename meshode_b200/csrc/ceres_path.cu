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

}  // namespace mo
