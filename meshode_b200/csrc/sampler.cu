// Trilinear distance lookup and its gradient: UniformGrid::DistanceFloat<float> /
// DistanceFloat<Jet<float,3>> (reference src/lib/uniformgrid.cc:85-150) behind
// DistanceFieldLoss_forward/backward (src/interface/distance_layer.cc:8-81), and
// UniformGrid::distance<double> / distance<Jet<double,3>> (uniformgrid.cc:18-83) behind
// DistanceLoss (src/lib/distanceloss.h:6-25).
//
// The arithmetic is the reference's, operation for operation, on a dual number with the
// operator definitions of ceres::Jet, using contraction-free IEEE operations, so value
// and partials agree bit for bit with the CPU evaluation.  One thread per vertex;
// vertices and gradients move through shared memory so that global traffic is 128-bit
// and coalesced; the eight corner fetches are read-only (L2-resident grid) gathers.
#include "common.cuh"
#include "sampler.cuh"

namespace mo {
namespace {

constexpr int kBlock = 256;

// mode bit 0: write value^2 (forward), bit 1: write 0.5*d(value^2) (backward)
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_distance_f32(const float* __restrict__ grid, const int n,
                                                         const float* __restrict__ V, const int nV,
                                                         float* __restrict__ out, float* __restrict__ grad,
                                                         const int vec_ok) {
  __shared__ __align__(16) float s_v[kBlock * 3];
  const int tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * kBlock;           // first vertex of this CTA
  const int cnt = (int)min((size_t)kBlock, (size_t)nV - base);
  const float* src = V + base * 3;
  // 768 floats per CTA start at a multiple of 3072 B: 16-byte aligned, so full CTAs move as float4
  if (cnt == kBlock && (vec_ok & 1)) {
    if (tid < kBlock * 3 / 4) reinterpret_cast<float4*>(s_v)[tid] = __ldg(reinterpret_cast<const float4*>(src) + tid);
  } else {
    for (int i = tid; i < cnt * 3; i += kBlock) s_v[i] = __ldg(src + i);
  }
  __syncthreads();
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  if (tid < cnt) {
    const float x = s_v[tid * 3], y = s_v[tid * 3 + 1], z = s_v[tid * 3 + 2];
    if (MODE & 2) {
      typedef Jet3<float> J;
      J vd = sample<J, float>(grid, n, J(x, 1.f, 0.f, 0.f), J(y, 0.f, 1.f, 0.f), J(z, 0.f, 0.f, 1.f));
      vd = vd * vd;                                          // distance_layer.cc:73
      if (MODE & 1) out[base + tid] = vd.a;
      g0 = (float)((double)vd.v0 * 0.5); g1 = (float)((double)vd.v1 * 0.5); g2 = (float)((double)vd.v2 * 0.5);   // :75-77
    } else {
      typedef Num<float> F;
      const F d = sample<F, float>(grid, n, F(x), F(y), F(z));
      out[base + tid] = __fmul_rn(d.a, d.a);                 // distance_layer.cc:29-33
    }
  }
  if (MODE & 2) {
    __syncthreads();
    if (tid < cnt) { s_v[tid * 3] = g0; s_v[tid * 3 + 1] = g1; s_v[tid * 3 + 2] = g2; }
    __syncthreads();
    float* dst = grad + base * 3;
    if (cnt == kBlock && (vec_ok & 2)) {
      if (tid < kBlock * 3 / 4) reinterpret_cast<float4*>(dst)[tid] = reinterpret_cast<const float4*>(s_v)[tid];
    } else {
      for (int i = tid; i < cnt * 3; i += kBlock) dst[i] = s_v[i];
    }
  }
}

__global__ void __launch_bounds__(kBlock) k_distance_f64(const double* __restrict__ grid, const int n,
                                                         const double* __restrict__ P, const int nP,
                                                         double* __restrict__ val, double* __restrict__ grad) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= nP) return;
  const double x = P[3 * (size_t)i], y = P[3 * (size_t)i + 1], z = P[3 * (size_t)i + 2];
  if (grad) {
    typedef Jet3<double> J;
    const J r = sample<J, double>(grid, n, J(x, 1.0, 0.0, 0.0), J(y, 0.0, 1.0, 0.0), J(z, 0.0, 0.0, 1.0));
    val[i] = r.a;
    grad[3 * (size_t)i] = r.v0; grad[3 * (size_t)i + 1] = r.v1; grad[3 * (size_t)i + 2] = r.v2;
  } else {
    typedef Num<double> F;
    val[i] = sample<F, double>(grid, n, F(x), F(y), F(z)).a;
  }
}

}  // namespace

int launch_distance_f32(const Template& T, const float* d_V, int n, float* d_out, float* d_grad, int mode, cudaStream_t s) {
  if (n == 0) return MO_OK;
  const int blocks = div_up(n, kBlock);
  const int vec_ok = (((uintptr_t)d_V & 15) == 0 ? 1 : 0) | (((uintptr_t)d_grad & 15) == 0 ? 2 : 0);
  if (mode == 1) k_distance_f32<1><<<blocks, kBlock, 0, s>>>(T.d_grid32, T.N, d_V, n, d_out, d_grad, vec_ok);
  else if (mode == 2) k_distance_f32<2><<<blocks, kBlock, 0, s>>>(T.d_grid32, T.N, d_V, n, d_out, d_grad, vec_ok);
  else k_distance_f32<3><<<blocks, kBlock, 0, s>>>(T.d_grid32, T.N, d_V, n, d_out, d_grad, vec_ok);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

int launch_distance_f64(const Template& T, const double* d_P, int n, double* d_val, double* d_grad, cudaStream_t s) {
  if (n == 0) return MO_OK;
  k_distance_f64<<<div_up(n, kBlock), kBlock, 0, s>>>(T.d_grid64, T.N, d_P, n, d_val, d_grad);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

}  // namespace mo
