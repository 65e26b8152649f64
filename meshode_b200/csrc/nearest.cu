// Exact nearest-vertex search: the cKDTree(src).query(tar, k=1) of the reference's ReverseLossLayer
// (src/python/layers/reverse_loss_layer.py:15-19), on the GPU.
//
// One thread per query point, candidate points streamed through shared memory in tiles (broadcast
// reads), squared distances in FP64 (float32 coordinates are exact in FP64, as in cKDTree, which
// works on doubles).  Lowest index wins exact ties.
#include "common.cuh"

namespace mo {
namespace {

constexpr int kBlock = 128;
constexpr int kTile = 1024;

__global__ void __launch_bounds__(kBlock) k_nearest_vertex(const float* __restrict__ Q, const int nQ,
                                                           const float* __restrict__ P, const int nP,
                                                           int* __restrict__ idx, double* __restrict__ dist2) {
  __shared__ double s_p[kTile * 3];
  const int q = blockIdx.x * kBlock + threadIdx.x;
  const bool live = q < nQ;
  const int qq = live ? q : nQ - 1;
  const double x = (double)Q[3 * (size_t)qq], y = (double)Q[3 * (size_t)qq + 1], z = (double)Q[3 * (size_t)qq + 2];
  double best = 1.0 / 0.0;
  int bi = -1;
  for (int base = 0; base < nP; base += kTile) {
    const int cnt = min(kTile, nP - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 3; i += kBlock) s_p[i] = (double)__ldg(P + 3 * (size_t)base + i);
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      const double dx = dsub(s_p[3 * j], x), dy = dsub(s_p[3 * j + 1], y), dz = dsub(s_p[3 * j + 2], z);
      const double d = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
      if (d < best) { best = d; bi = base + j; }   // strict: the lowest index wins ties
    }
  }
  if (live) {
    idx[q] = bi;
    if (dist2) dist2[q] = best;
  }
}

}  // namespace

int nearest_vertex(const float* d_Q, int nQ, const float* d_P, int nP, int* d_idx, double* d_dist2, cudaStream_t s) {
  if (nQ == 0) return MO_OK;
  k_nearest_vertex<<<div_up(nQ, kBlock), kBlock, 0, s>>>(d_Q, nQ, d_P, nP, d_idx, d_dist2);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

}  // namespace mo

extern "C" int mo_nearest_vertex(const float* d_Q, int nQ, const float* d_P, int nP, int* d_idx, double* d_dist2,
                                 mo_stream_t stream) {
  MO_REQUIRE(nQ >= 0 && nP > 0, "nearest vertex: empty point set");
  MO_REQUIRE(nQ == 0 || (d_Q && d_P && d_idx), "null pointer");
  return mo::nearest_vertex(d_Q, nQ, d_P, nP, d_idx, d_dist2, (cudaStream_t)stream);
}
