// Internal declarations shared by the translation units of libmeshode_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/meshode_b200.h"

namespace mo {

constexpr int kStatSlots = 32;   // the build kernel spreads its counters over this many slots of Template::d_stats

// What DeformParams holds in the reference (src/interface/deform_params.h:7-25), as device buffers.
struct Template {
  int device = 0;
  cudaStream_t free_stream = 0;   // the buffers return to the pool in this stream's order (set by mo_template_destroy*)
  int N = 0, nV = 0, nF = 0;
  int z0 = 0, z1 = 0;           // voxel slices this template's build computes (z-slab sharding)
  int tz_first = 0, tz_stride = 1;   // or: z-tile layers (4 slices each) tz_first, tz_first + tz_stride, ... (cyclic sharding)
  double* d_Vn = nullptr;       // [nV,3] normalised FP64 target vertices (Mesh::V_ after Normalize)
  int* d_F = nullptr;           // [nF,3]
  double* d_grid64 = nullptr;   // [N^3] z,y,x  (UniformGrid::voxel_distance_)
  float* d_grid32 = nullptr;    // [N^3] (float) of the above
  int* d_nearest = nullptr;     // [N^3] nearest triangle (igl's I)
  double* d_xf = nullptr;       // [4] scale, trans.x, trans.y, trans.z  (params.scale / params.trans)
  unsigned long long* d_stats = nullptr;   // [kStatSlots][8]: fp32 tests, fp64 tests, cull tests, error bits (slot 0 only), disc pre-tests; counters summed over the slots
  // one edge set per template (params.edge_offset / edge_lambda)
  int kind = MO_EDGES_NONE;
  int eV = 0, eF = 0, eE = 0, nEdges = 0;
  int2* d_ev = nullptr;         // [nEdges] (v0, v1)
  float* d_rest = nullptr;      // [nEdges,3]
  float* d_lambda = nullptr;    // [nEdges] (CAD only)
  int* d_csr_start = nullptr;   // [eV+1]
  int* d_csr_key = nullptr;     // [2*nEdges] 2*edge + side, ascending per vertex
  float4* d_inc = nullptr;      // [2*nEdges] per-incidence records in CSR order: rest vector, other endpoint | side << 31 (built on first backward)
  float* d_inc_lambda = nullptr;   // [2*nEdges] lambda per incidence (CAD only)
  float* d_v0 = nullptr;        // [eV,3] vertices at store time (rest = V0[v1] - V0[v0])
  float* d_cells = nullptr;     // [N^3][8] the eight corner values of every cell, one 32-byte record (deform engine; built on first use)
  unsigned* d_ell = nullptr;    // [ceil(ell_D/2)][eV] packed other endpoints per incident edge, built on first deform
  int ell_D = 0;
  unsigned* d_nbr = nullptr;    // [nbr_W][eV] distinct neighbours with multiplicities (fast deform loop), built on first use
  int nbr_W = 0;
};

// Stream-ordered allocation from the device's default memory pool (release threshold raised so
// that per-pair templates are recycled without driver calls).  dev_free returns the block to the
// pool in the order of the legacy default stream.
cudaError_t dev_alloc_bytes(void** p, size_t bytes, cudaStream_t s);
template <class T> inline cudaError_t dev_alloc(T** p, size_t count, cudaStream_t s) {
  return dev_alloc_bytes(reinterpret_cast<void**>(p), count * sizeof(T), s);
}
void dev_free(void* p, cudaStream_t s = 0);

// Stream-ordered scratch that is returned to the pool on every exit path of the function that owns it.
template <class T>
struct ScratchBuf {
  T* p = nullptr;
  cudaStream_t s = 0;
  ScratchBuf() = default;
  ScratchBuf(const ScratchBuf&) = delete;
  ScratchBuf& operator=(const ScratchBuf&) = delete;
  ~ScratchBuf() { if (p) cudaFreeAsync(p, s); }
  cudaError_t alloc(size_t count, cudaStream_t stream) {
    s = stream;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream);
  }
};

void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define MO_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return ::mo::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
// every kernel launch of this library is followed by MO_LAUNCH_CHECK, which also counts it
// (mo_launch_count: the "gpu_launches" figure of bench.py)
extern std::atomic<unsigned long long> g_launches;
extern std::atomic<int> g_build_stats;   // mo_build_stats_enable: distance-field builds count their tests
#define MO_LAUNCH_CHECK()                                             \
  do {                                                                \
    ::mo::g_launches.fetch_add(1, std::memory_order_relaxed);         \
    MO_CUDA(cudaGetLastError());                                      \
  } while (0)
#define MO_REQUIRE(cond, msg)                                                \
  do {                                                                       \
    if (!(cond)) { ::mo::set_error(std::string("bad argument: ") + msg); return MO_ERR_BAD_ARG; } \
  } while (0)

// sdf_build.cu
int build_field_from_f32(Template& T, const float* d_V, cudaStream_t s);
int build_field_from_normalized(Template& T, cudaStream_t s);
// sampler.cu
int launch_distance_f32(const Template& T, const float* d_V, int n, float* d_out, float* d_grad, int mode, cudaStream_t s);
int launch_distance_f64(const Template& T, const double* d_P, int n, double* d_val, double* d_grad, cudaStream_t s);
// edges.cu
void free_edges(Template& T, cudaStream_t s = 0);
int edges_store(Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                cudaStream_t s);
int edges_forward(const Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                  float* d_out, cudaStream_t s);
int edges_backward(const Template& T, const float* d_V, int nV, float* d_grad, cudaStream_t s);
int edges_backward_atomic(const Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                          int nE, float* d_grad, cudaStream_t s);
int loss_fused(const Template& TD, const Template* TE, const float* d_V, int nV, float w_edge, float mask_thr,
               double* d_loss, float* d_grad, cudaStream_t s);
int adam_loop_coop(const Template& TD, const Template* TE, float* d_V, int nV, float w_edge, float mask_thr,
                   const float2* d_sched, int iters, float w1, float b2, float w2, float eps, float* d_scratch,
                   cudaStream_t s);

// ceres_path.cu
int ceres_edges(int kind, const double* d_V, const double* d_R, int nV, const int* d_I, const double* d_rest, int nE,
                double lambda, double* d_res, double* d_jac, cudaStream_t s);
int ceres_problem(const Template* TD, int kind, const double* d_V, const double* d_R, int nV, const int* d_I,
                  const double* d_rest, int nE, double lambda, double* d_cost2, double* d_gV, double* d_gR, cudaStream_t s);
int ceres_solve(const Template* TD, int kind, double* d_V, double* d_R, int nV, const int* d_I, const double* d_rest, int nE,
                double lambda, int max_iters, int max_cg, double cg_tol, int verbose, double* h_summary, cudaStream_t s);

// deform.cu
int deform_batch_adam(Template* const* TD, Template* const* TE, float* const* h_V, int B, int iters, double lr,
                      double beta1, double beta2, double eps, int flags, cudaStream_t s);
int deform_adam_large(Template& TD, Template& TE, float* d_V, int nV, float w_edge, float mask_thr, int iters, double lr,
                      double beta1, double beta2, double eps, cudaStream_t s);

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// order-preserving float <-> uint maps for atomicMin/atomicMax
__device__ __forceinline__ unsigned f2o(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(unsigned o) {
  unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

// IEEE round-to-nearest FP64 arithmetic that nvcc will not contract into FMAs, so the
// FP64 parts match the host oracle (g++ -ffp-contract=off) bit for bit.
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

}  // namespace mo
