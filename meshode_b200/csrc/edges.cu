// Edge-rigidity losses: Store{Rigidity,Graph,Cad}Information and
// {Rigid,Graph,Cad}EdgeLoss_forward/backward (reference src/interface/rigid_layer.cc,
// graph_layer.cc, cad_layer.cc), plus the fused per-iteration loss of the Python layers
// (src/python/layers/{rigid,graph,graph2}_loss_layer.py).
//
// forward  : one thread per edge, indices as passed by the caller, vertex gathers from L2.
// backward : one thread per vertex walking a CSR (vertex -> incident (edge, side) keys in
//            ascending edge order) built at store time, so each vertex accumulates its
//            contributions in exactly the order of the reference's serial scatter loop:
//            atomic-free and bit-identical.  An edge-parallel variant with warp-aggregated
//            red.global.add.f32 is provided for callers whose connectivity changes.
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include "common.cuh"
#include "sampler.cuh"

namespace cg = cooperative_groups;

namespace mo {
namespace {

constexpr int kBlock = 256;

// endpoints of edge e in the reference's enumeration order
__device__ __forceinline__ int2 edge_endpoints(int kind, int e, const int* __restrict__ F, const int* __restrict__ E, int nE) {
  if (kind == MO_EDGES_GRAPH || (kind == MO_EDGES_CAD && e < nE)) return make_int2(__ldg(E + 2 * (size_t)e), __ldg(E + 2 * (size_t)e + 1));
  const int o = kind == MO_EDGES_CAD ? e - nE : e;
  const int f = o / 3, j = o - 3 * f;
  return make_int2(__ldg(F + 3 * (size_t)f + j), __ldg(F + 3 * (size_t)f + (j + 1) % 3));   // rigid_layer.cc:34-35
}

__global__ void k_edges_store(int kind, const float* __restrict__ V, int nV, const int* __restrict__ F,
                              const int* __restrict__ E, int nE, int nEdges, int2* __restrict__ ev,
                              float* __restrict__ rest, float* __restrict__ lambda, int* __restrict__ deg,
                              unsigned long long* __restrict__ stats) {
  const int e = blockIdx.x * kBlock + threadIdx.x;
  if (e >= nEdges) return;
  int2 v = edge_endpoints(kind, e, F, E, nE);
  if ((unsigned)v.x >= (unsigned)nV || (unsigned)v.y >= (unsigned)nV) {
    atomicOr(&stats[3], 4ull);
    ev[e] = make_int2(-1, -1);
    rest[3 * (size_t)e] = rest[3 * (size_t)e + 1] = rest[3 * (size_t)e + 2] = 0.f;
    if (lambda) lambda[e] = 0.f;
    return;
  }
  ev[e] = v;
  float r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r[k] = fsub(__ldg(V + 3 * (size_t)v.y + k), __ldg(V + 3 * (size_t)v.x + k));   // rigid_layer.cc:39-43
    rest[3 * (size_t)e + k] = r[k];
  }
  if (lambda) {   // cad_layer.cc:48-49: 2e-2 / (norm + 1e-8), double arithmetic, stored as float
    const float norm = __fsqrt_rn(fadd(fadd(fmul(r[0], r[0]), fmul(r[1], r[1])), fmul(r[2], r[2])));
    lambda[e] = (float)__ddiv_rn(2e-2, dadd((double)norm, 1e-8));
  }
  atomicAdd(&deg[v.x], 1);
  atomicAdd(&deg[v.y], 1);
}

__global__ void k_scan1(const int* __restrict__ in, int* __restrict__ out, int n) {   // one CTA, exclusive
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = min(n, tid * per), e = min(n, b + per);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += in[i];
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = s_warp[tid];
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (tid >= o) w += v; }
    s_warp[tid] = w;
  }
  __syncthreads();
  int run = incl - sum + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
  for (int i = b; i < e; ++i) { out[i] = run; run += in[i]; }
  if (tid == 1023) out[n] = s_warp[31];
}

__global__ void k_csr_fill(const int2* __restrict__ ev, int nEdges, const int* __restrict__ start, int* __restrict__ fill,
                           int* __restrict__ keys) {
  const int e = blockIdx.x * kBlock + threadIdx.x;
  if (e >= nEdges) return;
  const int2 v = ev[e];
  if (v.x < 0) return;
  keys[start[v.x] + atomicAdd(&fill[v.x], 1)] = 2 * e;       // side 0: "l_v0 -= r"
  keys[start[v.y] + atomicAdd(&fill[v.y], 1)] = 2 * e + 1;   // side 1: "l_v1 += r"
}

__global__ void k_csr_sort(const int* __restrict__ start, int nV, int* __restrict__ keys) {
  const int v = blockIdx.x * kBlock + threadIdx.x;
  if (v >= nV) return;
  int* a = keys + start[v];
  const int n = start[v + 1] - start[v];
  if (n <= 32) {
    for (int i = 1; i < n; ++i) {
      const int x = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
      a[j + 1] = x;
    }
  } else {   // heap sort for hubs
    for (int st = n / 2 - 1; st >= 0; --st) {
      int root = st;
      for (;;) {
        int ch = 2 * root + 1;
        if (ch >= n) break;
        if (ch + 1 < n && a[ch] < a[ch + 1]) ++ch;
        if (a[root] >= a[ch]) break;
        const int t = a[root]; a[root] = a[ch]; a[ch] = t; root = ch;
      }
    }
    for (int end = n - 1; end > 0; --end) {
      const int t0 = a[0]; a[0] = a[end]; a[end] = t0;
      int root = 0;
      for (;;) {
        int ch = 2 * root + 1;
        if (ch >= end) break;
        if (ch + 1 < end && a[ch] < a[ch + 1]) ++ch;
        if (a[root] >= a[ch]) break;
        const int t = a[root]; a[root] = a[ch]; a[ch] = t; root = ch;
      }
    }
  }
}

// {Rigid,Graph,Cad}EdgeLoss_forward
__global__ void k_edges_forward(int kind, const float* __restrict__ V, int nV, const int* __restrict__ F,
                                const int* __restrict__ E, int nE, int nEdges, const float* __restrict__ rest,
                                const float* __restrict__ lambda, float* __restrict__ out) {
  const int e = blockIdx.x * kBlock + threadIdx.x;
  if (e >= nEdges) return;
  const int2 v = edge_endpoints(kind, e, F, E, nE);
  const bool ok = (unsigned)v.x < (unsigned)nV && (unsigned)v.y < (unsigned)nV;
  const float lam = lambda ? lambda[e] : 1.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float l = 0.f;
    if (ok) {
      l = fsub(fsub(__ldg(V + 3 * (size_t)v.y + k), __ldg(V + 3 * (size_t)v.x + k)), rest[3 * (size_t)e + k]);   // rigid_layer.cc:80-82
      if (lambda) l = fmul(l, lam);                                                                              // cad_layer.cc:117-122
    }
    out[3 * (size_t)e + k] = fmul(l, l);
  }
}

// Per-incidence records in CSR order: (rest vector of the edge, other endpoint | side << 31), 16 bytes.  The
// backward pass of a vertex then streams its own contiguous records and gathers only V[other] (12 B) per
// incidence, instead of key -> edge endpoints -> both vertices + rest vector (four dependent gathers).
__global__ void k_csr_records(const int* __restrict__ keys, const int* __restrict__ start, int nV,
                              const int2* __restrict__ ev, const float* __restrict__ rest,
                              const float* __restrict__ lambda, float4* __restrict__ inc, float* __restrict__ inc_lambda) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= start[nV]) return;   // edges with an out-of-range endpoint own no keys
  const int key = keys[i];
  const int e = key >> 1, side = key & 1;
  const int2 v = ev[e];
  const unsigned other = (unsigned)(side ? v.x : v.y) | ((unsigned)side << 31);
  inc[i] = make_float4(rest[3 * (size_t)e], rest[3 * (size_t)e + 1], rest[3 * (size_t)e + 2], __uint_as_float(other));
  if (inc_lambda) inc_lambda[i] = lambda[e];
}

// one vertex's gradient from its records: the same operations in the same (ascending edge, side) order as
// gather_vertex below, i.e. as the reference's serial scatter loop
template <bool COHERENT = false>   // COHERENT: V changes during the kernel (persistent loop), read it through L2
__device__ __forceinline__ void gather_vertex_rec(const float* __restrict__ V, const float self[3],
                                                  const float4* __restrict__ inc, const float* __restrict__ inc_lambda,
                                                  int kb, int ke, float acc[3], double* edge_loss) {
  constexpr int B = 4;   // incidences whose record and endpoint loads are in flight together (accumulated in order)
  for (int i0 = kb; i0 < ke; i0 += B) {
    float4 rec[B];
    float lamv[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int i = min(i0 + j, ke - 1);
      rec[j] = __ldg(inc + i);
      lamv[j] = inc_lambda ? __ldg(inc_lambda + i) : 1.f;
    }
    float o[B][3];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const float* po = V + 3 * (size_t)(__float_as_uint(rec[j].w) & 0x7fffffffu);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[j][k] = COHERENT ? __ldcg(po + k) : __ldg(po + k);
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
      if (i0 + j < ke) {
        const bool side = (__float_as_uint(rec[j].w) >> 31) != 0u;
        const float lam = lamv[j];
        const float lam2 = inc_lambda ? fmul(lam, lam) : 1.f;   // cad_layer.cc:186-187
        const float rs[3] = {rec[j].x, rec[j].y, rec[j].z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          // r = (V[v1] - V[v0]) - rest with (v0, v1) = (self, other) on side 0 and (other, self) on side 1
          float r = fsub(side ? fsub(self[k], o[j][k]) : fsub(o[j][k], self[k]), rs[k]);
          if (edge_loss && !side) { const float l = inc_lambda ? fmul(r, lam) : r; *edge_loss += (double)fmul(l, l); }
          if (inc_lambda) r = fmul(r, lam2);
          acc[k] = side ? fadd(acc[k], r) : fsub(acc[k], r);        // rigid_layer.cc:123-128
        }
      }
    }
  }
}

__global__ void k_edges_backward_csr(const float* __restrict__ V, int nV, const float4* __restrict__ inc,
                                     const float* __restrict__ inc_lambda, const int* __restrict__ start,
                                     float* __restrict__ grad) {
  const int v = blockIdx.x * kBlock + threadIdx.x;
  if (v >= nV) return;
  float acc[3] = {0.f, 0.f, 0.f};
  const float self[3] = {V[3 * (size_t)v], V[3 * (size_t)v + 1], V[3 * (size_t)v + 2]};
  gather_vertex_rec(V, self, inc, inc_lambda, start[v], start[v + 1], acc, nullptr);
  grad[3 * (size_t)v] = acc[0]; grad[3 * (size_t)v + 1] = acc[1]; grad[3 * (size_t)v + 2] = acc[2];
}

// edge-parallel scatter; lanes of a warp that hit the same vertex are summed first
__global__ void k_edges_backward_atomic(int kind, const float* __restrict__ V, int nV, const int* __restrict__ F,
                                        const int* __restrict__ E, int nE, int nEdges, const float* __restrict__ rest,
                                        const float* __restrict__ lambda, float* __restrict__ grad) {
  const int e = blockIdx.x * kBlock + threadIdx.x;
  int2 v = make_int2(-1, -1);
  float r[3] = {0.f, 0.f, 0.f};
  if (e < nEdges) {
    v = edge_endpoints(kind, e, F, E, nE);
    if ((unsigned)v.x < (unsigned)nV && (unsigned)v.y < (unsigned)nV) {
      float lam2 = 1.f;
      if (lambda) { const float lam = lambda[e]; lam2 = fmul(lam, lam); }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        r[k] = fsub(fsub(__ldg(V + 3 * (size_t)v.y + k), __ldg(V + 3 * (size_t)v.x + k)), rest[3 * (size_t)e + k]);
        if (lambda) r[k] = fmul(r[k], lam2);
      }
    } else {
      v = make_int2(-1, -1);
    }
  }
  const cg::coalesced_group active = cg::coalesced_threads();
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int tgt = side ? v.y : v.x;
    const cg::coalesced_group same = cg::labeled_partition(active, tgt);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float s = cg::reduce(same, side ? r[k] : -r[k], cg::plus<float>());
      if (tgt >= 0 && same.thread_rank() == 0) atomicAdd(grad + 3 * (size_t)tgt + k, s);
    }
  }
}

// fused per-iteration loss: distance (Jet) + CSR edge gather + masked/weighted sum
__global__ void __launch_bounds__(kBlock) k_loss_fused(const float* __restrict__ grid, int n, const float* __restrict__ V,
                                                       int nV, const float4* __restrict__ inc,
                                                       const float* __restrict__ inc_lambda, const int* __restrict__ start,
                                                       float w_edge, float mask_thr, double* __restrict__ loss,
                                                       float* __restrict__ grad) {
  __shared__ double s_part[kBlock / 32];
  const int v = blockIdx.x * kBlock + threadIdx.x;
  double my = 0.0;
  if (v < nV) {
    typedef Jet3<float> J;
    const float x = V[3 * (size_t)v], y = V[3 * (size_t)v + 1], z = V[3 * (size_t)v + 2];
    J vd = sample<J, float>(grid, n, J(x, 1.f, 0.f, 0.f), J(y, 0.f, 1.f, 0.f), J(z, 0.f, 0.f, 1.f));
    vd = vd * vd;
    const float lossD = fmul(vd.a, 0.5f);                                    // rigid_loss_layer.py:11
    float gD[3] = {(float)((double)vd.v0 * 0.5), (float)((double)vd.v1 * 0.5), (float)((double)vd.v2 * 0.5)};
    if (mask_thr > 0.f && !(lossD < mask_thr)) gD[0] = gD[1] = gD[2] = 0.f;   // graph_loss_layer.py:18,40
    float gE[3] = {0.f, 0.f, 0.f};
    double le = 0.0;
    const float self[3] = {x, y, z};
    if (start) gather_vertex_rec(V, self, inc, inc_lambda, start[v], start[v + 1], gE, loss ? &le : nullptr);
    my = (double)lossD + 0.5 * le * (double)w_edge;
    if (grad) {
#pragma unroll
      for (int k = 0; k < 3; ++k) grad[3 * (size_t)v + k] = fadd(gD[k], fmul(gE[k], w_edge));   // rigid_loss_layer.py:27
    }
  }
  if (loss) {
    for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < kBlock / 32; ++i) t += s_part[i];
      atomicAdd(loss, t);
    }
  }
}

// The whole Adam loop of one mesh of any size in ONE cooperative launch (the launch-bound alternative is two
// launches per iteration): every thread owns the vertices gtid + k*T, positions are double buffered in global
// memory (a vertex's new position goes to the other buffer because its neighbours still gather the old one),
// one grid-wide barrier per iteration.  The arithmetic is k_loss_fused followed by the Adam step of deform.cu,
// operation for operation.  Vertex reads go through L2 (__ldcg): the read-only path is not coherent with the
// stores of other SMs.
__global__ void __launch_bounds__(kBlock) k_adam_loop_coop(const float* __restrict__ grid, int n, float* bufA, float* bufB,
                                                           int nV, const float4* __restrict__ inc,
                                                           const float* __restrict__ inc_lambda,
                                                           const int* __restrict__ start, float w_edge, float mask_thr,
                                                           const float2* __restrict__ sched, int iters, float w1, float b2,
                                                           float w2, float eps, float* __restrict__ mom) {
  cg::grid_group gridg = cg::this_grid();
  const int T = gridDim.x * kBlock;
  const int gtid = blockIdx.x * kBlock + threadIdx.x;
  float* cur = bufA;
  float* nxt = bufB;
  for (int it = 0; it < iters; ++it) {
    const float2 sc = __ldg(&sched[it]);
    for (int v = gtid; v < nV; v += T) {
      typedef Jet3<float> J;
      // everything that does not depend on another load is requested first: the iteration is a chain of L2 round trips
      const float p[3] = {__ldcg(cur + 3 * (size_t)v), __ldcg(cur + 3 * (size_t)v + 1), __ldcg(cur + 3 * (size_t)v + 2)};
      const int kb = start ? __ldg(start + v) : 0, ke = start ? __ldg(start + v + 1) : 0;
      float m0[3], v0[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { m0[k] = __ldcg(mom + 3 * (size_t)v + k); v0[k] = __ldcg(mom + 3 * (size_t)nV + 3 * (size_t)v + k); }
      J vd = sample<J, float>(grid, n, J(p[0], 1.f, 0.f, 0.f), J(p[1], 0.f, 1.f, 0.f), J(p[2], 0.f, 0.f, 1.f));
      vd = vd * vd;
      const float lossD = fmul(vd.a, 0.5f);
      float gD[3] = {(float)((double)vd.v0 * 0.5), (float)((double)vd.v1 * 0.5), (float)((double)vd.v2 * 0.5)};
      if (mask_thr > 0.f && !(lossD < mask_thr)) gD[0] = gD[1] = gD[2] = 0.f;
      float gE[3] = {0.f, 0.f, 0.f};
      if (start) gather_vertex_rec<true>(cur, p, inc, inc_lambda, kb, ke, gE, nullptr);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float g = fadd(gD[k], fmul(gE[k], w_edge));                    // rigid_loss_layer.py:27
        const float mi = __fmaf_rn(w1, fsub(g, m0[k]), m0[k]);                 // torch's _single_tensor_adam, as k_adam_step
        const float vi = __fmaf_rn(fmul(w2, g), g, fmul(v0[k], b2));
        __stcg(mom + 3 * (size_t)v + k, mi); __stcg(mom + 3 * (size_t)nV + 3 * (size_t)v + k, vi);
        const float denom = fadd(__fdiv_rn(__fsqrt_rn(vi), sc.y), eps);
        __stcg(nxt + 3 * (size_t)v + k, fadd(p[k], __fdiv_rn(fmul(sc.x, mi), denom)));
      }
    }
    gridg.sync();
    float* t = cur; cur = nxt; nxt = t;
  }
  if (cur != bufA) {   // odd iteration count: the result sits in the scratch buffer
    for (int i = gtid; i < 3 * nV; i += T) bufA[i] = __ldcg(cur + i);
  }
}

}  // namespace

static int ensure_records(const Template& Tc, cudaStream_t s);

// Cooperative persistent Adam loop; returns MO_ERR_STATE when the device cannot co-schedule the grid (the
// caller then falls back to two launches per iteration).
int adam_loop_coop(const Template& TD, const Template* TE, float* d_V, int nV, float w_edge, float mask_thr,
                   const float2* d_sched, int iters, float w1, float b2, float w2, float eps, float* d_scratch,
                   cudaStream_t s) {
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  MO_CUDA(cudaGetDevice(&dev));
  MO_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) return MO_ERR_STATE;
  MO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_adam_loop_coop, kBlock, 0));
  if (per_sm < 1) return MO_ERR_STATE;
  if (TE) {
    const int rc = ensure_records(*TE, s);
    if (rc != MO_OK) return rc;
  }
  const int blocks = std::min(div_up(nV, kBlock), sms * std::min(per_sm, 2));
  const float* grid = TD.d_grid32;
  int n = TD.N;
  float* bufB = d_scratch;                       // [3*nV] second position buffer
  float* mom = d_scratch + 3 * (size_t)nV;       // [6*nV] Adam moments (zeroed by the caller)
  const float4* inc = TE ? TE->d_inc : nullptr;
  const float* incl = TE ? TE->d_inc_lambda : nullptr;
  const int* start = TE ? TE->d_csr_start : nullptr;
  void* args[] = {(void*)&grid, (void*)&n, (void*)&d_V, (void*)&bufB, (void*)&nV, (void*)&inc, (void*)&incl, (void*)&start,
                  (void*)&w_edge, (void*)&mask_thr, (void*)&d_sched, (void*)&iters, (void*)&w1, (void*)&b2, (void*)&w2,
                  (void*)&eps, (void*)&mom};
  MO_CUDA(cudaLaunchCooperativeKernel((const void*)k_adam_loop_coop, dim3(blocks), dim3(kBlock), args, 0, s));
  MO_LAUNCH_CHECK();
  return MO_OK;
}

void free_edges(Template& T, cudaStream_t s) {
  dev_free(T.d_ev, s); dev_free(T.d_rest, s); dev_free(T.d_lambda, s); dev_free(T.d_csr_start, s); dev_free(T.d_csr_key, s);
  dev_free(T.d_v0, s); dev_free(T.d_ell, s); dev_free(T.d_nbr, s); dev_free(T.d_inc, s); dev_free(T.d_inc_lambda, s);
  T.d_inc = nullptr; T.d_inc_lambda = nullptr;
  T.d_ev = nullptr; T.d_rest = nullptr; T.d_lambda = nullptr; T.d_csr_start = nullptr; T.d_csr_key = nullptr;
  T.d_v0 = nullptr; T.d_ell = nullptr; T.ell_D = 0; T.d_nbr = nullptr; T.nbr_W = 0;
  T.kind = MO_EDGES_NONE; T.nEdges = 0;
}

static int edge_count(int kind, int nF, int nE) {
  return kind == MO_EDGES_RIGID ? 3 * nF : (kind == MO_EDGES_GRAPH ? nE : nE + 3 * nF);
}

int edges_store(Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                cudaStream_t s) {
  const int nEdges = edge_count(kind, nF, nE);
  // storage is re-used when the shape is unchanged (every iteration of a re-initialising caller)
  if (T.kind != kind || T.nEdges != nEdges || T.eV != nV) {
    free_edges(T, s);
    MO_CUDA(dev_alloc(&T.d_ev, (size_t)std::max(nEdges, 1), s));
    MO_CUDA(dev_alloc(&T.d_rest, 3 * (size_t)std::max(nEdges, 1), s));
    if (kind == MO_EDGES_CAD) MO_CUDA(dev_alloc(&T.d_lambda, (size_t)std::max(nEdges, 1), s));
    MO_CUDA(dev_alloc(&T.d_csr_start, (size_t)nV + 1, s));
    MO_CUDA(dev_alloc(&T.d_csr_key, 2 * (size_t)std::max(nEdges, 1), s));
    MO_CUDA(dev_alloc(&T.d_v0, 3 * (size_t)std::max(nV, 1), s));
  }
  if (T.d_ell) { dev_free(T.d_ell, s); T.d_ell = nullptr; T.ell_D = 0; }
  if (T.d_nbr) { dev_free(T.d_nbr, s); T.d_nbr = nullptr; T.nbr_W = 0; }
  if (T.d_inc) { dev_free(T.d_inc, s); T.d_inc = nullptr; }
  if (T.d_inc_lambda) { dev_free(T.d_inc_lambda, s); T.d_inc_lambda = nullptr; }
  if (nV > 0) MO_CUDA(cudaMemcpyAsync(T.d_v0, d_V, sizeof(float) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s));
  T.kind = kind; T.nEdges = nEdges; T.eV = nV; T.eF = nF; T.eE = nE;
  int* deg = nullptr;   // [nV] degree + [nV] fill cursor
  MO_CUDA(cudaMallocAsync(&deg, sizeof(int) * 2 * ((size_t)nV + 1), s));
  MO_CUDA(cudaMemsetAsync(deg, 0, sizeof(int) * 2 * ((size_t)nV + 1), s));
  if (nEdges > 0) {
    k_edges_store<<<div_up(nEdges, kBlock), kBlock, 0, s>>>(kind, d_V, nV, d_F, d_E, nE, nEdges, T.d_ev, T.d_rest, T.d_lambda,
                                                            deg, T.d_stats);
    MO_LAUNCH_CHECK();
  }
  k_scan1<<<1, 1024, 0, s>>>(deg, T.d_csr_start, nV);
  MO_LAUNCH_CHECK();
  if (nEdges > 0) {
    k_csr_fill<<<div_up(nEdges, kBlock), kBlock, 0, s>>>(T.d_ev, nEdges, T.d_csr_start, deg + nV + 1, T.d_csr_key);
    MO_LAUNCH_CHECK();
    k_csr_sort<<<div_up(nV, kBlock), kBlock, 0, s>>>(T.d_csr_start, nV, T.d_csr_key);
    MO_LAUNCH_CHECK();
  }
  MO_CUDA(cudaFreeAsync(deg, s));
  return MO_OK;
}

int edges_forward(const Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                  float* d_out, cudaStream_t s) {
  const int nEdges = edge_count(kind, nF, nE);
  if (nEdges == 0) return MO_OK;
  k_edges_forward<<<div_up(nEdges, kBlock), kBlock, 0, s>>>(kind, d_V, nV, d_F, d_E, nE, nEdges, T.d_rest, T.d_lambda, d_out);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

// per-incidence records of the stored edge set, built on the first backward pass that needs them
static int ensure_records(const Template& Tc, cudaStream_t s) {
  Template& T = const_cast<Template&>(Tc);   // a cache of data derived from the stored edges
  if (T.d_inc || T.nEdges == 0) return MO_OK;
  const int nKeys = 2 * T.nEdges;
  MO_CUDA(dev_alloc(&T.d_inc, (size_t)nKeys, s));
  if (T.d_lambda) MO_CUDA(dev_alloc(&T.d_inc_lambda, (size_t)nKeys, s));
  k_csr_records<<<div_up(nKeys, kBlock), kBlock, 0, s>>>(T.d_csr_key, T.d_csr_start, T.eV, T.d_ev, T.d_rest, T.d_lambda,
                                                         T.d_inc, T.d_inc_lambda);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

int edges_backward(const Template& T, const float* d_V, int nV, float* d_grad, cudaStream_t s) {
  if (nV == 0) return MO_OK;
  const int rc = ensure_records(T, s);
  if (rc != MO_OK) return rc;
  k_edges_backward_csr<<<div_up(nV, kBlock), kBlock, 0, s>>>(d_V, nV, T.d_inc, T.d_inc_lambda, T.d_csr_start, d_grad);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

int edges_backward_atomic(const Template& T, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                          int nE, float* d_grad, cudaStream_t s) {
  const int nEdges = edge_count(kind, nF, nE);
  MO_CUDA(cudaMemsetAsync(d_grad, 0, sizeof(float) * 3 * (size_t)nV, s));
  if (nEdges == 0) return MO_OK;
  k_edges_backward_atomic<<<div_up(nEdges, kBlock), kBlock, 0, s>>>(kind, d_V, nV, d_F, d_E, nE, nEdges, T.d_rest, T.d_lambda,
                                                                    d_grad);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

int loss_fused(const Template& TD, const Template* TE, const float* d_V, int nV, float w_edge, float mask_thr,
               double* d_loss, float* d_grad, cudaStream_t s) {
  if (d_loss) MO_CUDA(cudaMemsetAsync(d_loss, 0, sizeof(double), s));
  if (nV == 0) return MO_OK;
  if (TE) {
    const int rc = ensure_records(*TE, s);
    if (rc != MO_OK) return rc;
  }
  k_loss_fused<<<div_up(nV, kBlock), kBlock, 0, s>>>(TD.d_grid32, TD.N, d_V, nV, TE ? TE->d_inc : nullptr,
                                                     TE ? TE->d_inc_lambda : nullptr, TE ? TE->d_csr_start : nullptr, w_edge,
                                                     mask_thr, d_loss, d_grad);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

}  // namespace mo
