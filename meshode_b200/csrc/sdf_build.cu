// Distance-field build for sm_100a: Mesh::Normalize + Mesh::ConstructDistanceField
// (reference src/lib/mesh.cc:66-85, :106-152; the nearest-triangle search that the
// reference delegates to igl::point_mesh_squared_distance at mesh.cc:140).
//
// Pipeline (all on the caller's stream, no host synchronisation):
//   k_bbox / k_xform / k_vnorm   FP64 normalisation, bit-identical to the CPU arithmetic
//   k_tri_count / k_scan / k_tri_fill
//                                triangles binned by centroid into cells of 4^3 voxels;
//                                per-cell tight AABB; records written cell-sorted
//   k_sdf_tiles                  one CTA per 8^3-voxel tile, one lane per voxel:
//                                cells are swept ring by ring around the tile, culled
//                                against the tile's running upper bound; surviving
//                                triangle records are staged in shared memory; each warp
//                                (4x4x2 voxels) culls staged candidates against its own
//                                bound (lanes over candidates, warp-shuffle max / ballot)
//                                and runs the dense branch-free FP32 point-triangle test
//                                on the survivors (lanes over voxels, record broadcast from
//                                shared memory).  Candidates whose FP32 lower bound is
//                                within the rigorous error band of the running minimum are
//                                queued per lane and re-evaluated exactly in FP64
//                                (Ericson's closest point, the oracle's arithmetic), so the
//                                stored distance and nearest index are the FP64 result.
#include <cfloat>

#include "common.cuh"

namespace mo {
namespace {

constexpr int kTile = 8;          // voxels per tile edge in x and y
constexpr int kTileZ = 4;         // voxels per tile in z: 8 warps per CTA, two CTAs per SM hide each other's barriers
constexpr int kCellVox = 4;       // voxels per bin-cell edge
constexpr int kWarps = kTile * kTile * kTileZ / 32;
constexpr int kThreads = kWarps * 32;
constexpr int kCtasPerSm = 2;
constexpr int kCap = 768;         // triangle records staged per chunk (64 B record + 16 B bounding sphere each)
constexpr int kMaxRanges = 1024;  // cell ranges collected per pass
constexpr int kListCap = 12;      // per-lane queue of FP64 candidates

// |q_fp32 - q_exact| <= kA * |p-a|^2 + kB for coordinates inside the unit cube: record
// rounding moves the triangle by <= 3e-8 (=> 1.1e-7*sqrt(pp) <= 5e-6*pp + 5e-10), the
// arithmetic adds a few ulp of pp.  Both constants carry a >2x margin.
constexpr float kErrA = 1.2e-5f;
constexpr float kErrB = 1.2e-9f;

struct SdfArgs {
  int N, nc, ntile, ntz, tz0, z0, z1;
  float cs;                       // cell size in normalised units
  const unsigned* max_ext;        // bit pattern of the largest triangle AABB extent
  const int* cell_start;          // [ncell+1]
  const unsigned* cell_bb;        // [ncell*6] ordered-uint lo xyz, hi xyz
  const float4* rec32;            // [nF*4] cell-sorted FP32 records
  const float4* sph;              // [nF] cell-sorted bounding spheres (centre, radius rounded up)
  const double* rec64;            // [nF*9] cell-sorted FP64 vertices
  const int* tri_id;              // [nF] cell-sorted -> original triangle index
  double* grid64;
  float* grid32;
  int* nearest;
  unsigned long long* stats;
};

// ---------------------------------------------------------------------------------
// normalisation
// ---------------------------------------------------------------------------------
__global__ void k_bbox(const float* __restrict__ V, int nV, unsigned* __restrict__ bb) {
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};   // mesh.cc:69-71
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nV; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float x = V[3 * (size_t)i + j];
      if (x < mn[j]) mn[j] = x;   // explicit compares: NaNs are ignored like mesh.cc:73-76
      if (x > mx[j]) mx[j] = x;
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[j] = fminf(mn[j], __shfl_xor_sync(0xffffffffu, mn[j], o));
      mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      atomicMin(&bb[j], f2o(mn[j]));
      atomicMax(&bb[3 + j], f2o(mx[j]));
    }
  }
}

__global__ void k_xform(const unsigned* __restrict__ bb, double* __restrict__ xf) {
  double mn[3], mx[3];
  for (int j = 0; j < 3; ++j) { mn[j] = (double)o2f(bb[j]); mx[j] = (double)o2f(bb[3 + j]); }
  const double e0 = dsub(mx[0], mn[0]), e1 = dsub(mx[1], mn[1]), e2 = dsub(mx[2], mn[2]);
  const double m12 = e1 < e2 ? e2 : e1;            // std::max(a,b) = (a<b)?b:a
  const double m = e0 < m12 ? m12 : e0;
  const double scale = dmul(m, 1.1);               // mesh.cc:80-81
  xf[0] = scale;
  for (int j = 0; j < 3; ++j) xf[1 + j] = dsub(mn[j], dmul(0.05, scale));   // mesh.cc:82-83
}

__global__ void k_vnorm(const float* __restrict__ V, int n3, const double* __restrict__ xf, double* __restrict__ Vn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const double scale = xf[0], pos = xf[1 + i % 3];
  Vn[i] = __ddiv_rn(dsub((double)V[i], pos), scale);   // mesh.cc:84-85
}

__global__ void k_fill_grid(double* g64, float* g32, int* nearest, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  g64[i] = 1e30; g32[i] = 1e30f; nearest[i] = -1;   // uniformgrid.cc:9-17
}

// ---------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------
__global__ void k_init_cells(unsigned* __restrict__ cell_bb, int ncell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
#pragma unroll
  for (int j = 0; j < 3; ++j) { cell_bb[6 * (size_t)c + j] = 0xffffffffu; cell_bb[6 * (size_t)c + 3 + j] = 0u; }
}

__device__ __forceinline__ int cell_coord(double x, int N, int nc) {
  const double c = floor(x * (double)N * (1.0 / kCellVox));
  return c < 0.0 ? 0 : (c > (double)(nc - 1) ? nc - 1 : (int)c);
}

__global__ void k_tri_count(const double* __restrict__ Vn, const int* __restrict__ F, int nF, int nV, int N, int nc,
                            int* __restrict__ cell_count, unsigned* __restrict__ cell_bb, int* __restrict__ tri_cell,
                            unsigned* __restrict__ max_ext, unsigned long long* __restrict__ stats) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nF) return;
  const int i0 = F[3 * (size_t)t], i1 = F[3 * (size_t)t + 1], i2 = F[3 * (size_t)t + 2];
  if ((unsigned)i0 >= (unsigned)nV || (unsigned)i1 >= (unsigned)nV || (unsigned)i2 >= (unsigned)nV) {
    tri_cell[t] = -1; atomicOr(&stats[3], 1ull); return;
  }
  double lo[3], hi[3], ce[3];
  bool finite = true;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double a = Vn[3 * (size_t)i0 + j], b = Vn[3 * (size_t)i1 + j], c = Vn[3 * (size_t)i2 + j];
    lo[j] = fmin(a, fmin(b, c)); hi[j] = fmax(a, fmax(b, c)); ce[j] = (a + b + c) * (1.0 / 3.0);
    finite = finite && isfinite(a) && isfinite(b) && isfinite(c);
  }
  if (!finite) { tri_cell[t] = -1; atomicOr(&stats[3], 2ull); return; }
  const int cx = cell_coord(ce[0], N, nc), cy = cell_coord(ce[1], N, nc), cz = cell_coord(ce[2], N, nc);
  const int c = (cz * nc + cy) * nc + cx;
  float ext = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float l = __double2float_rd(lo[j]), h = __double2float_ru(hi[j]);
    atomicMin(&cell_bb[6 * (size_t)c + j], f2o(l));
    atomicMax(&cell_bb[6 * (size_t)c + 3 + j], f2o(h));
    ext = fmaxf(ext, __fsub_ru(h, l));
  }
  atomicMax(max_ext, __float_as_uint(ext));
  atomicAdd(&cell_count[c], 1);
  tri_cell[t] = c;
}

// exclusive scan of n ints by one CTA of 1024 threads
__global__ void k_scan(const int* __restrict__ in, int* __restrict__ out, int n) {
  __shared__ int s_warp[32];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b = tid * per, e = min(n, b + per);
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

__global__ void k_tri_fill(const double* __restrict__ Vn, const int* __restrict__ F, int nF,
                           const int* __restrict__ tri_cell, const int* __restrict__ cell_start,
                           int* __restrict__ cell_fill, float4* __restrict__ rec32, float4* __restrict__ sph,
                           double* __restrict__ rec64, int* __restrict__ tri_id) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nF) return;
  const int c = tri_cell[t];
  if (c < 0) return;
  const int slot = cell_start[c] + atomicAdd(&cell_fill[c], 1);
  const int i0 = F[3 * (size_t)t], i1 = F[3 * (size_t)t + 1], i2 = F[3 * (size_t)t + 2];
  double a[3], b[3], cc[3], ab[3], ac[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    a[j] = Vn[3 * (size_t)i0 + j]; b[j] = Vn[3 * (size_t)i1 + j]; cc[j] = Vn[3 * (size_t)i2 + j];
    ab[j] = b[j] - a[j]; ac[j] = cc[j] - a[j];
    rec64[9 * (size_t)slot + j] = a[j]; rec64[9 * (size_t)slot + 3 + j] = b[j]; rec64[9 * (size_t)slot + 6 + j] = cc[j];
  }
  const double e11 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  const double e12 = ab[0] * ac[0] + ab[1] * ac[1] + ab[2] * ac[2];
  const double e22 = ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2];
  const double bc2 = e11 - 2.0 * e12 + e22;
  const double det = e11 * e22 - e12 * e12;
  // "regular": sin^2 of the angle at a is > 1e-3 and no edge is degenerate.  Otherwise the
  // interior test is skipped (boundary distance only) and the inradius widens the error band.
  const bool regular = (e11 > 0.0) && (e22 > 0.0) && (bc2 > 0.0) && (det > 1e-3 * e11 * e22);
  float r3x;
  if (regular) {
    r3x = (float)(1.0 / det);
    if (!(r3x > 0.f) || isinf(r3x)) r3x = -0.f;
  } else {
    const double per = sqrt(fmax(e11, 0.0)) + sqrt(fmax(e22, 0.0)) + sqrt(fmax(bc2, 0.0));
    const double rin = per > 0.0 ? sqrt(fmax(det, 0.0)) / per : 0.0;   // 2*Area / perimeter
    r3x = -__double2float_ru(rin * 1.001);
  }
  const float i11 = e11 > 0.0 ? (float)(1.0 / e11) : 0.f;
  const float i22 = e22 > 0.0 ? (float)(1.0 / e22) : 0.f;
  const float ibc = bc2 > 0.0 ? (float)(1.0 / bc2) : 0.f;
  rec32[4 * (size_t)slot + 0] = make_float4((float)a[0], (float)a[1], (float)a[2], (float)e11);
  rec32[4 * (size_t)slot + 1] = make_float4((float)ab[0], (float)ab[1], (float)ab[2], (float)e12);
  rec32[4 * (size_t)slot + 2] = make_float4((float)ac[0], (float)ac[1], (float)ac[2], (float)e22);
  rec32[4 * (size_t)slot + 3] = make_float4(r3x, isinf(i11) ? 0.f : i11, isinf(i22) ? 0.f : i22, isinf(ibc) ? 0.f : ibc);
  // bounding sphere about the centroid; the radius absorbs the float rounding of the centre
  {
    double cx[3], r2 = 0.0;
    float cf[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { cx[j] = (a[j] + b[j] + cc[j]) * (1.0 / 3.0); cf[j] = (float)cx[j]; }
    const double* vs[3] = {a, b, cc};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double d2 = 0.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) { const double d = vs[k][j] - (double)cf[j]; d2 += d * d; }
      r2 = fmax(r2, d2);
    }
    // + 2e-7: float rounding of the query point (<= 5.2e-8) and of the centre, with margin
    sph[slot] = make_float4(cf[0], cf[1], cf[2], __double2float_ru(sqrt(r2) * 1.000001 + 2e-7));
  }
  tri_id[slot] = t;
}

// ---------------------------------------------------------------------------------
// FP32 point-triangle squared distance, branch-free apart from one warp-uniform test.
// Record: r0 = (a, e11) r1 = (ab, e12) r2 = (ac, e22) r3 = (1/det | -inradius, 1/e11, 1/e22, 1/|bc|^2)
// Returns q >= exact - err and sets err so that |q - exact| <= err for regular triangles,
// exact in [q - err, q] for flagged ones.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float tri_q(const float4 r0, const float4 r1, const float4 r2, const float4 r3,
                                       const float px, const float py, const float pz, float& err) {
  const float apx = px - r0.x, apy = py - r0.y, apz = pz - r0.z;
  const float pp = fmaf(apz, apz, fmaf(apy, apy, apx * apx));
  const float d1 = fmaf(r1.z, apz, fmaf(r1.y, apy, r1.x * apx));
  const float d2 = fmaf(r2.z, apz, fmaf(r2.y, apy, r2.x * apx));
  const float e11 = r0.w, e12 = r1.w, e22 = r2.w;
  const float m2d1 = -2.f * d1, m2d2 = -2.f * d2;
  // closest points on the three edges (clamped parameters)
  const float t1 = __saturatef(d1 * r3.y);
  const float q1 = fmaf(t1, fmaf(t1, e11, m2d1), pp);
  const float t2 = __saturatef(d2 * r3.z);
  const float q2 = fmaf(t2, fmaf(t2, e22, m2d2), pp);
  const float g = (d2 - d1) + (e11 - e12);           // bc . bp
  const float bc2 = fmaf(-2.f, e12, e11 + e22);
  const float u = __saturatef(g * r3.w);
  const float bp2 = (pp + m2d1) + e11;
  const float q3 = fmaf(u, fmaf(u, bc2, -2.f * g), bp2);
  const float qe = fminf(q1, fminf(q2, q3));
  err = fmaf(pp, kErrA, kErrB);
  float q;
  if (r3.x > 0.f) {   // same record for every lane: uniform branch
    const float s = (e22 * d1 - e12 * d2) * r3.x;
    const float t = (e11 * d2 - e12 * d1) * r3.x;
    // full quadratic form: second-order insensitive to errors in (s,t)
    const float us = fmaf(s, e11, fmaf(2.f * t, e12, m2d1));
    const float ut = fmaf(t, e22, m2d2);
    const float qf = fmaf(t, ut, fmaf(s, us, pp));
    const bool inside = (s >= 0.f) && (t >= 0.f) && (s + t <= 1.f);
    q = inside ? qf : qe;
  } else {
    q = qe;
    err = fmaf(-2.f * r3.x, sqrtf(fmaxf(qe, 0.f)), err);
  }
  return fmaxf(q, 0.f);
}

// Exact FP64 closest point (Ericson 5.1.5), the same operation sequence as the host
// oracle's point_triangle_sqr, with contraction-free arithmetic.
__device__ __forceinline__ double ddot(const double* a, const double* b) {
  return dadd(dadd(dmul(a[0], b[0]), dmul(a[1], b[1])), dmul(a[2], b[2]));
}
__device__ __forceinline__ double dsafe_div(double n, double d) { return d != 0.0 ? __ddiv_rn(n, d) : 0.0; }

__device__ __noinline__ double tri_exact64(const double* __restrict__ tv, const double px, const double py,
                                           const double pz) {
  double a[3], b[3], c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { a[j] = tv[j]; b[j] = tv[3 + j]; c[j] = tv[6 + j]; }
  const double p[3] = {px, py, pz};
  double ab[3], ac[3], ap[3], q[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { ab[j] = dsub(b[j], a[j]); ac[j] = dsub(c[j], a[j]); ap[j] = dsub(p[j], a[j]); }
  const double d1 = ddot(ab, ap), d2 = ddot(ac, ap);
  bool done = false;
  if (d1 <= 0.0 && d2 <= 0.0) { q[0] = a[0]; q[1] = a[1]; q[2] = a[2]; done = true; }
  double d3 = 0, d4 = 0, d5 = 0, d6 = 0, vc = 0, vb = 0;
  if (!done) {
    double bp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) bp[j] = dsub(p[j], b[j]);
    d3 = ddot(ab, bp); d4 = ddot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; done = true; }
  }
  if (!done) {
    vc = dsub(dmul(d1, d4), dmul(d3, d2));
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
      const double v = dsafe_div(d1, dsub(d1, d3));
#pragma unroll
      for (int j = 0; j < 3; ++j) q[j] = dadd(a[j], dmul(v, ab[j]));
      done = true;
    }
  }
  if (!done) {
    double cp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) cp[j] = dsub(p[j], c[j]);
    d5 = ddot(ab, cp); d6 = ddot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { q[0] = c[0]; q[1] = c[1]; q[2] = c[2]; done = true; }
  }
  if (!done) {
    vb = dsub(dmul(d5, d2), dmul(d1, d6));
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
      const double w = dsafe_div(d2, dsub(d2, d6));
#pragma unroll
      for (int j = 0; j < 3; ++j) q[j] = dadd(a[j], dmul(w, ac[j]));
      done = true;
    }
  }
  if (!done) {
    const double va = dsub(dmul(d3, d6), dmul(d5, d4));
    const double d43 = dsub(d4, d3), d56 = dsub(d5, d6);
    if (va <= 0.0 && d43 >= 0.0 && d56 >= 0.0) {
      const double w = dsafe_div(d43, dadd(d43, d56));
#pragma unroll
      for (int j = 0; j < 3; ++j) q[j] = dadd(b[j], dmul(w, dsub(c[j], b[j])));
    } else {
      const double sum = dadd(dadd(va, vb), vc);
      if (sum != 0.0) {
        const double denom = __ddiv_rn(1.0, sum);
        const double v = dmul(vb, denom), w = dmul(vc, denom);
#pragma unroll
        for (int j = 0; j < 3; ++j) q[j] = dadd(dadd(a[j], dmul(ab[j], v)), dmul(ac[j], w));
      } else {
        q[0] = a[0]; q[1] = a[1]; q[2] = a[2];
      }
    }
  }
  const double dx = dsub(p[0], q[0]), dy = dsub(p[1], q[1]), dz = dsub(p[2], q[2]);
  return dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
}

struct LaneState {
  double best64;   // exact minimum so far
  int best_id;     // its original triangle index (lowest on exact ties)
  float ub;        // rigorous FP32 upper bound of the exact minimum
  float sub;       // upper bound of sqrt(ub) (for the bounding-sphere pre-test)
  int cnt;         // queued FP64 candidates
  unsigned n64;
};

__device__ __forceinline__ void flush_queue(LaneState& st, const int* s_lid, const float* s_lq, const int tid,
                                            const SdfArgs& A, const double px, const double py, const double pz) {
  for (int k = 0; k < st.cnt; ++k) {
    if (s_lq[k * kThreads + tid] <= st.ub) {
      const int gi = s_lid[k * kThreads + tid];
      const double d = tri_exact64(A.rec64 + 9 * (size_t)gi, px, py, pz);
      const int id = A.tri_id[gi];
      st.n64++;
      if (d < st.best64 || (d == st.best64 && id < st.best_id)) { st.best64 = d; st.best_id = id; }
    }
  }
  st.cnt = 0;
  if (st.best_id >= 0) st.ub = fminf(st.ub, __double2float_ru(st.best64));
  st.sub = __fsqrt_ru(st.ub);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm) k_sdf_tiles(const SdfArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_rec = reinterpret_cast<float4*>(smem_raw);                 // kCap*4
  float4* s_sph = s_rec + kCap * 4;                                    // kCap
  int* s_gidx = reinterpret_cast<int*>(s_sph + kCap);                  // kCap
  int* s_lid = s_gidx + kCap;                                          // kListCap*kThreads
  float* s_lq = reinterpret_cast<float*>(s_lid + kListCap * kThreads); // kListCap*kThreads
  int* s_rstart = reinterpret_cast<int*>(s_lq + kListCap * kThreads);  // kMaxRanges
  int* s_rcnt = s_rstart + kMaxRanges;
  int* s_roff = s_rcnt + kMaxRanges;
  __shared__ int s_nr, s_total;
  __shared__ unsigned s_ub[2];
  __shared__ unsigned long long s_stats[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = A.N, nc = A.nc;
  const int tx = blockIdx.x % A.ntile, ty = (blockIdx.x / A.ntile) % A.ntile, tz = A.tz0 + blockIdx.x / (A.ntile * A.ntile);   // tz in units of kTileZ

  // this lane's voxel; the warp owns a 4x4x2 block of the tile
  const int bx = tx * kTile + (warp & 1) * 4, by = ty * kTile + ((warp >> 1) & 1) * 4, bz = tz * kTileZ + (warp >> 2) * 2;
  const int vx = bx + (lane & 3), vy = by + ((lane >> 2) & 3), vz = bz + (lane >> 4);
  const bool valid = vx < N && vy < N && vz < N && vz >= A.z0 && vz < A.z1;
  const double invN = 1.0 / (double)N;
  const double pxd = __ddiv_rn((double)vx, (double)N), pyd = __ddiv_rn((double)vy, (double)N),
               pzd = __ddiv_rn((double)vz, (double)N);   // mesh.cc:115-117
  const float px = (float)pxd, py = (float)pyd, pz = (float)pzd;
  const float wcx = (float)((bx + 1.5) * invN), wcy = (float)((by + 1.5) * invN), wcz = (float)((bz + 0.5) * invN);
  const float Rw = (float)(2.1795 * invN * 1.0001);   // half diagonal of the 3x3x1-interval sample box

  // tile sample box (clipped to the grid and the slab)
  const float tlo[3] = {(float)(tx * kTile * invN), (float)(ty * kTile * invN), (float)(max(tz * kTileZ, A.z0) * invN)};
  const float thi[3] = {(float)(min(tx * kTile + kTile - 1, N - 1) * invN), (float)(min(ty * kTile + kTile - 1, N - 1) * invN),
                        (float)(min(min(tz * kTileZ + kTileZ - 1, N - 1), A.z1 - 1) * invN)};

  LaneState st;
  st.best64 = DBL_MAX; st.best_id = -1; st.ub = __int_as_float(0x7f800000); st.sub = st.ub; st.cnt = 0; st.n64 = 0;
  unsigned n32 = 0, ncull = 0, nsph = 0;
  float thr_w = __int_as_float(0x7f800000);
  float ub_cta = __int_as_float(0x7f800000);
  const float max_ext = __uint_as_float(*A.max_ext);
  if (tid < 4) s_stats[tid] = 0ull;
  if (tid < 2) s_ub[tid] = 0u;
  int par = 0;

  constexpr int kCz = kTileZ / kCellVox;                // cells per tile in z
  const int cbx = 2 * tx, cby = 2 * ty, cbz = kCz * tz;   // the tile's 2 x 2 x kCz cell block
  for (int r = 0; r <= nc; ++r) {
    if (r >= 1) {
      const float lb = (float)(r - 1) * A.cs - max_ext;   // nothing binned in ring >= r is closer than this
      if (lb > 0.f && lb * lb * 0.9999f > ub_cta) break;
    }
    if (r >= 1) {   // ring r-1 already enclosed the whole cell grid
      const int q = r - 1;
      if (cbx - q <= 0 && cby - q <= 0 && cbz - q <= 0 && cbx + 1 + q >= nc - 1 && cby + 1 + q >= nc - 1 &&
          cbz + kCz - 1 + q >= nc - 1)
        break;
    }
    const int side = 2 + 2 * r, sidez = kCz + 2 * r;
    const int x0 = cbx - r, y0 = cby - r, z0c = cbz - r;
    const int nenum = side * side * sidez;
    for (int base = 0; base < nenum; base += kMaxRanges) {
      __syncthreads();
      if (tid == 0) { s_nr = 0; s_total = 0; }
      __syncthreads();
      const int lim = min(nenum, base + kMaxRanges);
      for (int i = base + tid; i < lim; i += kThreads) {
        const int ix = i % side, iy = (i / side) % side, iz = i / (side * side);
        if (r > 0 && ix > 0 && ix < side - 1 && iy > 0 && iy < side - 1 && iz > 0 && iz < sidez - 1) continue;
        const int cx = x0 + ix, cy = y0 + iy, cz = z0c + iz;
        if ((unsigned)cx >= (unsigned)nc || (unsigned)cy >= (unsigned)nc || (unsigned)cz >= (unsigned)nc) continue;
        const int c = (cz * nc + cy) * nc + cx;
        const int cs0 = A.cell_start[c], cnt = A.cell_start[c + 1] - cs0;
        if (cnt == 0) continue;
        float d2 = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float lo = o2f(A.cell_bb[6 * (size_t)c + j]), hi = o2f(A.cell_bb[6 * (size_t)c + 3 + j]);
          const float gap = fmaxf(0.f, fmaxf(lo - thi[j], tlo[j] - hi));
          d2 = fmaf(gap, gap, d2);
        }
        if (d2 * 0.9999f <= ub_cta) {
          const int slot = atomicAdd(&s_nr, 1);
          s_rstart[slot] = cs0; s_rcnt[slot] = cnt; s_roff[slot] = atomicAdd(&s_total, cnt);
        }
      }
      __syncthreads();
      const int nr = s_nr, total = s_total;
      for (int cb = 0; cb < total; cb += kCap) {
        // ---- stage up to kCap candidate records in shared memory -------------------
        for (int ri = warp; ri < nr; ri += kWarps) {
          const int off = s_roff[ri], cnt = s_rcnt[ri], start = s_rstart[ri];
          const int lo = max(off, cb), hi = min(off + cnt, cb + kCap);
          for (int q4 = lane; q4 < (hi - lo) * 4; q4 += 32) {
            const int rec = lo + (q4 >> 2), part = q4 & 3;
            const int g = start + (rec - off);
            s_rec[(rec - cb) * 4 + part] = __ldg(&A.rec32[4 * (size_t)g + part]);
            if (part == 0) s_gidx[rec - cb] = g;
            if (part == 1) s_sph[rec - cb] = __ldg(&A.sph[g]);
          }
        }
        __syncthreads();
        const int nrec = min(kCap, total - cb);
        // ---- per warp: cull against the warp bound, dense test on the survivors ------
        for (int b = 0; b < nrec; b += 32) {
          const int j = b + lane;
          const bool has = j < nrec;
          const int jr = has ? j : 0;
          float ec;
          const float qc = tri_q(s_rec[jr * 4], s_rec[jr * 4 + 1], s_rec[jr * 4 + 2], s_rec[jr * 4 + 3], wcx, wcy, wcz, ec);
          unsigned m = __ballot_sync(0xffffffffu, has && (qc - ec <= thr_w));
          ncull += has ? 1u : 0u;
          nsph += valid ? (unsigned)__popc(m) : 0u;
          while (m) {
            const int jj = b + __ffs(m) - 1;
            m &= m - 1;
            // per-voxel pre-test against the triangle's bounding sphere: |p - c| - rho is a lower bound of
            // the distance; if it exceeds every lane's upper bound the exact test is skipped for the warp
            const float4 sp = s_sph[jj];
            const float dx = px - sp.x, dy = py - sp.y, dz = pz - sp.z;
            const float dc2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float reach = st.sub + sp.w;
            if (!__any_sync(0xffffffffu, valid && !(dc2 > reach * reach * 1.000002f))) continue;
            n32 += valid ? 1u : 0u;
            float e;
            const float q = tri_q(s_rec[jj * 4], s_rec[jj * 4 + 1], s_rec[jj * 4 + 2], s_rec[jj * 4 + 3], px, py, pz, e);
            const float qlo = q - e;
            if (valid && qlo <= st.ub) {
              if (st.cnt == kListCap) flush_queue(st, s_lid, s_lq, tid, A, pxd, pyd, pzd);
              s_lid[st.cnt * kThreads + tid] = s_gidx[jj];
              s_lq[st.cnt * kThreads + tid] = qlo;
              st.cnt++;
            }
            if (q + e < st.ub) { st.ub = q + e; st.sub = __fsqrt_ru(st.ub); }
          }
          const float um = warp_max(valid ? st.ub : 0.f);
          const float su = sqrtf(um) + Rw;
          thr_w = su * su * 1.00001f;
        }
        const float um = warp_max(valid ? st.ub : 0.f);
        if (lane == 0) atomicMax(&s_ub[par], __float_as_uint(um));
        __syncthreads();
        ub_cta = __uint_as_float(s_ub[par]);
        par ^= 1;
        if (tid == 0) s_ub[par] = 0u;   // next chunk's slot; not touched again before two more barriers
      }
    }
  }

  // ---- exact FP64 evaluation of everything still queued, then store ------------------
  if (valid) {
    flush_queue(st, s_lid, s_lq, tid, A, pxd, pyd, pzd);
    const size_t o = ((size_t)vz * N + vy) * N + vx;
    const double d = st.best_id >= 0 ? __dsqrt_rn(st.best64) : 1e30;   // mesh.cc:146
    A.grid64[o] = d;
    A.grid32[o] = (float)d;
    A.nearest[o] = st.best_id;
  }
  unsigned long long a32 = n32, a64 = st.n64, ac = ncull, as = nsph;
  for (int o = 16; o > 0; o >>= 1) {
    a32 += __shfl_xor_sync(0xffffffffu, a32, o);
    a64 += __shfl_xor_sync(0xffffffffu, a64, o);
    ac += __shfl_xor_sync(0xffffffffu, ac, o);
    as += __shfl_xor_sync(0xffffffffu, as, o);
  }
  if (lane == 0) { atomicAdd(&s_stats[0], a32); atomicAdd(&s_stats[1], a64); atomicAdd(&s_stats[2], ac); atomicAdd(&s_stats[3], as); }
  __syncthreads();
  if (tid < 3) atomicAdd(&A.stats[tid], s_stats[tid]);
  if (tid == 3) atomicAdd(&A.stats[4], s_stats[3]);
}

constexpr size_t kSdfSmem = (size_t)kCap * 64 + (size_t)kCap * 16 + (size_t)kCap * 4 + (size_t)kListCap * kThreads * 8 + (size_t)kMaxRanges * 12;

int run_build(Template& T, cudaStream_t s) {
  const int N = T.N, nF = T.nF, nV = T.nV;
  const int ntile = div_up(N, kTile), nc = 2 * ntile;
  const size_t ncell = (size_t)nc * nc * nc;
  const size_t nvox = (size_t)N * N * N;

  int *cell_count = nullptr, *cell_start = nullptr, *tri_cell = nullptr, *tri_id = nullptr;
  unsigned *cell_bb = nullptr, *max_ext = nullptr;
  float4 *rec32 = nullptr, *sph = nullptr;
  double* rec64 = nullptr;
  // one scratch allocation, stream ordered
  const size_t b_count = 2 * ncell * sizeof(int);          // count + fill
  const size_t b_start = (ncell + 1) * sizeof(int);
  const size_t b_bb = 6 * ncell * sizeof(unsigned);
  const size_t b_tri = 2 * (size_t)nF * sizeof(int);       // tri_cell + tri_id
  const size_t b_r32 = (size_t)nF * 64, b_r64 = (size_t)nF * 72, b_sph = (size_t)nF * 16;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t total = al(b_count) + al(b_start) + al(b_bb) + al(b_tri) + al(b_r32) + al(b_r64) + al(b_sph) + 256;
  unsigned char* scratch = nullptr;
  MO_CUDA(cudaMallocAsync(&scratch, total, s));
  unsigned char* p = scratch;
  cell_count = (int*)p; p += al(b_count);
  cell_start = (int*)p; p += al(b_start);
  cell_bb = (unsigned*)p; p += al(b_bb);
  tri_cell = (int*)p; tri_id = tri_cell + nF; p += al(b_tri);
  rec32 = (float4*)p; p += al(b_r32);
  rec64 = (double*)p; p += al(b_r64);
  sph = (float4*)p; p += al(b_sph);
  max_ext = (unsigned*)p;
  int* cell_fill = cell_count + ncell;

  MO_CUDA(cudaMemsetAsync(cell_count, 0, b_count, s));
  k_init_cells<<<div_up((long long)ncell, 256), 256, 0, s>>>(cell_bb, (int)ncell);
  MO_LAUNCH_CHECK();
  MO_CUDA(cudaMemsetAsync(max_ext, 0, sizeof(unsigned), s));
  MO_CUDA(cudaMemsetAsync(T.d_stats, 0, 8 * sizeof(unsigned long long), s));

  k_tri_count<<<div_up(nF, 256), 256, 0, s>>>(T.d_Vn, T.d_F, nF, nV, N, nc, cell_count, cell_bb, tri_cell, max_ext, T.d_stats);
  MO_LAUNCH_CHECK();
  k_scan<<<1, 1024, 0, s>>>(cell_count, cell_start, (int)ncell);
  MO_LAUNCH_CHECK();
  k_tri_fill<<<div_up(nF, 256), 256, 0, s>>>(T.d_Vn, T.d_F, nF, tri_cell, cell_start, cell_fill, rec32, sph, rec64, tri_id);
  MO_LAUNCH_CHECK();

  if (T.z0 > 0 || T.z1 < N) {
    k_fill_grid<<<div_up((long long)nvox, 256), 256, 0, s>>>(T.d_grid64, T.d_grid32, T.d_nearest, nvox);
    MO_LAUNCH_CHECK();
  }

  SdfArgs A;
  A.N = N; A.nc = nc; A.ntile = ntile; A.z0 = T.z0; A.z1 = T.z1;
  A.ntz = div_up(N, kTileZ);
  A.tz0 = T.z0 / kTileZ;
  const int tz1 = (T.z1 - 1) / kTileZ;
  A.cs = (float)((double)kCellVox / N);
  A.max_ext = max_ext; A.cell_start = cell_start; A.cell_bb = cell_bb;
  A.rec32 = rec32; A.sph = sph; A.rec64 = rec64; A.tri_id = tri_id;
  A.grid64 = T.d_grid64; A.grid32 = T.d_grid32; A.nearest = T.d_nearest; A.stats = T.d_stats;
  static bool attr_set[64] = {};
  if (!attr_set[T.device & 63]) {
    MO_CUDA(cudaFuncSetAttribute(k_sdf_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSdfSmem));
    attr_set[T.device & 63] = true;
  }
  const int ntiles = ntile * ntile * (tz1 - A.tz0 + 1);
  k_sdf_tiles<<<ntiles, kThreads, kSdfSmem, s>>>(A);
  MO_LAUNCH_CHECK();
  MO_CUDA(cudaFreeAsync(scratch, s));
  return MO_OK;
}

}  // namespace

int build_field_from_f32(Template& T, const float* d_V, cudaStream_t s) {
  unsigned* bb = nullptr;
  MO_CUDA(cudaMallocAsync(&bb, 6 * sizeof(unsigned), s));
  MO_CUDA(cudaMemsetAsync(bb, 0xff, 3 * sizeof(unsigned), s));
  MO_CUDA(cudaMemsetAsync(bb + 3, 0, 3 * sizeof(unsigned), s));
  const int blocks = std::min(div_up(T.nV, 256), 296);
  k_bbox<<<blocks, 256, 0, s>>>(d_V, T.nV, bb);
  MO_LAUNCH_CHECK();
  k_xform<<<1, 1, 0, s>>>(bb, T.d_xf);
  MO_LAUNCH_CHECK();
  k_vnorm<<<div_up(3LL * T.nV, 256), 256, 0, s>>>(d_V, 3 * T.nV, T.d_xf, T.d_Vn);
  MO_LAUNCH_CHECK();
  MO_CUDA(cudaFreeAsync(bb, s));
  return run_build(T, s);
}

int build_field_from_normalized(Template& T, cudaStream_t s) { return run_build(T, s); }

}  // namespace mo
