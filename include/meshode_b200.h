/* meshode_b200.h -- C-ABI of libmeshode_b200.so (sm_100a CUDA implementation of
 * MeshODE's data-parallel hot path).
 *
 * Every entry point takes plain DEVICE pointers and sizes plus the CUDA stream
 * to enqueue on (a cudaStream_t passed as void*; NULL = legacy default stream).
 * Nothing here synchronises the host unless its comment says so.  All functions
 * return 0 (MO_OK) or a negative MO_ERR_* code; mo_last_error() gives the text
 * of the calling thread's last failure.  There is no CPU fallback: without a
 * CUDA device every compute entry returns MO_ERR_CUDA.
 *
 * Each declaration cites the reference interface (relative to the MeshODE
 * source tree) that it replaces.  Tensors of the reference API map to
 * (pointer, row count) pairs: V = float32 [n,3] contiguous, F = int32 [m,3],
 * E = int32 [e,2].
 */
#ifndef MESHODE_B200_H_
#define MESHODE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mo_stream_t; /* cudaStream_t */

enum {
  MO_OK = 0,
  MO_ERR_BAD_HANDLE = -1, /* param_id not created / already destroyed */
  MO_ERR_BAD_ARG = -2,    /* null pointer, negative size, size mismatch with stored edges */
  MO_ERR_CUDA = -3,       /* a CUDA runtime call failed (text in mo_last_error) */
  MO_ERR_STATE = -4       /* edges of that kind were never stored for this template */
};

/* which edge set a template holds; one set per template, each Store* overwrites it
 * (src/interface/rigid_layer.cc:29, graph_layer.cc:29, cad_layer.cc:33) */
enum { MO_EDGES_NONE = -1, MO_EDGES_RIGID = 0, MO_EDGES_GRAPH = 1, MO_EDGES_CAD = 2 };

int mo_version(void);
const char* mo_last_error(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches). */
unsigned long long mo_launch_count(void);
/* FFMA-chain microbenchmark used as the FP32 roofline denominator: every thread runs `iters`
 * dependent-free FMA groups (8 independent chains); flops = blocks*threads*iters*8*2.
 * d_sink (>= blocks*threads floats) keeps the result alive. */
int mo_microbench_fp32(int blocks, int threads, int iters, float* d_sink, mo_stream_t stream);
/* The same on packed operands (fma.rn.f32x2, two FMAs per instruction): flops = blocks*threads*iters*8*4.
 * sm_100 issues it at half rate, so it lands a few per cent above the scalar chain; bench.py takes the larger. */
int mo_microbench_fp32x2(int blocks, int threads, int iters, float* d_sink, mo_stream_t stream);
/* number of visible CUDA devices (0 when there is none). */
int mo_device_count(void);

/* ---- template = DeformParams (src/interface/deform_params.h:7-25) ------------ */

/* InitializeDeformTemplate(tensorV, tensorF, symmetry, grid_resolution) -> param_id
 * (src/interface/deform_params.cc:16-40): CopyTensorToMesh(normalize=1)
 * (src/interface/mesh_tensor.cc:47-85) + Mesh::Normalize (src/lib/mesh.cc:66-85) +
 * Mesh::ConstructDistanceField (src/lib/mesh.cc:106-152).  `symmetry` is accepted and
 * ignored, as in the reference (the reflection runs on a still-empty mesh, :28-33).
 * The distance field, nearest-triangle index, scale and translation stay on the
 * device; no host synchronisation. */
int mo_template_create(const float* d_V, int nV, const int* d_F, int nF, int symmetry, int grid_res,
                       mo_stream_t stream, int* out_param_id);

/* Same, but only voxel slices z in [z0, z1) are computed (z-slab sharding of one large
 * grid across GPUs); the other slices hold 1e30 (src/lib/uniformgrid.cc:9-17 initial
 * value) until the caller fills them, e.g. by an all-gather on the pointers returned
 * by mo_template_grid(). */
int mo_template_create_slab(const float* d_V, int nV, const int* d_F, int nF, int grid_res, int z0, int z1,
                            mo_stream_t stream, int* out_param_id);

/* Same, cyclic: only the z-tile layers first_layer, first_layer + layer_stride, ... are computed, a layer being
 * MO_LAYER_SLICES = 4 consecutive voxel slices (layer l = slices [4l, 4l+4)).  With layer_stride = number of
 * GPUs and first_layer = rank every GPU's share spans the whole z range, so the shares cost the same wherever
 * the surface lies (contiguous slabs through the middle of a shape cost more than polar ones).  The layers of
 * group j (slices [4*stride*j, 4*stride*(j+1))) are contiguous in the fields, rank r's piece at offset 4r
 * slices: one in-place all-gather per group assembles the field. */
enum { MO_LAYER_SLICES = 4 };
int mo_template_create_layers(const float* d_V, int nV, const int* d_F, int nF, int grid_res, int first_layer,
                              int layer_stride, mo_stream_t stream, int* out_param_id);

/* Mesh::ConstructDistanceField on an ALREADY normalised FP64 mesh (the C++ drivers'
 * path: ref.Normalize(); ref.ConstructDistanceField(grid) -- src/app/rigid_deform.cc:49-52).
 * scale/trans are recorded as given (Mesh::GetScale/GetTranslation). */
int mo_template_create_normalized(const double* d_Vn, int nV, const int* d_F, int nF, int grid_res, double scale,
                                  const double* h_trans3, mo_stream_t stream, int* out_param_id);

/* g_params entries are never freed in the reference (deform_params.cc:7-14); these are additive.
 * mo_template_destroy waits for the template's device to go idle (work that still reads the template may sit on
 * any stream) and frees; mo_template_destroy_async returns the buffers to the pool in the order of `stream` --
 * the caller guarantees that every use of the template was enqueued on, or is ordered before, that stream.
 * Entry points running concurrently with a destroy keep the template alive until they return. */
int mo_template_destroy(int param_id);
int mo_template_destroy_async(int param_id, mo_stream_t stream);

/* Host copies of the template's metadata (Mesh::GetScale / GetTranslation,
 * UniformGrid::Dimension).  Synchronises `stream`.  Any out pointer may be NULL. */
int mo_template_info(int param_id, mo_stream_t stream, int* grid_res, int* nV, int* nF, double* scale,
                     double* trans3);

/* Device pointers of the stored fields, all N^3 in [z][y][x] order
 * (UniformGrid::GetDistance(i=z, j=y, k=x), src/lib/uniformgrid.h:29-31):
 *   grid_f64  : sqrt of the FP64 squared distance, what UniformGrid stores;
 *   grid_f32  : (float)grid_f64, the value DistanceFloat<> fetches (uniformgrid.cc:122);
 *   nearest   : index into F of the nearest triangle (igl's `I`, src/lib/mesh.cc:138-140),
 *               lowest index on exact FP64 ties.
 * Any out pointer may be NULL. */
int mo_template_grid(int param_id, const double** d_grid_f64, const float** d_grid_f32, const int** d_nearest);

/* Copies voxel slices z in [z0, z1) between the template's fields and caller-owned full-size
 * (N^3) device arrays laid out the same way: direction 0 = template -> caller (read back,
 * e.g. into torch tensors), 1 = caller -> template (e.g. after an all-gather of z-slabs).
 * Any of the three array pointers may be NULL. Asynchronous on `stream`. */
int mo_template_copy_grid(int param_id, int direction, int z0, int z1, double* d_grid_f64, float* d_grid_f32,
                          int* d_nearest, mo_stream_t stream);

/* Device pointer of the normalised FP64 target vertices [nV,3] (Mesh::GetV after Normalize). */
int mo_template_vertices(int param_id, const double** d_Vn);

/* Distance-field builds started after mo_build_stats_enable(1) run the instrumented instantiation of the search
 * kernel, which counts its tests for mo_template_build_stats; the default (0) runs the same search without the
 * counters (they cost registers in the hot loops).  The counts are deterministic: an instrumented build of the same
 * input reports what the plain build executed.  Returns the previous setting.  Process-wide. */
int mo_build_stats_enable(int on);

/* Build statistics of the last grid build of this template (for roofline accounting; all zero unless the build ran
 * with mo_build_stats_enable(1)):
 * point-triangle tests executed in FP32 (74 FLOP each), exact FP64 re-evaluations, bounding-
 * cylinder tests of triangle clusters against a tile, a block or a voxel, and bounding-disc pre-tests of
 * single triangles against a voxel (both the same ~38 FLOP test, sdf_build.cu cyl_skip).
 * Synchronises `stream`.  Any out pointer may be NULL. */
int mo_template_build_stats(int param_id, mo_stream_t stream, unsigned long long* fp32_tests,
                            unsigned long long* fp64_tests, unsigned long long* cull_tests,
                            unsigned long long* disc_tests);

/* ---- NormalizeByTemplate / DenormalizeByTemplate (src/interface/normalize.cc:5-45) ---- */
/* in place on float32 [n,3]; inverse = 0: (v - trans)/scale, inverse = 1: v*scale + trans,
 * arithmetic in FP64 rounded once to float32 as in the reference. */
int mo_normalize_by_template(float* d_V, int n, int param_id, int inverse, mo_stream_t stream);

/* ---- DistanceFieldLoss (src/interface/distance_layer.cc) ---------------------------- */
/* DistanceFieldLoss_forward (:8-37): out[i] = DistanceFloat<float>(V[i])^2, float32 [n]. */
int mo_distance_forward(const float* d_V, int n, int param_id, float* d_out, mo_stream_t stream);
/* DistanceFieldLoss_backward (:39-81): out[i,:] = 0.5 * d(dist^2)/dV[i] via Jet<float,3>
 * arithmetic, float32 [n,3]. */
int mo_distance_backward(const float* d_V, int n, int param_id, float* d_grad, mo_stream_t stream);
/* both in one pass over V (additive; 28 B of HBM traffic per vertex instead of 16 + 24). */
int mo_distance_forward_backward(const float* d_V, int n, int param_id, float* d_out, float* d_grad,
                                 mo_stream_t stream);
/* UniformGrid::distance<double> / distance<Jet<double,3>> (src/lib/uniformgrid.cc:18-83), the
 * sampler behind DistanceLoss (src/lib/distanceloss.h:6-25): val [n] and, if d_grad != NULL,
 * the three partials [n,3]. */
int mo_distance_f64(const double* d_P, int n, int param_id, double* d_val, double* d_grad, mo_stream_t stream);

/* ---- edge rigidity (src/interface/{rigid,graph,cad}_layer.cc) ----------------------- */
/* Store{Rigidity,Graph,Cad}Information: rest vectors (and CAD lambda) from the current V.
 * kind RIGID uses F (3*nF directed face edges), GRAPH uses E, CAD uses E rows then face edges.
 * Unused index arrays may be NULL with count 0.  Also captures the connectivity for the
 * deterministic backward. */
int mo_edges_store(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                   int nE, mo_stream_t stream);
/* {Rigid,Graph,Cad}EdgeLoss_forward: out float32 [nEdges,3], nEdges = 3nF | nE | nE+3nF.
 * Reads the index arrays passed here, like the reference. */
int mo_edges_forward(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                     int nE, float* d_out, mo_stream_t stream);
/* {Rigid,Graph,Cad}EdgeLoss_backward: out float32 [nV,3].  Vertex-gather over the CSR built
 * by mo_edges_store, accumulating each vertex's incident edges in edge order, so the result
 * is bit-identical to the reference's serial scatter loop.  The counts must equal the
 * stored ones (MO_ERR_BAD_ARG otherwise).  NOTE: the connectivity is the one captured by
 * mo_edges_store -- d_F / d_E are only checked for their counts here (the reference walks the
 * caller's arrays, rigid_layer.cc:113-130; passing a different index array of the same length
 * than the one stored mixes foreign indices with the stored rest vectors there, a meaningless
 * result either way).  mo_edges_backward_atomic honours the arrays passed to it. */
int mo_edges_backward(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                      int nE, float* d_grad, mo_stream_t stream);
/* Same result up to float32 summation order: edge-parallel scatter with warp-aggregated
 * red.global.add.f32 on the index arrays passed here.  d_grad is zeroed first. */
int mo_edges_backward_atomic(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF,
                             const int* d_E, int nE, float* d_grad, mo_stream_t stream);

/* ---- fused per-iteration loss (src/python/layers/{rigid,graph,graph2}_loss_layer.py) --------------------- */
/* L = 0.5*sum(dist_fwd) + w_edge*0.5*sum(edge_fwd);  grad = mask*dist_bwd + w_edge*edge_bwd
 * (rigid_loss_layer.py:9-27 with w_edge = 1; graph_loss_layer.py:11-43 with w_edge =
 * rigidity^2 and mask_threshold = 0.5*0.03^2 on 0.5*dist_fwd; mask_threshold <= 0 disables
 * the mask).  dist_param_id selects the distance field, edge_param_id the stored edges
 * (graph_loss2_layer.py:18-19 uses two different templates).  d_loss receives one double
 * (sum accumulated in FP64); d_grad float32 [nV,3]. Either may be NULL. */
int mo_loss_forward_backward(int dist_param_id, int edge_param_id, const float* d_V, int nV, float w_edge,
                             float mask_threshold, double* d_loss, float* d_grad, mo_stream_t stream);

enum {
  MO_DEFORM_EXACT = 1,         /* reference order of the edge sum: bit-identical to the CPU loop */
  MO_DEFORM_CTA_ONLY = 2,      /* exact loop: never use the cluster-split kernel (A/B timing, tests) */
  MO_DEFORM_CLUSTER_ONLY = 4   /* exact loop: every pair on the cluster-split kernel (A/B timing, tests) */
};

/* ---- whole optimisation loops (src/python/rigid_deform.py:32-41) ----------------------- */
/* For each of B independent pairs: `iters` iterations of
 *     grad = DistanceFieldLoss_backward(V, dist_pid) + {Rigid,Graph}EdgeLoss_backward(V, edge_pid)
 *     torch.optim.Adam step (lr, betas, eps; float32 state, bias correction as in
 *     torch/optim/adam.py::_single_tensor_adam)
 * in ONE persistent kernel, one CTA per pair, V / rest positions / gradient resident in shared
 * memory.  h_dist_pids / h_edge_pids are HOST arrays of B param_ids (they may be equal, as in
 * rigid_loss_layer.py, or differ, as in graph_loss2_layer.py:18-19); h_dV is a HOST array of B
 * device pointers to normalised float32 [nV_i,3] vertices, updated in place.  Edges (RIGID or
 * GRAPH) must have been stored with mo_edges_store for nV_i <= 6144 vertices.
 * flags = MO_DEFORM_EXACT: bit-identical to running the per-call entry points and Adam in float32 on
 * the CPU (edge terms accumulated in the reference's order); this is what the parity gates and
 * bench.py use.  flags = 0 (fast, opt-in): the edge term is summed over distinct neighbours on
 * displacements U = V - V0 -- mathematically the same sum, one gather per neighbour, ~1e-10 per
 * term away from the reference's float32 order; Adam amplifies that to ~2e-4 Chamfer after 10 000
 * iterations, exactly what a 1-ulp change of the input does to the exact loop (tools/chaos_probe.py).
 * Scheduling of the exact loop: full waves of pairs run one CTA per pair; the B mod (SM count) pairs of a
 * partial wave (all of them when B is smaller than the SM count) run on thread-block clusters, ceil(nV/1024)
 * SMs per pair with the positions exchanged through distributed shared memory, when that finishes sooner.
 * Same arithmetic, same bits; MO_DEFORM_CTA_ONLY / MO_DEFORM_CLUSTER_ONLY pin one schedule. */
int mo_deform_batch_adam(const int* h_dist_pids, const int* h_edge_pids, float* const* h_dV, int B, int iters,
                         double lr, double beta1, double beta2, double eps, int flags, mo_stream_t stream);
/* Same loop for one pair of any size (state in HBM/L2, two launches per iteration), with the
 * graph layer's options: edge weight (rigidity^2) and distance-gradient mask threshold
 * (graph_loss_layer.py:18,40-42; mask_threshold <= 0 disables the mask). */
int mo_deform_adam_large(int dist_pid, int edge_pid, float* d_V, int nV, float w_edge, float mask_threshold,
                         int iters, double lr, double beta1, double beta2, double eps, mo_stream_t stream);

/* ---- nearest vertex (src/python/layers/reverse_loss_layer.py:15-19) --------------------- */
/* cKDTree(P).query(Q, k=1): for every query point Q[i] (float32 [nQ,3]) the index of the nearest
 * point of P (float32 [nP,3]) into d_idx (int32 [nQ]) and, if d_dist2 != NULL, the squared distance
 * (FP64, like cKDTree's double arithmetic).  Exact search; the lowest index wins exact ties. */
int mo_nearest_vertex(const float* d_Q, int nQ, const float* d_P, int nP, int* d_idx, double* d_dist2,
                      mo_stream_t stream);

/* ---- the Ceres loss terms of the C++ drivers (src/lib/deformer.cc), FP64 ------------------ */
enum { MO_CERES_EDGE = 0, MO_CERES_ADAPTIVE_EDGE = 1, MO_CERES_ROT_EDGE = 2 };
/* Residual blocks EdgeLoss (src/lib/edgeloss.h:8-33), AdaptiveEdgeLoss (:35-62) or EdgeLossWithRot
 * (:64-98) for nE edges: block e couples p1 = V[I[e,0]], p2 = V[I[e,1]] (and rot1 = R[I[e,0]], rot2 =
 * R[I[e,1]] for kind ROT; d_R may be NULL otherwise) with rest vector d_rest[e] (Deformer passes
 * v = V0[a] - V0[b], deformer.cc:44-49).  d_res receives [nE,3] ([nE,6] for ROT); d_jac, if not
 * NULL, the Jacobian ceres::AutoDiffCostFunction would return: [nE,3,6] over (p1,p2), or [nE,6,12]
 * over (p1,p2,rot1,rot2) for ROT, row-major. */
int mo_ceres_edges(int kind, const double* d_V, const double* d_R, int nV, const int* d_I, const double* d_rest,
                   int nE, double lambda, double* d_res, double* d_jac, mo_stream_t stream);
/* What ceres::Problem::Evaluate returns for the problems of Deformer::Deform / DeformWithRot /
 * DeformGraph (src/lib/deformer.cc:18-92, :94-171, :370-442): nV DistanceLoss blocks on the field
 * of dist_param_id (src/lib/distanceloss.h:6-25; dist_param_id < 0 leaves them out) plus nE edge
 * blocks of `kind`.  d_cost2[0] = 0.5*sum of squared distance residuals, d_cost2[1] = same for the
 * edge residuals; d_gV [nV,3] and d_gR [nV,3] (ROT only) receive the gradient J^T r.  Any output
 * may be NULL. */
int mo_ceres_problem(int dist_param_id, int kind, const double* d_V, const double* d_R, int nV, const int* d_I,
                     const double* d_rest, int nE, double lambda, double* d_cost2, double* d_gV, double* d_gR,
                     mo_stream_t stream);

/* ceres::Solve(options, &problem, &summary) for those problems with the options of the C++ drivers
 * (src/lib/deformer.cc:55-74, :135-153: max_num_iterations = 100, everything else Ceres' defaults:
 * trust-region Levenberg-Marquardt, initial radius 1e4, function / gradient / parameter tolerances
 * 1e-6 / 1e-10 / 1e-8, Jacobi scaling).  d_V [nV,3] (and d_R [nV,3] for kind ROT) are optimised in
 * place.  The LM step solves (J^T J + D^T D) delta = -g matrix-free with Jacobi-preconditioned CG
 * (relative residual cg_tolerance, default 1e-10; at most max_cg_iterations, default 4000) where Ceres
 * uses a sparse Cholesky factorisation.  h_summary (HOST, 10 doubles, may be NULL): initial cost,
 * final cost, final distance cost ("Vertices cost"), final edge cost ("Rigidity cost"), LM iterations,
 * accepted steps, total CG iterations, termination (0 function tolerance, 1 gradient tolerance, 2
 * parameter tolerance, 3 iteration limit, 4 five invalid steps in a row, 5 radius underflow), final
 * trust-region radius, final max |gradient|.  verbose != 0 prints Ceres-style progress lines.
 * Synchronises the stream (the LM loop reads its scalars back). */
int mo_ceres_solve(int dist_param_id, int kind, double* d_V, double* d_R, int nV, const int* d_I, const double* d_rest,
                   int nE, double lambda, int max_iterations, int max_cg_iterations, double cg_tolerance, int verbose,
                   double* h_summary, mo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MESHODE_B200_H_ */
