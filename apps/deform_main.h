// Shared body of the re-hosted rigid_deform / rigid_rot_deform drivers (reference
// src/app/rigid_deform.cc:12-63, src/app/rigid_rot_deform.cc:12-63): same command line, same console
// output, the distance field (Mesh::ConstructDistanceField) and Deformer::Deform / DeformWithRot
// (src/lib/deformer.cc:18-92 / :94-171) on the GPU through the C-ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "../include/meshode_b200.h"
#include "mesh_host.h"

namespace mo_app {

#define APP_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e__)); return 2; } \
  } while (0)
#define APP_MO(call)                                                                  \
  do {                                                                                \
    if ((call) != MO_OK) { fprintf(stderr, "%s: %s\n", #call, mo_last_error()); return 2; } \
  } while (0)

inline int deform_main(int argc, char** argv, const char* name, bool with_rot) {
  int GRID_RESOLUTION = 64;
  int MESH_RESOLUTION = 5000;
  if (argc < 5) {
    printf("./%s source.obj reference.obj output.obj [GRID_RESOLUTION=64] [MESH_RESOLUTION=5000] [lambda=1] [symmetry=0]\n", name);
    return 0;
  }
  Mesh src, ref;
  if (!src.ReadOBJ(argv[1]) || !ref.ReadOBJ(argv[2])) { fprintf(stderr, "cannot read the input meshes\n"); return 1; }
  int symmetry = 0;
  if (argc > 7) sscanf(argv[7], "%d", &symmetry);
  if (symmetry) ref.ReflectionSymmetrize();
  if (argc > 4) sscanf(argv[4], "%d", &GRID_RESOLUTION);
  if (argc > 5) sscanf(argv[5], "%d", &MESH_RESOLUTION);   // parsed and unused, as in the reference
  double lambda = 1;
  if (argc > 6) sscanf(argv[6], "%lf", &lambda);
  std::cout << "Source:\t\t" << "Num vertices: " << src.nV() << "\tNum faces: " << src.nF() << std::endl;
  std::cout << "Reference:\t" << "Num vertices: " << ref.nV() << "\tNum faces: " << ref.nF() << std::endl << std::endl;

  ref.Normalize();
  src.ApplyTransform(ref);

  // UniformGrid grid(GRID_RESOLUTION); ref.ConstructDistanceField(grid);
  double* d_ref = nullptr; int* d_refF = nullptr;
  APP_CUDA(cudaMalloc(&d_ref, sizeof(double) * ref.V.size()));
  APP_CUDA(cudaMalloc(&d_refF, sizeof(int) * ref.F.size()));
  APP_CUDA(cudaMemcpy(d_ref, ref.V.data(), sizeof(double) * ref.V.size(), cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemcpy(d_refF, ref.F.data(), sizeof(int) * ref.F.size(), cudaMemcpyHostToDevice));
  int pid = -1;
  APP_MO(mo_template_create_normalized(d_ref, ref.nV(), d_refF, ref.nF(), GRID_RESOLUTION, ref.scale, ref.pos, nullptr, &pid));

  // residual blocks of Deformer::Deform / DeformWithRot: one EdgeLoss per directed face edge,
  // v = V[F[i][j]] - V[F[i][(j+1)%3]] (deformer.cc:42-52 / :119-131)
  const int nV = src.nV(), nE = 3 * src.nF();
  std::vector<int> I(2 * (size_t)nE);
  std::vector<double> rest(3 * (size_t)nE);
  for (int i = 0; i < src.nF(); ++i) {
    for (int j = 0; j < 3; ++j) {
      const int a = src.F[3 * i + j], b = src.F[3 * i + (j + 1) % 3];
      if (a < 0 || a >= nV || b < 0 || b >= nV) { fprintf(stderr, "face %d references a missing vertex\n", i); return 1; }
      const size_t e = 3 * (size_t)i + j;
      I[2 * e] = a; I[2 * e + 1] = b;
      for (int k = 0; k < 3; ++k) rest[3 * e + k] = src.V[3 * (size_t)a + k] - src.V[3 * (size_t)b + k];
    }
  }
  double *d_V = nullptr, *d_R = nullptr, *d_rest = nullptr; int* d_I = nullptr;
  APP_CUDA(cudaMalloc(&d_V, sizeof(double) * 3 * (size_t)nV + 8));
  APP_CUDA(cudaMalloc(&d_R, sizeof(double) * 3 * (size_t)nV + 8));
  APP_CUDA(cudaMalloc(&d_rest, sizeof(double) * rest.size() + 8));
  APP_CUDA(cudaMalloc(&d_I, sizeof(int) * I.size() + 8));
  APP_CUDA(cudaMemcpy(d_V, src.V.data(), sizeof(double) * 3 * (size_t)nV, cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemset(d_R, 0, sizeof(double) * 3 * (size_t)nV));   // std::vector<double> rots(V.size() * 3, 0)
  APP_CUDA(cudaMemcpy(d_rest, rest.data(), sizeof(double) * rest.size(), cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemcpy(d_I, I.data(), sizeof(int) * I.size(), cudaMemcpyHostToDevice));

  double summary[10];
  const int max_cg = getenv("MESHODE_MAX_CG") ? atoi(getenv("MESHODE_MAX_CG")) : 0;
  APP_MO(mo_ceres_solve(pid, with_rot ? MO_CERES_ROT_EDGE : MO_CERES_EDGE, d_V, with_rot ? d_R : nullptr, nV, d_I, d_rest, nE,
                        lambda, /*max_num_iterations*/ 100, max_cg, 0.0, /*minimizer_progress_to_stdout*/ 1, summary, nullptr));
  std::cout << "Vertices cost: " << summary[2] << std::endl;
  std::cout << "Rigidity cost: " << summary[3] << std::endl;
  std::cout << "Final cost: " << summary[2] + summary[3] << std::endl;
  APP_CUDA(cudaMemcpy(src.V.data(), d_V, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToHost));
  std::cout << "Deformed" << std::endl;
  if (!src.WriteOBJ(argv[3])) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
  mo_template_destroy(pid);
  cudaFree(d_ref); cudaFree(d_refF); cudaFree(d_V); cudaFree(d_R); cudaFree(d_rest); cudaFree(d_I);
  return 0;
}

}  // namespace mo_app
