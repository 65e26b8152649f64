// cad_deform cad.obj reference.obj output.obj [GRID_RESOLUTION=64] [MESH_RESOLUTION=5000] [lambda=1] [symmetry=0]
// (reference src/app/cad_deform.cc:21-120, the branch without a flow file): clean-up + Subdivision of the CAD model,
// the reference's distance field, Deformer::DeformGraph on the deformation graph, Subdivision::LinearSolve, OBJ out.
//
// On the GPU through the C-ABI: Mesh::ConstructDistanceField (mo_template_create_normalized) and ceres::Solve of
// DeformGraph's problem (src/lib/deformer.cc:370-442: one DistanceLoss per graph node, one EdgeLoss per graph edge;
// mo_ceres_solve).  Host side, out of process (meshode_b200/cad_host.py, scipy in place of CGAL / Eigen): the
// subdivision with its neighbour pairs and deformation graph, and the sparse least-squares LinearSolve -- neither is
// on the hot path, and both are "parity unpinned" (Delaunay triangulations are not reproducible across libraries).
#include <unistd.h>

#include <climits>
#include <cstdint>
#include <string>

#include "deform_main.h"

namespace {

struct Bundle {
  std::vector<double> V, GV;
  std::vector<int> F, E, REF, GE;
};

bool read_bundle(const std::string& path, Bundle& b) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int32_t h[6];
  bool ok = fread(h, sizeof(int32_t), 6, f) == 6 && h[0] == 0x4d4f4344;
  if (ok) {
    b.V.resize(3 * (size_t)h[1]); b.F.resize(3 * (size_t)h[2]); b.E.resize(2 * (size_t)h[3]); b.REF.resize((size_t)h[1]);
    b.GV.resize(3 * (size_t)h[4]); b.GE.resize(2 * (size_t)h[5]);
    ok = fread(b.V.data(), 8, b.V.size(), f) == b.V.size() && fread(b.F.data(), 4, b.F.size(), f) == b.F.size() &&
         fread(b.E.data(), 4, b.E.size(), f) == b.E.size() && fread(b.REF.data(), 4, b.REF.size(), f) == b.REF.size() &&
         fread(b.GV.data(), 8, b.GV.size(), f) == b.GV.size() && fread(b.GE.data(), 4, b.GE.size(), f) == b.GE.size();
  }
  fclose(f);
  return ok;
}

bool write_bundle(const std::string& path, const Bundle& b) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const int32_t h[6] = {0x4d4f4344, (int32_t)(b.V.size() / 3), (int32_t)(b.F.size() / 3), (int32_t)(b.E.size() / 2),
                        (int32_t)(b.GV.size() / 3), (int32_t)(b.GE.size() / 2)};
  bool ok = fwrite(h, sizeof(int32_t), 6, f) == 6 && fwrite(b.V.data(), 8, b.V.size(), f) == b.V.size() &&
            fwrite(b.F.data(), 4, b.F.size(), f) == b.F.size() && fwrite(b.E.data(), 4, b.E.size(), f) == b.E.size() &&
            fwrite(b.REF.data(), 4, b.REF.size(), f) == b.REF.size() && fwrite(b.GV.data(), 8, b.GV.size(), f) == b.GV.size() &&
            fwrite(b.GE.data(), 4, b.GE.size(), f) == b.GE.size();
  return fclose(f) == 0 && ok;
}

// the host helper lives next to the library: <dir of this binary>/../cad_host.py
std::string helper_command() {
  char exe[PATH_MAX];
  const ssize_t n = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
  std::string dir = n > 0 ? std::string(exe, (size_t)n) : std::string("./cad_deform");
  dir = dir.substr(0, dir.find_last_of('/'));
  const char* py = getenv("MESHODE_PYTHON");
  return std::string(py ? py : "python3") + " '" + dir + "/../cad_host.py'";
}

}  // namespace

int main(int argc, char** argv) {
  using namespace mo_app;
  int GRID_RESOLUTION = 64;
  int MESH_RESOLUTION = 5000;
  if (argc < 5) {
    printf("./cad_deform cad.obj reference.obj output.obj "
           "[GRID_RESOLUTION=64] [MESH_RESOLUTION=5000] "
           "[lambda=1] [symmetry=0] [flow_output=filename]\n");
    return 0;
  }
  if (argc > 8) { fprintf(stderr, "cad_deform: the flow-file variant (DeformSubdivision with a callback) is not re-hosted\n"); return 1; }
  Mesh ref;
  if (!ref.ReadOBJ(argv[2])) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
  int symmetry = 0;
  if (argc > 7) sscanf(argv[7], "%d", &symmetry);
  if (symmetry) ref.ReflectionSymmetrize();

  // cad.RemoveDegenerated(); cad.MergeDuplex(); sub.Subdivide(cad, 2e-2); sub.ComputeGeometryNeighbors(1.5e-2);
  // sub.ComputeRepresentativeGraph(1e-2)   (cad_deform.cc:43-50), out of process
  const std::string tmp = std::string(argv[3]) + ".cad_host";
  const std::string helper = helper_command();
  if (system((helper + " prepare '" + argv[1] + "' '" + tmp + ".in'").c_str()) != 0) {
    fprintf(stderr, "cad_deform: the host-side subdivision failed (%s)\n", helper.c_str());
    return 1;
  }
  Bundle sub;
  if (!read_bundle(tmp + ".in", sub)) { fprintf(stderr, "cad_deform: cannot read %s.in\n", tmp.c_str()); return 1; }

  if (argc > 4) sscanf(argv[4], "%d", &GRID_RESOLUTION);
  if (argc > 5) sscanf(argv[5], "%d", &MESH_RESOLUTION);   // parsed and unused, as in the reference
  double lambda = 1;
  if (argc > 6) sscanf(argv[6], "%lf", &lambda);
  Mesh cad;   // the subdivided mesh
  cad.V = sub.V; cad.F = sub.F;
  std::cout << "Source:\t\t" << "Num vertices: " << cad.nV() << "\tNum faces: " << cad.nF() << std::endl;
  std::cout << "Reference:\t" << "Num vertices: " << ref.nV() << "\tNum faces: " << ref.nF() << std::endl << std::endl;

  ref.Normalize();
  cad.ApplyTransform(ref);                                 // sub.ApplyTransform(ref): the mesh ...
  for (size_t i = 0; i < sub.GV.size(); ++i) sub.GV[i] = (sub.GV[i] - ref.pos[i % 3]) / ref.scale;   // ... and the graph nodes (subdivision.cc:19-26)

  // UniformGrid grid(GRID_RESOLUTION); ref.ConstructDistanceField(grid);
  double* d_ref = nullptr; int* d_refF = nullptr;
  APP_CUDA(cudaMalloc(&d_ref, sizeof(double) * ref.V.size()));
  APP_CUDA(cudaMalloc(&d_refF, sizeof(int) * ref.F.size()));
  APP_CUDA(cudaMemcpy(d_ref, ref.V.data(), sizeof(double) * ref.V.size(), cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemcpy(d_refF, ref.F.data(), sizeof(int) * ref.F.size(), cudaMemcpyHostToDevice));
  int pid = -1;
  APP_MO(mo_template_create_normalized(d_ref, ref.nV(), d_refF, ref.nF(), GRID_RESOLUTION, ref.scale, ref.pos, nullptr, &pid));

  // Deformer::DeformGraph (deformer.cc:370-442): DistanceLoss per graph node, EdgeLoss(v = V[first] - V[second], lambda)
  // per graph edge, ceres::Solve with max_num_iterations = 100
  const int nG = (int)(sub.GV.size() / 3), nGE = (int)(sub.GE.size() / 2);
  std::vector<double> rest(3 * (size_t)nGE);
  for (int e = 0; e < nGE; ++e) {
    const int a = sub.GE[2 * e], b = sub.GE[2 * e + 1];
    if (a < 0 || a >= nG || b < 0 || b >= nG) { fprintf(stderr, "graph edge %d references a missing node\n", e); return 1; }
    for (int k = 0; k < 3; ++k) rest[3 * (size_t)e + k] = sub.GV[3 * (size_t)a + k] - sub.GV[3 * (size_t)b + k];
  }
  double *d_V = nullptr, *d_rest = nullptr; int* d_I = nullptr;
  APP_CUDA(cudaMalloc(&d_V, sizeof(double) * sub.GV.size() + 8));
  APP_CUDA(cudaMalloc(&d_rest, sizeof(double) * rest.size() + 8));
  APP_CUDA(cudaMalloc(&d_I, sizeof(int) * sub.GE.size() + 8));
  APP_CUDA(cudaMemcpy(d_V, sub.GV.data(), sizeof(double) * sub.GV.size(), cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemcpy(d_rest, rest.data(), sizeof(double) * rest.size(), cudaMemcpyHostToDevice));
  APP_CUDA(cudaMemcpy(d_I, sub.GE.data(), sizeof(int) * sub.GE.size(), cudaMemcpyHostToDevice));
  double summary[10];
  const int max_cg = getenv("MESHODE_MAX_CG") ? atoi(getenv("MESHODE_MAX_CG")) : 0;
  APP_MO(mo_ceres_solve(pid, MO_CERES_EDGE, d_V, nullptr, nG, d_I, d_rest, nGE, lambda, /*max_num_iterations*/ 100, max_cg, 0.0,
                        /*minimizer_progress_to_stdout*/ 1, summary, nullptr));
  std::cout << "Vertices cost: " << summary[2] << std::endl;
  std::cout << "Rigidity cost: " << summary[3] << std::endl;
  std::cout << "Final cost: " << summary[2] + summary[3] << std::endl;
  APP_CUDA(cudaMemcpy(sub.GV.data(), d_V, sizeof(double) * sub.GV.size(), cudaMemcpyDeviceToHost));

  // sub.LinearSolve(): LinearEstimation(V, F, neighbour pairs, representative_reference_, representative_vertices_)
  // with its default rigidity 2.0 (subdivision.cc:462-470, linear.h), out of process
  sub.V = cad.V;
  if (!write_bundle(tmp + ".lin", sub)) { fprintf(stderr, "cad_deform: cannot write %s.lin\n", tmp.c_str()); return 1; }
  if (system((helper + " linear '" + tmp + ".lin' '" + tmp + ".out' 2.0").c_str()) != 0) {
    fprintf(stderr, "cad_deform: the host-side linear solve failed\n");
    return 1;
  }
  {
    FILE* f = fopen((tmp + ".out").c_str(), "rb");
    const bool ok = f && fread(cad.V.data(), 8, cad.V.size(), f) == cad.V.size();
    if (f) fclose(f);
    if (!ok) { fprintf(stderr, "cad_deform: cannot read %s.out\n", tmp.c_str()); return 1; }
  }
  std::cout << "Deformed" << std::endl;
  if (!cad.WriteOBJ(argv[3])) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
  remove((tmp + ".in").c_str()); remove((tmp + ".lin").c_str()); remove((tmp + ".out").c_str());
  mo_template_destroy(pid);
  cudaFree(d_ref); cudaFree(d_refF); cudaFree(d_V); cudaFree(d_rest); cudaFree(d_I);
  return 0;
}
