// ref_iface_shim.cc -- C entry points around the REFERENCE's own interface loops, compiled where they lie under
// /root/reference (never copied): src/interface/distance_layer.cc, rigid_layer.cc, graph_layer.cc, cad_layer.cc,
// normalize.cc (with src/lib/uniformgrid.cc), against the stand-in third-party headers in oracle/stubs/ (Eigen, Ceres'
// Jet, and a torch::Tensor that is a plain host buffer).
//
// TEST INFRASTRUCTURE ONLY (output: oracle/_ref/libmeshode_ref.so, git-ignored).  It pins the oracle's restatement of
// SURVEY.md s8 rows a11-a16 (tests/test_oracle_ref.py) and generates tests/golden/golden_iface.npz
// (tests/golden/make_golden_iface.py).  Nothing in meshode_b200/ loads it.
//
// One piece of the reference is restated here rather than compiled: the DeformParams registry (CreateParams /
// GetParams, src/interface/deform_params.cc:7-14, a std::vector and two accessors).  Its translation unit also holds
// InitializeDeformTemplate, which needs libigl and CGAL through mesh.cc / mesh_tensor.cc and cannot be built here.
#include <cstddef>
#include <deque>

#include "cad_layer.h"        // reference: Store/CadEdgeLoss_* (src/interface/cad_layer.h)
#include "distance_layer.h"   // reference: DistanceFieldLoss_* (src/interface/distance_layer.h)
#include "graph_layer.h"      // reference: Store/GraphEdgeLoss_*
#include "normalize.h"        // reference: NormalizeByTemplate / DenormalizeByTemplate
#include "rigid_layer.h"      // reference: Store/RigidEdgeLoss_*

// ---- registry (deform_params.cc:7-14); a deque so that references stay valid while tests hold several ----------
static std::deque<DeformParams> g_params;
int CreateParams() {
  g_params.emplace_back();
  return (int)g_params.size() - 1;
}
DeformParams& GetParams(int param_id) { return g_params[(size_t)param_id]; }

// Mesh's members are defined in src/lib/mesh.cc (libigl, CGAL); DeformParams only needs construction
Mesh::Mesh() : scale_(1.0), pos_(0, 0, 0) {}

namespace {
torch::Tensor fview(const float* p, long long n, long long c) { return torch::Tensor::view(const_cast<float*>(p), {n, c}, torch::kFloat32); }
torch::Tensor fview1(const float* p, long long n) { return torch::Tensor::view(const_cast<float*>(p), {n}, torch::kFloat32); }
torch::Tensor iview(const int* p, long long n, long long c) { return torch::Tensor::view(const_cast<int*>(p), {n, c}, torch::kInt32); }
void copy_out(const torch::Tensor& t, float* out) { std::memcpy(out, t.storage().data(), sizeof(float) * (size_t)t.numel()); }
}  // namespace

extern "C" {

// what InitializeDeformTemplate leaves in a DeformParams (deform_params.cc:22-39): the grid filled through
// SetDistance(i=z, j=y, k=x) as Mesh::ConstructDistanceField does (mesh.cc:142-150), scale and translation
int ref_params_create(int N, const double* grid, double scale, const double* trans) {
  const int id = CreateParams();
  DeformParams& p = GetParams(id);
  p.grid = UniformGrid(N);
  size_t o = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j)
      for (int k = 0; k < N; ++k) p.grid.SetDistance(i, j, k, grid[o++]);
  p.scale = scale;
  p.trans = Vector3(trans[0], trans[1], trans[2]);
  return id;
}

void ref_iface_normalize(float* V, int n, int pid, int inverse) {
  if (inverse) DenormalizeByTemplate(fview(V, n, 3), pid);
  else NormalizeByTemplate(fview(V, n, 3), pid);
}
void ref_iface_dist_forward(const float* V, int n, int pid, float* out) { copy_out(DistanceFieldLoss_forward(fview(V, n, 3), pid), out); }
void ref_iface_dist_backward(const float* V, int n, int pid, float* out) { copy_out(DistanceFieldLoss_backward(fview(V, n, 3), pid), out); }

void ref_iface_rigid_store(const float* V, int n, const int* F, int m, int pid) { StoreRigidityInformation(fview(V, n, 3), iview(F, m, 3), pid); }
void ref_iface_rigid_forward(const float* V, int n, const int* F, int m, int pid, float* out) {
  copy_out(RigidEdgeLoss_forward(fview(V, n, 3), iview(F, m, 3), pid), out);
}
void ref_iface_rigid_backward(const float* V, int n, const int* F, int m, int pid, float* out) {
  copy_out(RigidEdgeLoss_backward(fview(V, n, 3), iview(F, m, 3), pid), out);
}

void ref_iface_graph_store(const float* V, int n, const int* E, int e, int pid) { StoreGraphInformation(fview(V, n, 3), iview(E, e, 2), pid); }
void ref_iface_graph_forward(const float* V, int n, const int* E, int e, int pid, float* out) {
  copy_out(GraphEdgeLoss_forward(fview(V, n, 3), iview(E, e, 2), pid), out);
}
void ref_iface_graph_backward(const float* V, int n, const int* E, int e, int pid, float* out) {
  copy_out(GraphEdgeLoss_backward(fview(V, n, 3), iview(E, e, 2), pid), out);
}

void ref_iface_cad_store(const float* V, int n, const int* F, int m, const int* E, int e, int pid) {
  StoreCadInformation(fview(V, n, 3), iview(F, m, 3), iview(E, e, 2), pid);
}
void ref_iface_cad_forward(const float* V, int n, const int* F, int m, const int* E, int e, int pid, float* out) {
  copy_out(CadEdgeLoss_forward(fview(V, n, 3), iview(F, m, 3), iview(E, e, 2), pid), out);
}
void ref_iface_cad_backward(const float* V, int n, const int* F, int m, const int* E, int e, int pid, float* out) {
  copy_out(CadEdgeLoss_backward(fview(V, n, 3), iview(F, m, 3), iview(E, e, 2), pid), out);
}
// the lambda the reference stored for CAD edge `i` (cad_layer.cc:48-49, :71-72)
float ref_iface_cad_lambda(int pid, int i) { return GetParams(pid).edge_lambda[(size_t)i]; }

}  // extern "C"
