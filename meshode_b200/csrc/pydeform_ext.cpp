// The compiled `pyDeform` module: a torch C++ extension with the 18 functions of the reference's pybind11 module
// (reference src/interface/pydeform.cc:14-39; signatures src/interface/{mesh_tensor,deform_params,normalize,
// linear_layer,distance_layer,rigid_layer,cad_layer,graph_layer}.h), every hot entry a thin call into the C-ABI of
// libmeshode_b200.so (include/meshode_b200.h).  torch tensors cross the boundary as device pointers on torch's
// current CUDA stream.  Like the reference's module it links libtorch, so `import torch` must come first.
//
// CUDA tensors are used in place (zero copy, results on the same device).  CPU tensors -- what the reference's
// scripts pass, src/python/rigid_deform.py:25-41 -- are staged through the current CUDA device and results come
// back on the CPU; in-place functions write back into the caller's storage.  There is no CPU implementation here:
// without a CUDA device every compute entry raises.  Unlike the reference, dtype / shape / contiguity / param_id
// are checked (TORCH_CHECK -> Python exceptions).
//
// Host-side entries: LoadMesh / SaveMesh are the OBJ reader and writer of apps/mesh_host.h (src/lib/mesh.cc:14-64);
// LoadCadMesh and SolveLinear (CGAL / Eigen sparse code in the reference, neither on the hot path) call the
// scipy restatements in meshode_b200/cadmesh.py and meshode_b200/linear.py.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <string>
#include <vector>

#include "../../apps/mesh_host.h"
#include "../../include/meshode_b200.h"

namespace {

namespace py = pybind11;

void ok(int rc) { TORCH_CHECK(rc == MO_OK, "pyDeform: ", mo_last_error()); }

void want(const torch::Tensor& t, c10::ScalarType dtype, int64_t cols, const char* name) {
  TORCH_CHECK(t.defined(), name, " is undefined");
  TORCH_CHECK(t.scalar_type() == dtype, name, " must have dtype ", dtype, ", got ", t.scalar_type());
  TORCH_CHECK(t.dim() == 2 && t.size(1) == cols, name, " must have shape [n, ", cols, "]");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

c10::Device work_device(const torch::Tensor& lead) {
  if (lead.is_cuda()) return lead.device();
  TORCH_CHECK(mo_device_count() > 0, "pyDeform: no CUDA device (this build has no CPU path)");
  return c10::Device(c10::kCUDA, c10::cuda::current_device());
}

// A tensor as the kernels see it: the caller's own storage when it is on the GPU, a device copy otherwise.
struct OnDevice {
  torch::Tensor t;
  bool staged = false;
  OnDevice() = default;
  OnDevice(const torch::Tensor& src, c10::Device dev) {
    staged = !(src.is_cuda() && src.device() == dev);
    t = staged ? src.detach().to(dev) : src.detach();
  }
  template <class T> T* ptr() const { return t.defined() && t.numel() > 0 ? t.data_ptr<T>() : nullptr; }
  int rows() const { return t.defined() ? (int)t.size(0) : 0; }
};

struct Launch {   // device guard + the stream torch is currently enqueueing on
  c10::cuda::CUDAGuard guard;
  void* stream;
  explicit Launch(c10::Device dev) : guard(dev), stream((void*)c10::cuda::getCurrentCUDAStream(dev.index()).stream()) {}
};

torch::Tensor back(const torch::Tensor& res, const torch::Tensor& like) { return like.is_cuda() ? res : res.cpu(); }

// ---- template (src/interface/deform_params.cc:16-40, normalize.cc:5-45) ------------------------------------
int InitializeDeformTemplate(torch::Tensor tensorV, torch::Tensor tensorF, int symmetry, int grid_resolution) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  want(tensorF, torch::kInt32, 3, "tensorF");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev), F(tensorF, dev);
  Launch L(dev);
  int pid = -1;
  ok(mo_template_create(V.ptr<float>(), V.rows(), F.ptr<int>(), F.rows(), symmetry, grid_resolution, L.stream, &pid));
  return pid;
}

void by_template(torch::Tensor tensorV, int param_id, int inverse) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev);
  {
    Launch L(dev);
    ok(mo_normalize_by_template(V.ptr<float>(), V.rows(), param_id, inverse, L.stream));
  }
  if (V.staged) {
    torch::NoGradGuard ng;
    tensorV.copy_(V.t);
  }
}
void NormalizeByTemplate(torch::Tensor tensorV, int param_id) { by_template(tensorV, param_id, 0); }
void DenormalizeByTemplate(torch::Tensor tensorV, int param_id) { by_template(tensorV, param_id, 1); }
void DestroyTemplate(int param_id) {   // additive: g_params entries are never freed in the reference
  if (mo_device_count() > 0) ok(mo_template_destroy_async(param_id, (void*)c10::cuda::getCurrentCUDAStream().stream()));
  else ok(mo_template_destroy(param_id));
}

// ---- distance field loss (src/interface/distance_layer.cc:8-81) ---------------------------------------------
torch::Tensor DistanceFieldLoss_forward(torch::Tensor tensorV, int param_id) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev);
  auto out = torch::empty({V.rows()}, V.t.options());
  Launch L(dev);
  ok(mo_distance_forward(V.ptr<float>(), V.rows(), param_id, out.numel() ? out.data_ptr<float>() : nullptr, L.stream));
  return back(out, tensorV);
}
torch::Tensor DistanceFieldLoss_backward(torch::Tensor tensorV, int param_id) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev);
  auto out = torch::empty({V.rows(), 3}, V.t.options());
  Launch L(dev);
  ok(mo_distance_backward(V.ptr<float>(), V.rows(), param_id, out.numel() ? out.data_ptr<float>() : nullptr, L.stream));
  return back(out, tensorV);
}

// ---- edge rigidity (src/interface/{rigid,graph,cad}_layer.cc) -----------------------------------------------
enum class EdgeOp { Store, Forward, Backward };

torch::Tensor edges(EdgeOp op, int kind, const torch::Tensor& tensorV, const torch::Tensor* tensorF,
                    const torch::Tensor* tensorE, int param_id) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  if (tensorF) want(*tensorF, torch::kInt32, 3, "tensorF");
  if (tensorE) want(*tensorE, torch::kInt32, 2, "tensorE");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev), F, E;
  if (tensorF) F = OnDevice(*tensorF, dev);
  if (tensorE) E = OnDevice(*tensorE, dev);
  const int nF = F.rows(), nE = E.rows();
  Launch L(dev);
  if (op == EdgeOp::Store) {
    ok(mo_edges_store(param_id, kind, V.ptr<float>(), V.rows(), F.ptr<int>(), nF, E.ptr<int>(), nE, L.stream));
    return torch::Tensor();
  }
  const int64_t n_edges = kind == MO_EDGES_RIGID ? 3 * (int64_t)nF : kind == MO_EDGES_GRAPH ? nE : nE + 3 * (int64_t)nF;
  auto out = torch::empty({op == EdgeOp::Forward ? n_edges : (int64_t)V.rows(), 3}, V.t.options());
  float* o = out.numel() ? out.data_ptr<float>() : nullptr;
  if (op == EdgeOp::Forward)
    ok(mo_edges_forward(param_id, kind, V.ptr<float>(), V.rows(), F.ptr<int>(), nF, E.ptr<int>(), nE, o, L.stream));
  else
    ok(mo_edges_backward(param_id, kind, V.ptr<float>(), V.rows(), F.ptr<int>(), nF, E.ptr<int>(), nE, o, L.stream));
  return back(out, tensorV);
}

void StoreRigidityInformation(torch::Tensor V, torch::Tensor F, int pid) { edges(EdgeOp::Store, MO_EDGES_RIGID, V, &F, nullptr, pid); }
torch::Tensor RigidEdgeLoss_forward(torch::Tensor V, torch::Tensor F, int pid) { return edges(EdgeOp::Forward, MO_EDGES_RIGID, V, &F, nullptr, pid); }
torch::Tensor RigidEdgeLoss_backward(torch::Tensor V, torch::Tensor F, int pid) { return edges(EdgeOp::Backward, MO_EDGES_RIGID, V, &F, nullptr, pid); }
void StoreGraphInformation(torch::Tensor V, torch::Tensor E, int pid) { edges(EdgeOp::Store, MO_EDGES_GRAPH, V, nullptr, &E, pid); }
torch::Tensor GraphEdgeLoss_forward(torch::Tensor V, torch::Tensor E, int pid) { return edges(EdgeOp::Forward, MO_EDGES_GRAPH, V, nullptr, &E, pid); }
torch::Tensor GraphEdgeLoss_backward(torch::Tensor V, torch::Tensor E, int pid) { return edges(EdgeOp::Backward, MO_EDGES_GRAPH, V, nullptr, &E, pid); }
void StoreCadInformation(torch::Tensor V, torch::Tensor F, torch::Tensor E, int pid) { edges(EdgeOp::Store, MO_EDGES_CAD, V, &F, &E, pid); }
torch::Tensor CadEdgeLoss_forward(torch::Tensor V, torch::Tensor F, torch::Tensor E, int pid) { return edges(EdgeOp::Forward, MO_EDGES_CAD, V, &F, &E, pid); }
torch::Tensor CadEdgeLoss_backward(torch::Tensor V, torch::Tensor F, torch::Tensor E, int pid) { return edges(EdgeOp::Backward, MO_EDGES_CAD, V, &F, &E, pid); }

// ---- additive: the fused per-iteration loss and the whole Adam loop --------------------------
std::vector<torch::Tensor> LossForwardBackward(torch::Tensor tensorV, int dist_param_id, int edge_param_id, double w_edge,
                                               double mask_threshold) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  const auto dev = work_device(tensorV);
  OnDevice V(tensorV, dev);
  auto loss = torch::empty({}, V.t.options().dtype(torch::kFloat64));
  auto grad = torch::empty({V.rows(), 3}, V.t.options());
  Launch L(dev);
  ok(mo_loss_forward_backward(dist_param_id, edge_param_id, V.ptr<float>(), V.rows(), (float)w_edge, (float)mask_threshold,
                              loss.data_ptr<double>(), grad.numel() ? grad.data_ptr<float>() : nullptr, L.stream));
  return {back(loss, tensorV), back(grad, tensorV)};
}

void DeformBatchAdam(std::vector<torch::Tensor> V, std::vector<int> pids, int iters, double lr) {
  TORCH_CHECK(V.size() == pids.size(), "one param_id per vertex tensor");
  if (V.empty()) return;
  std::vector<float*> ptr;
  for (auto& v : V) {
    want(v, torch::kFloat32, 3, "V[i]");
    TORCH_CHECK(v.is_cuda() && v.device() == V[0].device(), "DeformBatchAdam optimises in place: CUDA tensors on one device");
    ptr.push_back(v.data_ptr<float>());
  }
  Launch L(V[0].device());
  py::gil_scoped_release nogil;
  ok(mo_deform_batch_adam(pids.data(), pids.data(), ptr.data(), (int)V.size(), iters, lr, 0.9, 0.999, 1e-8, MO_DEFORM_EXACT,
                          L.stream));
}

// ---- host side (src/interface/mesh_tensor.cc:87-186, linear_layer.cc:5-84) -----------------------------------
std::vector<torch::Tensor> LoadMesh(const char* filename) {
  mo_app::Mesh m;
  TORCH_CHECK(m.ReadOBJ(filename), "LoadMesh: cannot read ", filename);
  auto V = torch::empty({m.nV(), 3}, torch::kFloat32);
  auto F = torch::empty({m.nF(), 3}, torch::kInt32);
  float* v = V.data_ptr<float>();
  for (size_t i = 0; i < m.V.size(); ++i) v[i] = (float)m.V[i];     // CopyMeshToTensor: FT -> float32 (mesh_tensor.cc:8-30)
  if (!m.F.empty()) std::memcpy(F.data_ptr<int>(), m.F.data(), sizeof(int) * m.F.size());
  return {V, F};
}

void SaveMesh(const char* filename, const torch::Tensor& tensorV, const torch::Tensor& tensorF) {
  want(tensorV, torch::kFloat32, 3, "tensorV");
  want(tensorF, torch::kInt32, 3, "tensorF");
  auto V = tensorV.detach().cpu().contiguous();
  auto F = tensorF.detach().cpu().contiguous();
  mo_app::Mesh m;                                                   // CopyTensorToMesh without normalisation
  m.V.assign(V.data_ptr<float>(), V.data_ptr<float>() + V.numel());
  m.F.assign(F.data_ptr<int>(), F.data_ptr<int>() + F.numel());
  TORCH_CHECK(m.WriteOBJ(filename, /*normalized=*/true), "SaveMesh: cannot write ", filename);
}

py::object host_module() { return py::module_::import("meshode_b200.pyDeform"); }

py::object LoadCadMesh(const char* filename) { return host_module().attr("LoadCadMesh")(filename); }

void SolveLinear(torch::Tensor tensorV, torch::Tensor tensorF, torch::Tensor tensorE, torch::Tensor tensorRef,
                 torch::Tensor tensorGraphV, double rigidity, int with_rot) {
  host_module().attr("SolveLinear")(tensorV, tensorF, tensorE, tensorRef, tensorGraphV, rigidity, with_rot);
}

}  // namespace

PYBIND11_MODULE(pyDeform, m) {
  m.doc() = "MeshODE's pyDeform on libmeshode_b200.so (sm_100a); same 18 functions as src/interface/pydeform.cc";
  m.def("LoadMesh", &LoadMesh);
  m.def("LoadCadMesh", &LoadCadMesh);
  m.def("SaveMesh", &SaveMesh);

  m.def("InitializeDeformTemplate", &InitializeDeformTemplate);
  m.def("NormalizeByTemplate", &NormalizeByTemplate);
  m.def("DenormalizeByTemplate", &DenormalizeByTemplate);
  m.def("SolveLinear", &SolveLinear);

  m.def("DistanceFieldLoss_forward", &DistanceFieldLoss_forward);
  m.def("DistanceFieldLoss_backward", &DistanceFieldLoss_backward);

  m.def("RigidEdgeLoss_forward", &RigidEdgeLoss_forward);
  m.def("RigidEdgeLoss_backward", &RigidEdgeLoss_backward);
  m.def("StoreRigidityInformation", &StoreRigidityInformation);

  m.def("CadEdgeLoss_forward", &CadEdgeLoss_forward);
  m.def("CadEdgeLoss_backward", &CadEdgeLoss_backward);
  m.def("StoreCadInformation", &StoreCadInformation);

  m.def("GraphEdgeLoss_forward", &GraphEdgeLoss_forward);
  m.def("GraphEdgeLoss_backward", &GraphEdgeLoss_backward);
  m.def("StoreGraphInformation", &StoreGraphInformation);

  // additive (not in the reference)
  m.def("DestroyTemplate", &DestroyTemplate);
  m.def("LossForwardBackward", &LossForwardBackward);
  m.def("DeformBatchAdam", &DeformBatchAdam);
  m.attr("__backend__") = "libmeshode_b200.so (compiled torch extension)";
}
