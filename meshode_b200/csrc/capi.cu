// C-ABI of libmeshode_b200.so: the handle table that replaces the reference's global
// g_params vector (src/interface/deform_params.cc:7-14) and the thin entry points declared
// in include/meshode_b200.h.
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mo {

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_build_stats{0};
static thread_local std::string t_error;
void set_error(const std::string& s) { t_error = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  t_error = buf;
  cudaGetLastError();   // clear the sticky flag of non-fatal errors
  return MO_ERR_CUDA;
}

cudaError_t dev_alloc_bytes(void** p, size_t bytes, cudaStream_t s) {
  static std::once_flag once[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::call_once(once[dev & 63], [dev]() {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
  });
  return cudaMallocAsync(p, bytes ? bytes : 1, s);
}
void dev_free(void* p, cudaStream_t s) {
  if (p) cudaFreeAsync(p, s);
}

namespace {

std::mutex g_mutex;
std::vector<std::shared_ptr<Template>> g_templates;   // index = param_id, like g_params

// Entry points hold the template through a shared_ptr for the duration of the call, so a concurrent
// mo_template_destroy cannot pull it from under them (the device buffers go when the last holder lets go).
// A template belongs to the device it was created on: using it while another device is current is refused
// here rather than left to fault with an illegal address inside a kernel.
typedef std::shared_ptr<Template> TemplateRef;
TemplateRef lookup(int pid) {
  TemplateRef T;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (pid >= 0 && pid < (int)g_templates.size()) T = g_templates[pid];
  }
  if (!T) {
    set_error("bad param_id " + std::to_string(pid));
    return nullptr;
  }
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = -1; }
  if (dev != T->device) {
    set_error("param_id " + std::to_string(pid) + " lives on CUDA device " + std::to_string(T->device) +
              ", the current device is " + std::to_string(dev));
    return nullptr;
  }
  return T;
}

void release(Template& T, cudaStream_t s = 0) {
  free_edges(T, s);
  dev_free(T.d_Vn, s); dev_free(T.d_F, s); dev_free(T.d_grid64, s); dev_free(T.d_grid32, s); dev_free(T.d_nearest, s);
  dev_free(T.d_xf, s); dev_free(T.d_stats, s); dev_free(T.d_cells, s);
  T.d_Vn = nullptr; T.d_F = nullptr; T.d_grid64 = nullptr; T.d_grid32 = nullptr; T.d_nearest = nullptr;
  T.d_xf = nullptr; T.d_stats = nullptr; T.d_cells = nullptr;
}

// the buffers of a destroyed template go back to the pool in the order of `free_stream` when the last holder
// (the table or an entry point still running with it) lets go
struct TemplateDeleter {
  void operator()(Template* T) const {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != T->device) cudaSetDevice(T->device);
    release(*T, T->free_stream);
    if (cur != T->device && cur >= 0) cudaSetDevice(cur);
    delete T;
  }
};

int allocate(Template& T, cudaStream_t s) {
  const size_t nvox = (size_t)T.N * T.N * T.N;
  MO_CUDA(cudaGetDevice(&T.device));
  MO_CUDA(dev_alloc(&T.d_Vn, 3 * (size_t)T.nV, s));
  MO_CUDA(dev_alloc(&T.d_F, 3 * (size_t)T.nF, s));
  MO_CUDA(dev_alloc(&T.d_grid64, nvox, s));
  MO_CUDA(dev_alloc(&T.d_grid32, nvox, s));
  MO_CUDA(dev_alloc(&T.d_nearest, nvox, s));
  MO_CUDA(dev_alloc(&T.d_xf, 4, s));
  MO_CUDA(dev_alloc(&T.d_stats, 8 * (size_t)kStatSlots, s));
  return MO_OK;
}

int publish(std::unique_ptr<Template>& T, int* out) {
  std::lock_guard<std::mutex> lock(g_mutex);
  g_templates.push_back(TemplateRef(T.release(), TemplateDeleter()));
  *out = (int)g_templates.size() - 1;
  return MO_OK;
}

int create_common(const float* d_V, const double* d_Vn, int nV, const int* d_F, int nF, int N, int z0, int z1, double scale,
                  const double* h_trans, cudaStream_t s, int* out, int tz_first = 0, int tz_stride = 1) {
  MO_REQUIRE(out != nullptr, "out_param_id is null");
  MO_REQUIRE((d_V != nullptr || d_Vn != nullptr) && d_F != nullptr, "null vertex / face pointer");
  MO_REQUIRE(nV > 0 && nF > 0, "empty mesh");
  MO_REQUIRE(N >= 2 && N <= 1024, "grid_resolution must be in [2, 1024]");
  MO_REQUIRE(0 <= z0 && z0 < z1 && z1 <= N, "bad z-slab");
  MO_REQUIRE(tz_stride >= 1 && tz_first >= 0 && (tz_stride == 1 ? tz_first == 0 : tz_first < tz_stride), "bad layer selection");
  std::unique_ptr<Template> T(new Template());
  T->N = N; T->nV = nV; T->nF = nF; T->z0 = z0; T->z1 = z1; T->tz_first = tz_first; T->tz_stride = tz_stride;
  int rc = allocate(*T, s);
  if (rc != MO_OK) { release(*T, s); return rc; }
  cudaError_t e = cudaMemcpyAsync(T->d_F, d_F, sizeof(int) * 3 * (size_t)nF, cudaMemcpyDeviceToDevice, s);
  if (e == cudaSuccess && d_Vn) {
    e = cudaMemcpyAsync(T->d_Vn, d_Vn, sizeof(double) * 3 * (size_t)nV, cudaMemcpyDeviceToDevice, s);
    const double xf[4] = {scale, h_trans ? h_trans[0] : 0.0, h_trans ? h_trans[1] : 0.0, h_trans ? h_trans[2] : 0.0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(T->d_xf, xf, sizeof(xf), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // xf lives on this frame
  }
  if (e != cudaSuccess) { release(*T, s); return cuda_fail(e, "template upload", __FILE__, __LINE__); }
  rc = d_Vn ? build_field_from_normalized(*T, s) : build_field_from_f32(*T, d_V, s);
  if (rc != MO_OK) { release(*T, s); return rc; }
  return publish(T, out);
}

}  // namespace
}  // namespace mo

namespace mo {
__global__ void k_normalize(float* __restrict__ V, int n3, const double* __restrict__ xf, int inverse) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const double scale = xf[0], t = xf[1 + i % 3];
  const double v = (double)V[i];
  // normalize.cc:21 : (v - trans)/scale ; normalize.cc:42 : v*scale + trans   (FP64, rounded once)
  V[i] = inverse ? (float)dadd(dmul(v, scale), t) : (float)__ddiv_rn(dsub(v, t), scale);
}
}  // namespace mo

using namespace mo;

extern "C" {

int mo_version(void) { return 1; }
unsigned long long mo_launch_count(void) { return g_launches.load(); }
const char* mo_last_error(void) { return t_error.c_str(); }
int mo_build_stats_enable(int on) { return g_build_stats.exchange(on ? 1 : 0); }
int mo_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int mo_template_create(const float* d_V, int nV, const int* d_F, int nF, int symmetry, int grid_res, mo_stream_t stream,
                       int* out_param_id) {
  (void)symmetry;   // no-op in the reference: deform_params.cc:28-33
  return create_common(d_V, nullptr, nV, d_F, nF, grid_res, 0, grid_res, 1.0, nullptr, (cudaStream_t)stream, out_param_id);
}

int mo_template_create_slab(const float* d_V, int nV, const int* d_F, int nF, int grid_res, int z0, int z1,
                            mo_stream_t stream, int* out_param_id) {
  return create_common(d_V, nullptr, nV, d_F, nF, grid_res, z0, z1, 1.0, nullptr, (cudaStream_t)stream, out_param_id);
}

int mo_template_create_layers(const float* d_V, int nV, const int* d_F, int nF, int grid_res, int first_layer, int layer_stride,
                              mo_stream_t stream, int* out_param_id) {
  return create_common(d_V, nullptr, nV, d_F, nF, grid_res, 0, grid_res, 1.0, nullptr, (cudaStream_t)stream, out_param_id,
                       first_layer, layer_stride);
}

int mo_template_create_normalized(const double* d_Vn, int nV, const int* d_F, int nF, int grid_res, double scale,
                                  const double* h_trans3, mo_stream_t stream, int* out_param_id) {
  return create_common(nullptr, d_Vn, nV, d_F, nF, grid_res, 0, grid_res, scale, h_trans3, (cudaStream_t)stream, out_param_id);
}

static int destroy_common(int param_id, cudaStream_t s, bool device_sync) {
  TemplateRef T;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (param_id < 0 || param_id >= (int)g_templates.size() || !g_templates[param_id]) {
      set_error("bad param_id " + std::to_string(param_id));
      return MO_ERR_BAD_HANDLE;
    }
    T.swap(g_templates[param_id]);
  }
  if (device_sync) {
    // no stream given: work that still reads the template may sit on any (non-blocking) stream of its device,
    // which the legacy default stream does not order -- wait for the device, like cudaFree does
    int cur = -1;
    MO_CUDA(cudaGetDevice(&cur));
    if (cur != T->device) MO_CUDA(cudaSetDevice(T->device));
    const cudaError_t e = cudaDeviceSynchronize();
    if (cur != T->device) cudaSetDevice(cur);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceSynchronize", __FILE__, __LINE__);
  }
  T->free_stream = s;
  return MO_OK;   // T goes out of scope: freed now, or when a concurrent entry point drops its reference
}

int mo_template_destroy(int param_id) { return destroy_common(param_id, 0, true); }

int mo_template_destroy_async(int param_id, mo_stream_t stream) { return destroy_common(param_id, (cudaStream_t)stream, false); }

int mo_template_info(int param_id, mo_stream_t stream, int* grid_res, int* nV, int* nF, double* scale, double* trans3) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  if (grid_res) *grid_res = T->N;
  if (nV) *nV = T->nV;
  if (nF) *nF = T->nF;
  if (scale || trans3) {
    double xf[4];
    MO_CUDA(cudaMemcpyAsync(xf, T->d_xf, sizeof(xf), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (scale) *scale = xf[0];
    if (trans3) { trans3[0] = xf[1]; trans3[1] = xf[2]; trans3[2] = xf[3]; }
  }
  return MO_OK;
}

int mo_template_grid(int param_id, const double** d_grid_f64, const float** d_grid_f32, const int** d_nearest) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  if (d_grid_f64) *d_grid_f64 = T->d_grid64;
  if (d_grid_f32) *d_grid_f32 = T->d_grid32;
  if (d_nearest) *d_nearest = T->d_nearest;
  return MO_OK;
}

int mo_template_copy_grid(int param_id, int direction, int z0, int z1, double* d_grid_f64, float* d_grid_f32,
                          int* d_nearest, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(0 <= z0 && z0 <= z1 && z1 <= T->N, "bad slice range");
  MO_REQUIRE(direction == 0 || direction == 1, "direction must be 0 or 1");
  const size_t off = (size_t)z0 * T->N * T->N, cnt = (size_t)(z1 - z0) * T->N * T->N;
  cudaStream_t s = (cudaStream_t)stream;
  if (cnt == 0) return MO_OK;
  if (direction == 1 && T->d_cells) { dev_free(T->d_cells, s); T->d_cells = nullptr; }   // derived from grid_f32
  if (d_grid_f64)
    MO_CUDA(cudaMemcpyAsync(direction ? (void*)(T->d_grid64 + off) : (void*)(d_grid_f64 + off),
                            direction ? (const void*)(d_grid_f64 + off) : (const void*)(T->d_grid64 + off),
                            cnt * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (d_grid_f32)
    MO_CUDA(cudaMemcpyAsync(direction ? (void*)(T->d_grid32 + off) : (void*)(d_grid_f32 + off),
                            direction ? (const void*)(d_grid_f32 + off) : (const void*)(T->d_grid32 + off),
                            cnt * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (d_nearest)
    MO_CUDA(cudaMemcpyAsync(direction ? (void*)(T->d_nearest + off) : (void*)(d_nearest + off),
                            direction ? (const void*)(d_nearest + off) : (const void*)(T->d_nearest + off),
                            cnt * sizeof(int), cudaMemcpyDeviceToDevice, s));
  return MO_OK;
}

int mo_template_vertices(int param_id, const double** d_Vn) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  if (d_Vn) *d_Vn = T->d_Vn;
  return MO_OK;
}

int mo_template_build_stats(int param_id, mo_stream_t stream, unsigned long long* fp32_tests, unsigned long long* fp64_tests,
                            unsigned long long* cull_tests, unsigned long long* disc_tests) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  unsigned long long all[8 * kStatSlots], h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  MO_CUDA(cudaMemcpyAsync(all, T->d_stats, sizeof(all), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  MO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (int sl = 0; sl < kStatSlots; ++sl)
    for (int k = 0; k < 8; ++k) h[k] += all[8 * sl + k];
  h[3] = all[3];   // error bits live in slot 0
  if (fp32_tests) *fp32_tests = h[0];
  if (fp64_tests) *fp64_tests = h[1];
  if (cull_tests) *cull_tests = h[2];
  if (disc_tests) *disc_tests = h[4];
  if (h[3]) {
    set_error("template holds invalid input: " + std::string((h[3] & 1) ? "[face index out of range] " : "") +
              ((h[3] & 2) ? "[non-finite vertex] " : "") + ((h[3] & 4) ? "[edge index out of range]" : ""));
    return MO_ERR_BAD_ARG;
  }
  return MO_OK;
}

int mo_distance_forward(const float* d_V, int n, int param_id, float* d_out, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(n >= 0 && (n == 0 || (d_V && d_out)), "null pointer");
  return launch_distance_f32(*T, d_V, n, d_out, nullptr, 1, (cudaStream_t)stream);
}

int mo_distance_backward(const float* d_V, int n, int param_id, float* d_grad, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(n >= 0 && (n == 0 || (d_V && d_grad)), "null pointer");
  return launch_distance_f32(*T, d_V, n, nullptr, d_grad, 2, (cudaStream_t)stream);
}

int mo_distance_forward_backward(const float* d_V, int n, int param_id, float* d_out, float* d_grad, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(n >= 0 && (n == 0 || (d_V && d_out && d_grad)), "null pointer");
  return launch_distance_f32(*T, d_V, n, d_out, d_grad, 3, (cudaStream_t)stream);
}

int mo_distance_f64(const double* d_P, int n, int param_id, double* d_val, double* d_grad, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(n >= 0 && (n == 0 || (d_P && d_val)), "null pointer");
  return launch_distance_f64(*T, d_P, n, d_val, d_grad, (cudaStream_t)stream);
}

static int check_edge_args(int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE) {
  MO_REQUIRE(kind == MO_EDGES_RIGID || kind == MO_EDGES_GRAPH || kind == MO_EDGES_CAD, "unknown edge kind");
  MO_REQUIRE(nV >= 0 && nF >= 0 && nE >= 0, "negative size");
  MO_REQUIRE(nV == 0 || d_V, "null vertex pointer");
  MO_REQUIRE(kind == MO_EDGES_GRAPH || nF == 0 || d_F, "null face pointer");
  MO_REQUIRE(kind == MO_EDGES_RIGID || nE == 0 || d_E, "null edge pointer");
  return MO_OK;
}

static void canon(int kind, int& nF, int& nE) {
  if (kind == MO_EDGES_RIGID) nE = 0;
  if (kind == MO_EDGES_GRAPH) nF = 0;
}

int mo_edges_store(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                   mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  int rc = check_edge_args(kind, d_V, nV, d_F, nF, d_E, nE);
  if (rc != MO_OK) return rc;
  canon(kind, nF, nE);
  return edges_store(*T, kind, d_V, nV, d_F, nF, d_E, nE, (cudaStream_t)stream);
}

static int check_stored(const Template& T, int kind, int nV, int nF, int nE, bool need_csr) {
  if (T.kind != kind) { set_error("edges of this kind were not stored for this template"); return MO_ERR_STATE; }
  MO_REQUIRE(nF == T.eF && nE == T.eE, "edge counts differ from the stored ones");
  MO_REQUIRE(!need_csr || nV == T.eV, "vertex count differs from the stored one");
  return MO_OK;
}

int mo_edges_forward(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                     float* d_out, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  int rc = check_edge_args(kind, d_V, nV, d_F, nF, d_E, nE);
  if (rc != MO_OK) return rc;
  canon(kind, nF, nE);
  rc = check_stored(*T, kind, nV, nF, nE, false);
  if (rc != MO_OK) return rc;
  MO_REQUIRE(T->nEdges == 0 || d_out, "null output pointer");
  return edges_forward(*T, kind, d_V, nV, d_F, nF, d_E, nE, d_out, (cudaStream_t)stream);
}

int mo_edges_backward(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E, int nE,
                      float* d_grad, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  int rc = check_edge_args(kind, d_V, nV, d_F, nF, d_E, nE);
  if (rc != MO_OK) return rc;
  canon(kind, nF, nE);
  rc = check_stored(*T, kind, nV, nF, nE, true);
  if (rc != MO_OK) return rc;
  MO_REQUIRE(nV == 0 || d_grad, "null output pointer");
  return edges_backward(*T, d_V, nV, d_grad, (cudaStream_t)stream);
}

int mo_edges_backward_atomic(int param_id, int kind, const float* d_V, int nV, const int* d_F, int nF, const int* d_E,
                             int nE, float* d_grad, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  int rc = check_edge_args(kind, d_V, nV, d_F, nF, d_E, nE);
  if (rc != MO_OK) return rc;
  canon(kind, nF, nE);
  rc = check_stored(*T, kind, nV, nF, nE, false);
  if (rc != MO_OK) return rc;
  MO_REQUIRE(nV == 0 || d_grad, "null output pointer");
  return edges_backward_atomic(*T, kind, d_V, nV, d_F, nF, d_E, nE, d_grad, (cudaStream_t)stream);
}

int mo_loss_forward_backward(int dist_param_id, int edge_param_id, const float* d_V, int nV, float w_edge,
                             float mask_threshold, double* d_loss, float* d_grad, mo_stream_t stream) {
  const TemplateRef TD = lookup(dist_param_id);
  if (!TD) return MO_ERR_BAD_HANDLE;
  TemplateRef TE;
  if (edge_param_id >= 0) {
    TE = lookup(edge_param_id);
    if (!TE) return MO_ERR_BAD_HANDLE;
    if (TE->kind == MO_EDGES_NONE) { set_error("no edges stored for edge_param_id"); return MO_ERR_STATE; }
    MO_REQUIRE(nV == TE->eV, "vertex count differs from the stored one");
  }
  MO_REQUIRE(nV >= 0 && (nV == 0 || d_V), "null pointer");
  return loss_fused(*TD, TE.get(), d_V, nV, w_edge, mask_threshold, d_loss, d_grad, (cudaStream_t)stream);
}


int mo_deform_batch_adam(const int* h_dist_pids, const int* h_edge_pids, float* const* h_dV, int B, int iters, double lr,
                         double beta1, double beta2, double eps, int flags, mo_stream_t stream) {
  MO_REQUIRE(B >= 0 && iters >= 0, "negative count");
  MO_REQUIRE(B == 0 || (h_dist_pids && h_edge_pids && h_dV), "null pointer");
  std::vector<TemplateRef> hold(2 * (size_t)B);
  std::vector<Template*> td(B), te(B);
  for (int i = 0; i < B; ++i) {
    hold[2 * i] = lookup(h_dist_pids[i]);
    hold[2 * i + 1] = lookup(h_edge_pids[i]);
    td[i] = hold[2 * i].get();
    te[i] = hold[2 * i + 1].get();
    if (!td[i] || !te[i]) return MO_ERR_BAD_HANDLE;
    MO_REQUIRE(h_dV[i] != nullptr, "null vertex pointer");
  }
  return deform_batch_adam(td.data(), te.data(), h_dV, B, iters, lr, beta1, beta2, eps, flags, (cudaStream_t)stream);
}

int mo_deform_adam_large(int dist_pid, int edge_pid, float* d_V, int nV, float w_edge, float mask_threshold, int iters,
                         double lr, double beta1, double beta2, double eps, mo_stream_t stream) {
  const TemplateRef TD = lookup(dist_pid);
  const TemplateRef TE = lookup(edge_pid);
  if (!TD || !TE) return MO_ERR_BAD_HANDLE;
  if (TE->kind == MO_EDGES_NONE) { set_error("no edges stored for edge_pid"); return MO_ERR_STATE; }
  MO_REQUIRE(nV == TE->eV && (nV == 0 || d_V) && iters >= 0, "vertex count differs from the stored one / null pointer");
  return deform_adam_large(*TD, *TE, d_V, nV, w_edge, mask_threshold, iters, lr, beta1, beta2, eps, (cudaStream_t)stream);
}

int mo_ceres_edges(int kind, const double* d_V, const double* d_R, int nV, const int* d_I, const double* d_rest, int nE,
                   double lambda, double* d_res, double* d_jac, mo_stream_t stream) {
  MO_REQUIRE(kind == MO_CERES_EDGE || kind == MO_CERES_ADAPTIVE_EDGE || kind == MO_CERES_ROT_EDGE, "unknown Ceres edge kind");
  MO_REQUIRE(nV >= 0 && nE >= 0, "negative size");
  MO_REQUIRE(nE == 0 || (d_V && d_I && d_rest), "null pointer");
  MO_REQUIRE(kind != MO_CERES_ROT_EDGE || nE == 0 || d_R, "EdgeLossWithRot needs the rotation parameters");
  return ceres_edges(kind, d_V, d_R, nV, d_I, d_rest, nE, lambda, d_res, d_jac, (cudaStream_t)stream);
}

int mo_ceres_problem(int dist_param_id, int kind, const double* d_V, const double* d_R, int nV, const int* d_I,
                     const double* d_rest, int nE, double lambda, double* d_cost2, double* d_gV, double* d_gR,
                     mo_stream_t stream) {
  MO_REQUIRE(kind == MO_CERES_EDGE || kind == MO_CERES_ADAPTIVE_EDGE || kind == MO_CERES_ROT_EDGE, "unknown Ceres edge kind");
  MO_REQUIRE(nV >= 0 && nE >= 0, "negative size");
  MO_REQUIRE(nV == 0 || d_V, "null vertex pointer");
  MO_REQUIRE(nE == 0 || (d_I && d_rest), "null edge pointer");
  MO_REQUIRE(kind != MO_CERES_ROT_EDGE || nE == 0 || d_R, "EdgeLossWithRot needs the rotation parameters");
  TemplateRef TD;
  if (dist_param_id >= 0) {
    TD = lookup(dist_param_id);
    if (!TD) return MO_ERR_BAD_HANDLE;
  }
  return ceres_problem(TD.get(), kind, d_V, d_R, nV, d_I, d_rest, nE, lambda, d_cost2, d_gV, d_gR, (cudaStream_t)stream);
}

int mo_ceres_solve(int dist_param_id, int kind, double* d_V, double* d_R, int nV, const int* d_I, const double* d_rest,
                   int nE, double lambda, int max_iterations, int max_cg_iterations, double cg_tolerance, int verbose,
                   double* h_summary, mo_stream_t stream) {
  MO_REQUIRE(kind == MO_CERES_EDGE || kind == MO_CERES_ADAPTIVE_EDGE || kind == MO_CERES_ROT_EDGE, "unknown Ceres edge kind");
  MO_REQUIRE(nV >= 0 && nE >= 0 && max_iterations >= 0, "negative size");
  MO_REQUIRE(nV == 0 || d_V, "null vertex pointer");
  MO_REQUIRE(nE == 0 || (d_I && d_rest), "null edge pointer");
  MO_REQUIRE(kind != MO_CERES_ROT_EDGE || nV == 0 || d_R, "EdgeLossWithRot needs the rotation parameters");
  TemplateRef TD;
  if (dist_param_id >= 0) {
    TD = lookup(dist_param_id);
    if (!TD) return MO_ERR_BAD_HANDLE;
  }
  return ceres_solve(TD.get(), kind, d_V, d_R, nV, d_I, d_rest, nE, lambda, max_iterations, max_cg_iterations, cg_tolerance, verbose,
                     h_summary, (cudaStream_t)stream);
}

int mo_normalize_by_template(float* d_V, int n, int param_id, int inverse, mo_stream_t stream) {
  const TemplateRef T = lookup(param_id);
  if (!T) return MO_ERR_BAD_HANDLE;
  MO_REQUIRE(n >= 0 && (n == 0 || d_V), "null pointer");
  if (n == 0) return MO_OK;
  k_normalize<<<div_up(3LL * n, 256), 256, 0, (cudaStream_t)stream>>>(d_V, 3 * n, T->d_xf, inverse);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

}  // extern "C"
