// FP32 roofline denominator: a register-resident FFMA chain (no memory traffic), timed by the
// caller with CUDA events.  MEASURED_PEAKS.json carries no FP32 entry, so bench.py measures it.
#include "common.cuh"

namespace mo {
__global__ void k_ffma_peak(int iters, float* __restrict__ sink) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
        a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f + blockIdx.x * 1e-9f;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
// the same chain on packed operands (fma.rn.f32x2 -> FFMA2, two FMAs per instruction at half the issue rate):
// the highest FP32 rate the SM reaches, a few per cent above the scalar chain
__global__ void k_ffma2_peak(int iters, float* __restrict__ sink) {
  unsigned long long a[8], m, c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = threadIdx.x * 1e-3f + j, y = x + 0.5f;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a[j]) : "f"(x), "f"(y));
  }
  {
    const float mm = 0.999f, cc = 1e-3f + blockIdx.x * 1e-9f;
    asm("mov.b64 %0, {%1,%1};" : "=l"(m) : "f"(mm));
    asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(cc));
  }
#pragma unroll 4
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[j]) : "l"(m), "l"(c));
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[j])); s += x + y; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace mo

extern "C" int mo_microbench_fp32x2(int blocks, int threads, int iters, float* d_sink, mo_stream_t stream) {
  MO_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && d_sink, "bad microbench arguments");
  mo::k_ffma2_peak<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, d_sink);
  MO_LAUNCH_CHECK();
  return MO_OK;
}

extern "C" int mo_microbench_fp32(int blocks, int threads, int iters, float* d_sink, mo_stream_t stream) {
  MO_REQUIRE(blocks > 0 && threads > 0 && threads <= 1024 && iters > 0 && d_sink, "bad microbench arguments");
  mo::k_ffma_peak<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, d_sink);
  MO_LAUNCH_CHECK();
  return MO_OK;
}
