// Throughput probe: scalar FFMA chains against packed fma.rn.f32x2 (FFMA2) chains on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f32x2_probe tools/probe/f32x2_probe.cu && build/f32x2_probe
// Development tool (decides whether packing two FP32 operations per instruction is worth a kernel rewrite).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_scalar(int iters, float* sink) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3f + j;
  const float m = 0.999f, c = 1e-3f + blockIdx.x * 1e-9f;
#pragma unroll 4
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
  float s = 0;
  for (int j = 0; j < 8; ++j) s += a[j];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(int iters, float* sink) {
  unsigned long long a[8], m, c;
  for (int j = 0; j < 8; ++j) {
    const float x = threadIdx.x * 1e-3f + j, y = x + 0.5f;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a[j]) : "f"(x), "f"(y));
  }
  { const float mm = 0.999f, cc = 1e-3f + blockIdx.x * 1e-9f;
    asm("mov.b64 %0, {%1,%1};" : "=l"(m) : "f"(mm));
    asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(cc)); }
#pragma unroll 4
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[j]) : "l"(m), "l"(c));
  float s = 0;
  for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[j])); s += x + y; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// half the warps' issue slots taken by integer work, the rest FFMA or FFMA2: does packing free issue slots?
__global__ void k_mixed(int iters, float* sink, int packed) {
  unsigned long long a[4], m, c;
  float b[4];
  unsigned z = threadIdx.x;
  for (int j = 0; j < 4; ++j) { b[j] = threadIdx.x * 1e-3f + j; asm("mov.b64 %0, {%1,%1};" : "=l"(a[j]) : "f"(b[j])); }
  { const float mm = 0.999f, cc = 1e-3f; asm("mov.b64 %0, {%1,%1};" : "=l"(m) : "f"(mm)); asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(cc)); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (packed) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[j]) : "l"(m), "l"(c));
      else { b[j] = fmaf(b[j], 0.999f, 1e-3f); asm volatile("" : "+f"(b[j])); b[j] = fmaf(b[j], 0.999f, 1e-3f); asm volatile("" : "+f"(b[j])); }
      z = z * 1664525u + 1013904223u; z ^= z >> 7;     // ALU / IMAD work sharing the issue slots
    }
  }
  float s = (float)z;
  for (int j = 0; j < 4; ++j) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[j])); s += x + y + b[j]; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 15;
  float* sink; cudaMalloc(&sink, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); k_scalar<<<blocks, threads>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * threads * iters * 8;
    printf("scalar FFMA : %.3f ms  %.1f G thread-inst/s  %.1f TFLOP/s\n", ms, inst / ms * 1e-6, 2 * inst / ms * 1e-9);
    cudaEventRecord(e0); k_packed<<<blocks, threads>>>(iters, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms  %.1f G thread-inst/s  %.1f TFLOP/s\n", ms, inst / ms * 1e-6, 4 * inst / ms * 1e-9);
    for (int p = 0; p < 2; ++p) {
      cudaEventRecord(e0); k_mixed<<<blocks, threads>>>(iters / 4, sink, p); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      printf("mixed (%s + integer work, same FLOPs): %.3f ms\n", p ? "FFMA2    " : "2 x FFMA ", ms);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
