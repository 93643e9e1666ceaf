// MUFU (XU pipe) throughput on this GPU: results per clock and SM for ex2.approx.ftz, rcp.approx.ftz, tanh.approx, and for a software
// reciprocal (integer seed + 2 / 3 Newton steps on the FMA pipe).  148 x 4 CTAs of 256 threads, 8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe_mufu.cu -o tools/_bin/probe_mufu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ float op(float x) {
  float y;
  if (MODE == 0) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  if (MODE == 1) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  if (MODE == 2) { asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
  // software reciprocal of x >= 1: seed from the exponent bits, Newton steps r <- r (2 - x r)
  float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
  r = r * fmaf(-x, r, 2.0f);
  r = r * fmaf(-x, r, 2.0f);
  if (MODE == 4) r = r * fmaf(-x, r, 2.0f);
  return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed + 0.001f * (threadIdx.x + 37 * j);
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = op<MODE>(v[j]) + 1.25f;       // keeps the argument in [1, ~3]; one FADD rides along
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out + gridDim.x * blockDim.x)[0] = t1 - t0;
}
template <int MODE> void run(const char* name, float* out, int grid) {
  const int iters = 4096;
  k<MODE><<<grid, 256>>>(out, 64, 1.5f);
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a)); k<MODE><<<grid, 256>>>(out, iters, 1.5f); CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  long long cyc; CK(cudaMemcpy(&cyc, out + grid * 256, 8, cudaMemcpyDeviceToHost));
  const double ops = (double)grid * 256 * 8 * iters;
  printf("%-28s %7.2f results / clk / SM   (%.1f G results/s, kernel %.3f ms, %lld cycles)\n", name, ops / 148.0 / (double)cyc, ops / ms / 1e6, ms, cyc);
}
int main() {
  float* out; CK(cudaMalloc(&out, (148 * 8 * 256 + 16) * 4));
  for (int occ : {4, 8}) {
    printf("--- %d CTAs of 256 threads per SM\n", occ);
    run<0>("ex2.approx.ftz.f32", out, 148 * occ);
    run<1>("rcp.approx.ftz.f32", out, 148 * occ);
    run<2>("tanh.approx.f32", out, 148 * occ);
    run<3>("software rcp, 2 Newton steps", out, 148 * occ);
    run<4>("software rcp, 3 Newton steps", out, 148 * occ);
  }
  return 0;
}
