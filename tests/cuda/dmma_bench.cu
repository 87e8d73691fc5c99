// tests/cuda/dmma_bench.cu — throughput of mma.sync.aligned.m8n8k4.f64 vs DFMA on this GPU (probe, not part of libuba).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int TILES>
__global__ void k_dmma(double* out, int iters) {
  double c[TILES][2];
  for (int i = 0; i < TILES; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < TILES; i++) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < TILES; i++) s += c[i][0] + c[i][1];
  if (s == 1.2345) out[0] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, int iters) {
  double c[ILP];
  for (int i = 0; i < ILP; i++) c[i] = threadIdx.x + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i];
  if (s == 1.2345) out[0] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 64);
  const int it = 100000;
  for (int warps : {1, 2, 4, 8, 16}) {
    float ms = timeit([&] { k_dmma<10><<<148, warps * 32>>>(out, it); });
    printf("DMMA m8n8k4 x10 tiles, %2d warps/SM: %.2f TFLOP/s\n", warps, 2.0 * 256 * 10 * it * warps * 148 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma<2><<<148, warps * 32>>>(out, it); });
    printf("DMMA m8n8k4 x2 tiles,  %2d warps/SM: %.2f TFLOP/s\n", warps, 2.0 * 256 * 2 * it * warps * 148 / (ms * 1e-3) / 1e12);
  }
  for (int warps : {4, 8, 16, 32}) {
    float ms = timeit([&] { k_dfma<8><<<148, warps * 32>>>(out, it); });
    printf("DFMA ilp 8, %2d warps/SM: %.2f TFLOP/s\n", warps, 2.0 * 8 * 32 * it * warps * 148 / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dfma<2><<<148, warps * 32>>>(out, it); });
    printf("DFMA ilp 2, %2d warps/SM: %.2f TFLOP/s\n", warps, 2.0 * 2 * 32 * it * warps * 148 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
