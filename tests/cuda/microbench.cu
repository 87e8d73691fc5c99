// tests/cuda/microbench.cu — latency probes used to size the kernels (not part of libuba).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

template <int OP>
__global__ void chain(double* out, long long* cyc, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-9, b = 1.0000001, c = 1e-9;
  __shared__ double sm[64];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 64);
  __syncthreads();
  long long t0 = clock64();
  int idx = 0;
  for (int i = 0; i < iters; i++) {
    if (OP == 0) a = fma(a, b, c);
    else if (OP == 1) a = a * b;
    else if (OP == 2) a = a + c;
    else if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); a = r + 1.5; }
    else if (OP == 4) a = fast_rcp(a) + 1.5;
    else if (OP == 5) a = 1.0 / a + 1.5;
    else if (OP == 6) a = rsqrt(a) + 1.5;
    else if (OP == 7) a = sqrt(a) + 1.5;
    else if (OP == 8) { idx = (int)sm[idx]; }
    else if (OP == 9) { __syncthreads(); }
    else if (OP == 10) { sm[threadIdx.x & 63] = a; __syncthreads(); a = sm[(threadIdx.x + 1) & 63] + 1.0; __syncthreads(); }
    else if (OP == 11) { float f = (float)a; f = fmaf(f, 1.0001f, 1e-3f); a = f; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = a + idx; }
}

__global__ void tput(double* out, int iters, int ilp) {
  double a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-3 + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < ilp) a[k] = fma(a[k], b, c);
  }
  double s = 0; for (int k = 0; k < 8; k++) s += a[k];
  if (s == 1234.5) out[0] = s;
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  const char* names[] = {"DFMA dep", "DMUL dep", "DADD dep", "MUFU.RCP64H+DADD", "fast_rcp+DADD", "IEEE div+DADD", "rsqrt+DADD", "sqrt+DADD",
                         "LDS.64 dep (+cvt)", "__syncthreads", "STS+bar+LDS+bar", "cvt f64->f32->f64 + FFMA"};
  const int iters = 20000;
  for (int threads : {32, 256}) {
    printf("-- block of %d threads, cycles per iteration\n", threads);
    for (int op = 0; op < 12; op++) {
      long long h = 0;
      switch (op) {
        case 0: chain<0><<<1, threads>>>(out, cyc, iters, 1.0); break; case 1: chain<1><<<1, threads>>>(out, cyc, iters, 1.0); break;
        case 2: chain<2><<<1, threads>>>(out, cyc, iters, 1.0); break; case 3: chain<3><<<1, threads>>>(out, cyc, iters, 1.3); break;
        case 4: chain<4><<<1, threads>>>(out, cyc, iters, 1.3); break; case 5: chain<5><<<1, threads>>>(out, cyc, iters, 1.3); break;
        case 6: chain<6><<<1, threads>>>(out, cyc, iters, 1.3); break; case 7: chain<7><<<1, threads>>>(out, cyc, iters, 1.3); break;
        case 8: chain<8><<<1, threads>>>(out, cyc, iters, 1.3); break; case 9: chain<9><<<1, threads>>>(out, cyc, iters, 1.3); break;
        case 10: chain<10><<<1, threads>>>(out, cyc, iters, 1.3); break; case 11: chain<11><<<1, threads>>>(out, cyc, iters, 1.3); break;
      }
      cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("   %-28s %8.1f\n", names[op], (double)h / iters);
    }
  }
  // DFMA throughput per SM: warps x ilp
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) for (int ilp : {1, 2, 4, 8}) {
    const int it = 200000;
    tput<<<148, warps * 32>>>(out, 1000, ilp);
    cudaEventRecord(e0); tput<<<148, warps * 32>>>(out, it, ilp); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("   DFMA tput: %2d warps/SM ilp %d -> %.2f TFLOP/s\n", warps, ilp, 2.0 * it * ilp * warps * 32 * 148 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
