// tests/emu/cuda_runtime.h — stand-in for the CUDA runtime used ONLY by the host emulation build
// (tests/emu/Makefile -> libuba_emu.so).  It lets the thread-independent kernels and the host
// driver of libuba run serially on a CPU so that host logic (index tables, launch sequencing, the
// LM controller) can be debugged on a box without a GPU.  It is test infrastructure: the package
// never loads libuba_emu.so and libuba.so is never compiled against this header.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) double2 { double x, y; };
template <typename T> inline T __ldcs(const T* p) { return *p; }
namespace uba_emu { extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim; }
#define threadIdx uba_emu::t_threadIdx
#define blockIdx uba_emu::t_blockIdx
#define blockDim uba_emu::t_blockDim
#define gridDim uba_emu::t_gridDim

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotReady = 600 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp { int major, minor; };

inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n ? n : 1); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
constexpr int cudaEventDisableTiming = 2;
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, int, int) { *s = nullptr; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, int) { *e = nullptr; return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
inline int atomicSub(int* p, int v) { const int o = *p; *p = o - v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline int __shfl_sync(unsigned, int v, int) { return v; }
inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
inline bool __all_sync(unsigned, bool p) { return p; }
using std::isfinite;
using std::min;
using std::max;
