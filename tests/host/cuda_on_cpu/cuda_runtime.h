// TEST INFRASTRUCTURE ONLY -- a stand-in for <cuda_runtime.h> that lets g++ compile the column-physics sources of
// isca_b200/csrc/physics*.cu for the HOST (tests/host/build_phys_cpu.py rewrites their `kernel<<<grid, block, smem, stream>>>(args)`
// launches into ISCA_CPU_LAUNCH(kernel, grid, block, args)).  "Device" memory is host memory, a stream is a null pointer, a kernel
// launch is an OpenMP loop over the blocks with the threads of a block run one after the other (none of these kernels synchronises
// inside a block).  The result, tests/host/_build/libisca_phys_cpu.so, exports the same C ABI as the product for the column
// physics; it is loaded only by tests (`pytest -m "not gpu"` runs the physics parity tests against it) and by bench.py's CPU
// baseline leg -- never by the isca_b200 package, whose library fails at create time without a CUDA device.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#define __global__
#define __device__
#define __host__
#define __constant__ static
#define __forceinline__ inline
#define __launch_bounds__(...)

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
typedef double* cudaEvent_t;

template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(std::malloc(n ? n : 1)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host build"; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline double isca_cpu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new double(0.0); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { *e = isca_cpu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*b - *a); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }

struct IscaCpuDim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local IscaCpuDim3 threadIdx, blockIdx, blockDim, gridDim;

template <class T> inline T __ldg(const T* p) { return *p; }
inline int atomicExch(int* a, int v) { int o; 
#pragma omp atomic capture
  { o = *a; *a = v; }
  return o; }
inline int atomicOr(int* a, int v) { int o;
#pragma omp atomic capture
  { o = *a; *a |= v; }
  return o; }

namespace isca_cpu {
extern double kernel_ms;                     // wall-clock time spent inside launches (bench.py's CPU baseline reads it: kernels only,
                                             // without the host-array staging copies of the C ABI entry points)
template <class F> inline void launch(long grid, long block, F&& body) {
  const double t0 = isca_cpu_now_ms();
#pragma omp parallel for schedule(dynamic, 2)
  for (long b = 0; b < grid; ++b) {
    gridDim.x = (unsigned)grid; blockDim.x = (unsigned)block; blockIdx.x = (unsigned)b;
    for (long t = 0; t < block; ++t) { threadIdx.x = (unsigned)t; body(); }
  }
  kernel_ms += isca_cpu_now_ms() - t0;
}
}  // namespace isca_cpu
#define ISCA_CPU_LAUNCH(kernel, grid, block, ...) isca_cpu::launch((long)(grid), (long)(block), [&]() { kernel(__VA_ARGS__); })
