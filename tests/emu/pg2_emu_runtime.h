// tests/emu/pg2_emu_runtime.h -- TEST INFRASTRUCTURE.
// A few dozen lines that let the CUDA sources under pagan2_msa_b200/csrc compile with plain g++ so that
// the engine's host logic (packing, grouping, band geometry, unpacking) and the per-thread kernel bodies
// can be exercised by the CPU test-suite in a container without a GPU.  Kernels become serial loops over
// (block, thread); there is no concurrency and no performance meaning.  Built only into
// tests/_emu/libpg2_emu.so by tests/emu/Makefile; the product library is built by nvcc and never sees
// this header (PG2_HOST_EMU is not defined there).
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <chrono>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
static inline double4 make_double4(double x, double y, double z, double w) { double4 r = {x, y, z, w}; return r; }
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
struct int4 { int x, y, z, w; };
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }

typedef int cudaError_t;
typedef int cudaStream_t;
struct pg2_emu_event { std::chrono::steady_clock::time_point t; };
typedef pg2_emu_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; size_t totalGlobalMem; };

static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p); p->major = 10; p->minor = 0; p->multiProcessorCount = 4; strcpy(p->name, "host-emulation");
    p->totalGlobalMem = (size_t)8 << 30; return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new pg2_emu_event(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
