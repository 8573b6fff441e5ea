// simt_emu.h -- TEST INFRASTRUCTURE ONLY: a single-threaded SIMT emulator that lets g++ compile and run the
// repo's CUDA sources (ev2gym_b200/csrc/ev2b.cu + *.cuh, built with -DEV2B_SIMT_EMU) on the CPU, so that kernel
// logic (indexing, barrier placement, shared-memory protocols, list maintenance) can be checked against the
// oracle in the `-m "not gpu"` suite.  It is NOT a CPU fallback: nothing under ev2gym_b200/ loads the library
// built from it, and it is far too slow to be one (every CUDA thread is a ucontext fiber).
//
// What is emulated: a grid of CTAs run one after the other; the threads of a CTA are fibers scheduled round-robin
// (or in a seeded random order, SIMT_EMU_SEED) and switched only at __syncthreads / named barriers / __syncwarp /
// warp shuffles and votes, which block until every participating thread has arrived.  cp.async copies are DEFERRED
// until cp_async_wait_all, so a missing wait shows up as stale data.  cudaMalloc returns memory that ends flush
// against a PROT_NONE guard page (overruns fault), dynamic shared memory likewise.  A barrier that can never
// complete (a participating thread exited or waits elsewhere) aborts with a message instead of hanging.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <functional>

// ---- qualifiers ------------------------------------------------------------------------------------------------
#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define EV2B_NOINLINE __attribute__((noinline))     // (libstdc++ itself spells __attribute__((__noinline__)): do not touch __noinline__)
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) alignas(n)
#define __shared__ static
#define __restrict__

// ---- vector types ----------------------------------------------------------------------------------------------
struct alignas(8)  uint2   { unsigned x, y; };
struct alignas(8)  int2    { int x, y; };
struct alignas(16) uint4   { unsigned x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8)  float2  { float x, y; };
static inline uint2   make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4   make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float2  make_float2(float x, float y) { return float2{x, y}; }

namespace simt {
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body);
unsigned char *dyn_smem();
void syncthreads();
void bar_sync(int id, int nthreads);       // PTX bar.sync id, nthreads
void bar_arrive(int id, int nthreads);     // PTX bar.arrive id, nthreads
void syncwarp(unsigned mask);
uint64_t shfl(unsigned mask, uint64_t bits, int src_lane);
unsigned ballot(unsigned mask, int pred);
void cp_async(void *dst, const void *src, size_t n);
void cp_async_wait_all();
void *dev_alloc(size_t bytes);
void dev_free(void *p);
long kernel_launches();
}  // namespace simt

extern simt::Dim3 threadIdx, blockIdx, blockDim, gridDim;
#define EV2B_DYNAMIC_SMEM(name) unsigned char *name = simt::dyn_smem()

// ---- intrinsics --------------------------------------------------------------------------------------------------
static inline void __syncthreads() { simt::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::syncwarp(mask); }
template <typename T> static inline T __simt_shfl(unsigned mask, T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
    uint64_t b = 0; memcpy(&b, &v, sizeof(T));
    b = simt::shfl(mask, b, src);
    T r; memcpy(&r, &b, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int o) { return __simt_shfl(mask, v, (int)(threadIdx.x & 31u) ^ o); }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src) { return __simt_shfl(mask, v, src & 31); }
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) {
    const int lane = (int)(threadIdx.x & 31u); return __simt_shfl(mask, v, lane >= (int)d ? lane - (int)d : lane);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d) {
    const int lane = (int)(threadIdx.x & 31u); return __simt_shfl(mask, v, lane + (int)d < 32 ? lane + (int)d : lane);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return simt::ballot(mask, pred); }
// The emulator runs one fiber at a time: the "converged" set a thread can rely on is itself (opportunistic warp-level
// decisions -- "does ANY lane need this block?" -- stay correct with that answer: each lane asks for what it needs).
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }
static inline int __any_sync(unsigned mask, int pred) { return simt::ballot(mask, pred) != 0; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }          // one OS thread: trivially atomic
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline int atomicOr(int *p, int v) { const int o = *p; *p = o | v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
using std::min; using std::max;

// ---- the slice of the CUDA runtime API the host side uses ----------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamNonBlocking = 1,
       cudaEventDisableTiming = 2 };
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)simt::dev_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { simt::dev_free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "simt_emu error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 148; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

#define EV2B_LAUNCH(kern, grid, block, smem, stream, ...) \
    simt::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
