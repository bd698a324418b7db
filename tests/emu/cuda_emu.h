// CPU emulation shim for the CUDA kernels of libteal_b200 -- TEST INFRASTRUCTURE ONLY (never shipped, never timed).
//
// The kernels' .cu files are compiled as plain C++ (g++ -DTEAL_EMU -include cuda_emu.h) and executed with CUDA's
// execution model restated on the CPU:
//   * a thread is a fiber (ucontext); the fibers of one CTA run round-robin on ONE OS thread and switch only at
//     synchronisation points (__syncthreads, warp collectives, spin-wait loads), so a CTA is deterministic;
//   * __shared__ is `static thread_local`: one copy per OS thread = per running CTA;
//   * CTAs run on a pool of OS threads (all of them concurrently for persistent kernels that wait on each other),
//     global atomics / RED / acquire / release are real atomics;
//   * warp collectives exchange values through per-warp slots and complete when all 32 lanes have arrived.
// What this checks: indexing, masks, pipelines, protocols (counters, flags, tickets) -- the LOGIC of kernels that have
// not run on hardware yet.  What it cannot check: memory-model subtleties, performance, PTX semantics of the inline
// assembly (each asm helper has a hand-written C++ twin under #ifdef TEAL_EMU).
#pragma once
#ifndef TEAL_EMU
#error "cuda_emu.h is only for -DTEAL_EMU builds"
#endif

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <ucontext.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

// ---- qualifiers ------------------------------------------------------------------------------------------------
#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __grid_constant__
#undef __align__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __shared__ static thread_local
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct WarpSlots {
  unsigned gen = 0;
  int arrived = 0;
  unsigned long long in[32];
  unsigned long long out[2][32];
  int aux[32];
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  uint3 tid;
};

struct Cta {
  ucontext_t main;
  std::vector<Fiber> fibers;
  int nthreads = 0, alive = 0, cur = 0;
  unsigned bar_gen = 0;
  int bar_arrived = 0;
  WarpSlots warps[kMaxThreads / 32];
  uint3 bid;
  dim3 bdim, gdim;
  std::function<void()> body;
  void* dyn_smem = nullptr;
};

extern thread_local Cta* g_cta;

inline Fiber& self() { return g_cta->fibers[g_cta->cur]; }
inline void yield() {
  Cta* c = g_cta;
  swapcontext(&c->fibers[c->cur].ctx, &c->main);
}

// Runs `body` once per thread of a grid.  concurrent = every CTA gets its own OS thread and all of them are alive
// together (persistent kernels whose CTAs wait on each other); otherwise a small pool walks the grid.
void run_grid(dim3 grid, dim3 block, size_t dyn_smem_bytes, bool concurrent, const std::function<void()>& body);
void set_concurrent(bool on);
void set_max_workers(int n);
bool concurrent();

inline unsigned char* dyn_smem() { return static_cast<unsigned char*>(g_cta->dyn_smem); }
int sm_count();          // the emulated device: a few SMs, so that grids stay small
void set_sm_count(int n);
inline int lane_id() { return (int)(self().tid.x & 31u); }
inline WarpSlots& my_warp() { return g_cta->warps[self().tid.x >> 5]; }
inline int warp_lanes() {  // threads of this warp that exist in the CTA
  const int w = (int)(self().tid.x >> 5);
  const int left = g_cta->nthreads - w * 32;
  return left < 32 ? left : 32;
}

// every lane deposits `v`; the last arrival computes out[lane] = f(all inputs, lane) and releases the warp
template <class F>
inline unsigned long long warp_exchange(unsigned long long v, int aux, F&& combine) {
  WarpSlots& w = my_warp();
  const int lane = lane_id();
  const unsigned g = w.gen;
  w.in[lane] = v;
  w.aux[lane] = aux;
  if (++w.arrived == warp_lanes()) {
    for (int l = 0; l < 32; ++l) w.out[g & 1][l] = combine(w, l);
    w.arrived = 0;
    w.gen = g + 1;
  } else {
    while (w.gen == g) yield();
  }
  return w.out[g & 1][lane];
}

}  // namespace emu

// ---- built-in variables ----------------------------------------------------------------------------------------
// (objects, not macros: host code also has members called gridDim / blockDim)
namespace emu {
template <int WHICH, int AXIS>
struct BuiltinAxis {
  operator unsigned() const {
    const Cta* c = g_cta;
    const uint3 t = c->fibers[c->cur].tid;
    switch (WHICH) {
      case 0: return AXIS == 0 ? t.x : AXIS == 1 ? t.y : t.z;
      case 1: return AXIS == 0 ? c->bid.x : AXIS == 1 ? c->bid.y : c->bid.z;
      case 2: return AXIS == 0 ? c->bdim.x : AXIS == 1 ? c->bdim.y : c->bdim.z;
      default: return AXIS == 0 ? c->gdim.x : AXIS == 1 ? c->gdim.y : c->gdim.z;
    }
  }
};
template <int WHICH>
struct Builtin3 {
  BuiltinAxis<WHICH, 0> x;
  BuiltinAxis<WHICH, 1> y;
  BuiltinAxis<WHICH, 2> z;
};
}  // namespace emu
static const emu::Builtin3<0> threadIdx;
static const emu::Builtin3<1> blockIdx;
static const emu::Builtin3<2> blockDim;
static const emu::Builtin3<3> gridDim;
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

// ---- synchronisation -------------------------------------------------------------------------------------------
inline void __syncthreads() {
  emu::Cta* c = emu::g_cta;
  const unsigned g = c->bar_gen;
  if (++c->bar_arrived >= c->alive) {
    c->bar_arrived = 0;
    c->bar_gen = g + 1;
  } else {
    while (c->bar_gen == g) emu::yield();
  }
}
inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::warp_exchange(0, 0, [](emu::WarpSlots&, int) { return 0ull; });
}
inline unsigned __ballot_sync(unsigned, int pred) {
  return (unsigned)emu::warp_exchange(pred ? 1ull : 0ull, 0, [](emu::WarpSlots& w, int) {
    unsigned long long m = 0;
    const int n = emu::warp_lanes();
    for (int l = 0; l < n; ++l) m |= (w.in[l] & 1ull) << l;
    return m;
  });
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) {
  static_assert(sizeof(T) <= 8, "shfl of <= 8 bytes");
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  const unsigned long long r = emu::warp_exchange(bits, src & 31, [](emu::WarpSlots& w, int l) { return w.in[w.aux[l]]; });
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int = 32) {
  return __shfl_sync(m, v, emu::lane_id() ^ lane_mask);
}
template <class T>
inline T __shfl_down_sync(unsigned m, T v, unsigned d, int = 32) {
  const int src = emu::lane_id() + (int)d;
  return __shfl_sync(m, v, src < 32 ? src : emu::lane_id());
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { emu::yield(); }
inline long long clock64() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ((long long)ts.tv_sec * 1000000000ll + ts.tv_nsec) >> 6;  // slow clock: the kernels' ~2 s watchdogs become minutes
}
[[noreturn]] inline void __trap() {
  fprintf(stderr, "[emu] __trap() in block %u thread %u\n", (unsigned)blockIdx.x, (unsigned)threadIdx.x);
  abort();
}

// ---- integer / float intrinsics --------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }

// ---- memory ----------------------------------------------------------------------------------------------------
template <class T> inline void __stcg(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }
template <class T> inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <> inline uint2 __ldcg<uint2>(const uint2* p) { uint2 r; memcpy(&r, p, sizeof r); return r; }
template <> inline uint4 __ldcg<uint4>(const uint4* p) { uint4 r; memcpy(&r, p, sizeof r); return r; }
template <> inline float4 __ldcg<float4>(const float4* p) { float4 r; memcpy(&r, p, sizeof r); return r; }
template <> inline float2 __ldcg<float2>(const float2* p) { float2 r; memcpy(&r, p, sizeof r); return r; }
template <class T> inline T __ldg(const T* p) { return *p; }

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float atomicAdd(float* p, float v) {
  unsigned* u = reinterpret_cast<unsigned*>(p);
  unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED);
  for (;;) {
    float f;
    memcpy(&f, &old, 4);
    f += v;
    unsigned nu;
    memcpy(&nu, &f, 4);
    if (__atomic_compare_exchange_n(u, &old, nu, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
      memcpy(&f, &old, 4);
      return f;
    }
  }
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
  }
  return old;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
  }
  return old;
}
inline unsigned atomicSub(unsigned* p, unsigned v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED);
  return cmp;
}

// ---- launches --------------------------------------------------------------------------------------------------
template <class... KArgs, class... Args>
inline cudaError_t emu_launch_kernel_ex(const cudaLaunchConfig_t* cfg, void (*kern)(KArgs...), Args&&... args) {
  emu::run_grid(cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, emu::concurrent(), [&]() { kern(args...); });
  return cudaSuccess;
}
#define cudaLaunchKernelEx emu_launch_kernel_ex
template <class K>
inline cudaError_t emu_occupancy(int* out, K, int block, size_t) {
  *out = block <= 64 ? 4 : 2;  // a small "device": a handful of resident CTAs per SM
  return cudaSuccess;
}
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor emu_occupancy
// ---- the few runtime calls the host code makes ----------------------------------------------------------------------
inline cudaError_t emu_get_device(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t emu_device_get_attribute(int* v, cudaDeviceAttr a, int) {
  *v = (a == cudaDevAttrMultiProcessorCount) ? emu::sm_count() : 227 * 1024;
  return cudaSuccess;
}
inline cudaError_t emu_memset_async(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t emu_get_last_error() { return cudaSuccess; }
inline const char* emu_get_error_string(cudaError_t) { return "emulated CUDA error"; }
#define cudaGetDevice emu_get_device
#define cudaDeviceGetAttribute emu_device_get_attribute
#define cudaMemsetAsync emu_memset_async
#define cudaGetLastError emu_get_last_error
#define cudaGetErrorString emu_get_error_string
template <class K>
inline cudaError_t emu_func_set_attribute(K, int, int) { return cudaSuccess; }
#define cudaFuncSetAttribute emu_func_set_attribute
