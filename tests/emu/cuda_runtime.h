// TEST INFRASTRUCTURE -- host "warp emulator" stand-in for <cuda_runtime.h>.
//
// The container this repository is developed in has no GPU.  This header (found first through
// -Itests/emu) lets g++ compile the UNMODIFIED kernel sources of gwinferno_b200/csrc (kernels.cu,
// stream.cuh, api.cu) into tests/emu/libgwi_emu.so, in which every CUDA thread is a fiber: the
// threads of a block run cooperatively on one OS thread and switch at the warp / block
// synchronisation points (__syncwarp, __shfl_*_sync, __ballot_sync, __syncthreads); blocks of a grid
// run on a pool of OS threads.  It exists so that the index arithmetic, accumulator layouts,
// reductions and likelihood glue of the kernels can be checked against the oracle by the CPU test
// suite (tests/test_kernels_emulated.py) and so that experiment variants can be validated before
// they cost GPU time.  It is NOT a compute path of the product: libgwi.so contains none of it, the
// product binding refuses to load a library that exports gwi_emu_marker, and nothing outside tests/
// refers to it.  It models no timing, no memory hierarchy and no intra-warp lock-step (a data race
// between lanes that real hardware would expose is not necessarily reproduced).
#pragma once
#ifndef GWI_HOST_EMULATION
#define GWI_HOST_EMULATION 1
#endif

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <tuple>
#include <utility>

// ---- qualifiers ------------------------------------------------------------------------------
#define __global__
#define __grid_constant__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static const
// one block at a time per OS thread => block-shared storage is thread-local storage
#define __shared__ thread_local

using std::isinf;
using std::isnan;

// ---- vector types ------------------------------------------------------------------------------
struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 {
  double x, y;
};
struct alignas(16) ulonglong2 {
  unsigned long long x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

// ---- runtime API (tests/emu/emu_runtime.cpp; exported with C linkage so that the ctypes binding
//      can use the emulator library in libcudart's place) -------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct gwi_emu_stream;
struct gwi_emu_event;
typedef gwi_emu_stream* cudaStream_t;
typedef gwi_emu_event* cudaEvent_t;
struct gwi_emu_graph;
typedef gwi_emu_graph* cudaGraph_t;
typedef gwi_emu_graph* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount;
  size_t sharedMemPerBlockOptin;
};

extern "C" {
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, int kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, int kind, cudaStream_t st);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaGetLastError(void);
cudaError_t cudaFuncSetAttribute(const void* fn, int attr, int value);
const char* cudaGetErrorString(cudaError_t e);
/* stream capture: between Begin and End every launch / async copy is recorded instead of executed */
cudaError_t cudaStreamBeginCapture(cudaStream_t s, int mode);
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* graph);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* exec, cudaGraph_t graph, unsigned long long flags);
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t s);
cudaError_t cudaGraphDestroy(cudaGraph_t graph);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec);
int gwi_emu_marker(void);  // identifies the emulator build (the product binding refuses it)
}

// ---- fibers: synchronisation points ------------------------------------------------------------
namespace gwi_emu {
void warp_barrier();                              // all live lanes of the calling lane's warp
void warp_barrier_mask(unsigned mask);            // __syncwarp(mask): the live lanes named in `mask` (all of them must call it with the same mask)
void named_barrier(int id, int n_threads);        // bar.sync id, n_threads (n_threads a multiple of 32; ids 1..15)
unsigned warp_alive_mask();
void warp_yield();                                // spin-wait hint: let the other fibers of the block run
void block_barrier();                             // all live threads of the block
uint64_t warp_exchange(uint64_t mine, int from);  // publish `mine`, return lane `from`'s value (one barrier)
unsigned warp_ballot(bool pred);

struct Body {
  virtual void run() = 0;
  virtual ~Body() {}
};
void run_grid(dim3 grid, dim3 block, size_t dyn_smem, std::shared_ptr<Body> body);  // executes now, or records during stream capture

template <class F>
struct Launcher {
  dim3 grid, block;
  size_t smem;
  F fn;
  template <class... A>
  void operator()(A... args) {
    struct B : Body {
      F fn;
      std::tuple<A...> a;
      B(F f, A... x) : fn(f), a(x...) {}
      void run() override { std::apply(fn, a); }
    };
    run_grid(grid, block, smem, std::make_shared<B>(fn, args...));
  }
};
template <class F>
Launcher<F> make_launcher(dim3 g, dim3 b, size_t s, F fn) {
  return Launcher<F>{g, b, s, fn};
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  std::memcpy(&v, &b, sizeof(T));
  return v;
}
inline int lane_id() { return (int)(threadIdx.x & 31u); }
}  // namespace gwi_emu

#include <tuple>

// kernel<<<grid, block, smem, stream>>>(args...)  (see csrc/gwi_internal.h: GWI_LAUNCH)
#define GWI_EMU_LAUNCH(kernel, grid, block, smem, stream) gwi_emu::make_launcher((grid), (block), (size_t)(smem), (kernel))

// ---- device intrinsics -------------------------------------------------------------------------
inline void __syncthreads() { gwi_emu::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  if (mask == 0xffffffffu) gwi_emu::warp_barrier();
  else gwi_emu::warp_barrier_mask(mask);
}
// bar.sync id, n: kernels call this wrapper (under nvcc it is one line of inline PTX)
inline void gwi_named_barrier_sync(int id, int n_threads) { gwi_emu::named_barrier(id, n_threads); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { gwi_emu::warp_yield(); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  return gwi_emu::from_bits<T>(gwi_emu::warp_exchange(gwi_emu::to_bits(v), src & 31));
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int mask) {
  return gwi_emu::from_bits<T>(gwi_emu::warp_exchange(gwi_emu::to_bits(v), (gwi_emu::lane_id() ^ mask) & 31));
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta) {
  const int l = gwi_emu::lane_id();
  return gwi_emu::from_bits<T>(gwi_emu::warp_exchange(gwi_emu::to_bits(v), l - delta >= 0 ? l - delta : l));
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
  const int l = gwi_emu::lane_id();
  return gwi_emu::from_bits<T>(gwi_emu::warp_exchange(gwi_emu::to_bits(v), l + delta <= 31 ? l + delta : l));
}
inline unsigned __ballot_sync(unsigned, int pred) { return gwi_emu::warp_ballot(pred != 0); }
inline int __any_sync(unsigned, int pred) { return gwi_emu::warp_ballot(pred != 0) != 0u; }
inline int __all_sync(unsigned, int pred) { return gwi_emu::warp_ballot(pred == 0) == 0u; }
inline unsigned __activemask() { return gwi_emu::warp_alive_mask(); }  // fibers of a warp are never "diverged": every live lane
template <class T>
inline unsigned __match_any_sync(unsigned, T v) {
  // lanes holding the same value: 32 broadcasts (this is a test emulator, not a fast one)
  unsigned r = 0;
  const uint64_t mine = gwi_emu::to_bits(v);
  for (int l = 0; l < 32; ++l)
    if (gwi_emu::warp_exchange(mine, l) == mine && ((gwi_emu::warp_alive_mask() >> l) & 1u)) r |= 1u << l;
  return r;
}

template <class T>
inline T __ldg(const T* p) {
  return *p;
}
template <class T>
inline T __ldcg(const T* p) {
  return *p;
}

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {
  uint64_t* q = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
  for (;;) {
    const double nv = gwi_emu::from_bits<double>(old) + v;
    if (__atomic_compare_exchange_n(q, &old, gwi_emu::to_bits(nv), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return gwi_emu::from_bits<double>(old);
  }
}

inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int cmp, int v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }

inline double __hiloint2double(int hi, int lo) { return gwi_emu::from_bits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo); }
inline int __double2loint(double v) { return (int)(uint32_t)gwi_emu::to_bits(v); }
inline int __double2hiint(double v) { return (int)(uint32_t)(gwi_emu::to_bits(v) >> 32); }
inline double __longlong_as_double(long long v) { return gwi_emu::from_bits<double>((uint64_t)v); }
inline long long __double_as_longlong(double v) { return (long long)gwi_emu::to_bits(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned long long __cvta_generic_to_global(const void* p) { return (unsigned long long)(uintptr_t)p; }
