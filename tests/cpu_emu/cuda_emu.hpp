// TEST INFRASTRUCTURE — never part of the product.  A SIMT emulator that lets the library's own CUDA sources
// (csrc/*.cu, rewritten only at their <<<...>>> launch sites by preprocess.py) be compiled by g++ and run on host
// cores, so the kernels' control flow, indexing, shared-memory rings and barriers can be exercised by the ordinary
// parity tests in a container without a GPU.  liblbm_b200.so never links or loads any of this; the emulated build
// is a separate file (tests/cpu_emu/_build/liblbm_b200_emu.so) that only tests/test_emu.py loads, by explicit path.
//
// Model: one kernel launch = a loop over blocks (OpenMP across host threads); the threads of a block are fibers
// on one host thread; __syncthreads() and the warp shuffles are typed waits the block's scheduler resolves.
// `__shared__` becomes `static thread_local` (one block at a time per host thread).  The CUDA runtime calls the
// library makes (cudaMalloc, cudaMemcpyAsync, streams, events) are implemented in cuda_emu.cpp as an in-order,
// immediately-executing device whose fresh allocations are filled with NaNs.
#pragma once
#include <cuda_runtime.h>  // types and prototypes only; the definitions live in cuda_emu.cpp, libcudart is not linked

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <tuple>
#include <utility>

#undef __shared__
#define __shared__ static thread_local
#undef __constant__
#define __constant__
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace emu
{
struct LaunchCfg
{
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  LaunchCfg(dim3 g, dim3 b, size_t s = 0, cudaStream_t st = nullptr) : grid(g), block(b), smem(s), stream(st) {}
};
void run_grid(const LaunchCfg& cfg, const std::function<void()>& thread_body);
void* dyn_smem();
void sync_threads();
void spin_yield();  // polling loops (mbarrier waits) hand the host thread to the block's other fibers
uint64_t warp_exchange(uint64_t mine, int src_lane_delta, bool up, int width);

template <class... P, class... A>
inline void launch(const LaunchCfg& cfg, void (*kernel)(P...), A&&... args)
{
  std::tuple<std::decay_t<P>...> held(static_cast<std::decay_t<P>>(std::forward<A>(args))...);
  run_grid(cfg, [&held, kernel]() { std::apply(kernel, held); });
}
}  // namespace emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
constexpr int warpSize = 32;

inline void __syncthreads() { emu::sync_threads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::warp_exchange(0, 0, false, 32); }  // a real barrier of the warp's fibers (all lanes must call it)

template <class T>
inline T __ldg(const T* p) { return *p; }

template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
  static_assert(sizeof(T) <= 8, "emulated shuffles move at most 8 bytes");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  raw = emu::warp_exchange(raw, (int)delta, false, width);
  std::memcpy(&v, &raw, sizeof(T));
  return v;
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
  static_assert(sizeof(T) <= 8, "emulated shuffles move at most 8 bytes");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  raw = emu::warp_exchange(raw, (int)delta, true, width);
  std::memcpy(&v, &raw, sizeof(T));
  return v;
}

inline double atomicAdd(double* a, double v)
{
  double old = *reinterpret_cast<volatile double*>(a), want;
  uint64_t o, w;
  do
  {
    want = old + v;
    std::memcpy(&o, &old, 8);
    std::memcpy(&w, &want, 8);
    if (__atomic_compare_exchange_n(reinterpret_cast<uint64_t*>(a), &o, w, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return old;
    std::memcpy(&old, &o, 8);
  } while (true);
}
inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v)
{
  unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline long long __double_as_longlong(double x) { long long r; std::memcpy(&r, &x, 8); return r; }
inline double __longlong_as_double(long long x) { double r; std::memcpy(&r, &x, 8); return r; }
inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }

// CUDA's global min / max overload set (mixed integer widths are common in the kernels)
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }
inline long long min(int a, long long b) { return a < b ? a : b; }
inline long long max(int a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline double min(double a, double b) { return std::fmin(a, b); }
inline double max(double a, double b) { return std::fmax(a, b); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __drcp_rn(double x) { return 1.0 / x; }
inline double __dsqrt_rn(double x) { return std::sqrt(x); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __double2int_rn(double x) { return (int)std::nearbyint(x); }
inline int __double2int_rd(double x) { return (int)std::floor(x); }
// un-fused arithmetic (the kernels use these where the reference's summation order must not be contracted)
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }

template <class... P>
inline cudaError_t cudaFuncSetAttribute(void (*)(P...), cudaFuncAttribute, int) { return cudaSuccess; }
