// TEST INFRASTRUCTURE (not part of the product, never loaded by jaxsso_b200): a functional SIMT emulator that
// lets g++ compile the product's CUDA kernel headers (csrc/*.cuh, with -DJSSO_EMU) and run them on the CPU, so
// that kernel LOGIC (indexing, warp shuffles, shared-memory staging, fixed-order reductions) can be checked
// against the oracle in the CPU test suite, where there is no GPU.  It says nothing about performance and is not a
// fallback: the product library (libjsso.so) is built by nvcc only and refuses to run without a device.
//
// Model: CTAs run one after another; inside a CTA every CUDA thread is a std::thread.  __syncthreads /
// __syncwarp are std::barrier phases; a warp shuffle is "publish my value, barrier, read the source lane,
// barrier".  A thread that returns from the kernel drops out of its barriers (arrive_and_drop), like an exited
// lane.  Limits: full-mask warp primitives only; no inter-CTA waiting (grid.sync only with one CTA; the
// peer-memory spin loops are not emulated); cp.async is a synchronous copy, so a missing wait is NOT detected.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

namespace emu {

struct WarpCtx {
  std::barrier<> bar;
  alignas(64) unsigned char slot[32][16];
  explicit WarpCtx(int n) : bar(n) {}
};
struct CtaCtx {
  std::barrier<> bar;
  std::vector<unsigned char> dyn;
  CtaCtx(int n, size_t smem) : bar(n), dyn(smem + 64) {}
};
struct Tls {
  uint3 tid{0, 0, 0}, bid{0, 0, 0};
  dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
  int lane = 0;
  WarpCtx* w = nullptr;
  CtaCtx* c = nullptr;
};
inline thread_local Tls tls;

inline void* dyn_smem() {
  uintptr_t p = (uintptr_t)tls.c->dyn.data();
  return (void*)((p + 63) & ~(uintptr_t)63);
}

// launch(grid, block, dynamic shared bytes, [&] { kernel(args...); })
template <class F>
void launch(unsigned grid, unsigned block, size_t smem, F&& body) {
  const unsigned n_warp = (block + 31) / 32;
  for (unsigned b = 0; b < grid; ++b) {
    CtaCtx cta((int)block, smem);
    std::vector<std::unique_ptr<WarpCtx>> warps;
    for (unsigned w = 0; w < n_warp; ++w)
      warps.emplace_back(new WarpCtx((int)std::min(32u, block - 32 * w)));
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
      th.emplace_back([&, t, b] {
        tls.tid = uint3{t, 0, 0};
        tls.bid = uint3{b, 0, 0};
        tls.bdim = dim3(block, 1, 1);
        tls.gdim = dim3(grid, 1, 1);
        tls.lane = (int)(t & 31);
        tls.w = warps[t >> 5].get();
        tls.c = &cta;
        body();
        tls.w->bar.arrive_and_drop();
        tls.c->bar.arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
  }
}

template <class T>
inline T shfl_from(T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  WarpCtx& w = *tls.w;
  std::memcpy(w.slot[tls.lane], &v, sizeof(T));
  w.bar.arrive_and_wait();
  T r;
  std::memcpy(&r, w.slot[src], sizeof(T));
  w.bar.arrive_and_wait();
  return r;
}

}  // namespace emu

#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.bid)
#define blockDim (emu::tls.bdim)
#define gridDim (emu::tls.gdim)

inline void __syncthreads() { emu::tls.c->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::tls.w->bar.arrive_and_wait(); }

template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = emu::tls.lane, base = lane & ~(width - 1);
  return emu::shfl_from(v, base + (src & (width - 1)));
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = emu::tls.lane, base = lane & ~(width - 1), s = lane + (int)d;
  return emu::shfl_from(v, s < base + width ? s : lane);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = emu::tls.lane, base = lane & ~(width - 1), s = lane - (int)d;
  return emu::shfl_from(v, s >= base ? s : lane);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  const int lane = emu::tls.lane, base = lane & ~(width - 1), s = lane ^ m;
  return emu::shfl_from(v, s < base + width ? s : lane);
}

template <class T>
inline T __ldg(const T* p) { return *p; }

inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicInc(unsigned* p, unsigned lim) {   // old >= lim ? 0 : old + 1
  unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  for (;;) {
    const unsigned nv = (old >= lim) ? 0u : old + 1u;
    if (__atomic_compare_exchange_n(p, &old, nv, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return old;
  }
}

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::sqrt;
template <class A, class B>
inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <class A, class B>
inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }

// cooperative groups: only what cg_persistent_kernel uses, and only for a single CTA
#define _COOPERATIVE_GROUPS_H_
namespace cooperative_groups {
struct grid_group {
  void sync() const {
    if (emu::tls.gdim.x != 1) { std::fprintf(stderr, "emu: grid.sync() needs a one-CTA grid\n"); std::abort(); }
    __syncthreads();
  }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
