// TEST INFRASTRUCTURE (not part of the product, never loaded by jaxsso_b200): a functional SIMT emulator that
// lets g++ compile the product's CUDA kernel headers (csrc/*.cuh, with -DJSSO_EMU) and run them on the CPU, so
// that kernel LOGIC (indexing, warp shuffles, shared-memory staging, fixed-order reductions) can be checked
// against the oracle in the CPU test suite, where there is no GPU.  It says nothing about performance and is not a
// fallback: the product library (libjsso.so) is built by nvcc only and refuses to run without a device.
//
// Model: CTAs run one after another; inside a CTA every CUDA thread is a cooperative fiber.
// __syncthreads / __syncwarp are barrier phases; a warp shuffle is "publish my value, barrier, read the source
// lane, barrier".  A thread that returns from the kernel drops out of its barriers, like an exited lane.  Limits: full-mask warp primitives only; no inter-CTA waiting (grid.sync only with one CTA; the
// peer-memory spin loops are not emulated); cp.async is a synchronous copy, so a missing wait is NOT detected.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
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
#define __shared__ static thread_local   /* one copy per rank thread: CTAs of different ranks run concurrently */
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// stack switch without a system call (swapcontext saves the signal mask with one): callee-saved registers only
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define EMU_ASAN 1
#else
#define EMU_ASAN 0
#endif

namespace emu {

// ---- cooperative fibers: every CUDA thread of the running CTA is a fiber (own stack, hand-written switch) on the launching OS thread.
// A barrier is "arrive; if not last, yield until the generation changes"; the scheduler resumes fibers round-robin,
// warp by warp, so a warp-level barrier completes after one pass over its 32 lanes.  Deterministic, no OS threads,
// no futexes (an OS-thread version spent milliseconds per CTA waking 256 threads).
struct Barrier {
  int expected = 0, arrived = 0;
  unsigned long long gen = 0;
};
struct Fiber;
struct WarpCtx {
  Barrier bar;
  alignas(64) unsigned char slot[32][16];
};
struct CtaCtx {
  Barrier bar;
  unsigned char* dyn = nullptr;   // dynamic shared memory of the running CTA (buffer owned by the scheduler)
};
struct Fiber {
  void* sp = nullptr;
  void* stack_lo = nullptr;   // AddressSanitizer needs the bounds of the stack it switches to
  uint3 tid{0, 0, 0}, bid{0, 0, 0};
  dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
  int lane = 0;
  WarpCtx* w = nullptr;
  CtaCtx* c = nullptr;
  bool done = false;
  const std::function<void()>* body = nullptr;
};
struct Sched {
  void* main_sp = nullptr;
  const void* main_lo = nullptr; size_t main_size = 0;
  std::vector<Fiber> fibers;
  std::vector<unsigned char*> stacks;
  std::vector<unsigned char> dyn;   // reused by every CTA; refilled with 0xFF (NaN doubles) so that a read of
                                    // shared memory nobody wrote shows up in the results
  Fiber* cur = nullptr;
};
inline Sched& sched() { static thread_local Sched s; return s; }
constexpr size_t kStack = 512 * 1024;

// the running fiber of this OS thread
#define EMU_CUR (*emu::sched().cur)

inline void yield() {
  Sched& S = sched();
#if EMU_ASAN
  void* fake = nullptr;
  __sanitizer_start_switch_fiber(&fake, S.main_lo, S.main_size);
#endif
  emu_switch(&S.cur->sp, S.main_sp);
#if EMU_ASAN
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
}
inline void barrier_wait(Barrier& b) {
  if (++b.arrived >= b.expected) { b.arrived = 0; ++b.gen; return; }
  const unsigned long long g = b.gen;
  while (b.gen == g) yield();
}
inline void barrier_drop(Barrier& b) {
  --b.expected;
  if (b.expected > 0 && b.arrived >= b.expected) { b.arrived = 0; ++b.gen; }
}
inline void fiber_entry() {
#if EMU_ASAN
  __sanitizer_finish_switch_fiber(nullptr, &sched().main_lo, &sched().main_size);
#endif
  Fiber& f = *sched().cur;
  (*f.body)();
  barrier_drop(f.w->bar);
  barrier_drop(f.c->bar);
  f.done = true;
#if EMU_ASAN
  __sanitizer_start_switch_fiber(nullptr, sched().main_lo, sched().main_size);   // this fiber never comes back
#endif
  emu_switch(&f.sp, sched().main_sp);
  std::abort();   // a finished fiber is never resumed
}

inline void* dyn_smem() {
  uintptr_t p = (uintptr_t)EMU_CUR.c->dyn;
  return (void*)((p + 63) & ~(uintptr_t)63);
}

inline unsigned dimx(unsigned v) { return v; }
inline unsigned dimx(int v) { return (unsigned)v; }
inline unsigned dimx(long long v) { return (unsigned)v; }
inline unsigned dimx(size_t v) { return (unsigned)v; }
inline unsigned dimx(const dim3& v) { return v.x; }
// Launches of different OS threads (rank threads) run CONCURRENTLY -- the peer-memory kernels spin on flags that
// another rank's kernel sets -- so everything a launch touches is per thread: the scheduler, the fiber stacks and
// the `__shared__` statics (thread_local).

// stream capture (CUDA graphs): while a thread captures, launches are recorded (closures hold their arguments BY
// VALUE, like a real launch) instead of executed; replay runs them in order
struct GraphNode { unsigned grid, block; size_t smem; std::function<void()> body; };
struct Graph { std::vector<GraphNode> nodes; };
inline thread_local Graph* capturing = nullptr;

// launch(grid, block, dynamic shared bytes, [=] { kernel(args...); })
template <class G, class B, class F>
void launch(G grid_, B block_, size_t smem, F&& body_) {
  if (capturing) {
    capturing->nodes.push_back(GraphNode{dimx(grid_), dimx(block_), smem, std::function<void()>(body_)});
    return;
  }
  {
    // EMU_JITTER=<max microseconds>: a random pause before every launch, different per rank thread, so that the
    // peer-memory protocols (flags, double-buffered arenas) are exercised with ranks running far ahead of each other
    static const int jitter = [] { const char* e = std::getenv("EMU_JITTER"); return e ? std::atoi(e) : 0; }();
    if (jitter > 0) {
      static thread_local unsigned long long rs = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(uintptr_t)&rs;
      rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
      if ((rs & 7) == 0) std::this_thread::sleep_for(std::chrono::microseconds(rs % (unsigned)jitter));
    }
  }
  const unsigned grid = dimx(grid_), block = dimx(block_);
  const unsigned n_warp = (block + 31) / 32;
  const std::function<void()> body = body_;
  Sched& S = sched();
  if (S.fibers.size() < block) S.fibers.resize(block);
  while (S.stacks.size() < block) S.stacks.push_back((unsigned char*)std::malloc(kStack));
  for (unsigned b = 0; b < grid; ++b) {
    CtaCtx cta;
    cta.bar.expected = (int)block;
    if (S.dyn.size() < smem + 64) S.dyn.resize(smem + 64);
    std::memset(S.dyn.data(), 0xFF, smem + 64);
    cta.dyn = S.dyn.data();
    std::vector<WarpCtx> warps(n_warp);
    for (unsigned w = 0; w < n_warp; ++w) warps[w].bar.expected = (int)std::min(32u, block - 32 * w);
    for (unsigned t = 0; t < block; ++t) {
      Fiber& f = S.fibers[t];
      f.tid = uint3{t, 0, 0}; f.bid = uint3{b, 0, 0};
      f.bdim = dim3(block, 1, 1); f.gdim = dim3(grid, 1, 1);
      f.lane = (int)(t & 31); f.w = &warps[t >> 5]; f.c = &cta; f.done = false; f.body = &body;
      // initial frame: six zeroed callee-saved registers, then the entry address for `ret`; after the ret
      // rsp = top - 8, the alignment a function sees on entry
      void** top = (void**)(((uintptr_t)S.stacks[t] + kStack) & ~(uintptr_t)15);
      void** sp0 = top - 8;
      for (int k = 0; k < 6; ++k) sp0[k] = nullptr;
      sp0[6] = (void*)&fiber_entry;
      sp0[7] = nullptr;
      f.sp = sp0;
      f.stack_lo = S.stacks[t];
    }
    unsigned live = block;
    while (live > 0) {
      unsigned progressed = 0;
      for (unsigned t = 0; t < block; ++t) {
        Fiber& f = S.fibers[t];
        if (f.done) continue;
        S.cur = &f;
#if EMU_ASAN
        void* fake = nullptr;
        __sanitizer_start_switch_fiber(&fake, f.stack_lo, kStack);
#endif
        emu_switch(&S.main_sp, f.sp);
#if EMU_ASAN
        __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
        if (f.done) { --live; }
        ++progressed;
      }
      if (!progressed) break;
    }
    S.cur = nullptr;
  }
}

template <class T>
inline T shfl_from(T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  Fiber& f = EMU_CUR;
  WarpCtx& w = *f.w;
  std::memcpy(w.slot[f.lane], &v, sizeof(T));
  barrier_wait(w.bar);
  T r;
  std::memcpy(&r, w.slot[src], sizeof(T));
  barrier_wait(w.bar);
  return r;
}

}  // namespace emu

#define threadIdx (EMU_CUR.tid)
#define blockIdx (EMU_CUR.bid)
#define blockDim (EMU_CUR.bdim)
#define gridDim (EMU_CUR.gdim)

inline void __syncthreads() { emu::barrier_wait(EMU_CUR.c->bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier_wait(EMU_CUR.w->bar); }

template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = EMU_CUR.lane, base = lane & ~(width - 1);
  return emu::shfl_from(v, base + (src & (width - 1)));
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = EMU_CUR.lane, base = lane & ~(width - 1), s = lane + (int)d;
  return emu::shfl_from(v, s < base + width ? s : lane);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = EMU_CUR.lane, base = lane & ~(width - 1), s = lane - (int)d;
  return emu::shfl_from(v, s >= base ? s : lane);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  const int lane = EMU_CUR.lane, base = lane & ~(width - 1), s = lane ^ m;
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
    if (EMU_CUR.gdim.x != 1) { std::fprintf(stderr, "emu: grid.sync() needs a one-CTA grid\n"); std::abort(); }
    __syncthreads();
  }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
