// CPU SIMT emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// Purpose: let the `-m "not gpu"` suite execute the *source code* of the product's CUDA kernels
// (femcy_b200/csrc/*_kernels.cuh) on the CPU, so that indexing, barrier structure, reduction order and
// the peer-window protocols are checked against the oracle in a container without a GPU.  Nothing under
// femcy_b200/ includes, links or loads this; the product library has no CPU path (see femcy_b200/_lib.py).
//
// Model
//   * one OS thread runs one thread block at a time; every CUDA thread of the block is a fiber with its own stack
//     (x86-64: a syscall-free register/stack switch; elsewhere and under ASan: ucontext), scheduled round-robin
//     (or in random order, SIMT_SHUFFLE) and switched only at synchronisation points;
//   * __syncthreads / __syncwarp / __shfl_*_sync / cooperative grid.sync are real barriers over the live
//     fibers of the block / warp / grid (exited threads count as arrived, as on Volta+);
//   * `__shared__` is `static thread_local`: one copy per OS thread == per running block;
//   * plain launches distribute the blocks over a small pool of OS threads; cooperative launches run ALL
//     blocks concurrently (one OS thread per block) with a pthread barrier behind grid.sync;
//   * several launches may run concurrently from different host threads ("ranks" of the multi-GPU tests):
//     peer memory is just host memory;
//   * global atomics are real atomics; spin loops call simt::yield() (FEMCY_SPIN_PAUSE) so that the other
//     fibers of the block make progress.
// Not modelled: memory-model weakness (x86 is stronger than the GPU), bank conflicts, timing.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>

// Fiber switch.  glibc's swapcontext() saves / restores the signal mask with two system calls per switch; a kernel
// with a few barriers switches each of its threads tens of times, so the emulation would spend most of its time in the
// kernel of the host (and much more under syscall-intercepting sandboxes).  On x86-64 a 12-instruction switch of the
// callee-saved registers and the stack pointer replaces it; ucontext stays as the portable fallback and under
// AddressSanitizer (SIMT_USE_UCONTEXT).
#if defined(__x86_64__) && !defined(__SANITIZE_ADDRESS__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_FAST_SWITCH 1
extern "C" void simt_switch_ctx(void** save_sp, void* load_sp);
asm(".text\n"
    ".p2align 4\n"
    ".globl simt_switch_ctx\n"
    ".type simt_switch_ctx,@function\n"
    "simt_switch_ctx:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size simt_switch_ctx,.-simt_switch_ctx\n");
#endif

#include <atomic>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

namespace simt {

struct uint3_ { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

enum { RUNNABLE = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, WAIT_GRID = 3, DONE = 4 };

struct GridState {
  dim3 grid, block;
  bool cooperative = false;
  size_t dyn_smem = 0;            // bytes of dynamic shared memory of this launch (third <<<>>> argument)
  pthread_barrier_t bar;
};

struct BlockState;
struct PendingCopy { void* dst; const void* src; int bytes; unsigned group; };
struct Fiber {
  std::vector<PendingCopy> cp_pending;   // cp.async copies issued but not yet waited for (performed AT the wait)
  unsigned cp_group = 0;                 // index of the open (uncommitted) group
#ifdef SIMT_FAST_SWITCH
  void* sp = nullptr;
#else
  ucontext_t ctx;
#endif
  uint3_ tid;
  int lin = 0;
  int state = RUNNABLE;
  BlockState* blk = nullptr;
};

struct BlockState {
  GridState* g = nullptr;
  uint3_ bid;
  int nthreads = 0, nwarps = 0;
  std::vector<Fiber> fibers;
#ifdef SIMT_FAST_SWITCH
  void* sched_sp = nullptr;
#else
  ucontext_t sched;
#endif
  int live = 0, wait_block = 0, wait_grid = 0;
  std::vector<int> warp_live, warp_wait;
  std::vector<uint64_t> shfl;
  const std::function<void()>* body = nullptr;
  char* stacks = nullptr;
  size_t stack_bytes = 0;
  char* dyn = nullptr;            // this block's dynamic shared memory: a heap block of EXACTLY the launch's size, so
  size_t dyn_bytes = 0;           // that an AddressSanitizer build catches a kernel that outgrows what the host asked for
};

inline thread_local BlockState* tl_block = nullptr;
inline thread_local Fiber* tl_fiber = nullptr;

static const size_t kStack = 256 * 1024;

inline void to_scheduler() {
  Fiber* f = tl_fiber;
#ifdef SIMT_FAST_SWITCH
  simt_switch_ctx(&f->sp, f->blk->sched_sp);
#else
  swapcontext(&f->ctx, &f->blk->sched);
#endif
}

inline void yield() { to_scheduler(); }   // state stays RUNNABLE

// cp.async emulation: copies are queued per thread and performed when the thread waits for their group
// (cp.async.wait_group N = all but the N most recently committed groups are complete), never earlier -- a kernel that
// reads a stage before waiting for it reads whatever the buffer held before.
inline void cp_async(void* dst, const void* src, int bytes) {
  Fiber* f = tl_fiber;
  f->cp_pending.push_back(PendingCopy{dst, src, bytes, f->cp_group});
}
inline void cp_async_commit() { tl_fiber->cp_group++; }
inline void cp_async_wait(unsigned keep) {
  Fiber* f = tl_fiber;
  size_t w = 0;
  for (size_t i = 0; i < f->cp_pending.size(); ++i) {
    const PendingCopy c = f->cp_pending[i];
    if (c.group + keep < f->cp_group) memcpy(c.dst, c.src, (size_t)c.bytes);
    else f->cp_pending[w++] = c;
  }
  f->cp_pending.resize(w);
}

inline void release(BlockState* B, int what, int warp) {
  for (auto& f : B->fibers) {
    if (f.state != what) continue;
    if (what == WAIT_WARP && (f.lin >> 5) != warp) continue;
    f.state = RUNNABLE;
  }
}

inline void barrier_warp() {
  Fiber* f = tl_fiber;
  BlockState* B = f->blk;
  int w = f->lin >> 5;
  if (++B->warp_wait[w] >= B->warp_live[w]) {
    B->warp_wait[w] = 0;
    release(B, WAIT_WARP, w);
    return;
  }
  f->state = WAIT_WARP;
  to_scheduler();
}

inline void barrier_block() {
  Fiber* f = tl_fiber;
  BlockState* B = f->blk;
  if (++B->wait_block >= B->live) {
    B->wait_block = 0;
    release(B, WAIT_BLOCK, 0);
    return;
  }
  f->state = WAIT_BLOCK;
  to_scheduler();
}

inline void barrier_grid() {
  Fiber* f = tl_fiber;
  BlockState* B = f->blk;
  if (!B->g->cooperative) {
    fprintf(stderr, "simt: grid.sync() in a non-cooperative launch\n");
    abort();
  }
  ++B->wait_grid;
  f->state = WAIT_GRID;
  to_scheduler();   // the scheduler performs the OS-level barrier once every live fiber waits
}

// called by the trampoline when the kernel body returns for this fiber
inline void fiber_exit() {
  Fiber* f = tl_fiber;
  BlockState* B = f->blk;
  int w = f->lin >> 5;
  f->state = DONE;
  --B->live;
  --B->warp_live[w];
  if (B->warp_live[w] > 0 && B->warp_wait[w] >= B->warp_live[w]) { B->warp_wait[w] = 0; release(B, WAIT_WARP, w); }
  if (B->live > 0 && B->wait_block >= B->live) { B->wait_block = 0; release(B, WAIT_BLOCK, 0); }
  to_scheduler();
}

inline void trampoline() {
  (*tl_fiber->blk->body)();
  fiber_exit();
}

inline void run_block(BlockState& B, GridState* g, unsigned bx, unsigned by, unsigned bz, const std::function<void()>& body) {
  B.g = g;
  B.bid.x = bx; B.bid.y = by; B.bid.z = bz;
  B.nthreads = (int)(g->block.x * g->block.y * g->block.z);
  B.nwarps = (B.nthreads + 31) / 32;
  B.body = &body;
  if (B.stack_bytes < (size_t)B.nthreads * kStack) {
    if (B.stacks) munmap(B.stacks, B.stack_bytes);
    B.stack_bytes = (size_t)B.nthreads * kStack;
    B.stacks = (char*)mmap(nullptr, B.stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (B.stacks == MAP_FAILED) { perror("simt: mmap"); abort(); }
  }
  if (B.dyn_bytes != g->dyn_smem || (g->dyn_smem && !B.dyn)) {
    free(B.dyn);
    B.dyn = g->dyn_smem ? (char*)aligned_alloc(16, (g->dyn_smem + 15) & ~(size_t)15) : nullptr;
    B.dyn_bytes = g->dyn_smem;
  }
  if (B.dyn) memset(B.dyn, 0xFF, B.dyn_bytes);     // NaN-poison: shared memory is uninitialised at block start on the GPU
  B.fibers.assign(B.nthreads, Fiber());
  B.warp_live.assign(B.nwarps, 0);
  B.warp_wait.assign(B.nwarps, 0);
  B.shfl.assign(B.nthreads, 0);
  B.live = B.nthreads; B.wait_block = 0; B.wait_grid = 0;
  for (int t = 0; t < B.nthreads; ++t) {
    Fiber& f = B.fibers[t];
    f.blk = &B;
    f.lin = t;
    f.tid.x = t % g->block.x;
    f.tid.y = (t / g->block.x) % g->block.y;
    f.tid.z = t / (g->block.x * g->block.y);
    f.state = RUNNABLE;
    B.warp_live[t >> 5]++;
#ifdef SIMT_FAST_SWITCH
    {
      // initial frame: six zeroed callee-saved registers, then the entry point as return address of simt_switch_ctx;
      // at the entry the stack pointer must be 8 mod 16 (as after a call), and the trampoline never returns
      char* top = B.stacks + (size_t)(t + 1) * kStack;
      uintptr_t sp = ((uintptr_t)top & ~(uintptr_t)15) - 8;       // value of rsp at the trampoline's entry (8 mod 16)
      void** frame = (void**)sp;
      frame[0] = nullptr;                                         // fake return address of the trampoline
      frame[-1] = (void*)trampoline;                              // popped by `ret`
      for (int r = 2; r <= 7; ++r) frame[-r] = nullptr;           // r15 .. rbp
      f.sp = (void*)(frame - 7);
    }
#else
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = B.stacks + (size_t)t * kStack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
#endif
  }
  tl_block = &B;
  long idle = 0;
  // SIMT_SHUFFLE=<seed>: visit the runnable fibers in a pseudo-random order (changes every round) instead of thread-id
  // order, so that a missing barrier between a producer and a consumer thread shows up whichever of them has the
  // lower id
  static const char* shuffle_env = getenv("SIMT_SHUFFLE");
  std::vector<int> order(B.nthreads);
  for (int t = 0; t < B.nthreads; ++t) order[t] = t;
  uint64_t rng = shuffle_env ? (uint64_t)atoll(shuffle_env) * 0x9E3779B97F4A7C15ull + bx * 7919u + by * 104729u + 1 : 0;
  while (B.live > 0) {
    bool ran = false;
    if (shuffle_env) {
      for (int t = B.nthreads - 1; t > 0; --t) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        int u = (int)(rng % (uint64_t)(t + 1));
        int tmp = order[t]; order[t] = order[u]; order[u] = tmp;
      }
    }
    for (int ti = 0; ti < B.nthreads; ++ti) {
      const int t = order[ti];
      Fiber& f = B.fibers[t];
      if (f.state != RUNNABLE) continue;
      tl_fiber = &f;
#ifdef SIMT_FAST_SWITCH
      simt_switch_ctx(&B.sched_sp, f.sp);
#else
      swapcontext(&B.sched, &f.ctx);
#endif
      ran = true;
    }
    if (B.live == 0) break;
    if (!ran) {
      if (B.wait_grid > 0 && B.wait_grid >= B.live) {
        pthread_barrier_wait(&g->bar);
        B.wait_grid = 0;
        release(&B, WAIT_GRID, 0);
      } else {
        fprintf(stderr, "simt: deadlock in block (%u,%u): live %d, at block barrier %d, at grid barrier %d\n", bx, by,
                B.live, B.wait_block, B.wait_grid);
        abort();
      }
    } else if ((++idle & 1023) == 0) {
      sched_yield();
    }
  }
  tl_fiber = nullptr;
  tl_block = nullptr;
}

inline void free_block(BlockState& B) {
  free(B.dyn);
  B.dyn = nullptr;
  B.dyn_bytes = 0;
  if (B.stacks) munmap(B.stacks, B.stack_bytes);
  B.stacks = nullptr;
  B.stack_bytes = 0;
}

// Launch `body` once per CUDA thread.  cooperative: all blocks concurrently (grid.sync allowed).
inline void* dyn_smem() { return tl_block->dyn; }       // `extern __shared__` of the running block

inline void launch(dim3 grid, dim3 block, bool cooperative, const std::function<void()>& body, size_t dyn_smem_bytes = 0) {
  GridState g;
  g.grid = grid; g.block = block; g.cooperative = cooperative; g.dyn_smem = dyn_smem_bytes;
  int64_t nblocks = (int64_t)grid.x * grid.y * grid.z;
  if (nblocks <= 0) return;
  if (cooperative) {
    pthread_barrier_init(&g.bar, nullptr, (unsigned)nblocks);
    std::vector<std::thread> th;
    for (int64_t b = 0; b < nblocks; ++b)
      th.emplace_back([&, b]() {
        BlockState B;
        run_block(B, &g, (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((int64_t)grid.x * grid.y)), body);
        free_block(B);
      });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g.bar);
    return;
  }
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  if (hw > 8) hw = 8;
  int64_t nworkers = nblocks < (int64_t)hw ? nblocks : (int64_t)hw;
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    BlockState B;
    for (;;) {
      int64_t b = next.fetch_add(1);
      if (b >= nblocks) break;
      run_block(B, &g, (unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((int64_t)grid.x * grid.y)), body);
    }
    free_block(B);
  };
  if (nworkers == 1) { worker(); return; }
  std::vector<std::thread> th;
  for (int64_t w = 0; w < nworkers; ++w) th.emplace_back(worker);
  for (auto& t : th) t.join();
}

template <typename T>
inline T shfl_generic(T v, int src_lane_in_warp) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8-byte values");
  Fiber* f = tl_fiber;
  BlockState* B = f->blk;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  B->shfl[f->lin] = bits;
  barrier_warp();
  int src = (f->lin & ~31) | (src_lane_in_warp & 31);
  if (src >= B->nthreads) src = f->lin;
  uint64_t got = B->shfl[src];
  barrier_warp();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace simt

// ---- CUDA surface ------------------------------------------------------------------------------
using simt::dim3;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static thread_local
#define threadIdx (simt::tl_fiber->tid)
#define blockIdx (simt::tl_block->bid)
#define blockDim (simt::tl_block->g->block)
#define gridDim (simt::tl_block->g->grid)

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
struct int4_ { int x, y, z, w; };
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

inline void __syncthreads() { simt::barrier_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::barrier_warp(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return simt::shfl_generic(v, (simt::tl_fiber->lin & 31) ^ lane_mask);
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane) { return simt::shfl_generic(v, src_lane); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  int l = simt::tl_fiber->lin & 31;
  return simt::shfl_generic(v, (l + (int)delta < 32) ? l + (int)delta : l);
}

template <typename T> inline T __ldcs(const T* p) { return *(const volatile T*)p; }
template <typename T> inline T __ldcg(const T* p) { return *(const volatile T*)p; }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
template <typename T> inline void __stcg(T* p, T v) { *p = v; }
template <> inline double2 __ldcs<double2>(const double2* p) { return *p; }
template <> inline double2 __ldcg<double2>(const double2* p) { return *p; }

inline double atomicAdd(double* p, double v) {
  uint64_t* u = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(u, __ATOMIC_RELAXED);
  for (;;) {
    double d;
    memcpy(&d, &old, 8);
    d += v;
    uint64_t nw;
    memcpy(&nw, &d, 8);
    if (__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {
      double r;
      memcpy(&r, &old, 8);
      return r;
    }
  }
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
  return old;
}

inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }

namespace cooperative_groups {
struct grid_group {
  void sync() const { simt::barrier_grid(); }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
