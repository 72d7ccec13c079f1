// Portability prelude of the kernel headers (*_kernels.cuh, elem_math.cuh).
//
// Product build (nvcc, sm_100a): the CUDA runtime headers and PTX system-scope accessors.
// FEMCY_SIMT_EMU (g++, tests/simt only): the CPU SIMT emulation that lets the not-gpu test suite execute the
// kernel source for logic checks.  The product library is never built with FEMCY_SIMT_EMU and has no CPU path.
#pragma once
#include <stdint.h>

struct femcy_d4 { double x, y, z, w; };   // one 32-byte node-sector record

#ifdef FEMCY_SIMT_EMU
#include "simt.h"
#define FEMCY_SPIN_PAUSE() simt::yield()
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  __atomic_store_n(p, v, __ATOMIC_RELAXED);
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  return __atomic_load_n(p, __ATOMIC_RELAXED);
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
// asynchronous global -> shared copies: the emulation DEFERS them to the wait, so a kernel that reads a buffer before
// waiting for its group sees stale (NaN-poisoned) shared memory, as it could on the GPU
template <int BYTES> inline void femcy_cp_async(void* smem_dst, const void* gsrc) { simt::cp_async(smem_dst, gsrc, BYTES); }
inline void femcy_cp_async_commit() { simt::cp_async_commit(); }
template <int KEEP> inline void femcy_cp_async_wait() { simt::cp_async_wait(KEEP); }
inline void femcy_red_add_f64(double* p, double v) { atomicAdd(p, v); }
// mbarrier + bulk global -> shared loads, emulated with the barrier word as
//   {phase parity : 1 | pending arrivals : 15 | arrival count of init : 16 | signed tx bytes : 32}
// (fibers of a block interleave only at yields, so plain read-modify-write is enough).  A phase completes -- the parity
// flips and the pending count re-arms -- when both the pending arrivals and the transaction bytes reach zero.
inline void femcy_mbar_settle_(unsigned long long* bar) {
  unsigned long long v = *bar;
  if (((v >> 48) & 0x7fffull) == 0ull && (unsigned)(v & 0xffffffffull) == 0u) {
    unsigned long long init = (v >> 32) & 0xffffull;
    *bar = ((v ^ (1ull << 63)) & (1ull << 63)) | (init << 48) | (init << 32);
  }
}
inline void femcy_mbar_init(unsigned long long* bar, unsigned count) {
  *bar = ((unsigned long long)count << 48) | ((unsigned long long)count << 32);
}
inline void femcy_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  unsigned long long v = *bar;
  int tx = (int)(unsigned)(v & 0xffffffffull) + (int)bytes;
  unsigned long long pend = ((v >> 48) & 0x7fffull) - 1ull;
  *bar = (v & (1ull << 63)) | ((pend & 0x7fffull) << 48) | (v & (0xffffull << 32)) | (unsigned)tx;
  femcy_mbar_settle_(bar);
}
inline void femcy_bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  memcpy(sdst, gsrc, bytes);
  unsigned long long v = *bar;
  int tx = (int)(unsigned)(v & 0xffffffffull) - (int)bytes;
  *bar = (v & 0xffffffff00000000ull) | (unsigned)tx;
  femcy_mbar_settle_(bar);
}
inline void femcy_bulk_load_stream(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  femcy_bulk_load(sdst, gsrc, bytes, bar);
}
inline void femcy_mbar_wait(unsigned long long* bar, unsigned parity) {
  while (((*(volatile unsigned long long*)bar) >> 63) == (unsigned long long)(parity & 1u)) simt::yield();
}
// bulk shared -> global store by ONE thread (the emulation copies at once; the caller has synchronised the block)
inline void femcy_bulk_store(void* gdst, const void* ssrc, unsigned bytes) { memcpy(gdst, ssrc, bytes); }
inline void femcy_fence_async_smem() {}
inline femcy_d4 femcy_ld256_nc(const double* p) { femcy_d4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3]; return v; }
inline unsigned long long femcy_globaltimer() { return 0ull; }
#else
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#define FEMCY_SPIN_PAUSE() ((void)0)
// system-scope (NVLink peer / host visible) 8-byte accessors; an aligned 8-byte access is single-copy atomic
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// device-scope acquire / release on a 32-bit word (grid-wide generation flags)
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// asynchronous global -> shared copies (LDGSTS): BYTES in {4, 8, 16}, both addresses BYTES-aligned; completion is per
// thread -- wait for the group, then a barrier, before other threads read the data
template <int BYTES>
__device__ __forceinline__ void femcy_cp_async(void* smem_dst, const void* gsrc) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async copies 4, 8 or 16 bytes");
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void femcy_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int KEEP>
__device__ __forceinline__ void femcy_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(KEEP) : "memory"); }
// fp64 reduction without a return value (RED: fire-and-forget).  Inside the cooperative PCG kernels ptxas turns a plain
// atomicAdd whose result is unused into ATOMG (which waits for the L2's reply); the explicit red keeps it a reduction.
__device__ __forceinline__ void femcy_red_add_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
// nanosecond device clock (the same time base on every SM): phase stamps inside the persistent PCG kernel
__device__ __forceinline__ unsigned long long femcy_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// 256-bit read-only global load (sm_100: LDG.E.256.CONSTANT): one instruction per 32-byte record; p is 32-byte aligned
__device__ __forceinline__ femcy_d4 femcy_ld256_nc(const double* p) {
  femcy_d4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
// bulk shared -> global store (TMA engine, non-tensor: cp.async.bulk, UBLKCP in the SASS), issued by ONE thread after the
// block has synchronised on the tile: 16-byte aligned addresses, size a multiple of 16.  Returns when the source may
// be reused (wait_group.read); the global writes complete asynchronously, ordered before the end of the kernel.
// every thread that wrote the tile with ordinary stores executes this BEFORE the block barrier that precedes the bulk
// store: it makes the thread's generic-proxy shared-memory writes visible to the async proxy (the TMA engine)
__device__ __forceinline__ void femcy_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void femcy_bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
  const unsigned src = (unsigned)__cvta_generic_to_shared(ssrc);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the tile -> async proxy
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// mbarrier + bulk global -> shared loads (TMA engine, non-tensor): a block initialises the barrier with ONE pending arrival,
// that thread arrives with the total byte count, any thread issues copies that complete_tx on the barrier, everybody
// waits on the phase parity.  16-byte aligned addresses, sizes multiples of 16.
__device__ __forceinline__ void femcy_mbar_init(unsigned long long* bar, unsigned count) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void femcy_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void femcy_bulk_load(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
// the same with an L2 evict-first hint: a stream that is read once per pass must not push the reused vectors out of L2
__device__ __forceinline__ void femcy_bulk_load_stream(void* sdst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b), "l"(pol) : "memory");
}
__device__ __forceinline__ void femcy_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  // bounded: a transaction count that never completes (a bug) must end in a launch failure, not in a hung GPU
  for (int tries = 0; tries < (1 << 22); ++tries) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    if (ok) return;
  }
  asm volatile("trap;");
}
#endif
