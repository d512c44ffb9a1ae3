// simt.h — one source, two targets.
//
// Under nvcc this header is (almost) empty: the kernels use the real CUDA built-ins.
// Under a plain host compiler (g++ -DSMRT_SIMT_EMULATION) it provides a tiny SIMT emulator: one OS thread per CUDA
// thread of ONE block at a time, __syncthreads()/named barriers as pthread barriers, warp shuffles through a per-warp
// mailbox.  The emulator exists so that the device code can be exercised by the CPU test-suite in the authoring
// container, which has nvcc but no GPU (tests/test_simt_emulation.py).  It is test infrastructure: the shipped
// library (libsmrt_dort_b200.so) is compiled by nvcc for sm_100a only and contains none of it.
#pragma once

#ifdef __CUDACC__

#include <cuda_runtime.h>
#define SMRT_DEV __device__ __forceinline__
#define SMRT_HD __host__ __device__ __forceinline__
#define SMRT_DEV_NOINLINE __device__ __noinline__
#define SMRT_GLOBAL __global__
#define SMRT_SHARED __shared__
#define SMRT_RESTRICT __restrict__

SMRT_DEV void smrt_named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// a routine that is not inlined (one copy serves several kernel instantiations) receives generic pointers: where the
// caller knows they point to shared memory, saying so turns LD.E / ST.E + 64-bit address arithmetic into LDS / STS
#define SMRT_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
// an opaque identity: the value must be kept (register or spill), it cannot be re-derived inside a loop
#define SMRT_KEEP_INT(x) asm volatile("" : "+r"(x))

// a == b ? x : y as ONE predicated select the optimiser cannot rewrite (it would turn a chain of them over the elements
// of a register array into a dynamically indexed, i.e. local-memory, array)
SMRT_DEV double smrt_select_eq(int a, int b, double x, double y) {
  double r;
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %3, %4;\n selp.f64 %0, %1, %2, p;\n}\n" : "=d"(r) : "d"(x), "d"(y), "r"(a), "r"(b));
  return r;
}

// ---- single-instruction fp64 approximations (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20 relative error): seeds that the
// callers refine with Newton steps where they need more
SMRT_DEV double smrt_rcp_approx(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
SMRT_DEV double smrt_rsqrt_approx(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// ---- asynchronous 8-byte copies global -> shared (LDGSTS), grouped; `valid` = false writes zeros without reading
SMRT_DEV void smrt_cp_async8(double* dst_smem, const double* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src),
               "r"(valid ? 8 : 0)
               : "memory");
}
SMRT_DEV void smrt_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SMRT_DEV void smrt_cp_async_wait() {  // at most N of this thread's groups still in flight
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMA bulk copy (cp.async.bulk, 1-D, global -> shared) completed on an mbarrier --------------------------------
typedef unsigned long long smrt_mbar_t;
SMRT_DEV unsigned smrt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SMRT_DEV void smrt_mbar_init(smrt_mbar_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smrt_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: order prior generic-proxy accesses of the destination before the async-proxy writes, arm the barrier with
// the byte count and issue the copies (each a multiple of 16 bytes, 16-byte aligned)
SMRT_DEV void smrt_bulk_load2(smrt_mbar_t* bar, void* dst0, const void* src0, void* dst1, const void* src1,
                              unsigned bytes_each) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smrt_smem_u32(bar)), "r"(2u * bytes_each)
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smrt_smem_u32(dst0)),
               "l"(src0), "r"(bytes_each), "r"(smrt_smem_u32(bar))
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smrt_smem_u32(dst1)),
               "l"(src1), "r"(bytes_each), "r"(smrt_smem_u32(bar))
               : "memory");
}
// single copy under one barrier phase
SMRT_DEV void smrt_bulk_load1(smrt_mbar_t* bar, void* dst, const void* src, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smrt_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smrt_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smrt_smem_u32(bar))
               : "memory");
}
// every thread that wrote global memory a later bulk copy of this CTA will read: order its (generic-proxy) stores before the
// async-proxy reads; followed by a block barrier, then the issuing thread
SMRT_DEV void smrt_fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// hint: bring [src, src + bytes) into L2 (bytes a multiple of 16)
SMRT_DEV void smrt_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// every consumer thread: wait for the phase with the given parity
SMRT_DEV void smrt_mbar_wait(smrt_mbar_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smrt_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

#else  // ------------------------------------------------------------------------------------------ host emulation

#include <pthread.h>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define SMRT_DEV inline
#define SMRT_HD inline
#define SMRT_DEV_NOINLINE inline
#define SMRT_GLOBAL
#define SMRT_SHARED static
#define SMRT_RESTRICT __restrict__
#define __launch_bounds__(...)  /* nothing */
#define __forceinline__ inline
#define __device__
#define __host__
#define __global__
#define __constant__ static

struct double2 {
  double x, y;
};
struct simt_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
extern thread_local simt_dim3 threadIdx;
extern thread_local simt_dim3 blockIdx;
extern simt_dim3 blockDim;
extern simt_dim3 gridDim;

namespace simt {
struct BlockState {
  int nthreads = 0;
  pthread_barrier_t block_barrier;
  std::vector<pthread_barrier_t> warp_barrier;        // one per warp
  std::vector<std::vector<uint64_t>> warp_mailbox;    // [warp][32]
  std::map<std::pair<int, int>, pthread_barrier_t*> named;
  std::mutex named_mutex;
  std::atomic<int> or_flag[2];
};
extern BlockState* g_block;

// Run `body()` for every thread of every block of the grid; blocks run one after the other.
void launch(unsigned grid, unsigned block, const std::function<void()>& body);
unsigned char* dynamic_smem(size_t bytes);  // per-launch scratch, shared by the threads of the current block
}  // namespace simt

void __syncthreads();
int __syncthreads_or(int pred);
void __syncwarp(unsigned mask = 0xffffffffu);
void simt_group_barrier(unsigned mask);
void __threadfence();
void smrt_named_barrier(int id, int nthreads);
#define SMRT_ASSUME_SHARED(p) ((void)0)
#define SMRT_KEEP_INT(x) ((void)0)

uint64_t simt_shfl_raw(unsigned mask, uint64_t v, int src_lane);
int __any_sync(unsigned mask, int pred);
unsigned __ballot_sync(unsigned mask, int pred);
unsigned __reduce_max_sync(unsigned mask, unsigned v);
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __double2hiint(double v) {
  uint64_t b;
  std::memcpy(&b, &v, 8);
  return (int)(b >> 32);
}

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src_lane, int width = 32) {
  static_assert(sizeof(T) <= 8, "shuffle of <= 8 byte types only");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  int lane = threadIdx.x & 31;
  int base = lane & ~(width - 1);
  raw = simt_shfl_raw(mask, raw, base + (src_lane & (width - 1)));
  T out;
  std::memcpy(&out, &raw, sizeof(T));
  return out;
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int width = 32) {
  int lane = threadIdx.x & 31;
  return __shfl_sync(m, v, (lane ^ lane_mask) & (width - 1) | (lane & ~(width - 1)), 32);
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int width = 32) {
  int lane = threadIdx.x & 31;
  int src = lane + (int)delta;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return __shfl_sync(m, v, src, 32);
}

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int* p, int compare, int v) {
  __atomic_compare_exchange_n(p, &compare, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return compare;  // the value seen (== compare on success)
}
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}

// emulation of the TMA bulk copy: the issuing thread copies synchronously; the wait is a no-op (the kernels place a
// block barrier between the wait and the first use, which the emulator needs for visibility)
typedef unsigned long long smrt_mbar_t;
inline void smrt_mbar_init(smrt_mbar_t* bar, unsigned) { *bar = 0; }
inline void smrt_bulk_load2(smrt_mbar_t*, void* dst0, const void* src0, void* dst1, const void* src1,
                            unsigned bytes_each) {
  std::memcpy(dst0, src0, bytes_each);
  std::memcpy(dst1, src1, bytes_each);
}
inline void smrt_bulk_load1(smrt_mbar_t*, void* dst, const void* src, unsigned bytes) { std::memcpy(dst, src, bytes); }
inline void smrt_prefetch_l2(const void*, unsigned) {}
inline void smrt_fence_async_global() {}
inline void smrt_mbar_wait(smrt_mbar_t*, unsigned) {}

inline double smrt_select_eq(int a, int b, double x, double y) { return a == b ? x : y; }
// emulation of the asynchronous copies: done at once
inline void smrt_cp_async8(double* dst, const double* src, bool valid) { *dst = valid ? *src : 0.0; }
inline void smrt_cp_async_commit() {}
template <int N>
inline void smrt_cp_async_wait() {}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
// the device versions are ~20-bit seeds: keep only a float mantissa so that the emulation tests the same tolerance
inline double smrt_approx_round(double v) {
  int e;
  double m = std::frexp(v, &e);
  return std::ldexp((double)(float)m, e);
}
inline double smrt_rcp_approx(double x) { return smrt_approx_round(1.0 / x); }
inline double smrt_rsqrt_approx(double x) { return smrt_approx_round(1.0 / std::sqrt(x)); }
inline double fma_(double a, double b, double c) { return std::fma(a, b, c); }
inline void sincos(double x, double* s, double* c) {
  *s = std::sin(x);
  *c = std::cos(x);
}
inline void sincospi(double x, double* s, double* c) {
  // exact at the multiples of 1/2 like the CUDA intrinsic
  double r = std::fmod(x, 2.0);
  if (r == 0.0) { *s = 0.0; *c = 1.0; }
  else if (r == 0.5) { *s = 1.0; *c = 0.0; }
  else if (r == 1.0) { *s = 0.0; *c = -1.0; }
  else if (r == 1.5) { *s = -1.0; *c = 0.0; }
  else { *s = std::sin(3.141592653589793238462643383279502884 * r); *c = std::cos(3.141592653589793238462643383279502884 * r); }
}
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }

#endif
