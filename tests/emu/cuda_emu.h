// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny single-OS-thread CUDA execution emulator: every CUDA thread of a CTA is a
// cooperative fiber, CTAs of a grid run one after another.  It exists so that the
// product's .cu sources (lossless-audio-codec_b200/csrc) can be compiled with g++
// (-x c++ -include this file) into tests/emu/liblac_b200_emu.so and debugged to
// parity against the oracle in a container that has no GPU.  The product library
// never includes this header and never falls back to it; GPU tests use the real
// nvcc-built liblac_b200.so.
//
// Supported subset: 1-D/2-D grids and 1-D blocks, static and dynamic shared memory,
// __syncthreads(_or/_and/_count), __syncwarp, full warp shuffles / ballots / votes,
// integer atomics, the bit/arith intrinsics the codec uses, and the handful of
// runtime calls (malloc/memcpy/memset/streams/events) the C-ABI layer makes.
// Set LACB_EMU_SHUFFLE=1 to randomise the fiber schedule (shakes out missing
// barriers that a fixed round-robin order would hide).
#pragma once
#ifndef LACB_EMU
#define LACB_EMU 1
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static
// Static shared memory: CTAs run one at a time, so one static copy is enough.
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace emu {

struct Fiber {
  void* sp = nullptr;
  bool done = false;
  unsigned tid = 0;
};

struct WarpX {
  uint64_t slots[32];
  uint64_t snap[2][32];
  uint32_t arrived = 0;
  uint32_t gen = 0;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<WarpX> warps;
  unsigned nthreads = 0, alive = 0;
  unsigned bar_arrived = 0, bar_gen = 0;
  unsigned bar_acc_or[2] = {0, 0}, bar_acc_and[2] = {1, 1}, bar_acc_cnt[2] = {0, 0};
  std::function<void()> body;
};

extern Block* g_block;
extern Fiber* g_cur;
extern void* g_sched_sp;
extern uint3 g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
extern uint64_t g_progress;

extern "C" void emu_switch(void** save_sp, void* new_sp);
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

struct TidProxy {
  operator uint3() const { return uint3{g_cur->tid, 0, 0}; }
};
struct ThreadIdxT {
  struct X { operator unsigned() const { return g_cur->tid; } } x;
  struct Z { operator unsigned() const { return 0; } } y, z;
};
struct BlockIdxT {
  struct X { operator unsigned() const { return g_blockIdx.x; } } x;
  struct Y { operator unsigned() const { return g_blockIdx.y; } } y;
  struct Z { operator unsigned() const { return g_blockIdx.z; } } z;
};
struct BlockDimT {
  struct X { operator unsigned() const { return g_blockDim.x; } } x;
  struct Y { operator unsigned() const { return 1; } } y, z;
};
struct GridDimT {
  struct X { operator unsigned() const { return g_gridDim.x; } } x;
  struct Y { operator unsigned() const { return g_gridDim.y; } } y;
  struct Z { operator unsigned() const { return g_gridDim.z; } } z;
};

// --- warp collective core -------------------------------------------------
inline const uint64_t* warp_exchange(unsigned mask, uint64_t v) {
  Block& b = *g_block;
  const unsigned tid = g_cur->tid, lane = tid & 31u;
  WarpX& w = b.warps[tid >> 5];
  // restrict the mask to lanes that exist in this (possibly partial) warp
  const unsigned wbase = tid & ~31u;
  const unsigned lanes = std::min(32u, b.nthreads - wbase);
  const unsigned exist = lanes == 32 ? 0xFFFFFFFFu : ((1u << lanes) - 1u);
  mask &= exist;
  w.slots[lane] = v;
  w.arrived |= 1u << lane;
  const uint32_t mygen = w.gen;
  ++g_progress;
  if ((w.arrived & mask) == mask) {
    memcpy(w.snap[mygen & 1u], w.slots, sizeof w.slots);
    w.arrived &= ~mask;
    w.gen++;
  } else {
    while (w.gen == mygen) yield();
  }
  return w.snap[mygen & 1u];
}

}  // namespace emu

#define threadIdx (emu::ThreadIdxT{})
#define blockIdx (emu::BlockIdxT{})
#define blockDim (emu::BlockDimT{})
#define gridDim (emu::GridDimT{})
static const int warpSize = 32;

// --- barriers ---------------------------------------------------------------
static inline int emu_bar(int pred, int kind) {
  emu::Block& b = *emu::g_block;
  const unsigned g = b.bar_gen & 1u;
  b.bar_acc_or[g] |= (pred != 0);
  b.bar_acc_and[g] &= (pred != 0);
  b.bar_acc_cnt[g] += (pred != 0);
  const unsigned mygen = b.bar_gen;
  ++emu::g_progress;
  if (++b.bar_arrived >= b.alive) {
    b.bar_arrived = 0;
    const unsigned ng = (mygen + 1u) & 1u;
    b.bar_acc_or[ng] = 0;
    b.bar_acc_and[ng] = 1;
    b.bar_acc_cnt[ng] = 0;
    b.bar_gen++;
  } else {
    while (b.bar_gen == mygen) emu::yield();
  }
  return kind == 0 ? (int)b.bar_acc_or[g] : kind == 1 ? (int)b.bar_acc_and[g] : (int)b.bar_acc_cnt[g];
}
static inline void __syncthreads() { emu_bar(0, 0); }
static inline int __syncthreads_or(int p) { return emu_bar(p, 0); }
static inline int __syncthreads_and(int p) { return emu_bar(p, 1); }
static inline int __syncthreads_count(int p) { return emu_bar(p, 2); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::warp_exchange(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// --- shuffles / votes ---------------------------------------------------------
template <typename T>
static inline uint64_t emu_bits(T v) {
  uint64_t r = 0;
  static_assert(sizeof(T) <= 8, "shuffle payload too large");
  memcpy(&r, &v, sizeof(T));
  return r;
}
template <typename T>
static inline T emu_unbits(uint64_t r) {
  T v;
  memcpy(&v, &r, sizeof(T));
  return v;
}
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const unsigned lane = emu::g_cur->tid & 31u;
  const uint64_t* s = emu::warp_exchange(mask, emu_bits(v));
  const unsigned base = lane & ~(unsigned)(width - 1);
  return emu_unbits<T>(s[base + ((unsigned)src & (unsigned)(width - 1))]);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const unsigned lane = emu::g_cur->tid & 31u;
  const uint64_t* s = emu::warp_exchange(mask, emu_bits(v));
  const unsigned base = lane & ~(unsigned)(width - 1);
  return (lane - base >= delta) ? emu_unbits<T>(s[lane - delta]) : v;
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const unsigned lane = emu::g_cur->tid & 31u;
  const uint64_t* s = emu::warp_exchange(mask, emu_bits(v));
  const unsigned base = lane & ~(unsigned)(width - 1);
  return (lane - base + delta < (unsigned)width) ? emu_unbits<T>(s[lane + delta]) : v;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  const unsigned lane = emu::g_cur->tid & 31u;
  const uint64_t* s = emu::warp_exchange(mask, emu_bits(v));
  const unsigned src = lane ^ (unsigned)lanemask;
  const unsigned base = lane & ~(unsigned)(width - 1);
  return (src - base < (unsigned)width) ? emu_unbits<T>(s[src]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  const uint64_t* s = emu::warp_exchange(mask, (uint64_t)(pred != 0));
  emu::Block& b = *emu::g_block;
  const unsigned wbase = emu::g_cur->tid & ~31u;
  const unsigned lanes = std::min(32u, b.nthreads - wbase);
  unsigned r = 0;
  for (unsigned i = 0; i < lanes; ++i)
    if (((mask >> i) & 1u) && s[i]) r |= 1u << i;
  return r;
}
// warp reductions (REDUX on sm_80+): every existing lane named in the mask contributes
static inline unsigned emu_reduce(unsigned mask, unsigned v, int op) {
  const uint64_t* s = emu::warp_exchange(mask, (uint64_t)v);
  emu::Block& b = *emu::g_block;
  const unsigned wbase = emu::g_cur->tid & ~31u;
  const unsigned lanes = std::min(32u, b.nthreads - wbase);
  unsigned r = op == 2 ? 0u : 0u;
  bool first = true;
  for (unsigned i = 0; i < lanes; ++i) {
    if (!((mask >> i) & 1u)) continue;
    const unsigned x = (unsigned)s[i];
    if (first) { r = x; first = false; continue; }
    if (op == 0) r += x;
    else if (op == 1) r |= x;
    else if (op == 2) r = x > r ? x : r;
    else r = x < r ? x : r;
  }
  return r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, 0); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, 1); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, 2); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, 3); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
  emu::Block& b = *emu::g_block;
  const unsigned wbase = emu::g_cur->tid & ~31u;
  const unsigned lanes = std::min(32u, b.nthreads - wbase);
  const unsigned exist = lanes == 32 ? 0xFFFFFFFFu : ((1u << lanes) - 1u);
  return (__ballot_sync(mask, pred) & mask & exist) == (mask & exist);
}
static inline unsigned __activemask() { return 0xFFFFFFFFu; }

// --- atomics (single OS thread: plain read-modify-write) ----------------------
template <typename T, typename V>
static inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <typename T, typename V>
static inline T atomicOr(T* p, V v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <typename T, typename V>
static inline T atomicAnd(T* p, V v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <typename T, typename V>
static inline T atomicMax(T* p, V v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <typename T, typename V>
static inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <typename T, typename V>
static inline T atomicExch(T* p, V v) { T o = *p; *p = (T)v; return o; }
template <typename T, typename V>
static inline T atomicCAS(T* p, V c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// --- intrinsics ------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(v);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 0xFu;
    unsigned byte = (unsigned)((src >> (8 * (sel & 7u))) & 0xFFu);
    if (sel & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
  sh &= 31u;
  return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned sh) {
  if (sh >= 32u) return lo;
  return sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  sh &= 31u;
  return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// --- runtime API subset -------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaHostAllocDefault = 0, cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t totalGlobalMem; int major, minor; };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
template <typename T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
enum { cudaHostRegisterPortable = 1 };
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return 0; }
static inline cudaError_t cudaHostUnregister(void*) { return 0; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = 0; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof *p);
  p->multiProcessorCount = 2;
  strcpy(p->name, "emu");
  p->major = 10;
  return 0;
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{0}; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }

// Dynamic shared memory: `extern __shared__ T name[];` is spelled through this macro
// in the product sources (LACB_DYN_SMEM) so both builds agree.
#define LACB_EMU_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_dyn_smem)

// Kernel launch: LACB_LAUNCH(kernel, grid, block, smem, stream, args...)
#define LACB_EMU_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
