// lacb_common.cuh -- shared device primitives of the B200 LAC block codec.
//
// Everything here is integer/bit work; nothing uses tensor cores (the path is
// ALU/HBM bound).  Reference citations are relative to the upstream
// tree (audexdev/Lossless-Audio-Codec v1.4.0).
#pragma once
#include <stdint.h>

#ifdef LACB_EMU
// CPU fiber emulator build (tests/emu): test infrastructure, never shipped.
#define LACB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  LACB_EMU_LAUNCH(kernel, grid, block, smem, stream, __VA_ARGS__)
#define LACB_DYN_SMEM(type, name) LACB_EMU_DYN_SMEM(type, name)
#else
#include <cuda_runtime.h>
#define LACB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define LACB_DYN_SMEM(type, name)                                  \
  extern __shared__ __align__(16) unsigned char name##_raw_[];     \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

namespace lacb {

// src/codec/block/constants.hpp:6-15, block/encoder.cpp:41-59, rice/rice.hpp:12-13
constexpr uint32_t kMaxBlock = 16384;
constexpr uint32_t kMinPart = 32;
constexpr uint32_t kMaxPartOrder = 8;
constexpr uint32_t kZrMinRun = 4;
constexpr uint32_t kZrRunK = 2;
constexpr uint32_t kDriftWin = 256;
constexpr uint32_t kMicroWin = 96;
enum : uint32_t { MODE_RICE = 0, MODE_ZR = 1, MODE_BIN = 2, MODE_STATIC = 3 };
enum : uint32_t { PRED_FIXED = 0, PRED_FIR = 1, PRED_LPC = 2 };
constexpr uint32_t kFull = 0xFFFFFFFFu;

typedef unsigned long long u64;
typedef long long i64;

// Optional phase clocks (-DLACB_PHASE_CLK, tools/phase_clk.py): cycles every warp spends in
// each numbered phase of k_analyze, accumulated in a global table.  Compiled out of the product.
#if defined(LACB_PHASE_CLK) && !defined(LACB_EMU)
__device__ unsigned long long g_phase_clk[64 + 3 * 32];  // [64 + 32 * i + warp]: busy cycles of phases 6 / 8 / 10 per warp
struct PhState { long long last[32]; int base; };
__device__ __forceinline__ PhState* ph_state() { __shared__ PhState st; return &st; }
__device__ __forceinline__ void ph_init(int base) {
  PhState* st = ph_state();
  if ((threadIdx.x & 31u) == 0u) st->last[threadIdx.x >> 5] = clock64();
  if (threadIdx.x == 0u) st->base = base;
}
__device__ __forceinline__ void ph_base(int base) { if (threadIdx.x == 0u) ph_state()->base = base; }
__device__ __forceinline__ void ph_mark(int id) {
  PhState* st = ph_state();
  if ((threadIdx.x & 31u) == 0u && st->base >= 0) {
    const long long t = clock64();
    atomicAdd(&g_phase_clk[st->base + id], (unsigned long long)(t - st->last[threadIdx.x >> 5]));
    if (st->base == 0 && (id == 6 || id == 8 || id == 10))
      atomicAdd(&g_phase_clk[64 + 32 * ((id - 6) >> 1) + (threadIdx.x >> 5)], (unsigned long long)(t - st->last[threadIdx.x >> 5]));
    st->last[threadIdx.x >> 5] = t;
  }
}
#define LACB_PH_INIT(b) ph_init(b)
#define LACB_PH_BASE(b) ph_base(b)
#define LACB_PH(id) ph_mark(id)
#else
#define LACB_PH_INIT(b)
#define LACB_PH_BASE(b)
#define LACB_PH(id)
#endif

// zig-zag, block/encoder.cpp:61-65 and rice/rice.cpp:7-15
__device__ __forceinline__ uint32_t zz32(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }
__device__ __forceinline__ int32_t unzz32(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }

// c + a * b with a signed 32 x 32 -> 64 bit multiply-add: spelled as the PTX instruction so the
// operands stay 32-bit registers (the C form makes the compiler keep sign-extended 64-bit
// copies of every sample, which doubles the register footprint of the FIR loops)
__device__ __forceinline__ i64 mad_wide(int32_t a, int32_t b, i64 c) {
#ifdef LACB_EMU
  return (i64)((u64)((i64)a * (i64)b) + (u64)c);
#else
  i64 d;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
#endif
}

__device__ __forceinline__ uint32_t bitwidth64(u64 v) { return 64u - (uint32_t)__clzll((i64)v); }

// Conflict-free shared-memory layout for "thread t owns E consecutive words":
// 16-byte chunks are XOR-swizzled inside groups of 64 chunks so that the LDS.128
// a quarter-warp issues for chunk c of 8 consecutive threads hits 8 distinct bank
// groups for E = 8, 16 or 32.
__device__ __forceinline__ uint32_t swz_chunk(uint32_t q) { return q ^ ((q >> 3) & 7u); }
__device__ __forceinline__ uint32_t swz(uint32_t i) { return (swz_chunk(i >> 2) << 2) | (i & 3u); }

// ---------------------------------------------------------------------------
// Adaptive Rice parameter, division-free closed forms.
//
// Exact base k of the adaptive models without a division or a search:
//   mean = floor(N / c), k = mean <= 1 ? 0 : min(31, bit_width(mean - 1))
// With M = N - c >= c:  bit_width(mean - 1) = 1 + max{ s : (M >> s) >= c }, and that s is
// bit_width(M) - bit_width(c), minus one when the shifted value falls short of c.
__device__ __forceinline__ uint32_t kbase_clz(u64 N, uint32_t c) {
  if (N < 2ull * c) return 0u;
  const u64 M = N - c;
  const uint32_t s0 = bitwidth64(M) - (32u - (uint32_t)__clz((int)c));
  const uint32_t t = (uint32_t)(M >> s0);
  const uint32_t kb = 1u + s0 - (t < c ? 1u : 0u);
  return kb > 31u ? 31u : kb;
}

// the same for N < 2^31 (sums of a whole block of 24-bit audio residuals stay far below), all 32-bit
__device__ __forceinline__ uint32_t kbase_clz32(uint32_t N, uint32_t c) {
  if (N < 2u * c) return 0u;
  const uint32_t M = N - c;
  const uint32_t s0 = (uint32_t)__clz((int)c) - (uint32_t)__clz((int)M);  // bit_width(M) - bit_width(c), M >= c
  return 1u + s0 - ((M >> s0) < c ? 1u : 0u);                             // <= 31 because M < 2^31
}

// rice_bits_for_unsigned, block/encoder.cpp:67-70
__device__ __forceinline__ u64 rice_cost(uint32_t u, uint32_t k) {
  const uint32_t q = (k >= 31u) ? 0u : (u >> k);
  return (u64)q + 1u + k;
}

// ---------------------------------------------------------------------------
// Block-wide primitives.  `scratch` holds 33 entries per scanned value; every call
// ends with a barrier so the scratch can be reused immediately.
// Sub-blocks.  A 256-sample stereo probe is analysed by ONE warp (NT == 32), and many probe warps share a CTA so that
// they can be kept in step through the code (k_analyze): inside NT == 32 code the "thread index" is the lane, a "block
// barrier" is a warp barrier and a "block vote" a warp vote.  Every other NT is a real CTA.
template <int NT>
__device__ __forceinline__ uint32_t blk_tid() { return NT == 32 ? (threadIdx.x & 31u) : threadIdx.x; }
template <int NT>
__device__ __forceinline__ void blk_sync() {
  if (NT == 32) __syncwarp();
  else __syncthreads();
}
template <int NT>
__device__ __forceinline__ int blk_sync_or(int p) {
  if (NT == 32) return __any_sync(kFull, p);
  return __syncthreads_or(p);
}
#define LACB_TID (::lacb::blk_tid<NT>())
#define LACB_SYNC() ::lacb::blk_sync<NT>()
#define LACB_SYNC_OR(p) ::lacb::blk_sync_or<NT>(p)

template <int NT>
__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64* scratch, u64* total) {
  const uint32_t tid = LACB_TID, lane = tid & 31u, w = tid >> 5;
  constexpr int NW = (NT + 31) / 32;
  u64 inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 y = __shfl_up_sync(kFull, inc, d);
    if (lane >= (uint32_t)d) inc += y;
  }
  if (NW == 1) {
    if (total) *total = __shfl_sync(kFull, inc, 31);
    return inc - v;
  }
  if (lane == 31u) scratch[w] = inc;
  LACB_SYNC();
  u64 ws = (lane < (uint32_t)NW) ? scratch[lane] : 0ull;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 y = __shfl_up_sync(kFull, ws, d);
    if (lane >= (uint32_t)d) ws += y;
  }
  const u64 wprev = __shfl_sync(kFull, ws, (int)((w + 31u) & 31u));
  const u64 tot = __shfl_sync(kFull, ws, NW - 1);
  LACB_SYNC();
  if (total) *total = tot;
  return (w ? wprev : 0ull) + inc - v;
}

template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* scratch, uint32_t* total) {
  const uint32_t tid = LACB_TID, lane = tid & 31u, w = tid >> 5;
  constexpr int NW = (NT + 31) / 32;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFull, inc, d);
    if (lane >= (uint32_t)d) inc += y;
  }
  if (NW == 1) {
    if (total) *total = __shfl_sync(kFull, inc, 31);
    return inc - v;
  }
  if (lane == 31u) scratch[w] = inc;
  LACB_SYNC();
  uint32_t ws = (lane < (uint32_t)NW) ? scratch[lane] : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFull, ws, d);
    if (lane >= (uint32_t)d) ws += y;
  }
  const uint32_t wprev = __shfl_sync(kFull, ws, (int)((w + 31u) & 31u));
  const uint32_t tot = __shfl_sync(kFull, ws, NW - 1);
  LACB_SYNC();
  if (total) *total = tot;
  return (w ? wprev : 0u) + inc - v;
}

// Exclusive sum scan and exclusive running maximum (identity -1) of one value pair per thread
// in a single pass: one barrier between the warp level and the block level.  `scratch` holds
// 48 entries.  TRAIL adds the closing barrier that protects the scratch; callers that reach
// another barrier before the scratch is written again pass false.
template <int NT, bool TRAIL>
__device__ __forceinline__ void block_scan_sum_max(u64 v, int32_t m, u64* scratch, u64* ex_sum, u64* total,
                                                   int32_t* ex_max) {
  const uint32_t tid = LACB_TID, lane = tid & 31u, w = tid >> 5;
  constexpr int NW = (NT + 31) / 32;
  u64 inc = v;
  int32_t im = m;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 y = __shfl_up_sync(kFull, inc, d);
    const int32_t ym = __shfl_up_sync(kFull, im, d);
    if (lane >= (uint32_t)d) {
      inc += y;
      if (ym > im) im = ym;
    }
  }
  int32_t exm = __shfl_up_sync(kFull, im, 1);
  if (lane == 0u) exm = -1;
  if (NW == 1) {
    *total = __shfl_sync(kFull, inc, 31);
    *ex_sum = inc - v;
    *ex_max = exm;
    LACB_SYNC();  // callers rely on one barrier inside the scan
    return;
  }
  int32_t* smax = reinterpret_cast<int32_t*>(scratch + 32);
  if (lane == 31u) {
    scratch[w] = inc;
    smax[w] = im;
  }
  LACB_PH(2);
  LACB_SYNC();
  LACB_PH(3);
  u64 ws = (lane < (uint32_t)NW) ? scratch[lane] : 0ull;
  int32_t wm = (lane < (uint32_t)NW) ? smax[lane] : -1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 y = __shfl_up_sync(kFull, ws, d);
    const int32_t ym = __shfl_up_sync(kFull, wm, d);
    if (lane >= (uint32_t)d) {
      ws += y;
      if (ym > wm) wm = ym;
    }
  }
  const u64 wprev = __shfl_sync(kFull, ws, (int)((w + 31u) & 31u));
  const int32_t wmprev = __shfl_sync(kFull, wm, (int)((w + 31u) & 31u));
  *total = __shfl_sync(kFull, ws, NW - 1);
  if (TRAIL) LACB_SYNC();
  *ex_sum = (w ? wprev : 0ull) + inc - v;
  if (w && wmprev > exm) exm = wmprev;
  *ex_max = exm;
}

// Warp sums with the REDUX unit (one instruction per 32-bit reduction instead of five
// shuffle + add steps).  The 64-bit form splits the value at bit 24: both halves of the
// per-lane values summed here (bit costs, sums of zig-zag values over E samples) are far
// below 2^51, so neither partial sum can wrap.
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) { return __reduce_add_sync(kFull, v); }
__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
  const uint32_t lo = __reduce_add_sync(kFull, (uint32_t)v & 0xFFFFFFu);
  const uint32_t hi = __reduce_add_sync(kFull, (uint32_t)(v >> 24));
  return ((u64)hi << 24) + lo;
}
// full-width (wrapping) 64-bit warp sum, for values that use all 64 bits
__device__ __forceinline__ u64 warp_sum_u64_full(u64 v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
  return v;
}

// 64-bit sums in shared memory fed by native 32-bit atomics (a 64-bit shared-memory atomicAdd is a compare-and-swap
// loop, and the block totals are hit by all 32 warps at once): word 0 of the slot collects the low 24 bits of every
// addend, word 1 the rest.  Good for at most 256 addends below 2^56 per slot; a zeroed u64 is an empty slot.
__device__ __forceinline__ void split_sum_add(u64* slot, u64 v) {
  uint32_t* w = reinterpret_cast<uint32_t*>(slot);
  atomicAdd(&w[0], (uint32_t)v & 0xFFFFFFu);
  atomicAdd(&w[1], (uint32_t)(v >> 24));
}
__device__ __forceinline__ u64 split_sum_get(const u64* slot) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(slot);
  return ((u64)w[1] << 24) + w[0];
}

// ---------------------------------------------------------------------------
// Bit-plane population counts.
//
// sum_i (u_i >> k) for every k follows from the per-bit-plane set counts c_b:
//   T_0 = sum u,   T_{k+1} = (T_k - c_k) >> 1          (exact, any 32-bit u)
// which turns estimate_static_k / estimate_initial_k (block/encoder.cpp:121-188,
// ~57% of the CPU encoder's time) into a handful of logic ops per sample.
// Counts are kept packed two planes per 32-bit word (16-bit fields; a block has at
// most 16384 samples): word w = c_{2w} | c_{2w+1} << 16, planes 0..15.
struct PlaneCounts {
  uint32_t w[8];
};

// Vertical (bit-sliced) population count of E words with carry-save adders:
// bit b of V[i] is bit i of the number of inputs that have bit b set.
template <int E>
__device__ __forceinline__ void csa_count(const uint32_t (&u)[E], uint32_t (&V)[5]) {
  static_assert(E == 8 || E == 16, "E must be 8 or 16");
#define LACB_FA(a, b, c, s, cy)            \
  {                                        \
    const uint32_t _a = (a), _b = (b), _c = (c); \
    s = _a ^ _b ^ _c;                      \
    cy = (_a & _b) | (_c & (_a ^ _b));     \
  }
  if constexpr (E == 16) {
    uint32_t s0, s1, s2, s3, s4, s5, s6, c0, c1, c2, c3, c4, c5, c6;
    LACB_FA(u[0], u[1], u[2], s0, c0);
    LACB_FA(u[3], u[4], u[5], s1, c1);
    LACB_FA(u[6], u[7], u[8], s2, c2);
    LACB_FA(u[9], u[10], u[11], s3, c3);
    LACB_FA(u[12], u[13], u[14], s4, c4);
    LACB_FA(s0, s1, s2, s5, c5);
    LACB_FA(s3, s4, u[15], s6, c6);
    V[0] = s5 ^ s6;
    const uint32_t c7 = s5 & s6;
    uint32_t t0, t1, t2, d0, d1, d2;
    LACB_FA(c0, c1, c2, t0, d0);
    LACB_FA(c3, c4, c5, t1, d1);
    LACB_FA(c6, c7, t0, t2, d2);
    V[1] = t1 ^ t2;
    const uint32_t d3 = t1 & t2;
    uint32_t e0, f0;
    LACB_FA(d0, d1, d2, e0, f0);
    V[2] = e0 ^ d3;
    const uint32_t f1 = e0 & d3;
    V[3] = f0 ^ f1;
    V[4] = f0 & f1;
  } else {
    uint32_t s0, s1, c0, c1, s2, c2;
    LACB_FA(u[0], u[1], u[2], s0, c0);
    LACB_FA(u[3], u[4], u[5], s1, c1);
    LACB_FA(s0, s1, u[6], s2, c2);
    V[0] = s2 ^ u[7];
    const uint32_t c3 = s2 & u[7];
    uint32_t t0, d0;
    LACB_FA(c0, c1, c2, t0, d0);
    V[1] = t0 ^ c3;
    const uint32_t d1 = t0 & c3;
    V[2] = d0 ^ d1;
    V[3] = d0 & d1;
    V[4] = 0u;
  }
#undef LACB_FA
}

// bit-sliced counters -> packed per-plane counts (planes 0..15)
__device__ __forceinline__ void planes_from_sliced(const uint32_t (&V)[5], PlaneCounts& pc) {
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    uint32_t lo = 0u, hi = 0u;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      lo |= ((V[i] >> (2 * w)) & 1u) << i;
      hi |= ((V[i] >> (2 * w + 1)) & 1u) << i;
    }
    pc.w[w] = lo | (hi << 16);
  }
}
__device__ __forceinline__ uint32_t plane_get(const PlaneCounts& pc, int b) {
  return (pc.w[b >> 1] >> ((b & 1) * 16)) & 0xFFFFu;
}
// adds the bit planes of one value
__device__ __forceinline__ void planes_add_value(PlaneCounts& pc, uint32_t u) {
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const uint32_t t = u >> (2 * w);
    pc.w[w] += (t & 1u) + ((t & 2u) << 15);
  }
}

// argmin_k (T_k + cnt*(1+k)) for k in [0,kmax], first minimum wins
// (estimate_initial_k kmax=12, estimate_static_k kmax=15; block/encoder.cpp:121-180).
__device__ __forceinline__ uint32_t best_static_k(u64 T0, const PlaneCounts& pc, uint32_t cnt, int kmax,
                                                  u64* bits_out) {
  u64 T = T0, best = ~0ull;
  uint32_t bk = 0;
  for (int k = 0; k <= kmax; ++k) {
    const u64 cost = T + (u64)cnt * (uint32_t)(1 + k);
    if (cost < best) {
      best = cost;
      bk = (uint32_t)k;
    }
    T = (T - plane_get(pc, k)) >> 1;
  }
  if (bits_out) *bits_out = best;
  return bk;
}

// The same argmin evaluated by a whole warp, one k per lane (every lane gets the result):
// T_k = (T_0 - sum_{b<k} c_b 2^b) >> k, the subtracted sum being an exclusive prefix over the
// lanes (below 2^30: a block has at most 2^14 samples and only planes 0..15 are counted).
// `words` = the eight packed plane-count words.
__device__ __forceinline__ uint32_t warp_best_static_k(u64 T0, const uint32_t* words, uint32_t cnt, int kmax,
                                                       u64* bits_out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t c = lane < 16u ? ((words[lane >> 1] >> ((lane & 1u) * 16u)) & 0xFFFFu) : 0u;
  const uint32_t part = lane < 16u ? (c << lane) : 0u;
  uint32_t inc = part;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFull, inc, d);
    if (lane >= (uint32_t)d) inc += y;
  }
  const u64 T = (T0 - (u64)(inc - part)) >> (lane & 15u);
  u64 key = ((T + (u64)cnt * (lane + 1u)) << 5) | lane;  // cost < 2^47
  if (lane > (uint32_t)kmax) key = ~0ull;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const u64 y = __shfl_xor_sync(kFull, key, d);
    if (y < key) key = y;
  }
  if (bits_out) *bits_out = key >> 5;
  return (uint32_t)key & 31u;
}

}  // namespace lacb
