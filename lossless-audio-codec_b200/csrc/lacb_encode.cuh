// lacb_encode.cuh -- channel-block analysis and emission for the LAC encoder.
//
// One CTA encodes one channel-block (<= 16384 samples).  Thread t owns the E
// consecutive samples [t*E, t*E+E); every per-sample quantity of the reference's
// serial estimators is rewritten as a block-wide scan so nothing but the 12-step
// Levinson recursion (done by a separate kernel) is sequential:
//   * Block::Encoder::encode          src/codec/block/encoder.cpp:313-838
//   * estimate_initial_k/static_k     src/codec/block/encoder.cpp:121-188  (bit-plane counts)
//   * estimate_residual_costs         src/codec/block/encoder.cpp:201-263  (closed-form k series)
//   * Rice::adapt_k                   src/codec/rice/rice.hpp:45-114       
//   * LPC::compute_residual_q15       src/codec/lpc/lpc.cpp:38-61,188-229
#pragma once
#include "lacb_common.cuh"

namespace lacb {

// PCM planes resident in device memory.
struct PcmSrc {
  const int32_t* L;
  const int32_t* R;  // nullptr for mono
  u64 frames;
};

// Quantised LPC analysis of one channel-block (output of the Levinson kernel).
struct LpcQ {
  int16_t coef[5][13];  // candidate c = order 4+2c, coefficients 1..12 (index 0 unused)
  int8_t used[5];       // analyze_block_q15's used_order (0 = unstable)
  int8_t pad[5];
};

// Decision record of one channel-block: everything the emitter needs.
struct ChanRec {
  uint32_t bytes;      // encoded size of this channel-block (byte padded)
  uint32_t bits;       // exact bit count before padding
  uint8_t type;        // 0 fixed, 1 FIR, 2 LPC
  uint8_t order;       // order byte as written
  uint8_t p;           // partition order
  uint8_t taps;        // taps actually used for the residual (LPC: used_order)
  int16_t coef[13];    // Q15 coefficients 1..order (LPC)
  uint8_t part[256];   // (mode << 5) | k per partition
  uint16_t pad;
  uint32_t cand_lo[11];  // debug: low 32 bits of every candidate's best_bits (0xFFFFFFFF = skipped)
  // decision dump (lacb_last_encode_decisions): the winner's estimates and the total of every partition level
  uint32_t lvl_bits[9];  // byte-rounded total of level p (block/encoder.cpp:527-529); [0] = unpartitioned; 0 = not evaluated
  uint8_t base_mode, has_run, pad2[2];
  u64 est_best, est_rice, est_zr, est_bin, est_stat;  // block/encoder.cpp:824-835, 457-466
};

__device__ __forceinline__ int32_t load_sample(const PcmSrc& s, int kind, u64 idx) {
  // kind: 0 left/mono, 1 right, 2 mid, 3 side (simd/neon.cpp:14-30: wrapping add, arithmetic >> 1)
  const int32_t l = s.L[idx];
  if (kind == 0) return l;
  const int32_t r = s.R[idx];
  if (kind == 1) return r;
  if (kind == 2) return (int32_t)((uint32_t)l + (uint32_t)r) >> 1;
  return (int32_t)((uint32_t)l - (uint32_t)r);
}

// partition geometry, block/encoder.cpp:93-119
__device__ __forceinline__ uint32_t max_partition_order(uint32_t n) {
  uint32_t mp = 0;
  for (uint32_t p = 1; p <= kMaxPartOrder; ++p) {
    if ((n >> p) < kMinPart) break;
    mp = p;
  }
  return mp;
}

// Best predictor candidate so far (block-uniform; kept in shared memory, thread 0 updates it).
struct BestCand {
  u64 rice, zr, bin, stat, best;
  uint32_t type, order, taps, ci, k_init, k_stat, has_run;
  uint32_t have;
};

struct AMisc {
  BestCand best;
  u64 tot_rice, tot_zr, tot_bin, u_total, p_first, stat_bits, red64;
  uint32_t cnt_tot[8], cnt_first[8];
  uint32_t k_init, k_stat;  // initial / static k of the candidate (block_static_k)
  uint32_t cand_lb[11];    // exact lower bound of every candidate's cost (pre-pass of k_analyze)
  uint32_t cand_n;         // candidates that exist
  uint8_t cand_order[12];  // evaluation order: ascending (bound, index)
  uint32_t hq_n, hq_kb_n;  // entries in the hard-chunk queue (bias pairs / base-k pairs)
  uint32_t hasrun_bits[8];
  uint32_t hasrun_all[16];  // fused levels: has-run bit of every segment, indexed by table id
  u64 lvl_total[9];         // fused levels: byte-rounded total of every level
};

template <int NT, int E>
struct ASmem {
  static constexpr uint32_t CAP = NT * E;
  // partition tables are sized for the block capacity: 16384 samples -> levels 0..8 (511 segments),
  // a 256-sample probe -> levels 0..3 (15 segments); small tables let many probe CTAs share an SM
  static constexpr uint32_t MAXP = CAP >= 8192u ? 8u : CAP >= 4096u ? 7u : CAP >= 2048u ? 6u : CAP >= 1024u ? 5u
                                   : CAP >= 512u ? 4u : CAP >= 256u ? 3u : CAP >= 128u ? 2u : CAP >= 64u ? 1u : 0u;
  static constexpr uint32_t MAXSEG = (2u << MAXP) - 1u;
  static constexpr uint32_t FBS = (1u << MAXP) + 2u;  // stride of the three per-segment cost arrays
  static constexpr size_t oX = 0;
  static constexpr size_t oU = oX + (size_t)CAP * 4;
  static constexpr size_t oPthr = oU + (size_t)CAP * 4;
  static constexpr size_t oCpre = oPthr + (((size_t)(NT + 1) * 8 + 15) & ~(size_t)15);
  static constexpr size_t oFlg = oCpre + (size_t)(NT + 1) * 32;
  // The K plane (one byte per sample) shares the Cpre area in full-size layouts, behind FbAll: the plane-count prefix
  // exists only from the winner's prepare() to the segment tables (one barrier later it is dead), the K plane is
  // written by the candidate passes before that and by the per-level / final passes after it.  The 16 KB this saves
  // bring the layout from 215.6 to 199.2 KB, under the 196 KB carve-out: the SM keeps 32 KB of L1 instead of none, and
  // the kernel's register spills (which otherwise go to L2) hit it.
  static constexpr bool kAliasK = NT >= 64;
  static constexpr size_t oKalias = oCpre + ((3 * (size_t)(MAXSEG + 1) * 8 + 15) & ~(size_t)15);
  static_assert(!kAliasK || oKalias + (size_t)NT * E <= oFlg, "K plane must fit inside Cpre behind FbAll");
  static constexpr size_t oKpub = kAliasK ? oKalias : oFlg + (size_t)NT * 4;
  static constexpr size_t oScr = oFlg + (size_t)NT * 4 + (kAliasK ? 0 : (size_t)NT * E);
  static constexpr size_t oSegP = oScr + 64 * 8;
  static constexpr size_t oSegStat = oSegP + (MAXSEG + 1) * 8;
  static constexpr size_t oSelBits = oSegStat + (MAXSEG + 1) * 8;
  static constexpr size_t oFb = oSelBits + (MAXSEG + 1) * 8;
  static constexpr size_t oMisc = oFb + 3 * (size_t)FBS * 8;
  static constexpr size_t oSegK = oMisc + ((sizeof(AMisc) + 15) & ~(size_t)15);
  static constexpr size_t oSelMK = oSegK + (MAXSEG + 1) * 2;
  static constexpr size_t oHardQ = (oSelMK + (MAXSEG + 1) + 15) & ~(size_t)15;
  static constexpr size_t BYTES = oHardQ + (size_t)NT * 2;

  unsigned char* base;
  __device__ int32_t* X() const { return reinterpret_cast<int32_t*>(base + oX); }
  __device__ uint32_t* U() const { return reinterpret_cast<uint32_t*>(base + oU); }
  __device__ u64* Pthr() const { return reinterpret_cast<u64*>(base + oPthr); }
  __device__ PlaneCounts* Cpre() const { return reinterpret_cast<PlaneCounts*>(base + oCpre); }
  // per-segment cost sums of all levels (3 x (MAXSEG + 1) u64): aliases Cpre, which is dead once the segment
  // tables are built
  __device__ u64* FbAll() const { return reinterpret_cast<u64*>(base + oCpre); }
  static_assert(3u * (MAXSEG + 1u) * 8u <= (size_t)(NT + 1) * 32u || NT < 64, "FbAll must fit inside Cpre");
  __device__ uint32_t* Flg() const { return reinterpret_cast<uint32_t*>(base + oFlg); }
  __device__ uint32_t* Kpl() const { return reinterpret_cast<uint32_t*>(base + oKpub); }  // k per sample, 1 byte each
  __device__ u64* Scr() const { return reinterpret_cast<u64*>(base + oScr); }
  __device__ u64* SegP() const { return reinterpret_cast<u64*>(base + oSegP); }
  __device__ u64* SegStat() const { return reinterpret_cast<u64*>(base + oSegStat); }
  __device__ u64* SelBits() const { return reinterpret_cast<u64*>(base + oSelBits); }
  __device__ u64* Fb() const { return reinterpret_cast<u64*>(base + oFb); }
  __device__ AMisc* Misc() const { return reinterpret_cast<AMisc*>(base + oMisc); }
  __device__ uint16_t* SegK() const { return reinterpret_cast<uint16_t*>(base + oSegK); }
  __device__ uint8_t* SelMK() const { return reinterpret_cast<uint8_t*>(base + oSelMK); }
  __device__ uint16_t* HardQ() const { return reinterpret_cast<uint16_t*>(base + oHardQ); }  // chunks awaiting k_bias_pair
};

// ---------------------------------------------------------------------------
// load the channel-block into the swizzled X plane (zero padded to CAP)
__device__ __forceinline__ int32_t combine_sample(int kind, int32_t l, int32_t r) {
  if (kind == 0) return l;
  if (kind == 1) return r;
  if (kind == 2) return (int32_t)((uint32_t)l + (uint32_t)r) >> 1;
  return (int32_t)((uint32_t)l - (uint32_t)r);
}

// Loads the channel-block into the swizzled X plane (zero padded to CAP) and returns the OR of
// the sample magnitudes this thread saw.  All of a thread's 128-bit loads are issued before the
// first one is consumed: with one resident CTA per SM nothing else hides the HBM latency.
template <int NT, int E>
__device__ __forceinline__ uint32_t load_block(const ASmem<NT, E>& sm, const PcmSrc& src, int kind, u64 start,
                                               uint32_t n) {
  int32_t* X = sm.X();
  uint32_t mag = 0u;  // OR of the magnitudes (v >= 0 ? v : ~v) of the samples this thread loaded
  constexpr int Q = E / 4;  // 16-byte chunks per thread
  const bool vec_ok = ((reinterpret_cast<uint64_t>(src.L + start) & 15ull) == 0ull) &&
                      (kind == 0 || (reinterpret_cast<uint64_t>(src.R + start) & 15ull) == 0ull);
  if (vec_ok) {
    const int4* L4 = reinterpret_cast<const int4*>(src.L + start);
    const int4* R4 = reinterpret_cast<const int4*>((kind ? src.R : src.L) + start);
    int4 a[Q], b[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const uint32_t q = (uint32_t)k * NT + LACB_TID;
      a[k] = make_int4(0, 0, 0, 0);
      b[k] = make_int4(0, 0, 0, 0);
      if (q * 4u + 4u <= n) {
        if (kind != 1) a[k] = L4[q];
        if (kind != 0) b[k] = R4[q];
      } else if (q * 4u < n) {  // the chunk straddling the end of the block
        int32_t t[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
        for (uint32_t m = 0; m < 4u; ++m)
          if (q * 4u + m < n) {
            if (kind != 1) t[m] = src.L[start + q * 4u + m];
            if (kind != 0) w[m] = src.R[start + q * 4u + m];
          }
        a[k] = make_int4(t[0], t[1], t[2], t[3]);
        b[k] = make_int4(w[0], w[1], w[2], w[3]);
      }
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const uint32_t q = (uint32_t)k * NT + LACB_TID;
      int4 v;
      v.x = combine_sample(kind, a[k].x, b[k].x);
      v.y = combine_sample(kind, a[k].y, b[k].y);
      v.z = combine_sample(kind, a[k].z, b[k].z);
      v.w = combine_sample(kind, a[k].w, b[k].w);
      mag |= (uint32_t)(v.x ^ (v.x >> 31)) | (uint32_t)(v.y ^ (v.y >> 31)) | (uint32_t)(v.z ^ (v.z >> 31)) |
             (uint32_t)(v.w ^ (v.w >> 31));
      reinterpret_cast<int4*>(X)[swz_chunk(q)] = v;
    }
    return mag;
  }
  for (uint32_t i = LACB_TID; i < ASmem<NT, E>::CAP; i += NT) {
    int32_t v = 0;
    if (i < n) v = load_sample(src, kind, start + i);
    mag |= (uint32_t)(v ^ (v >> 31));
    X[swz(i)] = v;
  }
  return mag;
}

// x[0..11] = the 12 samples before the thread's chunk (zero before the block start),
// x[12..12+E) = the thread's own samples
template <int NT, int E>
__device__ __forceinline__ void load_items(const ASmem<NT, E>& sm, int32_t (&x)[E + 12]) {
  const int4* X4 = reinterpret_cast<const int4*>(sm.X());
  const int q0 = (int)LACB_TID * (E / 4);
#pragma unroll
  for (int c = -3; c < E / 4; ++c) {
    int4 v = make_int4(0, 0, 0, 0);
    if (q0 + c >= 0) v = X4[swz_chunk((uint32_t)(q0 + c))];
    x[12 + 4 * c + 0] = v.x;
    x[12 + 4 * c + 1] = v.y;
    x[12 + 4 * c + 2] = v.z;
    x[12 + 4 * c + 3] = v.w;
  }
}

// Fixed / FIR residuals, block/encoder.cpp:265-309.  The fixed predictors are exact
// in wrapping 32-bit arithmetic (the reference truncates the int64 difference).
// Samples past the block end and the predictor warm-up (the first `order` samples are verbatim) concern at most two
// threads of a block: the main loops carry no per-sample tests, the two threads patch their chunk afterwards.
template <int E>
__device__ __forceinline__ void residual_tail(uint32_t g0, uint32_t n, int32_t (&r)[E]) {
  if (g0 + (uint32_t)E > n) {
#pragma unroll
    for (int j = 0; j < E; ++j)
      if (g0 + j >= n) r[j] = 0;
  }
}
template <int E>
__device__ __forceinline__ void residual_fixed(const int32_t (&x)[E + 12], uint32_t g0, uint32_t n, int order,
                                               int32_t (&r)[E]) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t x0 = (uint32_t)x[12 + j], x1 = (uint32_t)x[11 + j], x2 = (uint32_t)x[10 + j],
                   x3 = (uint32_t)x[9 + j], x4 = (uint32_t)x[8 + j];
    uint32_t v;
    switch (order) {
      case 1: v = x0 - x1; break;
      case 2: v = x0 - 2u * x1 + x2; break;
      case 3: v = x0 - 3u * x1 + 3u * x2 - x3; break;
      case 4: v = x0 - 4u * x1 + 6u * x2 - 4u * x3 + x4; break;
      default: v = x0; break;
    }
    r[j] = (int32_t)v;
  }
  if (g0 == 0u) {
#pragma unroll
    for (int j = 0; j < 4 && j < E; ++j)
      if (j < order) r[j] = x[12 + j];
  }
  residual_tail<E>(g0, n, r);
}
template <int E>
__device__ __forceinline__ void residual_fir(const int32_t (&x)[E + 12], uint32_t g0, uint32_t n, int32_t (&r)[E]) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const i64 p = (3ll * (i64)x[11 + j] - (i64)x[10 + j]) >> 2;
    r[j] = (int32_t)(uint32_t)(u64)((i64)x[12 + j] - p);
  }
  if (g0 == 0u) {
    r[0] = x[12];
    if (E > 1) r[1] = x[13];
  }
  residual_tail<E>(g0, n, r);
}
// LPC residual with `TAPS` Q15 taps, lpc.cpp:38-61; returns true if any residual leaves int32
// CHECK = false: the caller knows every |x| < 2^26, so |x - (sum >> 15)| < 2^26 + 12 * 2^26 * 2^15 / 2^15 < 2^31
// and the int32 range test of compute_residual_q15 cannot fire.
template <int E, int TAPS, bool CHECK>
__device__ __forceinline__ bool residual_lpc_t(const int32_t (&x)[E + 12], uint32_t g0, uint32_t n,
                                               const int16_t* c, int32_t (&r)[E]) {
  int32_t cf[TAPS + 1];
#pragma unroll
  for (int t = 1; t <= TAPS; ++t) cf[t] = c[t];
  bool ovf = false;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    i64 acc = 0;
#pragma unroll
    for (int t = 1; t <= TAPS; ++t) acc = mad_wide(cf[t], x[12 + j - t], acc);
    const i64 d = (i64)x[12 + j] - (acc >> 15);
    if (CHECK && g0 + j < n && (d < -2147483648ll || d > 2147483647ll)) ovf = true;
    r[j] = (int32_t)d;
  }
  residual_tail<E>(g0, n, r);
  return ovf;
}
template <int E, bool CHECK = true>
__device__ __forceinline__ bool residual_lpc(const int32_t (&x)[E + 12], uint32_t g0, uint32_t n,
                                             const int16_t* c, int taps, int32_t (&r)[E]) {
  switch (taps) {
    case 4: return residual_lpc_t<E, 4, CHECK>(x, g0, n, c, r);
    case 6: return residual_lpc_t<E, 6, CHECK>(x, g0, n, c, r);
    case 8: return residual_lpc_t<E, 8, CHECK>(x, g0, n, c, r);
    case 10: return residual_lpc_t<E, 10, CHECK>(x, g0, n, c, r);
    case 12: return residual_lpc_t<E, 12, CHECK>(x, g0, n, c, r);
    default: break;
  }
  // odd / short orders (Levinson stopped early): generic, coefficients past `taps` ignored
  bool ovf = false;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    i64 acc = 0;
#pragma unroll
    for (int t = 1; t <= 12; ++t)
      if (t <= taps) acc = mad_wide((int32_t)c[t], x[12 + j - t], acc);
    const i64 d = (i64)x[12 + j] - (acc >> 15);
    const bool in = g0 + j < n;
    if (in && (d < -2147483648ll || d > 2147483647ll)) ovf = true;
    r[j] = in ? (int32_t)d : 0;
  }
  return ovf;
}

// ---------------------------------------------------------------------------
// Per-residual preparation: zig-zag, U plane, per-thread prefix of u, last-nonzero
// scan and bit-plane counts (totals only, or the full per-thread prefix for the
// partition search when FULL).  Nothing per-sample stays in registers afterwards:
// later phases re-read u from the U plane (4 LDS.128 per thread), which is what keeps
// the 1024-thread CTA inside its 64-register budget without spilling.
template <int NT, int E>
struct Prep {
  u64 Pex;          // sum of u over all samples before this thread's chunk
  int32_t lnz_ex;   // index of the last non-zero residual before the chunk (-1: none)
  uint32_t zmask;   // bit j: sample g0+j exists and its residual is zero
  uint32_t any4;    // block-uniform: the residual contains a run of >= 4 zeros somewhere
  uint32_t cls;     // bit 0: some sample of the chunk has u <= 4; bits 8..13: bit width of the OR of the chunk's u
};

template <int NT, int E>
__device__ __forceinline__ void load_u(const ASmem<NT, E>& sm, uint32_t (&u)[E]) {
  const uint4* U4 = reinterpret_cast<const uint4*>(sm.U());
#pragma unroll
  for (int c = 0; c < E / 4; ++c) {
    const uint4 v = U4[swz_chunk(LACB_TID * (E / 4) + c)];
    u[4 * c] = v.x; u[4 * c + 1] = v.y; u[4 * c + 2] = v.z; u[4 * c + 3] = v.w;
  }
}

template <int NT, int E>
__device__ __forceinline__ uint32_t zero_lookahead(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n);

// LIGHT (the emitter): only what the token walk needs -- U plane, prefix of u, last-non-zero scan, run flag;
// no bit-plane counts and no lower bound.
template <int NT, int E, bool FULL, bool LIGHT = false>
__device__ __forceinline__ void prepare(const ASmem<NT, E>& sm, const int32_t (&r)[E], uint32_t n, Prep<NT, E>& pr) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  AMisc* mi = sm.Misc();
  uint32_t u[E];
  u64 S = 0;
  int32_t lastnz = -1;
  uint32_t zmask = 0, umin = 0xFFFFFFFFu, uor = 0u;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t uu = zz32(r[j]);  // residuals past the block end are 0
    u[j] = uu;
    S += uu;
    uor |= uu;
    umin = uu < umin ? uu : umin;  // a chunk that reaches past the block end reads as "has small values": general walk
    if (uu) lastnz = (int32_t)(g0 + j);
    zmask |= (uu == 0u ? 1u : 0u) << j;
  }
  if (g0 + (uint32_t)E > n) zmask &= g0 < n ? (1u << (n - g0)) - 1u : 0u;  // only samples that exist count as zeros
  pr.zmask = zmask;
  pr.cls = (umin <= 4u ? 1u : 0u) | ((32u - (uint32_t)__clz((int)uor)) << 8);
  uint4* U4 = reinterpret_cast<uint4*>(sm.U());
#pragma unroll
  for (int c = 0; c < E / 4; ++c)
    U4[swz_chunk(tid * (E / 4) + c)] = make_uint4(u[4 * c], u[4 * c + 1], u[4 * c + 2], u[4 * c + 3]);
  if (!FULL && !LIGHT && tid < 8) {
    mi->cnt_tot[tid] = 0u;
    mi->cnt_first[tid] = 0u;
  }
  uint32_t V[5] = {0u, 0u, 0u, 0u, 0u};
  if (!LIGHT) csa_count<E>(u, V);
  // one pass: prefix of u and index of the last non-zero residual before the chunk.  Its
  // barrier also orders the U plane and the zeroed counters before everything below.
  u64 total;
  block_scan_sum_max<NT, FULL>(S, lastnz, sm.Scr(), &pr.Pex, &total, &pr.lnz_ex);
  sm.Pthr()[tid] = pr.Pex;
  if (tid == 0) {
    sm.Pthr()[NT] = total;
    mi->u_total = total;
    if (NT * E <= 256) mi->p_first = total;
  }
  if (NT * E > 256 && g0 == 256u) mi->p_first = pr.Pex;

  PlaneCounts pc;
  if (!LIGHT) planes_from_sliced(V, pc);
  if (LIGHT) {
    // nothing to count
  } else if (FULL) {
    // exclusive prefix of the eight packed plane-count words over the threads, all eight in one two-level
    // scan (two barriers instead of sixteen); the per-warp totals go through the Fb area, idle at this point
    PlaneCounts* Cp = sm.Cpre();
    constexpr int NW = (NT + 31) / 32;
    const uint32_t lane = tid & 31u, wid = tid >> 5;
    uint32_t inc[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      inc[w] = pc.w[w];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, inc[w], d);
        if (lane >= (uint32_t)d) inc[w] += y;
      }
    }
    uint32_t* wtot = reinterpret_cast<uint32_t*>(sm.Fb());  // [warp][8]
    if (NW > 1) {
      if (lane == 31u) {
#pragma unroll
        for (int w = 0; w < 8; ++w) wtot[wid * 8u + w] = inc[w];
      }
      LACB_SYNC();
    }
    PlaneCounts ex;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      uint32_t before = 0u, total = __shfl_sync(kFull, inc[w], 31);
      if (NW > 1) {
        uint32_t ws = lane < (uint32_t)NW ? wtot[lane * 8u + w] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t y = __shfl_up_sync(kFull, ws, d);
          if (lane >= (uint32_t)d) ws += y;
        }
        const uint32_t wprev = __shfl_sync(kFull, ws, (int)((wid + 31u) & 31u));
        before = wid ? wprev : 0u;
        total = __shfl_sync(kFull, ws, NW - 1);
      }
      ex.w[w] = before + inc[w] - pc.w[w];
      if (tid == 0) Cp[NT].w[w] = total;
    }
    Cp[tid] = ex;
    LACB_SYNC();
  } else {
    const bool first = g0 < 256u;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint32_t t = warp_sum_u32(pc.w[w]);
      if ((tid & 31u) == 0u && t) atomicAdd(&mi->cnt_tot[w], t);
      if ((tid >> 5) * 32u * E < 256u) {  // warp-uniform: this warp overlaps the first 256 samples
        const uint32_t f = warp_sum_u32(first ? pc.w[w] : 0u);
        if ((tid & 31u) == 0u && f) atomicAdd(&mi->cnt_first[w], f);
      }
    }
  }
  {
    // any run of >= 4 zeros starting inside this chunk (look-ahead reaches 4 samples into the
    // next chunk)?  Without one, no partition of any level can use a zero-run token and all
    // run bookkeeping is skipped.
    const uint32_t zm = zero_lookahead(sm, pr, n);
    const uint32_t r4 = zm & (zm >> 1) & (zm >> 2) & (zm >> 3) & ((1u << E) - 1u);
    LACB_PH(4);
    pr.any4 = (uint32_t)LACB_SYNC_OR((int)(r4 != 0u));
    LACB_PH(5);
  }
}

// prefix of u / of the plane counts at an arbitrary sample position (any thread)
template <int NT, int E>
__device__ __forceinline__ u64 prefix_u(const ASmem<NT, E>& sm, uint32_t pos) {
  const uint32_t t = pos / E, rem = pos - t * E;
  u64 v = sm.Pthr()[t];
  const uint32_t* U = sm.U();
  for (uint32_t m = 0; m < rem; ++m) v += U[swz(t * E + m)];
  return v;
}
template <int NT, int E>
__device__ __forceinline__ void prefix_counts(const ASmem<NT, E>& sm, uint32_t pos, PlaneCounts& pc) {
  const uint32_t t = pos / E, rem = pos - t * E;
  pc = sm.Cpre()[t];
  const uint32_t* U = sm.U();
  for (uint32_t m = 0; m < rem; ++m) planes_add_value(pc, U[swz(t * E + m)]);
}

// ---------------------------------------------------------------------------
// Segment geometry of one thread at partition level p (p == 0: whole block).
struct SegGeom {
  uint32_t sidA;   // table index of the segment holding the first item
  uint32_t a0;     // its first sample
  uint32_t bnd;    // first sample of the next segment (0xFFFFFFFF when none)
  uint32_t s0;     // segment number inside the level
  bool fast;       // all E samples exist and lie in segment A: per-item bounds checks vanish
};
template <int E>
__device__ __forceinline__ SegGeom seg_geom(uint32_t g0, uint32_t n, uint32_t p) {
  SegGeom g;
  if (p == 0u) {
    g.sidA = 0u;
    g.a0 = 0u;
    g.bnd = 0xFFFFFFFFu;
    g.s0 = 0u;
  } else {
    const uint32_t base = n >> p, cnt = 1u << p;
    uint32_t s0 = g0 / base;
    if (s0 > cnt - 1u) s0 = cnt - 1u;
    g.s0 = s0;
    g.sidA = cnt - 1u + s0;
    g.a0 = s0 * base;
    g.bnd = (s0 + 1u < cnt) ? g.a0 + base : 0xFFFFFFFFu;
  }
  g.fast = (g0 + E <= n) && (g.bnd >= g0 + E);
  return g;
}

// k series of one level.  On return the K plane holds, for every sample i, the Rice
// parameter the model yields AFTER consuming sample i (the k used for sample i+1 unless
// that sample starts a segment).
//   STATEFUL  : Rice::adapt_k with drift and micro windows (rice.hpp:45-114), p = 0
//   !STATEFUL : adapt_k_stateless (block/encoder.cpp:72-77), restarted per segment
// Returns false only with DEFER: the chunk's k moves inside it and fits the 32-bit path; nothing was
// computed and the caller queues the chunk for k_base_pair.
template <int NT, int E, bool STATEFUL, bool FAST, bool DEFER>
__device__ __forceinline__ bool k_series_thread(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n,
                                                const SegGeom& sg, uint32_t (&kpk)[E / 4], uint32_t& flg_out) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  uint32_t u[E];
  load_u<NT, E>(sm, u);
  const u64 PaA = STATEFUL ? 0ull : sm.SegP()[sg.sidA];
  const u64 PaB = (!FAST && !STATEFUL && sg.bnd != 0xFFFFFFFFu) ? sm.SegP()[sg.sidA + 1u] : 0ull;
  u64 Pin = pr.Pex;
  uint32_t flg = 0u;
#pragma unroll
  for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = 0u;
  if (FAST) {
    // One evaluation for the whole chunk when k provably does not change inside it:
    // N_j is non-decreasing and c_j increases by one, so
    //   N_first >= lo * c_last  and  N_last < hi * c_first   (lo = 2^(k-1)+1, hi = 2^k+1)
    // bound every (N_j, c_j) of the chunk inside the k cell of the first sample.
    const u64 S = sm.Pthr()[tid + 1u] - pr.Pex;
    const uint32_t c_first = g0 - sg.a0 + 1u, c_last = c_first + (uint32_t)E - 1u;
    const u64 rel = pr.Pex - PaA;
    const u64 N_first = rel + u[0] + (c_first >> 1), N_last = rel + S + (c_last >> 1);
    const uint32_t kb0 = kbase_clz(N_first, c_first);
    bool uniform;
    if (kb0 == 0u) {
      uniform = N_last < 2ull * c_first;
    } else {
      const u64 lo = (1ull << (kb0 - 1u)) + 1ull, hi = (1ull << kb0) + 1ull;
      uniform = (N_first >= lo * c_last) && (kb0 == 31u || N_last < hi * c_first);
    }
    if (uniform) {
#pragma unroll
      for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = kb0 * 0x01010101u;
      if (STATEFUL) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          const uint32_t q = (kb0 >= 31u) ? 0u : (u[j] >> kb0);
          flg |= ((q > 3u) ? (1u << j) : 0u) | ((q == 0u) ? (1u << (16 + j)) : 0u);
        }
      }
      flg_out = flg;
      return true;
    }
    if (N_last < 0x80000000ull) {
      // k moves inside the chunk (block / segment start, level change): per-sample closed form,
      // in 32-bit arithmetic when the running sum allows it
      if (DEFER) {
        flg_out = 0u;
        return false;
      }
      uint32_t N32 = (uint32_t)rel;
#pragma unroll
      for (int j = 0; j < E; ++j) {
        N32 += u[j];
        const uint32_t c = c_first + (uint32_t)j;
        const uint32_t kb = kbase_clz32(N32 + (c >> 1), c);
        kpk[j >> 2] |= kb << (8 * (j & 3));
        if (STATEFUL) {
          const uint32_t q = u[j] >> kb;
          flg |= ((q > 3u) ? (1u << j) : 0u) | ((q == 0u) ? (1u << (16 + j)) : 0u);
        }
      }
      flg_out = flg;
      return true;
    }
  }
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t idx = g0 + j;
    Pin += u[j];
    const bool inB = !FAST && idx >= sg.bnd;
    const uint32_t a = inB ? sg.bnd : sg.a0;
    const uint32_t c = idx - a + 1u;
    const u64 N = Pin - (inB ? PaB : PaA) + (c >> 1);
    const uint32_t kb = kbase_clz(N, c);
    kpk[j >> 2] |= kb << (8 * (j & 3));
    if (STATEFUL) {
      const uint32_t q = (kb >= 31u) ? 0u : (u[j] >> kb);
      if (FAST || idx < n) flg |= ((q > 3u) ? (1u << j) : 0u) | ((q == 0u) ? (1u << (16 + j)) : 0u);
    }
  }
  flg_out = flg;
  return true;
}

// Base k (and the stateful model's flag words) of two chunks at once, one lane per sample: the
// block-wide counterpart of the per-sample loop above for chunks whose k moves inside them.  Only chunks
// that lie inside one segment and whose running sum stays below 2^31 are sent here.  Writes the k bytes
// of the chunk to the K plane and, for the stateful model, its flag word.
template <int NT, int E, bool STATEFUL>
__device__ __forceinline__ void k_base_pair(const ASmem<NT, E>& sm, uint32_t tA, uint32_t tB, uint32_t n, uint32_t p) {
  const uint32_t lane = LACB_TID & 31u, j = lane & 15u;
  const uint32_t t = lane < 16u ? tA : tB;
  const bool live = t < (uint32_t)NT;
  const uint32_t tc = live ? t : 0u;
  const uint32_t g0 = tc * E;
  const uint32_t u = sm.U()[swz(g0 + j)];
  uint32_t p1 = u;
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const uint32_t y = __shfl_up_sync(kFull, p1, d, 16);
    if (j >= (uint32_t)d) p1 += y;
  }
  const SegGeom sg = seg_geom<E>(g0, n, STATEFUL ? 0u : p);
  const u64 PaA = STATEFUL ? 0ull : sm.SegP()[sg.sidA];
  const uint32_t rel = (uint32_t)(sm.Pthr()[tc] - PaA);
  const uint32_t c = g0 + j - sg.a0 + 1u;
  const uint32_t kb = kbase_clz32(rel + p1 + (c >> 1), c);
  if (live) reinterpret_cast<uint8_t*>(sm.Kpl())[g0 + j] = (uint8_t)kb;
  if (STATEFUL) {
    const uint32_t q = u >> kb;
    const uint32_t bl = __ballot_sync(kFull, q > 3u), bz = __ballot_sync(kFull, q == 0u);
    const uint32_t sh = lane & 16u;
    if (live && j == 0u) sm.Flg()[tc] = ((bl >> sh) & 0xFFFFu) | (((bz >> sh) & 0xFFFFu) << 16);
  }
}

// 4 mask bits -> 4 bytes of 0 / 1
__device__ __forceinline__ uint32_t expand_nibble(uint32_t m) { return ((m & 0xFu) * 0x00204081u) & 0x01010101u; }
// 0x01 in every byte of a packed k word (bytes <= 31) that is not zero
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t w) { return ((w + 0x7F7F7F7Fu) >> 7) & 0x01010101u; }

// Bias of the stateful model (Rice::adapt_k, rice.hpp:60-112) on top of the base k series.
//
// Most chunks never need the per-sample recurrences:
//  * drift rule (256-sample window mean against the running mean): with the window sum
//    bounded by [W0 - S_leave, W0 + S_own] and (N, c) bounded by their values at the two
//    ends of the chunk, one pair of comparisons proves the rule yields the same value
//    (0, +1 or -1) for all E samples;
//  * micro rule (flags of the last 96 samples): the counts at the start of the chunk are
//    exact sums of whole threads' flag words; the "large" rule is decided for the whole
//    chunk by its extreme counts, the "zero" rule, which hovers around its threshold on
//    stationary signals, by a 16-step count over two flag words.
// The bias is then applied to the four packed k words with byte-parallel arithmetic.
// Chunks where a bound fails (level changes, block start) run the exact per-sample loop.
// Returns true when kpk holds the biased k of the chunk.  With SLOW = false a chunk that needs the
// per-sample evaluation is left untouched and false is returned (the caller queues it for k_bias_pair).
template <int NT, int E, bool SLOW>
__device__ __forceinline__ bool k_bias_thread(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t flg,
                                              uint32_t (&kpk)[E / 4]) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  const uint32_t* Flg = sm.Flg();
  constexpr int D = (int)kMicroWin / E;       // threads spanned by the 96-sample micro window
  constexpr int DW = (int)kDriftWin / E;      // threads spanned by the 256-sample drift window
  static_assert(E <= 16, "flag words hold 16 + 16 bits");
  uint32_t fullL = 0u, fullZ = 0u;
#pragma unroll
  for (int d = 1; d < D; ++d) {
    const uint32_t w = ((int)tid - d >= 0) ? Flg[tid - d] : 0u;
    fullL += (uint32_t)__popc(w & 0xFFFFu);
    fullZ += (uint32_t)__popc(w >> 16);
  }
  const uint32_t part = ((int)tid - D >= 0) ? Flg[tid - D] : 0u;
  const int tt = (int)tid - DW;
  u64 wprev = tt >= 0 ? sm.Pthr()[tt] : 0ull;  // becomes the inclusive prefix at item j of thread tt
  const uint32_t c_first = g0 + 1u, c_last = g0 + (uint32_t)E;
  // exact window counts over the 96 samples before item 0
  const uint32_t L0 = fullL + (uint32_t)__popc(part & 0xFFFFu), Z0 = fullZ + (uint32_t)__popc(part >> 16);
  {
    bool hard = false;
    int drift = 0;
    // Chunks in which a rule switches on (sample counts 96 and 256 are multiples of E, so it
    // switches on at the last sample): that one sample is evaluated exactly on its own.
    bool patch_drift = false, patch_micro = false;
    int drift_last = 0, micro_last = 0;
    if (c_last >= kDriftWin) {
      if (c_first < kDriftWin) {
        patch_drift = true;  // c_last == 256: the window is everything so far
        const u64 Pin = sm.Pthr()[tid + 1u];
        const u64 N = Pin + (c_last >> 1);
        if (N >= (u64)c_last) {
          const u64 lm = (Pin - sm.Pthr()[tt + 1] + 128ull) >> 8;
          const u64 tA = (3ull * lm + 3ull) >> 2, tB = lm + 2ull + lm / 3ull;
          if (N < tA * c_last) drift_last = 1;
          else if (N >= tB * c_last) drift_last = -1;
        }
      } else {
        const u64 S_own = sm.Pthr()[tid + 1u] - pr.Pex;
        const u64 S_leave = sm.Pthr()[tt + 1] - wprev;
        const u64 W0 = pr.Pex - wprev;
        const u64 lm_max = (W0 + S_own + 128ull) >> 8, lm_min = (W0 - S_leave + 128ull) >> 8;
        const u64 N_lo = pr.Pex + (c_first >> 1), N_hi = pr.Pex + S_own + (c_last >> 1);
        const u64 tA_max = (3ull * lm_max + 3ull) >> 2, tB_min = lm_min + 2ull + lm_min / 3ull;
        if ((N_lo >= tA_max * c_last && N_hi < tB_min * c_first) || N_hi < (u64)c_first) {
          drift = 0;  // inside the dead band everywhere, or the running mean is 0 (rule off) everywhere
        } else {
          const u64 tA_min = (3ull * lm_min + 3ull) >> 2, tB_max = lm_max + 2ull + lm_max / 3ull;
          if (N_hi < tA_min * c_first && N_lo >= (u64)c_last) drift = 1;
          else if (N_lo >= tB_max * c_last) drift = -1;
          else hard = true;
        }
      }
    }
    bool l_on = false;
    uint32_t mz = 0u;  // samples where the zero rule fires (and the large rule does not)
    if (!hard && c_last >= kMicroWin) {
      const uint32_t fl = flg & 0xFFFFu, fz = flg >> 16, pl = part & 0xFFFFu, pz = part >> 16;
      if (c_first < kMicroWin) {
        patch_micro = true;  // c_last == 96
        const uint32_t Ll = L0 + (uint32_t)__popc(fl) - (uint32_t)__popc(pl);
        const uint32_t Zl = Z0 + (uint32_t)__popc(fz) - (uint32_t)__popc(pz);
        micro_last = Ll >= 72u ? 1 : (Zl >= 77u ? -1 : 0);
      } else {
        if (L0 + (uint32_t)__popc(fl) < 72u) l_on = false;
        else if (L0 >= 72u + (uint32_t)__popc(pl)) l_on = true;
        else hard = true;
        if (!hard && !l_on) {
          if (Z0 + (uint32_t)__popc(fz) < 77u) {
            mz = 0u;
          } else if (Z0 >= 77u + (uint32_t)__popc(pz)) {
            mz = 0xFFFFu;
          } else {
            uint32_t z = Z0;
#pragma unroll
            for (int j = 0; j < E; ++j) {
              z += ((fz >> j) & 1u) - ((pz >> j) & 1u);
              mz |= (z >= 77u ? 1u : 0u) << j;
            }
          }
        }
      }
    }
    if (!hard) {
      // bias = uniform part bu, minus one where dm is set
      int bu;
      uint32_t dm;
      if (l_on) { bu = drift + 1 < 1 ? drift + 1 : 1; dm = 0u; }
      else if (drift < 0) { bu = -1; dm = 0u; }
      else { bu = drift; dm = mz; }
      const uint32_t kb_last = kpk[E / 4 - 1] >> 24;
      if (bu > 0) {
        // k + 1 must stay <= 31 in every byte, else the exact loop does the clamping
#pragma unroll
        for (int c4 = 0; c4 < E / 4; ++c4) hard = hard || (((kpk[c4] + 0x61616161u) & 0x80808080u) != 0u);
      }
      if (!hard) {
#pragma unroll
        for (int c4 = 0; c4 < E / 4; ++c4) {
          uint32_t w = kpk[c4];
          if (bu > 0) w += 0x01010101u;
          else if (bu < 0) w -= nonzero_bytes(w);
          if (dm) w -= expand_nibble(dm >> (4 * c4)) & nonzero_bytes(w);
          kpk[c4] = w;
        }
        if (patch_drift || patch_micro) {
          int b;
          if (patch_micro) {
            b = micro_last;  // the drift rule is still off at sample 96
          } else {
            const bool z_last = (mz >> (E - 1)) & 1u;
            b = l_on ? (drift_last + 1 < 1 ? drift_last + 1 : 1)
                     : (z_last ? (drift_last - 1 > -1 ? drift_last - 1 : -1) : drift_last);
          }
          int k = (int)kb_last + b;
          k = k < 0 ? 0 : (k > 31 ? 31 : k);
          kpk[E / 4 - 1] = (kpk[E / 4 - 1] & 0x00FFFFFFu) | ((uint32_t)k << 24);
        }
        return true;
      }
    }
  }
  if (!SLOW) return false;
  uint32_t u[E];
  load_u<NT, E>(sm, u);
  const uint4* U4 = reinterpret_cast<const uint4*>(sm.U());
  u64 Pin = pr.Pex;
  uint32_t uw[4] = {0u, 0u, 0u, 0u};
  // sliding micro-window counts: start with the window that ends just before item 0
  uint32_t L = L0, Z = Z0;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t c = g0 + j + 1u;
    Pin += u[j];
    if ((j & 3) == 0) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (tt >= 0) v = U4[swz_chunk((uint32_t)tt * (E / 4) + (uint32_t)(j >> 2))];
      uw[0] = v.x; uw[1] = v.y; uw[2] = v.z; uw[3] = v.w;
    }
    wprev += uw[j & 3];
    // item j enters the 96-sample window, item j of thread tid-D leaves it
    L += ((flg >> j) & 1u) - ((part >> j) & 1u);
    Z += ((flg >> (16 + j)) & 1u) - ((part >> (16 + j)) & 1u);
    const u64 N = Pin + (c >> 1);
    const uint32_t kb = (kpk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
    int bias = 0;
    if (c >= kDriftWin && N >= (u64)c) {
      const u64 ws = Pin - wprev;          // sum of the last 256 u (wprev == 0 when tt < 0, c == 256)
      const u64 lm = (ws + 128ull) >> 8;
      const u64 tA = (3ull * lm + 3ull) >> 2;
      if (N < tA * c) {
        bias = 1;
      } else {
        const u64 tB = lm + 2ull + lm / 3ull;  // floor((4 lm + 3) / 3) + 1
        if (N >= tB * c) bias = -1;
      }
    }
    if (c >= kMicroWin) {
      if (L * 4u >= 288u) bias = bias + 1 < 1 ? bias + 1 : 1;
      else if (Z * 5u >= 384u) bias = bias - 1 > -1 ? bias - 1 : -1;
    }
    int k = (int)kb + bias;
    k = k < 0 ? 0 : (k > 31 ? 31 : k);
    kpk[j >> 2] = (kpk[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | ((uint32_t)k << (8 * (j & 3)));
  }
  return true;
}

// The exact per-sample bias of two chunks at once, one lane per sample (lanes 0..15: chunk tA, lanes
// 16..31: chunk tB; an id >= NT means "none").  Chunks that k_bias_thread could not settle with its
// chunk-level proofs are collected block-wide and dealt to the warps in pairs, so the cost of the hard
// chunks (level changes, block start) is spread over all 32 warps instead of stalling the barrier behind
// the few warps that own them.  Reads the base k bytes of the chunk from the K plane and replaces them by
// the biased k; same arithmetic as the per-sample loop above (rice.hpp:60-112).
template <int NT, int E>
__device__ __forceinline__ void k_bias_pair(const ASmem<NT, E>& sm, uint32_t tA, uint32_t tB) {
  constexpr uint32_t D = kMicroWin / E, DW = kDriftWin / E;
  const uint32_t lane = LACB_TID & 31u, j = lane & 15u;
  const uint32_t t = lane < 16u ? tA : tB;
  const bool live = t < (uint32_t)NT;
  const uint32_t tc = live ? t : 0u;  // idle half-warps follow along on chunk 0 and drop the result
  const uint32_t g0 = tc * E;
  const uint32_t* U = sm.U();
  const uint32_t u = U[swz(g0 + j)];
  const bool has_w = tc >= DW;
  const uint32_t u2 = has_w ? U[swz((tc - DW) * E + j)] : 0u;
  u64 p1 = u, p2 = u2;  // inclusive prefixes inside the chunk and inside the chunk leaving the drift window
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const u64 y1 = __shfl_up_sync(kFull, p1, d, 16), y2 = __shfl_up_sync(kFull, p2, d, 16);
    if (j >= (uint32_t)d) {
      p1 += y1;
      p2 += y2;
    }
  }
  const u64 Pin = sm.Pthr()[tc] + p1;
  const u64 wprev = has_w ? sm.Pthr()[tc - DW] + p2 : 0ull;
  const uint32_t c = g0 + j + 1u;
  const u64 N = Pin + (c >> 1);
  uint8_t* Kb = reinterpret_cast<uint8_t*>(sm.Kpl());
  const uint32_t kb = Kb[g0 + j];
  int bias = 0;
  if (c >= kDriftWin && N >= (u64)c) {
    const u64 lm = (Pin - wprev + 128ull) >> 8;
    const u64 tA2 = (3ull * lm + 3ull) >> 2;
    if (N < tA2 * c) {
      bias = 1;
    } else {
      const u64 tB2 = lm + 2ull + lm / 3ull;
      if (N >= tB2 * c) bias = -1;
    }
  }
  if (c >= kMicroWin) {
    const uint32_t* Flg = sm.Flg();
    uint32_t L = 0u, Z = 0u;
#pragma unroll
    for (uint32_t d = 1; d <= D; ++d) {  // the 96 samples before item 0: threads t-6 .. t-1
      const uint32_t w = tc >= d ? Flg[tc - d] : 0u;
      L += (uint32_t)__popc(w & 0xFFFFu);
      Z += (uint32_t)__popc(w >> 16);
    }
    const uint32_t own = Flg[tc], part = tc >= D ? Flg[tc - D] : 0u;
    const uint32_t m = (2u << j) - 1u;  // items 0..j have entered, the same items of thread t-6 have left
    L += (uint32_t)__popc(own & m) - (uint32_t)__popc(part & m);
    Z += (uint32_t)__popc((own >> 16) & m) - (uint32_t)__popc((part >> 16) & m);
    if (L * 4u >= 288u) bias = bias + 1 < 1 ? bias + 1 : 1;
    else if (Z * 5u >= 384u) bias = bias - 1 > -1 ? bias - 1 : -1;
  }
  int k = (int)kb + bias;
  k = k < 0 ? 0 : (k > 31 ? 31 : k);
  if (live) Kb[g0 + j] = (uint8_t)k;
}

// Initial / static k of the whole block from the block totals (one k per lane), for the candidate passes.
// Only thread 0 (first sample of the walk) and the bookkeeping thread consume them, two barriers later.
template <int NT, int E>
__device__ __forceinline__ void block_static_k(const ASmem<NT, E>& sm, uint32_t n) {
  AMisc* mi = sm.Misc();
  u64 sb;
  const uint32_t ki = warp_best_static_k(mi->p_first, mi->cnt_first, n < 256u ? n : 256u, 12, nullptr);
  const uint32_t ks = warp_best_static_k(mi->u_total, mi->cnt_tot, n, 15, &sb);
  if ((LACB_TID & 31u) == 0u) {
    mi->k_init = ki;
    mi->k_stat = ks;
    mi->stat_bits = sb;
  }
}

// Appends this thread's chunk to the block-wide queue of chunks that need per-sample work
// (one shared-memory atomic per warp).  Warp collective.
template <int NT, int E>
__device__ __forceinline__ void queue_push(const ASmem<NT, E>& sm, uint32_t* counter, bool want) {
  const uint32_t lane = LACB_TID & 31u;
  const uint32_t m = __ballot_sync(kFull, want);
  if (!m) return;
  uint32_t base = 0u;
  if (lane == 0u) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(kFull, base, 0);
  if (want) sm.HardQ()[base + (uint32_t)__popc(m & ((1u << lane) - 1u))] = (uint16_t)LACB_TID;
}

// The k series of one level, left in the K plane (ends with a barrier).
//
// Stage 1, base k: one evaluation per chunk where k provably stays put (k_series_thread); the
// chunks where it moves are queued block-wide and dealt to all warps in pairs, one lane per sample
// (k_base_pair) -- unless most chunks are like that (short segments of the deep partition levels),
// in which case every owner runs the per-sample loop itself, the cheaper form then.
// Stage 2, stateful model only: bias by chunk-level proofs (k_bias_thread), the chunks that need
// the per-sample recurrences again queued and evaluated in pairs (k_bias_pair).
// Single-warp builds (the 256-sample probes) have nobody to share with and keep the per-thread loops.
//
// STATK: the last warp also evaluates block_static_k, in the interval where the other warps work
// on the queued chunks.
template <int NT, int E, bool STATEFUL, bool STATK = false>
__device__ __forceinline__ void k_series(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n,
                                         const SegGeom& sg, uint32_t p = 0u) {
  constexpr bool COOP = (NT >= 64) && (E == 16);
  constexpr uint32_t NW = NT / 32;
  const uint32_t tid = LACB_TID, warp = tid >> 5;
  AMisc* mi = sm.Misc();
  uint32_t* K = sm.Kpl() + tid * (E / 4);
  const uint16_t* hq = sm.HardQ();
  uint32_t kpk[E / 4];
  uint32_t flg;
  auto store_k = [&]() {
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) K[c4] = kpk[c4];
  };

  // ---- stage 1: base k (+ flag word)
  bool have;
  if (sg.fast) have = k_series_thread<NT, E, STATEFUL, true, COOP>(sm, pr, n, sg, kpk, flg);
  else have = k_series_thread<NT, E, STATEFUL, false, false>(sm, pr, n, sg, kpk, flg);
  if constexpr (COOP) {
    queue_push<NT, E>(sm, &mi->hq_kb_n, !have);
    if (have && !STATEFUL) store_k();
    LACB_SYNC();  // queue complete
    const uint32_t nh = mi->hq_kb_n;
    const bool pairs = nh <= (uint32_t)NT / 8u;
    if (pairs) {
      for (uint32_t i = warp * 2u; i < nh; i += NW * 2u)
        k_base_pair<NT, E, STATEFUL>(sm, hq[i], i + 1u < nh ? hq[i + 1u] : 0xFFFFu, n, p);
    } else if (!have) {
      k_series_thread<NT, E, STATEFUL, true, false>(sm, pr, n, sg, kpk, flg);
      if (!STATEFUL) store_k();
      have = true;
    }
    if (STATK && warp == NW - 1u) block_static_k<NT, E>(sm, n);
    if (!STATEFUL) {
      LACB_PH(8);
      LACB_SYNC();
      LACB_PH(9);
      if (tid == 0u) mi->hq_kb_n = 0u;  // consumed; the next pushes are at least one barrier away
      return;
    }
    if (have) sm.Flg()[tid] = flg;
    if (tid == 0u) mi->hq_n = 0u;
    LACB_SYNC();  // flag words (and the k bytes of the queued chunks) complete
    if (tid == 0u) mi->hq_kb_n = 0u;
    if (!have) {  // pick up what the pairs produced for this chunk
#pragma unroll
      for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = K[c4];
      flg = sm.Flg()[tid];
    }
  } else {
    if (!STATEFUL) {
      store_k();
      LACB_PH(8);
      LACB_SYNC();
      LACB_PH(9);
      return;
    }
    if (STATK && warp == NW - 1u) block_static_k<NT, E>(sm, n);
    sm.Flg()[tid] = flg;
    LACB_PH(6);
    LACB_SYNC();
    LACB_PH(7);
  }

  // ---- stage 2 (stateful model): bias
  const bool done = k_bias_thread<NT, E, !COOP>(sm, pr, flg, kpk);
  store_k();  // biased k, or the base k of a chunk left for the pairs
  if constexpr (COOP) {
    queue_push<NT, E>(sm, &mi->hq_n, !done);
    LACB_PH(8);
    LACB_SYNC();
    LACB_PH(9);
    const uint32_t nh = mi->hq_n;
    for (uint32_t i = warp * 2u; i < nh; i += NW * 2u)
      k_bias_pair<NT, E>(sm, hq[i], i + 1u < nh ? hq[i + 1u] : 0xFFFFu);
  }
  LACB_PH(8);
  LACB_SYNC();
  LACB_PH(9);
}

// Zero mask of the thread's E samples plus the next 4 (look-ahead for run detection),
// clipped to the block end.
template <int NT, int E>
__device__ __forceinline__ uint32_t zero_lookahead(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  uint32_t zm = pr.zmask;
  if (tid + 1u < (uint32_t)NT) {
    const uint4 v = reinterpret_cast<const uint4*>(sm.U())[swz_chunk((tid + 1u) * (E / 4))];
    const uint32_t nx = g0 + E;
    if (nx + 0u < n && v.x == 0u) zm |= 1u << (E + 0);
    if (nx + 1u < n && v.y == 0u) zm |= 1u << (E + 1);
    if (nx + 2u < n && v.z == 0u) zm |= 1u << (E + 2);
    if (nx + 3u < n && v.w == 0u) zm |= 1u << (E + 3);
  }
  return zm;
}

// Token of one sample under (mode, k): head bits, unary quotient, tail bits.
// Emission rules, block/encoder.cpp:585-771 with Rice::encode (rice.cpp:17-32) for the
// signed codes and write_rice_unsigned (encoder.cpp:79-87) for static / run lengths.
struct Token {
  uint32_t head, hlen;  // <= 3 bits
  uint32_t q;           // unary ones
  uint32_t tail, tlen;  // terminator + remainder, or a raw field (<= 32 bits)
};
__device__ __forceinline__ Token token_rice_signed(uint32_t u, uint32_t k, uint32_t head, uint32_t hlen) {
  Token t;
  t.head = head;
  t.hlen = hlen;
  t.q = u >> k;  // k <= 31 here (Rice::encode guards k >= 32 only)
  t.tail = k ? (u & ((1u << k) - 1u)) : 0u;
  t.tlen = k + 1u;  // the 0 terminator is the tail's leading bit
  return t;
}
__device__ __forceinline__ Token token_rice_unsigned(uint32_t u, uint32_t k, uint32_t head, uint32_t hlen) {
  Token t;
  t.head = head;
  t.hlen = hlen;
  t.q = (k >= 31u) ? 0u : (u >> k);
  t.tail = k ? (u & ((1u << k) - 1u)) : 0u;
  t.tlen = k + 1u;
  return t;
}

// Walks the thread's samples with the k series in place (K plane) and calls
//   f(j, idx, inB, u, k, is_zero, run_len_if_last /*0 unless this sample closes a run >= 4*/, in_long_run)
template <int NT, int E, bool FAST, bool ZR, typename F>
__device__ __forceinline__ void walk_thread(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n,
                                            const SegGeom& sg, uint32_t kinitA, uint32_t kinitB, F&& f) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  uint32_t u[E];
  load_u<NT, E>(sm, u);
  uint32_t kpk[E / 4];
  {
    const uint32_t* K = sm.Kpl() + tid * (E / 4);
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = K[c4];
  }
  const uint32_t zm_all = ZR ? zero_lookahead(sm, pr, n) : 0u;
  // samples at or after the segment boundary do not extend a run that started before it
  uint32_t zmA = zm_all;
  if (ZR && sg.bnd != 0xFFFFFFFFu && sg.bnd - g0 < (uint32_t)(E + 4)) zmA &= (1u << (sg.bnd - g0)) - 1u;
  uint32_t z = 0u;  // zeros immediately before g0 inside the current segment
  if (ZR) {
    const uint32_t byscan = g0 - 1u - (uint32_t)pr.lnz_ex;  // lnz_ex >= -1
    const uint32_t byseg = g0 - sg.a0;
    z = byscan < byseg ? byscan : byseg;
  }
  // k of the sample before this chunk (the last byte of the previous thread's K words)
  uint32_t kprev = tid ? (sm.Kpl()[tid * (E / 4) - 1u] >> 24) : 0u;
  if (g0 == sg.a0) kprev = kinitA;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t idx = g0 + j;
    if (FAST || idx < n) {
      const bool inB = !FAST && idx >= sg.bnd;
      uint32_t k = (j == 0) ? kprev : ((kpk[(j - 1) >> 2] >> (8 * ((j - 1) & 3))) & 0xFFu);
      if (!FAST && idx == sg.bnd) { k = kinitB; z = 0u; }
      if (ZR) {
        const uint32_t zm = inB ? zm_all : zmA;
        const bool is_zero = (zm >> j) & 1u;
        z = is_zero ? z + 1u : 0u;
        // zeros following this sample inside the segment, at most 4 looked at
        const uint32_t fwd = (uint32_t)__ffs((int)~(zm >> (j + 1))) - 1u;
        const bool long_run = is_zero && (z + fwd >= kZrMinRun);
        const uint32_t closes = (long_run && fwd == 0u) ? z : 0u;
        f(j, idx, inB, u[j], k, is_zero, closes, long_run);
      } else {
        f(j, idx, inB, u[j], k, u[j] == 0u, 0u, false);
      }
    }
  }
}
template <int NT, int E, typename F>
__device__ __forceinline__ void walk_items(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n,
                                           const SegGeom& sg, uint32_t kinitA, uint32_t kinitB, F&& f) {
  if (LACB_TID * E >= n) return;
  if (pr.any4) {
    if (sg.fast) walk_thread<NT, E, true, true>(sm, pr, n, sg, kinitA, kinitB, f);
    else walk_thread<NT, E, false, true>(sm, pr, n, sg, kinitA, kinitB, f);
  } else {
    if (sg.fast) walk_thread<NT, E, true, false>(sm, pr, n, sg, kinitA, kinitB, f);
    else walk_thread<NT, E, false, false>(sm, pr, n, sg, kinitA, kinitB, f);
  }
}

// estimate_residual_costs (block/encoder.cpp:201-263) for one level.  STATEFUL writes
// block totals to AMisc; otherwise per-segment prefix values go to Fb (three arrays) and
// the has-run bits to AMisc::hasrun_bits.
template <int NT, int E, bool STATEFUL>
__device__ __forceinline__ uint32_t cost_pass(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n, uint32_t p,
                                              uint32_t kinit_stateful) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  AMisc* mi = sm.Misc();
  const SegGeom sg = seg_geom<E>(g0, n, STATEFUL ? 0u : p);
  if (tid == 0) {
    mi->tot_rice = 0ull;
    mi->tot_zr = 0ull;
    mi->tot_bin = 0ull;
  }
  if (tid < 8) mi->hasrun_bits[tid] = 0u;
  // Full blocks: every chunk lies inside one segment and segments are whole groups of threads, so the
  // per-segment sums are accumulated directly (Fb as three arrays of sums) instead of through three block scans.
  const bool direct = !STATEFUL && n == (uint32_t)(NT * E);
  if (direct && tid < (1u << p)) {
    u64* Fz = sm.Fb();
    Fz[tid] = 0ull;
    Fz[ASmem<NT, E>::FBS + tid] = 0ull;
    Fz[2u * ASmem<NT, E>::FBS + tid] = 0ull;
  }
  k_series<NT, E, STATEFUL, STATEFUL>(sm, pr, n, sg, STATEFUL ? 0u : p);  // ends with a barrier
  uint32_t kinitA, kinitB = 0u;
  if (STATEFUL) {
    kinitA = mi->k_init;  // published by the last warp before the barriers inside k_series
    (void)kinit_stateful;
  } else {
    kinitA = sm.SegK()[sg.sidA] & 0xFFu;
    if (sg.bnd != 0xFFFFFFFFu) kinitB = sm.SegK()[sg.sidA + 1u] & 0xFFu;
  }
  u64 riceA = 0, zrA = 0, binA = 0, riceB = 0, zrB = 0, binB = 0;
  uint32_t runA = 0, runB = 0;
  // Plain chunks -- all E samples in one segment, every u > 4 (so no zero, no +-1/+-2 bin code)
  // and below 2^24 -- cost rice = sum (u >> k) + k + 1 per sample, and the other two modes
  // exactly two bits more per sample unless a zero-run escape (u > 8 << k, i.e. some q >= 8
  // here) can occur; those and all other chunks take the general walk.
  bool plain = sg.fast && !(pr.cls & 1u) && (pr.cls >> 8) <= 24u;
  if (plain) {
    uint32_t u[E];
    load_u<NT, E>(sm, u);
    const uint32_t* K = sm.Kpl() + tid * (E / 4);
    uint32_t kprev = tid ? (K[-1] >> 24) : 0u;
    if (g0 == sg.a0) kprev = kinitA;
    uint32_t acc = 0u, qor = 0u, ksum = 0u;  // acc is only used when every q < 8
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) {
      const uint32_t kw = K[c4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = 4 * c4 + b;
        const uint32_t k = kprev;
        kprev = (kw >> (8 * b)) & 0xFFu;
        const uint32_t q = (k >= 31u) ? 0u : (u[j] >> k);
        qor |= q;
        acc += q;
        ksum += k;
      }
    }
    if (qor < 8u) {
      riceA = (u64)(acc + ksum + (uint32_t)E);
      zrA = binA = riceA + 2ull * (uint32_t)E;
    } else {
      plain = false;
    }
  }
  // Silent chunks -- all E samples and the four after them zero, no segment boundary within
  // reach: every sample sits inside a zero run that closes later (zero-run cost 0, no run
  // closed here), takes 2 bits as a bin code and 1 + k as a Rice code.
  bool silent = false;
  if (!plain && sg.fast && pr.any4 && (pr.zmask & ((1u << E) - 1u)) == ((1u << E) - 1u) &&
      (sg.bnd == 0xFFFFFFFFu || sg.bnd >= g0 + (uint32_t)E + 4u)) {
    const uint32_t zm = zero_lookahead(sm, pr, n);
    if ((zm >> E) == 0xFu) {
      const uint32_t* K = sm.Kpl() + tid * (E / 4);
      uint32_t kprev = tid ? (K[-1] >> 24) : 0u;
      if (g0 == sg.a0) kprev = kinitA;
      uint32_t ksum = kprev;
#pragma unroll
      for (int c4 = 0; c4 < E / 4; ++c4) {
        const uint32_t kw = K[c4];
        ksum += (kw & 0xFFu) + ((kw >> 8) & 0xFFu) + ((kw >> 16) & 0xFFu) + ((kw >> 24) & 0xFFu);
      }
      ksum -= K[E / 4 - 1] >> 24;  // the last k belongs to the next chunk's first sample
      riceA = (u64)(ksum + (uint32_t)E);
      binA = 2ull * (uint32_t)E;
      zrA = 0ull;
      silent = true;
    }
  }
  if (!plain && !silent)
  walk_items<NT, E>(sm, pr, n, sg, kinitA, kinitB,
                    [&](int, uint32_t, bool inB, uint32_t u, uint32_t k, bool is_zero, uint32_t closes,
                        bool long_run) {
                      const u64 rc = rice_cost(u, k);
                      u64 bin, zr;
                      if (is_zero) bin = 2ull;
                      else if (u <= 4u) bin = 3ull;
                      else bin = 2ull + rc;
                      if (!pr.any4) {
                        zr = 0ull;  // never read: no segment can have a run, so zero-run mode is not a candidate
                      } else if (!is_zero) {
                        const uint32_t esc = 1u << (k + 3u < 24u ? k + 3u : 24u);
                        zr = 2ull + ((u > esc) ? 32ull : rc);
                      } else if (long_run) {
                        zr = closes ? 2ull + rice_cost(closes - kZrMinRun, kZrRunK) : 0ull;
                      } else {
                        zr = 2ull + rc;
                      }
                      if (inB) { riceB += rc; zrB += zr; binB += bin; runB |= (closes != 0u); }
                      else { riceA += rc; zrA += zr; binA += bin; runA |= (closes != 0u); }
                    });
  if (STATEFUL) {
    const u64 r = warp_sum_u64(riceA), z = warp_sum_u64(zrA), b = warp_sum_u64(binA);
    if ((tid & 31u) == 0u) {
      split_sum_add(&mi->tot_rice, r);
      if (pr.any4) split_sum_add(&mi->tot_zr, z);  // read only when the block has a run
      split_sum_add(&mi->tot_bin, b);
    }
    LACB_PH(10);
    const uint32_t any_run = (uint32_t)LACB_SYNC_OR((int)runA);  // also publishes the totals
    LACB_PH(11);
    return any_run;
  } else if (direct) {
    u64* Fb = sm.Fb();
    LACB_PH(10);
    const uint32_t per_seg = (uint32_t)NT >> p;  // threads per segment: 512 ... 4
    u64 a = riceA, b = zrA, c = binA;
    uint32_t run = runA;
    if (per_seg >= 32u) {
      a = warp_sum_u64(a);
      b = warp_sum_u64(b);
      c = warp_sum_u64(c);
      run = __reduce_or_sync(kFull, run);
    } else {
      for (uint32_t d = per_seg >> 1; d > 0u; d >>= 1) {  // groups of 16 / 8 / 4 lanes
        a += __shfl_xor_sync(kFull, a, (int)d);
        b += __shfl_xor_sync(kFull, b, (int)d);
        c += __shfl_xor_sync(kFull, c, (int)d);
        run |= __shfl_xor_sync(kFull, run, (int)d);
      }
    }
    const uint32_t lead = per_seg >= 32u ? 31u : per_seg - 1u;
    if ((tid & lead) == 0u) {
      atomicAdd(&Fb[sg.s0], a);
      if (pr.any4) atomicAdd(&Fb[ASmem<NT, E>::FBS + sg.s0], b);
      atomicAdd(&Fb[2u * ASmem<NT, E>::FBS + sg.s0], c);
      if (run) atomicOr(&mi->hasrun_bits[sg.s0 >> 5], 1u << (sg.s0 & 31u));
    }
    LACB_PH(14);
    LACB_SYNC();
    LACB_PH(15);
    return 1u;  // Fb holds sums, not prefixes
  } else {
    u64* Fb = sm.Fb();
    const uint32_t cnt = 1u << p;
    u64 tot;
    LACB_PH(10);
    const bool ownsA = (sg.a0 == g0);                                  // segment starts exactly at this thread
    const bool ownsB = (sg.bnd != 0xFFFFFFFFu && sg.bnd - g0 < (uint32_t)E && sg.bnd > g0);
    u64 ex = block_excl_scan_u64<NT>(riceA + riceB, sm.Scr(), &tot);
    if (ownsA) Fb[sg.s0] = ex;
    if (ownsB) Fb[sg.s0 + 1u] = ex + riceA;
    if (tid == 0) Fb[cnt] = tot;
    ex = block_excl_scan_u64<NT>(zrA + zrB, sm.Scr(), &tot);
    if (ownsA) Fb[ASmem<NT, E>::FBS + sg.s0] = ex;
    if (ownsB) Fb[ASmem<NT, E>::FBS + sg.s0 + 1u] = ex + zrA;
    if (tid == 0) Fb[ASmem<NT, E>::FBS + cnt] = tot;
    ex = block_excl_scan_u64<NT>(binA + binB, sm.Scr(), &tot);
    if (ownsA) Fb[2u * ASmem<NT, E>::FBS + sg.s0] = ex;
    if (ownsB) Fb[2u * ASmem<NT, E>::FBS + sg.s0 + 1u] = ex + binA;
    if (tid == 0) Fb[2u * ASmem<NT, E>::FBS + cnt] = tot;
    if (runA) atomicOr(&mi->hasrun_bits[sg.s0 >> 5], 1u << (sg.s0 & 31u));
    if (runB) atomicOr(&mi->hasrun_bits[(sg.s0 + 1u) >> 5], 1u << ((sg.s0 + 1u) & 31u));
    LACB_PH(14);
    LACB_SYNC();
    LACB_PH(15);
    return 0u;
  }
}

// ---------------------------------------------------------------------------
// All partition levels of a FULL block in one sweep (block/encoder.cpp:486-545).
//
// The stateless model restarts at every segment start, so the k series and the costs of a chunk depend on the
// level only through (a) the start of the segment the chunk lies in and (b) whether the chunk is the last one of
// its segment (zero runs are clipped at the boundary).  Segment starts nest: going one level deeper, a chunk in a
// left child keeps its start, and once a chunk is last-in-segment it stays so.  A thread therefore walks the
// levels with its samples in registers and re-evaluates its chunk only when (start, last, initial k) changes -- 5.5 of the 8
// levels on average instead of 8, each without the K plane, the hard-chunk queue or any block barrier (the k of the
// sample before the chunk is one more closed-form evaluation instead of a neighbour's K word).  The per-segment sums
// of every level are accumulated side by side (FbAll, aliasing the plane-count prefix that is dead by now) and the
// level choice is made after a single barrier.
template <int NT, int E, bool ZR, typename F>
__device__ __forceinline__ void walk_regs(const uint32_t (&u)[E], const uint32_t (&kpk)[E / 4], uint32_t kprev,
                                          uint32_t zm, uint32_t z, F&& f) {
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const uint32_t k = (j == 0) ? kprev : ((kpk[(j - 1) >> 2] >> (8 * ((j - 1) & 3))) & 0xFFu);
    if (ZR) {
      const bool is_zero = (zm >> j) & 1u;
      z = is_zero ? z + 1u : 0u;
      const uint32_t fwd = (uint32_t)__ffs((int)~(zm >> (j + 1))) - 1u;
      const bool long_run = is_zero && (z + fwd >= kZrMinRun);
      const uint32_t closes = (long_run && fwd == 0u) ? z : 0u;
      f(j, u[j], k, is_zero, closes, long_run);
    } else {
      f(j, u[j], k, u[j] == 0u, 0u, false);
    }
  }
}

// base k of a chunk that lies inside one segment starting at a0 (prefix of u there: Pa), samples in registers
template <int NT, int E>
__device__ __forceinline__ void k_chunk_stateless(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, const uint32_t (&u)[E],
                                                  uint32_t a0, u64 Pa, uint32_t (&kpk)[E / 4]) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  const u64 S = sm.Pthr()[tid + 1u] - pr.Pex;
  const uint32_t c_first = g0 - a0 + 1u, c_last = c_first + (uint32_t)E - 1u;
  const u64 rel = pr.Pex - Pa;
  const u64 N_first = rel + u[0] + (c_first >> 1), N_last = rel + S + (c_last >> 1);
  const uint32_t kb0 = kbase_clz(N_first, c_first);
  bool uniform;
  if (kb0 == 0u) {
    uniform = N_last < 2ull * c_first;
  } else {
    const u64 lo = (1ull << (kb0 - 1u)) + 1ull, hi = (1ull << kb0) + 1ull;
    uniform = (N_first >= lo * c_last) && (kb0 == 31u || N_last < hi * c_first);
  }
  if (uniform) {
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = kb0 * 0x01010101u;
    return;
  }
#pragma unroll
  for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = 0u;
  if (N_last < 0x80000000ull) {
    uint32_t N32 = (uint32_t)rel;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      N32 += u[j];
      const uint32_t c = c_first + (uint32_t)j;
      kpk[j >> 2] |= kbase_clz32(N32 + (c >> 1), c) << (8 * (j & 3));
    }
    return;
  }
  u64 N = rel;
#pragma unroll
  for (int j = 0; j < E; ++j) {
    N += u[j];
    const uint32_t c = c_first + (uint32_t)j;
    kpk[j >> 2] |= kbase_clz(N + (c >> 1), c) << (8 * (j & 3));
  }
}

// k series of the chunk in registers plus the k in force for its first sample: the segment's initial k at a
// segment start, else the model after the sample before the chunk (one more closed-form evaluation)
template <int NT, int E>
__device__ __forceinline__ void chunk_k_stateless(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, const uint32_t (&u)[E],
                                                  uint32_t sid, uint32_t a0, uint32_t (&kpk)[E / 4], uint32_t& kprev) {
  const uint32_t g0 = LACB_TID * E;
  const u64 Pa = sm.SegP()[sid];
  k_chunk_stateless<NT, E>(sm, pr, u, a0, Pa, kpk);
  if (g0 == a0) {
    kprev = sm.SegK()[sid] & 0xFFu;
  } else {
    const uint32_t c = g0 - a0;
    kprev = kbase_clz(pr.Pex - Pa + (c >> 1), c);
  }
}

// Token walk of the thread's chunk at partition level p >= 1 of a FULL block, k series in registers:
//   f(u, k, is_zero, closes, long_run, mk)   with mk = (mode << 5) | k of the chunk's segment (SelMK)
// `kinit_from_mk`: the emitter has no SegK table; the initial k of an adaptive segment is the k field of its mk.
template <int NT, int E, bool KINIT_FROM_MK, typename F>
__device__ __forceinline__ void chunk_walk_stateless(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n,
                                                     uint32_t p, F&& f) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  const uint32_t base = n >> p, cnt = 1u << p, per_seg = (uint32_t)NT >> p;
  const uint32_t s0 = tid / per_seg, a0 = s0 * base, sid = cnt - 1u + s0;
  const bool last = (s0 + 1u < cnt) && (g0 + (uint32_t)E == a0 + base);
  uint32_t u[E];
  load_u<NT, E>(sm, u);
  const uint32_t mk = sm.SelMK()[sid];
  uint32_t kpk[E / 4], kprev;
  if ((mk >> 5) == MODE_STATIC) {  // the adaptive series is not used: every sample takes the segment's static k
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) kpk[c4] = 0u;
    kprev = 0u;
  } else {
    chunk_k_stateless<NT, E>(sm, pr, u, sid, a0, kpk, kprev);
    if (KINIT_FROM_MK && g0 == a0) kprev = mk & 31u;
  }
  constexpr uint32_t kAll = (1u << E) - 1u;
  if (pr.any4) {
    const uint32_t zm_all = zero_lookahead(sm, pr, n);
    const uint32_t byscan = g0 - 1u - (uint32_t)pr.lnz_ex, byseg = g0 - a0;
    walk_regs<NT, E, true>(u, kpk, kprev, last ? (zm_all & kAll) : zm_all, byscan < byseg ? byscan : byseg,
                           [&](int, uint32_t uu, uint32_t k, bool is_zero, uint32_t closes, bool long_run) {
                             f(uu, k, is_zero, closes, long_run, mk);
                           });
  } else {
    walk_regs<NT, E, false>(u, kpk, kprev, 0u, 0u,
                            [&](int, uint32_t uu, uint32_t k, bool is_zero, uint32_t closes, bool long_run) {
                              f(uu, k, is_zero, closes, long_run, mk);
                            });
  }
}

// costs of one chunk of a full block under the stateless model of the segment [a0, ...) -- the three tiers of
// cost_pass with the k series in registers.  `last` = the chunk ends where its segment ends.
template <int NT, int E>
__device__ __forceinline__ void chunk_costs_stateless(const ASmem<NT, E>& sm, const Prep<NT, E>& pr,
                                                      const uint32_t (&u)[E], uint32_t zm_all, uint32_t sid, uint32_t a0,
                                                      bool last, u64& rice, u64& zr, u64& bin, uint32_t& run) {
  const uint32_t tid = LACB_TID, g0 = tid * E;
  uint32_t kpk[E / 4], kprev;
  chunk_k_stateless<NT, E>(sm, pr, u, sid, a0, kpk, kprev);
  run = 0u;
  bool plain = !(pr.cls & 1u) && (pr.cls >> 8) <= 24u;
  if (plain) {
    uint32_t acc = 0u, qor = 0u, ksum = 0u, kp = kprev;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      const uint32_t k = kp;
      kp = (kpk[j >> 2] >> (8 * (j & 3))) & 0xFFu;
      const uint32_t q = (k >= 31u) ? 0u : (u[j] >> k);
      qor |= q;
      acc += q;
      ksum += k;
    }
    if (qor < 8u) {
      rice = (u64)(acc + ksum + (uint32_t)E);
      zr = bin = rice + 2ull * (uint32_t)E;
      return;
    }
  }
  constexpr uint32_t kAll = (1u << E) - 1u;
  if (pr.any4 && (pr.zmask & kAll) == kAll && !last && (zm_all >> E) == 0xFu) {  // silent chunk inside a longer run
    uint32_t ksum = kprev;
#pragma unroll
    for (int c4 = 0; c4 < E / 4; ++c4) {
      const uint32_t kw = kpk[c4];
      ksum += (kw & 0xFFu) + ((kw >> 8) & 0xFFu) + ((kw >> 16) & 0xFFu) + ((kw >> 24) & 0xFFu);
    }
    ksum -= kpk[E / 4 - 1] >> 24;
    rice = (u64)(ksum + (uint32_t)E);
    bin = 2ull * (uint32_t)E;
    zr = 0ull;
    return;
  }
  u64 r = 0, z = 0, b = 0;
  uint32_t rn = 0u;
  auto body = [&](int, uint32_t uu, uint32_t k, bool is_zero, uint32_t closes, bool long_run) {
    const u64 rc = rice_cost(uu, k);
    u64 bc, zc;
    if (is_zero) bc = 2ull;
    else if (uu <= 4u) bc = 3ull;
    else bc = 2ull + rc;
    if (!pr.any4) {
      zc = 0ull;
    } else if (!is_zero) {
      const uint32_t esc = 1u << (k + 3u < 24u ? k + 3u : 24u);
      zc = 2ull + ((uu > esc) ? 32ull : rc);
    } else if (long_run) {
      zc = closes ? 2ull + rice_cost(closes - kZrMinRun, kZrRunK) : 0ull;
    } else {
      zc = 2ull + rc;
    }
    r += rc;
    z += zc;
    b += bc;
    rn |= (closes != 0u);
  };
  if (pr.any4) {
    // zeros right before the chunk inside the segment; look-ahead clipped at the segment end
    const uint32_t byscan = g0 - 1u - (uint32_t)pr.lnz_ex, byseg = g0 - a0;
    const uint32_t z0 = byscan < byseg ? byscan : byseg;
    walk_regs<NT, E, true>(u, kpk, kprev, last ? (zm_all & kAll) : zm_all, z0, body);
  } else {
    walk_regs<NT, E, false>(u, kpk, kprev, 0u, 0u, body);
  }
  rice = r;
  zr = z;
  bin = b;
  run = rn;
}

// Sums of all levels 1..max_p into FbAll[mode * (MAXSEG + 1) + sid] and the has-run bits of every segment into
// AMisc::hasrun_all.  Full blocks only (n == NT * E).  Ends with a barrier.
template <int NT, int E>
__device__ __forceinline__ void levels_fused(const ASmem<NT, E>& sm, const Prep<NT, E>& pr, uint32_t n, uint32_t max_p) {
  constexpr uint32_t STR = ASmem<NT, E>::MAXSEG + 1u;
  const uint32_t tid = LACB_TID, g0 = tid * E;
  AMisc* mi = sm.Misc();
  u64* Fa = sm.FbAll();
  for (uint32_t i = tid; i < 3u * STR; i += NT) Fa[i] = 0ull;
  if (tid < 16u) mi->hasrun_all[tid] = 0u;
  LACB_SYNC();
  uint32_t u[E];
  load_u<NT, E>(sm, u);
  const uint32_t zm_all = pr.any4 ? zero_lookahead(sm, pr, n) : 0u;
  uint32_t prevA = 0xFFFFFFFFu, prevKin = 0xFFFFFFFFu;
  bool prevLast = false;
  u64 rice = 0, zr = 0, bin = 0;
  uint32_t run = 0u;
  for (uint32_t p = 1u; p <= max_p; ++p) {
    const uint32_t base = n >> p, cnt = 1u << p;
    const uint32_t per_seg = (uint32_t)NT >> p;  // threads per segment: 512 ... 4
    const uint32_t s0 = tid / per_seg;
    const uint32_t a0 = s0 * base, sid = cnt - 1u + s0;
    const bool last = (s0 + 1u < cnt) && (g0 + (uint32_t)E == a0 + base);  // a boundary follows the chunk
    // a chunk that opens its segment also depends on the segment's initial k, which looks at the first
    // min(256, length) samples and so changes with the level once segments are shorter than that
    const uint32_t kin = (g0 == a0) ? (uint32_t)(sm.SegK()[sid] & 0xFFu) : 0xFFFFFFFFu;
    if (a0 != prevA || last != prevLast || kin != prevKin) {
      chunk_costs_stateless<NT, E>(sm, pr, u, zm_all, sid, a0, last, rice, zr, bin, run);
      prevA = a0;
      prevLast = last;
      prevKin = kin;
    }
    u64 a = rice, b = zr, c = bin;
    uint32_t rn = run;
    if (per_seg >= 32u) {
      a = warp_sum_u64(a);
      b = warp_sum_u64(b);
      c = warp_sum_u64(c);
      rn = __reduce_or_sync(kFull, rn);
    } else {
      for (uint32_t d = per_seg >> 1; d > 0u; d >>= 1) {  // groups of 16 / 8 / 4 lanes
        a += __shfl_xor_sync(kFull, a, (int)d);
        b += __shfl_xor_sync(kFull, b, (int)d);
        c += __shfl_xor_sync(kFull, c, (int)d);
        rn |= __shfl_xor_sync(kFull, rn, (int)d);
      }
    }
    const uint32_t lead = per_seg >= 32u ? 31u : per_seg - 1u;
    if ((tid & lead) == 0u) {
      split_sum_add(&Fa[sid], a);
      if (pr.any4) split_sum_add(&Fa[STR + sid], b);
      split_sum_add(&Fa[2u * STR + sid], c);
      if (rn) atomicOr(&mi->hasrun_all[sid >> 5], 1u << (sid & 31u));
    }
  }
  LACB_SYNC();
}

// Exact emitted bit count of the thread's samples for the final decision.
// part_mk[s] = (mode << 5) | k of segment s at level p.
__device__ __forceinline__ Token make_token(uint32_t mode, uint32_t kstatic, uint32_t u, uint32_t k, bool is_zero,
                                            uint32_t closes, bool long_run, bool* emit) {
  *emit = true;
  if (mode == MODE_RICE) return token_rice_signed(u, k, 0u, 0u);
  if (mode == MODE_STATIC) return token_rice_unsigned(u, kstatic, 0u, 0u);
  if (mode == MODE_BIN) {
    Token t;
    t.q = 0u;
    if (is_zero) { t.head = 0u; t.hlen = 2u; t.tail = 0u; t.tlen = 0u; return t; }
    if (u <= 2u) { t.head = (1u << 1) | (u & 1u); t.hlen = 3u; t.tail = 0u; t.tlen = 0u; return t; }   // +-1: tag 01 + sign
    if (u <= 4u) { t.head = (2u << 1) | (u & 1u); t.hlen = 3u; t.tail = 0u; t.tlen = 0u; return t; }   // +-2: tag 10 + sign
    return token_rice_signed(u, k, 3u, 2u);
  }
  // zero-run mode
  if (is_zero && long_run) {
    if (!closes) { *emit = false; Token t = {0u, 0u, 0u, 0u, 0u}; return t; }
    return token_rice_unsigned(closes - kZrMinRun, kZrRunK, 1u, 2u);
  }
  const uint32_t esc = 1u << (k + 3u < 24u ? k + 3u : 24u);
  if (u > esc) { Token t; t.head = 2u; t.hlen = 2u; t.q = 0u; t.tail = u; t.tlen = 32u; return t; }
  return token_rice_signed(u, k, 0u, 2u);
}

}  // namespace lacb
