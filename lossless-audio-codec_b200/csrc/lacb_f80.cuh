// lacb_f80.cuh -- software x87 double-extended arithmetic (64-bit significand,
// round-to-nearest-even) and the Levinson-Durbin recursion built on it.
//
// The reference runs LPC analysis in `long double` (src/codec/lpc/lpc.hpp:11-32,
// lpc.cpp:98-154), which on the x86-64 Linux build that defines the ".lac bytes"
// is the 80-bit x87 format.  Reproducing the Q15 coefficients bit for bit on a GPU
// therefore needs the same significand width and rounding; this file provides
// exactly the operations the recursion uses (+, -, *, /, <, int64 conversion and
// the final round-to-double of lpc.cpp:73-78).  Exponent range is not limited to
// 15 bits: the recursion cannot get near the x87 overflow/underflow thresholds
// (inputs are integers below 2^63 and every divisor is >= 1e-8).
#pragma once
#include "lacb_common.cuh"

namespace lacb {

struct f80 {
  u64 m;      // significand, MSB set unless the value is zero
  int32_t e;  // value = (-1)^s * m * 2^(e-63)
  uint32_t s;
};

__device__ __forceinline__ f80 f80_zero() { return f80{0ull, 0, 0u}; }
__device__ __forceinline__ f80 f80_make(u64 m, int32_t e, uint32_t s) { return f80{m, e, s}; }
__device__ __forceinline__ f80 f80_neg(f80 a) {
  a.s ^= 1u;
  return a;
}
__device__ __forceinline__ f80 f80_from_i64(i64 v) {
  if (v == 0) return f80_zero();
  const uint32_t s = v < 0;
  u64 m = s ? (0ull - (u64)v) : (u64)v;
  const int lz = __clzll((i64)m);
  return f80{m << lz, 63 - lz, s};
}

// round a 128-bit magnitude (hi:lo, hi's MSB set) with an extra sticky flag to 64 bits, RNE
__device__ __forceinline__ f80 f80_round(u64 hi, u64 lo, uint32_t sticky, int32_t e, uint32_t s) {
  const uint32_t rbit = (uint32_t)(lo >> 63);
  const uint32_t st = ((lo << 1) != 0ull) | sticky;
  if (rbit && (st || (hi & 1ull))) {
    hi += 1ull;
    if (hi == 0ull) {
      hi = 1ull << 63;
      e += 1;
    }
  }
  return f80{hi, e, s};
}

__device__ __forceinline__ f80 f80_mul(f80 a, f80 b) {
  if (a.m == 0ull || b.m == 0ull) return f80{0ull, 0, a.s ^ b.s};
  u64 hi = __umul64hi(a.m, b.m), lo = a.m * b.m;
  int32_t e = a.e + b.e + 1;
  if (!(hi >> 63)) {  // product in [2^126, 2^127): normalise by one bit
    hi = (hi << 1) | (lo >> 63);
    lo <<= 1;
    e -= 1;
  }
  return f80_round(hi, lo, 0u, e, a.s ^ b.s);
}

// magnitude compare: -1, 0, +1
__device__ __forceinline__ int f80_cmp_mag(f80 a, f80 b) {
  if (a.m == 0ull || b.m == 0ull) return (a.m != 0ull) - (b.m != 0ull);
  if (a.e != b.e) return a.e > b.e ? 1 : -1;
  return (a.m > b.m) - (a.m < b.m);
}

__device__ __forceinline__ f80 f80_add(f80 a, f80 b) {
  if (b.m == 0ull) return a;
  if (a.m == 0ull) return b;
  if (f80_cmp_mag(a, b) < 0) {
    const f80 t = a;
    a = b;
    b = t;
  }
  // |a| >= |b|; align b to a's exponent in a 128-bit window (a.m in the high word)
  const uint32_t d = (uint32_t)(a.e - b.e);
  u64 bhi, blo;
  uint32_t sticky = 0u;
  if (d == 0u) {
    bhi = b.m;
    blo = 0ull;
  } else if (d < 64u) {
    bhi = b.m >> d;
    blo = b.m << (64u - d);
  } else if (d == 64u) {
    bhi = 0ull;
    blo = b.m;
  } else if (d < 128u) {
    bhi = 0ull;
    blo = b.m >> (d - 64u);
    sticky = (b.m << (128u - d)) != 0ull;
  } else {
    bhi = 0ull;
    blo = 0ull;
    sticky = 1u;
  }
  u64 hi, lo;
  int32_t e = a.e;
  if (a.s == b.s) {
    lo = blo;
    hi = a.m + bhi;
    if (hi < a.m) {  // carry out: shift right by one
      sticky |= (uint32_t)(lo & 1ull);
      lo = (lo >> 1) | (hi << 63);
      hi = (hi >> 1) | (1ull << 63);
      e += 1;
    }
    return f80_round(hi, lo, sticky, e, a.s);
  }
  // subtraction: a - (b + epsilon) when sticky
  lo = 0ull - blo;
  hi = a.m - bhi - (blo != 0ull);
  if (sticky) {  // borrow one unit of the lowest kept bit; the remainder stays sticky
    if (lo == 0ull) hi -= 1ull;
    lo -= 1ull;
  }
  if (hi == 0ull && lo == 0ull) return f80{0ull, 0, 0u};  // exact cancellation -> +0 (RNE)
  if (hi == 0ull) {
    hi = lo;
    lo = 0ull;
    e -= 64;
  }
  const int lz = __clzll((i64)hi);
  if (lz) {
    hi = (hi << lz) | (lo >> (64 - lz));
    lo <<= lz;
    e -= lz;
  }
  return f80_round(hi, lo, sticky, e, a.s);
}
__device__ __forceinline__ f80 f80_sub(f80 a, f80 b) { return f80_add(a, f80_neg(b)); }

// a / b, b != 0
__device__ __forceinline__ f80 f80_div(f80 a, f80 b) {
  if (a.m == 0ull) return f80{0ull, 0, a.s ^ b.s};
  // quotient of the significands, 64 bits with the MSB set, by restoring division
  u64 rem = a.m, q = 0ull;
  int32_t e = a.e - b.e;
  // first bit
  if (rem >= b.m) {
    rem -= b.m;
    q = 1ull;
  } else {
    e -= 1;
    // rem < b.m: shift in one more bit so that the first quotient bit is 1
    const uint32_t top = (uint32_t)(rem >> 63);
    rem <<= 1;
    if (top || rem >= b.m) rem -= b.m;  // always true here since 2*a.m >= 2^64 > b.m
    q = 1ull;
  }
  for (int i = 0; i < 63; ++i) {
    const uint32_t top = (uint32_t)(rem >> 63);
    rem <<= 1;
    q <<= 1;
    if (top || rem >= b.m) {
      rem -= b.m;
      q |= 1ull;
    }
  }
  // rounding: compare 2*rem with b.m
  uint32_t up = 0u;
  if (rem != 0ull) {
    const u64 other = b.m - rem;  // rem < b.m
    if (rem > other) up = 1u;
    else if (rem == other) up = (uint32_t)(q & 1ull);
  }
  if (up) {
    q += 1ull;
    if (q == 0ull) {
      q = 1ull << 63;
      e += 1;
    }
  }
  return f80{q, e, a.s ^ b.s};
}

// a < b (signed)
__device__ __forceinline__ bool f80_lt(f80 a, f80 b) {
  const bool az = a.m == 0ull, bz = b.m == 0ull;
  if (az && bz) return false;
  if (az) return !b.s;
  if (bz) return a.s != 0u;
  if (a.s != b.s) return a.s != 0u;
  const int c = f80_cmp_mag(a, b);
  return a.s ? (c > 0) : (c < 0);
}

// quantize_coeff_q15 (lpc.cpp:73-78): (double)a, * 32768.0, std::round (half away
// from zero), clamp to int16.
__device__ __forceinline__ int32_t f80_quant_q15(f80 a) {
  if (a.m == 0ull) return 0;
  if (a.e < -18) return 0;                 // |a| < 2^-17: rounds to 0 whatever the mantissa
  if (a.e >= 16) return a.s ? -32768 : 32767;
  // (double)a: round the 64-bit significand to 53 bits, RNE
  u64 m = a.m >> 11;
  const u64 rest = a.m & 0x7FFull;
  int32_t e = a.e;
  if (rest > 0x400ull || (rest == 0x400ull && (m & 1ull))) {
    m += 1ull;
    if (m >> 53) {
      m >>= 1;
      e += 1;
    }
  }
  // value = m * 2^(e-52); scaled = m * 2^(e-37); round half away from zero
  const int sh = 37 - e;  // scaled = m / 2^sh
  i64 r;
  if (sh <= 0) {
    r = (sh < -10) ? (i64)1 << 40 : (i64)(m << (-sh));
  } else if (sh > 54) {
    r = 0;
  } else {
    const u64 ip = m >> sh;
    const u64 half = (m >> (sh - 1)) & 1ull;
    r = (i64)(ip + half);
  }
  if (a.s) r = -r;
  if (r < -32768) r = -32768;
  if (r > 32767) r = 32767;
  return (int32_t)r;
}

// Levinson-Durbin with per-order snapshots.
//
// LPC::levinson_durbin (lpc.cpp:98-154) run once to order 12; analyze_block_q15
// (lpc.cpp:156-186) asks for orders {4,6,8,10,12} separately, but iteration i only
// reads state produced by iterations < i, so the coefficients after iteration `o`
// of one order-12 run are those of a separate order-`o` run.  R[0..12] are the exact
// int64 autocorrelations (lpc.cpp:80-96).  Outputs, for candidate c (order 4+2c):
// used[c] (0 => unstable, candidate dropped) and coef[c][1..12] in Q15.
__device__ inline void levinson_q15(const i64* R, int max_order, int16_t (*coef)[13], int8_t* used) {
  // 0.999L and 1e-8L as the x87 bit patterns g++ emits for those literals
  const f80 k999 = f80_make(0xFFBE76C8B4395810ull, -1, 0u);
  const f80 eps = f80_make(0xABCC77118461CEFDull, -27, 0u);
  const f80 one = f80_make(1ull << 63, 0, 0u);
  f80 Rf[13], a[13], prevA[13];
  for (int i = 0; i <= 12; ++i) {
    Rf[i] = f80_from_i64(i <= max_order ? R[i] : 0);
    a[i] = f80_zero();
    prevA[i] = f80_zero();
  }
  if (f80_lt(Rf[0], one)) Rf[0] = one;  // lpc.cpp:172-174
  for (int c = 0; c < 5; ++c) {
    used[c] = 0;
    for (int i = 0; i <= 12; ++i) coef[c][i] = 0;
  }
  f80 E = Rf[0];
  int achieved = 0;
  bool stopped = f80_lt(E, eps);
  for (int i = 1; i <= 12; ++i) {
    if (!stopped && i <= max_order) {
      f80 acc = f80_zero();
      for (int j = 1; j < i; ++j) acc = f80_add(acc, f80_mul(prevA[j], Rf[i - j]));
      if (f80_lt(E, eps)) {
        stopped = true;
      } else {
        f80 ki = f80_div(f80_sub(Rf[i], acc), E);
        if (f80_lt(k999, ki)) ki = k999;
        if (f80_lt(ki, f80_neg(k999))) ki = f80_neg(k999);
        const f80 e_new = f80_mul(f80_sub(one, f80_mul(ki, ki)), E);
        if (f80_lt(e_new, eps)) {
          stopped = true;
        } else {
          a[i] = ki;
          for (int j = 1; j < i; ++j) a[j] = f80_sub(prevA[j], f80_mul(ki, prevA[i - j]));
          for (int j = 1; j <= i; ++j) prevA[j] = a[j];
          E = e_new;
          achieved = i;
        }
      }
    } else {
      stopped = true;
    }
    // snapshot for the candidate whose order is i, or for every later candidate once stopped
    if (i >= 4 && (i & 1) == 0) {
      const int c = (i - 4) >> 1;
      if (i <= max_order) {
        used[c] = (int8_t)achieved;
        for (int j = 1; j <= achieved; ++j) coef[c][j] = (int16_t)f80_quant_q15(a[j]);
      }
    }
  }
}

}  // namespace lacb
