// TEST INFRASTRUCTURE: operation-level check of the software x87 double-extended arithmetic
// (lossless-audio-codec_b200/csrc/lacb_f80.cuh, compiled here through the CPU emulator header) against the
// host's native `long double`, which on x86-64 Linux is the 80-bit x87 format the reference's Levinson
// recursion runs in (src/codec/lpc/lpc.hpp:11-32, lpc.cpp:98-154).  Every result is compared bit for bit:
// sign, 64-bit significand and exponent.  usage: f80_check [operations]   (exit status 0 = all equal)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cuda_emu.h"
#include "../../lossless-audio-codec_b200/csrc/lacb_f80.cuh"

using lacb::f80;
typedef unsigned long long u64;
typedef long long i64;

static u64 rng_state = 0x9E3779B97F4A7C15ull;
static u64 rnd() {
  u64 x = rng_state;
  x ^= x << 13;
  x ^= x >> 7;
  x ^= x << 17;
  return rng_state = x;
}

static f80 from_ld(long double v) {
  static_assert(sizeof(long double) >= 10, "long double must be the x87 80-bit type");
  unsigned char b[16];
  std::memcpy(b, &v, sizeof v);
  u64 m;
  uint16_t se;
  std::memcpy(&m, b, 8);
  std::memcpy(&se, b + 8, 2);
  if (m == 0) return lacb::f80_make(0, 0, se >> 15);
  return lacb::f80_make(m, (int32_t)(se & 0x7FFF) - 16383, se >> 15);
}
static long double to_ld(f80 a) {
  if (a.m == 0) return a.s ? -0.0L : 0.0L;
  const long double mag = std::ldexp((long double)a.m, a.e - 63);  // exact: m has 64 bits, long double holds 64
  return a.s ? -mag : mag;
}
static bool same(f80 a, long double want, const char* what, long double x, long double y) {
  const f80 w = from_ld(want);
  const bool ok = (a.m == 0 && w.m == 0) ? true : (a.m == w.m && a.e == w.e && a.s == w.s);
  if (!ok)
    std::fprintf(stderr, "%s mismatch: x=%La y=%La  got m=%016llx e=%d s=%u  want m=%016llx e=%d s=%u\n", what, x, y,
                 a.m, a.e, a.s, w.m, w.e, w.s);
  return ok;
}
// an integer of random width up to 63 bits (autocorrelation values: R[0] above 2^53 included), random sign
static i64 rand_int() {
  const unsigned width = 1 + (unsigned)(rnd() % 63);
  i64 v = (i64)(rnd() >> (64 - width));
  if ((rnd() & 7) == 0) v = (i64)((1ull << 53) + (rnd() >> 12));  // just above the double range
  return (rnd() & 1) ? -v : v;
}
static int16_t quant_ref(double c) {  // LPC::quantize_coeff_q15, lpc.cpp:73-78
  double scaled = std::round(c * 32768.0);
  if (scaled < -32768.0) scaled = -32768.0;
  if (scaled > 32767.0) scaled = 32767.0;
  return (int16_t)scaled;
}

int main(int argc, char** argv) {
  const long long want_ops = argc > 1 ? std::atoll(argv[1]) : 10000000ll;
  long long ops = 0, bad = 0;
  while (ops < want_ops && bad < 10) {
    // operands the recursion meets: integers, their quotients (reflection coefficients), products and sums
    const i64 ia = rand_int(), ib = rand_int();
    const long double la = (long double)ia, lb = (long double)ib;
    const f80 fa = lacb::f80_from_i64(ia), fb = lacb::f80_from_i64(ib);
    bad += !same(fa, la, "from_i64", la, 0);
    bad += !same(lacb::f80_add(fa, fb), la + lb, "add(int,int)", la, lb);
    bad += !same(lacb::f80_sub(fa, fb), la - lb, "sub(int,int)", la, lb);
    bad += !same(lacb::f80_mul(fa, fb), la * lb, "mul(int,int)", la, lb);
    ops += 4;
    if (ib != 0) {
      const long double lq = la / lb;
      const f80 fq = lacb::f80_div(fa, fb);
      bad += !same(fq, lq, "div(int,int)", la, lb);
      // second generation: non-integers with full 64-bit significands
      const long double lp = lq * la, ls = lq + lb, ld2 = lq - la;
      const f80 fp = lacb::f80_mul(fq, fa), fsum = lacb::f80_add(fq, fb), fd2 = lacb::f80_sub(fq, fa);
      bad += !same(fp, lp, "mul(frac,int)", lq, la);
      bad += !same(fsum, ls, "add(frac,int)", lq, lb);
      bad += !same(fd2, ld2, "sub(frac,int)", lq, la);
      ops += 4;
      if (lp != 0.0L) {
        bad += !same(lacb::f80_div(fsum, fp), ls / lp, "div(frac,frac)", ls, lp);
        bad += !same(lacb::f80_mul(fq, fq), lq * lq, "mul(frac,frac)", lq, lq);
        // 1 - k*k and error *= (1 - k*k): the update of lpc.cpp:131-140
        const long double one = 1.0L, l1 = one - lq * lq;
        const f80 f1 = lacb::f80_sub(lacb::f80_from_i64(1), lacb::f80_mul(fq, fq));
        bad += !same(f1, l1, "1-k*k", lq, lq);
        bad += !same(lacb::f80_mul(fa, f1), la * l1, "err*(1-k*k)", la, l1);
        bad += (lacb::f80_lt(fq, fp) != (lq < lp));
        bad += (lacb::f80_lt(fp, fq) != (lp < lq));
        ops += 6;
      }
      // quantisation of coefficient-sized values (and of wild ones: clamping)
      const long double lc = (rnd() & 3) ? std::ldexp(lq, -(int)(rnd() % 8)) / (std::fabs(lq) > 4 ? std::fabs(lq) : 1.0L) : lq;
      const f80 fc = from_ld(lc);
      if ((int32_t)quant_ref((double)lc) != lacb::f80_quant_q15(fc)) {
        std::fprintf(stderr, "quant_q15 mismatch: c=%La got %d want %d\n", lc, lacb::f80_quant_q15(fc), (int)quant_ref((double)lc));
        ++bad;
      }
      ops += 1;
      (void)to_ld;
    }
  }
  std::printf("%lld operations compared with native long double, %lld mismatches\n", ops, bad);
  return bad ? 1 : 0;
}
