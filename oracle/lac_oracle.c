/* oracle/lac_oracle.c -- TEST INFRASTRUCTURE ONLY (see lac_oracle.h).
 *
 * Plain C restatement of the reference LAC codec's block/frame algorithm.
 * Every function names the reference file:line it follows (paths relative to
 * the reference tree).  Straight-line, scalar, one sample at a time: this is
 * the *checker*, written for obviousness, not speed.  `long double` is the
 * x86-64 80-bit type here exactly as in the reference build.
 */
#define _GNU_SOURCE
#include "lac_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* constants: src/codec/block/constants.hpp:6-15, block/encoder.cpp:41-59 */
enum {
  ZR_MIN_RUN = 4,
  ZR_RUN_K = 2,
  MIN_PART = 32,
  MAX_PART_ORDER = 8,
  MODE_RICE = 0,
  MODE_ZR = 1,
  MODE_BIN = 2,
  MODE_STATIC = 3,
  PRED_FIXED = 0,
  PRED_FIR = 1,
  PRED_LPC = 2,
  DRIFT_WIN = 256, /* rice.hpp:12 */
  MICRO_WIN = 96   /* rice.hpp:13 */
};

static __thread char g_err[256];
const char* lao_last_error(void) { return g_err; }
void lao_free(void* p) { free(p); }

/* ------------------------------------------------------------------ */
/* MSB-first bit sink: src/codec/bitstream/bit_writer.cpp:15-119       */
typedef struct {
  uint8_t* buf;
  size_t cap;
  uint64_t nbits;
} bitsink;

static void sink_init(bitsink* s) {
  s->cap = 1024;
  s->buf = (uint8_t*)calloc(s->cap, 1);
  s->nbits = 0;
}
static void sink_put1(bitsink* s, uint32_t bit) {
  size_t byte = (size_t)(s->nbits >> 3);
  if (byte >= s->cap) {
    size_t ncap = s->cap * 2;
    s->buf = (uint8_t*)realloc(s->buf, ncap);
    memset(s->buf + s->cap, 0, ncap - s->cap);
    s->cap = ncap;
  }
  if (bit) s->buf[byte] |= (uint8_t)(0x80u >> (s->nbits & 7u));
  s->nbits++;
}
static void sink_put(bitsink* s, uint32_t value, int nbits) {
  for (int i = nbits - 1; i >= 0; --i) sink_put1(s, (i >= 32) ? 0u : ((value >> i) & 1u));
}
static void sink_ones(bitsink* s, uint32_t count) {
  while (count--) sink_put1(s, 1u);
}
static void sink_pad(bitsink* s) { /* flush_to_byte, bit_writer.cpp:105-111 */
  while (s->nbits & 7u) sink_put1(s, 0u);
}

/* ------------------------------------------------------------------ */
/* MSB-first bit source: src/codec/bitstream/bit_reader.hpp:40-202     */
typedef struct {
  const uint8_t* data;
  uint64_t size_bits;
  uint64_t pos;
  int error;
} bitsrc;

static void src_fail(bitsrc* r) { /* mark_error, bit_reader.hpp:50-54 */
  r->error = 1;
  r->pos = r->size_bits;
}
static uint64_t src_left(const bitsrc* r) { return r->error ? 0 : r->size_bits - r->pos; }
static uint32_t src_get(bitsrc* r, int nbits) { /* read_bits, :75-138 */
  if (nbits <= 0) return 0;
  if (r->error || r->pos >= r->size_bits || (uint64_t)nbits > r->size_bits - r->pos) {
    src_fail(r);
    return 0;
  }
  uint32_t v = 0;
  for (int i = 0; i < nbits; ++i) {
    uint64_t p = r->pos++;
    v = (v << 1) | ((r->data[p >> 3] >> (7u - (p & 7u))) & 1u);
  }
  return v;
}
/* read_unary_ones, bit_reader.hpp:140-172: count ones up to a zero terminator,
 * rejecting as soon as the count would pass max_ones; running off the end is an
 * error.  The reference checks the limit per byte-run, which is equivalent to
 * rejecting when the number of ones before the terminator (or before the end of
 * data) exceeds max_ones. */
static int src_unary(bitsrc* r, uint32_t max_ones, uint32_t* ones) {
  uint32_t n = 0;
  *ones = 0;
  while (r->pos < r->size_bits) {
    uint64_t p = r->pos;
    uint32_t bit = (r->data[p >> 3] >> (7u - (p & 7u))) & 1u;
    if (!bit) {
      r->pos++;
      *ones = n;
      return 1;
    }
    if (n == max_ones) { /* one more would exceed the limit */
      *ones = n;
      return 0;
    }
    n++;
    r->pos++;
  }
  *ones = n;
  src_fail(r);
  return 0;
}
static int src_skip_zero_pad(bitsrc* r) { /* consume_zero_padding_to_byte, :180-185 */
  while (r->pos & 7u) {
    if (src_get(r, 1) != 0u || r->error) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------ */
/* zig-zag + Rice: block/encoder.cpp:61-87, rice/rice.cpp:7-52         */
static uint32_t zz(int32_t r) { return ((uint32_t)r << 1) ^ (r < 0 ? 0xFFFFFFFFu : 0u); }
static int32_t unzz(uint32_t u) {
  if ((u & 1u) == 0u) return (int32_t)(u >> 1);
  return (int32_t)(-(int64_t)((u >> 1) + 1u));
}
static uint64_t rice_cost(uint32_t u, uint32_t k) { /* rice_bits_for_unsigned, encoder.cpp:67-70 */
  uint32_t q = (k >= 31u) ? 0u : (u >> k);
  return (uint64_t)q + 1u + k;
}
/* Rice::encode (rice.cpp:17-32): shifts guarded at k>=32 only. */
static void rice_put_signed(bitsink* s, int32_t v, uint32_t k) {
  uint32_t u = zz(v);
  uint32_t q = (k >= 32u) ? 0u : (u >> k);
  uint32_t rem = (k >= 32u) ? u : (u & (((uint32_t)1 << k) - 1u));
  sink_ones(s, q);
  sink_put1(s, 0);
  if (k > 0) sink_put(s, rem, (int)k);
}
/* write_rice_unsigned (encoder.cpp:79-87): quotient forced to 0 at k>=31. */
static void rice_put_unsigned(bitsink* s, uint32_t u, uint32_t k) {
  uint32_t q = (k >= 31u) ? 0u : (u >> k);
  sink_ones(s, q);
  sink_put1(s, 0);
  if (k > 0) sink_put(s, u & ((1u << k) - 1u), (int)k);
}

/* ------------------------------------------------------------------ */
/* adaptive k: rice/rice.hpp:15-114 (stateful), block/encoder.cpp:72-77 (stateless) */
typedef struct {
  uint64_t prev_sum;
  uint32_t win_idx, micro_idx, win_filled;
  uint64_t win_sum;
  uint32_t large_cnt, zero_cnt;
  uint32_t recent[DRIFT_WIN];
  uint8_t large[MICRO_WIN], zero[MICRO_WIN];
} kstate;

static uint32_t bit_width64(uint64_t v) {
  uint32_t w = 0;
  while (v) {
    ++w;
    v >>= 1;
  }
  return w;
}
static uint32_t k_stateless(uint64_t sum, uint32_t count) {
  if (count == 0) return 0;
  uint64_t mean = (sum + (count >> 1)) / count;
  if (mean <= 1) return 0;
  uint32_t k = bit_width64(mean - 1u);
  return k > 31u ? 31u : k;
}
static uint32_t k_stateful(uint64_t sum, uint32_t count, kstate* st) {
  if (count == 0) return 0;
  uint64_t cur = sum - st->prev_sum;
  st->prev_sum = sum;
  uint32_t mi = st->micro_idx;
  st->large_cnt -= st->large[mi];
  st->zero_cnt -= st->zero[mi];
  if (st->win_filled < DRIFT_WIN)
    st->win_filled++;
  else
    st->win_sum -= st->recent[st->win_idx];
  st->recent[st->win_idx] = (uint32_t)cur;
  st->win_sum += cur;

  uint64_t mean = (sum + (count >> 1)) / count;
  uint32_t k = 0;
  if (mean > 1) {
    k = bit_width64(mean - 1u);
    if (k > 31u) k = 31u;
  }
  uint32_t q = (k >= 31u) ? 0u : (uint32_t)(cur >> k);
  uint8_t is_large = q > 3u, is_zero = q == 0u;
  st->large_cnt += is_large;
  st->zero_cnt += is_zero;
  st->large[mi] = is_large;
  st->zero[mi] = is_zero;

  int bias = 0;
  if (st->win_filled > 0 && mean > 0) {
    uint64_t lm = (st->win_filled == DRIFT_WIN)
                      ? ((st->win_sum + (DRIFT_WIN >> 1)) >> 8)
                      : ((st->win_sum + (st->win_filled >> 1)) / st->win_filled);
    if (lm * 3 > mean * 4)
      bias = 1;
    else if (lm * 4 + 3 < mean * 3)
      bias = -1;
  }
  if (st->win_idx + 1 >= MICRO_WIN || st->win_filled >= MICRO_WIN) {
    uint32_t wsz = (st->win_filled >= MICRO_WIN) ? MICRO_WIN : st->win_filled;
    /* the counters are uint16 in the reference; the products are computed in int */
    if ((uint32_t)(uint16_t)st->large_cnt * 4 >= wsz * 3) {
      bias = bias + 1 < 1 ? bias + 1 : 1;
    } else if ((uint32_t)(uint16_t)st->zero_cnt * 5 >= wsz * 4) {
      bias = bias - 1 > -1 ? bias - 1 : -1;
    }
  }
  int bk = (int)k + bias;
  if (bk < 0) bk = 0;
  if (bk > 31) bk = 31;
  st->micro_idx = (st->micro_idx + 1u == MICRO_WIN) ? 0u : st->micro_idx + 1u;
  st->win_idx = (st->win_idx + 1u) & (DRIFT_WIN - 1u);
  return (uint32_t)bk;
}
static uint32_t k_next(uint64_t sum, uint32_t count, int stateless, kstate* st) {
  return stateless ? k_stateless(sum, count) : k_stateful(sum, count, st);
}

void lao_adaptive_k_series(const uint32_t* u, uint32_t n, uint32_t initial_k, int stateless,
                           uint32_t* k_out) {
  kstate st;
  memset(&st, 0, sizeof st);
  uint64_t sum = 0;
  uint32_t k = initial_k;
  for (uint32_t i = 0; i < n; ++i) {
    k_out[i] = k;
    sum += u[i];
    k = k_next(sum, i + 1, stateless, &st);
  }
}

/* ------------------------------------------------------------------ */
/* cost estimators: block/encoder.cpp:121-263                           */
static uint32_t est_initial_k(const int32_t* r, uint32_t n) { /* :121-158 */
  if (n == 0) return 0;
  uint32_t cnt = n < 256 ? n : 256;
  uint64_t cost[13];
  memset(cost, 0, sizeof cost);
  for (uint32_t i = 0; i < cnt; ++i) {
    uint32_t u = zz(r[i]);
    for (uint32_t k = 0; k <= 12; ++k) cost[k] += (uint64_t)(u >> k) + 1u + k;
  }
  uint32_t best = 0;
  uint64_t bc = UINT64_MAX;
  for (uint32_t k = 0; k <= 12; ++k)
    if (cost[k] < bc) {
      bc = cost[k];
      best = k;
    }
  return best; /* the mean-based fallback (:139-147) can never survive the argmin */
}
static uint32_t est_static_k(const int32_t* r, uint32_t n, uint64_t* bits_out) { /* :160-188 */
  uint64_t cost[16];
  memset(cost, 0, sizeof cost);
  if (n == 0) {
    if (bits_out) *bits_out = 0;
    return 0;
  }
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t u = zz(r[i]);
    for (uint32_t k = 0; k <= 15; ++k) cost[k] += rice_cost(u, k);
  }
  uint32_t best = 0;
  uint64_t bc = UINT64_MAX;
  for (uint32_t k = 0; k <= 15; ++k)
    if (cost[k] < bc) {
      bc = cost[k];
      best = k;
    }
  if (bits_out) *bits_out = bc;
  return best;
}
typedef struct {
  uint64_t rice, zr, bin;
  int has_run;
} seg_costs;
static seg_costs est_costs(const int32_t* r, uint32_t n, uint32_t k0, int stateless) { /* :201-263 */
  seg_costs c = {0, 0, 0, 0};
  if (n == 0) return c;
  kstate st;
  memset(&st, 0, sizeof st);
  uint32_t k = k0, count = 0, idx = 0;
  uint64_t sum = 0;
  while (idx < n) {
    uint32_t run = 0;
    while (idx + run < n && r[idx + run] == 0) ++run;
    if (run >= ZR_MIN_RUN) {
      c.has_run = 1;
      c.zr += 2 + rice_cost(run - ZR_MIN_RUN, ZR_RUN_K);
      for (uint32_t j = 0; j < run; ++j) {
        c.rice += rice_cost(0, k);
        c.bin += 2;
        ++count;
        k = k_next(sum, count, stateless, &st);
      }
      idx += run;
      continue;
    }
    int32_t v = r[idx];
    uint32_t u = zz(v);
    c.rice += rice_cost(u, k);
    if (v == 0)
      c.bin += 2;
    else if (v == 1 || v == -1 || v == 2 || v == -2)
      c.bin += 3;
    else
      c.bin += 2 + rice_cost(u, k);
    uint32_t esc = 1u << (k + 3u < 24u ? k + 3u : 24u);
    c.zr += 2 + ((u > esc) ? 32u : rice_cost(u, k));
    sum += u;
    ++count;
    k = k_next(sum, count, stateless, &st);
    ++idx;
  }
  return c;
}

/* ------------------------------------------------------------------ */
/* predictors: block/encoder.cpp:265-309, lpc/lpc.cpp:38-229            */
static void fixed_residual(const int32_t* x, uint32_t n, int order, int32_t* r) { /* :265-295 */
  for (uint32_t i = 0; i < n; ++i) {
    if ((int)i < order) {
      r[i] = x[i];
      continue;
    }
    int64_t p = 0;
    switch (order) {
      case 1: p = x[i - 1]; break;
      case 2: p = 2LL * x[i - 1] - x[i - 2]; break;
      case 3: p = 3LL * x[i - 1] - 3LL * x[i - 2] + x[i - 3]; break;
      case 4: p = 4LL * x[i - 1] - 6LL * x[i - 2] + 4LL * x[i - 3] - x[i - 4]; break;
      default: p = 0; break;
    }
    r[i] = (int32_t)(uint32_t)(uint64_t)((int64_t)x[i] - p); /* truncating cast */
  }
}
static void fir_residual(const int32_t* x, uint32_t n, int32_t* r) { /* :297-309 */
  for (uint32_t i = 0; i < n; ++i) {
    if (i < 2) {
      r[i] = x[i];
      continue;
    }
    int64_t p = (3LL * x[i - 1] - (int64_t)x[i - 2]) >> 2;
    r[i] = (int32_t)(uint32_t)(uint64_t)((int64_t)x[i] - p);
  }
}
static void autocorr(const int32_t* x, uint32_t n, int order, long double* R) { /* lpc.cpp:80-96 */
  for (int k = 0; k <= order; ++k) {
    int64_t s = 0;
    for (uint32_t i = (uint32_t)k; i < n; ++i) s += (int64_t)x[i] * (int64_t)x[i - (uint32_t)k];
    R[k] = (long double)s;
  }
}
static int levinson(const long double* R, int order, long double* a) { /* lpc.cpp:98-154 */
  const long double eps = 1e-8L;
  long double E[33], prevA[33];
  for (int i = 0; i <= order; ++i) {
    E[i] = 0.0L;
    prevA[i] = 0.0L;
    a[i] = 0.0L;
  }
  E[0] = R[0];
  if (!isfinite(E[0]) || E[0] < eps) return 0;
  int achieved = 0;
  for (int i = 1; i <= order; ++i) {
    long double acc = 0.0L;
    for (int j = 1; j < i; ++j) acc += prevA[j] * R[i - j];
    long double den = E[i - 1];
    if (!isfinite(den) || den < eps) break;
    long double ki = (R[i] - acc) / den;
    if (!isfinite(ki)) break;
    if (ki > 0.999L) ki = 0.999L;
    if (ki < -0.999L) ki = -0.999L;
    long double e_new = (1.0L - ki * ki) * E[i - 1];
    if (!isfinite(e_new) || e_new < eps) {
      achieved = i - 1;
      break;
    }
    a[i] = ki;
    for (int j = 1; j < i; ++j) a[j] = prevA[j] - ki * prevA[i - j];
    for (int j = 1; j <= i; ++j) prevA[j] = a[j];
    E[i] = e_new;
    achieved = i;
  }
  return achieved;
}
static int16_t quant_q15(double c) { /* lpc.cpp:73-78 */
  double s = round(c * 32768.0);
  if (s < -32768.0) s = -32768.0;
  if (s > 32767.0) s = 32767.0;
  return (int16_t)s;
}
static int lpc_analyze(const int32_t* x, uint32_t n, int order, int16_t* c) { /* lpc.cpp:156-186 */
  long double R[33], a[33];
  if (n == 0)
    for (int k = 0; k <= order; ++k) R[k] = 0.0L;
  else
    autocorr(x, n, order, R);
  if (R[0] < 1.0L) R[0] = 1.0L;
  int used = levinson(R, order, a);
  c[0] = 0;
  for (int i = 1; i <= order; ++i) c[i] = (i <= used) ? quant_q15((double)a[i]) : 0;
  return used;
}
int lao_lpc_analyze(const int32_t* pcm, uint32_t n, int order, int16_t* coeffs_out) {
  return lpc_analyze(pcm, n, order, coeffs_out);
}
static int lpc_try_residual(const int32_t* x, uint32_t n, const int16_t* c, int order,
                            int32_t* r) { /* lpc.cpp:38-61 */
  for (uint32_t i = 0; i < n; ++i) {
    int64_t acc = 0;
    int taps = order < (int)i ? order : (int)i;
    for (int t = 1; t <= taps; ++t) acc += (int64_t)c[t] * (int64_t)x[i - (uint32_t)t];
    int64_t d = (int64_t)x[i] - (acc >> 15);
    if (d < INT32_MIN || d > INT32_MAX) return 0;
    r[i] = (int32_t)d;
  }
  return 1;
}
/* compute_residual_q15 with used_order_inout (lpc.cpp:188-229): try `start`, then
 * the fallback orders {12,10,8,6,4} below it, then give up with order 0. */
static int lpc_residual(const int32_t* x, uint32_t n, const int16_t* c, int lpc_order, int start,
                        int32_t* r) {
  static const int fallback[5] = {12, 10, 8, 6, 4};
  int maxo = lpc_order;
  if (start > maxo) start = maxo;
  if (start < 0) start = 0;
  int attempts[8], na = 0;
  attempts[na++] = start;
  for (int f = 0; f < 5; ++f)
    if (fallback[f] < start && fallback[f] <= maxo) attempts[na++] = fallback[f];
  if (start != 0) attempts[na++] = 0;
  for (int t = 0; t < na; ++t) {
    if (attempts[t] <= 0) break;
    if (lpc_try_residual(x, n, c, attempts[t], r)) return attempts[t];
  }
  memcpy(r, x, sizeof(int32_t) * n);
  return 0;
}

/* ------------------------------------------------------------------ */
/* Block::Encoder::encode: block/encoder.cpp:313-838                    */
typedef struct {
  int type, order_param, used_order;
  uint64_t rice, zr, bin, stat, best;
  uint32_t k_init, k_stat;
  int has_run;
  int16_t coeffs[33];
  int32_t* res;
} cand;

static void score(cand* c, uint32_t n, int zero_run) { /* :337-351 */
  c->k_init = est_initial_k(c->res, n);
  seg_costs sc = est_costs(c->res, n, c->k_init, 0);
  c->rice = sc.rice;
  c->has_run = sc.has_run;
  c->zr = (zero_run && sc.has_run) ? sc.zr : sc.rice;
  c->bin = sc.bin;
  c->k_stat = est_static_k(c->res, n, &c->stat);
  uint64_t m = c->rice < c->stat ? c->rice : c->stat;
  uint64_t m2 = c->zr < c->bin ? c->zr : c->bin;
  c->best = m < m2 ? m : m2;
}

static uint32_t part_len(uint32_t n, uint32_t p, uint32_t idx) { /* :103-119 */
  if (p == 0) return n;
  uint32_t base = n >> p, cnt = 1u << p;
  return (idx + 1u == cnt) ? n - base * (cnt - 1u) : base;
}

typedef struct {
  uint8_t mode;
  uint32_t k;
  uint64_t bits;
  uint32_t len;
} pchoice;

int lao_block_encode(const int32_t* pcm, uint32_t n, int zero_run, int partitioning, uint8_t** out,
                     uint64_t* out_size, lao_block_info* info) {
  const int max_valid = (n > 1) ? (int)((n - 1 < 32) ? n - 1 : 32) : 0; /* :314-316 */
  int32_t* bufA = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
  int32_t* bufB = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
  cand best, cur;
  int have = 0, slot = 0;
  memset(&best, 0, sizeof best);
  if (info) {
    memset(info, 0, sizeof *info);
    for (int i = 0; i < 11; ++i) info->cand_best_bits[i] = UINT64_MAX;
  }

  /* candidate order: fixed 0..4, FIR, LPC 4,6,8,10,12 (:362-407) */
  for (int ci = 0; ci < 11; ++ci) {
    memset(&cur, 0, sizeof cur);
    cur.res = (have && best.res == bufA) ? bufB : bufA;
    if (ci <= 4) {
      cur.type = PRED_FIXED;
      cur.order_param = ci;
      fixed_residual(pcm, n, ci, cur.res);
    } else if (ci == 5) {
      cur.type = PRED_FIR;
      cur.order_param = 2;
      fir_residual(pcm, n, cur.res);
    } else {
      int co = 4 + 2 * (ci - 6);
      if (co > max_valid) continue;
      cur.type = PRED_LPC;
      cur.order_param = co;
      int used = lpc_analyze(pcm, n, co, cur.coeffs);
      if (used == 0) continue; /* :394-396 */
      cur.used_order = used;
      if (n > 0) cur.used_order = lpc_residual(pcm, n, cur.coeffs, co, used, cur.res);
      if (cur.used_order == 0) continue; /* :402-404 */
    }
    score(&cur, n, zero_run);
    if (info) info->cand_best_bits[ci] = cur.best;
    /* consider (:352-359): strictly fewer bits, or equal bits and lower type */
    if (!have || cur.best < best.best || (cur.best == best.best && cur.type < best.type)) {
      best = cur;
      have = 1;
    }
    (void)slot;
  }
  /* the "no candidate" fallback (:410-417) is unreachable: fixed-0 always scores */

  const int chosen_order = (best.type == PRED_LPC)
                               ? (best.used_order < max_valid
                                      ? (best.used_order > 1 ? best.used_order : 1)
                                      : (max_valid > 1 ? max_valid : 1))
                               : best.order_param; /* :421-423 */
  const int32_t* res = best.res;

  /* base (p = 0) mode, :432-456 */
  const int allow_zr = zero_run && best.has_run;
  uint8_t base_mode = MODE_RICE;
  uint64_t base_bits = best.rice;
  uint32_t base_k = best.k_init;
  if (allow_zr && best.zr <= base_bits) {
    base_bits = best.zr;
    base_mode = MODE_ZR;
  }
  if (best.bin < base_bits) {
    base_bits = best.bin;
    base_mode = MODE_BIN;
  }
  if (best.stat < base_bits) {
    base_bits = best.stat;
    base_mode = MODE_STATIC;
    base_k = best.k_stat;
  }

  pchoice* parts = (pchoice*)malloc(sizeof(pchoice) * LAO_MAX_PARTS);
  pchoice* trial = (pchoice*)malloc(sizeof(pchoice) * LAO_MAX_PARTS);
  uint32_t best_p = 0, nparts = 1;
  parts[0].mode = base_mode;
  parts[0].k = base_k;
  parts[0].bits = base_bits;
  parts[0].len = n;
  uint64_t best_total = base_bits + 8 + 7; /* :475-484 */
  best_total += (8u - (best_total & 7u)) & 7u;

  if (partitioning && n >= MIN_PART) { /* :486-545 */
    uint32_t max_p = 0;
    for (uint32_t p = 1; p <= MAX_PART_ORDER; ++p) {
      if ((n >> p) < MIN_PART) break;
      max_p = p;
    }
    for (uint32_t p = 1; p <= max_p; ++p) {
      uint32_t cnt = 1u << p, off = 0;
      uint64_t sum_bits = 0;
      for (uint32_t s = 0; s < cnt; ++s) {
        uint32_t len = part_len(n, p, s);
        const int32_t* seg = res + off;
        uint32_t ka = est_initial_k(seg, len);
        uint64_t sbits = 0;
        uint32_t ks = est_static_k(seg, len, &sbits);
        seg_costs sc = est_costs(seg, len, ka, 1);
        int azr = zero_run && sc.has_run;
        pchoice pc;
        pc.len = len;
        pc.k = ka;
        pc.mode = MODE_RICE;
        pc.bits = sc.rice;
        if (azr && sc.zr < pc.bits) {
          pc.mode = MODE_ZR;
          pc.bits = sc.zr;
        }
        if (sc.bin < pc.bits) {
          pc.mode = MODE_BIN;
          pc.bits = sc.bin;
        }
        if (sbits < pc.bits || sbits <= pc.bits + pc.bits / 20u) { /* :518, :190-192 */
          pc.mode = MODE_STATIC;
          pc.k = ks;
          pc.bits = sbits;
        }
        sum_bits += pc.bits;
        trial[s] = pc;
        off += len;
      }
      uint64_t total = sum_bits + 8 + 7ull * cnt;
      total += (8u - (total & 7u)) & 7u;
      uint64_t margin = best_total / 20u;
      if (total < best_total || (total <= best_total + margin && best_p == 0) ||
          (total == best_total && p < best_p)) { /* :538-540 */
        best_total = total;
        best_p = p;
        nparts = cnt;
        memcpy(parts, trial, sizeof(pchoice) * cnt);
      }
    }
  }

  /* emission, :554-822 */
  bitsink bw;
  sink_init(&bw);
  const int stateless = best_p > 0;
  uint8_t control = (uint8_t)((parts[0].mode & 3u) << 5);
  if (best_p > 0) control |= (uint8_t)(0x80u | (best_p & 0x0Fu));
  sink_put(&bw, (uint32_t)best.type, 8);
  sink_put(&bw, (uint32_t)chosen_order, 8);
  if (best.type == PRED_LPC)
    for (int i = 1; i <= chosen_order; ++i) sink_put(&bw, (uint16_t)best.coeffs[i], 16);
  sink_put(&bw, control, 8);
  for (uint32_t s = 0; s < nparts; ++s) {
    sink_put(&bw, parts[s].mode, 2);
    sink_put(&bw, parts[s].k, 5);
  }
  uint32_t off = 0;
  for (uint32_t s = 0; s < nparts; ++s) {
    const int32_t* seg = res + off;
    uint32_t len = parts[s].len, k = parts[s].k, count = 0;
    uint64_t sum = 0;
    kstate st;
    memset(&st, 0, sizeof st);
    if (parts[s].mode == MODE_RICE) { /* :585-600 */
      for (uint32_t i = 0; i < len; ++i) {
        rice_put_signed(&bw, seg[i], k);
        sum += zz(seg[i]);
        k = k_next(sum, i + 1, stateless, &st);
      }
    } else if (parts[s].mode == MODE_STATIC) { /* :602-607 */
      for (uint32_t i = 0; i < len; ++i) rice_put_unsigned(&bw, zz(seg[i]), k);
    } else if (parts[s].mode == MODE_BIN) { /* :609-667 */
      for (uint32_t i = 0; i < len; ++i) {
        int32_t v = seg[i];
        if (v == 0) {
          sink_put(&bw, 0, 2);
        } else if (v == 1 || v == -1) {
          sink_put(&bw, 1, 2);
          sink_put1(&bw, v < 0);
        } else if (v == 2 || v == -2) {
          sink_put(&bw, 2, 2);
          sink_put1(&bw, v < 0);
        } else {
          sink_put(&bw, 3, 2);
          rice_put_signed(&bw, v, k);
        }
        sum += zz(v);
        ++count;
        k = k_next(sum, count, stateless, &st);
      }
    } else { /* zero-run, :669-771 */
      uint32_t idx = 0;
      while (idx < len) {
        uint32_t run = 0;
        while (idx + run < len && seg[idx + run] == 0) ++run;
        if (run >= ZR_MIN_RUN) {
          sink_put(&bw, 1, 2);
          rice_put_unsigned(&bw, run - ZR_MIN_RUN, ZR_RUN_K);
          if (stateless) {
            count += run;
            k = k_stateless(sum, count);
          } else {
            for (uint32_t j = 0; j < run; ++j) {
              ++count;
              k = k_stateful(sum, count, &st);
            }
          }
          idx += run;
          continue;
        }
        uint32_t u = zz(seg[idx]);
        uint32_t esc = 1u << (k + 3u < 24u ? k + 3u : 24u);
        if (u > esc) {
          sink_put(&bw, 2, 2);
          sink_put(&bw, u, 32);
        } else {
          sink_put(&bw, 0, 2);
          rice_put_signed(&bw, seg[idx], k);
        }
        sum += u;
        ++count;
        k = k_next(sum, count, stateless, &st);
        ++idx;
      }
    }
    off += len;
  }
  sink_pad(&bw);

  if (info) {
    info->predictor_type = (uint32_t)best.type;
    info->order = (uint32_t)chosen_order;
    memcpy(info->coeffs, best.coeffs, sizeof info->coeffs);
    info->partition_order = best_p;
    info->n_parts = nparts;
    for (uint32_t s = 0; s < nparts; ++s) {
      info->part_mode[s] = parts[s].mode;
      info->part_k[s] = (uint8_t)parts[s].k;
    }
    info->est_total_bits = best_total;
  }
  *out = bw.buf;
  *out_size = bw.nbits >> 3;
  free(parts);
  free(trial);
  free(bufA);
  free(bufB);
  return 0;
}

/* ------------------------------------------------------------------ */
/* Block::Decoder::decode_into: block/decoder.cpp:64-520               */
static int get_rice_u(bitsrc* r, uint32_t k, uint32_t* value) { /* :74-83 */
  if (k > 31u) return 0;
  uint32_t q = 0;
  if (!src_unary(r, UINT32_MAX >> k, &q)) return 0;
  uint32_t rem = (k > 0) ? src_get(r, (int)k) : 0u;
  if (r->error) return 0;
  *value = (q << k) | rem;
  return 1;
}

static int decode_segment(bitsrc* r, uint32_t n, uint32_t k0, uint32_t mode, int32_t* res,
                          int stateless) { /* :104-306 */
  uint32_t k = k0, count = 0;
  uint64_t sum = 0;
  kstate st;
  memset(&st, 0, sizeof st);
  if (mode == MODE_RICE) {
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t u;
      if (!get_rice_u(r, k, &u)) return 0;
      res[i] = unzz(u);
      sum += u;
      ++count;
      k = k_next(sum, count, stateless, &st);
    }
    return 1;
  }
  if (mode == MODE_ZR) {
    uint32_t idx = 0;
    while (idx < n) {
      uint32_t tag = src_get(r, 2);
      if (r->error) return 0;
      if (tag > 2u) return 0;
      if (tag == 0u) {
        uint32_t u;
        if (!get_rice_u(r, k, &u)) return 0; /* the reference `break`s, then idx != samples */
        res[idx++] = unzz(u);
        sum += u;
        ++count;
        k = k_next(sum, count, stateless, &st);
      } else if (tag == 1u) {
        uint32_t enc;
        if (!get_rice_u(r, ZR_RUN_K, &enc) || enc > UINT32_MAX - ZR_MIN_RUN) return 0;
        uint32_t run = enc + ZR_MIN_RUN;
        if (run > n - idx) return 0;
        for (uint32_t j = 0; j < run; ++j) res[idx + j] = 0;
        idx += run;
        if (stateless) {
          count += run;
          k = k_stateless(sum, count);
        } else {
          for (uint32_t j = 0; j < run; ++j) {
            ++count;
            k = k_stateful(sum, count, &st);
          }
        }
      } else {
        uint32_t u = src_get(r, 32);
        if (r->error) return 0;
        int32_t v = unzz(u);
        res[idx++] = v;
        sum += zz(v);
        ++count;
        k = k_next(sum, count, stateless, &st);
      }
    }
    return idx == n;
  }
  if (mode == MODE_BIN) {
    uint32_t idx = 0;
    while (idx < n) {
      uint32_t tag = src_get(r, 2);
      if (r->error) return 0;
      int32_t v = 0;
      uint32_t u = 0;
      if (tag == 0u) {
        v = 0;
        u = 0;
      } else if (tag == 1u || tag == 2u) {
        uint32_t sign = src_get(r, 1);
        if (r->error) return 0;
        v = (tag == 1u) ? (sign ? -1 : 1) : (sign ? -2 : 2);
        u = zz(v);
      } else {
        if (!get_rice_u(r, k, &u)) return 0;
        v = unzz(u);
      }
      res[idx++] = v;
      sum += u;
      ++count;
      k = k_next(sum, count, stateless, &st);
    }
    return 1;
  }
  if (mode == MODE_STATIC) {
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t u;
      if (!get_rice_u(r, k0, &u)) return 0;
      res[i] = unzz(u);
    }
    return 1;
  }
  return 0;
}

static int restore(int32_t* x, uint32_t n, uint32_t type, int order, const int16_t* c) {
  /* decoder.cpp:308-403: in place, every reconstructed sample must fit int32 */
  if (type == PRED_FIXED) {
    if (order == 0) return 1;
    for (uint32_t i = (uint32_t)order; i < n; ++i) {
      int64_t p;
      switch (order) {
        case 1: p = x[i - 1]; break;
        case 2: p = 2LL * x[i - 1] - x[i - 2]; break;
        case 3: p = 3LL * x[i - 1] - 3LL * x[i - 2] + x[i - 3]; break;
        case 4: p = 4LL * x[i - 1] - 6LL * x[i - 2] + 4LL * x[i - 3] - x[i - 4]; break;
        default: return 0;
      }
      int64_t s = (int64_t)x[i] + p;
      if (s < INT32_MIN || s > INT32_MAX) return 0;
      x[i] = (int32_t)s;
    }
    return 1;
  }
  if (type == PRED_FIR) {
    for (uint32_t i = 2; i < n; ++i) {
      int64_t p = (3LL * x[i - 1] - (int64_t)x[i - 2]) >> 2;
      int64_t s = (int64_t)x[i] + p;
      if (s < INT32_MIN || s > INT32_MAX) return 0;
      x[i] = (int32_t)s;
    }
    return 1;
  }
  for (uint32_t i = 0; i < n; ++i) {
    int64_t acc = 0;
    int taps = order < (int)i ? order : (int)i;
    for (int t = 1; t <= taps; ++t) acc += (int64_t)c[t] * (int64_t)x[i - (uint32_t)t];
    int64_t s = (acc >> 15) + (int64_t)x[i];
    if (s < INT32_MIN || s > INT32_MAX) return 0;
    x[i] = (int32_t)s;
  }
  return 1;
}

static int block_decode(bitsrc* r, uint32_t n, int32_t* out) {
  if (n == 0 || n > LAO_MAX_BLOCK || !out) return 0;
  uint32_t type = src_get(r, 8);
  int order = (int)src_get(r, 8);
  if (r->error) return 0;
  if (type > 2u) return 0;
  if (type == PRED_LPC) {
    if (order <= 0 || order > 32 || (uint32_t)order >= n) return 0;
  } else if (type == PRED_FIR) {
    if (order != 2) return 0;
  } else if (order < 0 || order > 4) {
    return 0;
  }
  int16_t c[33];
  memset(c, 0, sizeof c);
  if (type == PRED_LPC)
    for (int i = 1; i <= order; ++i) {
      c[i] = (int16_t)(uint16_t)src_get(r, 16);
      if (r->error) return 0;
    }
  uint32_t control = src_get(r, 8);
  if (r->error) return 0;
  if (control & 0x10u) return 0;
  int pflag = (control & 0x80u) != 0;
  uint32_t p = control & 0x0Fu, cmode = (control >> 5) & 3u;
  if (pflag && p == 0) return 0;
  if (!pflag && p != 0) return 0;
  if (p > MAX_PART_ORDER) return 0;
  if (p > 0 && (n >> p) < MIN_PART) return 0;
  uint32_t cnt = (p == 0) ? 1u : (1u << p);
  if (part_len(n, p, cnt - 1u) == 0) return 0;
  uint8_t modes[LAO_MAX_PARTS];
  uint32_t ks[LAO_MAX_PARTS];
  for (uint32_t i = 0; i < cnt; ++i) {
    modes[i] = (uint8_t)src_get(r, 2);
    ks[i] = src_get(r, 5);
    if (r->error) return 0;
  }
  if (modes[0] != cmode) return 0;
  uint32_t off = 0;
  for (uint32_t i = 0; i < cnt; ++i) {
    uint32_t len = part_len(n, p, i);
    if (!decode_segment(r, len, ks[i], modes[i], out + off, p > 0)) return 0;
    off += len;
  }
  if (off != n) return 0;
  if (!src_skip_zero_pad(r)) return 0;
  return restore(out, n, type, order, c);
}

int lao_block_decode(const uint8_t* data, uint64_t size, uint32_t block_size, int32_t* out,
                     uint64_t* bits_consumed) {
  bitsrc r = {data, size * 8u, 0, 0};
  int ok = block_decode(&r, block_size, out);
  if (bits_consumed) *bits_consumed = ok ? r.pos : 0;
  return ok;
}

/* ------------------------------------------------------------------ */
/* stereo proxy: lac/encoder.cpp:31-57,114-197                          */
static uint64_t sat_add(uint64_t a, uint64_t b) { return (b > UINT64_MAX - a) ? UINT64_MAX : a + b; }
static uint64_t zz64(int64_t v) {
  return v >= 0 ? ((uint64_t)v << 1) : ((((uint64_t)(-(v + 1))) << 1) | 1u);
}
static uint64_t proxy_bits(uint64_t sum, uint64_t count) { /* :42-57 */
  if (count == 0) return 0;
  uint64_t mean = (sum + (count >> 1)) / count;
  uint32_t k = 0;
  while (k < 31u && ((uint64_t)1 << k) < mean) ++k;
  return sat_add(sum >> k, count * (uint64_t)(k + 1u));
}
uint32_t lao_stereo_proxy(const int32_t* L, const int32_t* R, uint32_t n) { /* :126-197 */
  uint64_t raw[4] = {0, 0, 0, 0}, dif[4] = {0, 0, 0, 0}, anti[4] = {0, 0, 0, 0};
  int64_t prev[4] = {0, 0, 0, 0};
  for (uint32_t i = 0; i < n; ++i) {
    int64_t v[4];
    v[0] = L[i];
    v[1] = R[i];
    v[2] = (v[0] + v[1]) >> 1;
    v[3] = v[0] - v[1];
    for (int c = 0; c < 4; ++c) {
      raw[c] = sat_add(raw[c], zz64(v[c]));
      if (i == 0) {
        dif[c] = zz64(v[c]);
        anti[c] = dif[c];
      } else {
        dif[c] = sat_add(dif[c], zz64(v[c] - prev[c]));
        anti[c] = sat_add(anti[c], zz64(v[c] + prev[c]));
      }
      prev[c] = v[c];
    }
  }
  uint64_t bits[4];
  int nondiff = 0;
  for (int c = 0; c < 4; ++c) {
    uint64_t rb = proxy_bits(raw[c], n), db = proxy_bits(dif[c], n), ab = proxy_bits(anti[c], n);
    uint64_t m = rb < db ? rb : db;
    bits[c] = m < ab ? m : ab;
    if (rb < db || ab < db) nondiff = 1;
  }
  uint64_t lr = sat_add(bits[0], bits[1]), ms = sat_add(bits[2], bits[3]);
  uint64_t smaller = lr < ms ? lr : ms;
  uint64_t diff = lr >= ms ? lr - ms : ms - lr;
  uint32_t choose_ms = ms < lr;
  uint32_t uncertain = smaller == 0 || diff == 0 || nondiff || diff <= smaller / 100u;
  return choose_ms | (uncertain << 1);
}

/* ------------------------------------------------------------------ */
/* LAC::Encoder::encode: lac/encoder.cpp:215-466                        */
typedef struct {
  uint8_t* p;
  uint64_t n;
} blob;

static void blob_append(blob* b, const uint8_t* src, uint64_t n) {
  b->p = (uint8_t*)realloc(b->p, b->n + n + 1);
  memcpy(b->p + b->n, src, n);
  b->n += n;
}
static void ms_split(const int32_t* L, const int32_t* R, uint32_t n, int32_t* M, int32_t* S) {
  /* simd/neon.cpp:14-30 */
  for (uint32_t i = 0; i < n; ++i) {
    int32_t sum = (int32_t)((uint32_t)L[i] + (uint32_t)R[i]);
    M[i] = sum >> 1;
    S[i] = (int32_t)((uint32_t)L[i] - (uint32_t)R[i]);
  }
}
static blob enc_pair(const int32_t* a, const int32_t* b, uint32_t n, int zr, int part) {
  blob o = {NULL, 0};
  uint8_t* t;
  uint64_t tn;
  lao_block_encode(a, n, zr, part, &t, &tn, NULL);
  blob_append(&o, t, tn);
  free(t);
  if (b) {
    lao_block_encode(b, n, zr, part, &t, &tn, NULL);
    blob_append(&o, t, tn);
    free(t);
  }
  return o;
}
static blob enc_lr(const int32_t* L, const int32_t* R, uint64_t start, uint32_t n, int zr, int part) {
  return enc_pair(L + start, R ? R + start : NULL, n, zr, part);
}
static blob enc_ms(const int32_t* L, const int32_t* R, uint64_t start, uint32_t n, int zr, int part) {
  int32_t* M = (int32_t*)malloc(sizeof(int32_t) * n);
  int32_t* S = (int32_t*)malloc(sizeof(int32_t) * n);
  ms_split(L + start, R + start, n, M, S);
  blob o = enc_pair(M, S, n, zr, part);
  free(M);
  free(S);
  return o;
}

typedef struct {
  const int32_t *L, *R;
  uint64_t frames;
  uint32_t mode; /* effective stereo mode */
  int zr, part;
  uint32_t nblocks;
  blob* outs;
  uint32_t next;
  pthread_mutex_t mu;
} enc_job;

static blob encode_one_block(const enc_job* j, uint32_t bi) { /* encode_block lambda :270-383 */
  uint64_t start = (uint64_t)bi * LAO_MAX_BLOCK;
  uint32_t n = (uint32_t)((j->frames - start < LAO_MAX_BLOCK) ? j->frames - start : LAO_MAX_BLOCK);
  if (!j->R) return enc_lr(j->L, NULL, start, n, j->zr, j->part);
  if (j->mode == 1) return enc_ms(j->L, j->R, start, n, j->zr, j->part);
  if (j->mode == 0) return enc_lr(j->L, j->R, start, n, j->zr, j->part);
  uint32_t d = lao_stereo_proxy(j->L + start, j->R + start, n);
  int choose_ms = d & 1u;
  blob sel = {NULL, 0};
  if (d & 2u) {
    if (n <= 4096u) {
      blob lr = enc_lr(j->L, j->R, start, n, j->zr, j->part);
      blob ms = enc_ms(j->L, j->R, start, n, j->zr, j->part);
      choose_ms = ms.n < lr.n;
      if (choose_ms) {
        sel = ms;
        free(lr.p);
      } else {
        sel = lr;
        free(ms.p);
      }
    } else {
      uint64_t ps[3] = {start, start + (n - 256u) / 2u, start + n - 256u};
      uint64_t lrs = 0, mss = 0;
      for (int t = 0; t < 3; ++t) {
        blob a = enc_lr(j->L, j->R, ps[t], 256, j->zr, j->part);
        blob b = enc_ms(j->L, j->R, ps[t], 256, j->zr, j->part);
        lrs += a.n;
        mss += b.n;
        free(a.p);
        free(b.p);
      }
      choose_ms = mss < lrs;
    }
  }
  blob o = {NULL, 0};
  uint8_t flag = (uint8_t)(choose_ms ? 1 : 0);
  blob_append(&o, &flag, 1);
  if (!sel.p) sel = choose_ms ? enc_ms(j->L, j->R, start, n, j->zr, j->part)
                              : enc_lr(j->L, j->R, start, n, j->zr, j->part);
  blob_append(&o, sel.p, sel.n);
  free(sel.p);
  return o;
}
static void* enc_worker(void* arg) {
  enc_job* j = (enc_job*)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    uint32_t bi = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (bi >= j->nblocks) return NULL;
    j->outs[bi] = encode_one_block(j, bi);
  }
}
static void put_be32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24);
  p[1] = (uint8_t)(v >> 16);
  p[2] = (uint8_t)(v >> 8);
  p[3] = (uint8_t)v;
}
static int depth_ok(int64_t s, uint32_t depth) { /* lac/encoder.cpp:85-93 */
  if (depth == 16) return s >= -32768 && s <= 32767;
  if (depth == 24) return s >= -0x800000 && s <= 0x7FFFFF;
  return 0;
}

int lao_encode(const int32_t* left, const int32_t* right, uint64_t frames, uint32_t sample_rate,
               uint32_t bit_depth, uint32_t stereo_mode, int zero_run, int partitioning,
               uint32_t threads, uint8_t** out, uint64_t* out_size) {
  /* argument checks :220-241 */
  if (!left || frames == 0) return -1;
  if (!(sample_rate == 44100 || sample_rate == 48000 || sample_rate == 96000 ||
        sample_rate == 192000))
    return -1;
  if (bit_depth != 16 && bit_depth != 24) return -1;
  if (stereo_mode > 2) return -1;
  for (uint64_t i = 0; i < frames; ++i) {
    if (!depth_ok(left[i], bit_depth)) return -1;
    if (right && !depth_ok(right[i], bit_depth)) return -1;
  }
  enc_job j;
  memset(&j, 0, sizeof j);
  j.L = left;
  j.R = right;
  j.frames = frames;
  j.mode = right ? stereo_mode : 0;
  j.zr = zero_run;
  j.part = partitioning;
  j.nblocks = (uint32_t)((frames + LAO_MAX_BLOCK - 1) / LAO_MAX_BLOCK); /* plan_blocks :59-69 */
  j.outs = (blob*)calloc(j.nblocks, sizeof(blob));
  pthread_mutex_init(&j.mu, NULL);
  uint32_t nt = threads ? threads : 1;
  if (nt > j.nblocks) nt = j.nblocks;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  for (uint32_t t = 0; t < nt; ++t) pthread_create(&th[t], NULL, enc_worker, &j);
  for (uint32_t t = 0; t < nt; ++t) pthread_join(th[t], NULL);

  uint64_t total = 10 + 4 + 8ull * j.nblocks;
  for (uint32_t b = 0; b < j.nblocks; ++b) total += j.outs[b].n;
  uint8_t* o = (uint8_t*)malloc(total);
  /* frame header, frame/frame_header.hpp:25-36 */
  o[0] = 0x4C;
  o[1] = 0x41;
  o[2] = 3;
  o[3] = right ? 2 : 1;
  o[4] = (uint8_t)j.mode;
  o[5] = (uint8_t)((sample_rate >> 8) & 0xFF);
  o[6] = (uint8_t)(sample_rate & 0xFF);
  o[7] = (uint8_t)((sample_rate >> 16) & 0xFF);
  o[8] = (uint8_t)bit_depth;
  o[9] = 0;
  put_be32(o + 10, j.nblocks); /* block table :445-453 */
  uint64_t pos = 14 + 8ull * j.nblocks;
  for (uint32_t b = 0; b < j.nblocks; ++b) {
    uint64_t start = (uint64_t)b * LAO_MAX_BLOCK;
    uint32_t n = (uint32_t)((frames - start < LAO_MAX_BLOCK) ? frames - start : LAO_MAX_BLOCK);
    put_be32(o + 14 + 8ull * b, n);
    put_be32(o + 18 + 8ull * b, (uint32_t)j.outs[b].n);
    memcpy(o + pos, j.outs[b].p, j.outs[b].n);
    pos += j.outs[b].n;
    free(j.outs[b].p);
  }
  free(j.outs);
  pthread_mutex_destroy(&j.mu);
  *out = o;
  *out_size = total;
  return 0;
}

/* ------------------------------------------------------------------ */
/* LAC::Decoder::decode: lac/decoder.cpp:76-303                         */
static uint32_t get_be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
#define DEC_FAIL(msg)                                  \
  do {                                                 \
    snprintf(g_err, sizeof g_err, "[decode-error] %s", msg); \
    goto fail;                                         \
  } while (0)

static int decode_frame_block(bitsrc* r, uint32_t n, uint32_t channels, uint32_t smode,
                              uint32_t depth, int32_t* l, int32_t* rr, uint32_t bi) {
  /* decode_block lambda :167-207 */
  int mid_side = 0;
  if (channels == 2 && smode == 2) {
    uint32_t flag = src_get(r, 8);
    if (r->error || flag > 1u) {
      snprintf(g_err, sizeof g_err, "[decode-error] invalid per-block stereo flag");
      return 0;
    }
    mid_side = flag == 1u;
  } else if (channels == 2 && smode == 1) {
    mid_side = 1;
  }
  if (!block_decode(r, n, l)) {
    snprintf(g_err, sizeof g_err, "[decode-error] block=%u channel=primary", bi);
    return 0;
  }
  if (channels == 2 && !block_decode(r, n, rr)) {
    snprintf(g_err, sizeof g_err, "[decode-error] block=%u channel=secondary", bi);
    return 0;
  }
  int ok = 1;
  if (channels == 1) {
    for (uint32_t i = 0; i < n; ++i) ok &= depth_ok(l[i], depth);
  } else if (mid_side) { /* reconstruct_mid_side_in_place :48-65 */
    for (uint32_t i = 0; i < n; ++i) {
      int64_t m = l[i], s = rr[i];
      int64_t L = m + ((s + (s & 1)) >> 1), R = L - s;
      if (!depth_ok(L, depth) || !depth_ok(R, depth)) {
        ok = 0;
        break;
      }
      l[i] = (int32_t)L;
      rr[i] = (int32_t)R;
    }
  } else {
    for (uint32_t i = 0; i < n; ++i) ok &= depth_ok(l[i], depth) & depth_ok(rr[i], depth);
  }
  if (!ok) {
    snprintf(g_err, sizeof g_err, "[decode-error] decoded sample outside PCM bit depth");
    return 0;
  }
  return 1;
}

int lao_decode(const uint8_t* data, uint64_t size, uint32_t threads, int32_t** left,
               int32_t** right, uint64_t* frames, uint32_t* channels, uint32_t* sample_rate,
               uint32_t* bit_depth, uint32_t* stereo_mode) {
  (void)threads;
  int32_t *L = NULL, *R = NULL;
  uint32_t *bsz = NULL, *bby = NULL;
  *left = *right = NULL;
  *frames = 0;
  g_err[0] = 0;
  if (!data || size == 0) DEC_FAIL("empty input");
  if (size < 10) DEC_FAIL("invalid frame header");
  /* FrameHeader::read/validate, frame_header.hpp:38-59 */
  uint32_t sync = ((uint32_t)data[0] << 8) | data[1], ver = data[2], ch = data[3], sm = data[4];
  uint32_t sr = (((uint32_t)data[5] << 8) | data[6]) | ((uint32_t)data[7] << 16);
  uint32_t depth = data[8], reserved = data[9];
  if (sync != 0x4C41 || (ver != 2 && ver != 3) || (ch != 1 && ch != 2) || (ch == 1 && sm != 0) ||
      sm > 2 || !(sr == 44100 || sr == 48000 || sr == 96000 || sr == 192000) ||
      (depth != 16 && depth != 24) || reserved != 0)
    DEC_FAIL("invalid frame header");
  const uint8_t* payload = data + 10;
  uint64_t pbytes = size - 10;
  if (pbytes < 4) DEC_FAIL("invalid block count");
  uint32_t nb = get_be32(payload);
  const uint32_t max_blocks = (uint32_t)(((1ull << 30) / 4 + 255) / 256);
  if (nb == 0 || nb > max_blocks) DEC_FAIL("invalid block count");
  uint32_t words = (ver >= 3) ? 2u : 1u;
  if (nb > ((pbytes - 4) * 8u) / (32u * words)) DEC_FAIL("truncated block size table");
  bsz = (uint32_t*)malloc(sizeof(uint32_t) * nb);
  bby = (uint32_t*)calloc(nb, sizeof(uint32_t));
  uint64_t total = 0, total_bytes = 0, tp = 4;
  for (uint32_t i = 0; i < nb; ++i) {
    uint32_t s = get_be32(payload + tp);
    tp += 4;
    if (s == 0 || s > LAO_MAX_BLOCK || (i + 1u < nb && s < 256u)) DEC_FAIL("invalid block size");
    total += s;
    if (total > 6912000000ull) DEC_FAIL("total samples exceed maximum");
    bsz[i] = s;
    if (ver >= 3) {
      uint32_t b = get_be32(payload + tp);
      tp += 4;
      if (b == 0) DEC_FAIL("invalid compressed block size");
      total_bytes += b;
      if (total_bytes > pbytes) DEC_FAIL("compressed block sizes exceed frame payload");
      bby[i] = b;
    }
  }
  if (total * ch * 4u > (1ull << 30)) DEC_FAIL("decoded PCM allocation exceeds maximum");
  {
    uint64_t wav = total * ch * (depth / 8u);
    if (36u + wav + (wav & 1u) > 0xFFFFFFFFull) DEC_FAIL("decoded WAV data exceeds RIFF limit");
  }
  L = (int32_t*)calloc(total + 1, sizeof(int32_t));
  if (ch == 2) R = (int32_t*)calloc(total + 1, sizeof(int32_t));
  if (ver < 3) { /* serial v2 :209-218 */
    bitsrc r = {payload + tp, (pbytes - tp) * 8u, 0, 0};
    uint64_t off = 0;
    for (uint32_t i = 0; i < nb; ++i) {
      if (!decode_frame_block(&r, bsz[i], ch, sm, depth, L + off, R ? R + off : NULL, i)) goto fail;
      off += bsz[i];
    }
    if (src_left(&r) != 0) DEC_FAIL("trailing frame payload");
  } else {
    uint64_t avail = pbytes - tp;
    if (total_bytes != avail) DEC_FAIL("compressed block sizes do not match frame payload");
    uint64_t off = 0, boff = 0;
    for (uint32_t i = 0; i < nb; ++i) {
      bitsrc r = {payload + tp + boff, (uint64_t)bby[i] * 8u, 0, 0};
      if (!decode_frame_block(&r, bsz[i], ch, sm, depth, L + off, R ? R + off : NULL, i)) goto fail;
      if (src_left(&r) != 0) {
        snprintf(g_err, sizeof g_err, "[decode-error] block=%u channel=trailing-payload", i);
        goto fail;
      }
      off += bsz[i];
      boff += bby[i];
    }
  }
  free(bsz);
  free(bby);
  *left = L;
  *right = R;
  *frames = total;
  *channels = ch;
  *sample_rate = sr;
  *bit_depth = depth;
  *stereo_mode = sm;
  return 0;
fail:
  free(bsz);
  free(bby);
  free(L);
  free(R);
  return -1;
}
