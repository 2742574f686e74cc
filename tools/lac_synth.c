/* tools/lac_synth.c -- deterministic synthetic PCM generator (benchmark/test input).
 *
 * Implements the integer-only sectioned signal of SURVEY.md Appendix C: two LCG
 * streams drive four 2^17-sample sections (AR(4) noise / triangle / sparse
 * digital silence / level-stepped white noise) so that every predictor,
 * partition order, residual mode and the uncertain-stereo probe path fire.
 * This is input generation only: it is neither the codec nor the oracle.
 *
 * Two entry points:
 *   lac_synth        the Appendix C stream from frame 0 (filter / triangle state runs through
 *                    the whole file): configs 1, 2, 3, 5 and every short test input.
 *   lac_synth_range  frames [f0, f0 + frames) of the RANGE-ADDRESSABLE variant used for config 4
 *                    (one 10 h file cut into block ranges, one per GPU): the same stream, except
 *                    that the AR / side-filter / triangle state is reset to zero at every
 *                    2^reset_log2-frame boundary ("super-section", 2^19 = the four sections once,
 *                    SURVEY.md Appendix C last paragraph).  The two LCG streams are NOT reset: they
 *                    are advanced to the range start by a closed-form jump (the number of draws
 *                    before a frame is a fixed function of the frame index).  Any rank can therefore
 *                    produce its range of the one file without generating what precedes it, and the
 *                    result equals the corresponding slice of the whole-file generation
 *                    (tests/test_synth_range.py).  With reset_log2 = 0 there is no reset and the
 *                    range is produced by running the recurrences from frame 0 (slow, exact).
 *                    For inputs shorter than 2^reset_log2 both entry points agree.
 */
#include <stdint.h>
#include <stddef.h>

static inline uint32_t lcg(uint32_t* s) {
  *s = *s * 1664525u + 1013904223u;
  return *s;
}
static inline int32_t n16(uint32_t s) { return (int32_t)((s >> 8) & 0xFFFFu) - 32768; }
static inline int32_t clampd(int64_t v, int depth) {
  int64_t lo = depth == 16 ? -32768 : -0x800000, hi = depth == 16 ? 32767 : 0x7FFFFF;
  return (int32_t)(v < lo ? lo : (v > hi ? hi : v));
}

/* state after n LCG steps from s: the n-fold composition of x -> a x + c, by squaring */
static uint32_t lcg_jump(uint32_t s, uint64_t n) {
  uint32_t a = 1664525u, c = 1013904223u;  /* current power of the map */
  uint32_t ra = 1u, rc = 0u;               /* accumulated map */
  while (n) {
    if (n & 1u) {
      ra = ra * a;
      rc = rc * a + c;
    }
    c = c * a + c;
    a = a * a;
    n >>= 1;
  }
  return ra * s + rc;
}

/* LCG draws (per stream) consumed by the frames before frame f: one per frame, plus the dither
 * draw of 24-bit output in every section but the third (sec 2) */
static uint64_t draws_before(uint64_t f, int depth) {
  if (depth != 24) return f;
  const uint64_t cyc = f >> 19, rem = f & ((1ull << 19) - 1ull);
  const uint64_t sec = rem >> 17, in = rem & ((1ull << 17) - 1ull);
  uint64_t dither = cyc * 3ull * (1ull << 17);
  dither += (sec < 2 ? sec : sec - 1) * (1ull << 17);   /* whole sections before this one, sec 2 excluded */
  if (sec != 2) dither += in;
  return f + dither;
}

typedef struct {
  uint32_t g, g2;
  int64_t y1, y2, y3, y4, z1, z2, tri, dir;
} SynthState;

static inline void synth_reset_filters(SynthState* st) {
  st->y1 = st->y2 = st->y3 = st->y4 = st->z1 = st->z2 = st->tri = 0;
  st->dir = 1;
}

static inline void synth_frame(SynthState* st, uint64_t i, int depth, int32_t* Lo, int32_t* Ro) {
  const int sec = (int)((i >> 17) & 3u);
  const int64_t w = n16(lcg(&st->g)), v = n16(lcg(&st->g2));
  int64_t l = 0, r = 0;
  if (sec == 0) {
    int64_t y = ((29491 * st->y1 - 19661 * st->y2 + 9830 * st->y3 - 6554 * st->y4) >> 15) + (w >> 3);
    st->y4 = st->y3; st->y3 = st->y2; st->y2 = st->y1; st->y1 = y;
    l = y;
    int64_t z = ((24576 * st->z1 - 8192 * st->z2) >> 15) + (v >> 4);
    st->z2 = st->z1; st->z1 = z;
    r = l + z;
  } else if (sec == 1) {
    st->tri += st->dir * 37;
    if (st->tri > 12000) st->dir = -1;
    if (st->tri < -12000) st->dir = 1;
    l = st->tri + (w >> 13);
    r = st->tri / 2 + (v >> 13);
  } else if (sec == 2) {
    const uint32_t ph = (uint32_t)(i & 1023u);
    l = ph < 512 ? 0 : (((w & 7) == 0) ? ((w >> 3) & 3) - 1 : 0);
    r = ph < 768 ? 0 : (((v & 7) == 0) ? ((v >> 3) & 3) - 2 : 0);
  } else {
    const int lv = (int)((i >> 11) & 7u);
    l = w >> lv;
    r = v >> (7 - lv);
  }
  if (depth == 24) {
    l = l * 256 + (sec == 2 ? 0 : (int64_t)(lcg(&st->g) >> 24));
    r = r * 256 + (sec == 2 ? 0 : (int64_t)(lcg(&st->g2) >> 24));
  }
  *Lo = clampd(l, depth);
  *Ro = clampd(r, depth);
}

static inline void synth_store(uint64_t o, int32_t L, int32_t R, int depth, int channels, int32_t* left,
                               int32_t* right, uint8_t* packed) {
  const int bps = depth / 8;
  if (left) left[o] = L;
  if (right) right[o] = R;
  if (packed) {
    uint8_t* p = packed + o * (uint64_t)(bps * channels);
    for (int b = 0; b < bps; ++b) p[b] = (uint8_t)((uint32_t)L >> (8 * b));
    if (channels == 2)
      for (int b = 0; b < bps; ++b) p[bps + b] = (uint8_t)((uint32_t)R >> (8 * b));
  }
}

/* Generates `frames` frames starting at frame 0 (the generator is sequential in i).
 * Any of left/right/packed may be NULL.  packed receives interleaved little-endian
 * depth/8-byte samples (channels = 1 or 2; mono emits the left stream only). */
void lac_synth(uint32_t seed, uint64_t frames, int depth, int channels,
               int32_t* left, int32_t* right, uint8_t* packed) {
  SynthState st;
  st.g = seed;
  st.g2 = seed ^ 0x9E3779B9u;
  synth_reset_filters(&st);
  for (uint64_t i = 0; i < frames; ++i) {
    int32_t L, R;
    synth_frame(&st, i, depth, &L, &R);
    synth_store(i, L, R, depth, channels, left, right, packed);
  }
}

/* Frames [f0, f0 + frames) of the range-addressable stream (see the file comment).  Output index 0
 * is frame f0. */
void lac_synth_range(uint32_t seed, uint64_t f0, uint64_t frames, int depth, int channels, int reset_log2,
                     int32_t* left, int32_t* right, uint8_t* packed) {
  SynthState st;
  synth_reset_filters(&st);
  uint64_t start = 0;  /* first frame actually generated: the last reset point at or before f0 */
  if (reset_log2 >= 19) start = (f0 >> reset_log2) << reset_log2;  /* whole section cycles only */
  else reset_log2 = 0;
  st.g = lcg_jump(seed, draws_before(start, depth));
  st.g2 = lcg_jump(seed ^ 0x9E3779B9u, draws_before(start, depth));
  const uint64_t mask = reset_log2 ? ((1ull << reset_log2) - 1ull) : ~0ull;
  for (uint64_t i = start; i < f0 + frames; ++i) {
    if (reset_log2 && (i & mask) == 0) synth_reset_filters(&st);
    int32_t L, R;
    synth_frame(&st, i, depth, &L, &R);
    if (i >= f0) synth_store(i - f0, L, R, depth, channels, left, right, packed);
  }
}
