/* tools/lac_synth.c -- deterministic synthetic PCM generator (benchmark/test input).
 *
 * Implements the integer-only sectioned signal of SURVEY.md Appendix C: two LCG
 * streams drive four 2^17-sample sections (AR(4) noise / triangle / sparse
 * digital silence / level-stepped white noise) so that every predictor,
 * partition order, residual mode and the uncertain-stereo probe path fire.
 * This is input generation only: it is neither the codec nor the oracle.
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

/* Generates `frames` frames starting at frame 0 (the generator is sequential in i).
 * Any of left/right/packed may be NULL.  packed receives interleaved little-endian
 * depth/8-byte samples (channels = 1 or 2; mono emits the left stream only). */
void lac_synth(uint32_t seed, uint64_t frames, int depth, int channels,
               int32_t* left, int32_t* right, uint8_t* packed) {
  uint32_t g = seed, g2 = seed ^ 0x9E3779B9u;
  int64_t y1 = 0, y2 = 0, y3 = 0, y4 = 0, z1 = 0, z2 = 0, tri = 0, dir = 1;
  const int bps = depth / 8;
  for (uint64_t i = 0; i < frames; ++i) {
    const int sec = (int)((i >> 17) & 3u);
    const int64_t w = n16(lcg(&g)), v = n16(lcg(&g2));
    int64_t l = 0, r = 0;
    if (sec == 0) {
      int64_t y = ((29491 * y1 - 19661 * y2 + 9830 * y3 - 6554 * y4) >> 15) + (w >> 3);
      y4 = y3; y3 = y2; y2 = y1; y1 = y;
      l = y;
      int64_t z = ((24576 * z1 - 8192 * z2) >> 15) + (v >> 4);
      z2 = z1; z1 = z;
      r = l + z;
    } else if (sec == 1) {
      tri += dir * 37;
      if (tri > 12000) dir = -1;
      if (tri < -12000) dir = 1;
      l = tri + (w >> 13);
      r = tri / 2 + (v >> 13);
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
      l = l * 256 + (sec == 2 ? 0 : (int64_t)(lcg(&g) >> 24));
      r = r * 256 + (sec == 2 ? 0 : (int64_t)(lcg(&g2) >> 24));
    }
    const int32_t L = clampd(l, depth), R = clampd(r, depth);
    if (left) left[i] = L;
    if (right) right[i] = R;
    if (packed) {
      uint8_t* p = packed + i * (uint64_t)(bps * channels);
      for (int b = 0; b < bps; ++b) p[b] = (uint8_t)((uint32_t)L >> (8 * b));
      if (channels == 2)
        for (int b = 0; b < bps; ++b) p[bps + b] = (uint8_t)((uint32_t)R >> (8 * b));
    }
  }
}
