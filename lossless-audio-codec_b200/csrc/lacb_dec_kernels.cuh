// lacb_dec_kernels.cuh -- decoder kernels (sm_100a).
//
// The .lac block bitstream is inherently serial inside a block: variable-length
// codes whose Rice parameter adapts to every decoded value, and no byte offsets for
// partitions or for the second channel (docs/format.md:18-27).  The
// parallelism is the block count, so the parser runs one thread per frame-block
// (32 independent streams per warp, each lane with its own 64-bit bit window), and
// everything that is data parallel (mid/side reconstruction, PCM range validation,
// interleave + 16/24-bit packing) is a separate bandwidth-bound kernel.
//   Block::Decoder::decode_into   src/codec/block/decoder.cpp:64-520
//   LAC::Decoder decode_block     src/codec/lac/decoder.cpp:167-207
//   reconstruct_mid_side_in_place src/codec/lac/decoder.cpp:48-65
//   pack_pcm_to_wav_bytes         src/main.cpp:150-182
#pragma once
#include "lacb_common.cuh"

namespace lacb {

enum : uint32_t {
  DERR_OK = 0,
  DERR_FLAG = 1,       // "invalid per-block stereo flag"
  DERR_PRIMARY = 2,    // "block=<i> channel=primary"
  DERR_SECONDARY = 3,  // "block=<i> channel=secondary"
  DERR_RANGE = 4,      // "decoded sample outside PCM bit depth"
  DERR_TRAILING = 5,   // "block=<i> channel=trailing-payload"
};

// MSB-first bit reader over [p, end); reads past the end deliver zeros and are
// detected through the consumed-bit count (BitReader, bitstream/bit_reader.hpp:40-202).
struct BitSrc {
  const uint8_t* p;    // next byte to load
  const uint8_t* end;
  u64 w;               // window, next bit at bit 63
  int avail;           // valid bits in w
  u64 consumed;        // bits handed out so far
};
__device__ __forceinline__ void bs_init(BitSrc& s, const uint8_t* p, const uint8_t* end) {
  s.p = p;
  s.end = end;
  s.w = 0ull;
  s.avail = 0;
  s.consumed = 0ull;
}
__device__ __forceinline__ void bs_refill(BitSrc& s) {
  while (s.avail <= 56) {
    const u64 byte = (s.p < s.end) ? (u64)(*s.p) : 0ull;
    s.p++;
    s.w |= byte << (56 - s.avail);
    s.avail += 8;
  }
}
// n in [0,32]
__device__ __forceinline__ uint32_t bs_get(BitSrc& s, uint32_t n) {
  if (n == 0u) return 0u;
  if (s.avail < (int)n) bs_refill(s);
  const uint32_t v = (uint32_t)(s.w >> (64u - n));
  s.w <<= n;
  s.avail -= (int)n;
  s.consumed += n;
  return v;
}
__device__ __forceinline__ u64 bs_size_bits(const BitSrc& s, const uint8_t* begin) { return (u64)(s.end - begin) * 8ull; }
// unary: count ones up to the 0 terminator; fails if more than max_ones or the data ends
__device__ __forceinline__ bool bs_unary(BitSrc& s, uint32_t max_ones, u64 limit_bits, uint32_t* ones) {
  uint32_t q = 0u;
  for (;;) {
    if (s.avail < 32) bs_refill(s);
    const uint32_t top = (uint32_t)(s.w >> 32);
    const uint32_t run = (uint32_t)__clz((int)~top);  // leading ones among the top 32 bits
    if (run < 32u) {
      q += run;
      s.w <<= (run + 1u);
      s.avail -= (int)(run + 1u);
      s.consumed += run + 1u;
      break;
    }
    q += 32u;
    s.w <<= 32;
    s.avail -= 32;
    s.consumed += 32u;
    if (s.consumed > limit_bits || q > max_ones) return false;
  }
  *ones = q;
  return q <= max_ones && s.consumed <= limit_bits;
}
__device__ __forceinline__ bool bs_rice(BitSrc& s, uint32_t k, u64 limit_bits, uint32_t* value) {
  // read_rice_unsigned, block/decoder.cpp:74-83
  if (k > 31u) return false;
  uint32_t q;
  if (!bs_unary(s, 0xFFFFFFFFu >> k, limit_bits, &q)) return false;
  const uint32_t rem = bs_get(s, k);
  if (s.consumed > limit_bits) return false;
  *value = (q << k) | rem;
  return true;
}

// Incremental adaptive-k state (rice.hpp:15-114) kept in registers.  The 256-entry
// ring of recent values is the already decoded residual array itself.
struct KState {
  u64 sum, win_sum;
  uint32_t count;
  uint32_t lg[3], zr[3];  // 96-bit shift registers of the large / zero flags (bit 0 = newest)
  uint32_t large_cnt, zero_cnt;
  uint32_t kb;            // previous base k (hint)
};
__device__ __forceinline__ void ks_reset(KState& st) {
  st.sum = st.win_sum = 0ull;
  st.count = 0u;
  st.lg[0] = st.lg[1] = st.lg[2] = 0u;
  st.zr[0] = st.zr[1] = st.zr[2] = 0u;
  st.large_cnt = st.zero_cnt = 0u;
  st.kb = 1u;
}
// advance the model by one sample with zig-zag value u; `old_u` is the value that
// leaves the 256-sample drift window (ignored while count <= 256)
template <bool STATELESS>
__device__ __forceinline__ uint32_t ks_step(KState& st, uint32_t u, uint32_t old_u) {
  st.sum += u;
  st.count++;
  const uint32_t c = st.count;
  const u64 N = st.sum + (c >> 1);
  const uint32_t kb = kbase_from(N, c, st.kb);
  st.kb = kb ? kb : 1u;
  if (STATELESS) return kb;
  st.win_sum += u;
  if (c > kDriftWin) st.win_sum -= old_u;
  const uint32_t q = (kb >= 31u) ? 0u : (u >> kb);
  const uint32_t is_l = q > 3u, is_z = q == 0u;
  st.large_cnt += is_l - (st.lg[2] >> 31);
  st.zero_cnt += is_z - (st.zr[2] >> 31);
  st.lg[2] = (st.lg[2] << 1) | (st.lg[1] >> 31);
  st.lg[1] = (st.lg[1] << 1) | (st.lg[0] >> 31);
  st.lg[0] = (st.lg[0] << 1) | is_l;
  st.zr[2] = (st.zr[2] << 1) | (st.zr[1] >> 31);
  st.zr[1] = (st.zr[1] << 1) | (st.zr[0] >> 31);
  st.zr[0] = (st.zr[0] << 1) | is_z;
  int bias = 0;
  if (c >= kDriftWin && N >= (u64)c) {
    const u64 lm = (st.win_sum + 128ull) >> 8;
    const u64 tA = (3ull * lm + 3ull) >> 2;
    if (N < tA * c) bias = 1;
    else {
      const u64 tB = lm + 2ull + lm / 3ull;  // floor((4 lm + 3) / 3) + 1
      if (N >= tB * c) bias = -1;
    }
  }
  if (c >= kMicroWin) {
    if (st.large_cnt * 4u >= 288u) bias = bias + 1 < 1 ? bias + 1 : 1;
    else if (st.zero_cnt * 5u >= 384u) bias = bias - 1 > -1 ? bias - 1 : -1;
  }
  int k = (int)kb + bias;
  return (uint32_t)(k < 0 ? 0 : (k > 31 ? 31 : k));
}

// decode_residual_segment, block/decoder.cpp:104-306.  `res` points at the segment.
template <bool STATELESS>
__device__ __forceinline__ bool decode_segment(BitSrc& s, u64 limit, uint32_t n, uint32_t k0, uint32_t mode,
                                               int32_t* res) {
  if (mode == MODE_STATIC) {
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t u;
      if (!bs_rice(s, k0, limit, &u)) return false;
      res[i] = unzz32(u);
    }
    return true;
  }
  KState st;
  ks_reset(st);
  uint32_t k = k0;
  uint32_t idx = 0u;
  while (idx < n) {
    uint32_t u = 0u;
    uint32_t run = 0u;  // > 0: a zero run of that length was decoded instead of one value
    if (mode == MODE_RICE) {
      if (!bs_rice(s, k, limit, &u)) return false;
    } else if (mode == MODE_BIN) {
      const uint32_t tag = bs_get(s, 2u);
      if (s.consumed > limit) return false;
      if (tag == 0u) {
        u = 0u;
      } else if (tag == 3u) {
        if (!bs_rice(s, k, limit, &u)) return false;
      } else {
        const uint32_t sign = bs_get(s, 1u);
        if (s.consumed > limit) return false;
        u = (tag == 1u) ? (sign ? 1u : 2u) : (sign ? 3u : 4u);
      }
    } else {  // MODE_ZR
      const uint32_t tag = bs_get(s, 2u);
      if (s.consumed > limit) return false;
      if (tag > 2u) return false;
      if (tag == 0u) {
        if (!bs_rice(s, k, limit, &u)) return false;
      } else if (tag == 2u) {
        u = bs_get(s, 32u);
        if (s.consumed > limit) return false;
      } else {
        uint32_t enc;
        if (!bs_rice(s, kZrRunK, limit, &enc) || enc > 0xFFFFFFFFu - kZrMinRun) return false;
        run = enc + kZrMinRun;
        if (run > n - idx) return false;
      }
    }
    if (run) {
      for (uint32_t j = 0; j < run; ++j) {
        res[idx] = 0;
        const uint32_t old_u = (!STATELESS && idx >= kDriftWin) ? zz32(res[idx - kDriftWin]) : 0u;
        k = ks_step<STATELESS>(st, 0u, old_u);
        ++idx;
      }
    } else {
      res[idx] = unzz32(u);
      const uint32_t old_u = (!STATELESS && idx >= kDriftWin) ? zz32(res[idx - kDriftWin]) : 0u;
      k = ks_step<STATELESS>(st, u, old_u);
      ++idx;
    }
  }
  return true;
}

// restore_*_in_place, block/decoder.cpp:308-403: every reconstructed sample must fit int32
__device__ __forceinline__ bool restore_block(int32_t* x, uint32_t n, uint32_t type, uint32_t order, const int16_t* c) {
  if (type == PRED_FIXED) {
    if (order == 0u) return true;
    i64 h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    for (uint32_t i = 0; i < n; ++i) {
      i64 s = x[i];
      if (i >= order) {
        i64 p;
        if (order == 1u) p = h1;
        else if (order == 2u) p = 2 * h1 - h2;
        else if (order == 3u) p = 3 * h1 - 3 * h2 + h3;
        else p = 4 * h1 - 6 * h2 + 4 * h3 - h4;
        s += p;
        if (s < -2147483648ll || s > 2147483647ll) return false;
        x[i] = (int32_t)s;
      }
      h4 = h3; h3 = h2; h2 = h1; h1 = s;
    }
    return true;
  }
  if (type == PRED_FIR) {
    i64 h1 = 0, h2 = 0;
    for (uint32_t i = 0; i < n; ++i) {
      i64 s = x[i];
      if (i >= 2u) {
        s += (3 * h1 - h2) >> 2;
        if (s < -2147483648ll || s > 2147483647ll) return false;
        x[i] = (int32_t)s;
      }
      h2 = h1; h1 = s;
    }
    return true;
  }
  if (order <= 12u) {
    // history before the block start is zero, which reproduces taps = min(order, i)
    int32_t cf[13];
#pragma unroll
    for (int t = 1; t <= 12; ++t) cf[t] = (uint32_t)t <= order ? (int32_t)c[t] : 0;
    int32_t h[13];
#pragma unroll
    for (int t = 0; t <= 12; ++t) h[t] = 0;
    for (uint32_t i = 0; i < n; ++i) {
      i64 acc = 0;
#pragma unroll
      for (int t = 1; t <= 12; ++t) acc += (i64)cf[t] * (i64)h[t];
      const i64 s = (acc >> 15) + (i64)x[i];
      if (s < -2147483648ll || s > 2147483647ll) return false;
      x[i] = (int32_t)s;
#pragma unroll
      for (int t = 12; t >= 2; --t) h[t] = h[t - 1];
      h[1] = (int32_t)s;
    }
    return true;
  }
  for (uint32_t i = 0; i < n; ++i) {  // orders 13..32: legal in the format, never produced by the encoder
    i64 acc = 0;
    const uint32_t taps = order < i ? order : i;
    for (uint32_t t = 1; t <= taps; ++t) acc += (i64)c[t] * (i64)x[i - t];
    const i64 s = (acc >> 15) + (i64)x[i];
    if (s < -2147483648ll || s > 2147483647ll) return false;
    x[i] = (int32_t)s;
  }
  return true;
}

__device__ __forceinline__ uint32_t part_len(uint32_t n, uint32_t p, uint32_t idx) {
  if (p == 0u) return n;
  const uint32_t base = n >> p, cnt = 1u << p;
  return (idx + 1u == cnt) ? n - base * (cnt - 1u) : base;
}

// Block::Decoder::decode_into (block/decoder.cpp:64-520)
__device__ __forceinline__ bool decode_channel_block(BitSrc& s, u64 limit, uint32_t n, int32_t* out) {
  if (n == 0u || n > kMaxBlock) return false;
  const uint32_t type = bs_get(s, 8u);
  const uint32_t order = bs_get(s, 8u);
  if (s.consumed > limit) return false;
  if (type > 2u) return false;
  if (type == PRED_LPC) {
    if (order == 0u || order > 32u || order >= n) return false;
  } else if (type == PRED_FIR) {
    if (order != 2u) return false;
  } else if (order > 4u) {
    return false;
  }
  int16_t c[33];
  for (int i = 0; i < 33; ++i) c[i] = 0;
  if (type == PRED_LPC) {
    for (uint32_t i = 1; i <= order; ++i) {
      c[i] = (int16_t)(uint16_t)bs_get(s, 16u);
      if (s.consumed > limit) return false;
    }
  }
  const uint32_t control = bs_get(s, 8u);
  if (s.consumed > limit) return false;
  if (control & 0x10u) return false;
  const bool pflag = (control & 0x80u) != 0u;
  const uint32_t p = control & 0x0Fu, cmode = (control >> 5) & 3u;
  if (pflag && p == 0u) return false;
  if (!pflag && p != 0u) return false;
  if (p > kMaxPartOrder) return false;
  if (p > 0u && (n >> p) < kMinPart) return false;
  const uint32_t cnt = 1u << p;
  // partition metadata is read up front (block/decoder.cpp:447-455), then the segments
  uint8_t mk[256];
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint32_t m = bs_get(s, 2u);
    const uint32_t k = bs_get(s, 5u);
    if (s.consumed > limit) return false;
    mk[i] = (uint8_t)((m << 5) | k);
  }
  if ((uint32_t)(mk[0] >> 5) != cmode) return false;
  uint32_t off = 0u;
  for (uint32_t i = 0; i < cnt; ++i) {
    const uint32_t len = part_len(n, p, i);
    const uint32_t m = mk[i] >> 5, k = mk[i] & 31u;
    const bool ok = p ? decode_segment<true>(s, limit, len, k, m, out + off)
                      : decode_segment<false>(s, limit, len, k, m, out + off);
    if (!ok) return false;
    off += len;
  }
  // consume_zero_padding_to_byte (bit_reader.hpp:180-185)
  const uint32_t padn = (uint32_t)((8u - (s.consumed & 7ull)) & 7ull);
  if (padn) {
    if (bs_get(s, padn) != 0u) return false;
    if (s.consumed > limit) return false;
  }
  return restore_block(out, n, type, order, c);
}

struct DecCfg {
  uint32_t channels, stereo_mode, bit_depth, n_blocks;
};

// K12: one thread per frame-block.  blk_fs[b] = first sample of block b, blk_size[b] its
// sample count, blk_boff[b] its byte offset inside `payload`, blk_bytes[b] its byte size.
__global__ void __launch_bounds__(128) k_decode_blocks(DecCfg cfg, const uint8_t* __restrict__ payload,
                                                       const u64* __restrict__ blk_fs,
                                                       const uint32_t* __restrict__ blk_size,
                                                       const u64* __restrict__ blk_boff,
                                                       const uint32_t* __restrict__ blk_bytes, int32_t* L, int32_t* R,
                                                       uint32_t* blk_err, uint8_t* blk_ms) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cfg.n_blocks) return;
  const uint8_t* begin = payload + blk_boff[b];
  BitSrc s;
  bs_init(s, begin, begin + blk_bytes[b]);
  const u64 limit = (u64)blk_bytes[b] * 8ull;
  const uint32_t n = blk_size[b];
  uint32_t err = DERR_OK;
  uint32_t ms = 0u;
  if (cfg.channels == 2u && cfg.stereo_mode == 2u) {
    const uint32_t flag = bs_get(s, 8u);
    if (s.consumed > limit || flag > 1u) err = DERR_FLAG;
    ms = flag == 1u;
  } else if (cfg.channels == 2u && cfg.stereo_mode == 1u) {
    ms = 1u;
  }
  if (!err && !decode_channel_block(s, limit, n, L + blk_fs[b])) err = DERR_PRIMARY;
  if (!err && cfg.channels == 2u && !decode_channel_block(s, limit, n, R + blk_fs[b])) err = DERR_SECONDARY;
  if (!err && s.consumed != limit) err = DERR_TRAILING;
  blk_err[b] = err;
  blk_ms[b] = (uint8_t)ms;
}

// M/S reconstruction + depth validation + optional interleaved packing; one CTA per block.
// out_packed == nullptr: planes are fixed up in place (LAC::Decoder::decode semantics).
__global__ void __launch_bounds__(256) k_finish_pcm(DecCfg cfg, const u64* __restrict__ blk_fs,
                                                    const uint32_t* __restrict__ blk_size, int32_t* L, int32_t* R,
                                                    uint32_t* blk_err, const uint8_t* __restrict__ blk_ms,
                                                    uint8_t* out_packed) {
  const int32_t lo = cfg.bit_depth == 16u ? -32768 : -8388608, hi = cfg.bit_depth == 16u ? 32767 : 8388607;
  const uint32_t bps = cfg.bit_depth / 8u;
  for (uint32_t b = blockIdx.x; b < cfg.n_blocks; b += gridDim.x) {
    const uint32_t e = blk_err[b];
    if (e != DERR_OK && e != DERR_TRAILING) continue;  // samples are garbage
    const u64 fs = blk_fs[b];
    const uint32_t n = blk_size[b];
    const bool ms = cfg.channels == 2u && blk_ms[b];
    uint32_t bad = 0u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      i64 l = L[fs + i], r = 0;
      if (cfg.channels == 2u) {
        r = R[fs + i];
        if (ms) {
          const i64 m = l, sd = r;
          l = m + ((sd + (sd & 1)) >> 1);
          r = l - sd;
        }
      }
      const bool ok = l >= lo && l <= hi && (cfg.channels == 1u || (r >= lo && r <= hi));
      bad |= !ok;
      if (ok) {
        if (ms) {
          L[fs + i] = (int32_t)l;
          R[fs + i] = (int32_t)r;
        }
        if (out_packed) {
          uint8_t* o = out_packed + (fs + i) * (u64)(bps * cfg.channels);
          for (uint32_t k = 0; k < bps; ++k) o[k] = (uint8_t)((uint32_t)(int32_t)l >> (8u * k));
          if (cfg.channels == 2u)
            for (uint32_t k = 0; k < bps; ++k) o[bps + k] = (uint8_t)((uint32_t)(int32_t)r >> (8u * k));
        }
      }
    }
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) blk_err[b] = DERR_RANGE;
  }
}

}  // namespace lacb
