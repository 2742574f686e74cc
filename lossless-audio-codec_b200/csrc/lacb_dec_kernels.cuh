// lacb_dec_kernels.cuh -- decoder kernels (sm_100a).
//
// The .lac block bitstream is inherently serial inside a block: variable-length
// codes whose Rice parameter adapts to every decoded value, and no byte offsets for
// partitions or for the second channel (docs/format.md:18-27).  The
// only independent units are the blocks, so:
//   k_parse_blocks   one warp per frame-block.  Partitioned (stateless-k) and static segments
//                    are decoded in speculative batches: lane 0 only places up to 32 token
//                    boundaries from a shared-memory ring of the bitstream assuming k stays
//                    put, the 32 lanes extract the tokens, recompute k with one prefix scan
//                    and commit the prefix that used the right k; anything unusual goes to the
//                    exact serial reader.  Unpartitioned adaptive segments (stateful model)
//                    use the serial reader throughout.
//   k_restore_blocks one thread per channel-block: fixed / FIR / LPC reconstruction,
//                    chunks staged through shared memory with cp.async.
//   k_finish_pcm     mid/side reconstruction, PCM range validation, interleave +
//                    16/24-bit packing; fully data parallel and bandwidth bound.
// Reference: Block::Decoder::decode_into src/codec/block/decoder.cpp:64-520,
// LAC::Decoder decode_block src/codec/lac/decoder.cpp:167-207,
// reconstruct_mid_side_in_place :48-65, pack_pcm_to_wav_bytes src/main.cpp:150-182.
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

// MSB-first bit reader (BitReader, bitstream/bit_reader.hpp:40-202), register resident:
// a 64-bit window `w` (next bit at bit 63, `avail` valid bits) plus the following
// big-endian word already loaded (`nextw`), so a refill is a shift/or and the cached load
// it triggers is for data needed 32+ bits later (off the dependent chain).  After
// rd_refill() at least 33 bits are valid, enough for a unary prefix of <= 31 ones with its
// terminator, or for any fixed field (<= 32 bits).  `base` is the 4-byte aligned address at
// or before the first byte; positions are bit offsets from it.  Reads past `end` return
// whatever follows (word index clamped to the caller's buffer) and are caught by the
// rd_over() checks at segment granularity, which is how the reference's "ran out of data"
// rejections are reproduced.
struct BitRd {
  const uint32_t* base;
  u64 w;
  uint32_t avail;   // valid bits in w
  uint32_t wi;      // index of the word held in nextw
  uint32_t nextw;   // raw (little-endian) prefetched word
  uint32_t last_word;
  u64 start, end;
};
__device__ __forceinline__ uint32_t rd_word_raw(const BitRd& r, uint32_t i) {
  i = i > r.last_word ? r.last_word : i;
  return __ldg(r.base + i);
}
__device__ __forceinline__ uint32_t rd_word(const BitRd& r, uint32_t i) { return __byte_perm(rd_word_raw(r, i), 0u, 0x0123); }
__device__ __forceinline__ u64 rd_pos(const BitRd& r) { return (u64)r.wi * 32ull - r.avail; }
__device__ __forceinline__ void rd_refill(BitRd& r) {
  if (r.avail <= 32u) {
    // nextw holds the little-endian word as loaded: swapping it here, one refill after the
    // load was issued, keeps the load latency off the dependent chain
    r.w |= (u64)__byte_perm(r.nextw, 0u, 0x0123) << (32u - r.avail);
    r.avail += 32u;
    r.wi += 1u;
    r.nextw = rd_word_raw(r, r.wi);
  }
}
__device__ __forceinline__ void rd_seek(BitRd& r, u64 pos) {
  const u64 wq = pos >> 5;
  const uint32_t w0 = wq > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)wq;
  const uint32_t sh = (uint32_t)pos & 31u;
  r.w = (u64)rd_word(r, w0) << (32u + sh);
  r.avail = 32u - sh;
  r.wi = w0 + 1u;
  r.nextw = rd_word_raw(r, r.wi);
  rd_refill(r);
}
__device__ __forceinline__ void rd_init(BitRd& r, const uint8_t* begin, u64 nbytes, const uint8_t* buf_end) {
  const uint64_t a = reinterpret_cast<uint64_t>(begin);
  r.base = reinterpret_cast<const uint32_t*>(a & ~(uint64_t)3);
  r.start = (a & 3ull) * 8ull;
  r.end = r.start + nbytes * 8ull;
  // last word that lies entirely inside the caller's buffer
  const uint64_t words = (reinterpret_cast<uint64_t>(buf_end) - (a & ~(uint64_t)3)) >> 2;
  r.last_word = words ? (uint32_t)(words - 1ull > 0xFFFFFFF0ull ? 0xFFFFFFF0ull : words - 1ull) : 0u;
  rd_seek(r, r.start);
}
__device__ __forceinline__ bool rd_over(const BitRd& r) { return rd_pos(r) > r.end; }
// n in [0,32]
__device__ __forceinline__ uint32_t rd_get(BitRd& r, uint32_t n) {
  rd_refill(r);
  const uint32_t v = n ? (uint32_t)(r.w >> (64u - n)) : 0u;
  r.w <<= n;
  r.avail -= n;
  return v;
}

// read_rice_unsigned (block/decoder.cpp:74-83): unary quotient limited to
// UINT32_MAX >> k ones (read_unary_ones, bit_reader.hpp:140-172), then k remainder bits.
// Does not test for the end of data on the fast path: callers check rd_over() per segment.
__device__ __forceinline__ bool rd_rice(BitRd& r, uint32_t k, uint32_t* value) {
  rd_refill(r);
  const uint32_t hi = (uint32_t)(r.w >> 32);
  uint32_t q;
  if (hi != 0xFFFFFFFFu) {  // terminator within the 32 bits in view
    q = (uint32_t)__clz((int)~hi);
    r.w <<= (q + 1u);
    r.avail -= q + 1u;
  } else {  // long unary run: walk word by word
    const uint32_t max_ones = 0xFFFFFFFFu >> k;
    q = 0u;
    for (;;) {
      rd_refill(r);
      const uint32_t top = (uint32_t)(r.w >> 32);
      const uint32_t run = (uint32_t)__clz((int)~top);
      if (run < 32u) {
        q += run;
        r.w <<= (run + 1u);
        r.avail -= run + 1u;
        break;
      }
      q += 32u;
      r.w <<= 32;
      r.avail -= 32u;
      if (q > max_ones || rd_over(r)) return false;
    }
  }
  rd_refill(r);
  const uint32_t rem = k ? (uint32_t)(r.w >> (64u - k)) : 0u;
  r.w <<= k;
  r.avail -= k;
  *value = (q << k) | rem;
  return q <= (0xFFFFFFFFu >> k);
}

// Incremental adaptive-k state (rice.hpp:15-114).  ring[] is the 256-entry window of
// recent zig-zag values (shared memory, one per warp).
struct KState {
  u64 sum, win_sum;
  uint32_t count;
  uint32_t lg[3], zr[3];  // 96-bit shift registers of the large / zero flags (bit 0 = newest)
  uint32_t large_cnt, zero_cnt;
  uint32_t kb;            // previous base k (search hint)
};
__device__ __forceinline__ void ks_reset(KState& st) {
  st.sum = st.win_sum = 0ull;
  st.count = 0u;
  st.lg[0] = st.lg[1] = st.lg[2] = 0u;
  st.zr[0] = st.zr[1] = st.zr[2] = 0u;
  st.large_cnt = st.zero_cnt = 0u;
  st.kb = 1u;
}
template <bool STATELESS>
__device__ __forceinline__ uint32_t ks_step(KState& st, uint32_t u, uint32_t* ring) {
  st.sum += u;
  const uint32_t c = ++st.count;
  const u64 N = st.sum + (c >> 1);
  const uint32_t kb = kbase_clz(N, c);
  if (STATELESS) return kb;
  const uint32_t slot = (c - 1u) & (kDriftWin - 1u);
  st.win_sum += u;
  if (c > kDriftWin) st.win_sum -= ring[slot];
  ring[slot] = u;
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
  const int k = (int)kb + bias;
  return (uint32_t)(k < 0 ? 0 : (k > 31 ? 31 : k));
}

// decode_residual_segment, block/decoder.cpp:104-306.  `res` points at the segment.
template <bool STATELESS>
__device__ __forceinline__ bool decode_segment(BitRd& r, uint32_t n, uint32_t k0, uint32_t mode, int32_t* res,
                                               uint32_t* ring) {
  if (mode == MODE_STATIC) {
    if (k0 > 31u) return false;
    for (uint32_t i = 0; i < n; ++i) {
      uint32_t u;
      if (!rd_rice(r, k0, &u)) return false;
      res[i] = unzz32(u);
      if ((i & 255u) == 255u && rd_over(r)) return false;  // bounds the work on truncated streams
    }
    return !rd_over(r);
  }
  KState st;
  ks_reset(st);
  uint32_t k = k0;
  uint32_t idx = 0u;
  if (mode == MODE_RICE) {
    while (idx < n) {
      uint32_t u;
      if (!rd_rice(r, k, &u)) return false;
      res[idx++] = unzz32(u);
      k = ks_step<STATELESS>(st, u, ring);
      if ((idx & 255u) == 0u && rd_over(r)) return false;
    }
    return !rd_over(r);
  }
  if (mode == MODE_BIN) {
    while (idx < n) {
      const uint32_t tag = rd_get(r, 2u);
      if (rd_over(r)) return false;
      uint32_t u = 0u;
      if (tag == 3u) {
        if (!rd_rice(r, k, &u)) return false;
      } else if (tag != 0u) {
        const uint32_t sign = rd_get(r, 1u);
        if (rd_over(r)) return false;
        u = (tag == 1u) ? (sign ? 1u : 2u) : (sign ? 3u : 4u);
      }
      res[idx++] = unzz32(u);
      k = ks_step<STATELESS>(st, u, ring);
    }
    return !rd_over(r);
  }
  // MODE_ZR
  while (idx < n) {
    const uint32_t tag = rd_get(r, 2u);
    if (rd_over(r)) return false;
    if (tag > 2u) return false;
    if (tag == 1u) {
      uint32_t enc;
      if (!rd_rice(r, kZrRunK, &enc) || enc > 0xFFFFFFFFu - kZrMinRun) return false;
      const uint32_t run = enc + kZrMinRun;
      if (run > n - idx) return false;
      if (STATELESS) {  // count += run; k = adapt_k_stateless(sum, count)  (block/decoder.cpp:208-211)
        for (uint32_t j = 0; j < run; ++j) res[idx++] = 0;
        st.count += run;
        k = kbase_clz(st.sum + (st.count >> 1), st.count);
      } else {
        for (uint32_t j = 0; j < run; ++j) {
          res[idx++] = 0;
          k = ks_step<STATELESS>(st, 0u, ring);
        }
      }
      continue;
    }
    uint32_t u;
    if (tag == 0u) {
      if (!rd_rice(r, k, &u)) return false;
    } else {
      u = rd_get(r, 32u);
      if (rd_over(r)) return false;
    }
    res[idx++] = unzz32(u);
    k = ks_step<STATELESS>(st, u, ring);
  }
  return !rd_over(r);
}

__device__ __forceinline__ uint32_t part_len(uint32_t n, uint32_t p, uint32_t idx) {
  if (p == 0u) return n;
  const uint32_t base = n >> p, cnt = 1u << p;
  return (idx + 1u == cnt) ? n - base * (cnt - 1u) : base;
}

// What the restore stage needs to know about a parsed channel-block.
struct ChanHdr {
  uint8_t type, order;
  int16_t coef[33];
};

// Per-warp scratch of the speculative token batches.
struct ParseScratch {
  // big-endian words of the bitstream as overlapped pairs: ring[i & 127] = (word i) << 32 | word i+1,
  // so the 32 bits at any bit position come from one 64-bit load and one funnel shift
  u64 ring[128];
  uint32_t tp[34];  // ring bit position of every token of the batch; [cnt] = end
};
// staged word range [lo, hi) of the ring (warp-uniform registers)
struct Stage {
  uint32_t lo, hi;
  // Words of the chunk [pend_at, pend_at + 32) as loaded (little-endian), requested at the previous
  // refill: a refill consumes loads that were issued ~64 tokens earlier instead of waiting for its own.
  // pend1 is lane 31's right-hand neighbour (first word behind the chunk).
  uint32_t pend_at, pend, pend1;
};

// Makes sure the 64 words from the one holding bit `pos` are in the ring: a batch of 32 tokens
// spans at most 32 * 59 + 31 bits and every token is looked at through a 2-word window.
__device__ __forceinline__ void stage_ensure(const BitRd& r, ParseScratch* sc, Stage& sg, u64 pos, uint32_t lane) {
  const u64 wq = pos >> 5;
  const uint32_t wi = wq > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)wq;
  if (wi < sg.lo || wi > sg.hi) sg.lo = sg.hi = wi;  // outside the staged range: start over
  bool any = false;
  while (sg.hi < wi + 64u) {  // the slots overwritten hold words below wi - 32
    const uint32_t i = sg.hi + lane;
    uint32_t raw0, raw1 = 0u;
    if (sg.pend_at == sg.hi) {
      raw0 = sg.pend;
      raw1 = sg.pend1;
    } else {
      raw0 = rd_word_raw(r, i);
      if (lane == 31u) raw1 = rd_word_raw(r, i + 1u);
    }
    const uint32_t w0 = __byte_perm(raw0, 0u, 0x0123);
    uint32_t w1 = __shfl_down_sync(kFull, w0, 1);
    if (lane == 31u) w1 = __byte_perm(raw1, 0u, 0x0123);
    sc->ring[i & 127u] = ((u64)w0 << 32) | w1;
    sg.hi += 32u;
    any = true;
  }
  if (sg.hi - sg.lo > 128u) sg.lo = sg.hi - 128u;
  if (any) {
    sg.pend_at = sg.hi;  // request the next chunk; nothing reads these registers before the next refill
    sg.pend = rd_word_raw(r, sg.hi + lane);
    sg.pend1 = lane == 31u ? rd_word_raw(r, sg.hi + 32u) : 0u;
    __syncwarp();
  }
}
// Ring bit positions: rp = ((word index & 127) << 5) + bit, allowed to run past 4096 (the slot
// index wraps in the address).  32 bits starting at ring position rp:
__device__ __forceinline__ uint32_t ring_peek(const u64* ring, uint32_t rp) {
  const u64 v = ring[(rp >> 5) & 127u];
  return __funnelshift_l((uint32_t)v, (uint32_t)(v >> 32), rp);
}

// position of the most significant set bit, 0xFFFFFFFF for 0 (PTX bfind.u32, SASS FLO.U32)
__device__ __forceinline__ uint32_t bfind_u32(uint32_t x) {
#ifdef LACB_EMU
  return 31u - (uint32_t)__clz((int)x);
#else
  uint32_t f;
  asm("bfind.u32 %0, %1;" : "=r"(f) : "r"(x));
  return f;
#endif
}

// One token by the exact serial reader (lane 0 only): value u, sample count w.
//   MODE_RICE / MODE_STATIC: Rice(k)             block/decoder.cpp:126-136, 296-303
//   MODE_ZR  : tag 00 Rice(k) | 01 run | 10 raw  block/decoder.cpp:138-257
//   MODE_BIN : tag 00 | 01 s | 10 s | 11 Rice(k) block/decoder.cpp:259-294
__device__ __forceinline__ bool parse_token(BitRd& r, uint32_t mode, uint32_t k, uint32_t* u, uint32_t* w) {
  *w = 1u;
  if (mode == MODE_RICE || mode == MODE_STATIC) return rd_rice(r, k, u);
  const uint32_t tag = rd_get(r, 2u);
  if (mode == MODE_BIN) {
    if (tag == 0u) { *u = 0u; return true; }
    if (tag == 3u) return rd_rice(r, k, u);
    const uint32_t sign = rd_get(r, 1u);
    *u = (tag == 1u) ? (sign ? 1u : 2u) : (sign ? 3u : 4u);
    return true;
  }
  if (tag > 2u) return false;
  if (tag == 0u) return rd_rice(r, k, u);
  if (tag == 2u) { *u = rd_get(r, 32u); return true; }
  uint32_t enc;
  if (!rd_rice(r, kZrRunK, &enc) || enc > 0xFFFFFFFFu - kZrMinRun) return false;
  *u = 0u;
  *w = enc + kZrMinRun;
  return true;
}

// Lane 0: boundaries of up to B tokens under parameter k, nothing else -- per token one
// 2-word window from the ring, a count-leading-ones and an add sit on the serial chain.
// Stops in front of a token the exact reader has to look at (unary run not terminated inside
// the view, reserved tag).  tp[0..cnt] = token starts / end, returns cnt.
__device__ __forceinline__ uint32_t walk_tokens(const u64* ring, uint32_t* tp, uint32_t rel, uint32_t mode,
                                                uint32_t k, uint32_t B) {
  uint32_t cnt = 0u;
  if (mode == MODE_RICE || mode == MODE_STATIC) {
    // No exit test on the chain: a unary run that does not end inside the 32-bit view counts as 32
    // ones (clz(0) = 32) and the walk simply goes on; the lane that extracts the token sees q >= 32,
    // flags it, and everything from there on is discarded and redone by the exact reader.
    // clz(x) = 31 - bfind(x) with bfind(0) = 0xFFFFFFFF: written out so that the constant part joins
    // k + 1 and one three-input add is left on the chain (ptxas keeps 31 - x as a separate add otherwise)
    const uint32_t k32 = k + 32u;
    for (; cnt + 4u <= B; cnt += 4u) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tp[cnt + i] = rel;
        rel = rel + k32 - bfind_u32(~ring_peek(ring, rel));
      }
    }
    for (; cnt < B; ++cnt) {
      tp[cnt] = rel;
      rel = rel + k32 - bfind_u32(~ring_peek(ring, rel));
    }
  } else if (mode == MODE_ZR) {
    for (; cnt < B; ++cnt) {
      tp[cnt] = rel;
      const uint32_t hi = ring_peek(ring, rel);
      const uint32_t tag = hi >> 30;
      if (tag == 3u) break;
      if (tag == 2u) {
        rel += 34u;
      } else {
        const uint32_t q = (uint32_t)__clz((int)~((hi << 2) | 3u));
        if (q >= 30u) break;
        rel += 3u + q + (tag ? kZrRunK : k);
      }
    }
  } else {  // MODE_BIN
    for (; cnt < B; ++cnt) {
      tp[cnt] = rel;
      const uint32_t hi = ring_peek(ring, rel);
      const uint32_t tag = hi >> 30;
      if (tag == 0u) {
        rel += 2u;
      } else if (tag != 3u) {
        rel += 3u;
      } else {
        const uint32_t q = (uint32_t)__clz((int)~((hi << 2) | 3u));
        if (q >= 30u) break;
        rel += 3u + q + k;
      }
    }
  }
  tp[cnt] = rel;
  return cnt;
}

// Warp-cooperative decode of one segment whose Rice parameter is either fixed (MODE_STATIC)
// or follows the STATELESS model (block/encoder.cpp:72-77: a function of the sum of u and
// the sample count only).
//
// Lane 0 walks the token boundaries of a batch assuming "k stays k"; lane t then extracts
// token t from its two boundaries, and one prefix scan over the batch gives every lane the k
// that follows its token.  The first lane whose k differs ends the valid prefix: its tokens
// are committed with coalesced stores and the batch restarts behind it with the new k.  A
// token the walker cannot place (long unary run, reserved tag, data or run running past the
// end) is handed to the exact serial reader once everything in front of it is verified, so
// every accept / reject verdict is the serial decoder's.  `pos` (absolute bit position, same
// in all lanes) is advanced past the segment.
__device__ __forceinline__ bool decode_segment_fast(BitRd& r, u64& pos, uint32_t n, uint32_t k0, uint32_t mode,
                                                    int32_t* res, ParseScratch* sc, Stage& stg, uint32_t lane) {
  const bool adaptive = mode != MODE_STATIC;
  u64 sum = 0ull;
  uint32_t count = 0u, idx = 0u, k = k0, B = adaptive ? 8u : 32u;
  if (k0 > 31u) return false;
  while (idx < n) {
    const uint32_t left = n - idx;
    const uint32_t want = k > 26u ? 0u : (B < left ? B : left);  // k > 26: the unary limit can bind, serial reader
    stage_ensure(r, sc, stg, pos, lane);
    // ring position of `pos`, and the absolute bit position of ring position 0
    const uint32_t rel0 = (uint32_t)pos & 4095u;
    const u64 ring_base = pos - rel0;
    uint32_t cnt = 0u;
    if (lane == 0u && want) cnt = walk_tokens(sc->ring, sc->tp, rel0, mode, k, want);
    cnt = __shfl_sync(kFull, cnt, 0);
    __syncwarp();
    uint32_t u = 0u, w = 0u;
    bool bad = false;
    if (lane < cnt) {
      const uint32_t s = sc->tp[lane], e = sc->tp[lane + 1u];
      bad = ring_base + e > r.end;
      w = 1u;
      if (mode == MODE_RICE || mode == MODE_STATIC) {
        const uint32_t rem = k ? ring_peek(sc->ring, e - k) >> (32u - k) : 0u;
        const uint32_t q = e - s - 1u - k;
        bad = bad || q >= 32u;  // the walker could not see the end of the unary run
        u = (q << k) | rem;
      } else {
        const uint32_t hs = ring_peek(sc->ring, s);
        const uint32_t tag = hs >> 30;
        if (mode == MODE_ZR && tag == 2u) {
          u = ring_peek(sc->ring, s + 2u);
        } else if (mode == MODE_ZR && tag == 1u) {
          const uint32_t rem = ring_peek(sc->ring, e - kZrRunK) >> (32u - kZrRunK);
          w = (((e - s - 3u - kZrRunK) << kZrRunK) | rem) + kZrMinRun;
        } else if (mode == MODE_BIN && tag != 3u) {
          const uint32_t sign = (hs >> 29) & 1u;
          u = tag == 0u ? 0u : (tag == 1u ? (sign ? 1u : 2u) : (sign ? 3u : 4u));
        } else {  // tag 00 (zero-run mode) / 11 (bin mode): Rice(k) behind the tag
          const uint32_t rem = k ? ring_peek(sc->ring, e - k) >> (32u - k) : 0u;
          u = ((e - s - 3u - k) << k) | rem;
        }
      }
    }
    uint32_t PW = w;
    u64 PU = u;
    uint32_t knext = k;
    if (mode == MODE_RICE && k <= 21u) {
      // plain adaptive Rice: one sample per token, and every token of the batch is below
      // 32 << k <= 2^26 (larger ones are flagged bad), so the prefix of u fits 32 bits
      uint32_t p32 = bad ? 0u : u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(kFull, p32, d);
        if (lane >= (uint32_t)d) p32 += y;
      }
      PU = p32;
      PW = lane + 1u;
      const uint32_t c = count + PW;
      knext = kbase_clz(sum + PU + (c >> 1), c);
    } else if (adaptive) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u64 yu = __shfl_up_sync(kFull, PU, d);
        const uint32_t yw = __shfl_up_sync(kFull, PW, d);
        if (lane >= (uint32_t)d) {
          PU += yu;
          PW += yw;
        }
      }
      const uint32_t c = count + PW;
      knext = c ? kbase_clz(sum + PU + (c >> 1), c) : 0u;
    } else {
      PW = lane + 1u;  // one sample per token, the scan is not needed
    }
    bad = bad || (lane < cnt && PW > left);
    const uint32_t mbad = __ballot_sync(kFull, bad);
    const uint32_t mism = __ballot_sync(kFull, lane < cnt && knext != k);
    uint32_t valid = mism ? (uint32_t)__ffs((int)mism) : cnt;  // tokens [0, valid) were read with the right k
    const uint32_t first_bad = mbad ? (uint32_t)__ffs((int)mbad) - 1u : cnt;
    if (valid > first_bad) valid = first_bad;
    uint32_t knew = k;
    if (valid) {
      if (lane < valid && w == 1u) res[idx + PW - 1u] = unzz32(u);
      uint32_t runs = __ballot_sync(kFull, lane < valid && w > 1u);
      while (runs) {  // zero runs are filled by the whole warp
        const int t = __ffs((int)runs) - 1;
        runs &= runs - 1u;
        const uint32_t o = __shfl_sync(kFull, PW - w, t), len = __shfl_sync(kFull, w, t);
        for (uint32_t j = lane; j < len; j += 32u) res[idx + o + j] = 0;
      }
      knew = __shfl_sync(kFull, knext, (int)valid - 1);
      sum += __shfl_sync(kFull, PU, (int)valid - 1);
      const uint32_t adv = __shfl_sync(kFull, PW, (int)valid - 1);
      count += adv;
      idx += adv;
      pos = ring_base + sc->tp[valid];
    }
    if (knew != k) {  // k moved behind token valid-1: what follows was read with the wrong k
      k = knew;
      B = valid < 4u ? 4u : valid;
      __syncwarp();
      continue;
    }
    if (valid == cnt && cnt == want && want) {  // a clean full batch
      B = B * 2u > 32u ? 32u : B * 2u;
      __syncwarp();
      continue;
    }
    if (idx >= n) break;
    // The next token is one the walker stopped at, or it runs past the data / the segment:
    // everything before it is verified, so the serial reader decides with the true k.
    uint32_t ok = 0u, su = 0u, sw = 0u;
    u64 npos = 0ull;
    if (lane == 0u) {
      rd_seek(r, pos);
      ok = parse_token(r, mode, k, &su, &sw) && !rd_over(r) && sw <= n - idx;
      npos = rd_pos(r);
    }
    ok = __shfl_sync(kFull, ok, 0);
    if (!ok) return false;
    su = __shfl_sync(kFull, su, 0);
    sw = __shfl_sync(kFull, sw, 0);
    pos = __shfl_sync(kFull, npos, 0);
    if (sw == 1u) {
      if (lane == 0u) res[idx] = unzz32(su);
    } else {
      for (uint32_t j = lane; j < sw; j += 32u) res[idx + j] = 0;
    }
    sum += su;
    count += sw;
    idx += sw;
    if (adaptive) k = kbase_clz(sum + (count >> 1), count);
    __syncwarp();
  }
  return true;
}

// Warp-cooperative decode of an UNPARTITIONED adaptive segment (stateful model: Rice::adapt_k with the drift and
// micro windows, rice.hpp:45-114) in Rice or bin mode, by the same speculation as decode_segment_fast: lane 0 places
// the boundaries of a batch assuming "k stays k", lane t extracts token t, and the k that follows every token is
// recomputed exactly from prefix sums --
//   base k     from sum + prefix of u and the sample count (closed form, kbase_clz),
//   drift rule from the 256-sample window sum: the running window sum plus the prefix of (u entering - u leaving),
//              the leaving value read from the 256-entry ring of committed samples,
//   micro rule from the counts of "large" / "zero" quotient flags over the last 96 samples: the running counts plus
//              popcounts of the batch's flag ballots minus those of the flags leaving (the top word of the 96-bit
//              flag history, bit-reversed so that bit t is the flag token t pushes out).
// The first lane whose k differs from the assumed one ends the valid prefix; the committed tokens update the ring,
// the flag history and the running sums, and the batch restarts behind them.  Tokens the walker cannot place go to
// the exact serial reader (parse_token) and are folded into the state one sample at a time.
// Zero-run mode is not handled here (a run token advances the model by many samples): the caller keeps the serial
// reader for it.
struct StatefulState {
  u64 sum, win_sum;
  uint32_t count;
  uint32_t lg[3], zr[3];  // flag history, bit 0 = newest (as KState)
  uint32_t Lc, Zc;        // set bits in lg / zr
};
__device__ __forceinline__ int stateful_bias(u64 N, uint32_t c, u64 win_sum, uint32_t Lc, uint32_t Zc) {
  int bias = 0;
  if (c >= kDriftWin && N >= (u64)c) {
    const u64 lm = (win_sum + 128ull) >> 8;
    const u64 tA = (3ull * lm + 3ull) >> 2;
    if (N < tA * c) bias = 1;
    else {
      const u64 tB = lm + 2ull + lm / 3ull;
      if (N >= tB * c) bias = -1;
    }
  }
  if (c >= kMicroWin) {
    if (Lc * 4u >= 288u) bias = bias + 1 < 1 ? bias + 1 : 1;
    else if (Zc * 5u >= 384u) bias = bias - 1 > -1 ? bias - 1 : -1;
  }
  return bias;
}
// one sample folded into the state by every lane alike (lane 0 writes the ring); returns the next k
__device__ __forceinline__ uint32_t stateful_step(StatefulState& st, uint32_t u, uint32_t* ring, uint32_t lane) {
  st.sum += u;
  const uint32_t c = ++st.count;
  const u64 N = st.sum + (c >> 1);
  const uint32_t kb = kbase_clz(N, c);
  const uint32_t slot = (c - 1u) & (kDriftWin - 1u);
  st.win_sum += u;
  if (c > kDriftWin) st.win_sum -= ring[slot];
  __syncwarp();
  if (lane == 0u) ring[slot] = u;
  __syncwarp();
  const uint32_t q = (kb >= 31u) ? 0u : (u >> kb);
  const uint32_t is_l = q > 3u, is_z = q == 0u;
  st.Lc += is_l - (st.lg[2] >> 31);
  st.Zc += is_z - (st.zr[2] >> 31);
  st.lg[2] = (st.lg[2] << 1) | (st.lg[1] >> 31);
  st.lg[1] = (st.lg[1] << 1) | (st.lg[0] >> 31);
  st.lg[0] = (st.lg[0] << 1) | is_l;
  st.zr[2] = (st.zr[2] << 1) | (st.zr[1] >> 31);
  st.zr[1] = (st.zr[1] << 1) | (st.zr[0] >> 31);
  st.zr[0] = (st.zr[0] << 1) | is_z;
  const int k = (int)kb + stateful_bias(N, c, st.win_sum, st.Lc, st.Zc);
  return (uint32_t)(k < 0 ? 0 : (k > 31 ? 31 : k));
}
// Not inlined (and everything passed by value) so that the register allocation of the partitioned path, the one
// every benchmark stream takes, is not touched by this one.  Returns the bit position behind the segment, or
// ~0 when the segment is rejected.
#ifdef LACB_EMU
#define LACB_NOINLINE
#else
#define LACB_NOINLINE __noinline__
#endif
__device__ LACB_NOINLINE u64 decode_segment_stateful(BitRd r, u64 pos, uint32_t n, uint32_t k0, uint32_t mode,
                                                     int32_t* res, uint32_t* ring, ParseScratch* sc, uint32_t lane) {
  constexpr u64 kReject = ~0ull;
  if (k0 > 31u) return kReject;
  Stage stg = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFEu, 0u, 0u};
  StatefulState st;
  st.sum = st.win_sum = 0ull;
  st.count = 0u;
  st.lg[0] = st.lg[1] = st.lg[2] = st.zr[0] = st.zr[1] = st.zr[2] = 0u;
  st.Lc = st.Zc = 0u;
  uint32_t idx = 0u, k = k0, B = 8u;
  const uint32_t le = 0xFFFFFFFFu >> (31u - lane);  // lanes 0..lane
  while (idx < n) {
    const uint32_t left = n - idx;
    const uint32_t want = k > 26u ? 0u : (B < left ? B : left);
    stage_ensure(r, sc, stg, pos, lane);
    const uint32_t rel0 = (uint32_t)pos & 4095u;
    const u64 ring_base = pos - rel0;
    uint32_t cnt = 0u;
    if (lane == 0u && want) cnt = walk_tokens(sc->ring, sc->tp, rel0, mode, k, want);
    cnt = __shfl_sync(kFull, cnt, 0);
    __syncwarp();
    uint32_t u = 0u;
    bool bad = false;
    if (lane < cnt) {
      const uint32_t s = sc->tp[lane], e = sc->tp[lane + 1u];
      bad = ring_base + e > r.end;
      if (mode == MODE_RICE) {
        const uint32_t rem = k ? ring_peek(sc->ring, e - k) >> (32u - k) : 0u;
        const uint32_t q = e - s - 1u - k;
        bad = bad || q >= 32u;
        u = (q << k) | rem;
      } else {  // MODE_BIN
        const uint32_t hs = ring_peek(sc->ring, s);
        const uint32_t tag = hs >> 30;
        if (tag != 3u) {
          const uint32_t sign = (hs >> 29) & 1u;
          u = tag == 0u ? 0u : (tag == 1u ? (sign ? 1u : 2u) : (sign ? 3u : 4u));
        } else {
          const uint32_t rem = k ? ring_peek(sc->ring, e - k) >> (32u - k) : 0u;
          u = ((e - s - 3u - k) << k) | rem;
        }
      }
    }
    const bool live = lane < cnt && !bad;
    // prefix of u and of (u - the sample leaving the drift window)
    const uint32_t c = st.count + lane + 1u;
    const uint32_t slot = (c - 1u) & (kDriftWin - 1u);
    const uint32_t leave = (live && c > kDriftWin) ? ring[slot] : 0u;
    u64 PU = live ? u : 0u;
    i64 PD = live ? (i64)u - (i64)leave : 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 yu = __shfl_up_sync(kFull, PU, d);
      const i64 yd = __shfl_up_sync(kFull, PD, d);
      if (lane >= (uint32_t)d) {
        PU += yu;
        PD += yd;
      }
    }
    const u64 N = st.sum + PU + (c >> 1);
    const uint32_t kb = kbase_clz(N, c);
    const uint32_t q = (kb >= 31u) ? 0u : (u >> kb);
    const uint32_t BL = __ballot_sync(kFull, live && q > 3u), BZ = __ballot_sync(kFull, live && q == 0u);
    const uint32_t outL = __brev(st.lg[2]), outZ = __brev(st.zr[2]);  // bit t: the flag token t pushes out
    const uint32_t Lc = st.Lc + (uint32_t)__popc(BL & le) - (uint32_t)__popc(outL & le);
    const uint32_t Zc = st.Zc + (uint32_t)__popc(BZ & le) - (uint32_t)__popc(outZ & le);
    const u64 ws = (u64)((i64)st.win_sum + PD);
    const int kk = (int)kb + stateful_bias(N, c, ws, Lc, Zc);
    const uint32_t knext = (uint32_t)(kk < 0 ? 0 : (kk > 31 ? 31 : kk));
    const uint32_t mbad = __ballot_sync(kFull, bad);
    const uint32_t mism = __ballot_sync(kFull, lane < cnt && knext != k);
    uint32_t valid = mism ? (uint32_t)__ffs((int)mism) : cnt;
    const uint32_t first_bad = mbad ? (uint32_t)__ffs((int)mbad) - 1u : cnt;
    if (valid > first_bad) valid = first_bad;
    uint32_t knew = k;
    if (valid) {
      if (lane < valid) {
        res[idx + lane] = unzz32(u);
        ring[slot] = u;
      }
      const int last = (int)valid - 1;
      knew = __shfl_sync(kFull, knext, last);
      st.sum += __shfl_sync(kFull, PU, last);
      st.win_sum = __shfl_sync(kFull, ws, last);
      st.Lc = __shfl_sync(kFull, Lc, last);
      st.Zc = __shfl_sync(kFull, Zc, last);
      st.count += valid;
      idx += valid;
      pos = ring_base + sc->tp[valid];
      // flag history: `valid` new flags, newest (token valid - 1) at bit 0
      const uint32_t insL = __brev(BL << (32u - valid)), insZ = __brev(BZ << (32u - valid));
      st.lg[2] = __funnelshift_lc(st.lg[1], st.lg[2], valid);
      st.lg[1] = __funnelshift_lc(st.lg[0], st.lg[1], valid);
      st.lg[0] = __funnelshift_lc(0u, st.lg[0], valid) | insL;
      st.zr[2] = __funnelshift_lc(st.zr[1], st.zr[2], valid);
      st.zr[1] = __funnelshift_lc(st.zr[0], st.zr[1], valid);
      st.zr[0] = __funnelshift_lc(0u, st.zr[0], valid) | insZ;
    }
    __syncwarp();
    if (knew != k) {
      k = knew;
      B = valid < 4u ? 4u : valid;
      continue;
    }
    if (valid == cnt && cnt == want && want) {
      B = B * 2u > 32u ? 32u : B * 2u;
      continue;
    }
    if (idx >= n) break;
    // a token the walker could not place (or one that runs past the data): the exact reader, one sample
    uint32_t ok = 0u, su = 0u, sw = 0u;
    u64 npos = 0ull;
    if (lane == 0u) {
      rd_seek(r, pos);
      ok = parse_token(r, mode, k, &su, &sw) && !rd_over(r);
      npos = rd_pos(r);
    }
    ok = __shfl_sync(kFull, ok, 0);
    if (!ok) return kReject;
    su = __shfl_sync(kFull, su, 0);
    pos = __shfl_sync(kFull, npos, 0);
    if (lane == 0u) res[idx] = unzz32(su);
    idx += 1u;
    k = stateful_step(st, su, ring, lane);
  }
  return pos;
}

// Header + residual part of Block::Decoder::decode_into (block/decoder.cpp:64-512): leaves
// the residual in `out` and the predictor description in `hdr`.  Warp collective: every
// lane calls it, lane 0 owns the reader.
__device__ __forceinline__ bool parse_channel_block(BitRd& r, uint32_t n, int32_t* out, ChanHdr* hdr, uint32_t* ring,
                                                    ParseScratch* sc, uint32_t lane) {
  if (n == 0u || n > kMaxBlock) return false;
  uint32_t ok = 1u, p = 0u;
  u64 table_pos = 0ull;
  if (lane == 0u) {
    ok = 0u;
    do {
      const uint32_t type = rd_get(r, 8u);
      const uint32_t order = rd_get(r, 8u);
      if (rd_over(r) || type > 2u) break;
      if (type == PRED_LPC) {
        if (order == 0u || order > 32u || order >= n) break;
      } else if (type == PRED_FIR) {
        if (order != 2u) break;
      } else if (order > 4u) {
        break;
      }
      hdr->type = (uint8_t)type;
      hdr->order = (uint8_t)order;
      bool bad = false;
      if (type == PRED_LPC) {
        for (uint32_t i = 1; i <= order && !bad; ++i) {
          hdr->coef[i] = (int16_t)(uint16_t)rd_get(r, 16u);
          bad = rd_over(r);
        }
      }
      if (bad) break;
      const uint32_t control = rd_get(r, 8u);
      if (rd_over(r) || (control & 0x10u)) break;
      const bool pflag = (control & 0x80u) != 0u;
      p = control & 0x0Fu;
      const uint32_t cmode = (control >> 5) & 3u;
      if ((pflag && p == 0u) || (!pflag && p != 0u) || p > kMaxPartOrder) break;
      if (p > 0u && (n >> p) < kMinPart) break;
      // the partition table sits in front of the tokens (block/decoder.cpp:447-455): remember
      // where it starts and read each entry when its segment comes up
      table_pos = rd_pos(r);
      const u64 tokens_pos = table_pos + 7ull * (1u << p);
      if (tokens_pos > r.end) break;
      if ((rd_get(r, 7u) >> 5) != cmode) break;
      ok = 1u;
    } while (false);
  }
  ok = __shfl_sync(kFull, ok, 0);
  p = __shfl_sync(kFull, p, 0);
  if (!ok) return false;
  table_pos = __shfl_sync(kFull, table_pos, 0);
  const uint32_t cnt = 1u << p;
  u64 pos = table_pos + 7ull * cnt;  // first token; kept identical in all lanes
  Stage stg = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFEu, 0u, 0u};
  uint32_t off = 0u;
  for (uint32_t i = 0; i < cnt; ++i) {
    uint32_t mk = 0u;
    if (lane == 0u) {
      BitRd t = r;  // table entry i (the reader is a handful of registers)
      rd_seek(t, table_pos + 7ull * i);
      mk = rd_get(t, 7u);
    }
    mk = __shfl_sync(kFull, mk, 0);
    const uint32_t len = part_len(n, p, i);
    const uint32_t mode = mk >> 5, k0 = mk & 31u;
    if (p || mode == MODE_STATIC) {
      ok = decode_segment_fast(r, pos, len, k0, mode, out + off, sc, stg, lane) ? 1u : 0u;
#ifndef LACB_STATEFUL_SERIAL  // (measurement builds only: the serial reader for every stateful segment)
    } else if (mode != MODE_ZR) {  // stateful adaptation (p == 0), one sample per token: speculative batches
      const u64 np = decode_segment_stateful(r, pos, len, k0, mode, out + off, ring, sc, lane);
      ok = np != ~0ull;
      if (ok) pos = np;
#endif
    } else {  // stateful zero-run mode: the serial reader with the full model
      if (lane == 0u) {
        rd_seek(r, pos);
        ok = decode_segment<false>(r, len, k0, mode, out + off, ring) ? 1u : 0u;
        pos = rd_pos(r);
      }
      ok = __shfl_sync(kFull, ok, 0);
      pos = __shfl_sync(kFull, pos, 0);
    }
    if (!ok) return false;
    off += len;
  }
  if (lane == 0u) {
    rd_seek(r, pos);
    // consume_zero_padding_to_byte (bit_reader.hpp:180-185)
    const uint32_t padn = (uint32_t)((8ull - ((rd_pos(r) - r.start) & 7ull)) & 7ull);
    if (padn) {
      if (rd_get(r, padn) != 0u) ok = 0u;
      if (rd_over(r)) ok = 0u;
    }
  }
  ok = __shfl_sync(kFull, ok, 0);
  return ok != 0u;
}

// Asynchronous 16-byte copies global -> shared (LDGSTS) with commit groups: unlike register
// prefetches, whose scoreboard waits also wait for the loads issued after them, a
// wait_group leaves exactly the wanted number of newer copies in flight.
__device__ __forceinline__ void cp_async16(int4* smem_dst, const int4* gsrc) {
#ifdef LACB_EMU
  *smem_dst = *gsrc;
#else
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef LACB_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef LACB_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// exact double arithmetic on integers below 2^53 (the LPC restore chain)
__device__ __forceinline__ double fma_exact(double a, double b, double c) {
#ifdef LACB_EMU
  return a * b + c;  // |a * b| < 2^53: the product is exact, so no fused operation is needed for exactness
#else
  return __fma_rn(a, b, c);
#endif
}
__device__ __forceinline__ double floor_exact(double x) {
#if defined(LACB_EMU) || defined(LACB_FLOOR_FRND)
  return floor(x);  // (on the device: FRND.F64.FLOOR, measured 0.69 ms per restore against 0.62 ms for the two adds)
#else
  // |x| < 2^51: x + (2^52 + 2^51) rounded towards minus infinity is floor(x) + (2^52 + 2^51) -- two adds on the FP64 pipe
  return __dadd_rn(__dadd_rd(x, 6755399441055744.0), -6755399441055744.0);
#endif
}
__device__ __forceinline__ int32_t d2i_sat(double v) {  // integer-valued v; saturates like cvt.rni.s32.f64
#ifdef LACB_EMU
  return v >= 2147483647.0 ? 2147483647 : (v <= -2147483648.0 ? (int32_t)(-2147483647 - 1) : (int32_t)v);
#else
  return __double2int_rn(v);
#endif
}

constexpr int kRestoreDepth = 6;  // chunks of 8 samples in flight per thread

// Runs the recurrence over x[0..n) in order, 8 samples at a time.  One thread owns the whole
// serial chain, so
//  * memory latency of the in-place pattern (load x[i] -> compute -> store x[i]) is hidden by
//    distance: chunk c + kRestoreDepth is requested (cp.async into the thread's own staging
//    slots, `stage` with `stride` int4 between slots) when chunk c is taken out;
//  * the chain itself is kept to what the next sample needs: step(i, value&, aux&) produces the
//    sample in wrapping 32-bit arithmetic, and the int32 range verdicts of the reference are
//    computed afterwards for the whole chunk by verify(i, residual, sample, aux, h1..h4) -- eight
//    mutually independent checks instead of a check inside every link of the chain.  The 32-bit
//    samples are exact as long as no sample has failed, and a failing sample ends the block.
// Chunks move as 128-bit vectors when the plane is 16-byte aligned (every block start of a
// regular stream is).
template <typename Step, typename Verify>
__device__ __forceinline__ bool restore_chunked(int32_t* x, uint32_t n, int4* stage, uint32_t stride, Step&& step,
                                                Verify&& verify) {
  constexpr int D = kRestoreDepth;
  const bool aligned = (reinterpret_cast<uint64_t>(x) & 15ull) == 0ull;
  int32_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;  // the four samples before the current chunk
  uint32_t i = 0;
  if (aligned) {
    const uint32_t nch = n >> 3;
    int4* x4 = reinterpret_cast<int4*>(x);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if ((uint32_t)d < nch) {
        cp_async16(stage + (2 * d) * stride, x4 + 2 * d);
        cp_async16(stage + (2 * d + 1) * stride, x4 + 2 * d + 1);
      }
      cp_async_commit();  // one group per chunk, empty groups keep the count uniform
    }
    // One compact loop body (a chunk of 8 samples, the staging slot by a running index): with the body unrolled over
    // the D slots the LPC loop was ~35 KB of code, and sharing an SM with the parser warps of other slices (host
    // pipelines) it ran up to six times slower depending on where the build happened to place it (instruction fetch).
    uint32_t d = 0u;
#pragma unroll 1
    for (uint32_t c = 0; c < nch; ++c) {
      {
        {
          cp_async_wait<D - 1>();  // the oldest group (chunk c) has landed
          const int4 lo = stage[(2 * d) * stride], hi = stage[(2 * d + 1) * stride];
          if (c + D < nch) {
            cp_async16(stage + (2 * d) * stride, x4 + 2 * (c + D));
            cp_async16(stage + (2 * d + 1) * stride, x4 + 2 * (c + D) + 1);
          }
          cp_async_commit();
          const int32_t in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          int32_t v[12];  // v[0..3] = the four samples before the chunk (oldest first), v[4..11] = the chunk
          v[0] = t4; v[1] = t3; v[2] = t2; v[3] = t1;
          i64 aux[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 + j] = in[j];
            aux[j] = 0;
            step(c * 8u + (uint32_t)j, v[4 + j], aux[j]);
          }
          bool ok = true;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ok = verify(c * 8u + (uint32_t)j, in[j], v[4 + j], aux[j], v[3 + j], v[2 + j], v[1 + j], v[j]) && ok;
          if (!ok) {
            cp_async_wait<0>();
            return false;
          }
          x4[2 * c] = make_int4(v[4], v[5], v[6], v[7]);
          x4[2 * c + 1] = make_int4(v[8], v[9], v[10], v[11]);
          t4 = v[8]; t3 = v[9]; t2 = v[10]; t1 = v[11];
          d = d + 1u == (uint32_t)D ? 0u : d + 1u;
        }
      }
    }
    cp_async_wait<0>();
    i = nch << 3;
  }
  for (; i < n; ++i) {
    const int32_t in = x[i];
    int32_t t = in;
    i64 aux = 0;
    step(i, t, aux);
    if (!verify(i, in, t, aux, t1, t2, t3, t4)) return false;
    x[i] = t;
    t4 = t3; t3 = t2; t2 = t1; t1 = t;
  }
  return true;
}

// Fixed predictor of order 1..4 (block/decoder.cpp:308-352): one to three wrapping 32-bit instructions
// per link of the chain; the verdict repeats the sum in 64 bits from the finished samples.
template <int ORDER>
__device__ __forceinline__ bool restore_fixed(int32_t* x, uint32_t n, int4* stage, uint32_t stride) {
  uint32_t h1 = 0, h2 = 0, h3 = 0, h4 = 0;
  return restore_chunked(
      x, n, stage, stride,
      [&](uint32_t i, int32_t& val, i64&) {
        uint32_t p;
        if (ORDER == 1) p = h1;
        else if (ORDER == 2) p = 2u * h1 - h2;
        else if (ORDER == 3) p = 3u * h1 - 3u * h2 + h3;
        else p = 4u * h1 - 6u * h2 + 4u * h3 - h4;
        const uint32_t s = (uint32_t)val + (i >= (uint32_t)ORDER ? p : 0u);  // the first ORDER samples are verbatim
        val = (int32_t)s;
        h4 = h3; h3 = h2; h2 = h1; h1 = s;
      },
      [](uint32_t i, int32_t in, int32_t out, i64, int32_t g1, int32_t g2, int32_t g3, int32_t g4) {
        i64 p;
        if (ORDER == 1) p = g1;
        else if (ORDER == 2) p = 2 * (i64)g1 - g2;
        else if (ORDER == 3) p = 3 * (i64)g1 - 3 * (i64)g2 + g3;
        else p = 4 * (i64)g1 - 6 * (i64)g2 + 4 * (i64)g3 - g4;
        return i < (uint32_t)ORDER || (i64)in + p == (i64)out;
      });
}

// restore_*_in_place, block/decoder.cpp:308-403: every reconstructed sample must fit int32
__device__ __forceinline__ bool restore_block(int32_t* x, uint32_t n, uint32_t type, uint32_t order, const int16_t* c,
                                              int4* stage, uint32_t stride, bool lpc_fp64) {
  if (type == PRED_FIXED) {
    switch (order) {
      case 0: return true;
      case 1: return restore_fixed<1>(x, n, stage, stride);
      case 2: return restore_fixed<2>(x, n, stage, stride);
      case 3: return restore_fixed<3>(x, n, stage, stride);
      default: return restore_fixed<4>(x, n, stage, stride);
    }
  }
  if (type == PRED_FIR) {
    int32_t h1 = 0, h2 = 0;
    return restore_chunked(
        x, n, stage, stride,
        [&](uint32_t i, int32_t& val, i64&) {
          const uint32_t t = (uint32_t)((3 * (i64)h1 - h2) >> 2);
          val = (int32_t)((uint32_t)val + (i >= 2u ? t : 0u));
          h2 = h1; h1 = val;
        },
        [](uint32_t i, int32_t in, int32_t out, i64, int32_t g1, int32_t g2, int32_t, int32_t) {
          return i < 2u || (i64)in + ((3 * (i64)g1 - g2) >> 2) == (i64)out;
        });
  }
  if (order <= 12u && !lpc_fp64) {
    // The integer form of the chain below (LACB_RESTORE_FP64=0): the comparison path, 0.98 ms against 0.60 ms.
    int32_t cf[13];
#pragma unroll
    for (int t = 1; t <= 12; ++t) cf[t] = (uint32_t)t <= order ? (int32_t)c[t] : 0;
    i64 A[14];
#pragma unroll
    for (int t = 0; t < 14; ++t) A[t] = 0;
    int32_t h1 = 0;  // the previous sample
    return restore_chunked(
        x, n, stage, stride,
        [&](uint32_t, int32_t& val, i64& aux) {
#pragma unroll
          for (int t = 1; t <= 12; ++t) A[t] = mad_wide(cf[t], h1, A[t]);
          aux = A[1] >> 15;  // the prediction, kept in 64 bits for the verdict
          val = (int32_t)((uint32_t)val + (uint32_t)aux);
          h1 = val;
#pragma unroll
          for (int t = 1; t <= 12; ++t) A[t] = A[t + 1];  // A[13] stays 0
        },
        [](uint32_t, int32_t in, int32_t out, i64 aux, int32_t, int32_t, int32_t, int32_t) {
          return aux + (i64)in == (i64)out;
        });
  }
  if (order <= 12u) {
    // Transposed (systolic) form of the predictor: A[t] collects the prediction of the sample t - 1 steps
    // ahead, and every finished sample s is folded into all twelve accumulators at once,
    //   A[t] += c_t * s   (t = 1..12),
    // twelve independent multiply-adds that depend on nothing but s.  When a sample's turn comes its sum
    // is already complete except for the c_1 term, so the sample-to-sample chain is one multiply-add, one
    // funnel shift (low word of the sum >> 15) and one 32-bit add; the direct form put a four-deep chain of
    // 64-bit multiply-adds plus two 64-bit additions in front of it.  Zero history before the block start
    // reproduces taps = min(order, i); taps beyond `order` have c_t = 0.  The sums are exact in 64 bits
    // (|c_t * s| < 2^46), so the order of the additions does not matter.
    // Measured: the kernel time does not move with the shape of the chain or the prefetch depth but drops
    // from 0.98 to 0.43 ms without the arithmetic -- with one warp per scheduler the twelve 32x32->64
    // multiply-adds per sample are paid at their issue cost whatever the number of active lanes.
    // Round 2, last session: the twelve multiply-adds run on the FP64 pipe.  A 32x32->64 integer multiply-add
    // (IMAD.WIDE) issues every ~5.5 cycles from the single warp a scheduler has here (measured: 66 cycles per sample
    // for twelve), a DFMA every 2; and every quantity of the chain is an integer below 2^53 in magnitude, so doubles
    // hold it EXACTLY: |c_t * s| <= 2^15 * 2^31, a sum of twelve < 2^50, its floor(/ 2^15) and the sample < 2^36.  The
    // prediction is floor(sum * 2^-15) (exact scaling, exact floor), the sample stays a double along the chain (the
    // conversions of the residual coming in and of the sample going out are off the chain), and the int32 range
    // verdict is a comparison of the exact double.
    // The coefficients carry the 2^-15 (an exact scaling: the sums become multiples of 2^-15 below 2^35), and the
    // residual -- an integer -- is added to the accumulator before the last term arrives, so the sample is
    // floor(A[1] + c_1 * s): the chain from sample to sample is one DFMA and one floor.
    double cf[13];
#pragma unroll
    for (int t = 1; t <= 12; ++t) cf[t] = (uint32_t)t <= order ? (double)c[t] * 0.000030517578125 : 0.0;
    double A[14];
#pragma unroll
    for (int t = 0; t < 14; ++t) A[t] = 0.0;
    double h1 = 0.0;  // the previous sample
    return restore_chunked(
        x, n, stage, stride,
        [&](uint32_t, int32_t& val, i64& aux) {
          const double pre = A[1] + (double)val;  // everything but the newest sample's term (off the chain)
#pragma unroll
          for (int t = 2; t <= 12; ++t) A[t] = fma_exact(cf[t], h1, A[t]);
          const double vd = floor_exact(fma_exact(cf[1], h1, pre));
          aux = (vd < -2147483648.0 || vd > 2147483647.0) ? 1 : 0;  // the reference's int32 range verdict
          val = d2i_sat(vd);
          h1 = vd;
#pragma unroll
          for (int t = 1; t <= 12; ++t) A[t] = A[t + 1];  // A[13] stays 0
        },
        [](uint32_t, int32_t, int32_t, i64 aux, int32_t, int32_t, int32_t, int32_t) { return aux == 0; });
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

struct DecCfg {
  uint32_t channels, stereo_mode, bit_depth, n_blocks;
  uint32_t lpc_fp64;  // LPC restore chain on the FP64 pipe (restore_block)
};

// K12a: one warp per frame-block, lane 0 parses.  blk_fs[b] = first sample of block b,
// blk_size[b] its sample count, blk_boff[b] its byte offset inside `payload`,
// blk_bytes[b] its byte size.  Residuals go to the planes, predictor headers to hdrs[2b+ch].
constexpr int kParseWarps = 1;
__global__ void __launch_bounds__(32 * kParseWarps, 32) k_parse_blocks(
    DecCfg cfg, const uint8_t* __restrict__ payload, u64 payload_bytes, const u64* __restrict__ blk_fs,
    const uint32_t* __restrict__ blk_size, const u64* __restrict__ blk_boff, const uint32_t* __restrict__ blk_bytes,
    int32_t* L, int32_t* R, ChanHdr* hdrs, uint32_t* blk_err, uint8_t* blk_ms) {
  __shared__ uint32_t rings[kParseWarps][kDriftWin];
  __shared__ ParseScratch scratch[kParseWarps];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t b = blockIdx.x * kParseWarps + warp;
  if (b >= cfg.n_blocks) return;
  uint32_t* ring = rings[warp];
  ParseScratch* sc = &scratch[warp];
  const uint8_t* begin = payload + blk_boff[b];
  BitRd r;
  rd_init(r, begin, blk_bytes[b], payload + payload_bytes);
  const uint32_t n = blk_size[b];
  uint32_t err = DERR_OK;
  uint32_t ms = 0u;
  if (cfg.channels == 2u && cfg.stereo_mode == 2u) {
    const uint32_t flag = rd_get(r, 8u);  // every lane reads the same byte
    if (rd_over(r) || flag > 1u) err = DERR_FLAG;
    ms = flag == 1u;
  } else if (cfg.channels == 2u && cfg.stereo_mode == 1u) {
    ms = 1u;
  }
  if (!err && !parse_channel_block(r, n, L + blk_fs[b], hdrs + (size_t)b * 2u, ring, sc, lane)) err = DERR_PRIMARY;
  if (!err && cfg.channels == 2u &&
      !parse_channel_block(r, n, R + blk_fs[b], hdrs + (size_t)b * 2u + 1u, ring, sc, lane))
    err = DERR_SECONDARY;
  if (lane == 0u) {
    if (!err && rd_pos(r) != r.end) err = DERR_TRAILING;  // checked after reconstruction in the reference
    blk_err[b] = err;
    blk_ms[b] = (uint8_t)ms;
  }
}

// Legacy v2 streams (lac/decoder.cpp:209-218): no per-block byte sizes, so the whole frame
// payload is one bit-serial chain.  A single warp walks the blocks in order with the same
// block parser; the first failing block stops the walk.  info[0] = blocks parsed OK,
// info[1] = 1 when payload bits remain after the last block ("trailing frame payload").
__global__ void __launch_bounds__(32) k_parse_serial(DecCfg cfg, const uint8_t* __restrict__ payload, u64 payload_bytes,
                                                     u64 padded_bytes, const u64* __restrict__ blk_fs,
                                                     const uint32_t* __restrict__ blk_size, int32_t* L, int32_t* R,
                                                     ChanHdr* hdrs, uint32_t* blk_err, uint8_t* blk_ms, uint32_t* info) {
  __shared__ uint32_t ring[kDriftWin];
  __shared__ ParseScratch sc;
  const uint32_t lane = threadIdx.x & 31u;
  BitRd r;
  rd_init(r, payload, payload_bytes, payload + padded_bytes);
  uint32_t done = 0u;
  for (uint32_t b = 0; b < cfg.n_blocks; ++b) {
    uint32_t err = DERR_OK, ms = 0u;
    if (cfg.channels == 2u && cfg.stereo_mode == 2u) {
      uint32_t flag = 0u, over = 0u;
      if (lane == 0u) {
        flag = rd_get(r, 8u);
        over = rd_over(r);
      }
      flag = __shfl_sync(kFull, flag, 0);
      over = __shfl_sync(kFull, over, 0);
      if (over || flag > 1u) err = DERR_FLAG;
      ms = flag == 1u;
    } else if (cfg.channels == 2u && cfg.stereo_mode == 1u) {
      ms = 1u;
    }
    const uint32_t n = blk_size[b];
    if (!err && !parse_channel_block(r, n, L + blk_fs[b], hdrs + (size_t)b * 2u, ring, &sc, lane)) err = DERR_PRIMARY;
    if (!err && cfg.channels == 2u &&
        !parse_channel_block(r, n, R + blk_fs[b], hdrs + (size_t)b * 2u + 1u, ring, &sc, lane))
      err = DERR_SECONDARY;
    if (lane == 0u) {
      blk_err[b] = err;
      blk_ms[b] = (uint8_t)ms;
    }
    if (err) break;
    ++done;
  }
  if (lane == 0u) {
    for (uint32_t b = done + 1u; b < cfg.n_blocks; ++b) blk_err[b] = DERR_PRIMARY;  // never reached; keeps them out of later stages
    info[0] = done;
    info[1] = (done == cfg.n_blocks && rd_pos(r) != r.end) ? 1u : 0u;
  }
}

// K12b: one thread per channel-block.  A reconstruction overflow is a primary / secondary
// channel failure (Block::Decoder::decode_into returns false), which outranks the
// trailing-payload error recorded by the parser.
//
// Channel-blocks are first grouped by predictor kind (fixed order 1..4, FIR, LPC <= 12,
// LPC > 12), each group padded to a warp boundary, so the 32 lanes of a restore warp run
// the same counted loop instead of serialising three different recurrences.
__device__ __forceinline__ uint32_t restore_key(const ChanHdr& h, uint32_t err, uint32_t ch) {
  if (err == DERR_FLAG || err == DERR_PRIMARY || (err == DERR_SECONDARY && ch == 1u)) return 0u;
  if (h.type == PRED_FIXED) return h.order;  // 0: nothing to do
  if (h.type == PRED_FIR) return 5u;
  return h.order <= 12u ? 6u : 7u;
}
__global__ void __launch_bounds__(1024) k_restore_order(DecCfg cfg, const ChanHdr* __restrict__ hdrs,
                                                        const uint32_t* __restrict__ blk_err, uint32_t* order,
                                                        uint32_t* n_order) {
  __shared__ uint32_t cnt[8], cur[8];
  const uint32_t tid = threadIdx.x, nj = cfg.n_blocks * cfg.channels;
  if (tid < 8u) cnt[tid] = 0u;
  __syncthreads();
  for (uint32_t j = tid; j < nj; j += blockDim.x) {
    const uint32_t b = j / cfg.channels, ch = j - b * cfg.channels;
    const uint32_t key = restore_key(hdrs[(size_t)b * 2u + ch], blk_err[b] & 0xFFu, ch);
    if (key) atomicAdd(&cnt[key], 1u);
  }
  __syncthreads();
  if (tid == 0u) {
    uint32_t pos = 0u;
    for (uint32_t k = 1; k < 8u; ++k) {
      cur[k] = pos;
      pos += (cnt[k] + 31u) & ~31u;
    }
    *n_order = pos;
  }
  __syncthreads();
  const uint32_t total = *n_order;
  for (uint32_t i = tid; i < total; i += blockDim.x) order[i] = 0xFFFFFFFFu;
  __syncthreads();
  for (uint32_t j = tid; j < nj; j += blockDim.x) {
    const uint32_t b = j / cfg.channels, ch = j - b * cfg.channels;
    const uint32_t key = restore_key(hdrs[(size_t)b * 2u + ch], blk_err[b] & 0xFFu, ch);
    if (key) order[atomicAdd(&cur[key], 1u)] = j;
  }
}
#ifndef LACB_RESTORE_TPB
#define LACB_RESTORE_TPB 64
#endif
constexpr uint32_t kRestoreTpb = LACB_RESTORE_TPB;  // threads (channel-blocks) per restore CTA
__global__ void __launch_bounds__(LACB_RESTORE_TPB) k_restore_blocks(DecCfg cfg, const u64* __restrict__ blk_fs,
                                                       const uint32_t* __restrict__ blk_size, int32_t* L, int32_t* R,
                                                       const ChanHdr* __restrict__ hdrs, uint32_t* blk_err,
                                                       const uint32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ n_order) {
  __shared__ int4 stage[2 * kRestoreDepth][kRestoreTpb];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_order) return;
  const uint32_t j = order[i];
  if (j == 0xFFFFFFFFu) return;
  const uint32_t b = j / cfg.channels, ch = j - b * cfg.channels;
  const ChanHdr* h = hdrs + (size_t)b * 2u + ch;
  int32_t* x = (ch ? R : L) + blk_fs[b];
  if (!restore_block(x, blk_size[b], h->type, h->order, h->coef, &stage[0][threadIdx.x], kRestoreTpb, cfg.lpc_fp64 != 0u))
    atomicOr(&blk_err[b], 0x100u << ch);
}
// Folds the restore verdicts (bits 8/9) into the per-block code in the reference's order:
// primary parse, primary reconstruction, secondary parse, secondary reconstruction.
__global__ void k_merge_restore_errors(DecCfg cfg, uint32_t* blk_err) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cfg.n_blocks) return;
  const uint32_t v = blk_err[b];
  if (!(v & 0x300u)) return;
  uint32_t e = v & 0xFFu;
  if (e != DERR_FLAG && e != DERR_PRIMARY) {
    if (v & 0x100u) e = DERR_PRIMARY;
    else if (e != DERR_SECONDARY && (v & 0x200u)) e = DERR_SECONDARY;
  }
  blk_err[b] = e;
}

// M/S reconstruction + depth validation + optional interleaved packing; one CTA per block.
// out_packed == nullptr: planes are fixed up in place (LAC::Decoder::decode semantics).
__global__ void __launch_bounds__(256) k_finish_pcm(DecCfg cfg, const u64* __restrict__ blk_fs,
                                                    const uint32_t* __restrict__ blk_size, int32_t* L, int32_t* R,
                                                    uint32_t* blk_err, const uint8_t* __restrict__ blk_ms,
                                                    uint8_t* out_packed, uint32_t keep_planes) {
  const int32_t lo = cfg.bit_depth == 16u ? -32768 : -8388608, hi = cfg.bit_depth == 16u ? 32767 : 8388607;
  const uint32_t bps = cfg.bit_depth / 8u, ch = cfg.channels;
  const uint32_t mask = cfg.bit_depth == 16u ? 0xFFFFu : 0xFFFFFFu;
  // the planes only have to be fixed up when somebody reads them afterwards
  const bool planes = keep_planes != 0u || out_packed == nullptr;
  for (uint32_t b = blockIdx.x; b < cfg.n_blocks; b += gridDim.x) {
    const uint32_t e = blk_err[b];
    if (e != DERR_OK && e != DERR_TRAILING) continue;  // samples are garbage
    const u64 fs = blk_fs[b];
    const uint32_t n = blk_size[b];
    const bool ms = ch == 2u && blk_ms[b];
    uint32_t bad = 0u;
    // one frame: mid/side -> left/right (lac/decoder.cpp:48-65), depth check, stores
    auto frame = [&](uint32_t i) {
      i64 l = L[fs + i], r = 0;
      if (ch == 2u) {
        r = R[fs + i];
        if (ms) {
          const i64 m = l, sd = r;
          l = m + ((sd + (sd & 1)) >> 1);
          r = l - sd;
        }
      }
      const bool ok = l >= lo && l <= hi && (ch == 1u || (r >= lo && r <= hi));
      bad |= !ok;
      if (!ok) return;
      if (ms && planes) {
        L[fs + i] = (int32_t)l;
        R[fs + i] = (int32_t)r;
      }
      if (out_packed) {
        uint8_t* o = out_packed + (fs + i) * (u64)(bps * ch);
        for (uint32_t k = 0; k < bps; ++k) o[k] = (uint8_t)((uint32_t)(int32_t)l >> (8u * k));
        if (ch == 2u)
          for (uint32_t k = 0; k < bps; ++k) o[bps + k] = (uint8_t)((uint32_t)(int32_t)r >> (8u * k));
      }
    };
    // four frames per thread with 128-bit plane accesses and whole-word packed stores when the block
    // starts on a multiple of four frames (every block of a regular stream does)
    const bool vec = (fs & 3ull) == 0ull && (reinterpret_cast<uint64_t>(L) & 15ull) == 0ull &&
                     (ch == 1u || (reinterpret_cast<uint64_t>(R) & 15ull) == 0ull) &&
                     (reinterpret_cast<uint64_t>(out_packed) & 3ull) == 0ull;
    const uint32_t n4 = vec ? n >> 2 : 0u;
    for (uint32_t q = threadIdx.x; q < n4; q += blockDim.x) {
      int4* L4 = reinterpret_cast<int4*>(L + fs) + q;
      int4* R4 = reinterpret_cast<int4*>(R + fs) + q;
      const int4 a = *L4;
      const int4 c = ch == 2u ? *R4 : make_int4(0, 0, 0, 0);
      i64 l[4] = {a.x, a.y, a.z, a.w}, r[4] = {c.x, c.y, c.z, c.w};
      bool ok = true;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (ms) {
          const i64 m = l[k], sd = r[k];
          l[k] = m + ((sd + (sd & 1)) >> 1);
          r[k] = l[k] - sd;
        }
        ok = ok && l[k] >= lo && l[k] <= hi && (ch == 1u || (r[k] >= lo && r[k] <= hi));
      }
      if (!ok) {  // rare: let the per-frame path decide which frames stand
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) frame(4u * q + k);
        continue;
      }
      if (ms && planes) {
        *L4 = make_int4((int32_t)l[0], (int32_t)l[1], (int32_t)l[2], (int32_t)l[3]);
        *R4 = make_int4((int32_t)r[0], (int32_t)r[1], (int32_t)r[2], (int32_t)r[3]);
      }
      if (out_packed) {
        uint32_t v[8];  // the samples of the four frames in stream order, masked to the sample width
        uint32_t nv = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[nv++] = (uint32_t)(int32_t)l[k] & mask;
          if (ch == 2u) v[nv++] = (uint32_t)(int32_t)r[k] & mask;
        }
        uint32_t* o = reinterpret_cast<uint32_t*>(out_packed + (fs + 4ull * q) * (u64)(bps * ch));
        if (bps == 2u) {
          for (uint32_t k = 0; k < nv; k += 2u) o[k >> 1] = v[k] | (v[k + 1u] << 16);
        } else {  // 24-bit: four samples fill three words
          for (uint32_t k = 0, w = 0; k < nv; k += 4u, w += 3u) {
            o[w] = v[k] | (v[k + 1u] << 24);
            o[w + 1u] = (v[k + 1u] >> 8) | (v[k + 2u] << 16);
            o[w + 2u] = (v[k + 2u] >> 16) | (v[k + 3u] << 8);
          }
        }
      }
    }
    for (uint32_t i = 4u * n4 + threadIdx.x; i < n; i += blockDim.x) frame(i);
    if (__syncthreads_or((int)bad) && threadIdx.x == 0) blk_err[b] = DERR_RANGE;
  }
}

}  // namespace lacb
