// lacb_enc_kernels.cuh -- encoder kernels (sm_100a): de-interleave, stereo proxy,
// autocorrelation, Levinson, channel-block analysis, block-offset scan, emission.
#pragma once
#include "lacb_encode.cuh"
#include "lacb_f80.cuh"

namespace lacb {

// Per-block flag word shared by the encoder stages.
enum : uint32_t {
  BF_CHOOSE_MS = 1u,     // mid/side selected (stereo)
  BF_UNCERTAIN = 2u,     // estimate_stereo_mode said "uncertain"
  BF_PROBE = 4u,         // uncertain and size > 4096: decided by 3x256-sample probes
  BF_BOTH = 8u,          // uncertain and size <= 4096: both pairs encoded, smaller wins
  BF_NEED_SHIFT = 4u,    // bits 4..7: channel kinds (L,R,M,S) that need a full analysis
};

struct EncCfg {
  uint32_t channels;      // 1 or 2
  uint32_t stereo_mode;   // effective: 0 LR, 1 MS, 2 auto (0 for mono)
  uint32_t zero_run;      // set_zero_run_enabled
  uint32_t partitioning;  // set_partitioning_enabled
  uint32_t n_blocks;
};

__device__ __forceinline__ uint32_t block_len(u64 frames, uint32_t b) {
  const u64 start = (u64)b * kMaxBlock;
  const u64 left = frames - start;
  return left < kMaxBlock ? (uint32_t)left : kMaxBlock;
}

struct JobDesc {
  u64 start;
  uint32_t n;
  int kind;
};
// Full-analysis slots: 4 per block (L,R,M,S).  Probe slots: 12 per block
// (3 positions x L,R,M,S; lac/encoder.cpp:343-353).
template <bool PROBE>
__device__ __forceinline__ void job_desc(const PcmSrc& src, uint32_t slot, JobDesc& jd) {
  if (PROBE) {
    const uint32_t b = slot / 12u, q = slot - b * 12u;
    const uint32_t nb = block_len(src.frames, b);
    const uint32_t t = q >> 2;
    const uint32_t off = t == 0u ? 0u : (t == 1u ? (nb - 256u) / 2u : nb - 256u);
    jd.start = (u64)b * kMaxBlock + off;
    jd.n = 256u;
    jd.kind = (int)(q & 3u);
    return;
  }
  const uint32_t b = slot >> 2, s = slot & 3u;
  jd.start = (u64)b * kMaxBlock;
  jd.n = block_len(src.frames, b);
  jd.kind = (int)s;
}

// Compacts the active analysis slots of every block into a job list (single CTA; the
// per-GPU table has at most ~10^5 blocks).  Probe slots are active for BF_PROBE blocks,
// full slots follow the need mask.
template <bool PROBE>
__global__ void __launch_bounds__(1024) k_build_jobs(EncCfg cfg, const uint32_t* blk_flags, uint32_t* jobs,
                                                     uint32_t* job_count) {
  __shared__ uint32_t scr[40];
  __shared__ uint32_t carry;
  const uint32_t tid = threadIdx.x;
  if (tid == 0u) carry = 0u;
  __syncthreads();
  for (uint32_t base = 0; base < cfg.n_blocks; base += 1024u) {
    const uint32_t b = base + tid;
    uint32_t mask = 0u;
    if (b < cfg.n_blocks) {
      const uint32_t f = blk_flags[b];
      mask = PROBE ? ((f & BF_PROBE) ? 0xFFFu : 0u) : ((f >> BF_NEED_SHIFT) & 0xFu);
    }
    uint32_t tot;
    const uint32_t ex = block_excl_scan_u32<1024>((uint32_t)__popc(mask), scr, &tot);
    uint32_t o = carry + ex;
    for (uint32_t s = 0; s < (PROBE ? 12u : 4u); ++s)
      if ((mask >> s) & 1u) jobs[o++] = b * (PROBE ? 12u : 4u) + s;
    __syncthreads();
    if (tid == 0u) carry += tot;
    __syncthreads();
  }
  if (tid == 0u) *job_count = carry;
}

// ---------------------------------------------------------------------------
// K1: packed little-endian interleaved PCM -> int32 planes.  A thread converts four frames: their
// packed bytes are a whole number of aligned 32-bit words (8, 12, 16 or 24 bytes), the planes are
// written with 128-bit stores; the frames behind the last full group of four are done one by one.
__device__ __forceinline__ int32_t sext_sample(uint32_t v, uint32_t bps) {
  return bps == 2u ? (int32_t)(int16_t)v : ((int32_t)(v << 8)) >> 8;
}
__global__ void k_deinterleave(const uint8_t* __restrict__ in, u64 frames, uint32_t channels, uint32_t bps,
                               int32_t* __restrict__ L, int32_t* __restrict__ R) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const uint32_t fb = channels * bps;  // bytes per frame
  const bool vec = (reinterpret_cast<uint64_t>(in) & 3ull) == 0ull && (reinterpret_cast<uint64_t>(L) & 15ull) == 0ull &&
                   (channels == 1u || (reinterpret_cast<uint64_t>(R) & 15ull) == 0ull);
  const u64 groups = vec ? frames >> 2 : 0ull;
  const uint32_t mask = bps == 2u ? 0xFFFFu : 0xFFFFFFu;
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in + g * 4ull * fb);
    uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};  // the samples of the four frames in stream order
    const uint32_t ns = 4u * channels;
    if (bps == 2u) {
      for (uint32_t k = 0; k < ns; k += 2u) {
        const uint32_t x = w[k >> 1];
        v[k] = x & mask;
        v[k + 1u] = x >> 16;
      }
    } else {  // 24-bit: three words hold four samples
      for (uint32_t k = 0, i = 0; k < ns; k += 4u, i += 3u) {
        const uint32_t a = w[i], b = w[i + 1u], c = w[i + 2u];
        v[k] = a & mask;
        v[k + 1u] = ((a >> 24) | (b << 8)) & mask;
        v[k + 2u] = ((b >> 16) | (c << 16)) & mask;
        v[k + 3u] = c >> 8;
      }
    }
    if (channels == 2u) {
      reinterpret_cast<int4*>(L)[g] = make_int4(sext_sample(v[0], bps), sext_sample(v[2], bps), sext_sample(v[4], bps),
                                                sext_sample(v[6], bps));
      reinterpret_cast<int4*>(R)[g] = make_int4(sext_sample(v[1], bps), sext_sample(v[3], bps), sext_sample(v[5], bps),
                                                sext_sample(v[7], bps));
    } else {
      reinterpret_cast<int4*>(L)[g] = make_int4(sext_sample(v[0], bps), sext_sample(v[1], bps), sext_sample(v[2], bps),
                                                sext_sample(v[3], bps));
    }
  }
  for (u64 f = groups * 4ull + (u64)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride) {
    const uint8_t* p = in + f * fb;
    for (uint32_t c = 0; c < channels; ++c) {
      int32_t v;
      if (bps == 2u) v = (int16_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
      else v = ((int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24))) >> 8;
      (c ? R : L)[f] = v;
      p += bps;
    }
  }
}

// depth-range validation of planar input (lac/encoder.cpp:85-93,232-241)
__global__ void k_validate(PcmSrc src, uint32_t depth, uint32_t* bad) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  const int32_t lo = depth == 16u ? -32768 : -8388608, hi = depth == 16u ? 32767 : 8388607;
  uint32_t any = 0u;
  for (u64 f = (u64)blockIdx.x * blockDim.x + threadIdx.x; f < src.frames; f += stride) {
    const int32_t l = src.L[f];
    any |= (l < lo) | (l > hi);
    if (src.R) {
      const int32_t r = src.R[f];
      any |= (r < lo) | (r > hi);
    }
  }
  if (__any_sync(kFull, (int)any) && (threadIdx.x & 31u) == 0u) atomicOr(bad, 1u);
}

// ---------------------------------------------------------------------------
// K2: estimate_stereo_mode (lac/encoder.cpp:126-197).  One CTA per block; 12
// saturating sums (raw / first difference / first sum for L, R, M, S).
__device__ __forceinline__ u64 zz64(i64 v) { return v >= 0 ? ((u64)v << 1) : ((((u64)(-(v + 1))) << 1) | 1ull); }
__device__ __forceinline__ u64 sat_add(u64 a, u64 b) { return (b > ~0ull - a) ? ~0ull : a + b; }
__device__ __forceinline__ u64 proxy_bits(u64 sum, u64 count) {
  if (count == 0ull) return 0ull;
  const u64 mean = (sum + (count >> 1)) / count;
  uint32_t k = 0u;
  while (k < 31u && (1ull << k) < mean) ++k;
  return sat_add(sum >> k, count * (u64)(k + 1u));
}
// The twelve sums of one block.  NARROW: every |sample| <= 2^24 (all 16 / 24-bit audio), so a term is below 2^27
// and the arithmetic fits 32 bits; a thread takes four consecutive frames per step (128-bit plane loads when the block
// is aligned) and folds its 32-bit partial sums into the 64-bit ones every fourth step (16 terms < 2^31).  The wide
// form is the reference's int64 arithmetic sample by sample.  Both return the OR of the sample magnitudes seen.
template <bool NARROW>
__device__ __forceinline__ uint32_t stereo_proxy_sums(const PcmSrc& src, u64 start, uint32_t n, u64 (&s)[12]) {
  const uint32_t tid = threadIdx.x;
  uint32_t mag = 0u;
#pragma unroll
  for (int i = 0; i < 12; ++i) s[i] = 0ull;
  if (NARROW) {
    const bool vec = ((reinterpret_cast<uint64_t>(src.L + start) | reinterpret_cast<uint64_t>(src.R + start)) & 15ull) == 0ull;
    uint32_t a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = 0u;
    uint32_t step = 0u;
    for (uint32_t q = tid; q * 4u < n; q += blockDim.x, ++step) {
      int32_t l[5], r[5];  // [0] = the frame before the quad (zero before the block start)
      const uint32_t i0 = q * 4u;
      l[0] = i0 ? src.L[start + i0 - 1u] : 0;
      r[0] = i0 ? src.R[start + i0 - 1u] : 0;
      if (vec && i0 + 4u <= n) {
        const int4 vl = reinterpret_cast<const int4*>(src.L + start)[q], vr = reinterpret_cast<const int4*>(src.R + start)[q];
        l[1] = vl.x; l[2] = vl.y; l[3] = vl.z; l[4] = vl.w;
        r[1] = vr.x; r[2] = vr.y; r[3] = vr.z; r[4] = vr.w;
      } else {
#pragma unroll
        for (uint32_t m = 0; m < 4u; ++m) {
          const bool in = i0 + m < n;
          l[1 + m] = in ? src.L[start + i0 + m] : 0;
          r[1 + m] = in ? src.R[start + i0 + m] : 0;
        }
      }
      mag |= (uint32_t)(l[0] ^ (l[0] >> 31)) | (uint32_t)(r[0] ^ (r[0] >> 31));
#pragma unroll
      for (int m = 1; m <= 4; ++m) {
        mag |= (uint32_t)(l[m] ^ (l[m] >> 31)) | (uint32_t)(r[m] ^ (r[m] >> 31));
        if (i0 + (uint32_t)m - 1u >= n) continue;  // past the block end (only in the last quad)
        const int32_t v[4] = {l[m], r[m], (l[m] + r[m]) >> 1, l[m] - r[m]};
        const int32_t pv[4] = {l[m - 1], r[m - 1], (l[m - 1] + r[m - 1]) >> 1, l[m - 1] - r[m - 1]};
        const bool first = i0 + (uint32_t)m - 1u == 0u;  // the block's first sample: no neighbour, all three forms are raw
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t raw = zz32(v[c]);
          a[c] += raw;
          a[4 + c] += first ? raw : zz32(v[c] - pv[c]);
          a[8 + c] += first ? raw : zz32(v[c] + pv[c]);
        }
      }
      if ((step & 3u) == 3u) {
#pragma unroll
        for (int i = 0; i < 12; ++i) { s[i] += a[i]; a[i] = 0u; }
      }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] += a[i];
    return mag;
  }
  for (uint32_t i = tid; i < n; i += blockDim.x) {
    i64 v[4], pv[4];
    const i64 l = src.L[start + i], r = src.R[start + i];
    v[0] = l; v[1] = r; v[2] = (l + r) >> 1; v[3] = l - r;
    if (i > 0u) {
      const i64 pl = src.L[start + i - 1u], prr = src.R[start + i - 1u];
      pv[0] = pl; pv[1] = prr; pv[2] = (pl + prr) >> 1; pv[3] = pl - prr;
    } else {
      pv[0] = pv[1] = pv[2] = pv[3] = 0;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // per-sample terms are < 2^35, so plain adds cannot wrap within a block
      s[c] += zz64(v[c]);
      s[4 + c] += (i == 0u) ? zz64(v[c]) : zz64(v[c] - pv[c]);
      s[8 + c] += (i == 0u) ? zz64(v[c]) : zz64(v[c] + pv[c]);
    }
  }
  return 0u;
}
__global__ void __launch_bounds__(256) k_stereo_proxy(PcmSrc src, EncCfg cfg, uint32_t* blk_flags) {
  __shared__ u64 acc[12];
  const uint32_t tid = threadIdx.x;
  for (uint32_t b = blockIdx.x; b < cfg.n_blocks; b += gridDim.x) {
    const uint32_t n = block_len(src.frames, b);
    const u64 start = (u64)b * kMaxBlock;
    if (tid < 12u) acc[tid] = 0ull;
    u64 s[12];
    // the 32-bit form first; a block that holds a sample beyond +-2^24 (planar int32 input that was not validated
    // against a bit depth) is summed again in 64 bits -- the vote is also the barrier behind the zeroing of acc
    const uint32_t mag = stereo_proxy_sums<true>(src, start, n, s);
    if (__syncthreads_or((int)((mag >> 24) != 0u))) stereo_proxy_sums<false>(src, start, n, s);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const u64 t = warp_sum_u64(s[i]);
      if ((tid & 31u) == 0u) atomicAdd(&acc[i], t);
    }
    __syncthreads();
    if (tid == 0u) {
      u64 bits[4];
      bool nondiff = false;
      for (int c = 0; c < 4; ++c) {
        const u64 rb = proxy_bits(acc[c], n), db = proxy_bits(acc[4 + c], n), ab = proxy_bits(acc[8 + c], n);
        const u64 m = rb < db ? rb : db;
        bits[c] = m < ab ? m : ab;
        if (rb < db || ab < db) nondiff = true;
      }
      const u64 lr = sat_add(bits[0], bits[1]), ms = sat_add(bits[2], bits[3]);
      const u64 smaller = lr < ms ? lr : ms;
      const u64 diff = lr >= ms ? lr - ms : ms - lr;
      const bool choose_ms = ms < lr;
      const bool uncertain = smaller == 0ull || diff == 0ull || nondiff || diff <= smaller / 100ull;
      uint32_t f = (choose_ms ? BF_CHOOSE_MS : 0u);
      if (uncertain) {
        f |= BF_UNCERTAIN;
        if (n > 4096u) f |= BF_PROBE;
        else f |= BF_BOTH | (0xFu << BF_NEED_SHIFT);
      } else {
        f |= (choose_ms ? 0xCu : 0x3u) << BF_NEED_SHIFT;
      }
      blk_flags[b] = f;
    }
    __syncthreads();
  }
}

// flags for mono / forced stereo modes
__global__ void k_plan_fixed(EncCfg cfg, uint32_t* blk_flags) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < cfg.n_blocks; b += gridDim.x * blockDim.x) {
    uint32_t f;
    if (cfg.channels == 1u) f = 0x1u << BF_NEED_SHIFT;
    else if (cfg.stereo_mode == 1u) f = BF_CHOOSE_MS | (0xCu << BF_NEED_SHIFT);
    else f = 0x3u << BF_NEED_SHIFT;
    blk_flags[b] = f;
  }
}

// After the probes: sum of the six LR and six MS probe sizes decides (lac/encoder.cpp:343-353)
__global__ void k_decide_probes(EncCfg cfg, uint32_t* blk_flags, const uint32_t* probe_bytes) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < cfg.n_blocks; b += gridDim.x * blockDim.x) {
    uint32_t f = blk_flags[b];
    if (!(f & BF_PROBE)) continue;
    u64 lr = 0, ms = 0;
    for (uint32_t t = 0; t < 3u; ++t) {
      lr += (u64)probe_bytes[b * 12u + t * 4u + 0u] + probe_bytes[b * 12u + t * 4u + 1u];
      ms += (u64)probe_bytes[b * 12u + t * 4u + 2u] + probe_bytes[b * 12u + t * 4u + 3u];
    }
    const bool choose_ms = ms < lr;
    f &= ~(BF_CHOOSE_MS | (0xFu << BF_NEED_SHIFT));
    f |= choose_ms ? (BF_CHOOSE_MS | (0xCu << BF_NEED_SHIFT)) : (0x3u << BF_NEED_SHIFT);
    blk_flags[b] = f;
  }
}

// ---------------------------------------------------------------------------
// K5: exact int64 autocorrelation, lags 0..12 (lpc.cpp:80-96), one CTA per job.
template <int NT, int E, bool PROBE>
__global__ void __launch_bounds__(NT) k_autocorr(PcmSrc src, const uint32_t* jobs, const uint32_t* job_count, i64* acor) {
  LACB_DYN_SMEM(unsigned char, smraw);
  __shared__ u64 red[13];
  int32_t* X = reinterpret_cast<int32_t*>(smraw);
  const uint32_t tid = threadIdx.x;
  const uint32_t nj = *job_count;
  for (uint32_t ji = blockIdx.x; ji < nj; ji += gridDim.x) {
    const uint32_t slot = jobs[ji];
    JobDesc jd;
    job_desc<PROBE>(src, slot, jd);
    {
      ASmem<NT, E> xs{smraw};  // only the X plane of the layout is used (and allocated) here
      load_block<NT, E>(xs, src, jd.kind, jd.start, jd.n);
    }
    if (tid < 13u) red[tid] = 0ull;
    __syncthreads();
    int32_t x[E + 12];
    {
      const int4* X4 = reinterpret_cast<const int4*>(X);
      const int q0 = (int)tid * (E / 4);
#pragma unroll
      for (int c = -3; c < E / 4; ++c) {
        int4 v = make_int4(0, 0, 0, 0);
        if (q0 + c >= 0) v = X4[swz_chunk((uint32_t)(q0 + c))];
        x[12 + 4 * c + 0] = v.x; x[12 + 4 * c + 1] = v.y; x[12 + 4 * c + 2] = v.z; x[12 + 4 * c + 3] = v.w;
      }
    }
    u64 s[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) s[k] = 0ull;
#pragma unroll
    for (int j = 0; j < E; ++j) {
      // samples past n and before 0 are zero in X, so every term outside the sums vanishes
#pragma unroll
      for (int k = 0; k < 13; ++k) s[k] = (u64)mad_wide(x[12 + j], x[12 + j - k], (i64)s[k]);
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      const u64 t = warp_sum_u64_full(s[k]);  // products of int32 samples: all 64 bits in use
      if ((tid & 31u) == 0u) atomicAdd(&red[k], t);
    }
    __syncthreads();
    if (tid < 13u) acor[(size_t)slot * 13u + tid] = (i64)red[tid];
    __syncthreads();
  }
}

// K5 for whole channel-blocks, streaming form.  The sums are exact wrapping int64, so any summation order gives the
// reference's value; nothing has to be staged: a thread reads its 16 samples and the 12 before them straight from the
// planes (the halo is its neighbour's own chunk, an L1 / L2 hit), several small CTAs share an SM so that the loads of
// one hide under the multiply-adds of the others, and a job costs one reduction however long the block is.  (The
// one-CTA-per-SM form above waits for its 64 KB block, computes, reduces: 0.52 ms for 7032 channel-blocks, six times
// its arithmetic.)
template <int NT>
__global__ void __launch_bounds__(NT, 2) k_autocorr_stream(PcmSrc src, const uint32_t* jobs, const uint32_t* job_count,
                                                           i64* acor) {
  constexpr int E = 16;
  __shared__ u64 red[NT / 32][13];
  const uint32_t tid = threadIdx.x;
  const uint32_t nj = *job_count;
  for (uint32_t ji = blockIdx.x; ji < nj; ji += gridDim.x) {
    const uint32_t slot = jobs[ji];
    JobDesc jd;
    job_desc<false>(src, slot, jd);
    const uint32_t n = jd.n;
    const int kind = jd.kind;
    const int32_t* pa = (kind == 1 ? src.R : src.L) + jd.start;
    const int32_t* pb = (kind >= 2 ? src.R : src.L) + jd.start;  // second operand of mid / side
    const bool vec_ok = ((reinterpret_cast<uint64_t>(pa) | reinterpret_cast<uint64_t>(pb)) & 15ull) == 0ull;
    u64 s[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) s[k] = 0ull;
    for (uint32_t c0 = tid; c0 * (uint32_t)E < n; c0 += NT) {
      int4 a[E / 4 + 3], b[E / 4 + 3];
      const int q0 = (int)c0 * (E / 4);
#pragma unroll
      for (int c = -3; c < E / 4; ++c) {
        const int q = q0 + c;
        int4 va = make_int4(0, 0, 0, 0), vb = make_int4(0, 0, 0, 0);
        if (q >= 0 && (uint32_t)q * 4u < n) {
          if (vec_ok && (uint32_t)q * 4u + 4u <= n) {
            va = reinterpret_cast<const int4*>(pa)[q];
            if (kind >= 2) vb = reinterpret_cast<const int4*>(pb)[q];
          } else {  // unaligned planes, or the quad straddling the end of the block
            int32_t t[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
            for (uint32_t m = 0; m < 4u; ++m)
              if ((uint32_t)q * 4u + m < n) {
                t[m] = pa[(uint32_t)q * 4u + m];
                if (kind >= 2) w[m] = pb[(uint32_t)q * 4u + m];
              }
            va = make_int4(t[0], t[1], t[2], t[3]);
            vb = make_int4(w[0], w[1], w[2], w[3]);
          }
        }
        a[c + 3] = va;
        b[c + 3] = vb;
      }
      int32_t x[E + 12];
#pragma unroll
      for (int c = 0; c < E / 4 + 3; ++c) {
        // kind 0 / 1: the plane itself; 2: mid, 3: side (combine_sample; a zero quad stays zero)
        const int k2 = kind >= 2 ? kind : 0;
        x[4 * c + 0] = combine_sample(k2, a[c].x, b[c].x);
        x[4 * c + 1] = combine_sample(k2, a[c].y, b[c].y);
        x[4 * c + 2] = combine_sample(k2, a[c].z, b[c].z);
        x[4 * c + 3] = combine_sample(k2, a[c].w, b[c].w);
      }
#pragma unroll
      for (int j = 0; j < E; ++j) {
#pragma unroll
        for (int k = 0; k < 13; ++k) s[k] = (u64)mad_wide(x[12 + j], x[12 + j - k], (i64)s[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) {
      const u64 t = warp_sum_u64_full(s[k]);
      if ((tid & 31u) == 0u) red[tid >> 5][k] = t;
    }
    __syncthreads();
    if (tid < 13u) {
      u64 t = 0ull;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) t += red[w][tid];
      acor[(size_t)slot * 13u + tid] = (i64)t;
    }
    __syncthreads();
  }
}

// K6: Levinson-Durbin + Q15 quantisation, one thread per job (lpc.cpp:98-186).
template <bool PROBE>
__global__ void k_levinson(PcmSrc src, const uint32_t* jobs, const uint32_t* job_count, const i64* acor, LpcQ* lpcq) {
  const uint32_t nj = *job_count;
  for (uint32_t ji = blockIdx.x * blockDim.x + threadIdx.x; ji < nj; ji += gridDim.x * blockDim.x) {
    const uint32_t slot = jobs[ji];
    JobDesc jd;
    job_desc<PROBE>(src, slot, jd);
    const uint32_t max_valid = jd.n > 1u ? (jd.n - 1u < 32u ? jd.n - 1u : 32u) : 0u;
    int max_order = 0;
    for (int co = 4; co <= 12; co += 2)
      if ((uint32_t)co <= max_valid) max_order = co;
    LpcQ out;
    for (int c = 0; c < 5; ++c) out.pad[c] = 0;
    if (max_order == 0) {
      for (int c = 0; c < 5; ++c) {
        out.used[c] = 0;
        for (int i = 0; i < 13; ++i) out.coef[c][i] = 0;
      }
    } else {
      i64 R[13];
      for (int i = 0; i < 13; ++i) R[i] = acor[(size_t)slot * 13u + i];
      levinson_q15(R, max_order, out.coef, out.used);
    }
    lpcq[slot] = out;
  }
}

// ---------------------------------------------------------------------------
// K4+K7+K8+K9: channel-block analysis.
template <int NT, int E>
__device__ __forceinline__ u64 block_sum_u64(const ASmem<NT, E>& sm, u64 v) {
  AMisc* mi = sm.Misc();
  if (LACB_TID == 0) mi->red64 = 0ull;
  LACB_SYNC();
  const u64 t = warp_sum_u64(v);
  if ((LACB_TID & 31u) == 0u && t) atomicAdd(&mi->red64, t);
  LACB_SYNC();
  const u64 r = mi->red64;
  LACB_SYNC();
  return r;
}

template <int NT, int E>
__device__ __forceinline__ bool compute_residual(const int32_t (&x)[E + 12], uint32_t g0, uint32_t n, uint32_t type,
                                                 uint32_t order, uint32_t taps, const int16_t* coef, int32_t (&r)[E]) {
  if (type == PRED_FIXED) {
    residual_fixed<E>(x, g0, n, (int)order, r);
    return false;
  }
  if (type == PRED_FIR) {
    residual_fir<E>(x, g0, n, r);
    return false;
  }
  return residual_lpc<E, false>(x, g0, n, coef, (int)taps, r);  // the analysis already settled the taps
}

//
// Gangs (NT == 32, the 256-sample stereo probes).  A probe is analysed by one warp, and its search is this same code --
// hundreds of KB of instructions.  As independent one-warp CTAs the 32 resident probes of an SM each sat somewhere else
// in that code and the kernel starved on instruction fetches (ncu: 37 warp-cycles of "no instruction" per issued
// instruction, issue slots 16 % busy).  So the probe warps of an SM form one CTA (a gang, blockDim.x / 32 sub-blocks
// with their own slice of the dynamic shared memory) and are put back in step by CTA barriers at the top of every job,
// every candidate evaluation and every partition level: at any moment they all run the same phase and fetch the same
// code.  Inside a sub-block "thread", "barrier" and "vote" mean lane, warp barrier and warp vote (blk_tid / blk_sync).
// The last, partly filled row of jobs is completed with copies of the last job (same result, written twice), so every
// sub-block passes the same barriers.
template <int NT, int E, bool PROBE>
__global__ void __launch_bounds__(NT == 32 ? 1024 : NT, 1) k_analyze(PcmSrc src, EncCfg cfg, const uint32_t* jobs, const uint32_t* job_count,
                                                const LpcQ* lpcq, ChanRec* recs, uint32_t* probe_bytes) {
  LACB_DYN_SMEM(unsigned char, smraw);
  constexpr bool GANG = NT == 32;
  const uint32_t gang = GANG ? blockDim.x >> 5 : 1u, sub = GANG ? threadIdx.x >> 5 : 0u;
  ASmem<NT, E> sm{smraw + (size_t)sub * ASmem<NT, E>::BYTES};
  AMisc* mi = sm.Misc();
  const uint32_t tid = LACB_TID, g0 = tid * E;
  const uint32_t nj = *job_count;
  for (uint32_t jrow = blockIdx.x * gang; jrow < nj; jrow += gridDim.x * gang) {
    const uint32_t ji = jrow + sub < nj ? jrow + sub : nj - 1u;
    if (GANG) __syncthreads();
    const uint32_t slot = jobs[ji];
    JobDesc jd;
    job_desc<PROBE>(src, slot, jd);
    const uint32_t n = jd.n;
    LACB_PH_INIT(PROBE ? -1 : 0);
    if (tid < 11u) mi->cand_lb[tid] = 0u;  // ordered before the pre-pass by the barrier of the vote below
    // block-uniform: some sample needs more than 26 magnitude bits (never true for 16 / 24-bit audio);
    // only then can an LPC residual leave int32 (lpc.cpp:38-61) and the fallback orders matter
    const bool xbig = LACB_SYNC_OR((int)((load_block<NT, E>(sm, src, jd.kind, jd.start, n) >> 26) != 0u)) != 0;
    LACB_PH(0);
    const uint32_t max_valid = n > 1u ? (n - 1u < 32u ? n - 1u : 32u) : 0u;
    const LpcQ* lq = lpcq + slot;

    // ---- which candidates exist: fixed 0..4, FIR, LPC 4,6,8,10,12 (block/encoder.cpp:362-407), and the taps an LPC
    // candidate's residual uses (compute_residual_q15's first attempt, lpc.cpp:188-229)
    auto cand_taps = [&](uint32_t ci) -> uint32_t {  // 0 = candidate does not exist; fixed / FIR report 1
      if (ci <= 5u) return 1u;
      const uint32_t c = ci - 6u, co = 4u + 2u * c;
      if (co > max_valid) return 0u;
      const uint32_t used = (uint32_t)lq->used[c];  // 0: unstable, block/encoder.cpp:394-396
      return used < co ? used : co;
    };

    // ---- pre-pass: an exact lower bound of every candidate's cost, before anything expensive.
    // Whatever the mode and k, a sample costs at least bit_width(u) + 1 bits (32 for u >= 2^31, where the k = 31
    // estimate drops the quotient; 3 for u = 4 as a bin code; 0 for a zero inside a run), so the sum over the
    // residual bounds min(rice, static, zero-run, bin) from below.  The candidates are then evaluated in ascending
    // order of their bound: the likely winner comes first, and as soon as a bound exceeds the best exact cost so
    // far that candidate and all later ones cannot win or tie (block/encoder.cpp:352-359) and are never evaluated --
    // they pay for this pass only, not for the residual / scan / bit-plane counts of a full evaluation.
    // Skipped when an LPC residual could leave int32 (|x| >= 2^26, never for 16 / 24-bit audio): the fallback
    // orders then need block-wide votes, and every candidate is simply evaluated in index order.
    if (!xbig) {
      auto bound = [&](const int32_t (&r)[E]) -> uint32_t {
        // |x| < 2^26 here, so every |residual| < 2^30, u < 2^31 and clz(u) >= 1: bit_width(u) + 1 = 33 - clz(u).
        // Zeros (0 bits inside a run; missing samples past the block end are zeros too) and fours (3 bits as a
        // bin code) need a correction, looked for only in chunks that hold a value <= 4.
        uint32_t a = 0u, umin = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < E; ++j) {
          const uint32_t uu = zz32(r[j]);
          a += (uint32_t)__clz((int)uu);
          umin = uu < umin ? uu : umin;
        }
        uint32_t lb = 33u * (uint32_t)E - a;
        if (umin <= 4u) {
#pragma unroll
          for (int j = 0; j < E; ++j) lb -= ((zz32(r[j]) & ~4u) == 0u) ? 1u : 0u;  // u == 0 or u == 4
        }
        return lb;
      };
      auto publish = [&](uint32_t ci, uint32_t lbt) {
        const uint32_t t = warp_sum_u32(lbt);
        if ((tid & 31u) == 0u && t) atomicAdd(&mi->cand_lb[ci], t);
      };
      {
        {
          int32_t x[E + 12];
          load_items<NT, E>(sm, x);
          int32_t r[E];
          residual_fixed<E>(x, g0, n, 0, r); publish(0u, bound(r));
          residual_fixed<E>(x, g0, n, 1, r); publish(1u, bound(r));
          residual_fixed<E>(x, g0, n, 2, r); publish(2u, bound(r));
          residual_fixed<E>(x, g0, n, 3, r); publish(3u, bound(r));
          residual_fixed<E>(x, g0, n, 4, r); publish(4u, bound(r));
          residual_fir<E>(x, g0, n, r); publish(5u, bound(r));
        }
#pragma unroll 1
        for (uint32_t ci = 6u; ci < 11u; ++ci) {
          const uint32_t taps = cand_taps(ci);
          if (taps == 0u) continue;
          // The samples are re-read from the X plane for every order: kept across the orders, the 28 of them next to
          // the coefficients, the accumulator and the 16 residuals overflow the 64-register budget -- a third of the
          // kernel's spill instructions came from here, and spills go to L2 (212 KB of shared memory leave ~no L1):
          // analyze 11.64 -> 11.27 ms.  (The bound straight from the taps, without the residual array, measured 11.35.)
          int32_t x[E + 12], r[E];
          load_items<NT, E>(sm, x);
          residual_lpc<E, false>(x, g0, n, lq->coef[ci - 6u], (int)taps, r);
          publish(ci, bound(r));
        }
      }
      LACB_SYNC();
    }
    if (tid < 11u) {
      // rank of candidate `tid` among the existing ones by (bound, index); without the pre-pass all bounds are 0
      const uint32_t mine = mi->cand_lb[tid];
      const bool exists = cand_taps(tid) != 0u;
      uint32_t rank = 0u, total = 0u;
      for (uint32_t cj = 0; cj < 11u; ++cj) {
        if (cand_taps(cj) == 0u) continue;
        ++total;
        const uint32_t other = mi->cand_lb[cj];
        if (other < mine || (other == mine && cj < tid)) ++rank;
      }
      if (exists) mi->cand_order[rank] = (uint8_t)tid;
      if (tid == 0u) mi->cand_n = total;
      if (!PROBE) recs[slot].cand_lo[tid] = exists ? 0xFFFFFFFEu : 0xFFFFFFFFu;  // overwritten when evaluated
    }
    if (tid == 0u) mi->best.have = 0u;
    if (tid == 0u) mi->hq_kb_n = 0u;
    LACB_SYNC();
    const uint32_t n_cand = mi->cand_n;
    bool have_best = false;
    u64 best_bits = 0ull;
    uint32_t best_ci = 0u;
    bool cand_done = false;
    for (uint32_t idx = 0;; ++idx) {
      uint32_t ci = 0u;
      if (idx >= n_cand) cand_done = true;
      if (!cand_done) {
        ci = mi->cand_order[idx];
        if (have_best && (u64)mi->cand_lb[ci] > best_bits) cand_done = true;  // this one and all later ones are out
      }
      if (GANG) {  // the gang goes on while any of its probes has a candidate left; the others wait here
        if (!__syncthreads_or((int)!cand_done)) break;
        if (cand_done) continue;
      } else if (cand_done) {
        break;
      }
      int32_t r[E];
      uint32_t type, order, taps = 0u;
      {
        // the samples are re-read from the X plane for every candidate instead of living in
        // 28 registers across the whole search
        int32_t x[E + 12];
        load_items<NT, E>(sm, x);
        if (ci <= 4u) {
          type = PRED_FIXED;
          order = ci;
          residual_fixed<E>(x, g0, n, (int)ci, r);
        } else if (ci == 5u) {
          type = PRED_FIR;
          order = 2u;
          residual_fir<E>(x, g0, n, r);
        } else {
          const uint32_t c = ci - 6u, co = 4u + 2u * c;
          type = PRED_LPC;
          order = co;
          // compute_residual_q15 attempts (lpc.cpp:188-229): used, then {12,10,8,6,4} below it
          uint32_t attempt = cand_taps(ci);
          if (!xbig) {
            residual_lpc<E, false>(x, g0, n, lq->coef[c], (int)attempt, r);
            taps = attempt;
            attempt = 0u;
          }
          while (attempt > 0u) {
            const bool ovf = residual_lpc<E, true>(x, g0, n, lq->coef[c], (int)attempt, r);
            if (!LACB_SYNC_OR((int)ovf)) {
              taps = attempt;
              break;
            }
            uint32_t next = 0u;
            for (uint32_t fo = 12u; fo >= 4u; fo -= 2u)
              if (fo < attempt && fo <= co) {
                next = fo;
                break;
              }
            attempt = next;
          }
          if (taps == 0u) continue;  // block/encoder.cpp:402-404
        }
      }
      Prep<NT, E> pr;
      LACB_PH(1);
      prepare<NT, E, false>(sm, r, n, pr);
      // (the block's initial / static k are evaluated inside the pass, see block_static_k)
      const uint32_t has_run = cost_pass<NT, E, true>(sm, pr, n, 0u, 0u);
      {
        // Every thread scores the candidate from the block totals (published by the barrier that ends the pass; the
        // next writes to them lie behind the barriers of the next evaluation) and keeps the running best in
        // registers, so the loop-exit test above is block-uniform without another barrier.  One lane keeps the
        // full record of the best candidate in shared memory for the partition search.
        const u64 stat = mi->stat_bits;
        const u64 rice = split_sum_get(&mi->tot_rice), bin = split_sum_get(&mi->tot_bin);
        const u64 zr = (cfg.zero_run && has_run) ? split_sum_get(&mi->tot_zr) : rice;  // block/encoder.cpp:343-345
        const u64 m1 = rice < stat ? rice : stat, m2 = zr < bin ? zr : bin;
        const u64 bb = m1 < m2 ? m1 : m2;
        // strictly fewer bits wins; on a tie the lower predictor type, then the earlier candidate (:352-359) --
        // the type never decreases with the index, so that is the lower index whatever the evaluation order
        if (!have_best || bb < best_bits || (bb == best_bits && ci < best_ci)) {
          have_best = true;
          best_bits = bb;
          best_ci = ci;
          if (tid == (uint32_t)(NT - 32)) {  // the last warp: the first warps are the loaded ones
            BestCand& best = mi->best;
            best.have = 1u;
            best.rice = rice; best.zr = zr; best.bin = bin; best.stat = stat; best.best = bb;
            best.type = type; best.order = order; best.taps = taps; best.ci = ci;
            best.k_init = mi->k_init; best.k_stat = mi->k_stat; best.has_run = has_run;
          }
        }
        if (!PROBE && tid == (uint32_t)(NT - 32)) recs[slot].cand_lo[ci] = (uint32_t)bb;
      }
    }
    LACB_PH(1);
    LACB_SYNC();
    LACB_PH_BASE(PROBE ? -1 : 20);
    const BestCand best = mi->best;
    if (GANG) __syncthreads();

    // winner residual again, with the full prefix structures for the partition search
    const int16_t* wcoef = best.type == PRED_LPC ? lq->coef[best.ci - 6u] : nullptr;
    int32_t r[E];
    {
      int32_t x[E + 12];
      load_items<NT, E>(sm, x);
      compute_residual<NT, E>(x, g0, n, best.type, best.order, best.taps, wcoef, r);
    }
    Prep<NT, E> pr;
    LACB_PH(1);
    prepare<NT, E, true>(sm, r, n, pr);

    const uint32_t max_p = (cfg.partitioning && n >= kMinPart) ? max_partition_order(n) : 0u;
    // per-segment initial k, static k / bits and prefix of u, all levels at once
    for (uint32_t sid = tid; sid < (2u << max_p) - 1u; sid += NT) {
      const uint32_t p = 31u - (uint32_t)__clz((int)(sid + 1u));
      const uint32_t s = sid + 1u - (1u << p);
      const uint32_t base = n >> p, cnt = 1u << p;
      const uint32_t a = s * base, b = (s + 1u == cnt) ? n : a + base;
      const uint32_t len = b - a, f = a + (len < 256u ? len : 256u);
      PlaneCounts ca, cf, cb;
      prefix_counts<NT, E>(sm, a, ca);
      prefix_counts<NT, E>(sm, f, cf);
      prefix_counts<NT, E>(sm, b, cb);
      const u64 pa = prefix_u<NT, E>(sm, a), pf = prefix_u<NT, E>(sm, f), pb = prefix_u<NT, E>(sm, b);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        cf.w[w] -= ca.w[w];
        cb.w[w] -= ca.w[w];
      }
      u64 sb;
      const uint32_t ki = best_static_k(pf - pa, cf, f - a, 12, nullptr);
      const uint32_t ks = best_static_k(pb - pa, cb, len, 15, &sb);
      sm.SegP()[sid] = pa;
      sm.SegStat()[sid] = sb;
      sm.SegK()[sid] = (uint16_t)(ki | (ks << 8));
    }
    LACB_PH(12);
    LACB_SYNC();
    LACB_PH(13);

    // base (p = 0) mode, block/encoder.cpp:432-484
    const bool allow_zr = cfg.zero_run && best.has_run;
    uint32_t base_mode = MODE_RICE, base_k = best.k_init;
    u64 base_bits = best.rice;
    if (allow_zr && best.zr <= base_bits) { base_bits = best.zr; base_mode = MODE_ZR; }
    if (best.bin < base_bits) { base_bits = best.bin; base_mode = MODE_BIN; }
    if (best.stat < base_bits) { base_bits = best.stat; base_mode = MODE_STATIC; base_k = best.k_stat; }
    u64 best_total = base_bits + 15ull;
    best_total += (8ull - (best_total & 7ull)) & 7ull;
    uint32_t best_p = 0u;
    if (tid == 0u) sm.SelMK()[0] = (uint8_t)((base_mode << 5) | base_k);
    if (!PROBE && tid < 9u) recs[slot].lvl_bits[tid] = tid ? 0u : (uint32_t)best_total;
    LACB_SYNC();

    // partition search, block/encoder.cpp:486-545
    const bool fused = (NT >= 64) && n == (uint32_t)(NT * E) && max_p >= 1u;
    if (fused) {
      // full blocks: every level in one sweep (levels_fused), then every segment's mode choice, then the levels' totals
      constexpr uint32_t STR = ASmem<NT, E>::MAXSEG + 1u;
      levels_fused<NT, E>(sm, pr, n, max_p);
      const u64* Fa = sm.FbAll();
      u64* SelBits = sm.SelBits();
      for (uint32_t sid = 1u + tid; sid < (2u << max_p) - 1u; sid += NT) {
        const u64 rice = split_sum_get(&Fa[sid]), zr = split_sum_get(&Fa[STR + sid]), bin = split_sum_get(&Fa[2u * STR + sid]);
        const bool hr = (mi->hasrun_all[sid >> 5] >> (sid & 31u)) & 1u;
        const uint32_t kk = sm.SegK()[sid];
        const u64 sbits = sm.SegStat()[sid];
        uint32_t mode = MODE_RICE, k = kk & 0xFFu;
        u64 bits = rice;
        if (cfg.zero_run && hr && zr < bits) { mode = MODE_ZR; bits = zr; }
        if (bin < bits) { mode = MODE_BIN; bits = bin; }
        if (sbits < bits || sbits <= bits + bits / 20ull) { mode = MODE_STATIC; k = kk >> 8; bits = sbits; }  // :518, :190-192
        sm.SelMK()[sid] = (uint8_t)((mode << 5) | k);
        SelBits[sid] = bits;
      }
      LACB_SYNC();
      const uint32_t warp = tid >> 5;
      if (warp >= 1u && warp <= max_p) {  // warp p adds up level p
        const uint32_t cnt = 1u << warp;
        u64 part_sum = 0ull;
        for (uint32_t s = tid & 31u; s < cnt; s += 32u) part_sum += SelBits[cnt - 1u + s];
        const u64 sum_bits = warp_sum_u64(part_sum);
        u64 total = sum_bits + 8ull + 7ull * cnt;
        total += (8ull - (total & 7ull)) & 7ull;
        if ((tid & 31u) == 0u) mi->lvl_total[warp] = total;
      }
      LACB_SYNC();
      for (uint32_t p = 1u; p <= max_p; ++p) {
        const u64 total = mi->lvl_total[p];
        if (!PROBE && tid == 0u) recs[slot].lvl_bits[p] = (uint32_t)total;
        const u64 margin = best_total / 20ull;
        if (total < best_total || (total <= best_total + margin && best_p == 0u) ||
            (total == best_total && p < best_p)) {  // :538-540
          best_total = total;
          best_p = p;
        }
      }
      LACB_PH(16);
    } else
    for (uint32_t p = 1u; p <= max_p; ++p) {
      if (GANG) __syncthreads();  // max_p is the same for every probe (256 samples, one configuration)
      const bool sums = cost_pass<NT, E, false>(sm, pr, n, p, 0u) != 0u;  // Fb: per-segment sums or prefixes
      const uint32_t cnt = 1u << p;
      const u64* Fb = sm.Fb();
      u64* SelBits = sm.SelBits();
      for (uint32_t s = tid; s < cnt; s += NT) {
        const uint32_t sid = cnt - 1u + s;
        const u64 rice = sums ? Fb[s] : Fb[s + 1u] - Fb[s];
        const u64 zr = sums ? Fb[ASmem<NT, E>::FBS + s] : Fb[ASmem<NT, E>::FBS + s + 1u] - Fb[ASmem<NT, E>::FBS + s];
        const u64 bin = sums ? Fb[2u * ASmem<NT, E>::FBS + s] : Fb[2u * ASmem<NT, E>::FBS + s + 1u] - Fb[2u * ASmem<NT, E>::FBS + s];
        const bool hr = (mi->hasrun_bits[s >> 5] >> (s & 31u)) & 1u;
        const uint32_t kk = sm.SegK()[sid];
        const u64 sbits = sm.SegStat()[sid];
        uint32_t mode = MODE_RICE, k = kk & 0xFFu;
        u64 bits = rice;
        if (cfg.zero_run && hr && zr < bits) { mode = MODE_ZR; bits = zr; }
        if (bin < bits) { mode = MODE_BIN; bits = bin; }
        if (sbits < bits || sbits <= bits + bits / 20ull) { mode = MODE_STATIC; k = kk >> 8; bits = sbits; }  // :518, :190-192
        sm.SelMK()[sid] = (uint8_t)((mode << 5) | k);
        SelBits[s] = bits;
      }
      LACB_SYNC();
      // every warp adds up the (at most 256) segment costs itself: one barrier instead of a block reduction
      u64 part_sum = 0ull;
      for (uint32_t s = tid & 31u; s < cnt; s += 32u) part_sum += SelBits[s];
      const u64 sum_bits = warp_sum_u64(part_sum);
      u64 total = sum_bits + 8ull + 7ull * cnt;
      total += (8ull - (total & 7ull)) & 7ull;
      if (!PROBE && tid == 0u) recs[slot].lvl_bits[p] = (uint32_t)total;
      const u64 margin = best_total / 20ull;
      if (total < best_total || (total <= best_total + margin && best_p == 0u) ||
          (total == best_total && p < best_p)) {  // :538-540
        best_total = total;
        best_p = p;
      }
      LACB_PH(16);
    }

    // exact emitted size of the chosen configuration (same token function as the emitter)
    const uint32_t chosen_order =
        best.type == PRED_LPC ? (best.taps < max_valid ? (best.taps > 1u ? best.taps : 1u) : (max_valid > 1u ? max_valid : 1u))
                              : best.order;  // block/encoder.cpp:421-423
    const uint32_t nparts = 1u << best_p;
    u64 tok_bits = 0ull;
    if (GANG) __syncthreads();
    if (fused && best_p >= 1u) {
      // full block, partitioned: the register walk of the level sweep, no K plane, no barrier
      chunk_walk_stateless<NT, E, true>(sm, pr, n, best_p,
                                        [&](uint32_t u, uint32_t k, bool is_zero, uint32_t closes, bool long_run, uint32_t m) {
                                          bool emit;
                                          const Token t = make_token(m >> 5, m & 31u, u, k, is_zero, closes, long_run, &emit);
                                          if (emit) tok_bits += (u64)t.hlen + t.q + t.tlen;
                                        });
    } else {
      const SegGeom sg = seg_geom<E>(g0, n, best_p);
      const uint8_t* mk = sm.SelMK();
      const uint32_t mkA = mk[sg.sidA], mkB = (sg.bnd != 0xFFFFFFFFu) ? mk[sg.sidA + 1u] : 0u;
      if (best_p == 0u) k_series<NT, E, true>(sm, pr, n, sg);
      else k_series<NT, E, false>(sm, pr, n, sg, best_p);
      walk_items<NT, E>(sm, pr, n, sg, mkA & 31u, mkB & 31u,
                        [&](int, uint32_t, bool inB, uint32_t u, uint32_t k, bool is_zero, uint32_t closes,
                            bool long_run) {
                          const uint32_t m = inB ? mkB : mkA;
                          bool emit;
                          const Token t = make_token(m >> 5, m & 31u, u, k, is_zero, closes, long_run, &emit);
                          if (emit) tok_bits += (u64)t.hlen + t.q + t.tlen;
                        });
    }
    LACB_PH(17);
    const u64 all_tok = block_sum_u64<NT, E>(sm, tok_bits);
    LACB_PH(18);
    const u64 bits = 16ull + (best.type == PRED_LPC ? 16ull * chosen_order : 0ull) + 8ull + 7ull * nparts + all_tok;
    const u64 bytes = (bits + 7ull) >> 3;

    if (PROBE) {
      if (tid == 0u) probe_bytes[slot] = bytes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bytes;
    } else {
      ChanRec* rec = recs + slot;
      for (uint32_t s = tid; s < 256u; s += NT) rec->part[s] = s < nparts ? sm.SelMK()[nparts - 1u + s] : 0;
      if (tid == 0u) {
        rec->bytes = bytes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bytes;
        rec->bits = bits > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)bits;
        rec->type = (uint8_t)best.type;
        rec->order = (uint8_t)chosen_order;
        rec->p = (uint8_t)best_p;
        rec->taps = (uint8_t)best.taps;
        for (int i = 0; i < 13; ++i) rec->coef[i] = (wcoef && i >= 1 && (uint32_t)i <= chosen_order) ? wcoef[i] : (int16_t)0;
        rec->pad = 0;
        rec->base_mode = (uint8_t)base_mode;
        rec->has_run = (uint8_t)(best.has_run != 0u);
        rec->pad2[0] = rec->pad2[1] = 0;
        rec->est_best = best.best; rec->est_rice = best.rice; rec->est_zr = best.zr; rec->est_bin = best.bin;
        rec->est_stat = best.stat;
      }
    }
    LACB_SYNC();
  }
}

// ---------------------------------------------------------------------------
// Per-block sizes and the final LR/MS pick for "encode both" blocks, then an exclusive
// scan of the block sizes (K11's device-wide prefix sum).  Single CTA: the table has at
// most ~10^5 entries per GPU.
__global__ void __launch_bounds__(1024) k_finalize_blocks(EncCfg cfg, uint32_t* blk_flags, const ChanRec* recs,
                                                          uint32_t* blk_bytes, u64* blk_off, u64* total_bytes,
                                                          uint32_t* err) {
  __shared__ u64 scr[40];
  __shared__ u64 carry;
  const uint32_t tid = threadIdx.x;
  if (tid == 0u) carry = 0ull;
  __syncthreads();
  for (uint32_t base = 0; base < cfg.n_blocks; base += 1024u) {
    const uint32_t b = base + tid;
    u64 sz = 0ull;
    if (b < cfg.n_blocks) {
      uint32_t f = blk_flags[b];
      const ChanRec* rc = recs + (size_t)b * 4u;
      if (cfg.channels == 1u) {
        sz = rc[0].bytes;
      } else {
        if (f & BF_BOTH) {  // lac/encoder.cpp:336-340: MS only if strictly smaller
          const u64 lr = (u64)rc[0].bytes + rc[1].bytes, ms = (u64)rc[2].bytes + rc[3].bytes;
          f = (f & ~BF_CHOOSE_MS) | (ms < lr ? BF_CHOOSE_MS : 0u);
          blk_flags[b] = f;
        }
        const uint32_t s0 = (f & BF_CHOOSE_MS) ? 2u : 0u;
        sz = (u64)rc[s0].bytes + rc[s0 + 1u].bytes + (cfg.stereo_mode == 2u ? 1ull : 0ull);
        if (rc[s0].bytes == 0xFFFFFFFFu || rc[s0 + 1u].bytes == 0xFFFFFFFFu) sz = 0x100000000ull;
      }
      if (sz == 0ull || sz > 0xFFFFFFFFull) {  // lac/encoder.cpp:447-450
        atomicOr(err, 1u);
        sz = 0ull;
      }
      blk_bytes[b] = (uint32_t)sz;
    }
    u64 tot;
    const u64 ex = block_excl_scan_u64<1024>(sz, scr, &tot);
    if (b < cfg.n_blocks) blk_off[b] = carry + ex;
    __syncthreads();
    if (tid == 0u) carry += tot;
    __syncthreads();
  }
  if (tid == 0u) *total_bytes = carry;
}

// ---------------------------------------------------------------------------
// K10: emission.  One CTA per (block, channel); tokens are positioned by a block-wide
// exclusive scan of their bit lengths and OR-ed into a shared-memory window that is
// word-aligned with the destination, then stored with coalesced 32-bit writes.
__device__ __forceinline__ void or_field(uint32_t* stg, uint32_t W, i64 rel, uint32_t value, uint32_t len) {
  if (len == 0u) return;
  if (rel + (i64)len <= 0 || rel >= (i64)W * 32) return;
  const i64 widx = rel >> 5;  // arithmetic shift = floor
  const uint32_t off = (uint32_t)(rel - widx * 32);
  const u64 v = (u64)value << (64u - len - off);
  const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  if (widx >= 0 && widx < (i64)W && hi) atomicOr(&stg[widx], hi);
  if (widx + 1 >= 0 && widx + 1 < (i64)W && lo) atomicOr(&stg[widx + 1], lo);
}
__device__ __forceinline__ void or_ones(uint32_t* stg, uint32_t W, i64 rel, uint32_t q) {
  if (q == 0u) return;
  i64 end = rel + (i64)q;  // exclusive
  if (end <= 0 || rel >= (i64)W * 32) return;
  if (rel < 0) rel = 0;
  if (end > (i64)W * 32) end = (i64)W * 32;
  uint32_t w0 = (uint32_t)(rel >> 5), w1 = (uint32_t)((end - 1) >> 5);
  const uint32_t m0 = 0xFFFFFFFFu >> (uint32_t)(rel & 31), m1 = 0xFFFFFFFFu << (31u - (uint32_t)((end - 1) & 31));
  if (w0 == w1) {
    atomicOr(&stg[w0], m0 & m1);
    return;
  }
  atomicOr(&stg[w0], m0);
  for (uint32_t w = w0 + 1u; w < w1; ++w) atomicOr(&stg[w], 0xFFFFFFFFu);
  atomicOr(&stg[w1], m1);
}

template <int NT, int E>
__global__ void __launch_bounds__(NT) k_emit(PcmSrc src, EncCfg cfg, const uint32_t* blk_flags, const ChanRec* recs,
                                             const u64* blk_off, uint8_t* payload) {
  LACB_DYN_SMEM(unsigned char, smraw);
  ASmem<NT, E> sm{smraw};
  const uint32_t tid = threadIdx.x, g0 = tid * E;
  const uint32_t njobs = cfg.n_blocks * cfg.channels;
  for (uint32_t job = blockIdx.x; job < njobs; job += gridDim.x) {
    const uint32_t b = job / cfg.channels, ch = job - b * cfg.channels;
    const uint32_t f = blk_flags[b];
    const uint32_t s0 = (cfg.channels == 2u && (f & BF_CHOOSE_MS)) ? 2u : 0u;
    const ChanRec* rec = recs + (size_t)b * 4u + s0 + ch;
    const uint32_t n = block_len(src.frames, b);
    const bool flagged = cfg.channels == 2u && cfg.stereo_mode == 2u;
    u64 out_off = blk_off[b] + (flagged ? 1ull : 0ull);
    if (ch == 1u) out_off += recs[(size_t)b * 4u + s0].bytes;
    if (flagged && ch == 0u && tid == 0u) payload[blk_off[b]] = (f & BF_CHOOSE_MS) ? 1 : 0;

    LACB_PH_INIT(-1);
    if (tid == 0u) sm.Misc()->hq_kb_n = 0u;
    load_block<NT, E>(sm, src, (int)(s0 + ch), (u64)b * kMaxBlock, n);
    __syncthreads();
    int32_t x[E + 12];
    load_items<NT, E>(sm, x);
    const uint32_t type = rec->type, order = rec->order, p = rec->p, nparts = 1u << p;
    int32_t r[E];
    compute_residual<NT, E>(x, g0, n, type, order, rec->taps, rec->coef, r);
    Prep<NT, E> pr;
    prepare<NT, E, false, true>(sm, r, n, pr);  // U plane, prefix of u, last-nonzero scan, run flag
    // segment tables of the chosen level
    for (uint32_t s = tid; s < nparts; s += NT) {
      const uint32_t sid = nparts - 1u + s;
      sm.SelMK()[sid] = rec->part[s];
      sm.SegP()[sid] = p ? prefix_u<NT, E>(sm, s * (n >> p)) : 0ull;
    }
    __syncthreads();
    const SegGeom sg = seg_geom<E>(g0, n, p);
    const uint32_t mkA = sm.SelMK()[sg.sidA], mkB = (sg.bnd != 0xFFFFFFFFu) ? sm.SelMK()[sg.sidA + 1u] : 0u;
    if (p == 0u) k_series<NT, E, true>(sm, pr, n, sg);
    else k_series<NT, E, false>(sm, pr, n, sg, p);
    u64 my_bits = 0ull;
    walk_items<NT, E>(sm, pr, n, sg, mkA & 31u, mkB & 31u,
                      [&](int, uint32_t, bool inB, uint32_t u, uint32_t k, bool is_zero, uint32_t closes, bool long_run) {
                        const uint32_t m = inB ? mkB : mkA;
                        bool emit;
                        const Token t = make_token(m >> 5, m & 31u, u, k, is_zero, closes, long_run, &emit);
                        if (emit) my_bits += (u64)t.hlen + t.q + t.tlen;
                      });
    u64 tok_total;
    const u64 my_ex = block_excl_scan_u64<NT>(my_bits, sm.Scr(), &tok_total);
    __syncthreads();
    const uint32_t hdr_fixed = 16u + (type == PRED_LPC ? 16u * order : 0u) + 8u;
    const u64 hdr_bits = (u64)hdr_fixed + 7ull * nparts;
    const u64 total_bits = hdr_bits + tok_total;
    const u64 total_bytes = (total_bits + 7ull) >> 3;  // == rec->bytes by construction

    // windows over the absolute bit range of this channel-block, aligned to destination words
    uint32_t* stg = reinterpret_cast<uint32_t*>(sm.X());
    constexpr uint32_t W = ASmem<NT, E>::CAP;  // words per window
    const u64 abs0 = out_off * 8ull;           // absolute bit address of the first bit
    const u64 absEnd = abs0 + total_bytes * 8ull;
    for (u64 wstart = abs0 & ~31ull; wstart < absEnd; wstart += (u64)W * 32ull) {
      // only the words this window will hand to the copy-out below (a 16384-sample block of 24-bit audio fills
      // less than half of the 64 KB plane)
      const u64 wleft = (absEnd - wstart + 31ull) >> 5;
      const uint32_t wuse = wleft < (u64)W ? (uint32_t)wleft : W;
      for (uint32_t i = tid; i < wuse; i += NT) stg[i] = 0u;
      __syncthreads();
      const i64 rel0 = (i64)(abs0 - wstart);  // staging position of bit 0 of the channel-block (may be negative)
      // header
      if (tid == 0u) {
        or_field(stg, W, rel0, type, 8u);
        or_field(stg, W, rel0 + 8, order, 8u);
        if (type == PRED_LPC)
          for (uint32_t i = 1; i <= order; ++i) or_field(stg, W, rel0 + 16 + 16 * (i64)(i - 1u), (uint16_t)rec->coef[i], 16u);
        const uint32_t m0 = sm.SelMK()[nparts - 1u] >> 5;
        const uint32_t control = ((m0 & 3u) << 5) | (p ? (0x80u | p) : 0u);  // block/encoder.cpp:773-778
        or_field(stg, W, rel0 + (i64)hdr_fixed - 8, control, 8u);
      }
      for (uint32_t s = tid; s < nparts; s += NT) {
        const uint32_t mk = sm.SelMK()[nparts - 1u + s];
        or_field(stg, W, rel0 + (i64)hdr_fixed + 7 * (i64)s, ((mk >> 5) << 5) | (mk & 31u), 7u);
      }
      // tokens
      const i64 t0 = rel0 + (i64)hdr_bits + (i64)my_ex;
      if (my_bits && t0 < (i64)W * 32 && t0 + (i64)my_bits > 0) {
        i64 pos = t0;
        walk_items<NT, E>(sm, pr, n, sg, mkA & 31u, mkB & 31u,
                          [&](int, uint32_t, bool inB, uint32_t u, uint32_t k, bool is_zero, uint32_t closes,
                              bool long_run) {
                            const uint32_t m = inB ? mkB : mkA;
                            bool emit;
                            const Token t = make_token(m >> 5, m & 31u, u, k, is_zero, closes, long_run, &emit);
                            if (!emit) return;
                            const uint32_t tot = t.hlen + t.q + t.tlen;
                            if (tot <= 32u) {  // the usual case: tag, unary run and tail go out as one field
                              const u64 v = ((u64)t.head << (t.q + t.tlen)) | ((((u64)1 << t.q) - 1ull) << t.tlen) | t.tail;
                              or_field(stg, W, pos, (uint32_t)v, tot);
                            } else {
                              or_field(stg, W, pos, t.head, t.hlen);
                              or_ones(stg, W, pos + t.hlen, t.q);
                              or_field(stg, W, pos + t.hlen + t.q, t.tail, t.tlen);
                            }
                            pos += (i64)tot;
                          });
      }
      __syncthreads();
      // copy out: staging word i <-> destination bytes [wstart/8 + 4i, +4)
      const u64 wbyte0 = wstart >> 3;
      const u64 lo = out_off, hi = out_off + total_bytes;
      for (uint32_t i = tid; i < W; i += NT) {
        const u64 b0 = wbyte0 + 4ull * i;
        if (b0 >= hi) break;
        if (b0 + 4ull <= lo) continue;
        const uint32_t v = stg[i];
        if (b0 >= lo && b0 + 4ull <= hi) {
          *reinterpret_cast<uint32_t*>(payload + b0) = __byte_perm(v, 0u, 0x0123);
        } else {
#pragma unroll
          for (uint32_t k = 0; k < 4u; ++k)
            if (b0 + k >= lo && b0 + k < hi) payload[b0 + k] = (uint8_t)(v >> (24u - 8u * k));
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace lacb
