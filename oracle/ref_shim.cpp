// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" wrapper over the *unmodified* reference classes so pytest /
// bench.py's cpu_baseline leg can call the real CPU codec through ctypes.
// It is compiled together with the reference sources where they lie under
// /root/reference (see oracle/Makefile); outputs go to oracle/_ref/ only.
//
// Wrapped reference entry points:
//   LAC::Encoder::encode            src/codec/lac/encoder.cpp:215
//   LAC::Decoder::decode            src/codec/lac/decoder.cpp:76
//   Block::Encoder::encode          src/codec/block/encoder.cpp:313
//   Block::Decoder::decode_into     src/codec/block/decoder.cpp:64
//   LPC::analyze_block_q15          src/codec/lpc/lpc.cpp:156
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "codec/bitstream/bit_reader.hpp"
#include "codec/block/decoder.hpp"
#include "codec/block/encoder.hpp"
#include "codec/lac/decoder.hpp"
#include "codec/lac/encoder.hpp"
#include "codec/lpc/lpc.hpp"

namespace {
thread_local std::string g_last_error;

uint8_t* dup_bytes(const std::vector<uint8_t>& v) {
  uint8_t* p = static_cast<uint8_t*>(std::malloc(v.empty() ? 1 : v.size()));
  if (p && !v.empty()) std::memcpy(p, v.data(), v.size());
  return p;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }
void ref_free(void* p) { std::free(p); }

// Returns 0 on success; *out is malloc'ed (ref_free). right may be NULL (mono).
int ref_encode(const int32_t* left, const int32_t* right, uint64_t frames,
               uint32_t sample_rate, uint32_t bit_depth, uint32_t stereo_mode,
               int zero_run, int partitioning, uint32_t threads,
               uint8_t** out, uint64_t* out_size) {
  try {
    std::vector<int32_t> l(left, left + frames);
    std::vector<int32_t> r;
    if (right) r.assign(right, right + frames);
    LAC::Encoder enc(12, static_cast<uint8_t>(stereo_mode), sample_rate,
                     static_cast<uint8_t>(bit_depth));
    enc.set_zero_run_enabled(zero_run != 0);
    enc.set_partitioning_enabled(partitioning != 0);
    enc.set_thread_count(threads);
    std::vector<uint8_t> bytes = enc.encode(l, r);
    *out = dup_bytes(bytes);
    *out_size = bytes.size();
    return *out ? 0 : -2;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

// Timed variants for bench.py's reference arm / cpu_baseline leg: the std::vector copies of the
// caller's planes (an artefact of calling the C++ interface through C) are made BEFORE the clock
// starts, so *seconds covers LAC::Encoder::encode / LAC::Decoder::decode and nothing else.
int ref_encode_timed(const int32_t* left, const int32_t* right, uint64_t frames,
                     uint32_t sample_rate, uint32_t bit_depth, uint32_t stereo_mode,
                     uint32_t threads, uint8_t** out, uint64_t* out_size, double* seconds) {
  try {
    std::vector<int32_t> l(left, left + frames);
    std::vector<int32_t> r;
    if (right) r.assign(right, right + frames);
    LAC::Encoder enc(12, static_cast<uint8_t>(stereo_mode), sample_rate,
                     static_cast<uint8_t>(bit_depth));
    enc.set_thread_count(threads);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> bytes = enc.encode(l, r);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    *out = dup_bytes(bytes);
    *out_size = bytes.size();
    return *out ? 0 : -2;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

// Decodes `data`, timing LAC::Decoder::decode only, and compares the result with the expected
// planes (*match = 1 when identical; expect_r NULL for mono).
int ref_decode_timed(const uint8_t* data, uint64_t size, uint32_t threads, const int32_t* expect_l,
                     const int32_t* expect_r, uint64_t frames, double* seconds, int* match) {
  try {
    LAC::Decoder dec;
    dec.set_thread_count(threads);
    std::vector<int32_t> l, r;
    const auto t0 = std::chrono::steady_clock::now();
    dec.decode(data, size, l, r, nullptr);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    bool ok = l.size() == frames && std::memcmp(l.data(), expect_l, sizeof(int32_t) * frames) == 0;
    if (expect_r) ok = ok && r.size() == frames && std::memcmp(r.data(), expect_r, sizeof(int32_t) * frames) == 0;
    else ok = ok && r.empty();
    *match = ok ? 1 : 0;
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

// Decodes a whole .lac stream. left/right are malloc'ed int32 planes.
int ref_decode(const uint8_t* data, uint64_t size, uint32_t threads,
               int32_t** left, int32_t** right, uint64_t* frames,
               uint32_t* channels, uint32_t* sample_rate, uint32_t* bit_depth,
               uint32_t* stereo_mode) {
  try {
    LAC::Decoder dec;
    dec.set_thread_count(threads);
    std::vector<int32_t> l, r;
    FrameHeader hdr;
    dec.decode(data, size, l, r, &hdr);
    *frames = l.size();
    *channels = hdr.channels;
    *sample_rate = hdr.sample_rate;
    *bit_depth = hdr.bit_depth;
    *stereo_mode = hdr.stereo_mode;
    *left = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * (l.size() + 1)));
    std::memcpy(*left, l.data(), sizeof(int32_t) * l.size());
    *right = nullptr;
    if (!r.empty()) {
      *right = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * r.size()));
      std::memcpy(*right, r.data(), sizeof(int32_t) * r.size());
    }
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

int ref_block_encode(const int32_t* pcm, uint32_t n, int zero_run, int partitioning,
                     uint8_t** out, uint64_t* out_size) {
  try {
    std::vector<int32_t> v(pcm, pcm + n);
    Block::Encoder enc(12);
    enc.set_zero_run_enabled(zero_run != 0);
    enc.set_partitioning_enabled(partitioning != 0);
    std::vector<uint8_t> bytes = enc.encode(v);
    *out = dup_bytes(bytes);
    *out_size = bytes.size();
    return *out ? 0 : -2;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

// Returns 1 on success (and bits consumed), 0 when the reference rejects.
int ref_block_decode(const uint8_t* data, uint64_t size, uint32_t block_size,
                     int32_t* out, uint64_t* bits_consumed) {
  BitReader br(data, static_cast<size_t>(size));
  Block::Decoder dec;
  const size_t before = br.bits_remaining();
  const bool ok = dec.decode_into(br, block_size, out);
  if (bits_consumed) *bits_consumed = ok ? before - br.bits_remaining() : 0;
  return ok ? 1 : 0;
}

// coeffs_out must hold order+1 int16. Returns used_order (0 => unstable).
int ref_lpc_analyze(const int32_t* pcm, uint32_t n, int order, int16_t* coeffs_out) {
  std::vector<int32_t> v(pcm, pcm + n);
  LPC lpc(order);
  std::vector<int16_t> c;
  int used = 0;
  lpc.analyze_block_q15(v, c, used);
  for (int i = 0; i <= order; ++i) coeffs_out[i] = c[static_cast<size_t>(i)];
  return used;
}

}  // extern "C"
