// lac_host.cpp -- host facade over the C ABI (see lac_host.hpp).
#include "lac_host.hpp"

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>

#include "../../include/lac_b200.h"
#include "wav_io.hpp"

#ifdef LACB_WITH_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

namespace {

constexpr uint32_t kMaxBlock = 16384;
constexpr uint64_t kMaxTotalSamples = 6912000000ull;        // lac/decoder.cpp:17-23
constexpr uint32_t kMaxBlockCount = (uint32_t)((kMaxDecodedPcmBytes / 4 + 255) / 256);
constexpr uint32_t kMinNonFinalBlock = 256;

bool rate_ok(uint32_t r) { return r == 44100 || r == 48000 || r == 96000 || r == 192000; }
bool depth_ok(uint8_t d) { return d == 16 || d == 24; }
bool sample_ok(int32_t v, uint8_t depth) {
  return depth == 16 ? (v >= -32768 && v <= 32767) : (v >= -8388608 && v <= 8388607);
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// One lacb_ctx per device, created on first use; a context runs one call at a time.
constexpr size_t kMaxDevices = 16;
struct DeviceSlot {
  std::mutex mu;
  lacb_ctx* ctx = nullptr;
};
DeviceSlot g_slots[kMaxDevices];
std::mutex g_slots_mu;

lacb_ctx* ctx_for(int device) {
  if (device < 0 || (size_t)device >= kMaxDevices)
    throw std::runtime_error("LAC B200 backend: device index " + std::to_string(device) + " out of range");
  std::lock_guard<std::mutex> lock(g_slots_mu);
  DeviceSlot& s = g_slots[device];
  if (!s.ctx) {
    const int rc = lacb_create(device, &s.ctx);
    if (rc != 0 || !s.ctx)
      throw std::runtime_error("LAC B200 backend: no usable CUDA device " + std::to_string(device) +
                               " (there is no CPU fallback)");
  }
  return s.ctx;
}

struct Shard {
  uint64_t first_frame = 0, frames = 0;
  uint32_t first_block = 0, blocks = 0;
};
std::vector<Shard> plan_shards(uint64_t frames, size_t devices) {
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  const uint32_t per = (uint32_t)((nb + devices - 1) / devices);
  std::vector<Shard> out;
  for (size_t d = 0; d < devices; ++d) {
    const uint32_t b0 = std::min<uint32_t>(nb, (uint32_t)d * per), b1 = std::min<uint32_t>(nb, (uint32_t)(d + 1) * per);
    if (b1 == b0) break;
    Shard s;
    s.first_block = b0;
    s.blocks = b1 - b0;
    s.first_frame = (uint64_t)b0 * kMaxBlock;
    s.frames = std::min<uint64_t>(frames, (uint64_t)b1 * kMaxBlock) - s.first_frame;
    out.push_back(s);
  }
  return out;
}

// How a call spreads over the box.  `threads` is the reference's worker cap (--threads / LAC_THREADS /
// set_thread_count, src/codec/lac/encoder.cpp:385-390, lac/decoder.cpp:236-242).  On the GPU path a worker is
// a host thread driving one device, and inside a device a slice of blocks in flight on its own stream, so the
// cap bounds devices first and then the slices per device: --threads=1 is one device, one stream, nothing
// overlapped; 0 = automatic.  The bytes never depend on it.
struct WorkerPlan {
  size_t devices = 1;
  uint32_t streams = 0;  // per device, 0 = automatic
};
WorkerPlan plan_workers(size_t requested_devices, size_t threads, uint64_t units) {
  WorkerPlan w;
  w.devices = std::min<size_t>(lacb_host::resolve_devices(requested_devices), kMaxDevices);
  if (units < w.devices) w.devices = (size_t)std::max<uint64_t>(1, units);
  if (threads > 0) {
    w.devices = std::min(w.devices, threads);
    w.streams = (uint32_t)std::min<size_t>(16, std::max<size_t>(1, threads / w.devices));
  }
  return w;
}

// Runs fn(d) for d in [0, n): inline for one shard, one host thread per device otherwise.  Exceptions are
// carried back to the caller (lowest shard wins).
template <typename F>
void for_each_device(size_t n, F&& fn) {
  if (n == 1) {
    fn((size_t)0);
    return;
  }
  std::vector<std::exception_ptr> errs(n);
  std::vector<std::thread> th;
  for (size_t d = 0; d < n; ++d)
    th.emplace_back([&, d] {
      try {
        fn(d);
      } catch (...) {
        errs[d] = std::current_exception();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (e) std::rethrow_exception(e);
}

// Global payload offsets from the per-device byte counts.  With more than one GPU the counts are exchanged with
// an NCCL all-gather over NVLink (one u64 per rank), as the sharded design calls for; the
// communicators, streams and buffers are created once per process and reused.  Every rank's gathered vector is
// checked against the host's own copy: any NCCL / CUDA failure or disagreement throws instead of silently
// falling through.
#ifdef LACB_WITH_NCCL
struct NcclRing {
  std::mutex mu;
  size_t n = 0;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> st;
  std::vector<uint64_t*> send, recv;
  void destroy() {
    for (size_t i = 0; i < n; ++i) {
      cudaSetDevice((int)i);
      if (send[i]) cudaFree(send[i]);
      if (recv[i]) cudaFree(recv[i]);
      if (st[i]) cudaStreamDestroy(st[i]);
      if (comms[i]) ncclCommDestroy(comms[i]);
    }
    n = 0;
    comms.clear();
    st.clear();
    send.clear();
    recv.clear();
  }
};
NcclRing g_nccl;  // lives until process exit (no teardown at static destruction: the CUDA runtime may be gone)
void nccl_ok(ncclResult_t r, const char* what) {
  if (r != ncclSuccess)
    throw std::runtime_error(std::string("LAC B200 backend: NCCL ") + what + ": " + ncclGetErrorString(r));
}
void cuda_ok(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("LAC B200 backend: ") + what + ": " + cudaGetErrorString(e));
}
#endif

std::vector<uint64_t> gather_counts(const std::vector<uint64_t>& mine_per_rank) {
  const size_t n = mine_per_rank.size();
  std::vector<uint64_t> all = mine_per_rank;
#ifdef LACB_WITH_NCCL
  if (n > 1) {
    std::lock_guard<std::mutex> lock(g_nccl.mu);
    if (g_nccl.n != n) {
      g_nccl.destroy();
      std::vector<int> devs(n);
      for (size_t i = 0; i < n; ++i) devs[i] = (int)i;
      g_nccl.comms.assign(n, nullptr);
      g_nccl.st.assign(n, nullptr);
      g_nccl.send.assign(n, nullptr);
      g_nccl.recv.assign(n, nullptr);
      g_nccl.n = n;
      nccl_ok(ncclCommInitAll(g_nccl.comms.data(), (int)n, devs.data()), "ncclCommInitAll");
      for (size_t i = 0; i < n; ++i) {
        cuda_ok(cudaSetDevice((int)i), "cudaSetDevice");
        cuda_ok(cudaStreamCreateWithFlags(&g_nccl.st[i], cudaStreamNonBlocking), "cudaStreamCreate");
        cuda_ok(cudaMalloc(&g_nccl.send[i], 8), "cudaMalloc");
        cuda_ok(cudaMalloc(&g_nccl.recv[i], 8 * n), "cudaMalloc");
      }
    }
    for (size_t i = 0; i < n; ++i) {
      cuda_ok(cudaSetDevice((int)i), "cudaSetDevice");
      cuda_ok(cudaMemcpyAsync(g_nccl.send[i], &mine_per_rank[i], 8, cudaMemcpyHostToDevice, g_nccl.st[i]),
              "cudaMemcpyAsync");
    }
    nccl_ok(ncclGroupStart(), "ncclGroupStart");
    for (size_t i = 0; i < n; ++i)
      nccl_ok(ncclAllGather(g_nccl.send[i], g_nccl.recv[i], 1, ncclUint64, g_nccl.comms[i], g_nccl.st[i]),
              "ncclAllGather");
    nccl_ok(ncclGroupEnd(), "ncclGroupEnd");
    std::vector<uint64_t> got(n * n);
    for (size_t i = 0; i < n; ++i) {
      cuda_ok(cudaSetDevice((int)i), "cudaSetDevice");
      cuda_ok(cudaMemcpyAsync(got.data() + i * n, g_nccl.recv[i], 8 * n, cudaMemcpyDeviceToHost, g_nccl.st[i]),
              "cudaMemcpyAsync");
    }
    for (size_t i = 0; i < n; ++i) {
      cuda_ok(cudaSetDevice((int)i), "cudaSetDevice");
      cuda_ok(cudaStreamSynchronize(g_nccl.st[i]), "cudaStreamSynchronize");
    }
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < n; ++j)
        if (got[i * n + j] != mine_per_rank[j])
          throw std::runtime_error(
              "LAC B200 backend: NCCL all-gather of the payload byte counts disagrees with the host");
    all.assign(got.begin(), got.begin() + n);
  }
#endif
  return all;
}

[[noreturn]] void throw_decode_error(const std::string& reason) { throw std::runtime_error("[decode-error] " + reason); }

struct ParsedFrame {
  FrameHeader hdr;
  std::vector<uint32_t> sizes, bytes;
  const uint8_t* payload = nullptr;
  uint64_t payload_bytes = 0, frames = 0;
};

// header + block table validation, lac/decoder.cpp:88-159 (same checks, same order, same text)
ParsedFrame parse_frame(const uint8_t* data, size_t size) {
  ParsedFrame pf;
  if (data == nullptr || size == 0) throw_decode_error("empty input");
  size_t hb = 0;
  if (!FrameHeader::parse(data, size, pf.hdr, hb)) throw_decode_error("invalid frame header");
  const bool v3 = pf.hdr.version >= 3;
  const uint8_t* body = data + hb;
  const uint64_t body_bytes = size - hb;
  if (body_bytes < 4) throw_decode_error("invalid block count");
  const uint32_t nb = be32(body);
  if (nb == 0 || nb > kMaxBlockCount) throw_decode_error("invalid block count");
  const uint32_t words = v3 ? 2u : 1u;  // v2 tables carry sample counts only
  if (nb > ((body_bytes - 4) * 8u) / (32u * words)) throw_decode_error("truncated block size table");
  pf.sizes.resize(nb);
  pf.bytes.resize(v3 ? nb : 0);
  uint64_t total = 0, total_bytes = 0;
  for (uint32_t i = 0; i < nb; ++i) {
    const uint32_t s = be32(body + 4 + 4ull * words * i);
    if (s == 0 || s > kMaxBlock || (i + 1 < nb && s < kMinNonFinalBlock)) throw_decode_error("invalid block size");
    total += s;
    if (total > kMaxTotalSamples) throw_decode_error("total samples exceed maximum");
    pf.sizes[i] = s;
    if (v3) {
      const uint32_t b = be32(body + 8 + 8ull * i);
      if (b == 0) throw_decode_error("invalid compressed block size");
      total_bytes += b;
      if (total_bytes > body_bytes) throw_decode_error("compressed block sizes exceed frame payload");
      pf.bytes[i] = b;
    }
  }
  pf.frames = total;
  pf.payload = body + 4 + 4ull * words * nb;
  pf.payload_bytes = body_bytes - 4 - 4ull * words * nb;
  return pf;
}

// Size limits of the reference decoders (lac/decoder.cpp:139-159 and the CLI fast path, src/main.cpp:247-262):
// both cap the decoded PCM at 1 GiB of int32 planes and the WAV at the classic RIFF size.  `allow_large` is the
// explicit opt-in (--allow-large, Decoder::set_allow_large) that lifts the two caps so that BASELINE configs 3
// and 4 can exist as files (the output is then RF64 where RIFF cannot hold it); MAX_TOTAL_SAMPLES
// and MAX_BLOCK_COUNT stay in force either way.
void check_decode_limits(const ParsedFrame& pf, bool allow_large) {
  uint64_t total_bytes = 0;
  for (uint32_t b : pf.bytes) total_bytes += b;
  if (!allow_large && pf.frames * pf.hdr.channels * 4ull > kMaxDecodedPcmBytes)
    throw_decode_error("decoded PCM allocation exceeds maximum");
  const uint64_t wav = pf.frames * pf.hdr.channels * (pf.hdr.bit_depth / 8u);
  if (!allow_large && 36u + wav + (wav & 1u) > 0xFFFFFFFFull) throw_decode_error("decoded WAV data exceeds RIFF limit");
  if (pf.hdr.version >= 3 && total_bytes != pf.payload_bytes)
    throw_decode_error("compressed block sizes do not match frame payload");
}

// reference message of a device verdict (lac/decoder.cpp:25-32,183-190,272-274) with the GLOBAL block index
std::string decode_error_text(const lacb_err& err, uint32_t first_block) {
  const uint32_t b = first_block + err.block_index;
  switch (err.reason) {
    case 1: return "[decode-error] invalid per-block stereo flag";
    case 2: return "[decode-error] block=" + std::to_string(b) + " channel=primary";
    case 3: return "[decode-error] block=" + std::to_string(b) + " channel=secondary";
    case 4: return "[decode-error] decoded sample outside PCM bit depth";
    case 5: return "[decode-error] block=" + std::to_string(b) + " channel=trailing-payload";
    default: return err.msg;
  }
}

// The decode_block pool of LAC::Decoder::decode (lac/decoder.cpp:236-291): the block table is cut into
// contiguous ranges, one per device, each decoded by its own host thread into its part of the output.  The
// first failing block in block order is the one reported, as in the reference.  v2 streams are one serial
// chain and stay on one device.
void run_decode(const ParsedFrame& pf, int layout, void* out_a, void* out_b, LAC::ThreadCollector* collector,
                size_t requested_devices, size_t threads) {
  lacb_dec_params prm{pf.hdr.bit_depth, pf.hdr.channels, pf.hdr.stereo_mode};
  const uint32_t nb = (uint32_t)pf.sizes.size();
  const bool v3 = !pf.bytes.empty();
  // fewer than ~2 waves of parser warps per device are latency bound: more devices would not help
  const WorkerPlan plan = plan_workers(v3 ? requested_devices : 1, threads, std::max<uint32_t>(1u, nb / 64u));
  const uint32_t per = (uint32_t)((nb + plan.devices - 1) / plan.devices);
  struct Part {
    uint32_t b0, b1;
    uint64_t frame0, byte0, bytes;
  };
  std::vector<Part> parts;
  {
    uint64_t f = 0, by = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += per) {
      Part p{b0, std::min(nb, b0 + per), f, by, 0};
      for (uint32_t b = p.b0; b < p.b1; ++b) {
        f += pf.sizes[b];
        if (v3) p.bytes += pf.bytes[b];
      }
      by += p.bytes;
      parts.push_back(p);
    }
  }
  const size_t fb = (size_t)pf.hdr.channels * (pf.hdr.bit_depth / 8u);
  std::vector<int> rcs(parts.size(), 0);
  std::vector<lacb_err> errs(parts.size());
  std::vector<std::string> msgs(parts.size());
  for_each_device(parts.size(), [&](size_t d) {
    const Part& p = parts[d];
    lacb_ctx* ctx = ctx_for((int)d);
    std::lock_guard<std::mutex> lock(g_slots[d].mu);
    if (collector) collector->record(std::this_thread::get_id());
    lacb_set_concurrency(ctx, plan.streams);
    void* a = layout == LACB_PLANAR_I32 ? (void*)(static_cast<int32_t*>(out_a) + p.frame0)
                                        : (void*)(static_cast<uint8_t*>(out_a) + p.frame0 * fb);
    void* b = (layout == LACB_PLANAR_I32 && out_b) ? (void*)(static_cast<int32_t*>(out_b) + p.frame0) : nullptr;
    errs[d] = lacb_err{};
    // v2 streams have no per-block byte sizes: NULL selects the serial walk over the whole payload
    rcs[d] = lacb_decode(ctx, &prm, pf.payload + p.byte0, v3 ? p.bytes : pf.payload_bytes, pf.sizes.data() + p.b0,
                         v3 ? pf.bytes.data() + p.b0 : nullptr, p.b1 - p.b0, layout, a, b, &errs[d]);
    if (rcs[d] != 0) msgs[d] = lacb_last_error(ctx);
  });
  for (size_t d = 0; d < parts.size(); ++d) {
    if (rcs[d] == 0) continue;
    if (rcs[d] == LACB_EDECODE) throw std::runtime_error(decode_error_text(errs[d], parts[d].b0));
    throw std::runtime_error("LAC B200 backend: " + msgs[d]);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
void FrameHeader::append_to(std::vector<uint8_t>& out) const {
  out.push_back((uint8_t)(sync >> 8));
  out.push_back((uint8_t)sync);
  out.push_back(version);
  out.push_back(channels);
  out.push_back(stereo_mode);
  out.push_back((uint8_t)(sample_rate >> 8));   // low 16 bits first, then the high byte
  out.push_back((uint8_t)sample_rate);
  out.push_back((uint8_t)(sample_rate >> 16));
  out.push_back(bit_depth);
  out.push_back(reserved);
}
bool FrameHeader::valid() const {
  return sync == 0x4C41 && (version == 2 || version == 3) && (channels == 1 || channels == 2) &&
         !(channels == 1 && stereo_mode != 0) && stereo_mode <= 2 && rate_ok(sample_rate) && depth_ok(bit_depth) &&
         reserved == 0;
}
bool FrameHeader::parse(const uint8_t* d, size_t size, FrameHeader& out, size_t& header_bytes) {
  if (!d || size < kBytes) return false;
  FrameHeader h;
  h.sync = (uint16_t)((d[0] << 8) | d[1]);
  h.version = d[2];
  h.channels = d[3];
  h.stereo_mode = d[4];
  h.sample_rate = (uint32_t)((d[5] << 8) | d[6]) | ((uint32_t)d[7] << 16);
  h.bit_depth = d[8];
  h.reserved = d[9];
  if (!h.valid()) return false;
  out = h;
  header_bytes = kBytes;
  return true;
}

void FrameHeader::write(BitWriter& w) const {
  std::vector<uint8_t> b;
  append_to(b);
  for (uint8_t x : b) w.write_bits(x, 8);
}
void FrameHeader::read(BitReader& r) {
  sync = (uint16_t)r.read_bits(16);
  version = (uint8_t)r.read_bits(8);
  channels = (uint8_t)r.read_bits(8);
  stereo_mode = (uint8_t)r.read_bits(8);
  const uint32_t lo = r.read_bits(16), hi = r.read_bits(8);
  bit_depth = (uint8_t)r.read_bits(8);
  reserved = (uint8_t)r.read_bits(8);
  sample_rate = lo | (hi << 16);
}

uint32_t BitReader::read_bits(int nbits) {
  if (nbits <= 0) return 0;
  if (error_ || pos_ >= size_ * 8 || (size_t)nbits > size_ * 8 - pos_) {
    mark_error();
    return 0;
  }
  uint32_t v = 0;
  for (int i = 0; i < nbits; ++i, ++pos_) {
    const uint32_t bit = (data_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1u;
    if (i < 32) v = (v << 1) | bit;
  }
  return v;
}
bool BitReader::read_unary_ones(uint32_t max_ones, uint32_t& ones) {
  // ones up to a 0 terminator; more than max_ones, or running out of data, fails
  uint32_t n = 0;
  ones = 0;
  while (!error_ && pos_ < size_ * 8) {
    const uint32_t bit = (data_[pos_ >> 3] >> (7 - (pos_ & 7))) & 1u;
    if (!bit) {
      ++pos_;
      ones = n;
      return true;
    }
    if (n == max_ones) {
      ones = n;
      return false;
    }
    ++n;
    ++pos_;
  }
  ones = n;
  mark_error();
  return false;
}
void BitReader::align_to_byte() {
  if (!error_) pos_ = (pos_ + 7) & ~(size_t)7;
  if (pos_ > size_ * 8) mark_error();
}
bool BitReader::consume_zero_padding_to_byte() {
  while (pos_ & 7) {
    if (read_bits(1) != 0u || error_) return false;
  }
  return true;
}
void BitReader::advance_bits(size_t n) {
  if (error_ || n > size_ * 8 - pos_) mark_error();
  else pos_ += n;
}

void BitWriter::write_bit(uint32_t bit) {
  cur_ = (uint8_t)((cur_ << 1) | (bit & 1u));
  if (++nbits_ == 8) {
    buffer_.push_back(cur_);
    cur_ = 0;
    nbits_ = 0;
  }
}
void BitWriter::write_bits(uint32_t value, int nbits) {
  for (int i = nbits - 1; i >= 0; --i) write_bit(i >= 32 ? 0u : (value >> i) & 1u);
}
void BitWriter::write_unary_ones(uint32_t ones) {
  for (uint32_t i = 0; i < ones; ++i) write_bit(1u);
}
void BitWriter::write_bytes(const uint8_t* data, size_t size) {
  for (size_t i = 0; i < size; ++i) write_bits(data[i], 8);
}
void BitWriter::flush_to_byte() {
  while (nbits_ != 0) write_bit(0u);
}
std::vector<uint8_t> BitWriter::take_buffer() {
  std::vector<uint8_t> out;  // complete bytes only, like get_buffer(); callers flush first
  out.swap(buffer_);
  return out;
}

// ---------------------------------------------------------------------------
namespace lacb_host {
lacb_ctx* shared_context(int device) { return ctx_for(device); }
std::mutex& shared_context_mutex(int device) { return g_slots[device].mu; }
int device_count() { return lacb_device_count(); }
size_t resolve_devices(size_t requested) {
  if (requested == 0) {
    const char* env = std::getenv("LAC_DEVICES");
    if (env && *env) requested = LAC::parse_thread_limit(env);
  }
  if (requested == 0) requested = 1;
  const int have = device_count();
  if (have <= 0) throw std::runtime_error("LAC B200 backend: no CUDA device visible (there is no CPU fallback)");
  return std::min<size_t>(requested, (size_t)have);
}
void warm_up(size_t requested) noexcept {
  try {
    const size_t n = std::min<size_t>(resolve_devices(requested), kMaxDevices);
    for (size_t d = 0; d < n; ++d) ctx_for((int)d);
  } catch (...) {
  }
}
}  // namespace lacb_host

namespace {

// Page-locking a freshly created output mapping costs more than it saves in a one-shot run (the kernel has to
// fault in and pin every page first: measured ~1 s per GB on the bench box, against ~0.1 s for the unpinned copy),
// so the file paths register their mappings only on request (LAC_PIN_FILES=1: long-lived processes that reuse them).
bool pin_files() {
  static const bool on = std::getenv("LAC_PIN_FILES") != nullptr && std::getenv("LAC_PIN_FILES")[0] == '1';
  return on;
}

// helper threads for host-side page work: the caller's --threads cap, else the machine's cores, at most 16
unsigned host_helpers(size_t thread_cap) {
  unsigned n = std::thread::hardware_concurrency();
  if (n == 0) n = 1;
  if (thread_cap && thread_cap < n) n = (unsigned)thread_cap;
  return std::min(n, 16u);
}

lacb_enc_params make_enc_params(uint32_t rate, uint8_t depth, uint8_t channels, uint8_t stereo_mode, bool zr, bool part) {
  lacb_enc_params prm{};
  prm.sample_rate = rate;
  prm.bit_depth = depth;
  prm.channels = channels;
  prm.stereo_mode = channels == 2 ? stereo_mode : 0;  // lac/encoder.cpp:247
  prm.zero_run_enabled = zr;
  prm.partitioning_enabled = part;
  prm.validate_range = 1;
  return prm;
}

// header + block count + table (lac/encoder.cpp:243-252,445-459) into dst (14 + 8 * nb bytes)
void write_frame_head(uint8_t* dst, const FrameHeader& hdr, uint64_t frames, const std::vector<uint32_t>& block_bytes) {
  std::vector<uint8_t> h;
  hdr.append_to(h);
  std::memcpy(dst, h.data(), FrameHeader::kBytes);
  const uint32_t nb = (uint32_t)block_bytes.size();
  auto put = [](uint8_t* p, uint32_t x) {
    p[0] = (uint8_t)(x >> 24);
    p[1] = (uint8_t)(x >> 16);
    p[2] = (uint8_t)(x >> 8);
    p[3] = (uint8_t)x;
  };
  put(dst + FrameHeader::kBytes, nb);
  uint8_t* t = dst + FrameHeader::kBytes + 4;
  for (uint32_t i = 0; i < nb; ++i, t += 8) {
    put(t, (uint32_t)std::min<uint64_t>(kMaxBlock, frames - (uint64_t)i * kMaxBlock));
    put(t + 4, block_bytes[i]);
  }
}

[[noreturn]] void throw_encode_failure(int rc, const std::string& msg, int layout, const void* a, const void* b,
                                       uint64_t frames, uint8_t bit_depth) {
  if (rc == LACB_EINVAL && layout == LACB_PLANAR_I32) {
    // reproduce the reference's message: first offending sample, left channel first (lac/encoder.cpp:232-241)
    const int32_t* l = static_cast<const int32_t*>(a);
    const int32_t* r = static_cast<const int32_t*>(b);
    for (uint64_t i = 0; i < frames; ++i)
      if (!sample_ok(l[i], bit_depth))
        throw std::invalid_argument("left sample at index " + std::to_string(i) +
                                    " is outside the configured PCM bit depth");
    for (uint64_t i = 0; r && i < frames; ++i)
      if (!sample_ok(r[i], bit_depth))
        throw std::invalid_argument("right sample at index " + std::to_string(i) +
                                    " is outside the configured PCM bit depth");
  }
  if (rc == LACB_ELIMIT) throw std::runtime_error("encoded block size is outside format limits");
  throw std::runtime_error("LAC B200 backend: " + msg);
}

// Several GPUs: every device encodes its block range with the PCM and the payload resident in its HBM
// (phase 1), the payload byte counts are all-gathered over NCCL to give every slab its global offset, and
// each device then copies its slab straight to that offset of the destination (phase 2) -- a vector, or the
// mapped output file -- so the host never concatenates anything (lac/encoder.cpp:445-465 by DMA).
// `place(total_payload_bytes)` returns where payload byte 0 goes, once the total is known.
template <typename Place>
uint64_t encode_sharded(const lacb_enc_params& prm, int layout, const void* a, const void* b, uint64_t frames,
                        const std::vector<Shard>& shards, uint32_t streams, LAC::ThreadCollector* collector,
                        std::vector<uint32_t>& block_bytes, Place&& place) {
  const size_t n = shards.size();
  const size_t bps = prm.bit_depth / 8u;
  struct Dev {
    void* d_a = nullptr;
    void* d_b = nullptr;
    const uint8_t* d_payload = nullptr;
    uint64_t bytes = 0;
    int rc = 0;
    std::string msg;
  };
  std::vector<Dev> dev(n);
  auto release = [&] {
    for (size_t d = 0; d < n; ++d) {
      lacb_ctx* ctx = ctx_for((int)d);
      if (dev[d].d_a) lacb_dev_free(ctx, dev[d].d_a);
      if (dev[d].d_b) lacb_dev_free(ctx, dev[d].d_b);
    }
  };
  for_each_device(n, [&](size_t d) {
    lacb_ctx* ctx = ctx_for((int)d);
    std::lock_guard<std::mutex> lock(g_slots[d].mu);
    if (collector) collector->record(std::this_thread::get_id());
    lacb_set_concurrency(ctx, streams);
    const Shard& s = shards[d];
    Dev& v = dev[d];
    auto fail = [&](int rc) {
      v.rc = rc;
      v.msg = lacb_last_error(ctx);
    };
    int rc;
    if (layout == LACB_PLANAR_I32) {
      if ((rc = lacb_dev_malloc(ctx, s.frames * 4, &v.d_a)) != 0) return fail(rc);
      if ((rc = lacb_memcpy_h2d(ctx, v.d_a, static_cast<const int32_t*>(a) + s.first_frame, s.frames * 4)) != 0) return fail(rc);
      if (b) {
        if ((rc = lacb_dev_malloc(ctx, s.frames * 4, &v.d_b)) != 0) return fail(rc);
        if ((rc = lacb_memcpy_h2d(ctx, v.d_b, static_cast<const int32_t*>(b) + s.first_frame, s.frames * 4)) != 0) return fail(rc);
      }
    } else {
      const uint64_t fbytes = (uint64_t)prm.channels * bps;
      if ((rc = lacb_dev_malloc(ctx, s.frames * fbytes + 4, &v.d_a)) != 0) return fail(rc);
      if ((rc = lacb_memcpy_h2d(ctx, v.d_a, static_cast<const uint8_t*>(a) + s.first_frame * fbytes, s.frames * fbytes)) != 0)
        return fail(rc);
    }
    lacb_err err{};
    const uint32_t* d_bb = nullptr;
    if ((rc = lacb_encode_device(ctx, &prm, layout, v.d_a, v.d_b, s.frames, &v.d_payload, &v.bytes, &d_bb, &err)) != 0)
      return fail(rc);
    if ((rc = lacb_memcpy_d2h(ctx, block_bytes.data() + s.first_block, d_bb, (uint64_t)s.blocks * 4)) != 0) return fail(rc);
  });
  for (size_t d = 0; d < n; ++d)
    if (dev[d].rc != 0) {
      release();
      throw_encode_failure(dev[d].rc, dev[d].msg, layout, a, b, frames, (uint8_t)prm.bit_depth);
    }
  std::vector<uint64_t> mine(n);
  for (size_t d = 0; d < n; ++d) mine[d] = dev[d].bytes;
  std::vector<uint64_t> counts;
  try {
    counts = gather_counts(mine);
  } catch (...) {
    release();
    throw;
  }
  std::vector<uint64_t> off(n + 1, 0);
  for (size_t d = 0; d < n; ++d) off[d + 1] = off[d] + counts[d];
  uint8_t* dst = place(off[n]);
  for_each_device(n, [&](size_t d) {
    lacb_ctx* ctx = ctx_for((int)d);
    std::lock_guard<std::mutex> lock(g_slots[d].mu);
    dev[d].rc = lacb_memcpy_d2h(ctx, dst + off[d], dev[d].d_payload, dev[d].bytes);
    if (dev[d].rc != 0) dev[d].msg = lacb_last_error(ctx);
  });
  release();
  for (size_t d = 0; d < n; ++d)
    if (dev[d].rc != 0) throw std::runtime_error("LAC B200 backend: " + dev[d].msg);
  return off[n];
}

}  // namespace

namespace LAC {

size_t parse_thread_limit(const char* value) {
  if (value == nullptr || value[0] == '\0') return 0;
  for (const char* p = value; *p; ++p)
    if (*p < '0' || *p > '9') throw std::invalid_argument("LAC_THREADS must be a positive integer");
  errno = 0;
  char* end = nullptr;
  const unsigned long long v = std::strtoull(value, &end, 10);
  if (errno != 0 || end == value || *end != '\0' || v == 0)
    throw std::invalid_argument("LAC_THREADS must be a positive integer");
  if (v > (unsigned long long)std::numeric_limits<size_t>::max()) throw std::invalid_argument("LAC_THREADS is too large");
  return (size_t)v;
}

Encoder::Encoder(uint8_t order, uint8_t stereo_mode, uint32_t sample_rate, uint8_t bit_depth, bool debug_lpc,
                 bool debug_stereo_est, bool debug_zr)
    : order_(order), stereo_mode_(stereo_mode), sample_rate_(sample_rate), bit_depth_(bit_depth),
      debug_lpc_(debug_lpc), debug_stereo_est_(debug_stereo_est), debug_zr_(debug_zr) {}

// The decision log of a whole file in the words of the reference's Debug build (LAC_DEBUG_LOG is compiled out of
// its Release build): [stereo-est] / [stereo-mode] (src/codec/lac/encoder.cpp:356-379), [zr-est], [part-est],
// [part-choose] and [debug-lpc] (src/codec/block/encoder.cpp:457-466, 527-551, 824-835), one block after the other
// and only for the channel-blocks that were emitted (the reference also logs its probe and both-pair encodes, in
// worker order).  `energy` (the Levinson error of an LPC winner) is not kept on the device and is left out.
void Encoder::print_decisions(lacb_ctx* ctx, uint8_t mode) const {
  uint32_t nb = 0;
  if (lacb_last_encode_decisions(ctx, nullptr, 0, &nb) != 0)
    throw std::runtime_error(std::string("LAC B200 backend: ") + lacb_last_error(ctx));
  std::vector<lacb_block_decision> dec(nb);
  if (lacb_last_encode_decisions(ctx, dec.data(), nb, &nb) != 0)
    throw std::runtime_error(std::string("LAC B200 backend: ") + lacb_last_error(ctx));
  std::ostringstream out;
  for (uint32_t b = 0; b < nb; ++b) {
    const lacb_block_decision& d = dec[b];
    const bool stereo = d.ch[1].bytes != 0;
    const bool ms = (d.flags & LACB_DEC_MS) != 0;
    for (int c = 0; c < (stereo ? 2 : 1); ++c) {
      const lacb_chan_decision& ch = d.ch[c];
      if (debug_zr_ && zero_run_enabled_)
        out << "[zr-est] block=" << b << " normal=" << ch.rice_bits << " zr=" << ch.zr_bits << " bin=" << ch.bin_bits
            << " static=" << ch.static_bits << " chosen=" << (int)ch.base_mode << " has_run=" << (int)ch.has_run << "\n";
      if (debug_partitions_ && partitioning_enabled_ && d.block_size >= 64) {
        for (uint32_t p = 1; p < 9; ++p)
          if (ch.level_bits[p])
            out << "[part-est] block=" << b << " p=" << p << " bits=" << ch.level_bits[p] << " partitions=" << (1u << p)
                << "\n";
        out << "[part-choose] block=" << b << " best_p=" << (uint32_t)ch.partition_order
            << " bits=" << ch.level_bits[ch.partition_order] << "\n";
      }
      if (debug_lpc_)
        out << "[debug-lpc] block=" << d.block_size << " chosen_order=" << (int)ch.order
            << " predictor=" << (int)ch.predictor_type << " est_bits=" << ch.est_bits << " rice_bits=" << ch.rice_bits
            << " zr_bits=" << ch.zr_bits << " bin_bits=" << ch.bin_bits << " part_order=" << (uint32_t)ch.partition_order
            << "\n";
    }
    if (debug_stereo_est_ && stereo) {
      if (mode == 2)
        out << "[stereo-est] block=" << b << " uncertain=" << ((d.flags & LACB_DEC_UNCERTAIN) ? 1 : 0)
            << " chosen=" << (ms ? "MS" : "LR") << "\n";
      out << "[stereo-mode] global=" << (int)mode << " block=" << b << " mode_used=" << (ms ? "MS" : "LR") << "\n";
    }
  }
  std::cerr << out.str();
}

void Encoder::check_config() const {
  if (!rate_ok(sample_rate_)) throw std::invalid_argument("unsupported sample rate: " + std::to_string(sample_rate_));
  if (!depth_ok(bit_depth_)) throw std::invalid_argument("unsupported bit depth: " + std::to_string(bit_depth_));
  if (stereo_mode_ > 2) throw std::invalid_argument("unsupported stereo mode: " + std::to_string(stereo_mode_));
}

std::vector<uint8_t> Encoder::encode(const std::vector<int32_t>& left, const std::vector<int32_t>& right,
                                     ThreadCollector* collector) {
  if (left.empty()) throw std::invalid_argument("left channel must not be empty");
  if (!right.empty() && right.size() != left.size())
    throw std::invalid_argument("right channel size (" + std::to_string(right.size()) +
                                ") must match left channel size (" + std::to_string(left.size()) + ")");
  check_config();
  return run(LACB_PLANAR_I32, left.data(), right.empty() ? nullptr : right.data(), left.size(),
             right.empty() ? 1 : 2, collector);
}

std::vector<uint8_t> Encoder::encode_packed(const uint8_t* pcm, uint64_t frames, uint8_t channels,
                                            ThreadCollector* collector) {
  if (!pcm || frames == 0) throw std::invalid_argument("left channel must not be empty");
  if (channels != 1 && channels != 2) throw std::invalid_argument("unsupported channel count");
  check_config();
  return run(LACB_PACKED_LE, pcm, nullptr, frames, channels, collector);
}

std::vector<uint8_t> Encoder::run(int layout, const void* a, const void* b, uint64_t frames, uint8_t channels,
                                  ThreadCollector* collector) {
  (void)order_;
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  WorkerPlan plan = plan_workers(device_count_, thread_count_, nb);
  if (debug_any()) {  // the decision log is read from one context's workspace: one device, one pass
    plan.devices = 1;
    plan.streams = 1;
  }
  const std::vector<Shard> shards = plan_shards(frames, plan.devices);
  const lacb_enc_params prm = make_enc_params(sample_rate_, bit_depth_, channels, stereo_mode_, zero_run_enabled_,
                                              partitioning_enabled_);
  FrameHeader hdr;
  hdr.channels = channels;
  hdr.stereo_mode = (uint8_t)prm.stereo_mode;
  hdr.sample_rate = sample_rate_;
  hdr.bit_depth = bit_depth_;
  const size_t head = FrameHeader::kBytes + 4 + 8ull * nb;
  std::vector<uint32_t> block_bytes(nb);
  std::vector<uint8_t> out;

  if (shards.size() == 1) {
    lacb_ctx* ctx = ctx_for(0);
    std::lock_guard<std::mutex> lock(g_slots[0].mu);
    if (collector) collector->record(std::this_thread::get_id());
    lacb_set_concurrency(ctx, plan.streams);
    uint8_t* slab = nullptr;
    uint64_t slab_bytes = 0;
    lacb_err err{};
    const int rc = lacb_encode(ctx, &prm, layout, a, b, frames, &slab, &slab_bytes, block_bytes.data(), &err);
    if (rc != 0) throw_encode_failure(rc, lacb_last_error(ctx), layout, a, b, frames, bit_depth_);
    if (debug_any()) print_decisions(ctx, (uint8_t)prm.stereo_mode);
    out.resize(head);
    out.reserve(head + slab_bytes);
    out.insert(out.end(), slab, slab + slab_bytes);
    lacb_free(slab);
  } else {
    encode_sharded(prm, layout, a, b, frames, shards, plan.streams, collector, block_bytes, [&](uint64_t total) {
      out.resize(head + total);
      return out.data() + head;
    });
  }
  write_frame_head(out.data(), hdr, frames, block_bytes);
  return out;
}

// The CLI's encode path: packed samples in (typically the mapped data chunk of the input WAV), the .lac written
// through a mapping of the output file -- page-locked when the kernel allows it -- so that payload bytes travel
// device -> file with no intermediate vector.  Returns the .lac size.
uint64_t Encoder::encode_packed_to_file(const uint8_t* pcm, uint64_t frames, uint8_t channels, const std::string& path,
                                        ThreadCollector* collector) {
  if (!pcm || frames == 0) throw std::invalid_argument("left channel must not be empty");
  if (channels != 1 && channels != 2) throw std::invalid_argument("unsupported channel count");
  check_config();
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  WorkerPlan plan = plan_workers(device_count_, thread_count_, nb);
  if (debug_any()) {
    plan.devices = 1;
    plan.streams = 1;
  }
  const std::vector<Shard> shards = plan_shards(frames, plan.devices);
  const lacb_enc_params prm = make_enc_params(sample_rate_, bit_depth_, channels, stereo_mode_, zero_run_enabled_,
                                              partitioning_enabled_);
  FrameHeader hdr;
  hdr.channels = channels;
  hdr.stereo_mode = (uint8_t)prm.stereo_mode;
  hdr.sample_rate = sample_rate_;
  hdr.bit_depth = bit_depth_;
  const uint64_t head = FrameHeader::kBytes + 4 + 8ull * nb;
  const uint64_t pcm_bytes = frames * channels * (bit_depth_ / 8u);
  // room for the payload: the PCM size plus the worst-case per-block overhead; an input that expands beyond
  // that (never seen: the static mode bounds a block near its raw size) falls back to the vector path
  const uint64_t cap = pcm_bytes + pcm_bytes / 8 + (uint64_t)nb * 64 + 4096;
  MappedFile mf;
  if (!mf.create(path, head + cap)) throw std::runtime_error("failed to create LAC output");
  // the pages of the new file are faulted in on helper threads while CUDA starts up (ctx_for waits for a
  // warm-up in flight); the payload rarely needs more than 3/4 of the PCM size, the rest faults on demand
  mf.populate_async(head + pcm_bytes - pcm_bytes / 4, host_helpers(thread_count_));
  lacb_ctx* ctx0 = ctx_for(0);
  mf.wait_populated();
  const bool pinned = pin_files() && lacb_host_register(ctx0, mf.data, mf.size) == 0;
  auto unpin = [&] {
    if (pinned) lacb_host_unregister(ctx0, mf.data);
  };
  std::vector<uint32_t> block_bytes(nb);
  uint64_t total = 0;
  try {
    if (shards.size() == 1) {
      std::lock_guard<std::mutex> lock(g_slots[0].mu);
      if (collector) collector->record(std::this_thread::get_id());
      lacb_set_concurrency(ctx0, plan.streams);
      lacb_err err{};
      const int rc = lacb_encode_to(ctx0, &prm, LACB_PACKED_LE, pcm, nullptr, frames, mf.data + head, cap, &total,
                                    block_bytes.data(), &err);
      if (rc == LACB_ENOMEM) {  // payload larger than the mapping: take the vector path
        unpin();
        mf.close();
        const std::vector<uint8_t> v = run(LACB_PACKED_LE, pcm, nullptr, frames, channels, collector);
        MappedFile out;
        if (!out.create(path, v.size())) throw std::runtime_error("failed to create LAC output");
        std::memcpy(out.data, v.data(), v.size());
        return v.size();
      }
      if (rc != 0) throw_encode_failure(rc, lacb_last_error(ctx0), LACB_PACKED_LE, pcm, nullptr, frames, bit_depth_);
      if (debug_any()) print_decisions(ctx0, (uint8_t)prm.stereo_mode);
    } else {
      total = encode_sharded(prm, LACB_PACKED_LE, pcm, nullptr, frames, shards, plan.streams, collector, block_bytes,
                             [&](uint64_t t) -> uint8_t* {
                               if (t > cap) throw std::runtime_error("LAC B200 backend: payload exceeds the output mapping");
                               return mf.data + head;
                             });
    }
  } catch (...) {
    unpin();
    throw;
  }
  write_frame_head(mf.data, hdr, frames, block_bytes);
  unpin();
  if (!mf.resize(head + total)) throw std::runtime_error("failed to size LAC output");
  return head + total;
}

void Decoder::decode(const uint8_t* data, size_t size, std::vector<int32_t>& left, std::vector<int32_t>& right,
                     FrameHeader* out_header) {
  left.clear();
  right.clear();
  const ParsedFrame pf = parse_frame(data, size);
  check_decode_limits(pf, allow_large_);
  std::vector<int32_t> l(pf.frames), r(pf.hdr.channels == 2 ? pf.frames : 0);
  run_decode(pf, LACB_PLANAR_I32, l.data(), r.empty() ? nullptr : r.data(), collector_, device_count_, thread_count_);
  left.swap(l);
  right.swap(r);
  if (out_header) *out_header = pf.hdr;
}

void Decoder::decode_packed(const uint8_t* data, size_t size, std::vector<uint8_t>& out, FrameHeader& hdr,
                            uint64_t& frames) {
  out.clear();
  const ParsedFrame pf = parse_frame(data, size);
  check_decode_limits(pf, allow_large_);
  std::vector<uint8_t> pcm(pf.frames * pf.hdr.channels * (pf.hdr.bit_depth / 8u));
  run_decode(pf, LACB_PACKED_LE, pcm.data(), nullptr, collector_, device_count_, thread_count_);
  out.swap(pcm);
  hdr = pf.hdr;
  frames = pf.frames;
}

// The CLI's decode fast path (src/main.cpp:184-430): the WAV is created at its final size and mapped, the
// header is written into the mapping and the device packs the samples straight into `mapped + header`
// (page-locked when the kernel allows it; otherwise the same pointer unregistered).  RF64 when the caller
// opted in to large files and classic RIFF cannot hold the data.
void Decoder::decode_packed_to_file(const uint8_t* data, size_t size, const std::string& path, FrameHeader& hdr,
                                    uint64_t& frames) {
  const ParsedFrame pf = parse_frame(data, size);
  check_decode_limits(pf, allow_large_);
  WavInfo info;
  info.channels = pf.hdr.channels;
  info.sample_rate = pf.hdr.sample_rate;
  info.bit_depth = pf.hdr.bit_depth;
  info.frames = pf.frames;
  const uint64_t pcm_bytes = pf.frames * pf.hdr.channels * (pf.hdr.bit_depth / 8u);
  const std::vector<uint8_t> head = wav_header(info, pcm_bytes);
  MappedFile mf;
  if (!mf.create(path, head.size() + pcm_bytes + (pcm_bytes & 1u))) throw std::runtime_error("failed to create WAV output");
  std::memcpy(mf.data, head.data(), head.size());
  mf.populate_async(mf.size, host_helpers(thread_count_));  // under the CUDA start-up, see encode_packed_to_file
  lacb_ctx* ctx0 = ctx_for(0);
  mf.wait_populated();
  const bool pinned = pin_files() && lacb_host_register(ctx0, mf.data, mf.size) == 0;
  try {
    run_decode(pf, LACB_PACKED_LE, mf.data + head.size(), nullptr, collector_, device_count_, thread_count_);
  } catch (...) {
    if (pinned) lacb_host_unregister(ctx0, mf.data);
    throw;
  }
  if (pinned) lacb_host_unregister(ctx0, mf.data);
  hdr = pf.hdr;
  frames = pf.frames;
}

}  // namespace LAC

namespace Block {

Encoder::Encoder(int order, bool, bool) : order_(order) {}

std::vector<uint8_t> Encoder::encode(const std::vector<int32_t>& pcm) {
  (void)order_;
  if (pcm.empty() || pcm.size() > kMaxBlock) throw std::invalid_argument("block size must be in [1, 16384]");
  lacb_ctx* ctx = ctx_for(0);
  std::lock_guard<std::mutex> lock(g_slots[0].mu);
  uint8_t* out = nullptr;
  uint64_t n = 0;
  const int rc = lacb_encode_block(ctx, pcm.data(), (uint32_t)pcm.size(), zero_run_enabled_, partitioning_enabled_,
                                   &out, &n);
  if (rc != 0) throw std::runtime_error(std::string("LAC B200 backend: ") + lacb_last_error(ctx));
  std::vector<uint8_t> v(out, out + n);
  lacb_free(out);
  return v;
}

bool Decoder::decode(BitReader& br, uint32_t block_size, std::vector<int32_t>& out) {
  if (block_size == 0 || block_size > kMaxBlock) return false;
  std::vector<int32_t> pcm(block_size);
  if (!decode_into(br, block_size, pcm.data())) return false;  // `out` untouched on failure
  out.swap(pcm);
  return true;
}

bool Decoder::decode_into(BitReader& br, uint32_t block_size, int32_t* out) {
  if (block_size == 0 || block_size > kMaxBlock || out == nullptr) return false;
  if (br.has_error()) return false;
  // the device reader starts at any bit position, like the reference's (block/decoder.cpp:64)
  const size_t byte = br.bit_position() >> 3, bit = br.bit_position() & 7u;
  lacb_ctx* ctx = ctx_for(0);
  std::lock_guard<std::mutex> lock(g_slots[0].mu);
  uint64_t bits = 0;
  int ran_out = 0;
  const int rc = lacb_decode_block_at(ctx, br.data() + byte, br.size_bytes() - byte, bit, block_size, out, &bits, &ran_out);
  if (rc != 1) {
    // a semantic reject leaves the reader usable; only running out of data puts it into its error state
    // (bitstream/bit_reader.hpp:40-60)
    if (rc < 0 || ran_out) br.mark_error();
    return false;
  }
  br.advance_bits((size_t)bits);
  return true;
}

}  // namespace Block
