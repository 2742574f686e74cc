// lac_host.cpp -- host facade over the C ABI (see lac_host.hpp).
#include "lac_host.hpp"

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>

#include "../../include/lac_b200.h"

#ifdef LACB_WITH_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

namespace {

constexpr uint32_t kMaxBlock = 16384;
constexpr uint64_t kMaxTotalSamples = 6912000000ull;        // lac/decoder.cpp:17-23
constexpr uint64_t kMaxDecodedPcmBytes = 1ull << 30;
constexpr uint32_t kMaxBlockCount = (uint32_t)((kMaxDecodedPcmBytes / 4 + 255) / 256);
constexpr uint32_t kMinNonFinalBlock = 256;

bool rate_ok(uint32_t r) { return r == 44100 || r == 48000 || r == 96000 || r == 192000; }
bool depth_ok(uint8_t d) { return d == 16 || d == 24; }
bool sample_ok(int32_t v, uint8_t depth) {
  return depth == 16 ? (v >= -32768 && v <= 32767) : (v >= -8388608 && v <= 8388607);
}
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24));
  v.push_back((uint8_t)(x >> 16));
  v.push_back((uint8_t)(x >> 8));
  v.push_back((uint8_t)x);
}

// One lacb_ctx per device, created on first use; a context runs one call at a time.
struct DeviceSlot {
  std::mutex mu;
  lacb_ctx* ctx = nullptr;
};
DeviceSlot g_slots[16];
std::mutex g_slots_mu;

lacb_ctx* ctx_for(int device) {
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

// Global payload offsets from the per-rank byte counts.  With more than one GPU the
// counts are exchanged with an NCCL all-gather over NVLink (one u64 per rank), as the
// sharded design calls for; each rank then knows every slab's offset.
std::vector<uint64_t> gather_counts(const std::vector<uint64_t>& mine_per_rank) {
  const size_t n = mine_per_rank.size();
  std::vector<uint64_t> all = mine_per_rank;
#ifdef LACB_WITH_NCCL
  if (n > 1) {
    std::vector<int> devs(n);
    for (size_t i = 0; i < n; ++i) devs[i] = (int)i;
    std::vector<ncclComm_t> comms(n);
    if (ncclCommInitAll(comms.data(), (int)n, devs.data()) == ncclSuccess) {
      std::vector<uint64_t*> send(n), recv(n);
      std::vector<cudaStream_t> st(n);
      for (size_t i = 0; i < n; ++i) {
        cudaSetDevice((int)i);
        cudaStreamCreate(&st[i]);
        cudaMalloc(&send[i], 8);
        cudaMalloc(&recv[i], 8 * n);
        cudaMemcpyAsync(send[i], &mine_per_rank[i], 8, cudaMemcpyHostToDevice, st[i]);
      }
      ncclGroupStart();
      for (size_t i = 0; i < n; ++i) ncclAllGather(send[i], recv[i], 1, ncclUint64, comms[i], st[i]);
      ncclGroupEnd();
      cudaSetDevice(0);
      cudaMemcpyAsync(all.data(), recv[0], 8 * n, cudaMemcpyDeviceToHost, st[0]);
      for (size_t i = 0; i < n; ++i) {
        cudaSetDevice((int)i);
        cudaStreamSynchronize(st[i]);
        cudaFree(send[i]);
        cudaFree(recv[i]);
        cudaStreamDestroy(st[i]);
        ncclCommDestroy(comms[i]);
      }
    }
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

void check_decode_limits(const ParsedFrame& pf, bool planes) {
  uint64_t total_bytes = 0;
  for (uint32_t b : pf.bytes) total_bytes += b;
  if (planes && pf.frames * pf.hdr.channels * 4ull > kMaxDecodedPcmBytes)
    throw_decode_error("decoded PCM allocation exceeds maximum");
  const uint64_t wav = pf.frames * pf.hdr.channels * (pf.hdr.bit_depth / 8u);
  if (36u + wav + (wav & 1u) > 0xFFFFFFFFull) throw_decode_error("decoded WAV data exceeds RIFF limit");
  if (pf.hdr.version >= 3 && total_bytes != pf.payload_bytes)
    throw_decode_error("compressed block sizes do not match frame payload");
}

void run_decode(const ParsedFrame& pf, int layout, void* out_a, void* out_b, LAC::ThreadCollector* collector) {
  lacb_dec_params prm{pf.hdr.bit_depth, pf.hdr.channels, pf.hdr.stereo_mode};
  lacb_err err{};
  lacb_ctx* ctx = ctx_for(0);
  std::lock_guard<std::mutex> lock(g_slots[0].mu);
  if (collector) collector->record(std::this_thread::get_id());
  // v2 streams have no per-block byte sizes: NULL selects the serial walk
  const int rc = lacb_decode(ctx, &prm, pf.payload, pf.payload_bytes, pf.sizes.data(),
                             pf.bytes.empty() ? nullptr : pf.bytes.data(),
                             (uint32_t)pf.sizes.size(), layout, out_a, out_b, &err);
  if (rc == LACB_EDECODE) throw std::runtime_error(err.msg);
  if (rc != 0) throw std::runtime_error(std::string("LAC B200 backend: ") + lacb_last_error(ctx));
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
}  // namespace lacb_host

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

Encoder::Encoder(uint8_t order, uint8_t stereo_mode, uint32_t sample_rate, uint8_t bit_depth, bool, bool, bool)
    : order_(order), stereo_mode_(stereo_mode), sample_rate_(sample_rate), bit_depth_(bit_depth) {}

std::vector<uint8_t> Encoder::encode(const std::vector<int32_t>& left, const std::vector<int32_t>& right,
                                     ThreadCollector* collector) {
  if (left.empty()) throw std::invalid_argument("left channel must not be empty");
  if (!right.empty() && right.size() != left.size())
    throw std::invalid_argument("right channel size (" + std::to_string(right.size()) +
                                ") must match left channel size (" + std::to_string(left.size()) + ")");
  if (!rate_ok(sample_rate_)) throw std::invalid_argument("unsupported sample rate: " + std::to_string(sample_rate_));
  if (!depth_ok(bit_depth_)) throw std::invalid_argument("unsupported bit depth: " + std::to_string(bit_depth_));
  if (stereo_mode_ > 2) throw std::invalid_argument("unsupported stereo mode: " + std::to_string(stereo_mode_));
  return run(LACB_PLANAR_I32, left.data(), right.empty() ? nullptr : right.data(), left.size(),
             right.empty() ? 1 : 2, collector);
}

std::vector<uint8_t> Encoder::encode_packed(const uint8_t* pcm, uint64_t frames, uint8_t channels,
                                            ThreadCollector* collector) {
  if (!pcm || frames == 0) throw std::invalid_argument("left channel must not be empty");
  if (channels != 1 && channels != 2) throw std::invalid_argument("unsupported channel count");
  if (!rate_ok(sample_rate_)) throw std::invalid_argument("unsupported sample rate: " + std::to_string(sample_rate_));
  if (!depth_ok(bit_depth_)) throw std::invalid_argument("unsupported bit depth: " + std::to_string(bit_depth_));
  if (stereo_mode_ > 2) throw std::invalid_argument("unsupported stereo mode: " + std::to_string(stereo_mode_));
  return run(LACB_PACKED_LE, pcm, nullptr, frames, channels, collector);
}

std::vector<uint8_t> Encoder::run(int layout, const void* a, const void* b, uint64_t frames, uint8_t channels,
                                  ThreadCollector* collector) {
  (void)order_;
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  const size_t devices = std::min<size_t>(lacb_host::resolve_devices(device_count_), nb);
  const std::vector<Shard> shards = plan_shards(frames, devices);
  const size_t bps = bit_depth_ / 8u;

  lacb_enc_params prm{};
  prm.sample_rate = sample_rate_;
  prm.bit_depth = bit_depth_;
  prm.channels = channels;
  prm.stereo_mode = channels == 2 ? stereo_mode_ : 0;
  prm.zero_run_enabled = zero_run_enabled_;
  prm.partitioning_enabled = partitioning_enabled_;
  prm.validate_range = 1;

  std::vector<uint32_t> block_bytes(nb);
  std::vector<uint8_t*> slabs(shards.size(), nullptr);
  std::vector<uint64_t> slab_bytes(shards.size(), 0);
  std::vector<int> rcs(shards.size(), 0);
  std::vector<std::string> msgs(shards.size());
  auto work = [&](size_t d) {
    try {
      lacb_ctx* ctx = ctx_for((int)d);
      std::lock_guard<std::mutex> lock(g_slots[d].mu);
      if (collector) collector->record(std::this_thread::get_id());
      const Shard& s = shards[d];
      const void* pa;
      const void* pb = nullptr;
      if (layout == LACB_PLANAR_I32) {
        pa = static_cast<const int32_t*>(a) + s.first_frame;
        if (b) pb = static_cast<const int32_t*>(b) + s.first_frame;
      } else {
        pa = static_cast<const uint8_t*>(a) + s.first_frame * channels * bps;
      }
      lacb_err err{};
      rcs[d] = lacb_encode(ctx, &prm, layout, pa, pb, s.frames, &slabs[d], &slab_bytes[d],
                           block_bytes.data() + s.first_block, &err);
      if (rcs[d] != 0) msgs[d] = lacb_last_error(ctx);
    } catch (const std::exception& e) {
      rcs[d] = LACB_ECUDA;
      msgs[d] = e.what();
    }
  };
  if (shards.size() == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (size_t d = 0; d < shards.size(); ++d) th.emplace_back(work, d);
    for (auto& t : th) t.join();
  }
  for (size_t d = 0; d < shards.size(); ++d) {
    if (rcs[d] == 0) continue;
    for (uint8_t* p : slabs) lacb_free(p);
    if (rcs[d] == LACB_EINVAL && layout == LACB_PLANAR_I32) {
      // reproduce the reference's message: first offending sample, left channel first
      const int32_t* l = static_cast<const int32_t*>(a);
      const int32_t* r = static_cast<const int32_t*>(b);
      for (uint64_t i = 0; i < frames; ++i)
        if (!sample_ok(l[i], bit_depth_))
          throw std::invalid_argument("left sample at index " + std::to_string(i) +
                                      " is outside the configured PCM bit depth");
      for (uint64_t i = 0; r && i < frames; ++i)
        if (!sample_ok(r[i], bit_depth_))
          throw std::invalid_argument("right sample at index " + std::to_string(i) +
                                      " is outside the configured PCM bit depth");
    }
    if (rcs[d] == LACB_ELIMIT) throw std::runtime_error("encoded block size is outside format limits");
    throw std::runtime_error("LAC B200 backend: " + msgs[d]);
  }

  // global offsets of the per-GPU slabs, then header + table + slabs in rank order
  const std::vector<uint64_t> counts = gather_counts(slab_bytes);
  uint64_t payload_total = 0;
  for (uint64_t c : counts) payload_total += c;
  std::vector<uint8_t> out;
  out.reserve(FrameHeader::kBytes + 4 + 8ull * nb + payload_total);
  FrameHeader hdr;
  hdr.channels = channels;
  hdr.stereo_mode = (uint8_t)prm.stereo_mode;
  hdr.sample_rate = sample_rate_;
  hdr.bit_depth = bit_depth_;
  hdr.append_to(out);
  put_be32(out, nb);
  for (uint32_t i = 0; i < nb; ++i) {
    const uint64_t start = (uint64_t)i * kMaxBlock;
    put_be32(out, (uint32_t)std::min<uint64_t>(kMaxBlock, frames - start));
    put_be32(out, block_bytes[i]);
  }
  for (size_t d = 0; d < shards.size(); ++d) {
    out.insert(out.end(), slabs[d], slabs[d] + counts[d]);
    lacb_free(slabs[d]);
  }
  return out;
}

void Decoder::decode(const uint8_t* data, size_t size, std::vector<int32_t>& left, std::vector<int32_t>& right,
                     FrameHeader* out_header) {
  left.clear();
  right.clear();
  (void)thread_count_;
  const ParsedFrame pf = parse_frame(data, size);
  check_decode_limits(pf, true);
  std::vector<int32_t> l(pf.frames), r(pf.hdr.channels == 2 ? pf.frames : 0);
  run_decode(pf, LACB_PLANAR_I32, l.data(), r.empty() ? nullptr : r.data(), collector_);
  left.swap(l);
  right.swap(r);
  if (out_header) *out_header = pf.hdr;
}

void Decoder::decode_packed(const uint8_t* data, size_t size, std::vector<uint8_t>& out, FrameHeader& hdr,
                            uint64_t& frames) {
  out.clear();
  const ParsedFrame pf = parse_frame(data, size);
  check_decode_limits(pf, false);
  std::vector<uint8_t> pcm(pf.frames * pf.hdr.channels * (pf.hdr.bit_depth / 8u));
  run_decode(pf, LACB_PACKED_LE, pcm.data(), nullptr, collector_);
  out.swap(pcm);
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
  if (br.has_error() || (br.bit_position() & 7u) != 0) return false;
  const size_t byte = br.bit_position() >> 3;
  lacb_ctx* ctx = ctx_for(0);
  std::lock_guard<std::mutex> lock(g_slots[0].mu);
  uint64_t bits = 0;
  const int rc = lacb_decode_block(ctx, br.data() + byte, br.size_bytes() - byte, block_size, out, &bits);
  if (rc != 1) {
    br.mark_error();
    return false;
  }
  br.advance_bits((size_t)bits);
  return true;
}

}  // namespace Block
