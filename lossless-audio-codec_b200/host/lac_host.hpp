// lac_host.hpp -- C++20 host facade over liblac_b200.so (include/lac_b200.h).
//
// Same class names, constructor/method signatures, exception types and messages as the
// reference's public codec interface, so code written against it (its CLI, its tests)
// relinks against the GPU path unchanged:
//   LAC::Encoder          src/codec/lac/encoder.hpp:12-43
//   LAC::Decoder          src/codec/lac/decoder.hpp:10-24
//   LAC::ThreadCollector  src/codec/lac/thread_collector.hpp:8-23
//   LAC::parse_thread_limit  src/codec/lac/thread_limit.hpp:10-28
//   Block::Encoder        src/codec/block/encoder.hpp:9-30
//   Block::Decoder        src/codec/block/decoder.hpp:9-15
//   FrameHeader           src/codec/frame/frame_header.hpp:10-77
// What stays on the host, as in the reference: argument validation, the 10-byte frame
// header, the block table, payload concatenation, table validation / limits on decode,
// and block-range scheduling across GPUs.  Everything per block runs on the device.
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

class BitReader;
class BitWriter;

struct FrameHeader {
  uint16_t sync = 0x4C41;
  uint8_t version = 3;
  uint8_t channels = 2;      // defaults as in frame/frame_header.hpp:19-27
  uint8_t stereo_mode = 2;
  uint32_t sample_rate = 44100;
  uint8_t bit_depth = 16;
  uint8_t reserved = 0;

  static constexpr size_t kBytes = 10;
  void append_to(std::vector<uint8_t>& out) const;
  void write(BitWriter& w) const;
  void read(BitReader& r);
  bool validate() const { return valid(); }
  bool valid() const;
  // parses and validates; header_bytes receives 10 on success
  static bool parse(const uint8_t* data, size_t size, FrameHeader& out, size_t& header_bytes);
  static bool parse(const uint8_t* data, size_t size, FrameHeader& out) {
    size_t n = 0;
    return parse(data, size, out, n);
  }
};

// MSB-first bit reader / writer with the reference's interface
// (bitstream/bit_reader.hpp:6-30, bit_writer.hpp:6-23).  Host-side utilities: the GPU
// decoder consumes whole byte ranges, these exist so code that hand-builds or inspects
// streams (the reference's unit tests) keeps compiling.
class BitReader {
 public:
  BitReader(const uint8_t* data, size_t size) : data_(data), size_(size) {}
  BitReader(const std::vector<uint8_t>& buf) : data_(buf.data()), size_(buf.size()) {}
  uint32_t read_bit() { return read_bits(1); }
  uint32_t read_bits(int nbits);
  bool read_unary_ones(uint32_t max_ones, uint32_t& ones);
  void align_to_byte();
  bool consume_zero_padding_to_byte();
  size_t bits_remaining() const { return error_ ? 0 : size_ * 8 - pos_; }
  bool has_error() const { return error_; }
  bool eof() const { return bits_remaining() == 0; }
  const uint8_t* data() const { return data_; }
  size_t size_bytes() const { return size_; }
  size_t bit_position() const { return pos_; }
  void advance_bits(size_t n);
  void mark_error() { error_ = true; pos_ = size_ * 8; }

 private:
  const uint8_t* data_;
  size_t size_;
  size_t pos_ = 0;
  bool error_ = false;
};

class BitWriter {
 public:
  BitWriter() = default;
  void write_bit(uint32_t bit);
  void write_bits(uint32_t value, int nbits);
  void write_unary_ones(uint32_t ones);
  void write_bytes(const uint8_t* data, size_t size);
  void reserve_bytes(size_t size) { buffer_.reserve(size); }
  void flush_to_byte();
  const std::vector<uint8_t>& get_buffer() const { return buffer_; }
  std::vector<uint8_t> take_buffer();

 private:
  std::vector<uint8_t> buffer_;  // complete bytes
  uint8_t cur_ = 0;              // partial byte, MSB first
  int nbits_ = 0;                // bits held in cur_
};

struct lacb_ctx;

namespace LAC {

class ThreadCollector {
 public:
  void record(std::thread::id id) {
    std::lock_guard<std::mutex> lock(mutex_);
    ids_.insert(id);
  }
  std::set<std::thread::id> snapshot() const {
    std::lock_guard<std::mutex> lock(mutex_);
    return ids_;
  }

 private:
  mutable std::mutex mutex_;
  std::set<std::thread::id> ids_;
};

// "LAC_THREADS must be a positive integer" semantics of the reference
size_t parse_thread_limit(const char* value);

class Encoder {
 public:
  Encoder(uint8_t order, uint8_t stereo_mode = 0, uint32_t sample_rate = 44100, uint8_t bit_depth = 16,
          bool debug_lpc = false, bool debug_stereo_est = false, bool debug_zr = false);

  std::vector<uint8_t> encode(const std::vector<int32_t>& left, const std::vector<int32_t>& right,
                              ThreadCollector* collector = nullptr);
  // GPU-path extension: packed little-endian interleaved samples straight from a WAV data
  // chunk (no host de-interleave)
  std::vector<uint8_t> encode_packed(const uint8_t* pcm, uint64_t frames, uint8_t channels,
                                     ThreadCollector* collector = nullptr);

  // GPU-path extension, the CLI's encode path: the .lac is written through a mapping of `path` (payload bytes go
  // device -> file, no intermediate vector).  Returns the .lac size.
  uint64_t encode_packed_to_file(const uint8_t* pcm, uint64_t frames, uint8_t channels, const std::string& path,
                                 ThreadCollector* collector = nullptr);

  void set_zero_run_enabled(bool enabled) { zero_run_enabled_ = enabled; }
  void set_partitioning_enabled(bool enabled) { partitioning_enabled_ = enabled; }
  void set_debug_partitions(bool enabled) { debug_partitions_ = enabled; }
  // worker cap of the reference (0 = automatic): bounds the devices used and the slices in flight per device
  void set_thread_count(size_t max_threads) { thread_count_ = max_threads; }
  void set_device_count(size_t devices) { device_count_ = devices; }  // 0 = LAC_DEVICES or 1

 private:
  void check_config() const;
  std::vector<uint8_t> run(int layout, const void* a, const void* b, uint64_t frames, uint8_t channels,
                           ThreadCollector* collector);
  uint8_t order_;  // accepted and ignored, exactly like the reference
  uint8_t stereo_mode_;
  uint32_t sample_rate_;
  uint8_t bit_depth_;
  bool zero_run_enabled_ = true;
  bool partitioning_enabled_ = true;
  bool debug_partitions_ = false;
  bool debug_lpc_ = false, debug_stereo_est_ = false, debug_zr_ = false;
  size_t thread_count_ = 0;
  size_t device_count_ = 0;
  bool debug_any() const { return debug_partitions_ || debug_lpc_ || debug_stereo_est_ || debug_zr_; }
  // prints the decision log of the encode that just ran on `ctx` (std::cerr, the reference's Debug-build lines)
  void print_decisions(lacb_ctx* ctx, uint8_t effective_stereo_mode) const;
};

class Decoder {
 public:
  explicit Decoder(ThreadCollector* collector = nullptr) : collector_(collector) {}
  void decode(const uint8_t* data, size_t size, std::vector<int32_t>& left, std::vector<int32_t>& right,
              FrameHeader* out_header = nullptr);
  // GPU-path extension: decode straight to packed WAV sample bytes (the CLI fast path,
  // src/main.cpp:184-430); `out` must hold frames * channels * bit_depth/8 bytes.
  void decode_packed(const uint8_t* data, size_t size, std::vector<uint8_t>& out, FrameHeader& hdr,
                     uint64_t& frames);
  // GPU-path extension, the CLI fast path with a mapped output (src/main.cpp:287-311,386-393): the WAV (RF64 when
  // large files are allowed and RIFF cannot hold it) is created at `path` and the device packs the samples
  // straight into the mapping.
  void decode_packed_to_file(const uint8_t* data, size_t size, const std::string& path, FrameHeader& hdr,
                             uint64_t& frames);
  void set_thread_count(size_t max_threads) { thread_count_ = max_threads; }
  void set_device_count(size_t devices) { device_count_ = devices; }  // 0 = LAC_DEVICES or 1
  // lifts the reference's 1 GiB decoded-PCM cap and the RIFF size cap; off by default
  void set_allow_large(bool allow) { allow_large_ = allow; }

 private:
  ThreadCollector* collector_;
  size_t thread_count_ = 0;
  size_t device_count_ = 0;
  bool allow_large_ = false;
};

}  // namespace LAC

namespace Block {

// block/constants.hpp:6-15
constexpr uint32_t MAX_BLOCK_SIZE = 16384;
constexpr uint32_t MIN_CANONICAL_NON_FINAL_BLOCK_SIZE = 256;
constexpr uint32_t ZERO_RUN_MIN_LENGTH = 4;
constexpr uint32_t ZERO_RUN_LENGTH_K = 2;
constexpr uint32_t MIN_PARTITION_SIZE = 32;
constexpr uint8_t MAX_PARTITION_ORDER = 8;
constexpr uint8_t PARTITION_FLAG = 0x80;
constexpr uint8_t RESIDUAL_RESERVED_MASK = 0x10;
constexpr uint8_t PARTITION_ORDER_SHIFT = 0;
constexpr uint8_t PARTITION_ORDER_MASK = 0x0F;

class Encoder {
 public:
  explicit Encoder(int order, bool debug_lpc = false, bool debug_zr = false);
  std::vector<uint8_t> encode(const std::vector<int32_t>& pcm);
  void set_zero_run_enabled(bool enabled) { zero_run_enabled_ = enabled; }
  void set_debug_block_index(size_t) {}
  void set_partitioning_enabled(bool enabled) { partitioning_enabled_ = enabled; }
  void set_debug_partitions(bool) {}

 private:
  int order_;
  bool zero_run_enabled_ = false;      // reference defaults (block/encoder.hpp:25-26)
  bool partitioning_enabled_ = false;
};

class Decoder {
 public:
  Decoder() = default;
  bool decode(BitReader& br, uint32_t block_size, std::vector<int32_t>& out);
  // Decodes from wherever the reader stands (any bit position).
  bool decode_into(BitReader& br, uint32_t block_size, int32_t* out);
};

}  // namespace Block

namespace lacb_host {
// number of CUDA devices the runtime sees (0 => every codec call throws: no CPU fallback)
int device_count();
// resolves --devices / LAC_DEVICES: 0 or unset => 1
size_t resolve_devices(size_t requested);
// creates the device contexts a call with `requested` devices will use (CUDA start-up: driver initialisation,
// context and module load), so a caller can overlap it with its own file work; errors surface at the first codec call
void warm_up(size_t requested) noexcept;
}  // namespace lacb_host
