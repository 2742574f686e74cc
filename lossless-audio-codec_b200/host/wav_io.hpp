// wav_io.hpp -- strict RIFF/WAVE PCM reader/writer of the host side.
//
// Accept/reject rules follow the reference parser (src/io/wav_io.cpp:162-278): RIFF size
// must match the file, one 16-byte PCM fmt chunk before one non-empty data chunk, 16/24-bit,
// 1-2 channels, 44.1/48/96/192 kHz, consistent block_align / byte_rate, unknown chunks
// skipped, odd chunks padded.  Unlike the reference it keeps the samples as the packed
// little-endian bytes of the data chunk (the GPU de-interleaves them)
// instead of issuing one ifstream read per sample.
#pragma once
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

struct WavInfo {
  uint16_t channels = 0;
  uint32_t sample_rate = 0;
  uint8_t bit_depth = 0;
  uint64_t frames = 0;
};

constexpr uint64_t kMaxDecodedPcmBytes = 1ull << 30;  // src/io/wav_io.cpp:13 (int32 planes)

// pcm receives frames * channels * bit_depth/8 bytes.  allow_large lifts the reference's
// 1 GiB decoded-PCM cap (needed for BASELINE configs 3 and 4).
bool read_wav_packed(const std::string& path, WavInfo& info, std::vector<uint8_t>& pcm, bool allow_large = false);
bool write_wav_packed(const std::string& path, const WavInfo& info, const uint8_t* pcm, uint64_t pcm_bytes);
// --- mapped-file forms (CLI fast paths) --------------------------------------
// A read-only mapping of a whole file, or a writable MAP_SHARED mapping of a file created at a given size.
struct MappedFile {
  uint8_t* data = nullptr;
  uint64_t size = 0;
  int fd = -1;
  bool writable = false;
  MappedFile() = default;
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  ~MappedFile() { close(); }
  bool open_read(const std::string& path);
  bool create(const std::string& path, uint64_t bytes);
  bool resize(uint64_t bytes);  // shrinks / grows the file; the mapping is dropped (call before close)
  void close();
  // Faults the first `bytes` of a writable mapping in on `threads` helper threads (MADV_POPULATE_WRITE per
  // range; a fresh file is otherwise faulted page by page by whoever copies into it, one thread, ~1.4 GB/s on
  // the bench box).  Returns at once; wait_populated() joins the helpers (close() and the destructor do too).
  void populate_async(uint64_t bytes, unsigned threads);
  void wait_populated();

 private:
  std::vector<std::thread> helpers_;
};
// Parses the RIFF / RF64 structure of a WAV held in memory (same accept / reject rules as read_wav_packed) and
// returns where the sample bytes are: nothing is copied.  RF64 (EBU Tech 3306: 'RF64' + 'ds64' chunk carrying the
// 64-bit riff / data sizes) is accepted only with allow_large, like every size above the reference's 1 GiB cap.
bool locate_wav_data(const uint8_t* file, uint64_t file_size, WavInfo& info, uint64_t& data_offset,
                     uint64_t& data_bytes, bool allow_large);
// Header of a WAV with pcm_bytes of samples: 44 bytes of classic RIFF, or 80 bytes of RF64 when the RIFF size
// field cannot hold it (only produced when the caller opted in to large files).
std::vector<uint8_t> wav_header(const WavInfo& info, uint64_t pcm_bytes, bool* rf64 = nullptr);

// int32 plane variants for callers that hold planes (LAC::Decoder::decode output)
bool write_wav_planes(const std::string& path, const WavInfo& info, const std::vector<int32_t>& left,
                      const std::vector<int32_t>& right);
void unpack_planes(const WavInfo& info, const uint8_t* pcm, std::vector<int32_t>& left, std::vector<int32_t>& right);

// The reference's int32-plane interface (src/io/wav_io.hpp:6-26).
bool read_wav(const std::string& path, std::vector<int32_t>& left, std::vector<int32_t>& right, uint16_t& channels,
              uint32_t& sample_rate, uint8_t& bit_depth);
bool write_wav(const std::string& path, const std::vector<int32_t>& left, const std::vector<int32_t>& right,
               uint16_t channels, uint32_t sample_rate, uint8_t bit_depth);
bool write_wav_unchecked_samples(const std::string& path, const std::vector<int32_t>& left,
                                 const std::vector<int32_t>& right, uint16_t channels, uint32_t sample_rate,
                                 uint8_t bit_depth);
