#include "wav_io.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {
uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
void put16(uint8_t* p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
bool rate_ok(uint32_t r) { return r == 44100 || r == 48000 || r == 96000 || r == 192000; }

struct File {
  FILE* f = nullptr;
  explicit File(const char* path, const char* mode) : f(std::fopen(path, mode)) {}
  ~File() { if (f) std::fclose(f); }
};

uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }
void put64(uint8_t* p, uint64_t v) { put32(p, (uint32_t)v); put32(p + 4, (uint32_t)(v >> 32)); }
}  // namespace

bool MappedFile::open_read(const std::string& path) {
  close();
  fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (::fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(); return false; }
  size = (uint64_t)st.st_size;
  if (size == 0) return true;  // empty file: valid handle, nothing mapped
  void* p = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  if (p == MAP_FAILED) { close(); return false; }
  data = static_cast<uint8_t*>(p);
  return true;
}
bool MappedFile::create(const std::string& path, uint64_t bytes) {
  close();
  fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
  if (fd < 0) return false;
  if (::ftruncate(fd, (off_t)bytes) != 0) { close(); return false; }
  size = bytes;
  writable = true;
  if (bytes == 0) return true;
  void* p = ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) { close(); return false; }
  data = static_cast<uint8_t*>(p);
  return true;
}
bool MappedFile::resize(uint64_t bytes) {
  if (fd < 0 || !writable) return false;
  wait_populated();
  if (data) ::munmap(data, size);
  data = nullptr;
  size = 0;
  return ::ftruncate(fd, (off_t)bytes) == 0;
}
void MappedFile::populate_async(uint64_t bytes, unsigned threads) {
  wait_populated();
  if (!data || !writable || size == 0) return;
  bytes = std::min<uint64_t>(bytes, size);
  threads = std::max(1u, std::min(threads, 32u));
  constexpr uint64_t kAlign = 2ull << 20;
  const uint64_t per = ((bytes + threads - 1) / threads + kAlign - 1) & ~(kAlign - 1);
  for (uint64_t off = 0; off < bytes; off += per) {
    uint8_t* p = data + off;
    const uint64_t len = std::min<uint64_t>(per, bytes - off);
    helpers_.emplace_back([p, len] {
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
      if (::madvise(p, len, MADV_POPULATE_WRITE) == 0) return;
      // older kernels: touch every page (a fresh file reads as zeros, so writing the byte back changes nothing)
      for (uint64_t i = 0; i < len; i += 4096) {
        volatile uint8_t* q = p + i;
        *q = *q;
      }
    });
  }
}
void MappedFile::wait_populated() {
  for (std::thread& t : helpers_)
    if (t.joinable()) t.join();
  helpers_.clear();
}
void MappedFile::close() {
  wait_populated();
  if (data) ::munmap(data, size);
  if (fd >= 0) ::close(fd);
  data = nullptr;
  size = 0;
  fd = -1;
  writable = false;
}

bool locate_wav_data(const uint8_t* file, uint64_t file_size, WavInfo& info, uint64_t& data_offset,
                     uint64_t& data_bytes, bool allow_large) {
  info = WavInfo{};
  data_offset = data_bytes = 0;
  if (!file || file_size < 12) return false;
  const bool rf64 = std::memcmp(file, "RF64", 4) == 0;
  if ((!rf64 && std::memcmp(file, "RIFF", 4) != 0) || std::memcmp(file + 8, "WAVE", 4) != 0) return false;
  if (rf64 && !allow_large) return false;  // the reference knows classic RIFF only (src/io/wav_io.cpp:190-200)
  uint64_t riff_size = le32(file + 4), ds64_data = 0;
  bool got_ds64 = false;
  bool got_fmt = false, got_data = false;
  uint16_t block_align = 0;
  uint64_t pos = 12;
  if (rf64) {  // 'ds64' must be the first chunk: riffSize, dataSize, sampleCount (u64 each), tableLength (u32)
    if (file_size < pos + 8 + 28 || std::memcmp(file + pos, "ds64", 4) != 0) return false;
    const uint32_t sz = le32(file + pos + 4);
    if (sz < 28 || (uint64_t)sz + (sz & 1u) > file_size - pos - 8) return false;
    riff_size = le64(file + pos + 8);
    ds64_data = le64(file + pos + 16);
    got_ds64 = true;
    pos += 8 + (uint64_t)sz + (sz & 1u);
  }
  if (riff_size + 8u != file_size) return false;
  while (pos < file_size) {
    if (file_size - pos < 8u) return false;
    const uint8_t* ch = file + pos;
    pos += 8u;
    uint64_t size = le32(ch + 4);
    if (got_ds64 && std::memcmp(ch, "data", 4) == 0 && size == 0xFFFFFFFFull) size = ds64_data;
    const uint64_t padded = size + (size & 1u);
    if (padded > file_size - pos) return false;
    if (std::memcmp(ch, "fmt ", 4) == 0) {
      if (got_fmt || got_data || size != 16u) return false;
      const uint8_t* f = file + pos;
      const uint16_t format = le16(f), channels = le16(f + 2), align = le16(f + 12), bits = le16(f + 14);
      const uint32_t rate = le32(f + 4), byte_rate = le32(f + 8);
      if (format != 1 || (bits != 16 && bits != 24) || !rate_ok(rate) || (channels != 1 && channels != 2)) return false;
      const uint16_t expect = (uint16_t)(channels * (bits / 8));
      if (align != expect || byte_rate != rate * expect) return false;
      info.channels = channels;
      info.sample_rate = rate;
      info.bit_depth = (uint8_t)bits;
      block_align = align;
      got_fmt = true;
    } else if (std::memcmp(ch, "data", 4) == 0) {
      if (!got_fmt || got_data || size == 0u || size % block_align != 0) return false;
      const uint64_t frames = size / block_align;
      if (!allow_large && frames * info.channels * 4ull > kMaxDecodedPcmBytes) return false;
      info.frames = frames;
      data_offset = pos;
      data_bytes = size;
      got_data = true;
    }
    pos += padded;
  }
  return got_fmt && got_data;
}

bool read_wav_packed(const std::string& path, WavInfo& info, std::vector<uint8_t>& pcm, bool allow_large) {
  info = WavInfo{};
  pcm.clear();
  MappedFile mf;
  if (!mf.open_read(path) || mf.size < 12) return false;
  uint64_t off = 0, bytes = 0;
  if (!locate_wav_data(mf.data, mf.size, info, off, bytes, allow_large)) return false;
  pcm.assign(mf.data + off, mf.data + off + bytes);
  return true;
}

std::vector<uint8_t> wav_header(const WavInfo& info, uint64_t pcm_bytes, bool* rf64_out) {
  const uint64_t pad = pcm_bytes & 1u;
  const bool rf64 = 36u + pcm_bytes + pad > 0xFFFFFFFFull;
  if (rf64_out) *rf64_out = rf64;
  const uint16_t align = (uint16_t)(info.channels * (info.bit_depth / 8));
  std::vector<uint8_t> h(rf64 ? 80 : 44, 0);
  uint8_t* p = h.data();
  if (rf64) {
    std::memcpy(p, "RF64", 4);
    put32(p + 4, 0xFFFFFFFFu);
    std::memcpy(p + 8, "WAVEds64", 8);
    put32(p + 16, 28);
    put64(p + 20, 72u + pcm_bytes + pad);  // riff size: everything after the first 8 bytes
    put64(p + 28, pcm_bytes);
    put64(p + 36, pcm_bytes / align);      // sample (frame) count
    put32(p + 44, 0);                      // no chunk-size table
    p += 36;                               // "fmt " starts at byte 48
  } else {
    std::memcpy(p, "RIFF", 4);
    put32(p + 4, (uint32_t)(36u + pcm_bytes + pad));
    std::memcpy(p + 8, "WAVE", 4);
  }
  std::memcpy(p + 12, "fmt ", 4);
  put32(p + 16, 16);
  put16(p + 20, 1);
  put16(p + 22, info.channels);
  put32(p + 24, info.sample_rate);
  put32(p + 28, info.sample_rate * align);
  put16(p + 32, align);
  put16(p + 34, info.bit_depth);
  std::memcpy(p + 36, "data", 4);
  put32(p + 40, rf64 ? 0xFFFFFFFFu : (uint32_t)pcm_bytes);
  return h;
}

bool write_wav_packed(const std::string& path, const WavInfo& info, const uint8_t* pcm, uint64_t pcm_bytes) {
  if ((info.channels != 1 && info.channels != 2) || (info.bit_depth != 16 && info.bit_depth != 24) ||
      !rate_ok(info.sample_rate))
    return false;
  const uint64_t pad = pcm_bytes & 1u;
  if (36u + pcm_bytes + pad > 0xFFFFFFFFull) return false;  // classic RIFF limit (src/io/wav_io.cpp:309)
  const std::vector<uint8_t> h = wav_header(info, pcm_bytes);
  File file(path.c_str(), "wb");
  if (!file.f) return false;
  if (std::fwrite(h.data(), 1, h.size(), file.f) != h.size()) return false;
  if (pcm_bytes && std::fwrite(pcm, 1, pcm_bytes, file.f) != pcm_bytes) return false;
  const uint8_t zero = 0;
  if (pad && std::fwrite(&zero, 1, 1, file.f) != 1) return false;
  return std::fflush(file.f) == 0;
}

void unpack_planes(const WavInfo& info, const uint8_t* pcm, std::vector<int32_t>& left, std::vector<int32_t>& right) {
  const uint32_t bps = info.bit_depth / 8u, ch = info.channels;
  left.resize(info.frames);
  right.resize(ch == 2 ? info.frames : 0);
  for (uint64_t i = 0; i < info.frames; ++i) {
    const uint8_t* p = pcm + i * bps * ch;
    for (uint32_t c = 0; c < ch; ++c, p += bps) {
      int32_t v;
      if (bps == 2) v = (int16_t)le16(p);
      else v = ((int32_t)(((uint32_t)p[0] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 24))) >> 8;
      (c ? right : left)[i] = v;
    }
  }
}

bool write_wav_planes(const std::string& path, const WavInfo& info, const std::vector<int32_t>& left,
                      const std::vector<int32_t>& right) {
  const uint32_t bps = info.bit_depth / 8u, ch = info.channels;
  std::vector<uint8_t> pcm((size_t)left.size() * bps * ch);
  for (size_t i = 0; i < left.size(); ++i) {
    uint8_t* p = pcm.data() + i * bps * ch;
    for (uint32_t c = 0; c < ch; ++c, p += bps) {
      const uint32_t v = (uint32_t)(c ? right[i] : left[i]);
      for (uint32_t k = 0; k < bps; ++k) p[k] = (uint8_t)(v >> (8 * k));
    }
  }
  return write_wav_packed(path, info, pcm.data(), pcm.size());
}

bool read_wav(const std::string& path, std::vector<int32_t>& left, std::vector<int32_t>& right, uint16_t& channels,
              uint32_t& sample_rate, uint8_t& bit_depth) {
  left.clear();
  right.clear();
  channels = 0;
  sample_rate = 0;
  bit_depth = 0;
  WavInfo info;
  std::vector<uint8_t> pcm;
  if (!read_wav_packed(path, info, pcm)) return false;
  unpack_planes(info, pcm.data(), left, right);
  channels = info.channels;
  sample_rate = info.sample_rate;
  bit_depth = info.bit_depth;
  return true;
}

static bool write_wav_common(const std::string& path, const std::vector<int32_t>& left,
                             const std::vector<int32_t>& right, uint16_t channels, uint32_t sample_rate,
                             uint8_t bit_depth, bool check) {
  if ((channels != 1 && channels != 2) || (bit_depth != 16 && bit_depth != 24) || !rate_ok(sample_rate)) return false;
  if (left.empty()) return false;
  if (channels == 2 && right.size() != left.size()) return false;
  if (channels == 1 && !right.empty()) return false;
  if (check) {
    const int32_t lo = bit_depth == 16 ? -32768 : -8388608, hi = bit_depth == 16 ? 32767 : 8388607;
    for (int32_t v : left)
      if (v < lo || v > hi) return false;
    for (int32_t v : right)
      if (v < lo || v > hi) return false;
  }
  WavInfo info;
  info.channels = channels;
  info.sample_rate = sample_rate;
  info.bit_depth = bit_depth;
  info.frames = left.size();
  return write_wav_planes(path, info, left, right);
}
bool write_wav(const std::string& path, const std::vector<int32_t>& left, const std::vector<int32_t>& right,
               uint16_t channels, uint32_t sample_rate, uint8_t bit_depth) {
  return write_wav_common(path, left, right, channels, sample_rate, bit_depth, true);
}
bool write_wav_unchecked_samples(const std::string& path, const std::vector<int32_t>& left,
                                 const std::vector<int32_t>& right, uint16_t channels, uint32_t sample_rate,
                                 uint8_t bit_depth) {
  return write_wav_common(path, left, right, channels, sample_rate, bit_depth, false);
}
