#include "wav_io.hpp"

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
}  // namespace

bool read_wav_packed(const std::string& path, WavInfo& info, std::vector<uint8_t>& pcm, bool allow_large) {
  info = WavInfo{};
  pcm.clear();
  File file(path.c_str(), "rb");
  if (!file.f) return false;
  if (std::fseek(file.f, 0, SEEK_END) != 0) return false;
  const long long end = std::ftell(file.f);
  if (end < 12) return false;
  const uint64_t file_size = (uint64_t)end;
  std::rewind(file.f);
  uint8_t hdr[12];
  if (std::fread(hdr, 1, 12, file.f) != 12) return false;
  if (std::memcmp(hdr, "RIFF", 4) != 0 || std::memcmp(hdr + 8, "WAVE", 4) != 0) return false;
  if ((uint64_t)le32(hdr + 4) + 8u != file_size) return false;

  bool got_fmt = false, got_data = false;
  uint16_t block_align = 0;
  uint64_t remaining = file_size - 12u;
  while (remaining > 0) {
    if (remaining < 8u) return false;
    uint8_t ch[8];
    if (std::fread(ch, 1, 8, file.f) != 8) return false;
    remaining -= 8u;
    const uint32_t size = le32(ch + 4);
    const uint64_t padded = (uint64_t)size + (size & 1u);
    if (padded > remaining) return false;
    if (std::memcmp(ch, "fmt ", 4) == 0) {
      if (got_fmt || got_data || size != 16u) return false;
      uint8_t f[16];
      if (std::fread(f, 1, 16, file.f) != 16) return false;
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
      pcm.resize(size);
      if (std::fread(pcm.data(), 1, size, file.f) != size) return false;
      info.frames = frames;
      got_data = true;
    } else {
      if (std::fseek(file.f, (long)size, SEEK_CUR) != 0) return false;
    }
    if ((size & 1u) && std::fseek(file.f, 1, SEEK_CUR) != 0) return false;
    remaining -= padded;
  }
  if (!got_fmt || !got_data) {
    pcm.clear();
    return false;
  }
  return true;
}

bool write_wav_packed(const std::string& path, const WavInfo& info, const uint8_t* pcm, uint64_t pcm_bytes) {
  if ((info.channels != 1 && info.channels != 2) || (info.bit_depth != 16 && info.bit_depth != 24) ||
      !rate_ok(info.sample_rate))
    return false;
  const uint64_t pad = pcm_bytes & 1u;
  if (36u + pcm_bytes + pad > 0xFFFFFFFFull) return false;  // classic RIFF limit (src/io/wav_io.cpp:309)
  uint8_t h[44];
  std::memcpy(h, "RIFF", 4);
  put32(h + 4, (uint32_t)(36u + pcm_bytes + pad));
  std::memcpy(h + 8, "WAVEfmt ", 8);
  put32(h + 16, 16);
  put16(h + 20, 1);
  put16(h + 22, info.channels);
  put32(h + 24, info.sample_rate);
  const uint16_t align = (uint16_t)(info.channels * (info.bit_depth / 8));
  put32(h + 28, info.sample_rate * align);
  put16(h + 32, align);
  put16(h + 34, info.bit_depth);
  std::memcpy(h + 36, "data", 4);
  put32(h + 40, (uint32_t)pcm_bytes);
  File file(path.c_str(), "wb");
  if (!file.f) return false;
  if (std::fwrite(h, 1, 44, file.f) != 44) return false;
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
