// lac_cli.cpp -- `lac_cli encode|decode|selftest` on the GPU path.
//
// Same workflow and flags as the reference CLI (src/main.cpp:600-918): --stereo-mode=lr|ms
// (default: auto per block), --threads=N / LAC_THREADS, --debug-threads, --no-partitioning;
// plus --devices=N / LAC_DEVICES to shard the block range across GPUs.  Files are written
// to a private temporary name and published with rename(), and input == output is refused.
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "lac_host.hpp"
#include "wav_io.hpp"

namespace {

void usage() {
  std::cerr << "Usage:\n"
            << "  lac_cli encode input.wav output.lac [--stereo-mode=lr|ms] [--threads=N] [--devices=N] "
               "[--debug-threads] [--no-partitioning] [--allow-large]\n"
            << "  lac_cli decode input.lac output.wav [--threads=N] [--devices=N] [--debug-threads] [--allow-large]\n"
            << "  lac_cli selftest\n"
            << "  lac_cli batch list.txt        (one encode/decode command per line, one process)\n";
}

size_t parse_count_flag(const std::string& arg, const char* name) {
  const std::string v = arg.substr(std::strlen(name));
  if (v.empty()) throw std::invalid_argument(std::string(name) + " requires a positive integer");
  for (char c : v)
    if (c < '0' || c > '9') throw std::invalid_argument(std::string(name) + " requires a positive integer");
  const unsigned long long n = std::stoull(v);
  if (n == 0) throw std::invalid_argument(std::string(name) + " requires a positive integer");
  return (size_t)n;
}

// Output is written inside a private (0700) temporary directory next to the destination and
// published with rename(): a failed run never touches an existing file, hard links and
// symlinks at the destination are replaced rather than written through, long file names and
// a restrictive umask work (the staging name is short and its mode is set explicitly).
struct Staged {
  std::filesystem::path final_path, tmp_dir, tmp_path;
  bool ok = false;
  explicit Staged(const std::string& p) : final_path(p) {
    std::filesystem::path parent = final_path.parent_path();
    if (parent.empty()) parent = ".";
    static int counter = 0;
    tmp_dir = parent / (".lacb-stage-" + std::to_string((long)getpid()) + "-" + std::to_string(counter++));
    if (::mkdir(tmp_dir.c_str(), 0700) == 0) {
      ::chmod(tmp_dir.c_str(), 0700);
      tmp_path = tmp_dir / "out";
      ok = true;
    }
  }
  bool publish() {
    if (!ok) return false;
    std::error_code ec;
    std::filesystem::rename(tmp_path, final_path, ec);
    return !ec;
  }
  ~Staged() {
    std::error_code ec;
    if (ok) {
      std::filesystem::remove(tmp_path, ec);
      std::filesystem::remove(tmp_dir, ec);
    }
  }
};

bool same_file(const std::string& a, const std::string& b) {
  if (a == b) return true;
  std::error_code ec;
  if (!std::filesystem::exists(b, ec)) return false;
  return std::filesystem::equivalent(a, b, ec) && !ec;
}

void print_threads(const char* label, const LAC::ThreadCollector& tc) {
  const auto ids = tc.snapshot();
  std::cout << label << ": " << ids.size() << " threads\n";
  for (const auto& id : ids) std::cout << "  " << id << "\n";
  std::cout << "GPU devices visible: " << lacb_host::device_count() << "\n";
}

int selftest() {
  const double pi = 3.14159265358979323846;
  for (uint32_t sr : {44100u, 48000u, 96000u, 192000u})
    for (uint8_t depth : {(uint8_t)16, (uint8_t)24}) {
      const size_t n = sr / 20 + 37;
      const double amp = depth == 16 ? 12000.0 : 2.6e6;
      std::vector<int32_t> l(n), r(n);
      for (size_t i = 0; i < n; ++i) {
        l[i] = (int32_t)std::lround(amp * std::sin(2 * pi * 440.0 * (double)i / sr));
        r[i] = (int32_t)std::lround(amp * 0.9 * std::sin(2 * pi * 443.0 * (double)i / sr));
      }
      for (int mode = 0; mode <= 3; ++mode) {  // LR, MS, auto, mono
        LAC::Encoder enc(12, (uint8_t)(mode == 3 ? 0 : mode), sr, depth);
        const std::vector<int32_t> none;
        const std::vector<uint8_t> bs = enc.encode(l, mode == 3 ? none : r);
        LAC::Decoder dec;
        std::vector<int32_t> dl, dr;
        FrameHeader hdr;
        const auto t0 = std::chrono::steady_clock::now();
        dec.decode(bs.data(), bs.size(), dl, dr, &hdr);
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        const bool ok = dl == l && (mode == 3 ? dr.empty() : dr == r) && hdr.sample_rate == sr && hdr.bit_depth == depth &&
                        hdr.channels == (mode == 3 ? 1 : 2) && hdr.stereo_mode == (mode == 3 ? 0 : mode);
        if (!ok) {
          std::cerr << "Selftest mismatch sr=" << sr << " depth=" << int(depth) << " mode=" << mode << "\n";
          return 1;
        }
        if (mode == 2)
          std::cout << "Selftest sr=" << sr << "Hz depth=" << int(depth) << " bytes=" << bs.size() << " decode_us=" << us << "\n";
      }
    }
  std::cout << "Selftest complete: adaptive block tests passed.\n";
  return 0;
}

}  // namespace

static int run_command(int argc, char** argv);

// `lac_cli batch list.txt`: one encode / decode command per line (same arguments as on the
// command line, whitespace separated, '#' starts a comment), executed in this process.  CUDA
// start-up (0.5 - 2 s per process) is paid once instead of once per file, which is what makes
// the GPU path worthwhile for collections of short files.  Exit status 1 if any line failed.
static int run_batch(const char* argv0, const char* list_path) {
  FILE* f = std::fopen(list_path, "r");
  if (!f) {
    std::cerr << "Failed to read batch list: " << list_path << "\n";
    return 1;
  }
  int failures = 0;
  char line[8192];
  while (std::fgets(line, sizeof line, f)) {
    std::vector<std::string> words;
    std::string cur;
    for (const char* p = line; *p && *p != '#'; ++p) {
      if (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') {
        if (!cur.empty()) words.push_back(cur);
        cur.clear();
      } else {
        cur.push_back(*p);
      }
    }
    if (!cur.empty()) words.push_back(cur);
    if (words.empty()) continue;
    if (words[0] == "batch") {
      std::cerr << "batch lists do not nest\n";
      ++failures;
      continue;
    }
    std::vector<char*> av;
    av.push_back(const_cast<char*>(argv0));
    for (std::string& w : words) av.push_back(w.data());
    if (run_command((int)av.size(), av.data()) != 0) ++failures;
  }
  std::fclose(f);
  return failures ? 1 : 0;
}

// LAC_TIMING=1: wall-clock milestones on stderr (where a short run spends its time)
static void milestone(const char* what) {
  static const bool on = std::getenv("LAC_TIMING") != nullptr;
  static const auto t0 = std::chrono::steady_clock::now();
  if (on)
    std::cerr << "[lac_cli " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()
              << " ms] " << what << "\n";
}

static int run_command(int argc, char** argv) {
  try {
    milestone("start");
    if (argc < 2) {
      usage();
      return 1;
    }
    const std::string cmd = argv[1];
    if (cmd == "selftest") return selftest();
    if (cmd == "batch" && argc == 3) return run_batch(argv[0], argv[2]);
    if ((cmd != "encode" && cmd != "decode") || argc < 4) {
      usage();
      return 1;
    }
    const std::string in_path = argv[2], out_path = argv[3];
    if (same_file(in_path, out_path)) {
      std::cerr << "Input and output paths must be different\n";
      return 1;
    }
    uint8_t stereo_mode = 2;
    size_t threads = 0, devices = 0;
    bool debug_threads = false, partitioning = true, allow_large = false;
    for (int i = 4; i < argc; ++i) {
      const std::string a = argv[i];
      if (a == "--stereo-mode=lr") stereo_mode = 0;
      else if (a == "--stereo-mode=ms") stereo_mode = 1;
      else if (a == "--stereo-mode=auto") stereo_mode = 2;
      else if (a.rfind("--threads=", 0) == 0) threads = parse_count_flag(a, "--threads=");
      else if (a.rfind("--devices=", 0) == 0) devices = parse_count_flag(a, "--devices=");
      else if (a == "--debug-threads") debug_threads = true;
      else if (a == "--no-partitioning") partitioning = false;
      else if (a == "--allow-large") allow_large = true;
      else if (a == "--debug-lpc" || a == "--debug-stereo-est" || a == "--debug-zr" || a == "--debug-partitions") {
      } else {
        std::cerr << "Unknown option: " << a << "\n";
        usage();
        return 1;
      }
    }
    if (threads == 0) threads = LAC::parse_thread_limit(std::getenv("LAC_THREADS"));

    if (cmd == "encode") {
      // the input WAV is mapped, not read: the device de-interleaves the data chunk where it lies
      MappedFile in;
      WavInfo info;
      uint64_t data_off = 0, data_bytes = 0;
      if (!in.open_read(in_path) || !locate_wav_data(in.data, in.size, info, data_off, data_bytes, allow_large)) {
        std::cerr << "Failed to read WAV: " << in_path << "\n";
        return 1;
      }
      milestone("wav mapped");
      LAC::ThreadCollector tc;
      LAC::Encoder enc(12, stereo_mode, info.sample_rate, info.bit_depth);
      enc.set_partitioning_enabled(partitioning);
      enc.set_thread_count(threads);
      enc.set_device_count(devices);
      Staged st(out_path);
      if (!st.ok) {
        std::cerr << "Failed to write LAC file: " << out_path << "\n";
        return 1;
      }
      uint64_t lac_bytes = 0;
      try {
        lac_bytes = enc.encode_packed_to_file(in.data + data_off, info.frames, (uint8_t)info.channels,
                                              st.tmp_path.string(), &tc);
      } catch (const std::runtime_error& e) {
        if (std::string(e.what()).rfind("failed to", 0) == 0) {
          std::cerr << "Failed to write LAC file: " << out_path << "\n";
          return 1;
        }
        throw;
      }
      milestone("encoded");
      if (!st.publish()) {
        std::cerr << "Failed to write LAC file: " << out_path << "\n";
        return 1;
      }
      milestone("lac written");
      std::cout << "Encoded " << in_path << " -> " << out_path << " (" << lac_bytes << " bytes)\n";
      if (debug_threads) print_threads("Thread usage", tc);
      return 0;
    }

    // decode: the .lac is mapped, the WAV is created at its final size, mapped, and filled by the device
    // (the reference's mmap fast path, src/main.cpp:184-430)
    MappedFile lac;
    if (!lac.open_read(in_path) || (!allow_large && lac.size > (1ull << 30))) {  // MAX_LAC_INPUT_BYTES, main.cpp:40
      std::cerr << "Failed to read LAC file: " << in_path << "\n";
      return 1;
    }
    milestone("lac mapped");
    LAC::ThreadCollector tc;
    LAC::Decoder dec(&tc);
    dec.set_thread_count(threads);
    dec.set_device_count(devices);
    dec.set_allow_large(allow_large);
    FrameHeader hdr;
    uint64_t frames = 0;
    Staged st(out_path);
    if (!st.ok) {
      std::cerr << "Failed to write WAV: " << out_path << "\n";
      return 1;
    }
    try {
      dec.decode_packed_to_file(lac.data, lac.size, st.tmp_path.string(), hdr, frames);
    } catch (const std::runtime_error& e) {
      if (std::string(e.what()).rfind("failed to", 0) == 0) {
        std::cerr << "Failed to write WAV: " << out_path << "\n";
        return 1;
      }
      std::cerr << "Decode failed: " << e.what() << "\n";
      return 1;
    }
    milestone("decoded");
    if (!st.publish()) {
      std::cerr << "Failed to write WAV: " << out_path << "\n";
      return 1;
    }
    milestone("wav written");
    std::cout << "Decoded " << in_path << " -> " << out_path << " (" << frames << " samples per channel)\n";
    if (debug_threads) print_threads("Decoder thread usage", tc);
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "Error: " << e.what() << "\n";
    return 1;
  }
}

int main(int argc, char** argv) { return run_command(argc, argv); }
