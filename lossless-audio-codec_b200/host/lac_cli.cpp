// lac_cli.cpp -- `lac_cli encode|decode|selftest` on the GPU path.
//
// Same workflow and flags as the reference CLI (src/main.cpp:600-918): --stereo-mode=lr|ms
// (default: auto per block), --threads=N / LAC_THREADS, --debug-threads, --no-partitioning;
// plus --devices=N / LAC_DEVICES to shard the block range across GPUs.  Files are written
// to a private temporary name and published with rename(), and input == output is refused.
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "lac_host.hpp"
#include "wav_io.hpp"

namespace {

void usage() {
  std::cerr << "Usage:\n"
            << "  lac_cli encode input.wav output.lac [--stereo-mode=lr|ms] [--threads=N] [--devices=N] "
               "[--debug-threads] [--debug-lpc] [--debug-stereo-est] [--debug-zr] [--debug-partitions] [--no-partitioning] [--allow-large]\n"
            << "  lac_cli decode input.lac output.wav [--threads=N] [--devices=N] [--debug-threads] [--allow-large]\n"
            << "  lac_cli selftest\n"
            << "  lac_cli batch list.txt        (one encode/decode command per line, one process)\n"
            << "  lac_cli serve [socket] [--devices=N]   (resident process; clients forward when LAC_SERVER=socket)\n"
            << "  lac_cli shutdown              (stops the server named by LAC_SERVER)\n";
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

// ---------------------------------------------------------------------------
// `lac_cli serve [socket]`: a resident process that keeps the CUDA contexts (and the warmed-up workspaces) alive and
// executes encode / decode commands sent by other `lac_cli` invocations.  A one-shot GPU command pays 0.7 - 5 s of
// CUDA start-up (driver + context + module, measured on the bench box: profiles/r2_cli_breakdown.jsonl) before its
// ~20 ms of work; with a server the same command line is forwarded over a unix socket and costs the work plus a
// process spawn.  Clients forward when LAC_SERVER names the socket (and fall back to running locally when nobody
// listens).  One request at a time; paths are made absolute by the client; stdout / stderr / exit status are the
// ones the command would have produced.
namespace {

std::string default_socket() {
  const char* e = std::getenv("LAC_SERVER");
  if (e && *e) return e;
  return "/tmp/lac_cli-" + std::to_string((long)::getuid()) + ".sock";
}
bool write_all(int fd, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  while (n) {
    const ssize_t w = ::write(fd, c, n);
    if (w <= 0) return false;
    c += w;
    n -= (size_t)w;
  }
  return true;
}
bool read_all(int fd, void* p, size_t n) {
  char* c = static_cast<char*>(p);
  while (n) {
    const ssize_t r = ::read(fd, c, n);
    if (r <= 0) return false;
    c += r;
    n -= (size_t)r;
  }
  return true;
}
bool send_str(int fd, const std::string& s) {
  const uint32_t n = (uint32_t)s.size();
  return write_all(fd, &n, 4) && write_all(fd, s.data(), n);
}
bool recv_str(int fd, std::string& s) {
  uint32_t n = 0;
  if (!read_all(fd, &n, 4) || n > (1u << 20)) return false;
  s.resize(n);
  return n == 0 || read_all(fd, s.data(), n);
}
int open_socket(const std::string& path, bool listen_side) {
  sockaddr_un addr{};
  if (path.size() >= sizeof addr.sun_path) return -1;
  addr.sun_family = AF_UNIX;
  std::strcpy(addr.sun_path, path.c_str());
  const int fd = ::socket(AF_UNIX, SOCK_STREAM, 0);
  if (fd < 0) return -1;
  if (listen_side) {
    ::unlink(path.c_str());
    const mode_t old = ::umask(0077);  // the socket is the owner's only
    const bool ok = ::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) == 0 && ::listen(fd, 16) == 0;
    ::umask(old);
    if (!ok) { ::close(fd); return -1; }
  } else if (::connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0) {
    ::close(fd);
    return -1;
  }
  return fd;
}

}  // namespace

static int run_server(const std::string& path, size_t devices) {
  const int ls = open_socket(path, true);
  if (ls < 0) {
    std::cerr << "Failed to listen on " << path << "\n";
    return 1;
  }
  lacb_host::warm_up(devices);
  std::cout << "lac_cli serving on " << path << " (" << lacb_host::device_count() << " GPU(s) visible)" << std::endl;
  for (;;) {
    const int fd = ::accept(ls, nullptr, nullptr);
    if (fd < 0) continue;
    std::string count;
    std::vector<std::string> words;
    bool ok = recv_str(fd, count);
    for (int i = 0, n = ok ? std::atoi(count.c_str()) : 0; ok && i < n; ++i) {
      std::string w;
      ok = recv_str(fd, w);
      words.push_back(w);
    }
    int rc = 1;
    bool stop = false;
    std::ostringstream out, err;
    if (ok && !words.empty()) {
      if (words[0] == "shutdown") {
        stop = true;
        rc = 0;
      } else if (words[0] == "encode" || words[0] == "decode") {
        std::vector<char*> av;
        std::string self = "lac_cli";
        av.push_back(self.data());
        for (std::string& w : words) av.push_back(w.data());
        std::streambuf* so = std::cout.rdbuf(out.rdbuf());
        std::streambuf* se = std::cerr.rdbuf(err.rdbuf());
        rc = run_command((int)av.size(), av.data());
        std::cout.rdbuf(so);
        std::cerr.rdbuf(se);
      } else {
        err << "lac_cli serve: only encode / decode are forwarded\n";
      }
    }
    send_str(fd, std::to_string(rc));
    send_str(fd, out.str());
    send_str(fd, err.str());
    ::close(fd);
    if (stop) break;
  }
  ::close(ls);
  ::unlink(path.c_str());
  return 0;
}

// Forwards `argv[1..]` to a server if one listens; returns -1 when the command has to run locally.
static int try_forward(int argc, char** argv) {
  if (!std::getenv("LAC_SERVER") || argc < 2) return -1;
  const std::string cmd = argv[1];
  if (cmd != "encode" && cmd != "decode" && cmd != "shutdown") return -1;
  const int fd = open_socket(default_socket(), false);
  if (fd < 0) return -1;
  std::vector<std::string> words;
  for (int i = 1; i < argc; ++i) {
    std::string w = argv[i];
    if ((i == 2 || i == 3) && cmd != "shutdown") w = std::filesystem::absolute(w).string();  // the server has its own cwd
    words.push_back(w);
  }
  bool ok = send_str(fd, std::to_string(words.size()));
  for (const std::string& w : words) ok = ok && send_str(fd, w);
  std::string rc, out, err;
  ok = ok && recv_str(fd, rc) && recv_str(fd, out) && recv_str(fd, err);
  ::close(fd);
  if (!ok) return -1;
  std::cout << out;
  std::cerr << err;
  return std::atoi(rc.c_str());
}

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
  if (on) {  // one write per line: the warm-up thread reports too
    std::ostringstream line;
    line << "[lac_cli " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()
         << " ms] " << what << "\n";
    std::cerr << line.str();
  }
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
    if (cmd == "serve") {
      size_t devs = 0;
      std::string path = default_socket();
      for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        if (a.rfind("--devices=", 0) == 0) devs = parse_count_flag(a, "--devices=");
        else path = a;
      }
      return run_server(path, devs);
    }
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
    bool debug_lpc = false, debug_stereo_est = false, debug_zr = false, debug_partitions = false;
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
      else if (a == "--debug-lpc") debug_lpc = true;
      else if (a == "--debug-stereo-est") debug_stereo_est = true;
      else if (a == "--debug-zr") debug_zr = true;
      else if (a == "--debug-partitions") debug_partitions = true;
      else {
        std::cerr << "Unknown option: " << a << "\n";
        usage();
        return 1;
      }
    }
    if (threads == 0) threads = LAC::parse_thread_limit(std::getenv("LAC_THREADS"));
    // CUDA start-up (driver initialisation, context, module load: the largest fixed cost of a one-shot run)
    // proceeds on its own thread while this one maps the input file
    std::thread warm([devices] {
      lacb_host::warm_up(devices);
      milestone("cuda ready");
    });
    struct Joiner {
      std::thread& t;
      ~Joiner() { if (t.joinable()) t.join(); }
    } joiner{warm};

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
      LAC::Encoder enc(12, stereo_mode, info.sample_rate, info.bit_depth, debug_lpc, debug_stereo_est, debug_zr);
      enc.set_debug_partitions(debug_partitions);
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
      if (debug_zr) {  // src/main.cpp:676-690: the same file without zero-run tokens, for comparison
        LAC::Encoder baseline(12, stereo_mode, info.sample_rate, info.bit_depth, debug_lpc, debug_stereo_est, false);
        baseline.set_zero_run_enabled(false);
        baseline.set_partitioning_enabled(partitioning);
        baseline.set_debug_partitions(debug_partitions);
        baseline.set_thread_count(threads);
        const std::vector<uint8_t> base = baseline.encode_packed(in.data + data_off, info.frames, (uint8_t)info.channels);
        const double gain = base.empty() ? 0.0 : (1.0 - (double)lac_bytes / (double)base.size()) * 100.0;
        std::cout << "[debug-zr] baseline_bytes=" << base.size() << " zr_bytes=" << lac_bytes << " gain=" << gain << "%\n";
      }
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

int main(int argc, char** argv) {
  // A resident server asks for more hardware queues than the default 8, so that the slice streams of the host pipelines
  // do not share one (read by the CUDA runtime when it starts; the library keeps twelve instead of four decode slices
  // in flight when it sees it).  Not for one-shot runs: the extra queues add 1 - 2.5 s to CUDA's start-up
  // (tools/cli_conn_check.py), far more than the 0.7 ms per decode they save.
  if (argc >= 2 && std::string(argv[1]) == "serve") setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  // A one-shot encode / decode on one GPU only needs that GPU: hiding the others from the CUDA runtime keeps
  // its initialisation from enumerating and mapping every device of an 8-GPU box (honours an existing setting).
  if (argc >= 2 && (std::string(argv[1]) == "encode" || std::string(argv[1]) == "decode")) {
    bool multi = std::getenv("LAC_DEVICES") != nullptr;
    for (int i = 4; i < argc; ++i)
      if (std::string(argv[i]).rfind("--devices=", 0) == 0 && std::string(argv[i]) != "--devices=1") multi = true;
    const int forwarded = try_forward(argc, argv);
    if (forwarded >= 0) return forwarded;
    if (!multi) ::setenv("CUDA_VISIBLE_DEVICES", "0", 0);
    // one-shot run: skip the CUDA context teardown (~0.3 s) once the outputs are published
    const int rc = run_command(argc, argv);
    std::cout.flush();
    std::cerr.flush();
    std::fflush(nullptr);
    ::_exit(rc);
  }
  if (argc == 2 && std::string(argv[1]) == "shutdown") {
    const int forwarded = try_forward(argc, argv);
    if (forwarded >= 0) return forwarded;
    std::cerr << "no server listening (LAC_SERVER)\n";
    return 1;
  }
  return run_command(argc, argv);
}
