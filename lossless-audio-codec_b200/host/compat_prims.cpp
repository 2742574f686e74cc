#include "compat_prims.hpp"

#include <algorithm>
#include <stdexcept>

#include "../../include/lac_b200.h"

namespace lacb_host {
lacb_ctx* shared_context(int device);  // lac_host.cpp
std::mutex& shared_context_mutex(int device);
}  // namespace lacb_host

namespace {
uint32_t zigzag(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }
int32_t unzigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }
uint32_t width_of(uint64_t v) {
  uint32_t w = 0;
  for (; v; v >>= 1) ++w;
  return w;
}
}  // namespace

void Rice::encode(BitWriter& w, int32_t value, uint32_t k) {
  const uint32_t u = zigzag(value);
  const uint32_t q = k >= 32u ? 0u : (u >> k);
  w.write_unary_ones(q);
  w.write_bit(0u);
  if (k > 0) w.write_bits(k >= 32u ? u : (u & ((1u << k) - 1u)), (int)k);
}

bool Rice::decode(BitReader& r, uint32_t k, int32_t& value) {
  if (k > 31u) return false;
  uint32_t q = 0;
  if (!r.read_unary_ones(0xFFFFFFFFu >> k, q)) return false;
  const uint32_t rem = k ? r.read_bits((int)k) : 0u;
  if (r.has_error()) return false;
  value = unzigzag((q << k) | rem);
  return true;
}

// Stateful adaptive Rice parameter: running mean of the zig-zag values, a +-1 drift bias
// from the mean of the last 256 values, and a +-1 micro bias from the share of large
// (q > 3) / zero quotients among the last 96.
uint32_t Rice::adapt_k(uint64_t sum, uint32_t count, AdaptState& st) {
  if (count == 0) return 0;
  const uint64_t cur = sum - st.previous_sum;
  st.previous_sum = sum;
  const uint32_t mi = st.micro_index;
  st.large_q_count = (uint16_t)(st.large_q_count - st.large_flags[mi]);
  st.zero_q_count = (uint16_t)(st.zero_q_count - st.zero_flags[mi]);
  if (st.window_filled < kDriftWindow) ++st.window_filled;
  else st.window_sum -= st.recent_u[st.window_index];
  st.recent_u[st.window_index] = (uint32_t)cur;
  st.window_sum += cur;

  const uint64_t mean = (sum + (count >> 1)) / count;
  uint32_t k = mean > 1 ? std::min<uint32_t>(31u, width_of(mean - 1)) : 0u;
  const uint32_t q = k >= 31u ? 0u : (uint32_t)(cur >> k);
  const uint8_t large = q > 3u, zero = q == 0u;
  st.large_q_count = (uint16_t)(st.large_q_count + large);
  st.zero_q_count = (uint16_t)(st.zero_q_count + zero);
  st.large_flags[mi] = large;
  st.zero_flags[mi] = zero;

  int bias = 0;
  if (st.window_filled > 0 && mean > 0) {
    const uint64_t local = st.window_filled == kDriftWindow
                               ? (st.window_sum + kDriftWindow / 2) >> 8
                               : (st.window_sum + (st.window_filled >> 1)) / st.window_filled;
    if (local * 3 > mean * 4) bias = 1;
    else if (local * 4 + 3 < mean * 3) bias = -1;
  }
  if (st.window_index + 1 >= kMicroWindow || st.window_filled >= kMicroWindow) {
    const uint32_t wsz = st.window_filled >= kMicroWindow ? kMicroWindow : st.window_filled;
    if ((uint32_t)st.large_q_count * 4 >= wsz * 3) bias = std::min(bias + 1, 1);
    else if ((uint32_t)st.zero_q_count * 5 >= wsz * 4) bias = std::max(bias - 1, -1);
  }
  st.micro_index = st.micro_index + 1 == kMicroWindow ? 0 : st.micro_index + 1;
  st.window_index = (st.window_index + 1) & (kDriftWindow - 1);
  return (uint32_t)std::clamp((int)k + bias, 0, 31);
}

bool LPC::analyze_block_q15(const std::vector<int32_t>& block, std::vector<int16_t>& coeffs_q15, int& used_order,
                            long double* energy_out) const {
  coeffs_q15.assign((size_t)order_ + 1, 0);
  used_order = 0;
  if (energy_out) {
    long double e = 0.0L;
    for (int32_t v : block) e += (long double)((int64_t)v * (int64_t)v);
    *energy_out = e;
  }
  if (order_ < 4 || order_ > 12 || (order_ & 1) || block.empty() || block.size() > 16384)
    throw std::invalid_argument("LPC(order): the GPU analysis covers the codec's orders 4, 6, 8, 10, 12");
  std::lock_guard<std::mutex> lock(lacb_host::shared_context_mutex(0));
  const int used = lacb_lpc_analyze(lacb_host::shared_context(0), block.data(), (uint32_t)block.size(), order_,
                                    coeffs_q15.data());
  if (used < 0) throw std::runtime_error("LAC B200 backend: lacb_lpc_analyze failed");
  used_order = used;
  return used > 0;
}

// open-loop Q15 residual with the overflow fallback to lower orders (12, 10, 8, 6, 4, then none)
void LPC::compute_residual_q15(const std::vector<int32_t>& block, const std::vector<int16_t>& c,
                               std::vector<int32_t>& residual, int* used_order_inout) const {
  const size_t n = block.size();
  residual.assign(n, 0);
  const int avail = std::min<int>(order_, (int)c.size() - 1);
  int start = used_order_inout ? std::clamp(*used_order_inout, 0, avail) : avail;
  std::vector<int> attempts{start};
  for (int f : {12, 10, 8, 6, 4})
    if (f < start && f <= avail) attempts.push_back(f);
  for (int ord : attempts) {
    if (ord <= 0) break;
    bool ok = true;
    for (size_t i = 0; i < n && ok; ++i) {
      int64_t acc = 0;
      const int taps = (int)std::min<size_t>((size_t)ord, i);
      for (int t = 1; t <= taps; ++t) acc += (int64_t)c[(size_t)t] * (int64_t)block[i - (size_t)t];
      const int64_t d = (int64_t)block[i] - (acc >> 15);
      ok = d >= INT32_MIN && d <= INT32_MAX;
      if (ok) residual[i] = (int32_t)d;
    }
    if (ok) {
      if (used_order_inout) *used_order_inout = ord;
      return;
    }
  }
  residual = block;
  if (used_order_inout) *used_order_inout = 0;
}

bool LPC::restore_from_residual_q15(const std::vector<int32_t>& residual, const std::vector<int16_t>& c,
                                    std::vector<int32_t>& out) const {
  const size_t n = residual.size();
  std::vector<int32_t> x(n);
  const int avail = std::min<int>(order_, (int)c.size() - 1);
  for (size_t i = 0; i < n; ++i) {
    int64_t acc = 0;
    const int taps = (int)std::min<size_t>((size_t)avail, i);
    for (int t = 1; t <= taps; ++t) acc += (int64_t)c[(size_t)t] * (int64_t)x[i - (size_t)t];
    const int64_t s = (acc >> 15) + (int64_t)residual[i];
    if (s < INT32_MIN || s > INT32_MAX) return false;
    x[i] = (int32_t)s;
  }
  out.swap(x);
  return true;
}
