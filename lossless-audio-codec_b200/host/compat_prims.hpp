// compat_prims.hpp -- host-side helpers with the reference's Rice / LPC class interfaces
// (src/codec/rice/rice.hpp:9-44, src/codec/lpc/lpc.hpp:5-32).
//
// They exist so code written against those headers -- in practice the reference's own unit
// tests, which hand-build and walk bitstreams -- compiles against the GPU facade.  They are
// not on the encode/decode path: Block::/LAC:: classes never call them, the device kernels
// implement the same rules (csrc/lacb_encode.cuh, lacb_dec_kernels.cuh, lacb_f80.cuh).
// LPC::analyze_block_q15 runs on the GPU (k_autocorr + k_levinson through lacb_lpc_analyze).
#pragma once
#include <array>
#include <cstdint>
#include <vector>

#include "lac_host.hpp"

class Rice {
 public:
  static constexpr uint32_t kDriftWindow = 256;
  static constexpr uint32_t kMicroWindow = 96;

  struct AdaptState {
    uint64_t previous_sum = 0;
    uint32_t window_index = 0;
    uint32_t micro_index = 0;
    uint32_t window_filled = 0;
    uint64_t window_sum = 0;
    uint16_t large_q_count = 0;
    uint16_t zero_q_count = 0;
    std::array<uint32_t, kDriftWindow> recent_u{};
    std::array<uint8_t, kMicroWindow> large_flags{};
    std::array<uint8_t, kMicroWindow> zero_flags{};
  };

  static void encode(BitWriter& w, int32_t value, uint32_t k);
  static bool decode(BitReader& r, uint32_t k, int32_t& value);
  static uint32_t adapt_k(uint64_t sum, uint32_t count, AdaptState& state);
};

class LPC {
 public:
  explicit LPC(int order) : order_(order) {}
  int get_order() const { return order_; }

  bool analyze_block_q15(const std::vector<int32_t>& block, std::vector<int16_t>& coeffs_q15, int& used_order,
                         long double* energy_out = nullptr) const;
  void compute_residual_q15(const std::vector<int32_t>& block, const std::vector<int16_t>& coeffs_q15,
                            std::vector<int32_t>& residual, int* used_order_inout = nullptr) const;
  bool restore_from_residual_q15(const std::vector<int32_t>& residual, const std::vector<int16_t>& coeffs_q15,
                                 std::vector<int32_t>& out_block) const;

 private:
  int order_;
};
