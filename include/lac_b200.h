/* include/lac_b200.h -- C ABI of liblac_b200.so, the B200 (sm_100a) implementation of
 * the LAC block-parallel encode/decode hot path.
 *
 * This is the drop-in boundary: the reference's host code (WAV / .lac container I/O,
 * block planning, CLI) keeps running on the CPU and calls these entry points where it
 * used to run its std::thread worker pools over Block::Encoder / Block::Decoder:
 *
 *   lacb_encode*        replaces the worker-pool region of LAC::Encoder::encode
 *                       (src/codec/lac/encoder.cpp:259-443: encode_block lambda,
 *                       estimate_stereo_mode, Block::Encoder::encode) and returns what
 *                       encoder.cpp:445-465 needs to write the block table + payload.
 *   lacb_decode*        replaces LAC::Decoder's decode_block pool
 *                       (src/codec/lac/decoder.cpp:167-292) and the CLI fast path's
 *                       inner loop (src/main.cpp:317-407), including mid/side
 *                       reconstruction, PCM range validation and WAV sample packing.
 *   lacb_encode_block   Block::Encoder::encode (src/codec/block/encoder.cpp:313).
 *   lacb_decode_block   Block::Decoder::decode_into (src/codec/block/decoder.cpp:64).
 *   lacb_lpc_analyze    LPC::analyze_block_q15 (src/codec/lpc/lpc.cpp:156).
 *
 * Plain pointers and sizes only; no exceptions cross this boundary.  Every function
 * returns 0 on success or a negative lacb_status; lacb_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device lacb_create fails.
 */
#ifndef LAC_B200_H
#define LAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lacb_ctx lacb_ctx; /* one context = one GPU + one stream + its workspaces */

typedef enum lacb_status {
  LACB_OK = 0,
  LACB_EINVAL = -1,  /* std::invalid_argument in the reference (lac/encoder.cpp:220-241) */
  LACB_EDECODE = -2, /* "[decode-error] ..." (lac/decoder.cpp:25-32) */
  LACB_ELIMIT = -3,  /* "encoded block size is outside format limits" (lac/encoder.cpp:447-450) */
  LACB_ECUDA = -4,   /* CUDA runtime failure */
  LACB_ENOMEM = -5
} lacb_status;

typedef enum lacb_layout {
  LACB_PLANAR_I32 = 0,   /* two int32 planes (left, right); right ignored for mono        */
  LACB_PACKED_LE = 1     /* interleaved little-endian 16/24-bit frames, as in a WAV chunk */
} lacb_layout;

typedef struct lacb_enc_params {
  uint32_t sample_rate;          /* carried for the caller's frame header only            */
  uint32_t bit_depth;            /* 16 or 24                                               */
  uint32_t channels;             /* 1 or 2                                                 */
  uint32_t stereo_mode;          /* 0 LR, 1 MS, 2 auto per block (--stereo-mode)           */
  uint32_t zero_run_enabled;     /* LAC::Encoder::set_zero_run_enabled (default 1)         */
  uint32_t partitioning_enabled; /* LAC::Encoder::set_partitioning_enabled (default 1)     */
  uint32_t validate_range;       /* 1: reject samples outside bit_depth like the reference */
} lacb_enc_params;

typedef struct lacb_dec_params {
  uint32_t bit_depth;   /* from the frame header */
  uint32_t channels;
  uint32_t stereo_mode;
} lacb_dec_params;

typedef struct lacb_err {
  int32_t code;          /* lacb_status */
  uint32_t block_index;  /* first failing block, when meaningful */
  uint32_t reason;       /* decoder: 1 stereo flag, 2 primary, 3 secondary, 4 range, 5 trailing payload */
  char msg[160];         /* reference-compatible message text */
} lacb_err;

/* Per-stage device times of the last call on this context (CUDA events on its stream). */
typedef struct lacb_timing {
  float h2d_ms, prep_ms, stereo_ms, lpc_ms, analyze_ms, finalize_ms, emit_ms, d2h_ms;
  float parse_ms, restore_ms, finish_ms;
  float total_ms;
} lacb_timing;

int lacb_create(int device, lacb_ctx** out);
void lacb_destroy(lacb_ctx* ctx);
const char* lacb_last_error(const lacb_ctx* ctx);
void lacb_free(void* p);
int lacb_device_count(void);
int lacb_get_timing(const lacb_ctx* ctx, lacb_timing* out);
/* Caps the slices of whole blocks the host-buffer paths keep in flight on internal streams (the
 * GPU path's counterpart of the reference's worker cap, LAC::Encoder::set_thread_count,
 * src/codec/lac/encoder.cpp:385-390): 0 = automatic, 1 = one stream (copy in, kernels, copy out,
 * nothing overlapped), n = at most n.  The bytes never depend on it. */
int lacb_set_concurrency(lacb_ctx* ctx, uint32_t max_slices_in_flight);

/* Encode `frames` frames held in HOST memory.  The block plan is the reference's:
 * fixed 16384-sample blocks, last one shorter (lac/encoder.cpp:59-69).
 *   pcm_a / pcm_b : LACB_PLANAR_I32 -> left / right int32 planes (pcm_b NULL for mono)
 *                   LACB_PACKED_LE  -> pcm_a = interleaved bytes, pcm_b unused
 *   payload_out   : malloc'ed concatenation of all block payloads (lacb_free)
 *   block_bytes   : caller array of ceil(frames/16384) entries, compressed bytes per block
 * Long inputs are processed as slices of whole blocks on internal streams, so that the
 * host<->device copies overlap the kernels; the bytes are those of a single pass.
 */
int lacb_encode(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a, const void* pcm_b,
                uint64_t frames, uint8_t** payload_out, uint64_t* payload_bytes, uint32_t* block_bytes,
                lacb_err* err);

/* Same, writing the payload into a caller buffer (use lacb_host_malloc memory to get full
 * PCIe speed).  Returns LACB_ENOMEM and the required size in *payload_bytes when
 * payload_cap is too small. */
int lacb_encode_to(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a, const void* pcm_b,
                   uint64_t frames, uint8_t* payload_buf, uint64_t payload_cap, uint64_t* payload_bytes,
                   uint32_t* block_bytes, lacb_err* err);

/* Same, with the PCM already resident on this context's device (d_a / d_b as pcm_a /
 * pcm_b above, device pointers) and the results left there.  *d_payload and
 * *d_block_bytes point into the context's workspace and stay valid until the next call
 * on the context. */
int lacb_encode_device(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* d_a, const void* d_b,
                       uint64_t frames, const uint8_t** d_payload, uint64_t* payload_bytes,
                       const uint32_t** d_block_bytes, lacb_err* err);

/* Device / pinned-host memory helpers for callers that keep data resident on the GPU
 * (all synchronous on the context's stream). */
int lacb_dev_malloc(lacb_ctx* ctx, uint64_t bytes, void** out);
int lacb_dev_free(lacb_ctx* ctx, void* p);
int lacb_host_malloc(lacb_ctx* ctx, uint64_t bytes, void** out);
int lacb_host_free(lacb_ctx* ctx, void* p);
/* page-lock / release caller memory (e.g. a mapped output file: payload slabs and decoded PCM are then
 * DMA-ed straight to their final place, the CLI fast path of src/main.cpp:287-311).  Registration can
 * fail for file mappings the kernel will not pin; the pointer still works unregistered, only slower. */
int lacb_host_register(lacb_ctx* ctx, void* p, uint64_t bytes);
int lacb_host_unregister(lacb_ctx* ctx, void* p);
int lacb_memcpy_h2d(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int lacb_memcpy_d2h(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int lacb_memcpy_d2d(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes);

/* Decode n_blocks blocks whose concatenated payloads start at `payload` (HOST memory).
 * block_sizes / block_bytes come from the v3 block table; block_bytes == NULL selects the
 * legacy v2 layout (no per-block byte sizes: the payload is walked as one serial chain,
 * src/codec/lac/decoder.cpp:209-218).  Output (HOST memory):
 *   LACB_PLANAR_I32 -> out_a / out_b int32 planes (out_b NULL for mono)
 *   LACB_PACKED_LE  -> out_a = interleaved little-endian bit_depth/8-byte samples
 */
int lacb_decode(lacb_ctx* ctx, const lacb_dec_params* prm, const uint8_t* payload, uint64_t payload_bytes,
                const uint32_t* block_sizes, const uint32_t* block_bytes, uint32_t n_blocks, int layout,
                void* out_a, void* out_b, lacb_err* err);

/* Device-resident variant: payload and the two tables are device pointers, the planes
 * are written to d_left / d_right (device; NULL = keep them in the context workspace),
 * optional packed output to d_packed.  d_payload must be readable up to payload_bytes
 * rounded up to a multiple of 16 (the bit reader loads aligned 32-bit words).  Output contents are
 * undefined when a decode entry point returns LACB_EDECODE (slices that finished before the failing one
 * have been written). */
int lacb_decode_device(lacb_ctx* ctx, const lacb_dec_params* prm, const uint8_t* d_payload, uint64_t payload_bytes,
                       const uint32_t* block_sizes_host, const uint32_t* block_bytes_host, uint32_t n_blocks,
                       int32_t* d_left, int32_t* d_right, uint8_t* d_packed, lacb_err* err);

/* Block-level hooks (what the reference's own unit tests exercise). */
int lacb_encode_block(lacb_ctx* ctx, const int32_t* pcm, uint32_t n, int zero_run, int partitioning,
                      uint8_t** out, uint64_t* out_size);
/* returns 1 when accepted (bits_consumed set), 0 when rejected, <0 on failure */
int lacb_decode_block(lacb_ctx* ctx, const uint8_t* data, uint64_t size, uint32_t block_size, int32_t* out,
                      uint64_t* bits_consumed);
/* Same, starting `bit_offset` bits into `data` (Block::Decoder::decode_into reads from wherever the
 * BitReader stands, src/codec/block/decoder.cpp:64); *ran_out = 1 when the rejection was "data ran out"
 * (the reference's reader is then in its error state, bitstream/bit_reader.hpp:40-60) rather than a
 * semantic reject. */
int lacb_decode_block_at(lacb_ctx* ctx, const uint8_t* data, uint64_t size, uint64_t bit_offset, uint32_t block_size,
                         int32_t* out, uint64_t* bits_consumed, int* ran_out);
/* coeffs_out[0..order]; returns used_order (0 = unstable) or <0 */
int lacb_lpc_analyze(lacb_ctx* ctx, const int32_t* pcm, uint32_t n, int order, int16_t* coeffs_out);

/* Debug: decision record of the last lacb_encode_block call. */
typedef struct lacb_block_info {
  uint32_t predictor_type, order, partition_order, n_parts, taps;
  int16_t coeffs[13];
  uint8_t part_mode[256], part_k[256];
  uint32_t cand_best_lo[11]; /* low 32 bits of each candidate's best cost; 0xFFFFFFFF = not applicable,
                                0xFFFFFFFE = dropped early: its lower bound exceeded the best cost so far */
  uint32_t bits;
} lacb_block_info;
int lacb_last_block_info(lacb_ctx* ctx, lacb_block_info* info);

/* Debug: the decisions of every block of the last lacb_encode / lacb_encode_to / lacb_encode_device call on this
 * context -- what the reference's Debug build logs with --debug-lpc, --debug-zr, --debug-partitions and
 * --debug-stereo-est (src/codec/block/encoder.cpp:457-466, 527-551, 824-835; src/codec/lac/encoder.cpp:356-379).
 * The records live in the context's workspace, so the call must have run as one pass: lacb_set_concurrency(ctx, 1)
 * before a host-buffer encode (the sliced pipeline reuses its workspaces and keeps the last slice only; the call
 * then fails with LACB_EINVAL).  ch[0] / ch[1] are the channel-blocks as emitted (left / right or mid / side). */
enum { LACB_DEC_MS = 1, LACB_DEC_UNCERTAIN = 2, LACB_DEC_PROBED = 4, LACB_DEC_BOTH = 8 };
typedef struct lacb_chan_decision {
  uint8_t predictor_type, order, partition_order, taps;
  uint8_t base_mode, has_run, pad[2];
  uint32_t bits, bytes;                                        /* exact emitted size */
  uint64_t est_bits, rice_bits, zr_bits, bin_bits, static_bits; /* the winning predictor's estimates */
  uint32_t level_bits[9];                                       /* total of partition level p; [0] = unpartitioned; 0 = not evaluated */
} lacb_chan_decision;
typedef struct lacb_block_decision {
  uint32_t flags;      /* LACB_DEC_* */
  uint32_t block_size; /* frames */
  lacb_chan_decision ch[2];
} lacb_block_decision;
int lacb_last_encode_decisions(lacb_ctx* ctx, lacb_block_decision* out, uint32_t capacity, uint32_t* n_blocks);

#ifdef __cplusplus
}
#endif
#endif /* LAC_B200_H */
