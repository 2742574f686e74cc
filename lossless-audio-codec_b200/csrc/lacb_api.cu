// lacb_api.cu -- C ABI (include/lac_b200.h) over the sm_100a kernels.
//
// Host side of the device boundary: workspace management, stage launches on the
// context's stream, and the two host<->device copies.  No codec arithmetic happens
// on the host and there is no CPU fallback.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "../../include/lac_b200.h"
#include "lacb_dec_kernels.cuh"
#include "lacb_enc_kernels.cuh"

using namespace lacb;

namespace {
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};
constexpr int FULL_NT = 1024, FULL_E = 16;  // 16384-sample channel-blocks: one CTA, 16 samples per thread
constexpr int PROBE_NT = 32, PROBE_E = 8;   // 256-sample stereo probes: one warp
constexpr size_t kFullSmem = ASmem<FULL_NT, FULL_E>::BYTES;
constexpr size_t kProbeSmem = ASmem<PROBE_NT, PROBE_E>::BYTES;
constexpr uint32_t kProbeGang = 32;  // probe warps per CTA
enum { EV_START = 0, EV_H2D, EV_PREP, EV_STEREO, EV_LPC, EV_ANALYZE, EV_FINAL, EV_EMIT, EV_D2H, EV_COUNT };
}  // namespace

struct lacb_ctx {
  int device = 0;
  int sms = 1;
  cudaStream_t stream = nullptr;
  std::string err;
  lacb_timing timing{};
  cudaEvent_t ev[EV_COUNT] = {};
  DevBuf packed_in, planeL, planeR, flags, jobs, jobs_p, counts, acor, acor_p, lpcq, lpcq_p, recs, probe_bytes,
      blk_bytes, blk_off, misc, payload;
  DevBuf d_payload, d_fs, d_size, d_boff, d_bytes, d_err, d_ms, d_L, d_R, d_packed, d_hdrs, d_order;
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  uint32_t* hmisc = nullptr;        // pinned: range error, size error, total bytes (2 words)
  cudaEvent_t ev_misc = nullptr;    // hmisc has landed
  std::vector<lacb_ctx*> kids;      // slice contexts of the pipelined host paths (own stream + workspace)
  lacb_block_info last_info{};
  // what lacb_last_encode_decisions reads: the analysis that last ran in THIS context's workspace
  uint32_t last_enc_blocks = 0, last_enc_channels = 0;
  uint64_t last_enc_frames = 0;
  uint32_t max_streams = 0;         // lacb_set_concurrency: 0 = automatic, n = at most n slices in flight
  // Slice contexts of the host pipelines launch the analysis / emit kernels with one CTA per job instead of one
  // persistent CTA per SM.  A persistent grid of slice i + 1 takes every SM the moment the grid of slice i drains, and
  // holds all of them to its end: the one-CTA kernels of slice i that follow its analysis (block sizes -> offsets, which
  // the host waits for before it can queue the emit kernel, the payload copy and the NEXT slice's input copy) then sat
  // behind the whole analysis of slice i + 1 (LACB_TRACE: "sizes known" arrived in pairs, ~2 ms of idle SMs per large
  // slice).  With short-lived CTAs an SM frees up every few microseconds and the older grid's kernels get it.
  bool per_job_grid = false;
  // ... and run them on a second stream of lower priority than everything else of the slice (copies, de-interleave,
  // stereo decision, job lists, autocorrelation, Levinson, block offsets): whenever an SM frees up, the small kernels of
  // this and of the NEXT slice go first, so they finish under the running analysis instead of queueing behind it with
  // their launch and dependency latencies exposed (6 slices x ~0.3 ms), and a slice's sizes reach the host the moment
  // its analysis ends.
  // The emit kernel sits in between: grids of one priority are served in launch order, and the analyses of the next
  // slices were launched before this slice's sizes were known -- at their priority every emit kernel would run after
  // the LAST analysis and the payload copies would pile up behind the pipeline instead of hiding under it.
  cudaStream_t stream_heavy = nullptr, stream_mid = nullptr;
  cudaEvent_t ev_x = nullptr;
};

namespace {

#define CK(call)                                                     \
  do {                                                               \
    cudaError_t _e = (call);                                         \
    if (_e != cudaSuccess) {                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e); \
      return LACB_ECUDA;                                             \
    }                                                                \
  } while (0)
#define CKR(expr)            \
  do {                       \
    const int _r = (expr);   \
    if (_r != 0) return _r;  \
  } while (0)

int ensure(lacb_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}
void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}
int ensure_pinned(lacb_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap) return 0;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_cap = 0;
  CK(cudaMallocHost(&ctx->pinned, bytes + 256));
  ctx->pinned_cap = bytes + 256;
  return 0;
}
void set_err(lacb_err* e, int code, uint32_t block, uint32_t reason, const char* msg) {
  if (!e) return;
  e->code = code;
  e->block_index = block;
  e->reason = reason;
  snprintf(e->msg, sizeof e->msg, "%s", msg);
}
float ev_ms(lacb_ctx* ctx, int a, int b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]);
  return ms;
}
template <typename T>
T* as(DevBuf& b) {
  return reinterpret_cast<T*>(b.p);
}
uint32_t lacb_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// First half of the device encoder (planes already on the device): everything up to the
// per-block sizes and payload offsets, whose totals are copied to pinned host memory
// asynchronously.  Nothing here waits for the device.
int encode_analysis(lacb_ctx* ctx, const lacb_enc_params* prm, const int32_t* dL, const int32_t* dR,
                    uint64_t frames, bool validate) {
  cudaStream_t st = ctx->stream;
  const uint64_t nb64 = (frames + kMaxBlock - 1) / kMaxBlock;
  const uint32_t nb = (uint32_t)nb64;
  EncCfg cfg;
  cfg.channels = prm->channels;
  cfg.stereo_mode = prm->channels == 2 ? prm->stereo_mode : 0u;  // lac/encoder.cpp:247
  cfg.zero_run = prm->zero_run_enabled;
  cfg.partitioning = prm->partitioning_enabled;
  cfg.n_blocks = nb;
  PcmSrc src{dL, prm->channels == 2 ? dR : nullptr, frames};
  const bool automs = cfg.stereo_mode == 2u;
  ctx->last_enc_blocks = nb;
  ctx->last_enc_channels = prm->channels;
  ctx->last_enc_frames = frames;

  CKR(ensure(ctx, ctx->flags, (size_t)nb * 4));
  CKR(ensure(ctx, ctx->jobs, (size_t)nb * 4 * 4));
  CKR(ensure(ctx, ctx->counts, 64));
  CKR(ensure(ctx, ctx->acor, (size_t)nb * 4 * 13 * 8));
  CKR(ensure(ctx, ctx->lpcq, (size_t)nb * 4 * sizeof(LpcQ)));
  CKR(ensure(ctx, ctx->recs, (size_t)nb * 4 * sizeof(ChanRec)));
  CKR(ensure(ctx, ctx->blk_bytes, (size_t)nb * 4));
  CKR(ensure(ctx, ctx->blk_off, (size_t)nb * 8));
  CKR(ensure(ctx, ctx->misc, 64));
  if (automs) {
    CKR(ensure(ctx, ctx->jobs_p, (size_t)nb * 12 * 4));
    CKR(ensure(ctx, ctx->acor_p, (size_t)nb * 12 * 13 * 8));
    CKR(ensure(ctx, ctx->lpcq_p, (size_t)nb * 12 * sizeof(LpcQ)));
    CKR(ensure(ctx, ctx->probe_bytes, (size_t)nb * 12 * 4));
  }
  uint32_t* counts = as<uint32_t>(ctx->counts);  // [0] full jobs, [1] probe jobs
  uint32_t* miscw = as<uint32_t>(ctx->misc);     // [0] range error, [1] size error, [2..3] total bytes
  CK(cudaMemsetAsync(ctx->misc.p, 0, 64, st));
  CK(cudaMemsetAsync(ctx->counts.p, 0, 64, st));

  const uint32_t wide = lacb_umin((uint32_t)((frames + 255) / 256), (uint32_t)ctx->sms * 8u);
  if (validate) {
    auto kv = k_validate;
    LACB_LAUNCH(kv, wide ? wide : 1u, 256, 0, st, src, prm->bit_depth, miscw);
  }
  CK(cudaEventRecord(ctx->ev[EV_PREP], st));

  // stereo decision (mode 2): proxy, then 3x256-sample probes for the uncertain blocks
  if (automs) {
    auto kp = k_stereo_proxy;
    LACB_LAUNCH(kp, lacb_umin(nb, (uint32_t)ctx->sms * 8u), 256, 0, st, src, cfg, as<uint32_t>(ctx->flags));
    auto kb = k_build_jobs<true>;
    LACB_LAUNCH(kb, 1, 1024, 0, st, cfg, as<uint32_t>(ctx->flags), as<uint32_t>(ctx->jobs_p), counts + 1);
    const uint32_t pgrid = lacb_umin(nb * 12u, (uint32_t)ctx->sms * 32u);
    auto ka = k_autocorr<PROBE_NT, PROBE_E, true>;
    LACB_LAUNCH(ka, pgrid, PROBE_NT, PROBE_NT * PROBE_E * 4, st, src, as<uint32_t>(ctx->jobs_p), counts + 1,
                as<i64>(ctx->acor_p));
    auto kl = k_levinson<true>;
    LACB_LAUNCH(kl, lacb_umin((nb * 12u + 63u) / 64u, (uint32_t)ctx->sms * 4u), 64, 0, st, src, as<uint32_t>(ctx->jobs_p),
                counts + 1, as<i64>(ctx->acor_p), as<LpcQ>(ctx->lpcq_p));
    // probe warps run in gangs of kProbeGang per CTA (see k_analyze)
    static const uint32_t forced_gang = [] {  // experiment knob, 1..32 sub-blocks per CTA
      const char* e = getenv("LACB_PROBE_GANG");
      const int v = e ? atoi(e) : 0;
      return (uint32_t)(v >= 1 && v <= 32 ? v : 0);
    }();
    // full gangs when every SM gets 32 probes or more; a short input spreads its probes over all SMs in smaller gangs
    uint32_t gang = kProbeGang;
    while (gang > 4u && nb * 12u <= (uint32_t)ctx->sms * (gang / 2u)) gang /= 2u;
    if (forced_gang) gang = forced_gang;
    const uint32_t ggrid = lacb_umin((nb * 12u + gang - 1u) / gang, (uint32_t)ctx->sms * (32u / gang));
    auto kz = k_analyze<PROBE_NT, PROBE_E, true>;
    LACB_LAUNCH(kz, ggrid, PROBE_NT * gang, kProbeSmem * gang, st, src, cfg, as<uint32_t>(ctx->jobs_p),
                counts + 1, as<LpcQ>(ctx->lpcq_p), (ChanRec*)nullptr, as<uint32_t>(ctx->probe_bytes));
    auto kd = k_decide_probes;
    LACB_LAUNCH(kd, lacb_umin((nb + 255u) / 256u, (uint32_t)ctx->sms), 256, 0, st, cfg, as<uint32_t>(ctx->flags),
                as<uint32_t>(ctx->probe_bytes));
  } else {
    auto kf = k_plan_fixed;
    LACB_LAUNCH(kf, lacb_umin((nb + 255u) / 256u, (uint32_t)ctx->sms), 256, 0, st, cfg, as<uint32_t>(ctx->flags));
  }
  CK(cudaEventRecord(ctx->ev[EV_STEREO], st));

  // LPC analysis and the channel-block search on the selected channels
  {
    auto kb = k_build_jobs<false>;
    LACB_LAUNCH(kb, 1, 1024, 0, st, cfg, as<uint32_t>(ctx->flags), as<uint32_t>(ctx->jobs), counts);
    const uint32_t fgrid = ctx->per_job_grid ? nb * 4u : lacb_umin(nb * 4u, (uint32_t)ctx->sms);
    auto ka = k_autocorr_stream<256>;
    LACB_LAUNCH(ka, lacb_umin(nb * 4u, (uint32_t)ctx->sms * 2u), 256, 0, st, src, as<uint32_t>(ctx->jobs), counts,
                as<i64>(ctx->acor));
    auto kl = k_levinson<false>;
    LACB_LAUNCH(kl, lacb_umin((nb * 4u + 63u) / 64u, (uint32_t)ctx->sms * 4u), 64, 0, st, src, as<uint32_t>(ctx->jobs),
                counts, as<i64>(ctx->acor), as<LpcQ>(ctx->lpcq));
    CK(cudaEventRecord(ctx->ev[EV_LPC], st));
    cudaStream_t sh = ctx->stream_heavy ? ctx->stream_heavy : st;
    if (sh != st) CK(cudaStreamWaitEvent(sh, ctx->ev[EV_LPC], 0));
    auto kz = k_analyze<FULL_NT, FULL_E, false>;
    LACB_LAUNCH(kz, fgrid, FULL_NT, kFullSmem, sh, src, cfg, as<uint32_t>(ctx->jobs), counts,
                as<LpcQ>(ctx->lpcq), as<ChanRec>(ctx->recs), (uint32_t*)nullptr);
    if (sh != st) {
      CK(cudaEventRecord(ctx->ev_x, sh));
      CK(cudaStreamWaitEvent(st, ctx->ev_x, 0));
    }
  }
  CK(cudaEventRecord(ctx->ev[EV_ANALYZE], st));

  // block sizes, final LR/MS pick of "encode both" blocks, payload offsets
  {
    auto kf = k_finalize_blocks;
    LACB_LAUNCH(kf, 1, 1024, 0, st, cfg, as<uint32_t>(ctx->flags), as<ChanRec>(ctx->recs),
                as<uint32_t>(ctx->blk_bytes), as<u64>(ctx->blk_off), reinterpret_cast<u64*>(miscw + 2), miscw + 1);
  }
  CK(cudaEventRecord(ctx->ev[EV_FINAL], st));
  if (!ctx->hmisc) {
    CK(cudaMallocHost(reinterpret_cast<void**>(&ctx->hmisc), 64));
    CK(cudaEventCreateWithFlags(&ctx->ev_misc, cudaEventDisableTiming));
  }
  CK(cudaMemcpyAsync(ctx->hmisc, ctx->misc.p, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev_misc, st));
  return 0;
}

// Second half of the device encoder: waits for the sizes of encode_analysis, sizes the payload
// buffer and launches the emitter.
int encode_emit(lacb_ctx* ctx, const lacb_enc_params* prm, const int32_t* dL, const int32_t* dR, uint64_t frames,
                uint64_t* total, lacb_err* err) {
  cudaStream_t st = ctx->stream;
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  EncCfg cfg;
  cfg.channels = prm->channels;
  cfg.stereo_mode = prm->channels == 2 ? prm->stereo_mode : 0u;
  cfg.zero_run = prm->zero_run_enabled;
  cfg.partitioning = prm->partitioning_enabled;
  cfg.n_blocks = nb;
  PcmSrc src{dL, prm->channels == 2 ? dR : nullptr, frames};
  const uint32_t* hmisc = ctx->hmisc;
  CK(cudaEventSynchronize(ctx->ev_misc));
  CK(cudaGetLastError());
  if (hmisc[0]) {
    ctx->err = "sample outside bit depth range";
    set_err(err, LACB_EINVAL, 0, 0, "sample outside bit depth range");
    return LACB_EINVAL;
  }
  if (hmisc[1]) {
    ctx->err = "encoded block size is outside format limits";
    set_err(err, LACB_ELIMIT, 0, 0, "encoded block size is outside format limits");
    return LACB_ELIMIT;
  }
  uint64_t tot;
  memcpy(&tot, hmisc + 2, 8);
  *total = tot;
  CKR(ensure(ctx, ctx->payload, (size_t)tot + 64));
  {
    // (everything queued on st so far has completed: the host has just read this slice's sizes)
    cudaStream_t sh = ctx->stream_mid ? ctx->stream_mid : st;
    auto ke = k_emit<FULL_NT, FULL_E>;
    LACB_LAUNCH(ke, ctx->per_job_grid ? nb * cfg.channels : lacb_umin(nb * cfg.channels, (uint32_t)ctx->sms), FULL_NT,
                kFullSmem, sh, src, cfg,
                as<uint32_t>(ctx->flags), as<ChanRec>(ctx->recs), as<u64>(ctx->blk_off), as<uint8_t>(ctx->payload));
    if (sh != st) {
      CK(cudaEventRecord(ctx->ev_x, sh));
      CK(cudaStreamWaitEvent(st, ctx->ev_x, 0));
    }
  }
  CK(cudaEventRecord(ctx->ev[EV_EMIT], st));
  return 0;
}

// Device part of the encoder: planes already on the device.  On success the payload is
// in ctx->payload, the per-block sizes in ctx->blk_bytes, *total the payload size.
int encode_on_device(lacb_ctx* ctx, const lacb_enc_params* prm, const int32_t* dL, const int32_t* dR,
                     uint64_t frames, bool validate, uint64_t* total, lacb_err* err) {
  CKR(encode_analysis(ctx, prm, dL, dR, frames, validate));
  return encode_emit(ctx, prm, dL, dR, frames, total, err);
}

void fill_enc_timing(lacb_ctx* ctx) {
  lacb_timing& t = ctx->timing;
  memset(&t, 0, sizeof t);
  t.h2d_ms = ev_ms(ctx, EV_START, EV_H2D);
  t.prep_ms = ev_ms(ctx, EV_H2D, EV_PREP);
  t.stereo_ms = ev_ms(ctx, EV_PREP, EV_STEREO);
  t.lpc_ms = ev_ms(ctx, EV_STEREO, EV_LPC);
  t.analyze_ms = ev_ms(ctx, EV_LPC, EV_ANALYZE);
  t.finalize_ms = ev_ms(ctx, EV_ANALYZE, EV_FINAL);
  t.emit_ms = ev_ms(ctx, EV_FINAL, EV_EMIT);
  t.d2h_ms = ev_ms(ctx, EV_EMIT, EV_D2H);
  t.total_ms = ev_ms(ctx, EV_START, EV_D2H);
}

bool params_ok(const lacb_enc_params* p) {
  return p && (p->channels == 1 || p->channels == 2) && (p->bit_depth == 16 || p->bit_depth == 24) &&
         p->stereo_mode <= 2;
}

__global__ void k_decode_one(const uint8_t* data, u64 size, u64 padded, u64 bit_off, uint32_t n, int32_t* out,
                             u64* result) {
  __shared__ uint32_t ring[kDriftWin];
  __shared__ ChanHdr hdr;
  __shared__ ParseScratch sc;
  __shared__ int4 stage1[2 * kRestoreDepth];
  const uint32_t lane = threadIdx.x & 31u;
  BitRd r;
  rd_init(r, data, size, data + padded);
  if (bit_off) rd_seek(r, r.start + bit_off);  // Block::Decoder::decode_into starts wherever the reader stands
  bool ok = parse_channel_block(r, n, out, &hdr, ring, &sc, lane);
  __syncwarp();
  if (lane != 0u) return;
  const bool ran_out = !ok && rd_over(r);
  if (ok) ok = restore_block(out, n, hdr.type, hdr.order, hdr.coef, stage1, 1u, true);
  result[0] = ok ? 1ull : 0ull;
  result[1] = ok ? rd_pos(r) - r.start - bit_off : 0ull;
  result[2] = ran_out ? 1ull : 0ull;
}

}  // namespace

extern "C" {

int lacb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int lacb_create(int device, lacb_ctx** out) {
  if (!out) return LACB_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return LACB_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return LACB_ECUDA;
  lacb_ctx* ctx = new lacb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return LACB_ECUDA;
  }
  ctx->sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 1;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return LACB_ECUDA;
  }
  for (int i = 0; i < EV_COUNT; ++i) cudaEventCreate(&ctx->ev[i]);
  cudaFuncSetAttribute(k_analyze<FULL_NT, FULL_E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)ASmem<FULL_NT, FULL_E>::BYTES);
  cudaFuncSetAttribute(k_analyze<PROBE_NT, PROBE_E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)(kProbeSmem * 32));
  // The decoder's kernels of different slices share SMs in the host pipeline (a restore CTA next to the parser warps
  // of later slices).  An SM's shared-memory carve-out is chosen from the preference of the kernel that first lands on
  // it and can only change when the SM is empty: with the default preference the parsers (2 KB each) left the SMs
  // configured too small for a restore CTA's 12 KB of staging, and the restore kernels of the early slices waited
  // until the parsers of ALL slices had retired (LACB_TRACE: restore of slice 0 done at 8.2 instead of 3.7 ms, decode
  // 13.7 instead of 10.4 ms) -- or not, depending on the build, because the driver's choice follows the kernels'
  // resource use.  All of them ask for the largest carve-out.
  cudaFuncSetAttribute(k_parse_blocks, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_parse_serial, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_restore_order, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_restore_blocks, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_merge_restore_errors, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_finish_pcm, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_emit<FULL_NT, FULL_E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)ASmem<FULL_NT, FULL_E>::BYTES);
  if (cudaGetLastError() != cudaSuccess) {
    lacb_destroy(ctx);
    return LACB_ECUDA;
  }
  *out = ctx;
  return LACB_OK;
}

void lacb_destroy(lacb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf* all[] = {&ctx->packed_in, &ctx->planeL, &ctx->planeR, &ctx->flags, &ctx->jobs, &ctx->jobs_p, &ctx->counts,
                   &ctx->acor, &ctx->acor_p, &ctx->lpcq, &ctx->lpcq_p, &ctx->recs, &ctx->probe_bytes,
                   &ctx->blk_bytes, &ctx->blk_off, &ctx->misc, &ctx->payload, &ctx->d_payload, &ctx->d_fs,
                   &ctx->d_size, &ctx->d_boff, &ctx->d_bytes, &ctx->d_err, &ctx->d_ms, &ctx->d_L, &ctx->d_R,
                   &ctx->d_packed, &ctx->d_hdrs, &ctx->d_order};
  for (DevBuf* b : all) release(*b);
  for (lacb_ctx* k : ctx->kids) lacb_destroy(k);
  if (ctx->hmisc) cudaFreeHost(ctx->hmisc);
  if (ctx->ev_misc) cudaEventDestroy(ctx->ev_misc);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < EV_COUNT; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->stream_heavy) cudaStreamDestroy(ctx->stream_heavy);
  if (ctx->stream_mid) cudaStreamDestroy(ctx->stream_mid);
  if (ctx->ev_x) cudaEventDestroy(ctx->ev_x);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lacb_last_error(const lacb_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
void lacb_free(void* p) { free(p); }
#if defined(LACB_PHASE_CLK) && !defined(LACB_EMU)
// debug builds only (tools/phase_clk.py): copies out and optionally clears the phase clock table
int lacb_debug_phase_clk(unsigned long long* out64, int reset) {
  cudaDeviceSynchronize();
  if (out64 && cudaMemcpyFromSymbol(out64, lacb::g_phase_clk, 160 * sizeof(unsigned long long)) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[160] = {0};
    cudaMemcpyToSymbol(lacb::g_phase_clk, z, sizeof z);
  }
  return 0;
}
#endif

int lacb_get_timing(const lacb_ctx* ctx, lacb_timing* out) {
  if (!ctx || !out) return LACB_EINVAL;
  *out = ctx->timing;
  return 0;
}

int lacb_encode_device(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* d_a, const void* d_b,
                       uint64_t frames, const uint8_t** d_payload, uint64_t* payload_bytes,
                       const uint32_t** d_block_bytes, lacb_err* err) {
  if (!ctx) return LACB_EINVAL;
  if (!params_ok(prm) || !d_a || frames == 0 || (layout == LACB_PLANAR_I32 && prm->channels == 2 && !d_b) ||
      (layout != LACB_PLANAR_I32 && layout != LACB_PACKED_LE) || (frames + kMaxBlock - 1) / kMaxBlock > 0xFFFFFFu) {
    ctx->err = "invalid encode arguments";
    set_err(err, LACB_EINVAL, 0, 0, "invalid encode arguments");
    return LACB_EINVAL;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[EV_START], st));
  CK(cudaEventRecord(ctx->ev[EV_H2D], st));
  const int32_t* dL = static_cast<const int32_t*>(d_a);
  const int32_t* dR = static_cast<const int32_t*>(d_b);
  bool validate = prm->validate_range != 0;
  if (layout == LACB_PACKED_LE) {
    CKR(ensure(ctx, ctx->planeL, frames * 4));
    if (prm->channels == 2) CKR(ensure(ctx, ctx->planeR, frames * 4));
    const uint32_t grid = lacb_umin((uint32_t)((frames + 255) / 256), (uint32_t)ctx->sms * 16u);
    auto kd = k_deinterleave;
    LACB_LAUNCH(kd, grid ? grid : 1u, 256, 0, st, static_cast<const uint8_t*>(d_a), (u64)frames, prm->channels,
                prm->bit_depth / 8u, as<int32_t>(ctx->planeL), as<int32_t>(ctx->planeR));
    dL = as<int32_t>(ctx->planeL);
    dR = as<int32_t>(ctx->planeR);
    validate = false;
  }
  uint64_t total = 0;
  CKR(encode_on_device(ctx, prm, dL, dR, frames, validate, &total, err));
  CK(cudaEventRecord(ctx->ev[EV_D2H], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  fill_enc_timing(ctx);
  if (d_payload) *d_payload = as<uint8_t>(ctx->payload);
  if (payload_bytes) *payload_bytes = total;
  if (d_block_bytes) *d_block_bytes = as<uint32_t>(ctx->blk_bytes);
  return 0;
}

int lacb_dev_malloc(lacb_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
}
int lacb_dev_free(lacb_ctx* ctx, void* p) {
  if (!ctx) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaFree(p));
  return 0;
}
int lacb_host_malloc(lacb_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMallocHost(out, bytes ? bytes : 1));
  return 0;
}
int lacb_host_free(lacb_ctx* ctx, void* p) {
  if (!ctx) return LACB_EINVAL;
  CK(cudaFreeHost(p));
  return 0;
}
// Page-locks caller memory (a mapped output file, a caller's own buffer) so that copies to / from it run at
// full PCIe speed and asynchronously.  Fails (LACB_ECUDA) where the kernel refuses to pin the pages (writable
// file mappings on most disk filesystems); callers then simply pass the pointer unregistered.
int lacb_host_register(lacb_ctx* ctx, void* p, uint64_t bytes) {
  if (!ctx || !p || !bytes) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();  // not sticky: clear it
    ctx->err = std::string("cudaHostRegister: ") + cudaGetErrorString(e);
    return LACB_ECUDA;
  }
  return 0;
}
int lacb_host_unregister(lacb_ctx* ctx, void* p) {
  if (!ctx || !p) return LACB_EINVAL;
  CK(cudaHostUnregister(p));
  return 0;
}
int lacb_memcpy_h2d(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int lacb_memcpy_d2h(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int lacb_memcpy_d2d(lacb_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---------------------------------------------------------------------------
// Pipelined host paths.  A long input is cut into slices of whole blocks that run on slice contexts
// (own stream, own workspace), so the host->device copy of one slice and the device->host copy of
// another run under the kernels of a third.  Blocks are independent, so the bytes are those of a
// single pass.  The encoder alternates between two contexts and sizes its slices as a multiple of
// the SM count of channel-block jobs (the persistent analysis grid then ends on a full wave); the
// decoder keeps dec_kids() slices in flight (see lacb_decode).
static uint32_t slice_blocks(const lacb_ctx* ctx, uint32_t channels) {
  static const long forced = getenv("LACB_SLICE_BLOCKS") ? atol(getenv("LACB_SLICE_BLOCKS")) : 0;  // tests / tuning
  if (forced > 0) return (uint32_t)forced;
  const uint32_t per_wave = ((uint32_t)ctx->sms + channels - 1u) / channels;  // blocks per wave of jobs
  return per_wave * 14u;
}
// slice contexts the host-buffer decode keeps in flight
// Four by default: the CUDA runtime maps streams onto 8 hardware queues unless CUDA_DEVICE_MAX_CONNECTIONS says
// otherwise, and slices whose streams share a queue serialise (6 - 16 slices in flight: 13 - 19 ms instead of 11.3 ms
// for 600 s of 24/96).  A process that raised the queue count before CUDA started (`lac_cli serve` and bench.py set 32) gets
// twelve: the first slice is a twelfth of the input, so the device-to-host copies start earlier (10.5 ms).
static uint32_t dec_kids() {
  static const long forced = getenv("LACB_DEC_KIDS") ? atol(getenv("LACB_DEC_KIDS")) : 0;  // tuning
  static const long conns = getenv("CUDA_DEVICE_MAX_CONNECTIONS") ? atol(getenv("CUDA_DEVICE_MAX_CONNECTIONS")) : 8;
  return forced >= 2 && forced <= 16 ? (uint32_t)forced : (conns >= 16 ? 12u : 4u);
}
static double trace_now() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static bool trace_on() {
  static const bool on = getenv("LACB_TRACE") != nullptr;
  return on;
}
#define TRACE(...)                                   \
  do {                                               \
    if (trace_on()) {                                \
      fprintf(stderr, "[lacb %.3f] ", trace_now());  \
      fprintf(stderr, __VA_ARGS__);                  \
      fprintf(stderr, "\n");                         \
    }                                                \
  } while (0)
static int ensure_kids(lacb_ctx* ctx, size_t count = 2) {
  while (ctx->kids.size() < count) {
    lacb_ctx* k = nullptr;
    const int rc = lacb_create(ctx->device, &k);
    if (rc != 0) {
      ctx->err = "cannot create slice context";
      return rc;
    }
    static const bool persistent = getenv("LACB_SLICE_PERSISTENT") != nullptr;  // A/B knobs
    static const bool one_prio = getenv("LACB_SLICE_ONE_PRIORITY") != nullptr;
    k->per_job_grid = !persistent;
    if (!persistent && !one_prio) {
      // the slice's own stream becomes a high-priority one, the heavy kernels get a stream of the lowest priority
      int least = 0, greatest = 0;
      cudaDeviceGetStreamPriorityRange(&least, &greatest);
      cudaStream_t hi = nullptr;
      if (greatest != least && cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, greatest) == cudaSuccess &&
          cudaStreamCreateWithPriority(&k->stream_heavy, cudaStreamNonBlocking, least) == cudaSuccess &&
          cudaStreamCreateWithPriority(&k->stream_mid, cudaStreamNonBlocking, (least + greatest) / 2) == cudaSuccess &&
          cudaEventCreateWithFlags(&k->ev_x, cudaEventDisableTiming) == cudaSuccess) {
        cudaStreamDestroy(k->stream);
        k->stream = hi;
      } else {
        if (hi) cudaStreamDestroy(hi);
        if (k->stream_heavy) cudaStreamDestroy(k->stream_heavy);
        if (k->stream_mid) cudaStreamDestroy(k->stream_mid);
        k->stream_heavy = k->stream_mid = nullptr;
        cudaGetLastError();
      }
    }
    ctx->kids.push_back(k);
  }
  return 0;
}
#define CKK(kid, expr)                \
  do {                                \
    const int _r = (expr);            \
    if (_r != 0) {                    \
      ctx->err = (kid)->err;          \
      cudaDeviceSynchronize();        \
      return _r;                      \
    }                                 \
  } while (0)

// host -> device copy of one slice plus the first half of its encode, nothing waits
static int enc_slice_begin(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a,
                           const void* pcm_b, uint64_t f0, uint64_t fr) {
  cudaStream_t st = ctx->stream;
  CKR(ensure(ctx, ctx->planeL, fr * 4));
  if (prm->channels == 2) CKR(ensure(ctx, ctx->planeR, fr * 4));
  bool validate = prm->validate_range != 0;
  CK(cudaEventRecord(ctx->ev[EV_START], st));
  if (layout == LACB_PLANAR_I32) {
    CK(cudaMemcpyAsync(ctx->planeL.p, static_cast<const int32_t*>(pcm_a) + f0, fr * 4, cudaMemcpyHostToDevice, st));
    if (prm->channels == 2)
      CK(cudaMemcpyAsync(ctx->planeR.p, static_cast<const int32_t*>(pcm_b) + f0, fr * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[EV_H2D], st));
  } else {
    const uint32_t bps = prm->bit_depth / 8;
    const size_t fb = (size_t)prm->channels * bps;
    CKR(ensure(ctx, ctx->packed_in, fr * fb + 4));
    CK(cudaMemcpyAsync(ctx->packed_in.p, static_cast<const uint8_t*>(pcm_a) + f0 * fb, fr * fb, cudaMemcpyHostToDevice,
                       st));
    CK(cudaEventRecord(ctx->ev[EV_H2D], st));
    const uint32_t grid = lacb_umin((uint32_t)((fr + 255) / 256), (uint32_t)ctx->sms * 16u);
    auto kd = k_deinterleave;
    LACB_LAUNCH(kd, grid ? grid : 1u, 256, 0, st, as<uint8_t>(ctx->packed_in), (u64)fr, prm->channels, bps,
                as<int32_t>(ctx->planeL), as<int32_t>(ctx->planeR));
    validate = false;
  }
  return encode_analysis(ctx, prm, as<int32_t>(ctx->planeL), as<int32_t>(ctx->planeR), fr, validate);
}

static int encode_host_sliced(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a,
                              const void* pcm_b, uint64_t frames, uint8_t* dst, uint64_t dst_cap,
                              uint8_t** payload_out, uint64_t* payload_bytes, uint32_t* block_bytes, lacb_err* err) {
  // Three slice contexts in rotation.  With two, the copy of slice i + 1 sat on the stream of slice i - 1 behind that
  // slice's emit kernel, which cannot start before the analysis of slice i (all SMs, all registers) has finished: the
  // copy therefore ran AFTER the analysis it was meant to hide under, ~2 ms of idle SMs per large slice.  The third
  // context's stream is free when slice i + 1 is queued (its last emit only waited for the analysis of slice i - 1).
  // Round 2, later: FOUR contexts, and the input of slice i + 2 is queued before the host waits for the sizes of slice
  // i -- with a look-ahead of one slice the copy of slice i + 1 only started when the sizes of slice i - 1 were known,
  // i.e. when its own analysis could already have begun (1.1 ms of idle SMs in front of the third slice).
  static const uint32_t NK = [] {  // tuning knob, 2..8 contexts
    const int v = getenv("LACB_ENC_KIDS") ? atoi(getenv("LACB_ENC_KIDS")) : 4;
    return (uint32_t)(v < 2 ? 2 : (v > 8 ? 8 : v));
  }();
  static const uint32_t AHEAD = [] {  // tuning knob, slices queued ahead (capped by the contexts below)
    const int v = getenv("LACB_ENC_AHEAD") ? atoi(getenv("LACB_ENC_AHEAD")) : 2;
    return (uint32_t)(v < 1 ? 1 : (v > 7 ? 7 : v));
  }();
  const uint32_t nkenc = ctx->max_streams >= 2u ? lacb_umin(NK, ctx->max_streams) : NK;  // the caller's cap on slices in flight
  const uint32_t ahead = lacb_umin(AHEAD ? AHEAD : 1u, nkenc - 1u);  // slices queued beyond the one being collected
  CKR(ensure_kids(ctx, nkenc));
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  // Slice plan: the first slices are short (2, 4, 8 waves of jobs) so that the kernels start after a
  // fraction of a millisecond of copying and every later copy hides under the slice before it.
  const uint32_t sb = slice_blocks(ctx, prm->channels);
  // The plan ends the way it starts: what follows the last analysis -- that slice's emit kernel and payload copy --
  // is exposed, so the last two slices are 4 and 2 waves.
  std::vector<uint32_t> cut{0u};  // cut[i] = first block of slice i, cut[ns] = nb
  const uint32_t wave = sb / 14u;
  uint32_t main_end = nb;  // blocks in front of the short closing slices
  if (getenv("LACB_SLICE_BLOCKS") == nullptr) {
    for (uint32_t w = 2u; w < 14u && cut.back() + wave * w + sb < nb; w *= 2u) cut.push_back(cut.back() + wave * w);
    if (cut.back() + sb + 6u * wave < nb) main_end = nb - 6u * wave;
  }
  {
    // the middle in equal slices of at most sb blocks (whole waves)
    const uint32_t mid = main_end - cut.back();
    const uint32_t nmid = (mid + sb - 1u) / sb;
    const uint32_t each = nmid ? ((mid + nmid - 1u) / nmid + wave - 1u) / (wave ? wave : 1u) * (wave ? wave : 1u) : 0u;
    while (each && cut.back() + each < main_end) cut.push_back(cut.back() + each);
    if (main_end > cut.back()) cut.push_back(main_end);
  }
  if (main_end < nb) {
    cut.push_back(main_end + 4u * wave);
    cut.push_back(nb);
  }
  const uint32_t ns = (uint32_t)cut.size() - 1u;
  uint8_t* host = dst;
  uint64_t cap = dst_cap;
  if (!dst) {  // library-owned result: start from the raw size, grow if a stream expands
    cap = frames * prm->channels * (prm->bit_depth / 8) + (uint64_t)nb * 8 + 4096;
    host = (uint8_t*)malloc(cap);
    if (!host) return LACB_ENOMEM;
  }
  auto slice_range = [&](uint32_t i, uint64_t* f0, uint64_t* fr) {
    const uint64_t b0 = cut[i], b1 = cut[i + 1u];
    *f0 = b0 * kMaxBlock;
    const uint64_t f1 = b1 * kMaxBlock < frames ? b1 * kMaxBlock : frames;
    *fr = f1 - *f0;
  };
  auto fail = [&](int rc) {
    cudaDeviceSynchronize();
    if (!dst) free(host);
    return rc;
  };
  uint64_t f0, fr;
  if (trace_on()) cudaEventRecord(ctx->ev[EV_START], ctx->stream);
  uint32_t begun = 0;  // slices whose input copy and analysis are queued
  auto begin_upto = [&](uint32_t want) -> int {  // queue slices [begun, want)
    for (; begun < want && begun < ns; ++begun) {
      uint64_t g0, gr;
      slice_range(begun, &g0, &gr);
      lacb_ctx* kn = ctx->kids[begun % nkenc];
      const int rc = enc_slice_begin(kn, prm, layout, pcm_a, pcm_b, g0, gr);
      if (rc != 0) { ctx->err = kn->err; return rc; }
      TRACE("enc slice %u queued", begun);
    }
    return 0;
  };
  {
    const int rc = begin_upto(ahead);
    if (rc != 0) return fail(rc);
  }
  uint64_t off = 0;
  bool overflow = false;
  // per-block sizes land in page-locked memory first: a device->host copy into the caller's
  // pageable array would block the host at every slice and serialise the pipeline
  uint32_t* bb_stage = nullptr;
  if (block_bytes) {
    const int rc = ensure_pinned(ctx, (size_t)nb * 4);
    if (rc != 0) return fail(rc);
    bb_stage = static_cast<uint32_t*>(ctx->pinned);
  }
  for (uint32_t i = 0; i < ns; ++i) {
    lacb_ctx* k = ctx->kids[i % nkenc];
    {  // the next slices: their copies run under the analysis of slice i, their analyses queue up behind it
      const int rc = begin_upto(i + 1u + ahead);
      if (rc != 0) return fail(rc);
    }
    slice_range(i, &f0, &fr);
    uint64_t total = 0;
    const int rc = encode_emit(k, prm, as<int32_t>(k->planeL), as<int32_t>(k->planeR), fr, &total, err);
    if (rc != 0) { ctx->err = k->err; return fail(rc); }
    TRACE("enc slice %u sizes known, emit queued (%llu bytes)", i, (unsigned long long)total);
    const uint32_t nbs = (uint32_t)((fr + kMaxBlock - 1) / kMaxBlock);
    if (off + total > cap) {
      if (dst) {
        overflow = true;  // keep going without copying: the caller is told the size it needs
      } else {
        cudaDeviceSynchronize();  // copies into the old buffer must have landed
        cap = (off + total) * 2;
        uint8_t* bigger = (uint8_t*)realloc(host, cap);
        if (!bigger) return fail(LACB_ENOMEM);
        host = bigger;
      }
    }
    if (!overflow)
      CKK(k, [&]() -> int {
        lacb_ctx* ctx = k;  // CK reports into the slice context
        CK(cudaMemcpyAsync(host + off, ctx->payload.p, total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
        return 0;
      }());
    if (bb_stage)
      CKK(k, [&]() -> int {
        lacb_ctx* ctx = k;
        CK(cudaMemcpyAsync(bb_stage + cut[i], ctx->blk_bytes.p, (size_t)nbs * 4, cudaMemcpyDeviceToHost,
                           ctx->stream));
        return 0;
      }());
    off += total;
  }
  for (lacb_ctx* k : ctx->kids) {
    if (cudaStreamSynchronize(k->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      ctx->err = "CUDA failure in the sliced encoder";
      return fail(LACB_ECUDA);
    }
  }
  TRACE("enc drained");
  if (trace_on()) {  // device timeline of the last slices (a slice context keeps the events of its last slice only)
    for (uint32_t i = ns > nkenc ? ns - nkenc : 0u; i < ns; ++i) {
      lacb_ctx* k = ctx->kids[i % nkenc];
      float t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const int evs[9] = {EV_START, EV_H2D, EV_PREP, EV_STEREO, EV_LPC, EV_ANALYZE, EV_FINAL, EV_EMIT, EV_D2H};
      for (int e = 0; e < 9; ++e) cudaEventElapsedTime(&t[e], ctx->ev[EV_START], k->ev[evs[e]]);
      TRACE("enc slice %u: start %.2f h2d %.2f prep %.2f stereo %.2f lpc %.2f analyze %.2f sizes %.2f emit %.2f d2h %.2f ms", i,
            t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]);
    }
  }
  memset(&ctx->timing, 0, sizeof ctx->timing);  // stage times are per slice context, not aggregated
  if (block_bytes) memcpy(block_bytes, bb_stage, (size_t)nb * 4);
  *payload_bytes = off;
  if (overflow) {
    ctx->err = "payload buffer too small";
    return LACB_ENOMEM;
  }
  if (payload_out) *payload_out = host;
  return 0;
}

static int encode_host(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a, const void* pcm_b,
                       uint64_t frames, uint8_t* dst, uint64_t dst_cap, uint8_t** payload_out,
                       uint64_t* payload_bytes, uint32_t* block_bytes, lacb_err* err) {
  if (!ctx) return LACB_EINVAL;
  if (!params_ok(prm) || !pcm_a || frames == 0 || (!payload_out && !dst) || !payload_bytes ||
      (layout == LACB_PLANAR_I32 && prm->channels == 2 && !pcm_b) ||
      (layout != LACB_PLANAR_I32 && layout != LACB_PACKED_LE) || (frames + kMaxBlock - 1) / kMaxBlock > 0xFFFFFFu) {
    ctx->err = "invalid encode arguments";
    set_err(err, LACB_EINVAL, 0, 0, "invalid encode arguments");
    return LACB_EINVAL;
  }
  if (payload_out) *payload_out = nullptr;
  *payload_bytes = 0;
  ctx->last_enc_blocks = 0;  // the sliced pipeline analyses in the slice contexts: nothing to report from here
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t nb = (uint32_t)((frames + kMaxBlock - 1) / kMaxBlock);
  if (ctx->max_streams != 1u && nb >= 2u * slice_blocks(ctx, prm->channels))  // one stream: plain H2D, kernels, D2H
    return encode_host_sliced(ctx, prm, layout, pcm_a, pcm_b, frames, dst, dst_cap, payload_out, payload_bytes,
                              block_bytes, err);
  CKR(ensure(ctx, ctx->planeL, frames * 4));
  if (prm->channels == 2) CKR(ensure(ctx, ctx->planeR, frames * 4));
  CK(cudaEventRecord(ctx->ev[EV_START], st));
  bool validate = prm->validate_range != 0;
  if (layout == LACB_PLANAR_I32) {
    CK(cudaMemcpyAsync(ctx->planeL.p, pcm_a, frames * 4, cudaMemcpyHostToDevice, st));
    if (prm->channels == 2) CK(cudaMemcpyAsync(ctx->planeR.p, pcm_b, frames * 4, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[EV_H2D], st));
  } else {
    const uint32_t bps = prm->bit_depth / 8;
    const size_t bytes = (size_t)frames * prm->channels * bps;
    CKR(ensure(ctx, ctx->packed_in, bytes + 4));
    CK(cudaMemcpyAsync(ctx->packed_in.p, pcm_a, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(ctx->ev[EV_H2D], st));
    const uint32_t grid = lacb_umin((uint32_t)((frames + 255) / 256), (uint32_t)ctx->sms * 16u);
    auto kd = k_deinterleave;
    LACB_LAUNCH(kd, grid ? grid : 1u, 256, 0, st, as<uint8_t>(ctx->packed_in), (u64)frames, prm->channels, bps,
                as<int32_t>(ctx->planeL), as<int32_t>(ctx->planeR));
    validate = false;  // a packed sample cannot leave its own bit depth
  }
  uint64_t total = 0;
  CKR(encode_on_device(ctx, prm, as<int32_t>(ctx->planeL), as<int32_t>(ctx->planeR), frames, validate, &total, err));
  uint8_t* host = dst;
  if (dst) {
    if (total > dst_cap) {  // tell the caller how much room the payload needs
      *payload_bytes = total;
      ctx->err = "payload buffer too small";
      return LACB_ENOMEM;
    }
  } else {
    host = (uint8_t*)malloc(total ? total : 1);
    if (!host) return LACB_ENOMEM;
  }
  const int rc_copy = [&]() -> int {
    CK(cudaMemcpyAsync(host, ctx->payload.p, total, cudaMemcpyDeviceToHost, st));
    if (block_bytes) CK(cudaMemcpyAsync(block_bytes, ctx->blk_bytes.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[EV_D2H], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return 0;
  }();
  if (rc_copy != 0) {
    if (!dst) free(host);  // the library-owned result buffer does not outlive a failed call
    return rc_copy;
  }
  fill_enc_timing(ctx);
  if (payload_out) *payload_out = host;
  *payload_bytes = total;
  return 0;
}

int lacb_encode(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a, const void* pcm_b,
                uint64_t frames, uint8_t** payload_out, uint64_t* payload_bytes, uint32_t* block_bytes,
                lacb_err* err) {
  if (!payload_out) return LACB_EINVAL;
  return encode_host(ctx, prm, layout, pcm_a, pcm_b, frames, nullptr, 0, payload_out, payload_bytes, block_bytes, err);
}

int lacb_encode_to(lacb_ctx* ctx, const lacb_enc_params* prm, int layout, const void* pcm_a, const void* pcm_b,
                   uint64_t frames, uint8_t* payload_buf, uint64_t payload_cap, uint64_t* payload_bytes,
                   uint32_t* block_bytes, lacb_err* err) {
  if (!payload_buf) return LACB_EINVAL;
  return encode_host(ctx, prm, layout, pcm_a, pcm_b, frames, payload_buf, payload_cap, nullptr, payload_bytes,
                     block_bytes, err);
}

int lacb_encode_block(lacb_ctx* ctx, const int32_t* pcm, uint32_t n, int zero_run, int partitioning, uint8_t** out,
                      uint64_t* out_size) {
  if (!ctx || !pcm || n == 0 || n > kMaxBlock || !out || !out_size) return LACB_EINVAL;
  lacb_enc_params prm{};
  prm.sample_rate = 44100;
  prm.bit_depth = 24;
  prm.channels = 1;
  prm.stereo_mode = 0;
  prm.zero_run_enabled = zero_run ? 1u : 0u;
  prm.partitioning_enabled = partitioning ? 1u : 0u;
  prm.validate_range = 0;  // Block::Encoder accepts any int32
  uint32_t bb = 0;
  const int rc = lacb_encode(ctx, &prm, LACB_PLANAR_I32, pcm, nullptr, n, out, out_size, &bb, nullptr);
  if (rc != 0) return rc;
  ChanRec rec;
  CK(cudaMemcpy(&rec, ctx->recs.p, sizeof rec, cudaMemcpyDeviceToHost));
  lacb_block_info& bi = ctx->last_info;
  memset(&bi, 0, sizeof bi);
  bi.predictor_type = rec.type;
  bi.order = rec.order;
  bi.partition_order = rec.p;
  bi.n_parts = 1u << rec.p;
  bi.taps = rec.taps;
  bi.bits = rec.bits;
  for (int i = 0; i < 13; ++i) bi.coeffs[i] = rec.coef[i];
  for (uint32_t i = 0; i < bi.n_parts; ++i) {
    bi.part_mode[i] = rec.part[i] >> 5;
    bi.part_k[i] = rec.part[i] & 31u;
  }
  for (int i = 0; i < 11; ++i) bi.cand_best_lo[i] = rec.cand_lo[i];
  return 0;
}

int lacb_last_encode_decisions(lacb_ctx* ctx, lacb_block_decision* out, uint32_t capacity, uint32_t* n_blocks) {
  if (!ctx || !n_blocks) return LACB_EINVAL;
  const uint32_t nb = ctx->last_enc_blocks;
  if (nb == 0u) {
    ctx->err = "no single-pass encode to report (lacb_set_concurrency(ctx, 1) before the call)";
    return LACB_EINVAL;
  }
  *n_blocks = nb;
  if (!out) return 0;
  if (capacity < nb) return LACB_ENOMEM;
  CK(cudaSetDevice(ctx->device));
  std::vector<ChanRec> recs((size_t)nb * 4);
  std::vector<uint32_t> flags(nb);
  CK(cudaMemcpy(recs.data(), ctx->recs.p, recs.size() * sizeof(ChanRec), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(flags.data(), ctx->flags.p, (size_t)nb * 4, cudaMemcpyDeviceToHost));
  for (uint32_t b = 0; b < nb; ++b) {
    lacb_block_decision& d = out[b];
    memset(&d, 0, sizeof d);
    const uint32_t f = flags[b];
    const bool ms = ctx->last_enc_channels == 2u && (f & BF_CHOOSE_MS);
    d.flags = (ms ? LACB_DEC_MS : 0u) | ((f & BF_UNCERTAIN) ? LACB_DEC_UNCERTAIN : 0u) |
              ((f & BF_PROBE) ? LACB_DEC_PROBED : 0u) | ((f & BF_BOTH) ? LACB_DEC_BOTH : 0u);
    const uint64_t left = ctx->last_enc_frames - (uint64_t)b * kMaxBlock;
    d.block_size = left < kMaxBlock ? (uint32_t)left : kMaxBlock;
    for (uint32_t c = 0; c < ctx->last_enc_channels; ++c) {
      const ChanRec& r = recs[(size_t)b * 4 + (ms ? 2u : 0u) + c];
      lacb_chan_decision& o = d.ch[c];
      o.predictor_type = r.type; o.order = r.order; o.partition_order = r.p; o.taps = r.taps;
      o.base_mode = r.base_mode; o.has_run = r.has_run;
      o.bits = r.bits; o.bytes = r.bytes;
      o.est_bits = r.est_best; o.rice_bits = r.est_rice; o.zr_bits = r.est_zr; o.bin_bits = r.est_bin;
      o.static_bits = r.est_stat;
      for (int p = 0; p < 9; ++p) o.level_bits[p] = r.lvl_bits[p];
    }
  }
  return 0;
}

int lacb_last_block_info(lacb_ctx* ctx, lacb_block_info* info) {
  if (!ctx || !info) return LACB_EINVAL;
  *info = ctx->last_info;
  return 0;
}

int lacb_lpc_analyze(lacb_ctx* ctx, const int32_t* pcm, uint32_t n, int order, int16_t* coeffs_out) {
  if (!ctx || !pcm || n == 0 || n > kMaxBlock || !coeffs_out || order < 4 || order > 12 || (order & 1))
    return LACB_EINVAL;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CKR(ensure(ctx, ctx->planeL, (size_t)n * 4));
  CKR(ensure(ctx, ctx->jobs, 16));
  CKR(ensure(ctx, ctx->counts, 64));
  CKR(ensure(ctx, ctx->acor, 4 * 13 * 8));
  CKR(ensure(ctx, ctx->lpcq, 4 * sizeof(LpcQ)));
  CK(cudaMemcpyAsync(ctx->planeL.p, pcm, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  const uint32_t one[2] = {0u, 1u};  // job list {slot 0}, count 1
  CK(cudaMemcpyAsync(ctx->jobs.p, &one[0], 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->counts.p, &one[1], 4, cudaMemcpyHostToDevice, st));
  PcmSrc src{as<int32_t>(ctx->planeL), nullptr, n};
  auto ka = k_autocorr_stream<256>;
  LACB_LAUNCH(ka, 1, 256, 0, st, src, as<uint32_t>(ctx->jobs), as<uint32_t>(ctx->counts), as<i64>(ctx->acor));
  auto kl = k_levinson<false>;
  LACB_LAUNCH(kl, 1, 64, 0, st, src, as<uint32_t>(ctx->jobs), as<uint32_t>(ctx->counts), as<i64>(ctx->acor),
              as<LpcQ>(ctx->lpcq));
  LpcQ q;
  CK(cudaMemcpyAsync(&q, ctx->lpcq.p, sizeof q, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  const int c = (order - 4) / 2;
  const uint32_t max_valid = n > 1u ? (n - 1u < 32u ? n - 1u : 32u) : 0u;
  for (int i = 0; i <= order; ++i) coeffs_out[i] = 0;
  if ((uint32_t)order > max_valid) return 0;
  for (int i = 1; i <= order; ++i) coeffs_out[i] = q.coef[c][i];
  return q.used[c];
}

// ---------------------------------------------------------------------------
static int decode_common(lacb_ctx* ctx, const lacb_dec_params* prm, const uint8_t* d_payload, uint64_t payload_bytes,
                         const uint32_t* block_sizes, const uint32_t* block_bytes, uint32_t n_blocks, int32_t* dL,
                         int32_t* dR, uint8_t* d_packed, lacb_err* err) {
  const bool serial = block_bytes == nullptr;  // v2 stream: one chain, no per-block byte sizes
  cudaStream_t st = ctx->stream;
  // tables: first sample and byte offset of every block (host prefix sums, O(n_blocks))
  const size_t tb = (size_t)n_blocks * (8 + 4 + 8 + 4);
  CKR(ensure_pinned(ctx, tb));
  u64* h_fs = reinterpret_cast<u64*>(ctx->pinned);
  u64* h_boff = h_fs + n_blocks;
  uint32_t* h_size = reinterpret_cast<uint32_t*>(h_boff + n_blocks);
  uint32_t* h_bytes = h_size + n_blocks;
  u64 fs = 0, bo = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    h_fs[b] = fs;
    h_boff[b] = bo;
    h_size[b] = block_sizes[b];
    h_bytes[b] = serial ? 0u : block_bytes[b];
    fs += block_sizes[b];
    bo += h_bytes[b];
  }
  if (bo > payload_bytes) {
    ctx->err = "compressed block sizes exceed frame payload";
    set_err(err, LACB_EDECODE, 0, 0, "[decode-error] compressed block sizes exceed frame payload");
    return LACB_EDECODE;
  }
  CKR(ensure(ctx, ctx->d_fs, (size_t)n_blocks * 8));
  CKR(ensure(ctx, ctx->d_boff, (size_t)n_blocks * 8));
  CKR(ensure(ctx, ctx->d_size, (size_t)n_blocks * 4));
  CKR(ensure(ctx, ctx->d_bytes, (size_t)n_blocks * 4));
  CKR(ensure(ctx, ctx->d_err, (size_t)n_blocks * 4));
  CKR(ensure(ctx, ctx->d_ms, (size_t)n_blocks));
  CK(cudaMemcpyAsync(ctx->d_fs.p, h_fs, (size_t)n_blocks * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_boff.p, h_boff, (size_t)n_blocks * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_size.p, h_size, (size_t)n_blocks * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_bytes.p, h_bytes, (size_t)n_blocks * 4, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(ctx->ev[EV_H2D], st));
  static const int lpc_fp64 = getenv("LACB_RESTORE_FP64") ? atoi(getenv("LACB_RESTORE_FP64")) : 1;  // A/B knob: 0 = integer chain
  DecCfg cfg{prm->channels, prm->stereo_mode, prm->bit_depth, n_blocks, (uint32_t)(lpc_fp64 != 0)};
  CKR(ensure(ctx, ctx->d_hdrs, (size_t)n_blocks * 2 * sizeof(ChanHdr)));
  if (serial) {
    CKR(ensure(ctx, ctx->misc, 64));
    auto ks = k_parse_serial;
    LACB_LAUNCH(ks, 1, 32, 0, st, cfg, d_payload, (u64)payload_bytes, (u64)((payload_bytes + 15ull) & ~15ull),
                as<u64>(ctx->d_fs), as<uint32_t>(ctx->d_size), dL, dR, as<ChanHdr>(ctx->d_hdrs),
                as<uint32_t>(ctx->d_err), as<uint8_t>(ctx->d_ms), as<uint32_t>(ctx->misc));
  } else {
    auto kp = k_parse_blocks;
    LACB_LAUNCH(kp, (n_blocks + kParseWarps - 1) / kParseWarps, 32 * kParseWarps, 0, st, cfg, d_payload,
                (u64)((payload_bytes + 15ull) & ~15ull), as<u64>(ctx->d_fs), as<uint32_t>(ctx->d_size),
                as<u64>(ctx->d_boff), as<uint32_t>(ctx->d_bytes), dL, dR, as<ChanHdr>(ctx->d_hdrs),
                as<uint32_t>(ctx->d_err), as<uint8_t>(ctx->d_ms));
  }
  CK(cudaEventRecord(ctx->ev[EV_LPC], st));
  CKR(ensure(ctx, ctx->d_order, ((size_t)n_blocks * 2 + 8 * 32 + 1) * 4));
  uint32_t* d_order = as<uint32_t>(ctx->d_order);
  auto ko = k_restore_order;
  // 256 threads, not 1024: in the host pipelines this one-CTA kernel starts while the parser warps of the other slices
  // fill the SMs (24 warps of 64 registers: 16 K registers free per SM); a 1024-thread CTA (20 K registers) did not fit
  // anywhere until parsers retired -- depending on how the parser CTAs had spread, a slice's restore waited up to 6 ms
  // (a decode that has the device to itself keeps the 1024-thread CTA: 17 us instead of 51)
  LACB_LAUNCH(ko, 1, ctx->per_job_grid ? 256 : 1024, 0, st, cfg, as<ChanHdr>(ctx->d_hdrs), as<uint32_t>(ctx->d_err),
              d_order + 1, d_order);
  CK(cudaEventRecord(ctx->ev[EV_STEREO], st));  // (decode: the restore order is known)
  auto kr = k_restore_blocks;
  LACB_LAUNCH(kr, (n_blocks * prm->channels + 8u * 32u + kRestoreTpb - 1u) / kRestoreTpb, kRestoreTpb, 0, st, cfg, as<u64>(ctx->d_fs),
              as<uint32_t>(ctx->d_size), dL, dR, as<ChanHdr>(ctx->d_hdrs), as<uint32_t>(ctx->d_err), d_order + 1,
              d_order);
  auto km = k_merge_restore_errors;
  LACB_LAUNCH(km, (n_blocks + 255u) / 256u, 256, 0, st, cfg, as<uint32_t>(ctx->d_err));
  CK(cudaEventRecord(ctx->ev[EV_ANALYZE], st));
  auto kf = k_finish_pcm;
  LACB_LAUNCH(kf, lacb_umin(n_blocks, (uint32_t)ctx->sms * 8u), 256, 0, st, cfg, as<u64>(ctx->d_fs),
              as<uint32_t>(ctx->d_size), dL, dR, as<uint32_t>(ctx->d_err), as<uint8_t>(ctx->d_ms), d_packed,
              (uint32_t)(dL != as<int32_t>(ctx->d_L)) /* caller-owned planes are part of the result */);
  CK(cudaEventRecord(ctx->ev[EV_EMIT], st));
  return 0;
}

static int decode_check_errors(lacb_ctx* ctx, uint32_t n_blocks, lacb_err* err, bool serial = false,
                               uint32_t base_block = 0) {
  std::vector<uint32_t> herr(n_blocks);
  uint32_t info[2] = {n_blocks, 0u};
  CK(cudaMemcpyAsync(herr.data(), ctx->d_err.p, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (serial) CK(cudaMemcpyAsync(info, ctx->misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaGetLastError());
  if (serial) {  // blocks after the first failure were never parsed
    for (uint32_t b = info[0] + 1u; b < n_blocks; ++b) herr[b] = DERR_OK;
  }
  for (uint32_t b = 0; b < n_blocks; ++b) {
    if (herr[b] == DERR_OK) continue;
    char msg[160];
    switch (herr[b]) {
      case DERR_FLAG: snprintf(msg, sizeof msg, "[decode-error] invalid per-block stereo flag"); break;
      case DERR_PRIMARY: snprintf(msg, sizeof msg, "[decode-error] block=%u channel=primary", base_block + b); break;
      case DERR_SECONDARY:
        snprintf(msg, sizeof msg, "[decode-error] block=%u channel=secondary", base_block + b);
        break;
      case DERR_RANGE: snprintf(msg, sizeof msg, "[decode-error] decoded sample outside PCM bit depth"); break;
      default:
        snprintf(msg, sizeof msg, "[decode-error] block=%u channel=trailing-payload", base_block + b);
        break;
    }
    ctx->err = msg;
    set_err(err, LACB_EDECODE, base_block + b, herr[b], msg);
    return LACB_EDECODE;
  }
  if (serial && info[1]) {
    ctx->err = "[decode-error] trailing frame payload";
    set_err(err, LACB_EDECODE, n_blocks, DERR_TRAILING, "[decode-error] trailing frame payload");
    return LACB_EDECODE;
  }
  return 0;
}

static void fill_dec_timing(lacb_ctx* ctx) {
  lacb_timing& t = ctx->timing;
  memset(&t, 0, sizeof t);
  t.h2d_ms = ev_ms(ctx, EV_START, EV_H2D);
  t.parse_ms = ev_ms(ctx, EV_H2D, EV_LPC);
  t.restore_ms = ev_ms(ctx, EV_LPC, EV_ANALYZE);
  t.finish_ms = ev_ms(ctx, EV_ANALYZE, EV_EMIT);
  t.d2h_ms = ev_ms(ctx, EV_EMIT, EV_D2H);
  t.total_ms = ev_ms(ctx, EV_START, EV_D2H);
}

// Block table checks shared by both decode entry points: sample counts in [1, 16384], at least one
// payload byte per block (v3 tables), byte sizes that fit the payload (lac/decoder.cpp:118-137).
static int validate_tables(lacb_ctx* ctx, const uint32_t* block_sizes, const uint32_t* block_bytes, uint32_t n_blocks,
                           uint64_t payload_bytes, u64* frames_out, lacb_err* err) {
  u64 frames = 0, bytes = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) {
    if (block_sizes[b] == 0 || block_sizes[b] > kMaxBlock) {
      ctx->err = "invalid block size";
      set_err(err, LACB_EDECODE, b, 0, "[decode-error] invalid block size");
      return LACB_EDECODE;
    }
    frames += block_sizes[b];
    if (block_bytes) {
      if (block_bytes[b] == 0) {
        ctx->err = "invalid compressed block size";
        set_err(err, LACB_EDECODE, b, 0, "[decode-error] invalid compressed block size");
        return LACB_EDECODE;
      }
      bytes += block_bytes[b];
      if (bytes > payload_bytes) {
        ctx->err = "compressed block sizes exceed frame payload";
        set_err(err, LACB_EDECODE, b, 0, "[decode-error] compressed block sizes exceed frame payload");
        return LACB_EDECODE;
      }
    }
  }
  if (frames_out) *frames_out = frames;
  return 0;
}

static bool dec_params_ok(const lacb_dec_params* p) {
  return p && (p->channels == 1 || p->channels == 2) && (p->bit_depth == 16 || p->bit_depth == 24) &&
         p->stereo_mode <= 2 && !(p->channels == 1 && p->stereo_mode != 0);
}

int lacb_decode_device(lacb_ctx* ctx, const lacb_dec_params* prm, const uint8_t* d_payload, uint64_t payload_bytes,
                       const uint32_t* block_sizes_host, const uint32_t* block_bytes_host, uint32_t n_blocks,
                       int32_t* d_left, int32_t* d_right, uint8_t* d_packed, lacb_err* err) {
  if (!ctx) return LACB_EINVAL;
  if (!dec_params_ok(prm) || !d_payload || !block_sizes_host || !block_bytes_host || n_blocks == 0 ||
      (d_left && prm->channels == 2 && !d_right)) {
    ctx->err = "invalid decode arguments";
    return LACB_EINVAL;
  }
  CK(cudaSetDevice(ctx->device));
  u64 frames = 0;
  CKR(validate_tables(ctx, block_sizes_host, block_bytes_host, n_blocks, payload_bytes, &frames, err));
  if (!d_left) {  // planes are an intermediate: keep them in the context's workspace
    CKR(ensure(ctx, ctx->d_L, frames * 4));
    if (prm->channels == 2) CKR(ensure(ctx, ctx->d_R, frames * 4));
    d_left = as<int32_t>(ctx->d_L);
    d_right = as<int32_t>(ctx->d_R);
  }
  CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
  CKR(decode_common(ctx, prm, d_payload, payload_bytes, block_sizes_host, block_bytes_host, n_blocks, d_left, d_right,
                    d_packed, err));
  CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
  const int rc = decode_check_errors(ctx, n_blocks, err);
  fill_dec_timing(ctx);
  return rc;
}

int lacb_decode(lacb_ctx* ctx, const lacb_dec_params* prm, const uint8_t* payload, uint64_t payload_bytes,
                const uint32_t* block_sizes, const uint32_t* block_bytes, uint32_t n_blocks, int layout, void* out_a,
                void* out_b, lacb_err* err) {
  if (!ctx) return LACB_EINVAL;
  if (!dec_params_ok(prm) || !payload || !block_sizes || n_blocks == 0 || !out_a ||
      (layout == LACB_PLANAR_I32 && prm->channels == 2 && !out_b) ||
      (layout != LACB_PLANAR_I32 && layout != LACB_PACKED_LE)) {
    ctx->err = "invalid decode arguments";
    return LACB_EINVAL;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  u64 frames = 0;
  CKR(validate_tables(ctx, block_sizes, block_bytes, n_blocks, payload_bytes, &frames, err));
  const uint32_t bps = prm->bit_depth / 8;
  // The parser (one warp per block) and the restore kernel (one thread per channel-block) are serial
  // chains: a launch of a few hundred blocks takes as long as one of a few thousand (measured: 2.9 ms
  // of parsing for 586 blocks, 3.55 ms for 3516), so slices cannot shorten the kernels.  What they
  // buy is the copies: with four slices in flight on four streams the host->device copy of the
  // payload and the device->host copy of the samples run under the chains of the other slices
  // (14.9 -> 12.4 ms for 600 s of 24/96 stereo; more streams than that start to share hardware queues).
  const uint32_t nk = ctx->max_streams ? lacb_umin(dec_kids(), ctx->max_streams) : dec_kids();
  const uint32_t dec_sb = getenv("LACB_DEC_SLICE_BLOCKS") ? (uint32_t)atol(getenv("LACB_DEC_SLICE_BLOCKS"))
                          : getenv("LACB_SLICE_BLOCKS")   ? slice_blocks(ctx, prm->channels)
                          : n_blocks < 1024u              ? n_blocks
                                                          : lacb_umin((uint32_t)ctx->sms * 32u, (n_blocks + nk - 1u) / nk);
  if (block_bytes && dec_sb > 0u && n_blocks > dec_sb && nk > 1u) {
    // pipelined: slices of whole blocks alternate between two slice contexts (see encode_host_sliced)
    CKR(ensure_kids(ctx, nk));
    const uint32_t sb = dec_sb;
    const uint32_t ns = (n_blocks + sb - 1u) / sb;
    u64 tot_bytes = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) tot_bytes += block_bytes[b];
    if (tot_bytes > payload_bytes) {
      ctx->err = "compressed block sizes exceed frame payload";
      set_err(err, LACB_EDECODE, 0, 0, "[decode-error] compressed block sizes exceed frame payload");
      return LACB_EDECODE;
    }
    u64 boff = 0, foff = 0;
    auto begin = [&](uint32_t i) -> int {
      lacb_ctx* k = ctx->kids[i % nk];
      const uint32_t b0 = i * sb, b1 = b0 + sb < n_blocks ? b0 + sb : n_blocks;
      u64 sbytes = 0, sframes = 0;
      for (uint32_t b = b0; b < b1; ++b) {
        sbytes += block_bytes[b];
        sframes += block_sizes[b];
      }
      lacb_ctx* ctx = k;  // CK / CKR report into the slice context
      CKR(ensure(ctx, ctx->d_payload, sbytes + 64));
      CKR(ensure(ctx, ctx->d_L, sframes * 4));
      if (prm->channels == 2) CKR(ensure(ctx, ctx->d_R, sframes * 4));
      if (layout == LACB_PACKED_LE) CKR(ensure(ctx, ctx->d_packed, sframes * prm->channels * bps));
      CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_payload.p, payload + boff, sbytes, cudaMemcpyHostToDevice, ctx->stream));
      CKR(decode_common(ctx, prm, as<uint8_t>(ctx->d_payload), sbytes, block_sizes + b0, block_bytes + b0, b1 - b0,
                        as<int32_t>(ctx->d_L), as<int32_t>(ctx->d_R),
                        layout == LACB_PACKED_LE ? as<uint8_t>(ctx->d_packed) : nullptr, err));
      if (layout == LACB_PACKED_LE) {
        CK(cudaMemcpyAsync(static_cast<uint8_t*>(out_a) + foff * prm->channels * bps, ctx->d_packed.p,
                           sframes * prm->channels * bps, cudaMemcpyDeviceToHost, ctx->stream));
      } else {
        CK(cudaMemcpyAsync(static_cast<int32_t*>(out_a) + foff, ctx->d_L.p, sframes * 4, cudaMemcpyDeviceToHost,
                           ctx->stream));
        if (prm->channels == 2)
          CK(cudaMemcpyAsync(static_cast<int32_t*>(out_b) + foff, ctx->d_R.p, sframes * 4, cudaMemcpyDeviceToHost,
                             ctx->stream));
      }
      CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
      boff += sbytes;
      foff += sframes;
      return 0;
    };
    // up to nk slices are in flight; a slice context is collected (in block order, so the first failing
    // block is the one reported) right before it is needed again, and at the end
    auto collect = [&](uint32_t i) -> int {
      lacb_ctx* k = ctx->kids[i % nk];
      const uint32_t b0 = i * sb, b1 = b0 + sb < n_blocks ? b0 + sb : n_blocks;
      const int r = decode_check_errors(k, b1 - b0, err, false, b0);  // waits for slice i only
      if (trace_on() && r == 0) {  // device timeline of the slice, relative to the start of the call
        float t[7] = {0, 0, 0, 0, 0, 0, 0};
        const int evs[7] = {EV_START, EV_H2D, EV_LPC, EV_STEREO, EV_ANALYZE, EV_EMIT, EV_D2H};
        for (int e = 0; e < 7; ++e) cudaEventElapsedTime(&t[e], ctx->ev[EV_START], k->ev[evs[e]]);
        TRACE("dec slice %u done: start %.2f h2d %.2f parse %.2f order %.2f restore %.2f finish %.2f d2h %.2f ms", i, t[0],
              t[1], t[2], t[3], t[4], t[5], t[6]);
      }
      if (r != 0) ctx->err = k->err;
      return r;
    };
    if (trace_on()) cudaEventRecord(ctx->ev[EV_START], ctx->stream);
    int rc = 0;
    uint32_t queued = 0, collected = 0;
    for (; queued < ns && rc == 0; ++queued) {
      if (queued >= nk) rc = collect(collected++);
      if (rc != 0) break;
      rc = begin(queued);
      if (rc != 0) ctx->err = ctx->kids[queued % nk]->err;
      TRACE("dec slice %u queued", queued);
    }
    while (rc == 0 && collected < queued) rc = collect(collected++);
    cudaDeviceSynchronize();
    memset(&ctx->timing, 0, sizeof ctx->timing);
    return rc;
  }
  CKR(ensure(ctx, ctx->d_payload, payload_bytes + 64));
  CKR(ensure(ctx, ctx->d_L, frames * 4));
  if (prm->channels == 2) CKR(ensure(ctx, ctx->d_R, frames * 4));
  if (layout == LACB_PACKED_LE) CKR(ensure(ctx, ctx->d_packed, frames * prm->channels * bps));
  CK(cudaEventRecord(ctx->ev[EV_START], st));
  CK(cudaMemcpyAsync(ctx->d_payload.p, payload, payload_bytes, cudaMemcpyHostToDevice, st));
  CKR(decode_common(ctx, prm, as<uint8_t>(ctx->d_payload), payload_bytes, block_sizes, block_bytes, n_blocks,
                    as<int32_t>(ctx->d_L), as<int32_t>(ctx->d_R),
                    layout == LACB_PACKED_LE ? as<uint8_t>(ctx->d_packed) : nullptr, err));
  const int rc = decode_check_errors(ctx, n_blocks, err, block_bytes == nullptr);
  if (rc != 0) return rc;
  if (layout == LACB_PACKED_LE) {
    CK(cudaMemcpyAsync(out_a, ctx->d_packed.p, frames * prm->channels * bps, cudaMemcpyDeviceToHost, st));
  } else {
    CK(cudaMemcpyAsync(out_a, ctx->d_L.p, frames * 4, cudaMemcpyDeviceToHost, st));
    if (prm->channels == 2) CK(cudaMemcpyAsync(out_b, ctx->d_R.p, frames * 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(ctx->ev[EV_D2H], st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  fill_dec_timing(ctx);
  return 0;
}

int lacb_decode_block_at(lacb_ctx* ctx, const uint8_t* data, uint64_t size, uint64_t bit_offset, uint32_t block_size,
                         int32_t* out, uint64_t* bits_consumed, int* ran_out) {
  if (!ctx || !out) return LACB_EINVAL;
  if (bits_consumed) *bits_consumed = 0;
  if (ran_out) *ran_out = 0;
  if (block_size == 0 || block_size > kMaxBlock) return 0;
  if (bit_offset > size * 8ull) {
    if (ran_out) *ran_out = 1;
    return 0;
  }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CKR(ensure(ctx, ctx->d_payload, size + 16));
  CKR(ensure(ctx, ctx->d_L, (size_t)block_size * 4));
  CKR(ensure(ctx, ctx->misc, 64));
  if (size) CK(cudaMemcpyAsync(ctx->d_payload.p, data, size, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->d_L.p, 0, (size_t)block_size * 4, st));
  auto kd = k_decode_one;
  LACB_LAUNCH(kd, 1, 32, 0, st, as<uint8_t>(ctx->d_payload), (u64)size, (u64)((size + 15ull) & ~15ull), (u64)bit_offset,
              block_size, as<int32_t>(ctx->d_L), as<u64>(ctx->misc));
  u64 res[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(res, ctx->misc.p, 24, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(out, ctx->d_L.p, (size_t)block_size * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (bits_consumed) *bits_consumed = res[1];
  if (ran_out) *ran_out = res[2] ? 1 : 0;
  return res[0] ? 1 : 0;
}

int lacb_decode_block(lacb_ctx* ctx, const uint8_t* data, uint64_t size, uint32_t block_size, int32_t* out,
                      uint64_t* bits_consumed) {
  return lacb_decode_block_at(ctx, data, size, 0, block_size, out, bits_consumed, nullptr);
}

int lacb_set_concurrency(lacb_ctx* ctx, uint32_t max_slices_in_flight) {
  if (!ctx) return LACB_EINVAL;
  ctx->max_streams = max_slices_in_flight;
  return 0;
}

}  // extern "C"
