#!/usr/bin/env python
"""bench.py -- LAC block codec hot path on B200: encode + decode PCM GB/s.

One "step" = one pass of the hot path over one batch: encode the batch to .lac block
payloads, then decode those payloads back to PCM.  Workload = BASELINE.json configs[1]
(10 min synthetic 24-bit 96 kHz stereo, forced --stereo-mode=ms) per GPU; with N GPUs
every rank holds its own 10-minute block range of an N x 10 min file (weak scaling) and
the only exchange is the NCCL all-gather of per-rank payload byte counts.

  value : PCM bytes through encode+decode per second, PCM resident in HBM (packed WAV
          sample bytes in, packed WAV sample bytes out), whole job over all ranks
  e2e   : the same through the C ABI with HOST buffers (H2D of PCM / payload and D2H of
          payload / PCM inside the timed region)
  roofline     : the dominant kernel (encoder channel-block analysis) against measured HBM copy bandwidth
  cpu_baseline : the reference CPU codec (oracle/_ref when built, else the C port) on the
                 box's host cores, on a bounded slice of the same workload

`--impl reference` times only the CPU codec (all host threads) on a bounded slice.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# The host pipelines of the library keep up to twelve slice streams in flight; the CUDA runtime maps streams onto 8
# hardware queues by default and slices that share a queue serialise.  Must be set before CUDA starts (`lac_cli serve`
# does the same); the library only goes beyond four decode slices in flight when it sees this setting.  It lengthens
# CUDA's start-up by 1 - 2.5 s (tools/cli_conn_check.py), which is why one-shot CLI runs leave it alone.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

RATE, DEPTH, CHANNELS, STEREO_MODE = 96000, 24, 2, 1
METRIC = "encode+decode PCM throughput (BASELINE: encode & decode PCM GB/s, byte-identical .lac vs CPU ref)"
UNIT = "GB/s"


_REAL_STDOUT = None


def quiet_stdout():
    """stdout must carry the one JSON line and nothing else, but libraries write to it too (NCCL prints its
    version banner there at communicator creation): everything else that goes to fd 1 is sent to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj) -> None:
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def synth_packed(seed: int, frames: int) -> np.ndarray:
    """SURVEY.md Appendix C generator (tools/lac_synth.c), packed 24-bit stereo bytes."""
    so = ROOT / "tools" / "liblac_synth.so"
    if not so.exists():
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", str(so), str(ROOT / "tools" / "lac_synth.c")])
    lib = C.CDLL(str(so))
    lib.lac_synth.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lac_synth.restype = None
    out = np.zeros(frames * CHANNELS * (DEPTH // 8), dtype=np.uint8)
    lib.lac_synth(seed, frames, DEPTH, CHANNELS, None, None, out.ctypes.data)
    return out


def unpack24(pk: np.ndarray):
    b = pk.reshape(-1, 2, 3).astype(np.int32)
    v = b[:, :, 0] | (b[:, :, 1] << 8) | (b[:, :, 2] << 16)
    v = (v << 8) >> 8
    return np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self) -> int:
        return len(self.lines)

    def stop(self, first: int = 0, last: int | None = None) -> dict:
        """Clocks / throttle reasons of the samples [first, last) (marks taken around the timed region); when
        the region was shorter than one sampling period the samples just around it are used."""
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        last = len(self.lines) if last is None else last
        if last <= first:
            first, last = max(0, first - 1), min(len(self.lines), last + 2)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines[first:last]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_codec():
    """(codec, kind): the compiled reference when oracle/_ref exists, else the C port."""
    import helpers as H
    if H.have_ref():
        return H.ref(), "reference"
    return H.oracle(), "port"


def cpu_roundtrip(codec, kind, l, r, threads):
    """One encode + decode of the planes on the host cores -> (encode s, decode s, .lac bytes).  With the compiled
    reference the clock runs inside the shim around LAC::Encoder::encode / LAC::Decoder::decode only (the
    std::vector copies a C caller needs are made before it starts, oracle/ref_shim.cpp); the decode is checked
    against the input there."""
    import helpers as H
    lp, rp = l.ctypes.data_as(H.i32p), r.ctypes.data_as(H.i32p)
    if kind == "reference":
        L = codec.lib
        L.ref_encode_timed.argtypes = [H.i32p, H.i32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.POINTER(H.u8p), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.ref_decode_timed.argtypes = [H.u8p, C.c_uint64, C.c_uint32, H.i32p, H.i32p, C.c_uint64,
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]
        out, n, te, td, match = H.u8p(), C.c_uint64(), C.c_double(), C.c_double(), C.c_int()
        rc = L.ref_encode_timed(lp, rp, l.size, RATE, DEPTH, STEREO_MODE, threads, C.byref(out), C.byref(n), C.byref(te))
        assert rc == 0, codec.last_error()
        rc = L.ref_decode_timed(out, n.value, threads, lp, rp, l.size, C.byref(td), C.byref(match))
        codec.free(out)
        assert rc == 0 and match.value == 1, "reference round trip does not restore the PCM"
        return te.value, td.value, n.value
    t0 = time.perf_counter()
    blob = codec.encode(l, r, RATE, DEPTH, STEREO_MODE, threads=threads)
    t1 = time.perf_counter()
    dl, dr, _ = codec.decode(blob, threads=threads)
    t2 = time.perf_counter()
    assert np.array_equal(dl, l) and np.array_equal(dr, r)
    return t1 - t0, t2 - t1, len(blob)


def cpu_sample(args):
    """The CPU legs run the SAME workload as the GPU arm (the whole `--seconds` of rank 0's input, ~3 s of work
    per step on 16 threads); `--cpu-seconds` can bound it further, and the line then says so."""
    import helpers as H
    secs = args.seconds if args.cpu_seconds <= 0 else min(args.cpu_seconds, args.seconds)
    l, r = H.synth(2, RATE * secs, DEPTH)
    nbytes = l.size * CHANNELS * (DEPTH // 8)
    what = (f"the whole workload ({secs} s, {nbytes / 1e6:.1f} MB PCM)" if secs == args.seconds
            else f"first {secs} s of the workload ({nbytes / 1e6:.1f} MB PCM)")
    return l, r, nbytes, secs, what


def run_reference(args, rank):
    if rank != 0:
        return
    codec, kind = cpu_codec()
    cores = os.cpu_count() or 1
    l, r, nbytes, secs, what = cpu_sample(args)
    for _ in range(args.warmup):
        cpu_roundtrip(codec, kind, l, r, cores)
    te = td = 0.0
    for _ in range(args.steps):
        a, b, _ = cpu_roundtrip(codec, kind, l, r, cores)
        te += a
        td += b
    val = nbytes * args.steps / (te + td) / 1e9
    sample = f"{what} per step, encode+decode, {cores} threads, timed around LAC::Encoder::encode / LAC::Decoder::decode"
    emit_line({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64", "data": "synthetic",
        "config": workload_config(args, secs),  # the workload this run timed (== the GPU arm's unless --cpu-seconds bounds it)
        "encode_gbs": nbytes * args.steps / te / 1e9, "decode_gbs": nbytes * args.steps / td / 1e9,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args, secs):
    return {"workload": f"BASELINE configs[1]: {secs} s synthetic 24-bit 96 kHz stereo WAV samples per GPU, "
                        "forced --stereo-mode=ms, encode then decode (SURVEY.md Appendix C generator, seed 2+rank)",
            "frames_per_gpu": RATE * secs, "block_samples": 16384, "stereo_mode": "ms",
            "l2": "inputs (345.6 MB PCM, ~213 MB payload per GPU) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"block-range sharding x{args.gpus}, NCCL all-gather of payload byte counts only"}


def run_c4(args, cd, rank, world, dist, torch):
    """BASELINE config 4 as specified: ONE 10 h 24-bit / 48 kHz stereo file (auto LR/MS), sharded by contiguous
    block range over the `world` ranks (ceil(n_blocks / world) blocks each), through the host-buffer C ABI.

    encode : every rank copies its PCM range host->device and encodes it, the ranks all-gather their payload byte
             counts over NCCL, and every rank DMAs its slab from HBM to its GLOBAL offset inside the one mapped
             output file and writes its table slice (rank 0 adds header + block count): src/codec/lac/encoder.cpp:445-465
             done by `world` processes.  All of that is inside the timed region.
    check  : every rank's slab and table slice against the reference's SHA-256 (tests/golden/golden_large.json,
             recorded from the unmodified reference by tools/make_golden_large.py), rank 0 the assembled file.
    decode : every rank decodes its slab back to packed PCM in host memory; compared with the input.
    Returns the `c4` sub-record on rank 0."""
    import hashlib
    import helpers as H
    g = json.loads((ROOT / "tests" / "golden" / "golden_large.json").read_text())["C4_full_10h_24_48k_auto"]
    frames_total, nb_total, depth, ch, mode = g["frames"], g["n_blocks"], g["depth"], g["channels"], g["stereo_mode"]
    shard = g["shards"][str(world)][rank]
    b0, nbr = shard["first_block"], shard["blocks"]
    f0 = b0 * 16384
    fr = min(frames_total, (b0 + nbr) * 16384) - f0
    fb = ch * (depth // 8)
    t_gen = time.perf_counter()
    h_in = cd.pinned(fr * fb)
    H.synth_lib().lac_synth_range(g["seed"], f0, fr, depth, ch, g["reset_log2"], None, None,
                                  h_in.ctypes.data_as(H.u8p))
    t_gen = time.perf_counter() - t_gen
    h_out = cd.pinned(fr * fb)
    d_pcm = cd.dev_malloc(fr * fb)
    sizes = np.full(nbr, 16384, dtype=np.uint32)
    sizes[-1] = fr - 16384 * (nbr - 1)
    shm = Path("/dev/shm")
    try:
        base = shm if shm.is_dir() and os.statvfs(shm).f_bavail * os.statvfs(shm).f_frsize > g["len"] + (1 << 30) else Path("/tmp")
    except OSError:
        base = Path("/tmp")
    out_path = base / f"lacb_c4_{os.environ.get('MASTER_PORT', 'single')}.lac"
    counts = torch.zeros(world, dtype=torch.int64, device="cuda") if dist else None
    mine = torch.zeros(1, dtype=torch.int64, device="cuda") if dist else None

    def barrier():
        if dist:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    if rank == 0:
        with open(out_path, "wb") as f:
            f.truncate(g["len"])
    barrier()
    # every rank maps the ONE output file; its slab is DMA-ed from HBM straight to its global offset in the
    # mapping (page-locked with lacb_host_register once the offset is known), so the host "concatenation" is the
    # device->host copy itself: no staging buffer, no memcpy
    import mmap
    fd = os.open(out_path, os.O_RDWR)
    mm = mmap.mmap(fd, g["len"])
    mm_arr = np.frombuffer(mm, dtype=np.uint8)
    base_ptr = mm_arr.ctypes.data
    state = {"registered": None, "bb": None, "off": 0}
    res = {}

    def encode_step():
        cd.h2d(d_pcm, h_in)
        d_payload, n, d_bb = cd.encode_device(d_pcm, 0, fr, depth, ch, mode)
        bb = cd.d2h(d_bb, nbr * 4, np.uint32)
        if dist:  # global slab offsets = exclusive scan of the gathered byte counts
            mine[0] = n
            dist.all_gather_into_tensor(counts, mine)
            allc = counts.cpu().numpy()
        else:
            allc = np.array([n], dtype=np.int64)
        t_a = time.perf_counter()
        off = 14 + 8 * nb_total + int(allc[:rank].sum())
        table = np.empty((nbr, 2), dtype=">u4")
        table[:, 0] = sizes
        table[:, 1] = bb
        mm_arr[14 + 8 * b0:14 + 8 * (b0 + nbr)] = np.frombuffer(table.tobytes(), dtype=np.uint8)
        if state["registered"] is None:
            lo = (base_ptr + off) & ~4095
            hi = (base_ptr + off + n + 4095) & ~4095
            state["registered"] = (lo, hi - lo) if cd.host_register(lo, hi - lo) else False
        cd.d2h_to(base_ptr + off, d_payload, n)
        if rank == 0:
            mm_arr[:14] = np.frombuffer(lacb_header(ch, mode, g["rate"], depth) + int(nb_total).to_bytes(4, "big"), dtype=np.uint8)
        state["bb"], state["off"] = bb, off
        return n, int(allc.sum()), time.perf_counter() - t_a

    def decode_step(n):  # reads the slab back out of the assembled file
        cd.decode_into(mm_arr[state["off"]:state["off"] + n], sizes, state["bb"], depth, ch, mode, h_out)

    n, total, _ = encode_step()  # warm-up (workspaces grow here) + the verified pass
    decode_step(n)
    barrier()
    ok_slab = n == shard["payload_bytes"] and hashlib.sha256(memoryview(mm_arr[state["off"]:state["off"] + n])).hexdigest() == shard["payload_sha256"]
    tb = np.empty((nbr, 2), dtype=">u4")
    tb[:, 0] = sizes
    tb[:, 1] = state["bb"]
    ok_table = hashlib.sha256(tb.tobytes()).hexdigest() == shard["table_sha256"]
    ok_rt = bool(np.array_equal(h_out, h_in))
    ok_file = True
    if rank == 0:
        h = hashlib.sha256()
        with open(out_path, "rb") as f:
            while True:
                chunk = f.read(1 << 26)
                if not chunk:
                    break
                h.update(chunk)
        ok_file = total == g["payload_bytes"] and h.hexdigest() == g["sha256"]
    steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    t_asm = 0.0
    for _ in range(steps):
        n, total, ta = encode_step()
        t_asm += ta
    barrier()
    t_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(steps):
        decode_step(n)
    barrier()
    t_dec = time.perf_counter() - t0
    if state["registered"]:
        cd.host_unregister(state["registered"][0])
    pinned_out = bool(state["registered"])
    del mm_arr
    try:
        mm.close()
    except BufferError:
        pass
    os.close(fd)
    cd.dev_free(d_pcm)
    flags = [ok_slab, ok_table, ok_rt, ok_file]
    if dist:
        t = torch.tensor([t_enc, t_dec, t_asm, t_gen] + [0.0 if f else 1.0 for f in flags], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_enc, t_dec, t_asm, t_gen = (float(x) for x in t[:4])
        flags = [float(x) == 0.0 for x in t[4:]]
    barrier()
    if rank == 0:
        try:
            os.unlink(out_path)
        except OSError:
            pass
        pcm = frames_total * fb
        res = {"workload": "BASELINE configs[3]: ONE 10 h synthetic 24-bit 48 kHz stereo file (auto LR/MS, range-addressable "
                           f"Appendix C stream, seed 4), {nb_total} blocks sharded by contiguous block range over {world} GPU(s), "
                           "host buffers through the C ABI; NCCL all-gather of payload byte counts -> global offsets; every "
                           "rank writes its table slice + slab into the one .lac",
               "pcm_bytes": pcm, "lac_bytes": g["len"], "steps": steps,
               "encode_gbs": pcm * steps / t_enc / 1e9, "decode_gbs": pcm * steps / t_dec / 1e9,
               "encode_decode_gbs": pcm * steps / (t_enc + t_dec) / 1e9,
               "assemble_ms_per_step": t_asm / steps * 1e3, "output_mapping_page_locked": pinned_out, "synth_s": t_gen,
               "slab_sha256_match_reference": flags[0], "table_sha256_match_reference": flags[1],
               "roundtrip_exact": flags[2], "assembled_lac_sha256_match_reference": flags[3],
               "reference_sha256": g["sha256"]}
        assert all(flags), f"config 4 parity failed: slab/table/roundtrip/file = {flags}"
    return res


def kernel_src_sha() -> str:
    """SHA-256 over the CUDA sources of the product (what a profile capture is keyed by)."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted((ROOT / "lossless-audio-codec_b200" / "csrc").glob("*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def lacb_header(channels, stereo_mode, rate, depth) -> bytes:
    """frame/frame_header.hpp:25-36 (version 3)."""
    return bytes([0x4C, 0x41, 3, channels, stereo_mode, (rate >> 8) & 0xFF, rate & 0xFF, (rate >> 16) & 0xFF, depth, 0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=int, default=600, help="audio seconds per GPU (configs[1] = 600)")
    ap.add_argument("--cpu-seconds", type=int, default=0,
                    help="bound the CPU legs to the first N audio seconds (0 = the whole workload, the default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c4", choices=["auto", "on", "off"], default="auto",
                    help="config-4 leg (the ONE 10 h file sharded over the ranks, SHA-checked against the reference): "
                         "auto = only when --gpus > 1")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only: the line then has no e2e)")
    ap.add_argument("--e2e-contexts", type=int, default=1, help="contexts (host threads) the e2e leg splits the blocks over")
    args = ap.parse_args()
    quiet_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from __graft_entry__ import load_package
    lacb = load_package()
    cd = lacb.Codec(local_rank)  # raises if liblac_b200.so or the GPU is missing: no fallback

    frames = RATE * args.seconds
    pk = synth_packed(2 + rank, frames)
    pcm_bytes = pk.size
    nb = (frames + 16383) // 16384
    sizes = np.full(nb, 16384, dtype=np.uint32)
    sizes[-1] = frames - 16384 * (nb - 1)

    d_pcm = cd.dev_malloc(pcm_bytes)
    d_out = cd.dev_malloc(pcm_bytes)
    cd.h2d(d_pcm, pk)
    if dist:
        counts = torch.zeros(world, dtype=torch.int64, device="cuda")
        mine = torch.zeros(1, dtype=torch.int64, device="cuda")

    def barrier():
        if dist:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    stats = {"enc_ms": 0.0, "dec_ms": 0.0, "analyze_ms": 0.0, "parse_ms": 0.0, "lac_bytes": 0, "enc_wall": 0.0,
             "dec_wall": 0.0, "stages": {}}

    def step_device(record):
        t0 = time.perf_counter()
        d_payload, nbytes, d_bb = cd.encode_device(d_pcm, 0, frames, DEPTH, CHANNELS, STEREO_MODE)
        te = cd.timing()
        bb = cd.d2h(d_bb, nb * 4, np.uint32)
        if dist:  # global payload offsets = exclusive scan of the gathered byte counts (C1)
            mine[0] = nbytes
            dist.all_gather_into_tensor(counts, mine)
            torch.cuda.synchronize()
        t1 = time.perf_counter()
        cd.decode_device(d_payload, nbytes, sizes, bb, DEPTH, CHANNELS, STEREO_MODE, d_packed=d_out)
        td = cd.timing()
        t2 = time.perf_counter()
        if record:
            stats["enc_ms"] += te["total_ms"]
            stats["dec_ms"] += td["total_ms"]
            stats["analyze_ms"] += te["analyze_ms"]
            stats["parse_ms"] += td["parse_ms"]
            stats["enc_wall"] += t1 - t0
            stats["dec_wall"] += t2 - t1
            stats["lac_bytes"] = nbytes
            for k, v in te.items():
                stats["stages"]["enc_" + k] = stats["stages"].get("enc_" + k, 0.0) + v
            for k, v in td.items():
                if k in ("parse_ms", "restore_ms", "finish_ms", "h2d_ms", "total_ms"):
                    stats["stages"]["dec_" + k] = stats["stages"].get("dec_" + k, 0.0) + v
        return nbytes, bb

    sampler = ClockSampler(local_rank)  # started before the warm-up: nvidia-smi needs ~0.1 s to deliver its first line
    sampler.start()
    for _ in range(args.warmup):
        step_device(False)
    # correctness of the very data being timed: round trip restores the PCM bit for bit
    back = cd.d2h(d_out, pcm_bytes)
    assert np.array_equal(back, pk), "device round trip does not restore the PCM"
    del back

    barrier()
    m0 = sampler.mark()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device(True)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop(m0, sampler.mark())

    # end to end through the C ABI with HOST buffers: every step copies the PCM host->device,
    # the payload device->host, the payload host->device again and the PCM device->host
    # (page-locked buffers owned by the caller, allocated once outside the timed region)
    h_in = cd.pinned(pcm_bytes)
    h_in[:] = pk
    h_payload = cd.pinned(pcm_bytes + (pcm_bytes >> 2) + 4096)
    h_out = cd.pinned(pcm_bytes)
    h_bb = np.zeros(nb, dtype=np.uint32)
    e2e_steps = max(1, min(args.steps, 5))
    if args.no_e2e:
        e2e_steps = 0
    # The block range is split over `--e2e-contexts` contexts on the same GPU, one host thread
    # each (the --threads of the GPU path): while one context's kernels run, the other's PCIe
    # copies proceed, the same overlap the reference gets from its worker pool.
    nctx = max(1, min(args.e2e_contexts, nb))
    cds = [cd] + [lacb.Codec(local_rank) for _ in range(nctx - 1)]
    per = (nb + nctx - 1) // nctx
    fb = CHANNELS * (DEPTH // 8)
    parts = []
    pay_off = 0
    for i in range(nctx):
        b0, b1 = min(nb, i * per), min(nb, (i + 1) * per)
        f0, f1 = b0 * 16384, min(frames, b1 * 16384)
        cap = (f1 - f0) * fb + ((f1 - f0) * fb >> 2) + 4096
        parts.append((b0, b1, f0 * fb, f1 * fb, pay_off, cap))
        pay_off += cap
    sizes_out = [0] * nctx

    e2e_split = [0.0, 0.0]  # wall seconds inside the encode / decode calls of context 0 (timed steps only)

    def work(i):
        b0, b1, p0, p1, po, cap = parts[i]
        c = cds[i]
        t_a = time.perf_counter()
        n = c.encode_into(h_in[p0:p1], h_payload[po:po + cap], h_bb[b0:b1], DEPTH, CHANNELS, STEREO_MODE)
        t_b = time.perf_counter()
        c.decode_into(h_payload[po:po + n], sizes[b0:b1], h_bb[b0:b1], DEPTH, CHANNELS, STEREO_MODE, h_out[p0:p1])
        if i == 0:
            e2e_split[0] += t_b - t_a
            e2e_split[1] += time.perf_counter() - t_b
        sizes_out[i] = n

    def step_host():
        if nctx == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(i,)) for i in range(nctx)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        return sum(sizes_out)

    lac_e2e = 0
    if e2e_steps:
        lac_e2e = step_host()
        assert np.array_equal(h_out, pk), "host round trip does not restore the PCM"
    barrier()
    e2e_split[0] = e2e_split[1] = 0.0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lac_e2e = step_host()
    barrier()
    e2e_wall = time.perf_counter() - t0

    # bare-copy ceiling of the e2e leg: the same page-locked buffers and the same bytes per step (PCM in, payload
    # out, payload in, PCM out) with no kernel in between, all ranks at once -- what the host's PCIe / memory
    # fabric alone allows; e2e is reported as a fraction of it
    copy_wall = 0.0
    if e2e_steps:
        d_tmp = cd.dev_malloc(pcm_bytes)
        lac_n = int(lac_e2e)

        def copy_step():
            cd.h2d(d_tmp, h_in)
            cd.d2h_to(h_payload.ctypes.data, d_tmp, lac_n)
            cd.h2d(d_tmp, h_payload[:lac_n])
            cd.d2h_to(h_out.ctypes.data, d_tmp, pcm_bytes)

        copy_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copy_step()
        barrier()
        copy_wall = time.perf_counter() - t0
        cd.dev_free(d_tmp)

    if dist:
        t = torch.tensor([wall, e2e_wall, copy_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, e2e_wall, copy_wall = float(t[0]), float(t[1]), float(t[2])

    c4 = None
    if args.c4 == "on" or (args.c4 == "auto" and world > 1):
        if not dist:
            import torch
        c4 = run_c4(args, cd, rank, world, dist, torch)

    if rank == 0:
        K = args.steps
        lac = stats["lac_bytes"]
        value = pcm_bytes * world * K / wall / 1e9
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM bytes of one k_analyze launch: from the committed `ncu --set full` capture of this same command
        # (profiles/r2_roofline_traffic.json, written by tools/ncu_traffic.py), but only while the kernel sources are
        # the ones that capture was taken from (SHA-256 of csrc/); after any kernel change this is null until re-captured
        traffic = None
        try:
            tj = json.loads((ROOT / "profiles" / "r2_roofline_traffic.json").read_text())
            if tj.get("frames_per_gpu") == frames and tj.get("kernel_src_sha256") == kernel_src_sha():
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        analyze_s = stats["analyze_ms"] / K / 1e3
        parse_s = stats["parse_ms"] / K / 1e3
        achieved = (pcm_bytes + lac) / analyze_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": wall / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/int64 (u8 bitstream)", "data": "synthetic",
            "config": workload_config(args, args.seconds),
            "encode_gbs": pcm_bytes * world * K / stats["enc_wall"] / 1e9,
            "decode_gbs": pcm_bytes * world * K / stats["dec_wall"] / 1e9,
            "encode_gbs_device_events": pcm_bytes * K / (stats["enc_ms"] / 1e3) / 1e9,
            "decode_gbs_device_events": pcm_bytes * K / (stats["dec_ms"] / 1e3) / 1e9,
            "compression_ratio": lac / pcm_bytes,
            "stage_ms_per_step": {k: round(v / K, 4) for k, v in sorted(stats["stages"].items())},
            "roofline": {"bound": "hbm", "kernel": "k_analyze<1024,16> (encoder channel-block search)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": pcm_bytes + lac},
            "roofline_decode": {"bound": "hbm", "kernel": "k_parse_blocks (serial bitstream parse, one warp per block)",
                                "achieved": (pcm_bytes + lac) / parse_s / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": (pcm_bytes + lac) / parse_s / 1e9 / peak},
            "e2e": None if not e2e_steps else {"value": pcm_bytes * world * e2e_steps / e2e_wall / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": pcm_bytes + lac_e2e, "d2h_bytes_per_step": lac_e2e + pcm_bytes,
                    "steps": e2e_steps, "contexts": nctx,
                    "encode_ms_per_step": e2e_split[0] / e2e_steps * 1e3,  # rank 0's calls, wall clock
                    "decode_ms_per_step": e2e_split[1] / e2e_steps * 1e3,
                    "copy_ceiling": pcm_bytes * world * e2e_steps / copy_wall / 1e9,
                    "frac_of_copy_ceiling": copy_wall / e2e_wall,
                    "copy_ceiling_note": "same pinned buffers and bytes per step (H2D PCM, D2H payload, H2D payload, D2H PCM), "
                                         "no kernels, serial copies on one stream per rank, all ranks at once"},
            # per step: k_deinterleave, k_plan_fixed, k_build_jobs, k_autocorr, k_levinson, k_analyze,
            # k_finalize_blocks, k_emit (encode) + k_parse_blocks, k_restore_order, k_restore_blocks,
            # k_merge_restore_errors, k_finish_pcm (decode)
            "gpu_launches": K * 13,
            "clocks": clocks,
        }
        if c4:
            line["c4"] = c4
        if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
            codec, kind = cpu_codec()
            cores = os.cpu_count() or 1
            l, r, sb, secs, what = cpu_sample(args)
            te, td, _ = cpu_roundtrip(codec, kind, l, r, cores)
            te2, td2, _ = cpu_roundtrip(codec, kind, l, r, cores)
            te, td = min(te, te2), min(td, td2)
            line["cpu_baseline"] = {"value": sb / (te + td) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                    "encode_gbs": sb / te / 1e9, "decode_gbs": sb / td / 1e9,
                                    "sample": f"{what}, best of 2, timed around LAC::Encoder::encode / LAC::Decoder::decode"}
        emit_line(line)
    cd.dev_free(d_pcm)
    cd.dev_free(d_out)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
