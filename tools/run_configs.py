"""Runs the five BASELINE.json configurations on cuda:0 and prints one JSON line per case.

Not the graded bench (that is bench.py on configs[1]); this records the other configs'
device-resident throughput and checks the size-independent property decode(encode(x)) == x
at full size.  Inputs are the SURVEY.md Appendix C generator.  Config 4 (10 h sharded over
8 GPUs) is represented by one GPU's 75-minute shard; config 3 runs at full size through
the library path (the reference CLI rejects it).

usage: python tools/run_configs.py [--quick] [--out profiles/file.jsonl]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package  # noqa: E402


def synth_packed(seed, frames, depth, channels):
    lib = C.CDLL(str(ROOT / "tools" / "liblac_synth.so"))
    lib.lac_synth.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lac_synth.restype = None
    out = np.zeros(frames * channels * (depth // 8), dtype=np.uint8)
    lib.lac_synth(seed, frames, depth, channels, None, None, out.ctypes.data)
    return out


ONLY = ""


def run_case(cd, name, seed, seconds, rate, depth, channels, mode, reps=3, decode=True):
    if ONLY and ONLY not in name:
        return None
    frames = rate * seconds
    pk = synth_packed(seed, frames, depth, channels)
    nb = (frames + 16383) // 16384
    sizes = np.full(nb, 16384, dtype=np.uint32)
    sizes[-1] = frames - 16384 * (nb - 1)
    d_in = cd.dev_malloc(pk.size)
    d_out = cd.dev_malloc(pk.size)
    cd.h2d(d_in, pk)
    enc_ms, dec_ms, stages = [], [], {}
    for _ in range(reps):
        d_payload, nbytes, d_bb = cd.encode_device(d_in, 0, frames, depth, channels, mode)
        te = cd.timing()
        enc_ms.append(te["total_ms"])
        if decode:
            bb = cd.d2h(d_bb, nb * 4, np.uint32)
            cd.decode_device(d_payload, nbytes, sizes, bb, depth, channels, mode, d_packed=d_out)
            td = cd.timing()
            dec_ms.append(td["total_ms"])
            stages = {"parse_ms": td["parse_ms"], "restore_ms": td["restore_ms"], "finish_ms": td["finish_ms"]}
        stages.update({"analyze_ms": te["analyze_ms"], "emit_ms": te["emit_ms"], "lpc_ms": te["lpc_ms"],
                       "stereo_ms": te["stereo_ms"]})
    ok = None
    if decode:
        ok = bool(np.array_equal(cd.d2h(d_out, pk.size), pk))
    cd.dev_free(d_in)
    cd.dev_free(d_out)
    e = min(enc_ms)
    line = {"config": name, "seconds": seconds, "rate": rate, "depth": depth, "channels": channels,
            "stereo_mode": ["lr", "ms", "auto"][mode], "pcm_mb": pk.size / 1e6, "lac_mb": nbytes / 1e6,
            "ratio": nbytes / pk.size, "blocks": int(nb), "encode_ms": e, "encode_gbs": pk.size / e / 1e6,
            "roundtrip_exact": ok, "stage_ms": {k: round(v, 3) for k, v in stages.items()}}
    if decode:
        d = min(dec_ms)
        line.update({"decode_ms": d, "decode_gbs": pk.size / d / 1e6,
                     "roofline_frac_encode": (pk.size + nbytes) / (stages["analyze_ms"] / 1e3) / 1e9 / 6549.8,
                     "roofline_frac_decode": (pk.size + nbytes) / (stages["parse_ms"] / 1e3) / 1e9 / 6549.8})
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1/10 durations")
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="run only the configs whose name contains this text")
    args = ap.parse_args()
    q = 10 if args.quick else 1
    global ONLY
    ONLY = args.only
    cd = load_package().Codec(0, os.environ.get("LACB_LIB"))  # LACB_LIB: an experimental build of the library
    lines = []
    t0 = time.time()
    lines.append(run_case(cd, "C1 60 s 16/44.1 stereo auto", 1, 60, 44100, 16, 2, 2))
    lines.append(run_case(cd, "C2 10 min 24/96 stereo ms", 2, 600 // q, 96000, 24, 2, 1))
    lines.append(run_case(cd, "C3 30 min 24/192 mono (library path)", 3, 1800 // q, 192000, 24, 1, 0))
    lines.append(run_case(cd, "C4 shard: 75 min of the 10 h 24/48 stereo auto file (1 of 8 GPUs)", 4, 4500 // q, 48000, 24,
                          2, 2, reps=2))
    for depth in (16, 24):
        for rate in (44100, 48000, 96000, 192000):
            lines.append(run_case(cd, f"C5 decode sweep {depth}/{rate // 1000}k stereo auto", 5, 300 // q, rate, depth, 2, 2,
                                  reps=2))
    print(f"# total {time.time() - t0:.1f} s", flush=True)
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines if x) + "\n")


if __name__ == "__main__":
    main()
