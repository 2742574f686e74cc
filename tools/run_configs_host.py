"""Host-buffer (C-ABI, `lacb_encode_to` / `lacb_decode`) pass over the large BASELINE configs on cuda:0.

Companion of run_configs.py (which times the device-resident path): every case goes through the
sliced host pipelines with page-locked caller buffers, and is checked two ways at full size --
the payload and the per-block byte counts must equal those of the single-pass device path on the
same input, and decode(encode(x)) must equal x.  One JSON line per case.

usage: python tools/run_configs_host.py [--quick] [--out profiles/file.jsonl]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from __graft_entry__ import load_package  # noqa: E402
from run_configs import synth_packed  # noqa: E402


def run_case(cd, name, seed, seconds, rate, depth, channels, mode, reps=3):
    frames = rate * seconds
    src = synth_packed(seed, frames, depth, channels)
    nb = (frames + 16383) // 16384
    sizes = np.full(nb, 16384, dtype=np.uint32)
    sizes[-1] = frames - 16384 * (nb - 1)
    pcm = cd.pinned(src.size)
    pcm[:] = src
    payload = cd.pinned(src.size + 64 * nb + 4096)
    back = cd.pinned(src.size)
    bb = np.zeros(nb, dtype=np.uint32)
    enc_s, dec_s, n = [], [], 0
    for _ in range(reps):
        t0 = time.perf_counter()
        n = cd.encode_into(pcm, payload, bb, depth, channels, mode)
        t1 = time.perf_counter()
        cd.decode_into(payload[:n], sizes, bb, depth, channels, mode, back)
        t2 = time.perf_counter()
        enc_s.append(t1 - t0)
        dec_s.append(t2 - t1)
    roundtrip = bool(np.array_equal(back, src))
    # single-pass device path on the same input
    d_in = cd.dev_malloc(src.size)
    cd.h2d(d_in, src)
    d_payload, nbytes, d_bb = cd.encode_device(d_in, 0, frames, depth, channels, mode)
    same = nbytes == n and bool(np.array_equal(cd.d2h(d_bb, nb * 4, np.uint32), bb)) and \
        bool(np.array_equal(cd.d2h(d_payload, nbytes), payload[:n]))
    cd.dev_free(d_in)
    e, d = min(enc_s), min(dec_s)
    line = {"config": name, "pcm_mb": src.size / 1e6, "lac_mb": n / 1e6, "blocks": int(nb),
            "host_encode_ms": e * 1e3, "host_encode_gbs": src.size / e / 1e9,
            "host_decode_ms": d * 1e3, "host_decode_gbs": src.size / d / 1e9,
            "host_encode_decode_gbs": src.size / (e + d) / 1e9,
            "payload_equals_device_path": same, "roundtrip_exact": roundtrip}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="1/10 durations")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    q = 10 if args.quick else 1
    mod = load_package()
    lines = []
    for case in (("C1 60 s 16/44.1 stereo auto", 1, 60, 44100, 16, 2, 2),
                 ("C2 10 min 24/96 stereo ms", 2, 600 // q, 96000, 24, 2, 1),
                 ("C3 30 min 24/192 mono (library path)", 3, 1800 // q, 192000, 24, 1, 0),
                 ("C4 shard: 75 min of the 10 h 24/48 stereo auto file (1 of 8 GPUs)", 4, 4500 // q, 48000, 24, 2, 2),
                 ("C5 24/192k stereo auto, 300 s", 5, 300 // q, 192000, 24, 2, 2)):
        cd = mod.Codec(0)  # fresh context per case: the page-locked buffers go with it
        lines.append(run_case(cd, *case))
        cd.close()
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines) + "\n")
    return 0 if all(x["payload_equals_device_path"] and x["roundtrip_exact"] for x in lines) else 1


if __name__ == "__main__":
    sys.exit(main())
