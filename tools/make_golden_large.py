#!/usr/bin/env python
"""Reference SHA-256 fixtures for the two BASELINE configs that only exist at full size:

  C3  30 min 24-bit / 192 kHz mono (345 600 000 frames, seed 3, Appendix C stream) -- whole .lac
  C4  10 h 24-bit / 48 kHz stereo, auto LR/MS (1 728 000 000 frames, seed 4, range-addressable
      stream of tools/lac_synth.c: lac_synth_range, reset_log2 = 19), encoded by the UNMODIFIED
      reference (oracle/_ref/liblac_ref.so, LAC::Encoder::encode -- the CLI refuses both)
      in eight block ranges; recorded are the SHA-256 of every rank's payload slab and table slice
      for the 1 / 2 / 4 / 8-way block-range sharding (ceil(n_blocks / N) blocks per rank, the product's
      plan_shards) and of the assembled 10 h .lac (header + table + slabs in rank order,
      src/codec/lac/encoder.cpp:445-465).

Blocks are independent, so the reference's encode of a block range IS that range of its
whole-file encode; tests/test_oracle_vs_ref.py::test_ranges_concatenate_to_whole_file pins that on a
short multi-range file.  Writes tests/golden/golden_large.json (a few KB).  ~6 minutes on 8 cores.

    python tools/make_golden_large.py [--only c3|c4]
"""
import argparse
import ctypes as C
import hashlib
import json
import struct
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402

MAX_BLOCK = 16384
C4 = dict(seed=4, frames=1_728_000_000, depth=24, rate=48000, channels=2, stereo_mode=2, reset_log2=H.C4_RESET_LOG2)
C3 = dict(seed=3, frames=345_600_000, depth=24, rate=192000, channels=1, stereo_mode=0)


def ref_encode_raw(ref, l, r, rate, depth, mode, threads):
    """(table bytes [n x 8], payload uint8 array) of the reference's encode of one range."""
    out, n = H.u8p(), C.c_uint64()
    rc = ref.lib.ref_encode(l.ctypes.data_as(H.i32p), r.ctypes.data_as(H.i32p) if r is not None else None, l.size, rate,
                            depth, mode, 1, 1, threads, C.byref(out), C.byref(n))
    assert rc == 0, ref.last_error()
    blob = np.ctypeslib.as_array(out, shape=(n.value,))
    nb = struct.unpack(">I", blob[10:14].tobytes())[0]
    header = blob[:10].tobytes()
    table = blob[14:14 + 8 * nb].tobytes()
    payload = blob[14 + 8 * nb:].copy()
    ref.free(out)
    return header, table, payload


def shard_plan(nb, world):
    per = (nb + world - 1) // world
    return [(min(nb, r * per), min(nb, (r + 1) * per)) for r in range(world)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["c3", "c4"])
    ap.add_argument("--threads", type=int, default=8)
    args = ap.parse_args()
    assert H.have_ref(), "oracle/_ref/liblac_ref.so is missing: run `make -C oracle ref` where /root/reference exists"
    ref = H.ref()
    path = ROOT / "tests" / "golden" / "golden_large.json"
    out = json.loads(path.read_text()) if path.exists() else {}
    out["generator"] = "tools/make_golden_large.py"
    out["source"] = "unmodified reference, oracle/_ref/liblac_ref.so (LAC::Encoder::encode)"

    if args.only in (None, "c3"):
        t0 = time.time()
        l, _ = H.synth(C3["seed"], C3["frames"], C3["depth"], 1)
        header, table, payload = ref_encode_raw(ref, l, None, C3["rate"], C3["depth"], 0, args.threads)
        nb = len(table) // 8
        h = hashlib.sha256(header + struct.pack(">I", nb) + table)
        h.update(payload)
        out["C3_full_1800s_24_192k_mono"] = dict(C3, len=14 + len(table) + payload.size, sha256=h.hexdigest(),
                                                 n_blocks=nb, payload_bytes=int(payload.size),
                                                 payload_sha256=hashlib.sha256(payload).hexdigest(),
                                                 table_sha256=hashlib.sha256(table).hexdigest())
        print("C3", out["C3_full_1800s_24_192k_mono"], f"{time.time() - t0:.0f} s", flush=True)
        del l, payload

    if args.only in (None, "c4"):
        t0 = time.time()
        frames = C4["frames"]
        nb = (frames + MAX_BLOCK - 1) // MAX_BLOCK
        tables, payloads, header = [], [], None
        for b0, b1 in shard_plan(nb, 8):
            f0, f1 = b0 * MAX_BLOCK, min(frames, b1 * MAX_BLOCK)
            l, r = H.synth_range(C4["seed"], f0, f1 - f0, C4["depth"], 2, C4["reset_log2"])
            header, table, payload = ref_encode_raw(ref, l, r, C4["rate"], C4["depth"], C4["stereo_mode"], args.threads)
            assert len(table) == 8 * (b1 - b0)
            tables.append(table)
            payloads.append(payload)
            print(f"C4 blocks [{b0}, {b1}): {payload.size} payload bytes, {time.time() - t0:.0f} s", flush=True)
            del l, r
        table = b"".join(tables)
        tab = np.frombuffer(table, dtype=">u4").reshape(nb, 2)
        assert int(tab[:, 0].astype(np.uint64).sum()) == frames
        boff = np.zeros(nb + 1, dtype=np.uint64)
        np.cumsum(tab[:, 1].astype(np.uint64), out=boff[1:])
        payload = np.concatenate(payloads)
        del payloads
        assert payload.size == int(boff[-1])
        rec = dict(C4, n_blocks=nb, payload_bytes=int(payload.size), table_sha256=hashlib.sha256(table).hexdigest(),
                   shards={})
        h = hashlib.sha256(header + struct.pack(">I", nb) + table)
        h.update(payload)
        rec["len"] = 14 + len(table) + int(payload.size)
        rec["sha256"] = h.hexdigest()
        for world in (1, 2, 4, 8):
            ranks = []
            for b0, b1 in shard_plan(nb, world):
                slab = payload[int(boff[b0]):int(boff[b1])]
                ranks.append(dict(first_block=b0, blocks=b1 - b0, payload_bytes=int(slab.size),
                                  payload_sha256=hashlib.sha256(slab).hexdigest(),
                                  table_sha256=hashlib.sha256(table[8 * b0:8 * b1]).hexdigest()))
            rec["shards"][str(world)] = ranks
        out["C4_full_10h_24_48k_auto"] = rec
        print("C4", {k: v for k, v in rec.items() if k != "shards"}, f"{time.time() - t0:.0f} s", flush=True)

    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
