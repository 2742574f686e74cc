"""Decode cost of UNPARTITIONED streams (every adaptive segment uses the stateful model, rice.hpp:45-114) against
the default partitioned streams of the same input: BASELINE configs[1] (600 s 24/96 forced M/S) encoded with
set_partitioning_enabled(false).  Prints one JSON line per variant (kernel stage times from CUDA events).
usage: python tools/nopart_decode.py [seconds]"""
import json, sys
sys.path.insert(0, "tests")
import numpy as np, helpers as H
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
cd = H.gpu_codec() if len(sys.argv) < 3 else H.lacb_module().Codec(0, sys.argv[2])  # [lib.so]: a measurement build
cd.set_concurrency(1)  # one pass, so that the stage times of the whole decode are reported
l, r, pk = H.synth(2, 96000 * secs, 24, want_packed=True)
for name, part in (("partitioned (default)", True), ("unpartitioned (--no-partitioning)", False)):
    payload, bb, sizes = cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2, partitioning=part)
    best = None
    for _ in range(3):
        out, = cd.decode_blocks(payload, sizes, bb, 24, 2, 1, packed=True)
        t = cd.timing()
        if best is None or t["parse_ms"] < best["parse_ms"]:
            best = dict(t)
    assert np.array_equal(out, pk), "round trip differs"
    print(json.dumps({"stream": name, "pcm_mb": pk.size / 1e6, "lac_mb": round(payload.size / 1e6, 1), "blocks": int(sizes.size),
                      "parse_ms": round(best["parse_ms"], 3), "restore_ms": round(best["restore_ms"], 3),
                      "finish_ms": round(best["finish_ms"], 3),
                      "decode_gbs_kernels": round(pk.size / 1e6 / (best["parse_ms"] + best["restore_ms"] + best["finish_ms"]), 1)}), flush=True)
