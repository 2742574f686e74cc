"""Timing experiment helper: encode-only stage times with an alternative build of the library."""
import sys
sys.path.insert(0, "tests")
import numpy as np, helpers as H
lib = sys.argv[1]
cd = H.lacb_module().Codec(0, lib)
l, r, pk = H.synth(2, 96000 * 120, 24, want_packed=True)
for it in range(3):
    try:
        cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2)
    except Exception as e:
        print("err", e)
    t = cd.timing()
print(lib.split("/")[-1], {k: round(v, 3) for k, v in t.items() if k in ("lpc_ms", "analyze_ms", "emit_ms")})
