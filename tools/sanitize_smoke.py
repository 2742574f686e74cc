"""Small encode + decode through the C ABI for compute-sanitizer runs (memcheck / racecheck):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Stereo auto (probe gangs, stereo proxy), forced MS and mono, full blocks and a short last block; checked against the oracle."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import helpers as H

cd = H.gpu_codec()
for seed, frames, depth, mode, ch in ((1, 5 * 16384 + 777, 16, 2, 2), (2, 2 * 16384 + 5, 24, 1, 2), (3, 16384 + 100, 24, 0, 1)):
    l, r = H.synth(seed, frames, depth)
    if ch == 1:
        r = None
    got = cd.encode(l, r, 48000, depth, mode) if ch == 2 else cd.encode(l, None, 48000, depth, 0)
    want = H.oracle().encode(l, r, 48000, depth, mode) if ch == 2 else H.oracle().encode(l, None, 48000, depth, 0)
    assert got == want, (seed, len(got), len(want))
    out = cd.decode(got)
    assert np.array_equal(out[0], l)
    print("ok", seed, frames, depth, mode, len(got), flush=True)
