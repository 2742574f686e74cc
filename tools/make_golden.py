#!/usr/bin/env python
"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/liblac_ref.so,
compiled from /root/reference by oracle/Makefile).  The reference ships no golden vectors of
its own, so these are outputs of the reference itself run in the build
container: for every entry of the deterministic parity corpus (tests/helpers.py) the SHA-256
and length of the bytes the reference produces, plus a few short streams verbatim (hex).

The fixtures travel to the GPU box, where /root/reference does not exist; tests/test_golden.py
checks the C oracle (CPU) and the CUDA path (GPU) against them.

    python tools/make_golden.py        # rewrites tests/golden/golden.json
"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402


def digest(b: bytes):
    return {"len": len(b), "sha256": hashlib.sha256(b).hexdigest()}


def main():
    if not H.have_ref():
        raise SystemExit("oracle/_ref/liblac_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
    ref = H.ref()
    out = {"generator": "tools/make_golden.py", "source": "unmodified reference, oracle/_ref/liblac_ref.so",
           "blocks": {}, "frames": {}, "synthetic": {}, "verbatim": {}}
    # Block::Encoder::encode for every corpus block under the four flag combinations
    for name, pcm in H.block_corpus().items():
        for zr in (True, False):
            for part in (True, False):
                b = ref.block_encode(pcm, zero_run=zr, partitioning=part)
                out["blocks"][f"{name}|zr={int(zr)}|part={int(part)}"] = digest(bytes(b))
    # LAC::Encoder::encode for the stereo corpus under the three stereo modes, and mono
    for name, (l, r, depth) in H.stereo_corpus().items():
        for mode in (0, 1, 2):
            b = ref.encode(l, r, 48000 if depth == 24 else 44100, depth, mode)
            out["frames"][f"{name}|mode={mode}"] = digest(bytes(b))
        b = ref.encode(l, None, 48000 if depth == 24 else 44100, depth, 0)
        out["frames"][f"{name}|mono"] = digest(bytes(b))
    # the BASELINE configs at reduced length (SURVEY Appendix C generator): seed, frames, depth, rate, channels, mode
    for key, (seed, frames, depth, rate, ch, mode) in {
        "C1_60s_16_44k_auto": (1, 2646000, 16, 44100, 2, 2),
        "C2_10s_24_96k_ms": (2, 960000, 24, 96000, 2, 1),
        "C3_5s_24_192k_mono": (3, 960000, 24, 192000, 1, 0),
        "C4_10s_24_48k_auto": (4, 480000, 24, 48000, 2, 2),
        "C2_full_600s_24_96k_ms": (2, 57600000, 24, 96000, 2, 1),  # BASELINE configs[1] at full size (213 135 210 bytes)
    }.items():
        l, r = H.synth(seed, frames, depth, ch)
        b = ref.encode(l, r if ch == 2 else None, rate, depth, mode, threads=8)
        out["synthetic"][key] = dict(digest(bytes(b)), seed=seed, frames=frames, depth=depth, rate=rate,
                                     channels=ch, stereo_mode=mode)
    # a few short streams verbatim, so a mismatch can be diffed without the reference
    corpus = H.block_corpus()
    for name in ("bin_fallback_64", "zeros_17", "noise_n33", "zr_sweep_n96"):
        out["verbatim"][f"block|{name}"] = bytes(ref.block_encode(corpus[name])).hex()
    l, r, depth = H.stereo_corpus()["short_17_24bit"]
    out["verbatim"]["frame|short_17_24bit|mode=2"] = bytes(ref.encode(l, r, 48000, depth, 2)).hex()
    path = ROOT / "tests" / "golden" / "golden.json"
    path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", path, {k: len(v) for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
