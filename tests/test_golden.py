"""Golden fixtures: outputs of the UNMODIFIED reference (oracle/_ref/liblac_ref.so), produced
in the build container by tools/make_golden.py and committed under tests/golden/.  They pin
the C oracle on any machine (CPU test) and the CUDA path on the GPU box, where neither
/root/reference nor a rebuild of oracle/_ref is needed for the comparison."""
import hashlib
import json

import numpy as np
import pytest

import helpers as H

GOLD = json.loads((H.ROOT / "tests" / "golden" / "golden.json").read_text())


def _dig(b: bytes):
    return {"len": len(b), "sha256": hashlib.sha256(b).hexdigest()}


def _check_blocks(codec):
    corpus = H.block_corpus()
    bad = []
    for key, want in GOLD["blocks"].items():
        name, zr, part = key.split("|")
        got = _dig(bytes(codec.block_encode(corpus[name], int(zr[-1]), int(part[-1]))))
        if got != want:
            bad.append((key, got["len"], want["len"]))
    assert not bad, bad[:5]


def _check_frames(codec):
    corpus = H.stereo_corpus()
    bad = []
    for key, want in GOLD["frames"].items():
        name, what = key.split("|")
        l, r, depth = corpus[name]
        rate = 48000 if depth == 24 else 44100
        if what == "mono":
            got = codec.encode(l, None, rate, depth, 0)
        else:
            got = codec.encode(l, r, rate, depth, int(what[-1]))
        if _dig(bytes(got)) != want:
            bad.append((key, len(got), want["len"]))
    assert not bad, bad[:5]


def _check_synthetic(codec, keys):
    for key in keys:
        g = GOLD["synthetic"][key]
        l, r = H.synth(g["seed"], g["frames"], g["depth"], g["channels"])
        got = codec.encode(l, r if g["channels"] == 2 else None, g["rate"], g["depth"], g["stereo_mode"])
        assert _dig(bytes(got)) == {"len": g["len"], "sha256": g["sha256"]}, key
        dl, dr, _ = codec.decode(bytes(got))
        assert np.array_equal(dl, l) and (g["channels"] == 1 or np.array_equal(dr, r))


def _check_verbatim(codec):
    blocks, stereo = H.block_corpus(), H.stereo_corpus()
    for key, hexbytes in GOLD["verbatim"].items():
        kind, name = key.split("|")[:2]
        if kind == "block":
            got = bytes(codec.block_encode(blocks[name], 1, 1))
        else:
            l, r, depth = stereo[name]
            got = bytes(codec.encode(l, r, 48000, depth, 2))
        assert got.hex() == hexbytes, key


# --- CPU: the oracle restatement against the reference's recorded outputs -------------------
def test_oracle_blocks_match_golden():
    _check_blocks(H.oracle())


def test_oracle_frames_match_golden():
    _check_frames(H.oracle())


def test_oracle_synthetic_match_golden():
    _check_synthetic(H.oracle(), ["C2_10s_24_96k_ms", "C3_5s_24_192k_mono", "C4_10s_24_48k_auto"])


def test_oracle_verbatim_match_golden():
    _check_verbatim(H.oracle())


# --- GPU: the CUDA path through the C ABI against the same fixtures ---------------------------
@pytest.mark.gpu
def test_gpu_blocks_match_golden():
    _check_blocks(H.gpu_codec())


@pytest.mark.gpu
def test_gpu_frames_match_golden():
    _check_frames(H.gpu_codec())


@pytest.mark.gpu
def test_gpu_synthetic_match_golden():
    """Includes BASELINE configs[1] at its full size: 600 s of 24-bit / 96 kHz stereo, forced mid/side,
    345.6 MB of PCM -> the reference's 213 135 210 bytes, SHA-256 recorded from the unmodified reference."""
    _check_synthetic(H.gpu_codec(), sorted(GOLD["synthetic"].keys()))


@pytest.mark.gpu
def test_gpu_verbatim_match_golden():
    _check_verbatim(H.gpu_codec())
