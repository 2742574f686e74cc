"""Pins the C restatement (oracle/lac_oracle.c) against the compiled, unmodified
reference (oracle/_ref/liblac_ref.so) on the deterministic corpus.  Skipped when
the reference build is absent (e.g. on a box without /root/reference); the golden
fixtures in tests/golden/ cover that case (test_oracle_golden.py)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")


@pytest.mark.parametrize("name", sorted(H.block_corpus().keys()))
def test_block_bytes_identical(name):
    pcm = H.block_corpus()[name]
    for zr, part in ((1, 1), (0, 1), (1, 0), (0, 0)):
        a = H.oracle().block_encode(pcm, zr, part)
        b = H.ref().block_encode(pcm, zr, part)
        assert a == b, f"{name} zr={zr} part={part}: oracle {len(a)}B vs ref {len(b)}B"
        ok, dec, bits = H.oracle().block_decode(a, len(pcm))
        ok2, dec2, bits2 = H.ref().block_decode(a, len(pcm))
        assert ok == ok2 and bits == bits2
        if name != "int32_noise_512":  # outside the PCM domain: residual wraps, both decoders reject
            assert ok and bits == 8 * len(a)
            assert np.array_equal(dec, pcm) and np.array_equal(dec2, pcm)


@pytest.mark.parametrize("name", sorted(H.stereo_corpus().keys()))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frame_bytes_identical(name, mode):
    l, r, depth = H.stereo_corpus()[name]
    a = H.oracle().encode(l, r, 48000, depth, mode)
    b = H.ref().encode(l, r, 48000, depth, mode, threads=4)
    assert a == b
    dl, dr, hdr = H.oracle().decode(a)
    rl, rr, rhdr = H.ref().decode(a, threads=2)
    assert hdr == rhdr and hdr["stereo_mode"] == mode
    assert np.array_equal(dl, l) and np.array_equal(dr, r)
    assert np.array_equal(rl, l) and np.array_equal(rr, r)


def test_mono_frame():
    l, _ = H.synth(3, 40000, 24, channels=1)
    a = H.oracle().encode(l, None, 192000, 24, 0)
    b = H.ref().encode(l, None, 192000, 24, 0)
    assert a == b
    dl, dr, hdr = H.oracle().decode(a)
    assert hdr["channels"] == 1 and dr.size == 0 and np.array_equal(dl, l)


def test_lpc_coefficients_identical():
    for name, pcm in H.block_corpus().items():
        for order in (4, 6, 8, 10, 12):
            if order > len(pcm) - 1:
                continue
            ua, ca = H.oracle().lpc_analyze(pcm, order)
            ub, cb = H.ref().lpc_analyze(pcm, order)
            assert ua == ub and np.array_equal(ca, cb), (name, order)


def _mutations(stream: bytes, rng, count):
    for _ in range(count):
        b = bytearray(stream)
        kind = rng.integers(0, 4)
        if kind == 0:
            i = rng.integers(0, len(b))
            b[i] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            b = b[: rng.integers(1, len(b))]
        elif kind == 2:
            b += bytes(rng.integers(0, 256, rng.integers(1, 4), dtype=np.uint8))
        else:
            i = rng.integers(0, len(b))
            b[i] = int(rng.integers(0, 256))
        yield bytes(b)


def test_block_decoder_accepts_and_rejects_like_reference():
    """Malformed channel blocks: same accept/reject verdict, same samples, same bit count."""
    rng = np.random.default_rng(5)
    corpus = H.block_corpus()
    for name in ("rand_amp1000_4096", "sparse_4096", "bin_fallback_64", "mixed_runs_spikes_2048",
                 "noise_n400", "zr_sweep_n256", "pm2_4096", "ar4_16384"):
        pcm = corpus[name]
        good = H.oracle().block_encode(pcm)
        for bad in _mutations(good, rng, 60):
            ok_a, dec_a, bits_a = H.oracle().block_decode(bad, len(pcm))
            ok_b, dec_b, bits_b = H.ref().block_decode(bad, len(pcm))
            assert ok_a == ok_b, name
            if ok_a:
                assert bits_a == bits_b and np.array_equal(dec_a, dec_b)


def test_frame_decoder_rejections_match_reference():
    rng = np.random.default_rng(6)
    l, r, depth = H.stereo_corpus()["walk_plus_noise"]
    good = H.oracle().encode(l[:20000], r[:20000], 44100, depth, 2)
    for bad in _mutations(good, rng, 150):
        try:
            out_a = H.oracle().decode(bad)
            err_a = None
        except RuntimeError as e:
            out_a, err_a = None, str(e)
        try:
            out_b = H.ref().decode(bad)
            err_b = None
        except RuntimeError as e:
            out_b, err_b = None, str(e)
        assert err_a == err_b
        if out_a is not None:
            assert np.array_equal(out_a[0], out_b[0]) and np.array_equal(out_a[1], out_b[1])


def test_config1_size_matches_survey():
    """60 s 16/44.1 stereo seed 1, auto stereo -> 5 217 578 bytes."""
    l, r = H.synth(1, 2_646_000, 16)
    b = H.ref().encode(l, r, 44100, 16, 2, threads=8)
    assert len(b) == 5_217_578


def test_ranges_concatenate_to_whole_file():
    """Blocks are independent: the reference's encodes of contiguous block ranges, assembled as
    header + table + slabs in range order (src/codec/lac/encoder.cpp:445-465), are its encode of the whole file.
    This is what lets tools/make_golden_large.py record config 4 (one 10 h file) range by range."""
    import struct
    frames = 37 * 16384 + 4321
    l, r = H.synth_range(4, 0, frames, 24)
    whole = H.ref().encode(l, r, 48000, 24, 2, threads=4)
    nb = (frames + 16383) // 16384
    for world in (2, 3, 8):
        per = (nb + world - 1) // world
        tables, slabs = [], []
        for k in range(world):
            b0, b1 = min(nb, k * per), min(nb, (k + 1) * per)
            if b0 == b1:
                continue
            f0, f1 = b0 * 16384, min(frames, b1 * 16384)
            rl, rr = H.synth_range(4, f0, f1 - f0, 24)  # the rank generates its own range of the one file
            part = H.ref().encode(rl, rr, 48000, 24, 2, threads=2)
            n = struct.unpack(">I", part[10:14])[0]
            assert n == b1 - b0 and part[:10] == whole[:10]
            tables.append(part[14:14 + 8 * n])
            slabs.append(part[14 + 8 * n:])
        assert whole == whole[:10] + struct.pack(">I", nb) + b"".join(tables) + b"".join(slabs)
    # the C restatement agrees on the same file
    assert H.oracle().encode(l, r, 48000, 24, 2) == whole
