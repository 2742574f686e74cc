"""CPU-only check of the product's CUDA sources: csrc/*.cu compiled against the fiber
emulator in tests/emu (test infrastructure) and compared with the oracle.  This is how
kernel logic is debugged where no GPU exists; the graded parity tests are the gpu-marked
ones, which run the nvcc-built library on a B200."""
import numpy as np
import pytest

import helpers as H

BLOCKS = ["noise_n1", "noise_n33", "noise_n257", "zr_sweep_n576", "bin_fallback_64", "sparse_4096", "noise_n8176",
          "ar4_16384", "level_steps_16384", "silence_then_noise_16384", "sine24_16384", "int32_noise_512",
          "random_walk_12288", "mixed_runs_spikes_2048"]


@pytest.fixture(scope="module")
def cd():
    return H.emu_codec()


@pytest.mark.parametrize("name", BLOCKS)
def test_block_bytes_identical(cd, name):
    pcm = H.block_corpus()[name]
    for zr, part in ((1, 1), (0, 1), (1, 0), (0, 0)):
        want = H.oracle().block_encode(pcm, zr, part)
        assert cd.block_encode(pcm, zr, part) == want, (name, zr, part)
    ok, dec, bits = cd.block_decode(want, len(pcm))
    ok2, dec2, bits2 = H.oracle().block_decode(want, len(pcm))
    assert (ok, bits) == (ok2, bits2) and (not ok or np.array_equal(dec, dec2))


@pytest.mark.parametrize("name", ["synth16", "synth24_sections", "short_17_24bit", "short_1024", "left_only"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frame_bytes_identical(cd, name, mode):
    l, r, depth = H.stereo_corpus()[name]
    want = H.oracle().encode(l, r, 48000, depth, mode)
    assert cd.encode(l, r, 48000, depth, mode) == want
    dl, dr, hdr = cd.decode(want)
    assert np.array_equal(dl, l) and np.array_equal(dr, r) and hdr["stereo_mode"] == mode


def test_lpc_coefficients_identical(cd):
    for name in ("ar4_16384", "sine24_16384", "noise_n33", "lownoise_16384", "ramp_16384"):
        pcm = H.block_corpus()[name]
        for order in (4, 6, 8, 10, 12):
            ua, ca = H.oracle().lpc_analyze(pcm, order)
            ub, cb = cd.lpc_analyze(pcm, order)
            assert ua == ub and np.array_equal(ca, cb), (name, order)


def test_random_blocks(cd):
    rng = np.random.default_rng(11)
    for it in range(40):
        n = int(rng.choice([rng.integers(1, 700), rng.integers(1, 16385), 256, 16384]))
        amp = 1 << int(rng.integers(0, 24))
        x = rng.integers(-amp, amp + 1, n)
        if it % 3 == 0:
            x[rng.random(n) < rng.random()] = 0
        if it % 5 == 0:
            x = np.cumsum(rng.integers(-amp // 64 - 1, amp // 64 + 2, n)).clip(-(1 << 23), (1 << 23) - 1)
        pcm = x.astype(np.int32)
        zr, part = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        want = H.oracle().block_encode(pcm, zr, part)
        assert cd.block_encode(pcm, zr, part) == want, (it, n, zr, part)
        ok, dec, _ = cd.block_decode(want, n)
        assert ok and np.array_equal(dec, pcm)


def test_decoder_error_messages(cd):
    rng = np.random.default_rng(6)
    l, r, depth = H.stereo_corpus()["walk_plus_noise"]
    good = H.oracle().encode(l[:20000], r[:20000], 44100, depth, 2)
    n = 0
    for _ in range(60):
        b = bytearray(good)
        i = int(rng.integers(10, len(b)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        try:
            a, ea = H.oracle().decode(bytes(b)), None
        except RuntimeError as e:
            a, ea = None, str(e)
        try:
            g, eg = cd.decode(bytes(b)), None
        except RuntimeError as e:
            g, eg = None, str(e)
        assert ea == eg
        n += ea is not None
        if a is not None:
            assert np.array_equal(a[0], g[0]) and np.array_equal(a[1], g[1])
    assert n >= 1


def test_block_decoder_fuzz_all_modes(cd):
    from test_gpu_parity import fuzz_block_decoder
    names = ["zr_sweep_n576", "bin_fallback_64", "sparse_4096", "mixed_runs_spikes_2048", "noise_n257", "pm2_4096",
             "int32_noise_512", "alternating_4096"]
    assert fuzz_block_decoder(cd, 7, names, 8) > 20


def test_restore_overflow_verdicts(cd):
    from test_gpu_parity import restore_overflow_fuzz
    assert restore_overflow_fuzz(cd, 8, 10) > 40


def test_lpc_restore_at_int32_limits(cd):
    from test_gpu_parity import lpc_fullscale_check
    accepted, rejected, peak = lpc_fullscale_check(cd, 120)
    assert accepted > 20 and rejected > 20 and peak > (1 << 30)


def test_sliced_host_pipeline():
    from test_gpu_parity import run_sliced
    run_sliced("emu_codec", 3)


def test_serial_v2_stream(cd):
    from test_gpu_parity import _to_v2
    l, r, depth = H.stereo_corpus()["synth16"]
    v2 = _to_v2(H.oracle().encode(l, r, 48000, depth, 2))
    dl, dr, hdr = cd.decode(v2)
    assert np.array_equal(dl, l) and np.array_equal(dr, r)
    with pytest.raises(RuntimeError, match="trailing frame payload"):
        cd.decode(v2 + b"\0")


def _shift_bits(data: bytes, k: int) -> bytes:
    """`data` behind k junk bits (ones), MSB first, zero padded to a byte"""
    v = (((1 << k) - 1) << (8 * len(data))) | int.from_bytes(data, "big")
    total = k + 8 * len(data)
    pad = (-total) % 8
    return (v << pad).to_bytes((total + pad) // 8, "big")


def test_block_decode_from_any_bit_position(cd):
    """Block::Decoder::decode_into reads from wherever the reader stands (block/decoder.cpp:64); a reject that ran
    out of data is told apart from a semantic one (the reference's reader is only then in its error state)."""
    corpus = H.block_corpus()
    for name in ("ar4_16384", "noise_n33", "zr_sweep_n96", "sparse_4096", "level_steps_16384"):
        pcm = corpus[name]
        blk = H.oracle().block_encode(pcm, True, True)
        for k in (1, 3, 7, 13):
            ok, out, bits, ran = cd.block_decode_at(_shift_bits(blk, k) + b"\xff" * 5, k, len(pcm))
            assert ok and not ran and np.array_equal(out, pcm), (name, k)
            assert (k + bits) % 8 == 0 and abs(bits - 8 * len(blk)) <= 7      # ends on a byte boundary of the buffer
        ok, _, _, ran = cd.block_decode_at(blk[: len(blk) // 2], 0, len(pcm))
        assert not ok and ran, name                      # truncated: data ran out
    bad = bytearray(H.oracle().block_encode(corpus["noise_n33"], True, True))
    bad[0] = 7                                            # predictor type 7: semantic reject, data is all there
    ok, _, _, ran = cd.block_decode_at(bytes(bad), 0, 33)
    assert not ok and not ran


def test_partition_levels_share_segment_starts(cd):
    """Fused partition-level sweep: a chunk keeps its costs from level to level only while (segment start, last-in-
    segment, initial k) are unchanged.  Triangle-section blocks of config 1 (two verbatim warm-up samples, then small
    residuals) have an initial k that changes once segments get shorter than 256 samples: the case that pins the key."""
    l, r = H.synth(1, 46 * 16384, 16)
    m = ((l.astype(np.int64) + r) >> 1).astype(np.int32)
    for b in (10, 44, 45):
        x = m[b * 16384:(b + 1) * 16384]
        assert cd.block_encode(x, True, True) == H.oracle().block_encode(x, True, True), b


def test_unpartitioned_file_round_trip(cd):
    """Unpartitioned streams: the stateful model by speculative batches (Rice / bin) and the serial reader (zero-run)."""
    for seed, depth in ((2, 24), (1, 16)):
        l, r = H.synth(seed, 3 * 16384 + 99, depth)
        want = H.oracle().encode(l, r, 48000, depth, 1, partitioning=False)
        assert cd.encode(l, r, 48000, depth, 1, partitioning=False) == want
        dl, dr, _ = cd.decode(want)
        assert np.array_equal(dl, l) and np.array_equal(dr, r)
