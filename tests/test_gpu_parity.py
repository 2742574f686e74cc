"""GPU parity tests: liblac_b200.so (sm_100a kernels through the C ABI) against the
C oracle on identical inputs.  Bit-exact is the bar: identical .lac bytes on encode,
identical PCM on decode, identical accept/reject verdicts on malformed streams."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cd():
    return H.gpu_codec()


@pytest.mark.parametrize("name", sorted(H.block_corpus().keys()))
def test_block_bytes_identical(cd, name):
    pcm = H.block_corpus()[name]
    for zr, part in ((1, 1), (0, 1), (1, 0), (0, 0)):
        want = H.oracle().block_encode(pcm, zr, part)
        got = cd.block_encode(pcm, zr, part)
        assert got == want, f"{name} zr={zr} part={part}: gpu {len(got)}B vs oracle {len(want)}B"
    ok, dec, bits = cd.block_decode(want, len(pcm))
    ok2, dec2, bits2 = H.oracle().block_decode(want, len(pcm))
    assert (ok, bits) == (ok2, bits2)
    if ok:
        assert np.array_equal(dec, dec2)


@pytest.mark.parametrize("name", sorted(H.stereo_corpus().keys()))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_frame_bytes_identical(cd, name, mode):
    l, r, depth = H.stereo_corpus()[name]
    want = H.oracle().encode(l, r, 48000, depth, mode)
    got = cd.encode(l, r, 48000, depth, mode)
    assert got == want
    dl, dr, hdr = cd.decode(want)
    assert hdr["stereo_mode"] == mode and hdr["bit_depth"] == depth
    assert np.array_equal(dl, l) and np.array_equal(dr, r)


def test_mono_frame(cd):
    l, _ = H.synth(3, 40000, 24, channels=1)
    want = H.oracle().encode(l, None, 192000, 24, 0)
    assert cd.encode(l, None, 192000, 24, 0) == want
    dl, dr, hdr = cd.decode(want)
    assert hdr["channels"] == 1 and dr.size == 0 and np.array_equal(dl, l)


def test_lpc_coefficients_identical(cd):
    mism = 0
    for name, pcm in H.block_corpus().items():
        for order in (4, 6, 8, 10, 12):
            if order > len(pcm) - 1:
                continue
            ua, ca = H.oracle().lpc_analyze(pcm, order)
            ub, cb = cd.lpc_analyze(pcm, order)
            mism += int(ua != ub or not np.array_equal(ca, cb))
    assert mism == 0


def test_packed_io_matches_planar(cd):
    for depth in (16, 24):
        l, r, pk = H.synth(5, 3 * 16384 + 777, depth, want_packed=True)
        p1, bb1, _ = cd.encode_blocks(l, r, depth, 2)
        p2, bb2, sz = cd.encode_blocks(None, None, depth, 2, packed=pk, channels=2)
        assert np.array_equal(p1, p2) and np.array_equal(bb1, bb2)
        out, = cd.decode_blocks(p2, sz, bb2, depth, 2, 2, packed=True)
        assert np.array_equal(out, pk)


def test_config1_full_file(cd):
    """BASELINE config 1: 60 s 16-bit 44.1 kHz stereo, auto LR/MS -- byte-identical to the
    CPU codec (5 217 578 bytes) and bit-exact round trip."""
    l, r = H.synth(1, 2_646_000, 16)
    got = cd.encode(l, r, 44100, 16, 2)
    assert len(got) == 5_217_578
    want = H.oracle().encode(l, r, 44100, 16, 2, threads=8)
    assert got == want
    dl, dr, _ = cd.decode(got)
    assert np.array_equal(dl, l) and np.array_equal(dr, r)


def test_config2_slice_forced_ms_24bit(cd):
    """A 30 s slice of config 2 (24-bit 96 kHz, forced MS) against the oracle, then the
    size-independent property on the same data: decode(encode(x)) == x."""
    l, r = H.synth(2, 96000 * 30, 24)
    got = cd.encode(l, r, 96000, 24, 1)
    assert got == H.oracle().encode(l, r, 96000, 24, 1, threads=8)
    dl, dr, _ = cd.decode(got)
    assert np.array_equal(dl, l) and np.array_equal(dr, r)


def test_config3_slice_mono_192k(cd):
    l, _ = H.synth(3, 192000 * 10, 24, channels=1)
    got = cd.encode(l, None, 192000, 24, 0)
    assert got == H.oracle().encode(l, None, 192000, 24, 0, threads=8)
    dl, _, _ = cd.decode(got)
    assert np.array_equal(dl, l)


def test_encoder_argument_errors(cd):
    l = np.zeros(100, np.int32)
    with pytest.raises(ValueError):
        cd.encode(np.zeros(0, np.int32), None)
    with pytest.raises(ValueError):
        cd.encode(l, np.zeros(99, np.int32))
    with pytest.raises(ValueError):
        cd.encode(l, None, sample_rate=12345)
    with pytest.raises(ValueError):
        cd.encode(l, None, bit_depth=20)
    with pytest.raises(ValueError):
        cd.encode(l, l, stereo_mode=3)
    bad = l.copy()
    bad[50] = 40000
    with pytest.raises(ValueError):
        cd.encode(bad, None, bit_depth=16)


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


def test_block_decoder_verdicts_match(cd):
    rng = np.random.default_rng(5)
    corpus = H.block_corpus()
    for name in ("rand_amp1000_4096", "sparse_4096", "bin_fallback_64", "mixed_runs_spikes_2048",
                 "noise_n400", "zr_sweep_n256", "pm2_4096", "ar4_16384"):
        pcm = corpus[name]
        good = H.oracle().block_encode(pcm)
        for bad in _mutations(good, rng, 40):
            ok_a, dec_a, bits_a = H.oracle().block_decode(bad, len(pcm))
            ok_b, dec_b, bits_b = cd.block_decode(bad, len(pcm))
            assert ok_a == ok_b, name
            if ok_a:
                assert bits_a == bits_b and np.array_equal(dec_a, dec_b)


def fuzz_block_decoder(cd, seed, names, flips):
    """Every residual mode (partitioned or not, zero-run on/off): intact, truncated, extended and
    bit-flipped streams must get the oracle's verdict, bit count and samples."""
    rng = np.random.default_rng(seed)
    corpus = H.block_corpus()
    rejected = 0
    for name in names:
        pcm = corpus[name]
        for zr, part in ((1, 1), (0, 1), (1, 0), (0, 0)):
            good = H.oracle().block_encode(pcm, zr, part)
            trials = [good, good[:-1], good[: len(good) // 2], good + b"\0"]
            for _ in range(flips):
                b = bytearray(good)
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                trials.append(bytes(b))
            for t in trials:
                ok_a, dec_a, bits_a = H.oracle().block_decode(t, len(pcm))
                ok_b, dec_b, bits_b = cd.block_decode(t, len(pcm))
                assert ok_a == ok_b, (name, zr, part, len(t))
                rejected += not ok_a
                if ok_a:
                    assert bits_a == bits_b and np.array_equal(dec_a, dec_b), (name, zr, part)
    return rejected


FUZZ_BLOCKS = ["zr_sweep_n576", "bin_fallback_64", "sparse_4096", "mixed_runs_spikes_2048", "noise_n257",
               "noise_n1000", "level_steps_16384", "pm2_4096", "lownoise_16384", "sparse_bursts_16384",
               "int32_noise_512", "zr_sweep_n320", "alternating_4096", "silence_then_noise_16384"]


def test_block_decoder_fuzz_all_modes(cd):
    assert fuzz_block_decoder(cd, 5, FUZZ_BLOCKS, 24) > 100


def test_frame_decoder_errors_match(cd):
    rng = np.random.default_rng(6)
    l, r, depth = H.stereo_corpus()["walk_plus_noise"]
    good = H.oracle().encode(l[:40000], r[:40000], 44100, depth, 2)
    for bad in _mutations(good, rng, 120):
        try:
            out_a, err_a = H.oracle().decode(bad), None
        except RuntimeError as e:
            out_a, err_a = None, str(e)
        try:
            out_b, err_b = cd.decode(bad), None
        except RuntimeError as e:
            out_b, err_b = None, str(e)
        assert err_a == err_b
        if out_a is not None:
            assert np.array_equal(out_a[0], out_b[0]) and np.array_equal(out_a[1], out_b[1])


def _to_v2(blob: bytes) -> bytes:
    """Same blocks as a legacy v2 frame: version byte 2, table of sample counts only."""
    import struct
    hdr = bytearray(blob[:10])
    hdr[2] = 2
    nb = struct.unpack(">I", blob[10:14])[0]
    tab = np.frombuffer(blob[14:14 + 8 * nb], dtype=">u4").reshape(nb, 2)
    return bytes(hdr) + struct.pack(">I", nb) + tab[:, 0].astype(">u4").tobytes() + blob[14 + 8 * nb:]


@pytest.mark.parametrize("name", ["synth16", "walk_plus_noise", "short_17_24bit"])
def test_serial_v2_streams(cd, name):
    """lac/decoder.cpp:209-218: v2 frames are one serial chain; same samples, same verdicts."""
    l, r, depth = H.stereo_corpus()[name]
    rng = np.random.default_rng(8)
    for mode in (0, 1, 2):
        v2 = _to_v2(H.oracle().encode(l, r, 48000, depth, mode))
        dl, dr, hdr = cd.decode(v2)
        assert np.array_equal(dl, l) and np.array_equal(dr, r)
        for bad in _mutations(v2, rng, 20):
            try:
                a, ea = H.oracle().decode(bad), None
            except RuntimeError as e:
                a, ea = None, str(e)
            try:
                g, eg = cd.decode(bad), None
            except RuntimeError as e:
                g, eg = None, str(e)
            assert ea == eg
            if a is not None:
                assert np.array_equal(a[0], g[0]) and np.array_equal(a[1], g[1])


SLICED_SCRIPT = """
import sys
sys.path.insert(0, {tests!r})
import numpy as np, helpers as H
cd = H.{codec}()
for depth, mode, ch in ((16, 2, 2), (24, 1, 2), (24, 0, 1)):
    l, r = H.synth(7, 7 * 16384 + 999, depth, ch)
    want = H.oracle().encode(l, r if ch == 2 else None, 48000, depth, mode)
    got = cd.encode(l, r if ch == 2 else None, 48000, depth, mode)
    assert got == want, (depth, mode, ch, len(got), len(want))
    dl, dr, hdr = cd.decode(want)
    assert np.array_equal(dl, l) and (ch == 1 or np.array_equal(dr, r))
    bad = bytearray(want); bad[len(bad) * 3 // 4] ^= 0x40
    try:
        a, ea = H.oracle().decode(bytes(bad)), None
    except RuntimeError as e:
        a, ea = None, str(e)
    try:
        g, eg = cd.decode(bytes(bad)), None
    except RuntimeError as e:
        g, eg = None, str(e)
    assert ea == eg, (ea, eg)
# incompressible input: the payload outgrows the library's first allocation in the sliced path (realloc),
# and a caller buffer that is too small must be reported with the size it needs
rng = np.random.default_rng(1)
l = rng.integers(-32768, 32768, 12 * 16384 + 5).astype(np.int32)
r = rng.integers(-32768, 32768, l.size).astype(np.int32)
want = H.oracle().encode(l, r, 44100, 16, 0)
assert len(want) > l.size * 4
assert cd.encode(l, r, 44100, 16, 0) == want
pk = np.empty(l.size * 4, dtype=np.uint8)
iv = np.stack([l, r], axis=1).astype("<i2").view(np.uint8).reshape(-1)
small = np.zeros(l.size * 2, dtype=np.uint8)
bb = np.zeros(13, dtype=np.uint32)
try:
    cd.encode_into(iv, small, bb, 16, 2, 0)
    raise SystemExit("small buffer accepted")
except RuntimeError as e:
    need = int(str(e).split("needs ")[1].split(" ")[0])
big = np.zeros(need, dtype=np.uint8)
n = cd.encode_into(iv, big, bb, 16, 2, 0)
assert n == need and int(bb.sum()) == n
cd.set_concurrency(1)  # one stream: the unsliced path, same bytes
n1 = cd.encode_into(iv, big, bb, 16, 2, 0)
assert n1 == n
cd.set_concurrency(0)
print("sliced ok")
"""


def run_sliced(codec: str, slice_blocks: int):
    """The pipelined host paths (slices of whole blocks alternating between two slice contexts)
    must give the bytes / samples / verdicts of a single pass; LACB_SLICE_BLOCKS forces small
    slices so that an 8-block input is cut into several."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, LACB_SLICE_BLOCKS=str(slice_blocks))
    out = subprocess.run([sys.executable, "-c", SLICED_SCRIPT.format(tests=str(H.ROOT / "tests"), codec=codec)],
                         env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "sliced ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("slice_blocks", [1, 3])
def test_sliced_host_pipeline(slice_blocks):
    run_sliced("gpu_codec", slice_blocks)


def test_repeatability(cd):
    """Races in the barrier-light analysis kernel or the hard-chunk queues would show up as run-to-run
    differences: the same input, encoded 12 times with all SMs busy, must give the oracle's bytes every time."""
    l, r = H.synth(4, 48000 * 20, 24)          # every signal section, auto stereo (probes + both paths)
    want = H.oracle().encode(l, r, 48000, 24, 2, threads=8)
    l2, r2 = H.synth(2, 96000 * 20, 24)
    want2 = H.oracle().encode(l2, r2, 96000, 24, 1, threads=8)
    for _ in range(12):
        assert cd.encode(l, r, 48000, 24, 2) == want
        assert cd.encode(l2, r2, 96000, 24, 1) == want2
    dl, dr, _ = cd.decode(want2)
    for _ in range(6):
        a, b, _ = cd.decode(want2)
        assert np.array_equal(a, dl) and np.array_equal(b, dr)
    assert np.array_equal(dl, l2) and np.array_equal(dr, r2)


def restore_overflow_fuzz(cd, rounds, flips):
    """Signals near int32 full scale whose predictors leave small residuals: a flipped residual bit makes the
    reconstruction leave int32, which Block::Decoder rejects (block/decoder.cpp:308-403).  The restore kernel
    keeps only 32-bit arithmetic on its sample-to-sample chain, so its 64-bit verdict is checked here."""
    rng = np.random.default_rng(3)
    rejected = 0
    for it in range(rounds):
        n = int(rng.choice([512, 2048, 4096]))
        t = np.arange(n)
        kind = it % 4
        if kind == 0:
            x = (2**31 - 2000 - 3 * t).astype(np.int64)
        elif kind == 1:
            x = (-(2**31) + 5000 + (t * t) // 50).astype(np.int64)
        elif kind == 2:
            x = np.round((2**31 - 10) * np.sin(2 * np.pi * t / 97.0)).astype(np.int64)
        else:
            x = np.round((2**31 - 10) * np.sin(2 * np.pi * t / 31.0) * np.cos(2 * np.pi * t / 411.0)).astype(np.int64)
        pcm = np.clip(x + rng.integers(-3, 4, n), -(2**31), 2**31 - 1).astype(np.int32)
        good = H.oracle().block_encode(pcm, 1, 1)
        assert cd.block_encode(pcm, 1, 1) == good
        for _ in range(flips):
            b = bytearray(good)
            b[int(rng.integers(3, len(b)))] ^= 1 << int(rng.integers(0, 8))
            ok_a, dec_a, bits_a = H.oracle().block_decode(bytes(b), n)
            ok_b, dec_b, bits_b = cd.block_decode(bytes(b), n)
            assert ok_a == ok_b, (it, kind)
            rejected += not ok_a
            if ok_a:
                assert bits_a == bits_b and np.array_equal(dec_a, dec_b)
    return rejected


def test_restore_overflow_verdicts(cd):
    assert restore_overflow_fuzz(cd, 24, 25) > 300


def lpc_fullscale_check(cd, count):
    """LPC restore at the int32 limits (the FP64 form of the chain must be exact there): verdicts and samples equal to
    the oracle decoder's on hand-built streams (helpers.craft_lpc_block)."""
    accepted = rejected = 0
    peak = 0
    for stream, n in H.lpc_fullscale_streams(count):
        ok_a, dec_a, bits_a = H.oracle().block_decode(stream, n)
        ok_b, dec_b, bits_b = cd.block_decode(stream, n)
        assert ok_a == ok_b
        if ok_a:
            assert bits_a == bits_b and np.array_equal(dec_a, dec_b)
            accepted += 1
            peak = max(peak, int(np.abs(dec_a.astype(np.int64)).max()))
        else:
            rejected += 1
    return accepted, rejected, peak


def test_lpc_restore_at_int32_limits(cd):
    accepted, rejected, peak = lpc_fullscale_check(cd, 300)
    assert accepted > 50 and rejected > 50 and peak > (1 << 30)


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


def test_concurrency_cap_same_bytes(cd):
    """lacb_set_concurrency (the GPU path's --threads): 1 = one stream, nothing overlapped; bytes never change.
    The input is long enough (2300 blocks) for the automatic plan to cut it into slices."""
    l, r, pk = H.synth(2, 2300 * 16384 + 777, 24, want_packed=True)
    ref = None
    try:
        for n in (1, 2, 0):
            cd.set_concurrency(n)
            payload, bb, sizes = cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2)
            (back,) = cd.decode_blocks(payload, sizes, bb, 24, 2, 1, packed=True)
            assert np.array_equal(back, pk)
            cur = (payload.tobytes(), bb.tobytes())
            assert ref is None or cur == ref
            ref = cur
    finally:
        cd.set_concurrency(0)


def test_unpartitioned_file_round_trip(cd):
    """Streams written with partitioning disabled: every adaptive segment is decoded with the stateful model
    (speculative batches for Rice / bin mode, the serial reader for zero-run mode)."""
    for seed, depth in ((2, 24), (1, 16)):
        l, r = H.synth(seed, 16 * 16384 + 99, depth)
        want = H.oracle().encode(l, r, 48000, depth, 1, partitioning=False)
        assert cd.encode(l, r, 48000, depth, 1, partitioning=False) == want
        dl, dr, _ = cd.decode(want)
        assert np.array_equal(dl, l) and np.array_equal(dr, r)
