"""lac_cli (C++20 host code over the C ABI) on a real GPU: same workflow as the reference
CLI, byte-identical .lac files, bit-exact WAV round trip."""
import os
import struct
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
CLI = H.PKG_DIR / "host" / "lac_cli"


def _write_wav(path, packed: np.ndarray, channels, rate, depth):
    align = channels * depth // 8
    n = packed.size
    hdr = b"RIFF" + struct.pack("<I", 36 + n + (n & 1)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate,
                                                                                      rate * align, align, depth)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", n) + packed.tobytes() + (b"\0" if n & 1 else b""))


@pytest.fixture(scope="module")
def cli():
    if not CLI.exists():
        subprocess.check_call(["make", "-s", "-C", str(H.PKG_DIR), "host"])
    return str(CLI)


def _run(*args):
    return subprocess.run(list(args), capture_output=True, text=True)


def test_selftest(cli):
    r = _run(cli, "selftest")
    assert r.returncode == 0, r.stderr
    assert "Selftest complete" in r.stdout


@pytest.mark.parametrize("depth,rate,channels,flag,mode", [(16, 44100, 2, None, 2), (24, 96000, 2, "--stereo-mode=ms", 1),
                                                           (24, 48000, 2, "--stereo-mode=lr", 0), (24, 192000, 1, None, 0)])
def test_encode_decode_matches_reference(cli, tmp_path, depth, rate, channels, flag, mode):
    frames = 5 * 16384 + 1717
    l, r, pk = H.synth(7, frames, depth, channels=channels, want_packed=True)
    wav, lac, back = tmp_path / "in.wav", tmp_path / "out.lac", tmp_path / "back.wav"
    _write_wav(wav, pk, channels, rate, depth)
    args = [cli, "encode", str(wav), str(lac), "--threads=2"] + ([flag] if flag else [])
    res = _run(*args)
    assert res.returncode == 0, res.stderr
    assert res.stdout.startswith(f"Encoded {wav} -> {lac} (")
    got = lac.read_bytes()
    want = H.oracle().encode(l, r if channels == 2 else None, rate, depth, mode)
    assert got == want
    if H.REF_CLI.exists():  # the unmodified reference CLI on the same file
        ref_lac = tmp_path / "ref.lac"
        rr = _run(str(H.REF_CLI), *args[1:3], str(ref_lac), *args[4:])
        assert rr.returncode == 0 and ref_lac.read_bytes() == got
    res = _run(cli, "decode", str(lac), str(back))
    assert res.returncode == 0, res.stderr
    assert f"({frames} samples per channel)" in res.stdout
    assert back.read_bytes() == wav.read_bytes()


def test_cli_rejections(cli, tmp_path):
    wav = tmp_path / "a.wav"
    l, r, pk = H.synth(1, 3000, 16, want_packed=True)
    _write_wav(wav, pk, 2, 44100, 16)
    assert _run(cli, "encode", str(wav), str(wav)).returncode == 1                      # same path
    assert _run(cli, "encode", str(wav), str(tmp_path / "x.lac"), "--threads=0").returncode == 1
    bad = tmp_path / "bad.wav"
    bad.write_bytes(wav.read_bytes()[:-5])
    out = tmp_path / "bad.lac"
    assert _run(cli, "encode", str(bad), str(out)).returncode == 1 and not out.exists()
    lac = tmp_path / "ok.lac"
    assert _run(cli, "encode", str(wav), str(lac)).returncode == 0
    broken = bytearray(lac.read_bytes())
    broken[30] ^= 0xFF
    (tmp_path / "broken.lac").write_bytes(bytes(broken[:-3]))
    res = _run(cli, "decode", str(tmp_path / "broken.lac"), str(tmp_path / "o.wav"))
    assert res.returncode == 1 and "Decode failed: [decode-error]" in res.stderr and not (tmp_path / "o.wav").exists()
    assert not [p for p in os.listdir(tmp_path) if ".tmp." in p]                       # no staged leftovers


def test_batch_mode(cli, tmp_path):
    """`lac_cli batch list.txt` runs many commands in one process (CUDA start-up paid once); results
    are those of the one-shot commands, a failing line gives exit status 1 but does not stop the rest."""
    lines = []
    want = {}
    for i, (depth, rate) in enumerate(((16, 44100), (24, 48000), (24, 96000))):
        l, r, pk = H.synth(20 + i, 2 * 16384 + 311 * i, depth, want_packed=True)
        wav = tmp_path / f"in{i}.wav"
        _write_wav(wav, pk, 2, rate, depth)
        lines.append(f"encode {wav} {tmp_path / f'o{i}.lac'}   # file {i}")
        lines.append(f"decode {tmp_path / f'o{i}.lac'} {tmp_path / f'b{i}.wav'}")
        want[i] = (H.oracle().encode(l, r, rate, depth, 2), wav.read_bytes())
    lst = tmp_path / "list.txt"
    lst.write_text("\n".join(lines) + "\n\n")
    res = _run(cli, "batch", str(lst))
    assert res.returncode == 0, res.stderr
    for i, (lac, wav) in want.items():
        assert (tmp_path / f"o{i}.lac").read_bytes() == lac
        assert (tmp_path / f"b{i}.wav").read_bytes() == wav
    lst.write_text(f"decode {tmp_path / 'missing.lac'} {tmp_path / 'x.wav'}\n" + lines[0] + "\n")
    (tmp_path / "o0.lac").unlink()
    res = _run(cli, "batch", str(lst))
    assert res.returncode == 1 and (tmp_path / "o0.lac").exists()


def test_threads_flag_is_honoured(cli, tmp_path):
    """--threads caps the workers (src/main.cpp:560-591): 1 = one device, one stream; the bytes do not depend on it."""
    l, r, pk = H.synth(3, 6 * 16384 + 5, 24, want_packed=True)
    wav = tmp_path / "in.wav"
    _write_wav(wav, pk, 2, 96000, 24)
    outs = []
    for t in ("--threads=1", "--threads=7"):
        lac = tmp_path / f"o{t[-1]}.lac"
        res = _run(cli, "encode", str(wav), str(lac), t, "--debug-threads")
        assert res.returncode == 0, res.stderr
        assert "Thread usage: 1 threads" in res.stdout  # one device => one worker, whatever the cap
        outs.append(lac.read_bytes())
        res = _run(cli, "decode", str(lac), str(tmp_path / "b.wav"), t, "--debug-threads")
        assert res.returncode == 0 and "Decoder thread usage: 1 threads" in res.stdout
        assert (tmp_path / "b.wav").read_bytes() == wav.read_bytes()
    assert outs[0] == outs[1] == H.oracle().encode(l, r, 96000, 24, 2)


def _rf64(packed: np.ndarray, channels, rate, depth) -> bytes:
    """EBU Tech 3306 RF64 form of a (small) WAV: ds64 chunk with the 64-bit sizes, 0xFFFFFFFF in the 32-bit fields"""
    align = channels * depth // 8
    n = packed.size
    ds64 = struct.pack("<QQQI", 72 + n + (n & 1), n, n // align, 0)
    return (b"RF64" + struct.pack("<I", 0xFFFFFFFF) + b"WAVEds64" + struct.pack("<I", 28) + ds64 + b"fmt " +
            struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * align, align, depth) + b"data" +
            struct.pack("<I", 0xFFFFFFFF) + packed.tobytes() + (b"\0" if n & 1 else b""))


def test_rf64_input_needs_allow_large(cli, tmp_path):
    l, r, pk = H.synth(5, 3 * 16384 + 1, 24, want_packed=True)
    wav, lac = tmp_path / "in.rf64.wav", tmp_path / "o.lac"
    wav.write_bytes(_rf64(pk, 2, 48000, 24))
    assert _run(cli, "encode", str(wav), str(lac)).returncode == 1 and not lac.exists()
    res = _run(cli, "encode", str(wav), str(lac), "--allow-large")
    assert res.returncode == 0, res.stderr
    assert lac.read_bytes() == H.oracle().encode(l, r, 48000, 24, 2)


def test_decode_size_cap_and_opt_in(cli, tmp_path):
    """A table that declares more than 1 GiB of int32 PCM is rejected before anything is allocated, exactly like the
    reference CLI (src/main.cpp:247-251); --allow-large is the explicit opt-in."""
    nb = 65536 + 512  # x 16384 samples x 4 bytes > 1 GiB, mono
    table = np.empty((nb, 2), dtype=">u4")
    table[:, 0], table[:, 1] = 16384, 1
    crafted = H.lacb_module().FrameHeader(1, 0, 48000, 16).pack() + struct.pack(">I", nb) + table.tobytes() + b"\0" * nb
    bad = tmp_path / "crafted.lac"
    bad.write_bytes(crafted)
    res = _run(cli, "decode", str(bad), str(tmp_path / "o.wav"))
    assert res.returncode == 1 and "decoded PCM allocation exceeds maximum" in res.stderr
    assert not (tmp_path / "o.wav").exists()
    res = _run(cli, "decode", str(bad), str(tmp_path / "o.wav"), "--allow-large")
    assert res.returncode == 1 and "[decode-error] block=0 channel=primary" in res.stderr  # now it gets as far as the device


def test_resident_server_forwards_commands(cli, tmp_path):
    """`lac_cli serve` keeps the CUDA context alive; clients with LAC_SERVER set forward encode / decode to it and get
    the same files, output text and exit status as a local run; without a listener they run locally."""
    import time
    sock = str(tmp_path / "s.sock")
    srv = subprocess.Popen([cli, "serve", sock], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        assert "serving on" in srv.stdout.readline()
        env = dict(os.environ, LAC_SERVER=sock)
        l, r, pk = H.synth(11, 4 * 16384 + 7, 24, want_packed=True)
        wav, lac, back = tmp_path / "in.wav", tmp_path / "o.lac", tmp_path / "b.wav"
        _write_wav(wav, pk, 2, 48000, 24)
        t0 = time.perf_counter()
        e = subprocess.run([cli, "encode", "in.wav", "o.lac"], capture_output=True, text=True, env=env, cwd=tmp_path)
        d = subprocess.run([cli, "decode", str(lac), str(back)], capture_output=True, text=True, env=env)
        dt = time.perf_counter() - t0
        assert e.returncode == 0 and d.returncode == 0, e.stderr + d.stderr
        assert e.stdout.startswith("Encoded ") and "samples per channel" in d.stdout
        assert lac.read_bytes() == H.oracle().encode(l, r, 48000, 24, 2) and back.read_bytes() == wav.read_bytes()
        assert dt < 2.0  # two forwarded commands: no CUDA start-up in either
        bad = subprocess.run([cli, "decode", str(wav), str(tmp_path / "x.wav")], capture_output=True, text=True, env=env)
        assert bad.returncode == 1 and "Decode failed" in bad.stderr
        assert subprocess.run([cli, "shutdown"], env=env, capture_output=True).returncode == 0
        assert srv.wait(timeout=30) == 0
        # nobody listens any more: the same command line runs locally
        e2 = subprocess.run([cli, "encode", str(wav), str(tmp_path / "o2.lac")], capture_output=True, text=True, env=env)
        assert e2.returncode == 0 and (tmp_path / "o2.lac").read_bytes() == lac.read_bytes()
    finally:
        if srv.poll() is None:
            srv.kill()
