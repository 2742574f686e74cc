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
