"""Block-range sharding across GPUs (runs only where >= 2 devices are visible)."""
import subprocess

import numpy as np
import pytest

import helpers as H
from test_cli_gpu import CLI, _write_wav

pytestmark = pytest.mark.gpu


def _ndev():
    return H.lacb_module().load_library().lacb_device_count()


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
def test_cli_two_devices_same_bytes(tmp_path):
    frames = 11 * 16384 + 99
    l, r, pk = H.synth(4, frames, 24, want_packed=True)
    wav = tmp_path / "in.wav"
    _write_wav(wav, pk, 2, 48000, 24)
    outs = []
    for dev in (1, 2):
        lac = tmp_path / f"out{dev}.lac"
        res = subprocess.run([str(CLI), "encode", str(wav), str(lac), f"--devices={dev}", "--debug-threads"],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        outs.append(lac.read_bytes())
        if dev == 2:
            assert "Thread usage: 2 threads" in res.stdout
    assert outs[0] == outs[1] == H.oracle().encode(l, r, 48000, 24, 2)


@pytest.mark.skipif(_ndev() < 2, reason="needs two GPUs")
def test_second_device_context():
    cd = H.lacb_module().Codec(1, H.GPU_SO)
    l, r = H.synth(9, 40000, 16)
    assert cd.encode(l, r, 44100, 16, 2) == H.oracle().encode(l, r, 44100, 16, 2)
