"""Whole-file decision dump (`lac_cli encode --debug-lpc --debug-zr --debug-partitions --debug-stereo-est`, fed by
lacb_last_encode_decisions) against the log of the reference's own Debug build (oracle/_ref/lac_cli_ref_debug: the
unmodified reference compiled without -DNDEBUG, so its LAC_DEBUG_LOG lines exist).  The lines must be the same
multiset: predictor, order, the four mode estimates, every partition level's total, the chosen level, the stereo
decision of every block.  (`energy=` of the reference's [debug-lpc] line is not kept on the device and is removed
before comparing; its [part-plan] / [part-samples] lines have no counterpart.)"""
import os
import re
import subprocess
from collections import Counter
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from test_cli_gpu import _write_wav

REF_DEBUG = H.ORACLE_DIR / "_ref" / "lac_cli_ref_debug"
KEEP = ("[zr-est]", "[part-est]", "[part-choose]", "[debug-lpc]", "[stereo-est]", "[stereo-mode]")
FLAGS = ["--debug-lpc", "--debug-zr", "--debug-partitions", "--debug-stereo-est"]


def _lines(text, prefixes=KEEP):
    out = []
    for ln in text.splitlines():
        if ln.startswith(prefixes):
            out.append(re.sub(r" energy=\S+", "", ln.strip()))
    return Counter(out)


def _emu_cli(tmp_path_factory):
    """lac_cli linked against the CPU emulator build of the kernels (test infrastructure)."""
    H.emu_codec()  # builds tests/emu/liblac_b200_emu.so when stale
    exe = tmp_path_factory.mktemp("emucli") / "lac_cli_emu"
    host = H.PKG_DIR / "host"
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-I/usr/local/cuda/include", str(host / "lac_cli.cpp"),
                           str(host / "lac_host.cpp"), str(host / "wav_io.cpp"), "-o", str(exe),
                           f"-L{H.ROOT / 'tests' / 'emu'}", "-llac_b200_emu", f"-Wl,-rpath,{H.ROOT / 'tests' / 'emu'}"])
    return str(exe)


@pytest.fixture(scope="module")
def emu_cli(tmp_path_factory):
    if not Path("/usr/local/cuda/include/cuda_runtime.h").exists():
        pytest.skip("CUDA headers needed to compile the host facade")
    return _emu_cli(tmp_path_factory)


def _compare(cli, tmp_path, frames, depth, rate):
    if not REF_DEBUG.exists():
        pytest.skip("oracle/_ref/lac_cli_ref_debug not built (reference tree absent)")
    l, r, pk = H.synth(5, frames, depth, want_packed=True)
    wav = tmp_path / "in.wav"
    _write_wav(wav, pk, 2, rate, depth)
    for mode in ("lr", "ms", None):
        extra = [f"--stereo-mode={mode}"] if mode else []
        ours = subprocess.run([cli, "encode", str(wav), str(tmp_path / "a.lac"), *extra, *FLAGS], capture_output=True, text=True)
        ref = subprocess.run([str(REF_DEBUG), "encode", str(wav), str(tmp_path / "b.lac"), "--threads=1", *extra, *FLAGS],
                             capture_output=True, text=True)
        assert ours.returncode == 0 and ref.returncode == 0, ours.stderr + ref.stderr
        assert (tmp_path / "a.lac").read_bytes() == (tmp_path / "b.lac").read_bytes()
        if mode:  # forced modes: every channel-block the reference encodes is emitted, so every line has a partner
            a, b = _lines(ours.stderr), _lines(ref.stderr)
            assert a == b, f"mode {mode}: only ours {list((a - b).items())[:3]}, only reference {list((b - a).items())[:3]}"
        else:     # auto: the reference also logs its probe / both-pair encodes; the stereo decisions must agree
            a, b = _lines(ours.stderr, ("[stereo-est]", "[stereo-mode]")), _lines(ref.stderr, ("[stereo-est]", "[stereo-mode]"))
            assert a == b and sum(a.values()) > 0
        dz_a = [ln for ln in ours.stdout.splitlines() if ln.startswith("[debug-zr]")]
        dz_b = [ln for ln in ref.stdout.splitlines() if ln.startswith("[debug-zr]")]
        assert dz_a == dz_b and len(dz_a) == 1


def test_decision_dump_matches_reference_debug_log_emulator(emu_cli, tmp_path):
    _compare(emu_cli, tmp_path, 3 * 16384 + 777, 16, 44100)


@pytest.mark.gpu
def test_decision_dump_matches_reference_debug_log(tmp_path):
    cli = H.PKG_DIR / "host" / "lac_cli"
    if not cli.exists():
        pytest.skip("lac_cli not built")
    _compare(str(cli), tmp_path, 40 * 16384 + 4321, 24, 96000)
