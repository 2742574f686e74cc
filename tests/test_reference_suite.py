"""The reference's OWN test executables (lac_tests, lac_rice_tests, lac_cli_tests), compiled
unchanged from /root/reference/tests/*.cpp against this repo's C++ facade
(tests/reftests/Makefile) instead of the reference's liblac:

  * emulator-linked build: runs here on the CPU (kernels through tests/emu), not gpu-marked;
  * liblac_b200.so-linked build: runs on the B200 box (the binaries travel with the snapshot).

Skipped when the binaries are absent and the reference tree is not there to build them."""
import subprocess

import pytest

import helpers as H

DIR = H.ROOT / "tests" / "reftests"
REF = H.Path("/root/reference")


def _ensure(kind):
    exe = DIR / "_build" / kind / "lac_tests"
    if not exe.exists() and (REF / "tests").is_dir():
        if kind == "emu":
            H.emu_codec()  # builds tests/emu/liblac_b200_emu.so
        subprocess.check_call(["make", "-s", "-C", str(DIR), kind])
    if not exe.exists():
        pytest.skip(f"tests/reftests/_build/{kind} not built and {REF} is absent")
    return exe.parent


def _run(exe, *args, timeout=900):
    r = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return r.stdout


def test_reference_suite_on_emulator():
    d = _ensure("emu")
    assert "rice tests ok" in _run(d / "lac_rice_tests")
    out = _run(d / "lac_tests")
    for marker in ("predictor selection tests ok", "decoder error tests ok", "stereo planner tests ok",
                   "block planner tests ok", "canonical block metadata tests ok", "e2e wav->lac->wav tests ok"):
        assert marker in out


@pytest.mark.gpu
def test_reference_suite_on_gpu():
    d = _ensure("gpu")
    assert "rice tests ok" in _run(d / "lac_rice_tests")
    out = _run(d / "lac_tests")
    for marker in ("predictor selection tests ok", "decoder error tests ok", "stereo planner tests ok",
                   "block planner tests ok", "canonical block metadata tests ok", "e2e wav->lac->wav tests ok"):
        assert marker in out


@pytest.mark.gpu
def test_reference_cli_suite_on_gpu():
    """tests/test_cli.cpp drives a lac_cli binary as a subprocess: give it ours."""
    d = _ensure("gpu")
    cli = H.PKG_DIR / "host" / "lac_cli"
    if not (d / "lac_cli_tests").exists() or not cli.exists():
        pytest.skip("lac_cli_tests / lac_cli not built")
    _run(d / "lac_cli_tests", str(cli))
