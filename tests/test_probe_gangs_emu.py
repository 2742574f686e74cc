"""Stereo probes as gangs (k_analyze<32, 8, true>): the bytes must not depend on how many probe
warps share a CTA, including gang sizes that leave the last row of jobs partly filled (completed with copies of the
last job).  The gang size is read once per process (LACB_PROBE_GANG), so every size runs in its own interpreter, on the
CPU emulator build of the same .cu sources."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

SNIPPET = r"""
import sys
sys.path.insert(0, r"%s")
import numpy as np
import helpers as H
cd = H.emu_codec()
l, r = H.synth(1, 5 * 16384 + 1234, 16)          # six blocks, several of them decided by probes
got = cd.encode(l, r, 44100, 16, 2)
want = H.oracle().encode(l, r, 44100, 16, 2)
assert got == want, (len(got), len(want))
dl, dr, hdr = cd.decode(got)
assert np.array_equal(dl, l) and np.array_equal(dr, r)
print("ok", len(got))
""" % str(ROOT / "tests")


@pytest.mark.parametrize("gang", ["1", "5", "7", "32"])
def test_probe_gang_size_does_not_change_the_bytes(gang):
    env = dict(os.environ, LACB_PROBE_GANG=gang)
    out = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok")
