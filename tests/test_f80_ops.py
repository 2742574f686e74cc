"""Operation-level check of the software x87 arithmetic behind the Levinson kernel (csrc/lacb_f80.cuh):
add, sub, mul, div, compare, int64 conversion and the Q15 quantisation against the host's native 80-bit
`long double` on >= 20 M random operands (integers up to 63 bits incl. R[0] > 2^53, quotients, products).
The harness (tests/emu/f80_check.cpp) compiles the same header through the CPU emulator."""
import platform
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(platform.machine() not in ("x86_64", "AMD64"), reason="needs the x87 long double of x86-64")
def test_f80_operations_match_native_long_double(tmp_path):
    exe = tmp_path / "f80_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-DLACB_EMU=1", f"-I{ROOT / 'tests' / 'emu'}", f"-I{ROOT / 'include'}",
                           str(ROOT / "tests" / "emu" / "f80_check.cpp"), str(ROOT / "tests" / "emu" / "cuda_emu.cpp"),
                           "-o", str(exe), "-Wno-unknown-pragmas"])
    res = subprocess.run([str(exe), "20000000"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 mismatches" in res.stdout
