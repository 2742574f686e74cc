"""The C-ABI library loads on a CPU-only box and exports every symbol include/lac_b200.h
declares; with no GPU it refuses to create a context (no CPU fallback)."""
import ctypes as C
import re
import subprocess

import helpers as H


def _declared():
    text = (H.ROOT / "include" / "lac_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lacb_[a-z0-9_]+)\s*\(", text)))


def _ensure_lib():
    if not H.GPU_SO.exists():
        subprocess.check_call(["make", "-s", "-C", str(H.PKG_DIR), "lib"])
    return H.GPU_SO


def test_header_symbols_exported():
    so = _ensure_lib()
    names = _declared()
    assert len(names) >= 20
    lib = C.CDLL(str(so))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(H.lacb_module().EXPORTS) <= set(names)


def test_no_cpu_fallback_without_gpu():
    so = _ensure_lib()
    lib = C.CDLL(str(so))
    if lib.lacb_device_count() > 0:
        return  # on a GPU box the gpu-marked tests exercise the library
    h = C.c_void_p()
    assert lib.lacb_create(0, C.byref(h)) != 0 and not h.value
    try:
        H.lacb_module().Codec(0, so)
    except RuntimeError as e:
        assert "no usable CUDA device" in str(e)
    else:
        raise AssertionError("Codec() must fail loudly without a CUDA device")


def test_product_does_not_reference_oracle_or_emulator():
    """The shipped sources never include/link the checker or the emulator."""
    for p in list((H.PKG_DIR / "csrc").iterdir()) + [H.PKG_DIR / "lacb.py", H.PKG_DIR / "sharding.py"]:
        t = p.read_text()
        assert "lac_oracle" not in t and "liblac_ref" not in t, p
        assert "cuda_emu.h" not in t, p
