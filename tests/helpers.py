"""Test-side bindings: the C oracle (oracle/liblac_oracle.so), the compiled
unmodified reference (oracle/_ref/liblac_ref.so, optional) and the synthetic
signal generator (tools/liblac_synth.so), plus the deterministic parity corpus.

Nothing here is product code: the product lives in lossless-audio-codec_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "liblac_oracle.so"
REF_SO = ORACLE_DIR / "_ref" / "liblac_ref.so"
REF_CLI = ORACLE_DIR / "_ref" / "lac_cli_ref"
SYNTH_SO = ROOT / "tools" / "liblac_synth.so"

i32p = C.POINTER(C.c_int32)
u8p = C.POINTER(C.c_uint8)


def _build_if_missing():
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (ORACLE_DIR / "lac_oracle.c").stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(ORACLE_DIR), "oracle"])
    if not SYNTH_SO.exists() or SYNTH_SO.stat().st_mtime < (ROOT / "tools" / "lac_synth.c").stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", str(SYNTH_SO),
                               str(ROOT / "tools" / "lac_synth.c")])


class BlockInfo(C.Structure):
    _fields_ = [
        ("predictor_type", C.c_uint32),
        ("order", C.c_uint32),
        ("coeffs", C.c_int16 * 33),
        ("partition_order", C.c_uint32),
        ("n_parts", C.c_uint32),
        ("part_mode", C.c_uint8 * 256),
        ("part_k", C.c_uint8 * 256),
        ("est_total_bits", C.c_uint64),
        ("cand_best_bits", C.c_uint64 * 11),
    ]


def _as_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(i32p)


def _take(lib_free, ptr, n, dtype=np.uint8):
    if not ptr:
        return np.zeros(0, dtype=dtype)
    arr = np.ctypeslib.as_array(ptr, shape=(int(n),)).copy() if n else np.zeros(0, dtype=dtype)
    lib_free(ptr)
    return arr.astype(dtype, copy=False)


class _Codec:
    """Common shape of the oracle ('lao_') and reference ('ref_') C APIs."""

    def __init__(self, path: Path, prefix: str):
        self.lib = C.CDLL(str(path))
        self.p = prefix
        L = self.lib
        g = lambda name: getattr(L, prefix + name)
        g("free").argtypes = [C.c_void_p]
        g("free").restype = None
        g("last_error").restype = C.c_char_p
        g("encode").argtypes = [i32p, i32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                C.c_int, C.c_int, C.c_uint32, C.POINTER(u8p), C.POINTER(C.c_uint64)]
        g("decode").argtypes = [u8p, C.c_uint64, C.c_uint32, C.POINTER(i32p), C.POINTER(i32p),
                                C.POINTER(C.c_uint64)] + [C.POINTER(C.c_uint32)] * 4
        if prefix == "lao_":
            g("block_encode").argtypes = [i32p, C.c_uint32, C.c_int, C.c_int, C.POINTER(u8p),
                                          C.POINTER(C.c_uint64), C.POINTER(BlockInfo)]
        else:
            g("block_encode").argtypes = [i32p, C.c_uint32, C.c_int, C.c_int, C.POINTER(u8p),
                                          C.POINTER(C.c_uint64)]
        g("block_decode").argtypes = [u8p, C.c_uint64, C.c_uint32, i32p, C.POINTER(C.c_uint64)]
        g("lpc_analyze").argtypes = [i32p, C.c_uint32, C.c_int, C.POINTER(C.c_int16)]
        self._g = g

    def free(self, ptr):
        self._g("free")(C.cast(ptr, C.c_void_p))

    def last_error(self) -> str:
        return (self._g("last_error")() or b"").decode()

    def encode(self, left, right=None, sample_rate=44100, bit_depth=16, stereo_mode=0,
               zero_run=True, partitioning=True, threads=1) -> bytes:
        la, lp = _as_i32(left)
        rp = None
        if right is not None and len(right):
            ra, rp = _as_i32(right)
        out = u8p()
        n = C.c_uint64()
        rc = self._g("encode")(lp, rp, len(la), sample_rate, bit_depth, stereo_mode,
                               int(zero_run), int(partitioning), threads, C.byref(out), C.byref(n))
        if rc != 0:
            raise ValueError(f"encode rejected rc={rc}: {self.last_error()}")
        return _take(self.free, out, n.value).tobytes()

    def decode(self, data: bytes, threads=1):
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data) if data else (C.c_uint8 * 1)()
        l, r = i32p(), i32p()
        frames = C.c_uint64()
        ch, sr, bd, sm = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        rc = self._g("decode")(C.cast(buf, u8p), len(data), threads, C.byref(l), C.byref(r),
                               C.byref(frames), C.byref(ch), C.byref(sr), C.byref(bd), C.byref(sm))
        if rc != 0:
            raise RuntimeError(self.last_error())
        left = _take(self.free, l, frames.value, np.int32)
        right = _take(self.free, r, frames.value, np.int32) if ch.value == 2 else np.zeros(0, np.int32)
        return left, right, dict(channels=ch.value, sample_rate=sr.value, bit_depth=bd.value,
                                 stereo_mode=sm.value)

    def block_encode(self, pcm, zero_run=True, partitioning=True, want_info=False):
        a, p = _as_i32(pcm)
        out = u8p()
        n = C.c_uint64()
        if self.p == "lao_":
            info = BlockInfo()
            rc = self._g("block_encode")(p, len(a), int(zero_run), int(partitioning), C.byref(out),
                                         C.byref(n), C.byref(info))
        else:
            info = None
            rc = self._g("block_encode")(p, len(a), int(zero_run), int(partitioning), C.byref(out),
                                         C.byref(n))
        assert rc == 0
        data = _take(self.free, out, n.value).tobytes()
        return (data, info) if want_info else data

    def block_decode(self, data: bytes, block_size: int):
        """Returns (ok, pcm, bits_consumed)."""
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data + b"\0" * (1 if not data else 0))
        out = np.zeros(max(1, block_size), dtype=np.int32)
        bits = C.c_uint64()
        ok = self._g("block_decode")(C.cast(buf, u8p), len(data), block_size,
                                     out.ctypes.data_as(i32p), C.byref(bits))
        return bool(ok), out[:block_size], bits.value

    def lpc_analyze(self, pcm, order: int):
        a, p = _as_i32(pcm)
        c = (C.c_int16 * (order + 1))()
        used = self._g("lpc_analyze")(p, len(a), order, c)
        return used, np.array(c[:], dtype=np.int16)


_oracle = None
_ref = None
_synth = None


def oracle() -> _Codec:
    global _oracle
    if _oracle is None:
        _build_if_missing()
        _oracle = _Codec(ORACLE_SO, "lao_")
        _oracle.lib.lao_stereo_proxy.argtypes = [i32p, i32p, C.c_uint32]
        _oracle.lib.lao_stereo_proxy.restype = C.c_uint32
        _oracle.lib.lao_adaptive_k_series.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32,
                                                      C.c_int, C.POINTER(C.c_uint32)]
    return _oracle


def have_ref() -> bool:
    return REF_SO.exists()


def ref() -> _Codec:
    global _ref
    if _ref is None:
        _ref = _Codec(REF_SO, "ref_")
    return _ref


def synth(seed: int, frames: int, depth: int, channels: int = 2, want_packed=False):
    """SURVEY Appendix-C signal. Returns (left, right[, packed_bytes])."""
    global _synth
    if _synth is None:
        _build_if_missing()
        _synth = C.CDLL(str(SYNTH_SO))
        _synth.lac_synth.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, i32p, i32p, u8p]
        _synth.lac_synth.restype = None
    left = np.zeros(frames, dtype=np.int32)
    right = np.zeros(frames, dtype=np.int32)
    packed = np.zeros(frames * channels * (depth // 8), dtype=np.uint8) if want_packed else None
    _synth.lac_synth(seed, frames, depth, channels, left.ctypes.data_as(i32p),
                     right.ctypes.data_as(i32p),
                     packed.ctypes.data_as(u8p) if want_packed else None)
    if channels == 1:
        right = np.zeros(0, dtype=np.int32)
    return (left, right, packed) if want_packed else (left, right)


C4_RESET_LOG2 = 19  # config 4's range-addressable stream: filter state reset every 2^19 frames (tools/lac_synth.c)


def synth_lib():
    global _synth
    if _synth is None:
        _build_if_missing()
        _synth = C.CDLL(str(SYNTH_SO))
        _synth.lac_synth.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, i32p, i32p, u8p]
        _synth.lac_synth.restype = None
    if not hasattr(_synth, "_range_ready"):
        _synth.lac_synth_range.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                           i32p, i32p, u8p]
        _synth.lac_synth_range.restype = None
        _synth._range_ready = True
    return _synth


def synth_range(seed: int, f0: int, frames: int, depth: int, channels: int = 2, reset_log2: int = C4_RESET_LOG2,
                planes=True, want_packed=False):
    """Frames [f0, f0+frames) of the range-addressable stream (config 4).  Returns (left, right[, packed]);
    planes=False skips the int32 planes (returns (None, None, packed))."""
    lib = synth_lib()
    left = np.zeros(frames, dtype=np.int32) if planes else None
    right = np.zeros(frames, dtype=np.int32) if planes else None
    packed = np.zeros(frames * channels * (depth // 8), dtype=np.uint8) if want_packed else None
    lib.lac_synth_range(seed, f0, frames, depth, channels, reset_log2,
                        left.ctypes.data_as(i32p) if planes else None,
                        right.ctypes.data_as(i32p) if planes else None,
                        packed.ctypes.data_as(u8p) if want_packed else None)
    if channels == 1 and planes:
        right = np.zeros(0, dtype=np.int32)
    return (left, right, packed) if want_packed else (left, right)


# ---------------------------------------------------------------------------
# deterministic parity corpus (shapes follow the reference's own test inputs;
# see SURVEY.md section 8(c) for the file:line of each family)
def lcg_stream(n, seed=1):
    out = np.empty(n, dtype=np.uint32)
    s = seed & 0xFFFFFFFF
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = s
    return out


def block_corpus():
    """name -> int32 array; sizes span 1..16384 and every residual mode."""
    rng = np.random.default_rng(1234)
    c = {}
    c["ramp_mod257_1024"] = (np.arange(1024) % 257).astype(np.int32)            # test_e2e.cpp:472
    st = lcg_stream(4096, 1)
    c["rand_amp1000_4096"] = ((st >> 5).astype(np.int64) % 1000).astype(np.int32)  # test_zerorun.cpp:19
    c["rand_amp_2p23_2048"] = ((lcg_stream(2048, 7) >> 5).astype(np.int64) % (1 << 23)).astype(np.int32)
    sp = np.zeros(4096, dtype=np.int32)
    sp[::11] = 1
    sp[::53] = -1
    c["sparse_4096"] = sp                                                         # test_zerorun.cpp:30
    bf = np.zeros(64, dtype=np.int32)
    for i in range(64):
        if i % 8 == 0:
            bf[i] = 1 << 23
        elif i % 3 == 0:
            bf[i] = 1
        elif i % 5 == 0:
            bf[i] = -1
    c["bin_fallback_64"] = bf                                                     # test_zerorun.cpp:39
    c["zeros_16384"] = np.zeros(16384, dtype=np.int32)
    c["zeros_17"] = np.zeros(17, dtype=np.int32)
    z = np.zeros(2048, dtype=np.int32)
    z[100:140] = rng.integers(-50, 50, 40)
    z[1000] = 1 << 23
    z[1001] = -(1 << 23)
    c["mixed_runs_spikes_2048"] = z
    t = np.arange(16384)
    c["sine_16384"] = np.round(12000 * np.sin(2 * np.pi * 440 * t / 44100)).astype(np.int32)
    c["sine24_16384"] = np.round(2.7e6 * np.sin(2 * np.pi * 443 * t / 96000)).astype(np.int32)
    c["ramp_16384"] = (t * 3 - 20000).astype(np.int32)
    c["noise16_16384"] = rng.integers(-32768, 32768, 16384).astype(np.int32)
    c["noise24_16384"] = rng.integers(-(1 << 23), 1 << 23, 16384).astype(np.int32)
    c["lownoise_16384"] = rng.integers(-3, 4, 16384).astype(np.int32)
    c["pm2_4096"] = rng.integers(-2, 3, 4096).astype(np.int32)
    ar = np.zeros(16384, dtype=np.int64)
    w = rng.integers(-4096, 4096, 16384)
    for i in range(16384):
        ar[i] = ((29491 * ar[i - 1] - 19661 * ar[i - 2] + 9830 * ar[i - 3] - 6554 * ar[i - 4]) >> 15) + w[i]
    c["ar4_16384"] = np.clip(ar, -(1 << 23), (1 << 23) - 1).astype(np.int32)
    steps = np.concatenate([rng.integers(-(1 << (3 + 2 * j)), 1 << (3 + 2 * j), 2048) for j in range(8)])
    c["level_steps_16384"] = steps.astype(np.int32)
    walk = np.cumsum(rng.integers(-300, 301, 12288))
    c["random_walk_12288"] = np.clip(walk, -32768, 32767).astype(np.int32)
    sil = np.zeros(16384, dtype=np.int32)
    sil[8192:] = rng.integers(-20000, 20000, 8192)
    c["silence_then_noise_16384"] = sil                                           # test_e2e.cpp:636
    for n in (1, 2, 3, 5, 13, 31, 32, 33, 63, 64, 65, 127, 128, 255, 256, 257, 400, 576, 1000, 4095, 8176):
        c[f"noise_n{n}"] = rng.integers(-2000, 2000, n).astype(np.int32)
    for n in (64, 96, 160, 256, 320, 576):                                        # test_zerorun.cpp:500-579
        zr = rng.integers(-3, 4, n).astype(np.int32)
        zr[n // 4: n // 4 + n // 3] = 0
        c[f"zr_sweep_n{n}"] = zr
    burst = np.zeros(16384, dtype=np.int32)
    idx = rng.integers(0, 16384, 300)
    burst[idx] = rng.integers(-2, 3, 300)
    c["sparse_bursts_16384"] = burst
    alt = np.where(np.arange(4096) % 2 == 0, 9000, -9000).astype(np.int32)      # test_e2e.cpp:735
    c["alternating_4096"] = alt
    sat = np.where(rng.random(2048) < 0.5, (1 << 23) - 1, -(1 << 23)).astype(np.int32)
    c["fullscale_toggle_2048"] = sat
    big = rng.integers(-(1 << 31), (1 << 31) - 1, 512).astype(np.int32)
    c["int32_noise_512"] = big
    # strongly predictable material near the int32 limits: LPC wins with samples of ~2^29 / 2^30, so the decoder's
    # LPC restore chain (FP64 form: sums of twelve 2^15 x 2^31 products) is exercised where exactness is tightest
    for name, shift in (("ar4_2p29_4096", 16), ("ar4_2p30_2048", 17)):
        nn = int(name.split("_")[-1])
        big_ar = c["ar4_16384"][:nn].astype(np.int64) << shift
        big_ar += rng.integers(-(1 << 20), 1 << 20, nn)
        c[name] = np.clip(big_ar, -(1 << 31), (1 << 31) - 1).astype(np.int32)
    return c


def stereo_corpus():
    """name -> (left, right, depth)."""
    rng = np.random.default_rng(99)
    c = {}
    n = 16384 + 37
    l = rng.integers(-20000, 20000, n).astype(np.int32)
    c["identical_multiblock"] = (l, l.copy(), 16)
    c["left_only"] = (l, np.zeros(n, np.int32), 16)
    c["anticorrelated"] = (l, (-l).clip(-32768, 32767).astype(np.int32), 16)
    walk = np.clip(np.cumsum(rng.integers(-200, 201, 3 * 16384)), -30000, 30000).astype(np.int32)
    c["walk_plus_noise"] = (walk, (walk + rng.integers(-40, 41, walk.size)).clip(-32768, 32767).astype(np.int32), 16)
    c["independent_noise"] = (rng.integers(-9000, 9000, 20000).astype(np.int32),
                              rng.integers(-9000, 9000, 20000).astype(np.int32), 16)
    t = np.arange(2 * 16384 + 5000)
    c["sines24"] = (np.round(2.7e6 * np.sin(2 * np.pi * 440 * t / 48000)).astype(np.int32),
                    np.round(2.5e6 * np.sin(2 * np.pi * 443 * t / 48000)).astype(np.int32), 24)
    c["short_1024"] = (rng.integers(-500, 500, 1024).astype(np.int32),
                       rng.integers(-500, 500, 1024).astype(np.int32), 16)
    c["short_17_24bit"] = (rng.integers(-(1 << 23), 1 << 23, 17).astype(np.int32),
                           rng.integers(-(1 << 23), 1 << 23, 17).astype(np.int32), 24)
    sl, sr = synth(1, 5 * 16384 + 1000, 16)
    c["synth16"] = (sl, sr, 16)
    # one block from each synthetic section at 24 bit
    sl, sr = synth(2, (3 << 17) + 2 * 16384, 24)
    pick = np.concatenate([np.arange(s << 17, (s << 17) + 16384) for s in range(4)] +
                          [np.arange((3 << 17) + 16384, (3 << 17) + 16384 + 5000)])
    c["synth24_sections"] = (sl[pick].copy(), sr[pick].copy(), 24)
    return c


# ---------------------------------------------------------------------------
# product bindings (lossless-audio-codec_b200/lacb.py; the directory name is not an
# importable identifier, so it is loaded by path)
import importlib.util as _ilu

PKG_DIR = ROOT / "lossless-audio-codec_b200"
GPU_SO = PKG_DIR / "liblac_b200.so"
EMU_SO = ROOT / "tests" / "emu" / "liblac_b200_emu.so"
_lacb_mod = None
_gpu_codec = None
_emu_codec = None


def lacb_module():
    global _lacb_mod
    if _lacb_mod is None:
        spec = _ilu.spec_from_file_location("lacb", str(PKG_DIR / "lacb.py"))
        _lacb_mod = _ilu.module_from_spec(spec)
        spec.loader.exec_module(_lacb_mod)
    return _lacb_mod


def gpu_codec():
    """The product: nvcc-built liblac_b200.so on cuda:0.  Raises if it cannot run."""
    global _gpu_codec
    if _gpu_codec is None:
        _gpu_codec = lacb_module().Codec(0, GPU_SO)
    return _gpu_codec


def emu_codec():
    """TEST INFRASTRUCTURE: the same .cu sources compiled against tests/emu/cuda_emu.h
    (CPU fiber emulator) so kernel logic can be checked where no GPU exists."""
    global _emu_codec
    if _emu_codec is None:
        src = [PKG_DIR / "csrc" / f for f in os.listdir(PKG_DIR / "csrc")] + [ROOT / "tests" / "emu" / "cuda_emu.h",
                                                                             ROOT / "tests" / "emu" / "cuda_emu.cpp"]
        if not EMU_SO.exists() or EMU_SO.stat().st_mtime < max(p.stat().st_mtime for p in src):
            subprocess.check_call(["make", "-s", "-C", str(PKG_DIR), "emu"])
        _emu_codec = lacb_module().Codec(0, EMU_SO)
    return _emu_codec


def craft_lpc_block(res, coefs, k=28) -> bytes:
    """A hand-built block stream: LPC predictor with the given Q15 coefficients, unpartitioned, static Rice (mode 3)
    with parameter k, residuals `res` (block/encoder.cpp:773-822 layout: type, order, coefficients, control byte,
    (mode:2, k:5), tokens, zero padding).  The reference's encoder never produces LPC blocks with samples near the int32
    limits (its autocorrelation wraps and the candidate is dropped), but its decoder accepts them."""
    bits = []

    def put(v, n):
        for i in range(n - 1, -1, -1):
            bits.append((v >> i) & 1)

    put(2, 8)
    put(len(coefs), 8)
    for c in coefs:
        put(int(c) & 0xFFFF, 16)
    put((3 & 3) << 5, 8)   # control byte: base mode 3 (static Rice), not partitioned
    put((3 << 5) | k, 7)   # the one partition: mode 3, k
    for r in res:
        r = int(r)
        u = (r << 1) if r >= 0 else (((-r - 1) << 1) | 1)
        q = u >> k
        if q:
            put((1 << q) - 1, q)
        put(0, 1)
        put(u & ((1 << k) - 1), k)
    while len(bits) % 8:
        bits.append(0)
    return bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))


def lpc_fullscale_streams(count, seed=21):
    """(stream, n) pairs: crafted LPC blocks of order 4..12 whose reconstruction runs up to (and sometimes beyond) the
    int32 limits -- impulses of 2^27 .. 2^30 into slowly decaying predictors."""
    rng = np.random.default_rng(seed)
    out = []
    for it in range(count):
        order = int(rng.choice([4, 6, 8, 10, 12]))
        n = int(rng.choice([96, 512, 1000]))
        # a stable-looking predictor: a dominant first tap below 1 and small higher taps
        coefs = [int(rng.integers(20000, 32000))] + [int(rng.integers(-9000, 9000)) for _ in range(order - 1)]
        res = rng.integers(-(1 << 16), 1 << 16, n)
        for _ in range(int(rng.integers(1, 6))):
            res[int(rng.integers(0, n))] = int(rng.choice([-1, 1])) * (1 << int(rng.integers(27, 31)))
        out.append((craft_lpc_block(res, coefs), n))
    return out
