"""ctypes binding of liblac_b200.so plus the host-side mirror of the reference's
frame codec interface (LAC::Encoder / LAC::Decoder, src/codec/lac/{encoder,decoder}.hpp).

Host responsibilities kept here, exactly as in the reference: argument validation,
the 10-byte frame header (src/codec/frame/frame_header.hpp:25-59), the big-endian
v3 block table (src/codec/lac/encoder.cpp:445-465) and its validation on decode
(src/codec/lac/decoder.cpp:88-159).  Everything per-block runs on the GPU through the
C ABI in include/lac_b200.h.  There is no CPU fallback: if the CUDA library is
missing or no device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import struct
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
DEFAULT_LIB = HERE / "liblac_b200.so"

MAX_BLOCK = 16384
LACB_PLANAR_I32, LACB_PACKED_LE = 0, 1
LACB_OK, LACB_EINVAL, LACB_EDECODE, LACB_ELIMIT, LACB_ECUDA, LACB_ENOMEM = 0, -1, -2, -3, -4, -5

SUPPORTED_RATES = (44100, 48000, 96000, 192000)
MAX_TOTAL_SAMPLES = 6_912_000_000          # lac/decoder.cpp:17-23
MAX_DECODED_PCM_BYTES = 1 << 30
MAX_BLOCK_COUNT = ((MAX_DECODED_PCM_BYTES // 4) + 255) // 256


class EncParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("sample_rate", "bit_depth", "channels", "stereo_mode",
                                          "zero_run_enabled", "partitioning_enabled", "validate_range")]


class DecParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("bit_depth", "channels", "stereo_mode")]


class Err(C.Structure):
    _fields_ = [("code", C.c_int32), ("block_index", C.c_uint32), ("reason", C.c_uint32), ("msg", C.c_char * 160)]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "prep_ms", "stereo_ms", "lpc_ms", "analyze_ms", "finalize_ms",
                                         "emit_ms", "d2h_ms", "parse_ms", "restore_ms", "finish_ms", "total_ms")]


class BlockInfo(C.Structure):
    _fields_ = [("predictor_type", C.c_uint32), ("order", C.c_uint32), ("partition_order", C.c_uint32),
                ("n_parts", C.c_uint32), ("taps", C.c_uint32), ("coeffs", C.c_int16 * 13),
                ("part_mode", C.c_uint8 * 256), ("part_k", C.c_uint8 * 256), ("cand_best_lo", C.c_uint32 * 11),
                ("bits", C.c_uint32)]


EXPORTS = ("lacb_create", "lacb_destroy", "lacb_last_error", "lacb_free", "lacb_device_count", "lacb_get_timing",
           "lacb_encode", "lacb_encode_to", "lacb_encode_device", "lacb_decode", "lacb_decode_device", "lacb_encode_block",
           "lacb_decode_block", "lacb_lpc_analyze", "lacb_last_block_info", "lacb_last_encode_decisions", "lacb_dev_malloc", "lacb_dev_free",
           "lacb_host_malloc", "lacb_host_free", "lacb_memcpy_h2d", "lacb_memcpy_d2h", "lacb_memcpy_d2d",
           "lacb_decode_block_at", "lacb_set_concurrency", "lacb_host_register", "lacb_host_unregister")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


def load_library(path=None) -> C.CDLL:
    path = Path(path) if path else DEFAULT_LIB
    if not path.exists():
        raise RuntimeError(f"{path} is missing: build it with `make -C {HERE} lib` (no CPU fallback exists)")
    lib = C.CDLL(str(path))
    for name in EXPORTS:
        getattr(lib, name)  # AttributeError if the ABI is incomplete
    lib.lacb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.lacb_destroy.argtypes = [C.c_void_p]
    lib.lacb_destroy.restype = None
    lib.lacb_last_error.argtypes = [C.c_void_p]
    lib.lacb_last_error.restype = C.c_char_p
    lib.lacb_free.argtypes = [C.c_void_p]
    lib.lacb_free.restype = None
    lib.lacb_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.lacb_encode.argtypes = [C.c_void_p, C.POINTER(EncParams), C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                C.POINTER(u8p), C.POINTER(C.c_uint64), u32p, C.POINTER(Err)]
    lib.lacb_encode_to.argtypes = [C.c_void_p, C.POINTER(EncParams), C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), u32p, C.POINTER(Err)]
    lib.lacb_encode_device.argtypes = [C.c_void_p, C.POINTER(EncParams), C.c_int, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                       C.POINTER(Err)]
    lib.lacb_decode.argtypes = [C.c_void_p, C.POINTER(DecParams), C.c_void_p, C.c_uint64, u32p, u32p, C.c_uint32,
                                C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Err)]
    lib.lacb_decode_device.argtypes = [C.c_void_p, C.POINTER(DecParams), C.c_void_p, C.c_uint64, u32p, u32p,
                                       C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Err)]
    lib.lacb_encode_block.argtypes = [C.c_void_p, i32p, C.c_uint32, C.c_int, C.c_int, C.POINTER(u8p),
                                      C.POINTER(C.c_uint64)]
    lib.lacb_decode_block.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, i32p, C.POINTER(C.c_uint64)]
    lib.lacb_decode_block_at.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, i32p,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    lib.lacb_set_concurrency.argtypes = [C.c_void_p, C.c_uint32]
    lib.lacb_lpc_analyze.argtypes = [C.c_void_p, i32p, C.c_uint32, C.c_int, C.POINTER(C.c_int16)]
    lib.lacb_last_block_info.argtypes = [C.c_void_p, C.POINTER(BlockInfo)]
    lib.lacb_dev_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.lacb_dev_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.lacb_host_malloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
    lib.lacb_host_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.lacb_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.lacb_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
    for fn in (lib.lacb_memcpy_h2d, lib.lacb_memcpy_d2h, lib.lacb_memcpy_d2d):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    return lib


class FrameHeader:
    """frame/frame_header.hpp:10-77."""
    SIZE = 10

    def __init__(self, channels=1, stereo_mode=0, sample_rate=44100, bit_depth=16, version=3):
        self.version, self.channels, self.stereo_mode = version, channels, stereo_mode
        self.sample_rate, self.bit_depth = sample_rate, bit_depth

    def pack(self) -> bytes:
        sr = self.sample_rate
        return bytes([0x4C, 0x41, self.version, self.channels, self.stereo_mode, (sr >> 8) & 0xFF, sr & 0xFF,
                      (sr >> 16) & 0xFF, self.bit_depth, 0])

    @staticmethod
    def parse(data: bytes) -> "FrameHeader | None":
        if len(data) < 10:
            return None
        sync = (data[0] << 8) | data[1]
        ver, ch, sm = data[2], data[3], data[4]
        sr = ((data[5] << 8) | data[6]) | (data[7] << 16)
        depth, reserved = data[8], data[9]
        ok = (sync == 0x4C41 and ver in (2, 3) and ch in (1, 2) and not (ch == 1 and sm != 0) and sm <= 2
              and sr in SUPPORTED_RATES and depth in (16, 24) and reserved == 0)
        return FrameHeader(ch, sm, sr, depth, ver) if ok else None


class DecodeError(RuntimeError):
    pass


class Codec:
    """One GPU context.  encode()/decode() mirror LAC::Encoder::encode / LAC::Decoder::decode."""

    def __init__(self, device: int = 0, lib_path=None):
        self.lib = load_library(lib_path)
        h = C.c_void_p()
        rc = self.lib.lacb_create(device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"lacb_create(device={device}) failed with {rc}: no usable CUDA device")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.lacb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return (self.lib.lacb_last_error(self.h) or b"").decode()

    def timing(self) -> dict:
        t = Timing()
        self.lib.lacb_get_timing(self.h, C.byref(t))
        return {n: getattr(t, n) for n, _ in Timing._fields_}

    # ---- device-resident path (inputs and outputs stay in HBM) -----------------------
    def dev_malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        if self.lib.lacb_dev_malloc(self.h, nbytes, C.byref(p)) != 0:
            raise RuntimeError("lacb_dev_malloc: " + self.last_error())
        return p.value

    def dev_free(self, ptr: int):
        self.lib.lacb_dev_free(self.h, C.c_void_p(ptr))

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        if self.lib.lacb_memcpy_h2d(self.h, C.c_void_p(dptr), arr.ctypes.data, arr.nbytes) != 0:
            raise RuntimeError("lacb_memcpy_h2d: " + self.last_error())

    def d2h(self, dptr: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if self.lib.lacb_memcpy_d2h(self.h, out.ctypes.data, C.c_void_p(dptr), nbytes) != 0:
            raise RuntimeError("lacb_memcpy_d2h: " + self.last_error())
        return out

    def d2h_to(self, host_ptr: int, dptr: int, nbytes: int):
        """device -> host copy to a raw host address (e.g. inside a mapped, registered output file)"""
        if self.lib.lacb_memcpy_d2h(self.h, C.c_void_p(host_ptr), C.c_void_p(dptr), nbytes) != 0:
            raise RuntimeError("lacb_memcpy_d2h: " + self.last_error())

    def host_register(self, host_ptr: int, nbytes: int) -> bool:
        return self.lib.lacb_host_register(self.h, C.c_void_p(host_ptr), nbytes) == 0

    def host_unregister(self, host_ptr: int):
        self.lib.lacb_host_unregister(self.h, C.c_void_p(host_ptr))

    def d2d(self, dst: int, src: int, nbytes: int):
        if self.lib.lacb_memcpy_d2d(self.h, C.c_void_p(dst), C.c_void_p(src), nbytes) != 0:
            raise RuntimeError("lacb_memcpy_d2d: " + self.last_error())

    def encode_device(self, d_a: int, d_b, frames: int, bit_depth, channels, stereo_mode, layout=LACB_PACKED_LE,
                      zero_run=True, partitioning=True):
        """PCM resident on the device -> (device payload pointer, payload bytes, device block-bytes pointer).
        The returned pointers live in the context workspace until the next call."""
        prm = EncParams(0, bit_depth, channels, stereo_mode, int(zero_run), int(partitioning), 0)
        dp, n, dbb, err = C.c_void_p(), C.c_uint64(), C.c_void_p(), Err()
        rc = self.lib.lacb_encode_device(self.h, C.byref(prm), layout, C.c_void_p(d_a),
                                         C.c_void_p(d_b) if d_b else None, frames, C.byref(dp), C.byref(n),
                                         C.byref(dbb), C.byref(err))
        if rc != 0:
            raise RuntimeError(f"lacb_encode_device rc={rc}: {self.last_error()}")
        return dp.value, n.value, dbb.value

    def decode_device(self, d_payload: int, payload_bytes: int, block_sizes, block_bytes, bit_depth, channels,
                      stereo_mode, d_packed: int = 0, d_left: int = 0, d_right: int = 0):
        bs = np.ascontiguousarray(block_sizes, dtype=np.uint32)
        bb = np.ascontiguousarray(block_bytes, dtype=np.uint32)
        prm = DecParams(bit_depth, channels, stereo_mode)
        err = Err()
        rc = self.lib.lacb_decode_device(self.h, C.byref(prm), C.c_void_p(d_payload), payload_bytes,
                                         bs.ctypes.data_as(u32p), bb.ctypes.data_as(u32p), bs.size,
                                         C.c_void_p(d_left) if d_left else None,
                                         C.c_void_p(d_right) if d_right else None,
                                         C.c_void_p(d_packed) if d_packed else None, C.byref(err))
        if rc == LACB_EDECODE:
            raise DecodeError(err.msg.decode())
        if rc != 0:
            raise RuntimeError(f"lacb_decode_device rc={rc}: {self.last_error()}")

    # ---- host buffers owned by the caller (pinned memory = full PCIe speed) ------------
    def pinned(self, nbytes: int) -> np.ndarray:
        """uint8 array over page-locked host memory (lives as long as the context)."""
        p = C.c_void_p()
        if self.lib.lacb_host_malloc(self.h, nbytes, C.byref(p)) != 0:
            raise RuntimeError("lacb_host_malloc: " + self.last_error())
        self._pinned = getattr(self, "_pinned", []) + [p]
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(max(nbytes, 1),))[:nbytes]

    def encode_into(self, pcm_packed: np.ndarray, payload_buf: np.ndarray, block_bytes: np.ndarray, bit_depth,
                    channels, stereo_mode, zero_run=True, partitioning=True) -> int:
        """Packed PCM (host) -> payload written into payload_buf; returns the payload size."""
        frames = pcm_packed.size // (channels * (bit_depth // 8))
        prm = EncParams(0, bit_depth, channels, stereo_mode, int(zero_run), int(partitioning), 0)
        n, err = C.c_uint64(), Err()
        rc = self.lib.lacb_encode_to(self.h, C.byref(prm), LACB_PACKED_LE, pcm_packed.ctypes.data, None, frames,
                                     payload_buf.ctypes.data, payload_buf.size, C.byref(n),
                                     block_bytes.ctypes.data_as(u32p), C.byref(err))
        if rc != 0:
            raise RuntimeError(f"lacb_encode_to rc={rc} (needs {n.value} bytes): {self.last_error()}")
        return n.value

    def decode_into(self, payload: np.ndarray, block_sizes, block_bytes, bit_depth, channels, stereo_mode,
                    out_packed: np.ndarray):
        bs = np.ascontiguousarray(block_sizes, dtype=np.uint32)
        bb = np.ascontiguousarray(block_bytes, dtype=np.uint32)
        prm = DecParams(bit_depth, channels, stereo_mode)
        err = Err()
        rc = self.lib.lacb_decode(self.h, C.byref(prm), payload.ctypes.data, payload.size, bs.ctypes.data_as(u32p),
                                  bb.ctypes.data_as(u32p), bs.size, LACB_PACKED_LE, out_packed.ctypes.data, None,
                                  C.byref(err))
        if rc == LACB_EDECODE:
            raise DecodeError(err.msg.decode())
        if rc != 0:
            raise RuntimeError(f"lacb_decode rc={rc}: {self.last_error()}")

    # ---- block payloads ------------------------------------------------------------
    def encode_blocks(self, left, right=None, bit_depth=16, stereo_mode=0, zero_run=True, partitioning=True,
                      packed=None, channels=None, sample_rate=44100, validate=True):
        """Returns (payload bytes as np.uint8 array, block_bytes uint32 array, block_sizes uint32 array)."""
        if packed is not None:
            packed = np.ascontiguousarray(packed, dtype=np.uint8)
            ch = channels
            frames = packed.size // (ch * (bit_depth // 8))
            a, b, layout = packed.ctypes.data, None, LACB_PACKED_LE
        else:
            left = np.ascontiguousarray(left, dtype=np.int32)
            ch = 2 if right is not None and len(right) else 1
            frames = left.size
            if ch == 2:
                right = np.ascontiguousarray(right, dtype=np.int32)
            a, b, layout = left.ctypes.data, (right.ctypes.data if ch == 2 else None), LACB_PLANAR_I32
        prm = EncParams(sample_rate, bit_depth, ch, stereo_mode, int(zero_run), int(partitioning), int(validate))
        nb = (frames + MAX_BLOCK - 1) // MAX_BLOCK
        bb = np.zeros(max(nb, 1), dtype=np.uint32)
        out, n, err = u8p(), C.c_uint64(), Err()
        rc = self.lib.lacb_encode(self.h, C.byref(prm), layout, a, b, frames, C.byref(out), C.byref(n),
                                  bb.ctypes.data_as(u32p), C.byref(err))
        if rc == LACB_EINVAL:
            raise ValueError(self.last_error())
        if rc != 0:
            raise RuntimeError(f"lacb_encode rc={rc}: {self.last_error()}")
        payload = np.ctypeslib.as_array(out, shape=(max(n.value, 1),))[: n.value].copy()
        self.lib.lacb_free(out)
        sizes = np.full(nb, MAX_BLOCK, dtype=np.uint32)
        sizes[-1] = frames - MAX_BLOCK * (nb - 1)
        return payload, bb[:nb], sizes

    def decode_blocks(self, payload, block_sizes, block_bytes, bit_depth, channels, stereo_mode, packed=False):
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        bs = np.ascontiguousarray(block_sizes, dtype=np.uint32)
        bb = np.ascontiguousarray(block_bytes, dtype=np.uint32) if block_bytes is not None else None
        bbp = bb.ctypes.data_as(u32p) if bb is not None else None  # None: serial v2 stream
        frames = int(bs.astype(np.uint64).sum())
        prm = DecParams(bit_depth, channels, stereo_mode)
        err = Err()
        if packed:
            out = np.zeros(frames * channels * (bit_depth // 8), dtype=np.uint8)
            rc = self.lib.lacb_decode(self.h, C.byref(prm), payload.ctypes.data, payload.size, bs.ctypes.data_as(u32p),
                                      bbp, bs.size, LACB_PACKED_LE, out.ctypes.data, None, C.byref(err))
            res = (out,)
        else:
            left = np.zeros(frames, dtype=np.int32)
            right = np.zeros(frames if channels == 2 else 0, dtype=np.int32)
            rc = self.lib.lacb_decode(self.h, C.byref(prm), payload.ctypes.data, payload.size, bs.ctypes.data_as(u32p),
                                      bbp, bs.size, LACB_PLANAR_I32, left.ctypes.data,
                                      right.ctypes.data if channels == 2 else None, C.byref(err))
            res = (left, right)
        if rc == LACB_EDECODE:
            raise DecodeError(err.msg.decode())
        if rc != 0:
            raise RuntimeError(f"lacb_decode rc={rc}: {self.last_error()}")
        return res

    # ---- whole .lac frames (LAC::Encoder::encode / LAC::Decoder::decode) --------------
    def encode(self, left, right=None, sample_rate=44100, bit_depth=16, stereo_mode=0, zero_run=True,
               partitioning=True) -> bytes:
        left = np.asarray(left)
        has_r = right is not None and len(right) > 0
        if left.size == 0:
            raise ValueError("left channel is empty")
        if has_r and len(right) != left.size:
            raise ValueError("channel size mismatch")
        if sample_rate not in SUPPORTED_RATES:
            raise ValueError("unsupported sample rate")
        if bit_depth not in (16, 24):
            raise ValueError("unsupported bit depth")
        if stereo_mode > 2:
            raise ValueError("invalid stereo mode")
        payload, bb, sizes = self.encode_blocks(left, right if has_r else None, bit_depth, stereo_mode, zero_run,
                                                partitioning, sample_rate=sample_rate)
        hdr = FrameHeader(2 if has_r else 1, stereo_mode if has_r else 0, sample_rate, bit_depth).pack()
        table = np.empty((sizes.size, 2), dtype=">u4")
        table[:, 0] = sizes
        table[:, 1] = bb
        return hdr + struct.pack(">I", sizes.size) + table.tobytes() + payload.tobytes()

    def decode(self, data: bytes):
        """Returns (left, right, header dict); raises DecodeError("[decode-error] ...")."""
        def fail(msg):
            raise DecodeError("[decode-error] " + msg)
        if not data:
            fail("empty input")
        hdr = FrameHeader.parse(data)
        if hdr is None:
            fail("invalid frame header")
        v3 = hdr.version >= 3
        body = memoryview(data)[10:]
        if len(body) < 4:
            fail("invalid block count")
        nb = struct.unpack(">I", body[:4])[0]
        if nb == 0 or nb > MAX_BLOCK_COUNT:
            fail("invalid block count")
        words = 2 if v3 else 1
        if nb > ((len(body) - 4) * 8) // (32 * words):
            fail("truncated block size table")
        table = np.frombuffer(body[4:4 + 4 * words * nb], dtype=">u4").reshape(nb, words).astype(np.uint64)
        sizes = table[:, 0]
        cbytes = table[:, 1] if v3 else np.ones(nb, dtype=np.uint64)
        bad = (sizes == 0) | (sizes > MAX_BLOCK)
        bad[:-1] |= sizes[:-1] < 256
        csum_s, csum_b = np.cumsum(sizes), np.cumsum(cbytes if v3 else np.zeros(nb, dtype=np.uint64))
        avail = len(body) - 4 - 4 * words * nb
        # errors are reported in table order, interleaved as the reference's loop does
        for i in range(nb) if (bad.any() or (cbytes == 0).any() or csum_s[-1] > MAX_TOTAL_SAMPLES
                               or csum_b[-1] > len(body)) else ():
            if bad[i]:
                fail("invalid block size")
            if csum_s[i] > MAX_TOTAL_SAMPLES:
                fail("total samples exceed maximum")
            if cbytes[i] == 0:
                fail("invalid compressed block size")
            if csum_b[i] > len(body):
                fail("compressed block sizes exceed frame payload")
        total = int(csum_s[-1])
        if total * hdr.channels * 4 > MAX_DECODED_PCM_BYTES:
            fail("decoded PCM allocation exceeds maximum")
        wav = total * hdr.channels * (hdr.bit_depth // 8)
        if 36 + wav + (wav & 1) > 0xFFFFFFFF:
            fail("decoded WAV data exceeds RIFF limit")
        if v3 and int(csum_b[-1]) != avail:
            fail("compressed block sizes do not match frame payload")
        payload = np.frombuffer(body[4 + 4 * words * nb:], dtype=np.uint8)
        if not v3:
            left, right = self.decode_blocks(payload, sizes.astype(np.uint32), None, hdr.bit_depth, hdr.channels,
                                             hdr.stereo_mode)
            return left, right, dict(channels=hdr.channels, sample_rate=hdr.sample_rate, bit_depth=hdr.bit_depth,
                                     stereo_mode=hdr.stereo_mode)
        payload = np.frombuffer(body[4 + 8 * nb:], dtype=np.uint8)
        left, right = self.decode_blocks(payload, sizes.astype(np.uint32), cbytes.astype(np.uint32), hdr.bit_depth,
                                         hdr.channels, hdr.stereo_mode)
        return left, right, dict(channels=hdr.channels, sample_rate=hdr.sample_rate, bit_depth=hdr.bit_depth,
                                 stereo_mode=hdr.stereo_mode)

    # ---- block-level hooks (Block::Encoder / Block::Decoder / LPC) --------------------
    def block_encode(self, pcm, zero_run=True, partitioning=True, want_info=False):
        a = np.ascontiguousarray(pcm, dtype=np.int32)
        out, n = u8p(), C.c_uint64()
        rc = self.lib.lacb_encode_block(self.h, a.ctypes.data_as(i32p), a.size, int(zero_run), int(partitioning),
                                        C.byref(out), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"lacb_encode_block rc={rc}: {self.last_error()}")
        data = np.ctypeslib.as_array(out, shape=(max(n.value, 1),))[: n.value].tobytes()
        self.lib.lacb_free(out)
        if want_info:
            info = BlockInfo()
            self.lib.lacb_last_block_info(self.h, C.byref(info))
            return data, info
        return data

    def block_decode(self, data: bytes, block_size: int):
        buf = np.frombuffer(data if data else b"\0", dtype=np.uint8)
        out = np.zeros(max(1, block_size), dtype=np.int32)
        bits = C.c_uint64()
        rc = self.lib.lacb_decode_block(self.h, buf.ctypes.data, len(data), block_size, out.ctypes.data_as(i32p),
                                        C.byref(bits))
        if rc < 0:
            raise RuntimeError(f"lacb_decode_block rc={rc}: {self.last_error()}")
        return bool(rc), out[:block_size], bits.value

    def block_decode_at(self, data: bytes, bit_offset: int, block_size: int):
        """Block::Decoder::decode_into from an arbitrary bit position -> (ok, pcm, bits consumed, ran_out)."""
        buf = np.frombuffer(data if data else b"\0", dtype=np.uint8)
        out = np.zeros(max(1, block_size), dtype=np.int32)
        bits, ran = C.c_uint64(), C.c_int()
        rc = self.lib.lacb_decode_block_at(self.h, buf.ctypes.data, len(data), bit_offset, block_size,
                                           out.ctypes.data_as(i32p), C.byref(bits), C.byref(ran))
        if rc < 0:
            raise RuntimeError(f"lacb_decode_block_at rc={rc}: {self.last_error()}")
        return bool(rc), out[:block_size], bits.value, bool(ran.value)

    def set_concurrency(self, n: int):
        self.lib.lacb_set_concurrency(self.h, n)

    def lpc_analyze(self, pcm, order: int):
        a = np.ascontiguousarray(pcm, dtype=np.int32)
        c = (C.c_int16 * (order + 1))()
        used = self.lib.lacb_lpc_analyze(self.h, a.ctypes.data_as(i32p), a.size, order, c)
        if used < 0:
            raise RuntimeError(f"lacb_lpc_analyze rc={used}")
        return used, np.array(c[:], dtype=np.int16)
