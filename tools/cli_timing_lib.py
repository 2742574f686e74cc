import struct
def write_wav(path, packed, channels, rate, depth):
    align = channels * depth // 8
    n = packed.size
    hdr = b"RIFF" + struct.pack("<I", 36 + n + (n & 1)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * align, align, depth)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", n)); f.write(packed.tobytes()); f.write(b"\0" if n & 1 else b"")
