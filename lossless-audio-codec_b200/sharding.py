"""Block-range sharding of one long encode across ranks (one process per GPU).

Blocks are independent fixed 16384-frame units (src/codec/lac/encoder.cpp:59-69), so a
long input shards by contiguous block range with no data-path collective.  The only
exchange is an all-gather of each rank's payload byte count (and, for the table, its
per-block byte sizes): global payload offsets are the exclusive scan of those counts and
rank 0 (the host) concatenates header + table + slabs in rank order, exactly what
src/codec/lac/encoder.cpp:445-465 does for its worker results.
"""
from __future__ import annotations

import struct

import numpy as np

MAX_BLOCK = 16384


def plan_shards(frames: int, world: int):
    """[(first_frame, n_frames)] per rank: contiguous block ranges of ceil(n_blocks/world) blocks."""
    nb = (frames + MAX_BLOCK - 1) // MAX_BLOCK
    per = (nb + world - 1) // world
    out = []
    for r in range(world):
        b0, b1 = min(nb, r * per), min(nb, (r + 1) * per)
        f0, f1 = b0 * MAX_BLOCK, min(frames, b1 * MAX_BLOCK)
        out.append((f0, max(0, f1 - f0)))
    return out


def exclusive_offsets(counts):
    counts = np.asarray(counts, dtype=np.uint64)
    off = np.zeros(counts.size + 1, dtype=np.uint64)
    np.cumsum(counts, out=off[1:])
    return off


def assemble_frame(header: bytes, block_sizes, block_bytes, slabs) -> bytes:
    """.lac v3 frame from per-rank results (rank order)."""
    sizes = np.concatenate([np.asarray(s, dtype=np.uint32) for s in block_sizes])
    cbytes = np.concatenate([np.asarray(b, dtype=np.uint32) for b in block_bytes])
    table = np.empty((sizes.size, 2), dtype=">u4")
    table[:, 0] = sizes
    table[:, 1] = cbytes
    return header + struct.pack(">I", sizes.size) + table.tobytes() + b"".join(bytes(s) for s in slabs)
