"""BASELINE configs 3 and 4 at FULL size on the GPU against SHA-256 fixtures recorded from the unmodified
reference (tests/golden/golden_large.json, tools/make_golden_large.py): config 3 (30 min 24/192 mono, 1.04 GB of
PCM) as a whole .lac, config 4 (the 10 h 24/48 stereo file) as the payload slab + table slice of one rank of the
8-way block-range sharding, input generated for that range only.  Decode is checked against the input."""
import hashlib
import json
import struct

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = json.loads((H.ROOT / "tests" / "golden" / "golden_large.json").read_text())


def test_config3_full_size_sha():
    g = GOLD["C3_full_1800s_24_192k_mono"]
    cd = H.gpu_codec()
    _, _, pk = H.synth_range(g["seed"], 0, g["frames"], 24, 1, reset_log2=0, planes=False, want_packed=True)
    payload, bb, sizes = cd.encode_blocks(None, None, 24, 0, packed=pk, channels=1, sample_rate=g["rate"])
    assert payload.size == g["payload_bytes"]
    assert hashlib.sha256(memoryview(payload)).hexdigest() == g["payload_sha256"]
    table = np.empty((sizes.size, 2), dtype=">u4")
    table[:, 0], table[:, 1] = sizes, bb
    assert hashlib.sha256(table.tobytes()).hexdigest() == g["table_sha256"]
    h = hashlib.sha256(H.lacb_module().FrameHeader(1, 0, g["rate"], 24).pack() + struct.pack(">I", sizes.size) + table.tobytes())
    h.update(memoryview(payload))
    assert h.hexdigest() == g["sha256"] and 14 + table.nbytes + payload.size == g["len"]
    (back,) = cd.decode_blocks(payload, sizes, bb, 24, 1, 0, packed=True)
    assert np.array_equal(back, pk)


@pytest.mark.parametrize("rank", [0, 5, 7])
def test_config4_rank_of_eight_sha(rank):
    g = GOLD["C4_full_10h_24_48k_auto"]
    sh = g["shards"]["8"][rank]
    f0 = sh["first_block"] * 16384
    fr = min(g["frames"], (sh["first_block"] + sh["blocks"]) * 16384) - f0
    cd = H.gpu_codec()
    _, _, pk = H.synth_range(g["seed"], f0, fr, 24, 2, g["reset_log2"], planes=False, want_packed=True)
    payload, bb, sizes = cd.encode_blocks(None, None, 24, 2, packed=pk, channels=2, sample_rate=g["rate"])
    assert payload.size == sh["payload_bytes"]
    assert hashlib.sha256(memoryview(payload)).hexdigest() == sh["payload_sha256"]
    table = np.empty((sizes.size, 2), dtype=">u4")
    table[:, 0], table[:, 1] = sizes, bb
    assert hashlib.sha256(table.tobytes()).hexdigest() == sh["table_sha256"]
    (back,) = cd.decode_blocks(payload, sizes, bb, 24, 2, 2, packed=True)
    assert np.array_equal(back, pk)
