"""The large fixtures (tests/golden/golden_large.json) are self-consistent, and the C oracle reproduces the first
blocks of config 4's range-addressable stream that the fixture's table digest covers (cheap CPU check; the full-size
comparison runs on the GPU, tests/test_gpu_large.py, and in bench.py's config-4 leg)."""
import json

import helpers as H

GOLD = json.loads((H.ROOT / "tests" / "golden" / "golden_large.json").read_text())


def test_shard_plans_cover_the_file():
    g = GOLD["C4_full_10h_24_48k_auto"]
    nb = g["n_blocks"]
    assert nb == (g["frames"] + 16383) // 16384 == 105469
    for world, ranks in g["shards"].items():
        per = (nb + int(world) - 1) // int(world)
        assert [r["first_block"] for r in ranks] == [min(nb, k * per) for k in range(int(world))]
        assert sum(r["blocks"] for r in ranks) == nb
        assert sum(r["payload_bytes"] for r in ranks) == g["payload_bytes"]
    assert g["len"] == 14 + 8 * nb + g["payload_bytes"]
    assert g["shards"]["1"][0]["table_sha256"] == g["table_sha256"]


def test_oracle_matches_first_range_prefix():
    """48 blocks of rank 0's range through the C oracle: same stream definition, same bytes as the product's
    multi-GPU test inputs (tests/test_multi_gpu.py uses the same generator call)."""
    g = GOLD["C4_full_10h_24_48k_auto"]
    frames = 48 * 16384
    l, r = H.synth_range(g["seed"], 0, frames, 24, 2, g["reset_log2"])
    a = H.oracle().encode(l, r, g["rate"], 24, 2)
    if H.have_ref():
        assert a == H.ref().encode(l, r, g["rate"], 24, 2, threads=4)
    assert len(a) > 14 + 8 * 48
