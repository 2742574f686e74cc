"""N>1 path on CPU: two gloo ranks each encode their block range (through the emulator
build of the kernels), all-gather their payload byte counts, and rank 0 assembles the
frame -- which must be byte-identical to the oracle's single-process encode."""
import os
import importlib.util

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

import helpers as H


def _sharding():
    spec = importlib.util.spec_from_file_location("sharding", str(H.PKG_DIR / "sharding.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _worker(rank, world, port, frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = _sharding()
    l, r = H.synth(4, frames, 24)
    f0, n = S.plan_shards(frames, world)[rank]
    cd = H.emu_codec()
    payload, bb, sizes = cd.encode_blocks(l[f0:f0 + n], r[f0:f0 + n], 24, 2)
    mine = torch.tensor([payload.size], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, mine)
    off = S.exclusive_offsets([int(c) for c in counts])
    # rank 0 collects the slabs (on the GPU box each rank's slab goes D2H over its own PCIe link)
    gathered = [None] * world
    dist.gather_object((payload.tobytes(), bb, sizes, int(off[rank])), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        hdr = H.lacb_module().FrameHeader(2, 2, 48000, 24).pack()
        pos = 0
        for slab, _, _, o in gathered:
            assert o == pos
            pos += len(slab)
        blob = S.assemble_frame(hdr, [g[2] for g in gathered], [g[1] for g in gathered], [g[0] for g in gathered])
        want = H.oracle().encode(l, r, 48000, 24, 2)
        q.put(blob == want)
    dist.destroy_process_group()


def test_two_rank_block_range_sharding():
    frames = 5 * 16384 + 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_plan_shards_covers_everything():
    S = _sharding()
    for frames in (1, 16384, 16385, 10 * 16384 + 5, 1_728_000_000):
        for world in (1, 2, 4, 8):
            plan = S.plan_shards(frames, world)
            assert sum(n for _, n in plan) == frames
            pos = 0
            for f0, n in plan:
                assert f0 == pos or n == 0
                assert f0 % 16384 == 0
                pos += n
