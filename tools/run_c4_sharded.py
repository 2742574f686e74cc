"""BASELINE config 4 on N GPUs of one box: the 10 h 24-bit / 48 kHz stereo file (auto LR/MS) cut into N
contiguous block ranges of 75 min (8 GPUs = the whole file), one process per GPU.  Launch with
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_c4_sharded.py
Each rank synthesises its own 75-minute range (generator seed 4 + rank: the Appendix C generator is
sequential, so ranges are independent signals of the same shape), encodes and decodes it on its GPU; the only
exchange is the NCCL all-gather of the per-rank payload byte counts from which every rank derives its global
payload offset.  Rank 0 prints one JSON line: aggregate PCM GB/s, max over ranks."""
import json, os, sys, time
from pathlib import Path
import numpy as np
import torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
from __graft_entry__ import load_package
from run_configs import synth_packed

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 4500
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keeps NCCL's banner off stdout
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cd = load_package().Codec(local)
rate, depth, ch, mode = 48000, 24, 2, 2
frames = rate * secs
pk = synth_packed(4 + rank, frames, depth, ch)
nb = (frames + 16383) // 16384
sizes = np.full(nb, 16384, dtype=np.uint32); sizes[-1] = frames - 16384 * (nb - 1)
d_in, d_out = cd.dev_malloc(pk.size), cd.dev_malloc(pk.size)
cd.h2d(d_in, pk)
counts = torch.zeros(world, dtype=torch.int64, device="cuda"); mine = torch.zeros(1, dtype=torch.int64, device="cuda")
def step():
    d_payload, nbytes, d_bb = cd.encode_device(d_in, 0, frames, depth, ch, mode)
    bb = cd.d2h(d_bb, nb * 4, np.uint32)
    mine[0] = nbytes
    dist.all_gather_into_tensor(counts, mine)
    torch.cuda.synchronize()
    t_mid = time.perf_counter()
    cd.decode_device(d_payload, nbytes, sizes, bb, depth, ch, mode, d_packed=d_out)
    return nbytes, t_mid
step()
assert np.array_equal(cd.d2h(d_out, pk.size), pk), "round trip differs"
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
K = 3
t0 = time.perf_counter(); te = 0.0
for _ in range(K):
    s0 = time.perf_counter(); nbytes, tm = step(); te += tm - s0
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
t = torch.tensor([wall, te, wall - te], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
offsets = torch.cumsum(counts, 0) - counts  # global payload offset of every rank's slab
if rank == 0:
    tot = pk.size * world
    print(json.dumps({"config": f"C4: {world} x {secs} s of the 10 h 24/48 stereo auto file, one range per GPU", "n_gpus": world,
                      "pcm_gb_total": tot / 1e9, "encode_decode_gbs": tot * K / float(t[0]) / 1e9,
                      "encode_gbs": tot * K / float(t[1]) / 1e9, "decode_gbs": tot * K / float(t[2]) / 1e9,
                      "payload_bytes_per_rank": counts.tolist(), "global_offsets": offsets.tolist(),
                      "roundtrip_exact": True}), flush=True)
dist.destroy_process_group()
