"""Host-path timing probe: lacb_encode_to / lacb_decode with host buffers, one context."""
import sys, time
sys.path.insert(0, "tests")
import numpy as np, helpers as H
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 600
import os
cd = H.lacb_module().Codec(0, os.environ["LACB_LIB"]) if os.environ.get("LACB_LIB") else H.gpu_codec()  # LACB_LIB: an experimental build
l, r, pk = H.synth(2, 96000 * secs, 24, want_packed=True)
frames = 96000 * secs; nb = (frames + 16383) // 16384
sizes = np.full(nb, 16384, dtype=np.uint32); sizes[-1] = frames - 16384 * (nb - 1)
if os.environ.get("PAGEABLE"):  # what a one-shot lac_cli hands over: plain (mapped) memory, first call included
    h_in = pk.copy(); h_pay = np.zeros(pk.size + (pk.size >> 2) + 4096, dtype=np.uint8); h_out = np.zeros(pk.size, dtype=np.uint8)
else:
    h_in = cd.pinned(pk.size); h_in[:] = pk
    h_pay = cd.pinned(pk.size + (pk.size >> 2) + 4096); h_out = cd.pinned(pk.size)
bb = np.zeros(nb, dtype=np.uint32)
for it in range(3):
    t0 = time.perf_counter(); n = cd.encode_into(h_in, h_pay, bb, 24, 2, 1); t1 = time.perf_counter()
    cd.decode_into(h_pay[:n], sizes, bb, 24, 2, 1, h_out); t2 = time.perf_counter()
    print(f"it{it}: encode {1e3*(t1-t0):.1f} ms ({pk.size/(t1-t0)/1e9:.2f} GB/s)  decode {1e3*(t2-t1):.1f} ms ({pk.size/(t2-t1)/1e9:.2f} GB/s)", flush=True)
assert np.array_equal(h_out, pk)
