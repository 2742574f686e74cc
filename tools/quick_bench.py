"""Early timing probe (not the graded bench): encode+decode one synthetic file on cuda:0."""
import sys, time
sys.path.insert(0, "tests")
import numpy as np, helpers as H
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cd = H.gpu_codec()
l, r, pk = H.synth(2, int(96000 * secs), 24, want_packed=True)
pcm_bytes = pk.size
for it in range(3):
    t0 = time.time(); payload, bb, sz = cd.encode_blocks(None, None, 24, mode, packed=pk, channels=2); t1 = time.time()
    te = cd.timing()
    out, = cd.decode_blocks(payload, sz, bb, 24, 2, mode, packed=True); t2 = time.time()
    td = cd.timing()
    print(f"it{it}: pcm {pcm_bytes/1e6:.1f} MB lac {payload.size/1e6:.1f} MB  enc wall {t1-t0:.3f}s ({pcm_bytes/(t1-t0)/1e9:.2f} GB/s)  dec wall {t2-t1:.3f}s ({pcm_bytes/(t2-t1)/1e9:.2f} GB/s) ok={np.array_equal(out, pk)}")
    print("   enc", {k: round(v, 3) for k, v in te.items() if v})
    print("   dec", {k: round(v, 3) for k, v in td.items() if v})
