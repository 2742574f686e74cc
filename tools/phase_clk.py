"""Per-phase warp cycles of k_analyze from a -DLACB_PHASE_CLK build of the library (debug aid).
usage: phase_clk.py build/var/lib_PHASE.so [seconds]"""
import ctypes as C, sys
sys.path.insert(0, "tests")
import numpy as np, helpers as H
lib = sys.argv[1]; secs = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 120
cd = H.lacb_module().Codec(0, lib)
dll = C.CDLL(lib)
l, r, pk = H.synth(2, 96000 * secs, 24, want_packed=True)
buf = (C.c_ulonglong * 160)()
cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2)
dll.lacb_debug_phase_clk(buf, 1)
cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2)
dll.lacb_debug_phase_clk(buf, 1)
v = np.array(buf[:], dtype=np.float64)
tot = v[:64].sum()
fast_names = {1: "residual + zz (loop top)", 4: "bit-plane counts", 8: "base k + flags", 9: "wait barrier", 2: "finalize prev + static k",
              6: "bias proofs", 3: "hard chunks (in-warp)", 10: "kprev + costs", 11: "warp sums / slots", 12: "pass 1 (11 residuals, bounds, warp scans)",
              13: "wait pass 1", 14: "pass 1 warp offsets", 15: "wait offsets", 0: "load_block"}
names = {0: "load_block", 1: "residual(+loop top)", 2: "prepare->scan", 3: "wait scan", 4: "prepare rest", 5: "wait any4",
         6: "base k + flags (1-warp builds)", 7: "wait flg", 8: "base k, flags, bias, pairs", 9: "wait K / queues", 10: "walk+reduce", 11: "wait totals",
         12: "seg tables", 13: "wait seg", 14: "level scans", 15: "wait level", 16: "level select", 17: "final size", 18: "final sum"}
print("analyze_ms", cd.timing()["analyze_ms"])
for base, tag in ((0, "candidates"), (20, "levels/final")):
    for i in range(20):
        if v[base + i]:
            nm = fast_names.get(i, names.get(i, '')) if base == 0 and "--fast" in sys.argv else names.get(i, '')
            print(f"{tag:13s} {i:2d} {nm:36s} {100 * v[base + i] / tot:6.2f}%")
for i, nm in enumerate(("kbase/flags", "bias", "walk")):
    w = v[64 + 32 * i: 96 + 32 * i]
    if not w.any():
        continue
    print(f"busy per warp, {nm}: " + " ".join(f"{x / w.mean():.2f}" for x in w))
