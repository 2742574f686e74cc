"""k_analyze time per section type of the Appendix C signal (AR noise / triangle / sparse silence / stepped noise).
usage: section_timing.py [--only=<section 0..3>] lib.so [lib2.so ...]   (--only: one section, e.g. under ncu)"""
import sys
only = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--only=")]
sys.argv = [a for a in sys.argv if not a.startswith("--only=")]
sys.path.insert(0, "tests")
import numpy as np, helpers as H
secs = 240
l, r = H.synth(2, 96000 * secs, 24)
nb = l.size // 16384
l = l[: nb * 16384].reshape(nb, 16384); r = r[: nb * 16384].reshape(nb, 16384)
sec = (np.arange(nb) * 16384 >> 17) & 3
names = ["AR(4) noise", "triangle", "sparse silence", "stepped noise"]
for lib in sys.argv[1:]:
    cd = H.lacb_module().Codec(0, lib)
    for s in (only or range(4)):
        ll = np.ascontiguousarray(l[sec == s]).reshape(-1); rr = np.ascontiguousarray(r[sec == s]).reshape(-1)
        pk = np.zeros(ll.size * 6, dtype=np.uint8)
        both = np.stack([ll, rr], axis=1).reshape(-1)
        b = both.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        pk = np.ascontiguousarray(b).reshape(-1)
        for _ in range(2):
            cd.encode_blocks(None, None, 24, 1, packed=pk, channels=2)
        t = cd.timing()
        print(lib.split("/")[-1], names[s], "blocks", int((sec == s).sum()), "analyze_ms %.3f" % t["analyze_ms"],
              "us/chanblock %.1f" % (1e3 * t["analyze_ms"] / (2 * (sec == s).sum())), flush=True)
