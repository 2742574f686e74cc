"""LAC_TIMING milestones of one-shot `lac_cli` runs on the C2 file, with and without page-locked file mappings
(LAC_PIN_FILES), plus the bare CUDA start-up (`lac_cli selftest`-free probe: a decode of a tiny file).
usage: python tools/cli_breakdown.py"""
import json, os, subprocess, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import helpers as H
from cli_timing_lib import write_wav
CLI = str(H.PKG_DIR / "host" / "lac_cli")
tmp = "/dev/shm/lacb_cli"; os.makedirs(tmp, exist_ok=True)
l, r, pk = H.synth(2, 96000 * 600, 24, want_packed=True)
write_wav(f"{tmp}/in.wav", pk, 2, 96000, 24)
l, r, pk1 = H.synth(1, 4096, 16, want_packed=True)
write_wav(f"{tmp}/tiny.wav", pk1, 2, 44100, 16)
def ms(cmd, env):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    return {"wall_ms": round((time.perf_counter() - t0) * 1e3, 1),
            "milestones": [ln.split("] ")[1] + "@" + ln.split(" ")[1] for ln in p.stderr.splitlines() if ln.startswith("[lac_cli")]}
for pin in ("0", "1"):
    env = dict(os.environ, LAC_TIMING="1", LAC_PIN_FILES=pin)
    for rep in range(2):
        print(json.dumps({"pin": pin, "encode": ms([CLI, "encode", f"{tmp}/in.wav", f"{tmp}/o.lac", "--stereo-mode=ms"], env),
                          "decode": ms([CLI, "decode", f"{tmp}/o.lac", f"{tmp}/o.wav"], env)}), flush=True)
env = dict(os.environ, LAC_TIMING="1")
for rep in range(3):
    print(json.dumps({"tiny": ms([CLI, "encode", f"{tmp}/tiny.wav", f"{tmp}/t.lac"], env)}), flush=True)
