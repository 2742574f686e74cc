"""lac_cli one-shot wall time against CUDA_DEVICE_MAX_CONNECTIONS (8 = CUDA default, 32 = what lac_cli sets).
usage (GPU box): python tools/cli_conn_check.py"""
import os, subprocess, sys, time, wave
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H

CLI = str(ROOT / "lossless-audio-codec_b200" / "host" / "lac_cli")
l, r, pk = H.synth(1, 44100 * 60, 16, want_packed=True)
wav = Path("/dev/shm/conn_c1.wav")
with wave.open(str(wav), "wb") as w:
    w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pk.tobytes())
for conns in ("8", "32", "8", "32"):
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS=conns)
    lac = f"/dev/shm/conn_{conns}.lac"
    out = f"/dev/shm/conn_{conns}.wav"
    te, td = [], []
    for _ in range(3):
        t0 = time.perf_counter(); subprocess.run([CLI, "encode", str(wav), lac], env=env, check=True, capture_output=True)
        t1 = time.perf_counter(); subprocess.run([CLI, "decode", lac, out], env=env, check=True, capture_output=True)
        t2 = time.perf_counter(); te.append(t1 - t0); td.append(t2 - t1)
    print(f"conns={conns}: encode {min(te):.3f} s  decode {min(td):.3f} s (best of 3)", flush=True)
