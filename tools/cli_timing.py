"""`lac_cli encode` / `lac_cli decode` wall times on the GPU box next to the unmodified reference CLI
(oracle/_ref/lac_cli_ref) on the same files: BASELINE configs[0] (60 s 16/44.1 auto) and configs[1]
(600 s 24/96 --stereo-mode=ms).  Writes one JSON line per config; files live in /dev/shm.
Measurement tool (like bench.py's cpu_baseline leg): the reference CLI is only timed and compared, never used by
the product.
usage: cli_timing.py [out.jsonl]"""
import json, os, struct, subprocess, sys, time
sys.path.insert(0, "tests")
import numpy as np, helpers as H

CLI = str(H.PKG_DIR / "host" / "lac_cli")
REF = str(H.REF_CLI)
tmp = "/dev/shm/lacb_cli" if os.path.isdir("/dev/shm") else "/tmp/lacb_cli"
os.makedirs(tmp, exist_ok=True)

def write_wav(path, packed, channels, rate, depth):
    align = channels * depth // 8
    n = packed.size
    hdr = b"RIFF" + struct.pack("<I", 36 + n + (n & 1)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * align, align, depth)
    with open(path, "wb") as f:
        f.write(hdr + b"data" + struct.pack("<I", n)); f.write(packed.tobytes()); f.write(b"\0" if n & 1 else b"")

def run(*args):
    t0 = time.perf_counter()
    r = subprocess.run(list(args), capture_output=True, text=True)
    return time.perf_counter() - t0, r

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
threads = os.cpu_count() or 1
for name, seed, secs, rate, depth, flag in (("C1 60 s 16/44.1 stereo auto", 1, 60, 44100, 16, None),
                                            ("C2 600 s 24/96 stereo --stereo-mode=ms", 2, 600, 96000, 24, "--stereo-mode=ms")):
    l, r, pk = H.synth(seed, rate * secs, depth, want_packed=True)
    wav = f"{tmp}/in.wav"; write_wav(wav, pk, 2, rate, depth)
    rec = {"config": name, "pcm_mb": pk.size / 1e6, "host_threads": threads}
    extra = [flag] if flag else []
    for tag, exe in (("gpu", CLI), ("ref", REF)):
        if not os.path.exists(exe):
            continue
        lac, back = f"{tmp}/{tag}.lac", f"{tmp}/{tag}.wav"
        best_e = best_d = 1e9
        for _ in range(2):
            te, re_ = run(exe, "encode", wav, lac, f"--threads={threads}", *extra)
            assert re_.returncode == 0, re_.stderr
            td, rd = run(exe, "decode", lac, back, f"--threads={threads}")
            assert rd.returncode == 0, rd.stderr
            best_e, best_d = min(best_e, te), min(best_d, td)
        assert open(back, "rb").read() == open(wav, "rb").read(), "round trip differs"
        rec[f"{tag}_encode_s"], rec[f"{tag}_decode_s"] = round(best_e, 3), round(best_d, 3)
        rec[f"{tag}_lac_bytes"] = os.path.getsize(lac)
    if "ref_lac_bytes" in rec:
        rec["lac_identical"] = open(f"{tmp}/gpu.lac", "rb").read() == open(f"{tmp}/ref.lac", "rb").read()
    out.write(json.dumps(rec) + "\n"); out.flush()
    print(rec, file=sys.stderr)

# many short files: one process per file (reference and GPU CLI) against `lac_cli batch`
l, r, pk = H.synth(1, 44100 * 60, 16, want_packed=True)
N = 16
for i in range(N):
    write_wav(f"{tmp}/s{i}.wav", pk, 2, 44100, 16)
rec = {"config": f"{N} files of 60 s 16/44.1 stereo auto, encode+decode each", "pcm_mb": N * pk.size / 1e6, "host_threads": threads}
lst = f"{tmp}/list.txt"
open(lst, "w").write("".join(f"encode {tmp}/s{i}.wav {tmp}/s{i}.lac\ndecode {tmp}/s{i}.lac {tmp}/s{i}.out.wav\n" for i in range(N)))
t, rr = run(CLI, "batch", lst)
assert rr.returncode == 0, rr.stderr
rec["gpu_batch_s"] = round(t, 3)
for tag, exe in (("gpu_per_process", CLI), ("ref_per_process", REF)):
    if not os.path.exists(exe):
        continue
    t0 = time.perf_counter()
    for i in range(N):
        assert subprocess.run([exe, "encode", f"{tmp}/s{i}.wav", f"{tmp}/s{i}.lac2", f"--threads={threads}"], capture_output=True).returncode == 0
        assert subprocess.run([exe, "decode", f"{tmp}/s{i}.lac2", f"{tmp}/s{i}.out2.wav", f"--threads={threads}"], capture_output=True).returncode == 0
    rec[tag + "_s"] = round(time.perf_counter() - t0, 3)
assert open(f"{tmp}/s3.out.wav", "rb").read() == open(f"{tmp}/s3.wav", "rb").read()
out.write(json.dumps(rec) + "\n"); out.flush()
print(rec, file=sys.stderr)
