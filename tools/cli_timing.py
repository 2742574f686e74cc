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
# a resident `lac_cli serve` for the "gpu_server" columns: the same command lines, forwarded (LAC_SERVER)
SOCK = f"{tmp}/srv.sock"
server = subprocess.Popen([CLI, "serve", SOCK], stdout=subprocess.PIPE, text=True)
assert "serving on" in server.stdout.readline()
SENV = dict(os.environ, LAC_SERVER=SOCK)

def run_env(env, *args):
    t0 = time.perf_counter()
    r = subprocess.run(list(args), capture_output=True, text=True, env=env)
    return time.perf_counter() - t0, r

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
    best_e = best_d = 1e9
    for _ in range(3):
        te, re_ = run_env(SENV, CLI, "encode", wav, f"{tmp}/srv.lac", f"--threads={threads}", *extra)
        assert re_.returncode == 0, re_.stderr
        td, rd = run_env(SENV, CLI, "decode", f"{tmp}/srv.lac", f"{tmp}/srv.wav", f"--threads={threads}")
        assert rd.returncode == 0, rd.stderr
        best_e, best_d = min(best_e, te), min(best_d, td)
    assert open(f"{tmp}/srv.wav", "rb").read() == open(wav, "rb").read(), "round trip differs (server)"
    assert open(f"{tmp}/srv.lac", "rb").read() == open(f"{tmp}/gpu.lac", "rb").read()
    rec["gpu_server_encode_s"], rec["gpu_server_decode_s"] = round(best_e, 3), round(best_d, 3)
    if "ref_lac_bytes" in rec:
        rec["lac_identical"] = open(f"{tmp}/gpu.lac", "rb").read() == open(f"{tmp}/ref.lac", "rb").read()
    out.write(json.dumps(rec) + "\n"); out.flush()
    print(rec, file=sys.stderr)

# where a one-shot run spends its time (LAC_TIMING=1 milestones of the GPU CLI on the C2 file)
env = dict(os.environ, LAC_TIMING="1")
bd = {}
for cmdline in ([CLI, "encode", wav, f"{tmp}/gpu.lac", "--stereo-mode=ms"], [CLI, "decode", f"{tmp}/gpu.lac", f"{tmp}/gpu.wav"]):
    r_ = subprocess.run(cmdline, capture_output=True, text=True, env=env)
    bd[cmdline[1]] = [ln.strip() for ln in r_.stderr.splitlines() if ln.startswith("[lac_cli")]
out.write(json.dumps({"config": "C2 milestones (ms since process start)", **bd}) + "\n"); out.flush()
print(bd, file=sys.stderr)

# BASELINE config 3 (30 min 24/192 mono, 1.04 GB) as FILES through the CLI: above the reference's 1 GiB decoded-PCM
# cap, so --allow-large; the .lac must have the SHA-256 the unmodified reference's library encode produced
# (tests/golden/golden_large.json); the reference CLI itself refuses the input
import hashlib
g3 = json.load(open("tests/golden/golden_large.json"))["C3_full_1800s_24_192k_mono"]
_, _, pk = H.synth_range(g3["seed"], 0, g3["frames"], 24, 1, reset_log2=0, planes=False, want_packed=True)
wav3 = f"{tmp}/c3.wav"; write_wav(wav3, pk, 1, g3["rate"], 24)
rec = {"config": "C3 1800 s 24/192 mono through lac_cli --allow-large (files)", "pcm_mb": pk.size / 1e6, "host_threads": threads}
te, re_ = run(CLI, "encode", wav3, f"{tmp}/c3.lac", "--allow-large"); assert re_.returncode == 0, re_.stderr
td, rd = run(CLI, "decode", f"{tmp}/c3.lac", f"{tmp}/c3.back.wav", "--allow-large"); assert rd.returncode == 0, rd.stderr
h = hashlib.sha256()
with open(f"{tmp}/c3.lac", "rb") as f:
    for chunk in iter(lambda: f.read(1 << 26), b""):
        h.update(chunk)
rec.update(gpu_encode_s=round(te, 3), gpu_decode_s=round(td, 3), gpu_lac_bytes=os.path.getsize(f"{tmp}/c3.lac"),
           lac_sha256_matches_reference_library=h.hexdigest() == g3["sha256"],
           roundtrip_identical=open(f"{tmp}/c3.back.wav", "rb").read() == open(wav3, "rb").read())
if os.path.exists(REF):
    tr, rr_ = run(REF, "encode", wav3, f"{tmp}/c3.ref.lac")
    rec["ref_cli"] = "rejected: " + (rr_.stderr.strip() or rr_.stdout.strip())[:80] if rr_.returncode else f"{tr:.3f} s"
assert rec["lac_sha256_matches_reference_library"] and rec["roundtrip_identical"], rec
for f_ in ("c3.wav", "c3.lac", "c3.back.wav"):
    os.unlink(f"{tmp}/{f_}")
del pk
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
t0 = time.perf_counter()
for i in range(N):
    assert subprocess.run([CLI, "encode", f"{tmp}/s{i}.wav", f"{tmp}/s{i}.lac3", f"--threads={threads}"], capture_output=True, env=SENV).returncode == 0
    assert subprocess.run([CLI, "decode", f"{tmp}/s{i}.lac3", f"{tmp}/s{i}.out3.wav", f"--threads={threads}"], capture_output=True, env=SENV).returncode == 0
rec["gpu_per_process_with_server_s"] = round(time.perf_counter() - t0, 3)
assert open(f"{tmp}/s5.out3.wav", "rb").read() == open(f"{tmp}/s5.wav", "rb").read()
subprocess.run([CLI, "shutdown"], env=SENV, capture_output=True)
server.wait(timeout=30)
assert open(f"{tmp}/s3.out.wav", "rb").read() == open(f"{tmp}/s3.wav", "rb").read()
out.write(json.dumps(rec) + "\n"); out.flush()
print(rec, file=sys.stderr)
