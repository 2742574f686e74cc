"""Extracts the DRAM traffic of the k_analyze launch from an `ncu --set full` report of
`python bench.py` (default workload) and writes profiles/r2_roofline_traffic.json, which
bench.py reports as roofline.traffic for as long as the kernel sources (SHA-256 of csrc/) are unchanged.
usage: ncu_traffic.py gpurun_out/prof_an_full.ncu-rep frames_per_gpu"""
import csv, io, json, subprocess, sys
sys.path.insert(0, ".")
from bench import kernel_src_sha
rep, frames = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "k_analyze<1024" not in d["Kernel Name"]:
        continue
    u = dict(zip(hdr, units))
    def to_bytes(key):
        v = float(d[key]); unit = u[key].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    rec = {"kernel": "k_analyze<1024,16,0>", "frames_per_gpu": frames,
           "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
           "duration_ms_under_ncu": float(d["gpu__time_duration.sum"]), "source": rep.split("/")[-1]}
    rec["dram_bytes_per_launch"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    if best is None or rec["duration_ms_under_ncu"] > best["duration_ms_under_ncu"]:
        best = rec  # the full-size launch (the e2e leg launches smaller ones)
best["kernel_src_sha256"] = kernel_src_sha()
json.dump(best, open("profiles/r2_roofline_traffic.json", "w"), indent=1)
print(best)
