"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; hdr = None
lines = collections.OrderedDict()
tot_inst = tot_samp = 0
with open(path, newline='') as f:
    for r in csv.reader(f):
        if not r: continue
        if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name': continue
        if r[0] == 'Line No': hdr = r; ii = hdr.index('Instructions Executed'); si = hdr.index('# Samples'); continue
        if r[0] == '': continue  # SASS row
        try:
            inst = float(r[ii]); samp = float(r[si])
        except ValueError:
            continue
        lines[(cur, int(r[0]))] = (inst, samp, r[1])
        tot_inst += inst; tot_samp += samp
print(f"total inst {tot_inst:.3e} samples {tot_samp:.0f}")
byfile = collections.Counter(); byfile_s = collections.Counter()
for (f_, l), (i, s, src) in lines.items(): byfile[f_] += i; byfile_s[f_] += s
for f_, i in byfile.most_common(): print(f"{f_:28s} inst {100*i/tot_inst:5.1f}%  samples {100*byfile_s[f_]/tot_samp:5.1f}%")
print()
for (f_, l), (i, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f_[:18]:18s}:{l:4d} inst {100*i/tot_inst:5.2f}% samp {100*s/tot_samp:5.2f}%  {src.strip()[:110]}")
