#!/usr/bin/env python
"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump: samples and executed instructions per
CUDA source line, top lines first, and per file.   python scripts/ncu_src_summary.py file.csv [ntop]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None
lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "" and hdr:
        try:
            i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed")
            lsb = hdr.index("stall_long_sb"); wt = hdr.index("stall_wait")
            lines.append((int(r[i_s] or 0), int(r[i_e] or 0), cur, int(r[0]), r[1].strip()[:110], int(r[lsb] or 0), int(r[wt] or 0)))
        except ValueError:
            pass
tot = sum(l[0] for l in lines); tote = sum(l[1] for l in lines)
print("total samples", tot, "instructions", tote)
byfile = collections.Counter(); bye = collections.Counter()
for l in lines: byfile[l[2]] += l[0]; bye[l[2]] += l[1]
for f, v in byfile.most_common(): print(f"  {f:28s} samples {v/tot:6.3f}  instr {bye[f]/tote:6.3f}")
for l in sorted(lines, reverse=True)[:ntop]:
    print(f"{l[0]/tot:6.3f} {l[1]/tote:6.3f} lsb={l[5]/max(l[0],1):4.2f} {l[2]}:{l[3]:4d}  {l[4]}")
