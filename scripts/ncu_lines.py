#!/usr/bin/env python
"""Stall samples of one ncu capture aggregated per CUDA source line:  python scripts/ncu_lines.py REP [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.Counter(); inst = collections.Counter(); text = {}
fname = None; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    # the combined view has duplicate 'Source' keys: first is CUDA text
    key = (fname, int(r[0]))
    text[key] = r[1].strip()
    s = r[hdr.index("# Samples")]; e = r[hdr.index("Instructions Executed")]
    if s.isdigit(): agg[key] += int(s)
    if e.isdigit(): inst[key] += int(e)
tot = sum(agg.values()) or 1
print("total samples", tot, " total warp instr", sum(inst.values()))
for key, n in agg.most_common(top):
    print("%5.1f%% %7d smp %9d inst  %s:%d  %s" % (100.0 * n / tot, n, inst[key], key[0], key[1], text[key][:110]))
