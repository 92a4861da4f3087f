#!/usr/bin/env python
"""Instruction mix (thread instructions per pixel) of the kernels matching a regex in an .ncu-rep.
usage: python tools/ncu_kernel_mix.py REPORT.ncu-rep KERNEL_REGEX PIXELS [launch_index]"""
import csv, io, re, subprocess, sys
from collections import Counter
rep, rx, px = sys.argv[1], sys.argv[2], float(sys.argv[3])
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
# several kernels may still be printed: split on the 'Kernel Name' header rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
for b in blocks[:1]:
    hdr, data = b["rows"][0], b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    mix, tot = Counter(), 0.0
    for r in data:
        try: t = float(r[ix["Thread Instructions Executed"]])
        except (ValueError, IndexError): continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if m: mix[m.group(1)] += t; tot += t
    print(b["name"][:100]); print(f"  {tot / px:.1f} thread instructions per pixel")
    print("  " + " ".join(f"{k} {v / px:.1f}" for k, v in mix.most_common(24)))
