#!/usr/bin/env python
"""Per-phase instruction mix and stall reasons of one kernel in an .ncu-rep (source page; phases = BAR.SYNC boundaries).

usage: python tools/ncu_phase_mix.py REPORT.ncu-rep PIXELS_PER_LAUNCH"""
import csv, io, re, subprocess, sys
from collections import Counter

def f(x):
    try: return float(x)
    except ValueError: return 0.0

def main(rep, pixels):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    src, smp, thr = ix["Source"], ix["# Samples"], ix["Thread Instructions Executed"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[src]]
    bounds = [0] + bars + [len(data)]
    tot_s = sum(f(r[smp]) for r in data) or 1.0
    tot_i = sum(f(r[thr]) for r in data)
    print(f"total thread instructions {tot_i:.3e} = {tot_i / pixels:.1f} per pixel; {tot_s:.0f} stall samples")
    for k, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        seg = data[a:b]
        ti = sum(f(r[thr]) for r in seg)
        if ti == 0: continue
        mix = Counter()
        for r in seg:
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[src])
            if m: mix[m.group(1)] += f(r[thr])
        ss = sum(f(r[smp]) for r in seg)
        agg = sorted(((s[6:], sum(f(r[ix[s]]) for r in seg)) for s in stalls), key=lambda kv: -kv[1])[:5]
        print(f"segment {k} (SASS lines {a}-{b}): {ti / pixels:6.1f} instr/px, {100 * ss / tot_s:5.1f} % of stall samples; "
              + " ".join(f"{n}={100 * v / max(ss, 1):.0f}%" for n, v in agg))
        print("    " + " ".join(f"{n} {v / pixels:.1f}" for n, v in mix.most_common(14)))

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
