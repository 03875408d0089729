#!/usr/bin/env python
"""Stall-reason samples per hot region of a kernel (from an .ncu-rep's source page, SASS view).
Usage: scripts/sass_stalls.py <rep> lo:hi [lo:hi ...]   (instruction index ranges as printed by sass_hot.py)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
iE = hdr.index("Instructions Executed")
data = rows[2:]
tot_all = sum(int(r[i] or 0) for r in data for i in cols)
for rng in sys.argv[2:]:
    lo, hi = map(int, rng.split(":"))
    sel = data[lo:hi + 1]
    ex = sum(int(r[iE]) for r in sel)
    agg = {hdr[i]: sum(int(r[i] or 0) for r in sel) for i in cols}
    tot = sum(agg.values())
    print(f"[{lo}..{hi}] executed {ex:.4g}, samples {tot} ({tot / tot_all:.3f} of all): " +
          " ".join(f"{k[6:]}:{v / tot:.3f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot >= 0.01))
