#!/usr/bin/env python
"""Hot regions of a kernel from an .ncu-rep (source page, SASS view): contiguous runs of instructions grouped by
their executed count, with the instruction mix of each run.  Usage: scripts/sass_hot.py <rep> [min_share] [--dump]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[iS].strip(), int(r[iE]), int(r[iSamp])) for r in rows[2:] if len(r) > iE]
total = sum(e for _, e, _ in ins)
tsamp = sum(s for _, _, s in ins)
print(f"instructions {len(ins)}, executed {total:.4g} warp-instr, samples {tsamp}")
# group consecutive instructions whose executed count is within 2 % of each other
runs, cur = [], []
for i, (s, e, sm) in enumerate(ins):
    if cur and abs(e - cur[-1][1]) > 0.02 * max(e, cur[-1][1], 1):
        runs.append(cur); cur = []
    cur.append((s, e, sm, i))
if cur: runs.append(cur)
for run in runs:
    ex = sum(e for _, e, _, _ in run)
    if ex < min_share * total: continue
    mix = collections.Counter(s.split()[1].split(".")[0] if s.startswith("@") else s.split()[0].split(".")[0] for s, _, _, _ in run)
    smp = sum(sm for _, _, sm, _ in run)
    print(f"[{run[0][3]:5d}..{run[-1][3]:5d}] n={len(run):4d} exec/instr={run[0][1]:.4g} share={ex/total:.3f} samples={smp/tsamp:.3f} "
          + " ".join(f"{k}:{v}" for k, v in mix.most_common(14)))
    if "--dump" in sys.argv:
        for s, e, sm, i in run: print(f"      {i:5d} {e:10d} {sm:6d}  {s}")
