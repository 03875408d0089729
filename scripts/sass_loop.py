#!/usr/bin/env python
"""Static look at a kernel's hot loop: dump the SASS of one kernel from libocd_b200.so (or an object), find the
backward branches and print the instruction mix of each loop body.
    scripts/sass_loop.py <file> <kernel-name-substring>"""
import collections, re, subprocess, sys

f, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    if pat not in dem:
        continue
    ins = []
    for line in blk.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(dem, "total instructions:", len(ins))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`\(\.L_x_\d+\)|BRA(?:\.\w+)*\s+.*0x([0-9a-f]+)", t)
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t)
            if m2 and int(m2.group(1), 16) in addr and int(m2.group(1), 16) < a:
                j = addr[int(m2.group(1), 16)]
                body = [x[1] for x in ins[j:i + 1]]
                if len(body) < 40:
                    continue
                mix = collections.Counter()
                for b in body:
                    b = re.sub(r"^@!?U?P\d+\s+", "", b)
                    mix[b.split()[0].split(".")[0]] += 1
                print(f"  loop {ins[j][0]:#x}..{a:#x}: {len(body)} instr:", dict(mix.most_common()))
