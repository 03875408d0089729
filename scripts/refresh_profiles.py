#!/usr/bin/env python
"""Turn the artefacts one GPU session left in gpurun_out/ (scripts/gpu_profile.sh <tag>, scripts/sweep.py,
bench.py, scripts/gpu_profile_sweep.sh <tag>) into the tracked summaries under profiles/.

    scripts/refresh_profiles.py <tag> <sweep.json> <bench.log> <bench_ref.log> [bench_n2.log bench_n4.log bench_n8.log]
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag, sweep, bench, ref = sys.argv[1:5]
scaling_logs = sys.argv[5:]
R = tag.split("_")[0] if tag.startswith("r0") else "r01"      # round prefix of the tracked files (r01, r02, ...)

shutil.copy(G / f"launches_{tag}.csv", P / f"{R}_launches.csv")
rows = [r for r in csv.reader(open(G / f"launches_{tag}.csv")) if r and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
with open(P / f"{R}_launch_shares.txt", "w") as f:
    f.write("# python bench.py --steps 5 --warmup 3 --no-extras under `ncu --metrics gpu__time_duration.sum "
            "--clock-control none`\n# launches  total_ms  share  kernel\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[0]:4d} {v[1]:10.3f} {100 * v[1] / tot:6.2f}%  {k[:100]}\n")


def summ(rep, out, title):
    o = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep), "--stalls"],
                       capture_output=True, text=True).stdout
    open(out, "w").write(f"# {title}\n# ncu --set full --clock-control none --import-source on; reduced by "
                         f"scripts/ncu_summary.py\n" + o)
    return o


main = summ(G / f"prof_solve_{tag}.ncu-rep", P / f"{R}_k_solve_ncu_summary.txt",
            "k_solve<5,1,3,fast,wide> at the bench shape (B = 2^20, H=5, 2 cars)")
for name, title in (("h5c6", "k_solve<5,5,3,fast,wide+step fence>: sweep point H=5, 6 cars, B = 2^20"),
                    ("h15c2", "k_solve<15,1,3,fast,wide> (Q kernel: saved step data through shared memory): sweep point H=15, 2 cars, B = 2^20"),
                    ("h15c6", "k_solve<15,5,3,fast,wide> (constant-segment-count adjoint): sweep point H=15, 6 cars, B = 2^20"),
                    ("h50c2", "k_solve<50,1,3,fast,wide> (constant-segment-count adjoint): sweep point H=50, 2 cars, B = 2^20")):
    rep = G / f"prof_{tag}_{name}.ncu-rep"
    if rep.exists():
        summ(rep, P / f"{R}_k_solve_{name}_ncu_summary.txt", title)
shutil.copy(G / sweep, P / f"{R}_sweep.json")

rd = wr = 0.0
for line in main.splitlines():
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if line.startswith(key):
            unit = line.split("[")[1].split("]")[0]
            val = float(line.split("['")[1].split("'")[0])
            val *= {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[unit]
            if "read" in key:
                rd = val
            else:
                wr = val
t = json.load(open(P / "traffic.json"))
t["k_solve_bytes_per_launch"] = int(rd + wr)
json.dump(t, open(P / "traffic.json", "w"), indent=1)


def last_json(path):
    return json.loads(open(G / path).read().strip().splitlines()[-1])


d = {"note": f"{R} kernels; full 1-GPU line, the --impl reference arm from the same box, and the multi-GPU "
             "lines given (torchrun, one rank per GPU)",
     "bench": last_json(bench), "reference_arm": last_json(ref)}
lines = [d["bench"]] + [last_json(x) for x in scaling_logs]
d["scaling"] = [{"n_gpus": x["n_gpus"], "value": x["value"], "ms_per_step": x["ms_per_step"], "e2e": x["e2e"]["value"],
                 "cmaes_ms_per_generation": x["cmaes"]["ms_per_generation"], "clocks": x["clocks"]} for x in lines]
json.dump(d, open(P / f"{R}_bench.json", "w"), indent=1)
print("profiles/ refreshed from tag", tag)
