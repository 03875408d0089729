import os, sys
sys.argv = sys.argv[:1]
exec(open('/root/repo/scripts/tuning/form_sweep.py').read().split("for H, C, lr in")[0])
for H, C, lr, sizes in ((5, 2, 0.1, (98304, 131072, 196608, 262144, 524288)), (15, 2, 0.03, (40960, 49152, 131072)), (32, 2, 0.003, (16384, 65536, 262144)), (50, 2, 0.003, (131072, 262144)), (50, 6, 0.003, (65536, 262144)), (15, 6, 0.03, (32768, 65536)), (5, 6, 0.1, (65536, 262144, 1048576))):
    for B in sizes:
        row = {f or "auto": round(run(H, C, B, f, lr), 4) for f in ("throughput", "latency")}
        print(H, C, B, row, flush=True)
