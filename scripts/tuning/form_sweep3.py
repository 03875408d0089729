"""Round 2: every form of the compile-time H = 15 / 50 kernels (and many cars) over the mid-size batches, ms per launch."""
import os, sys
sys.argv = sys.argv[:1]
exec(open('/root/repo/scripts/tuning/form_sweep.py').read().split("for H, C, lr in")[0])
for H, C, lr, sizes in ((15, 2, 0.02, (8192, 16384, 24576, 32768, 49152, 65536, 98304, 131072, 262144)),
                        (50, 2, 0.0003, (8192, 16384, 32768, 65536, 131072, 262144)),
                        (15, 6, 0.02, (16384, 32768, 65536, 131072, 262144)),
                        (50, 6, 0.0003, (16384, 32768, 65536, 131072)),
                        (15, 3, 0.02, (16384, 65536, 262144))):
    for B in sizes:
        row = {f or "auto": round(run(H, C, B, f, lr), 4) for f in ("", "throughput", "latency", "wide")}
        best = min((v, k) for k, v in row.items() if k != "auto")
        print(H, C, B, row, "best:", best[1], "auto/best = %.3f" % (row["auto"] / best[0]), flush=True)
