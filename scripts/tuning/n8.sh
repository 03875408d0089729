#!/bin/bash
# 8-GPU session (gpurun --gpus 8): topology, end-to-end scaling with and without the per-rank CPU binding, then the bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
nvidia-smi topo -m > $O/topo_n8.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)\|model name" > $O/lscpu_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29511 scripts/tuning/e2e_scale.py 2>/dev/null | tee $O/e2e_scale_n8_bind.log
$TR --master-port 29512 scripts/tuning/e2e_scale.py --no-bind 2>/dev/null | tee $O/e2e_scale_n8_nobind.log
$TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench rc=$?"
python scripts/tuning/e2e_scale.py 2>/dev/null | tee $O/e2e_scale_n1.log
