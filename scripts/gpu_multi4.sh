#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N" | tee gpurun_out/multi4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -n 1 gpurun_out/bench_${N}gpu.json | tee -a gpurun_out/multi4.log | cut -c1-300
timeout 600 python scripts/longform.py small 1.0 2> gpurun_out/longform.err | tail -n 1 | tee -a gpurun_out/multi4.log
tail -3 gpurun_out/longform.err
