#!/bin/bash
# round 2: bench.py under torchrun on 8 GPUs (what the driver's SCALE step runs at N = 8), with extras
cd "$(dirname "$0")/.."
O=gpurun_out/run8gpu; mkdir -p $O
nvidia-smi -L | wc -l | tee $O/gpus.txt
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err) 2>&1 | tail -3
tail -c 2500 $O/bench_8gpu.json; tail -5 $O/bench_8gpu.err
