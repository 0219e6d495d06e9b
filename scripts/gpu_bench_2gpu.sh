#!/bin/bash
# round 2, GPU run 6 (2 GPUs): library multi-GPU test, bench.py under torchrun with its extras (strong split, library_dp, other configs)
cd "$(dirname "$0")/.."
O=gpurun_out/run6; mkdir -p $O
nvidia-smi -L | tee $O/gpus.txt

(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err) 2>&1 | tail -3
tail -c 3000 $O/bench_2gpu.json; grep "bench +" $O/bench_2gpu.err


