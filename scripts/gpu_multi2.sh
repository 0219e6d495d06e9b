#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi2.log
timeout 600 python -m pytest tests/test_gpu_api.py tests/test_gpu_multi.py -x -q -p no:cacheprovider 2>&1 | tail -6 | tee -a gpurun_out/multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2 GPUs', round(d['value'],1), round(d['e2e']['value'],1), d['stages'])" | tee -a gpurun_out/multi2.log
tail -2 gpurun_out/bench_2gpu.err
