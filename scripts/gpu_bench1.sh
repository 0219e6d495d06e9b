#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=300 bash scripts/gpu_bringup.sh tests/test_gpu_decoder.py
echo "=== smoke" | tee -a gpurun_out/bringup.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/bringup.log
echo "=== bench base b64" | tee -a gpurun_out/bringup.log
timeout 600 python bench.py --arch base --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base64.json 2> gpurun_out/bench_base64.err; tail -c 3000 gpurun_out/bench_base64.json | tee -a gpurun_out/bringup.log; tail -5 gpurun_out/bench_base64.err
echo "=== bench small b256" | tee -a gpurun_out/bringup.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_small256.json 2> gpurun_out/bench_small256.err; tail -c 3000 gpurun_out/bench_small256.json | tee -a gpurun_out/bringup.log; tail -5 gpurun_out/bench_small256.err
echo "=== ncu launch list (base b64, 1 step)" | tee -a gpurun_out/bringup.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_base64.csv python bench.py --arch base --batch 64 --steps 1 --warmup 3 --new-tokens 12 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
