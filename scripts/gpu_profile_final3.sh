#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')" > /dev/null
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | cut -c1-400
timeout 840 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/launches_bench_small256_v3.csv python bench.py --steps 1 --warmup 3 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench_final3.log 2>&1
tail -n 1 gpurun_out/ncu_bench_final3.log | cut -c1-120
