#!/bin/bash
# what the driver does at round end, plus the bench: full gpu suite, smoke, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== full gpu suite" | tee gpurun_out/fulltests.log
timeout 1800 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -8 | tee -a gpurun_out/fulltests.log
echo "=== smoke" | tee -a gpurun_out/fulltests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300 | tee -a gpurun_out/fulltests.log
echo "=== bench (default)" | tee -a gpurun_out/fulltests.log
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | tee -a gpurun_out/fulltests.log; tail -3 gpurun_out/bench_default.err
