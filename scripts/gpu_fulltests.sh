#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== full gpu suite" | tee gpurun_out/fulltests.log
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/fulltests.log
echo "=== smoke" | tee -a gpurun_out/fulltests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a gpurun_out/fulltests.log
