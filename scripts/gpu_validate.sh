#!/bin/bash
# Full validation on one GPU: pytest -m gpu, smoke(), default bench (what the driver runs at round end), reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out/validate; mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -6) 2>&1 | tee $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
(time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err) 2>&1 | tail -3
tail -c 600 $O/bench_default.json; echo; grep "bench +" $O/bench_default.err
