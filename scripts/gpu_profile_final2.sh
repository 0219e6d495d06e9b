#!/bin/bash
# Round-1 final evidence for profiles/: launch list of the (shortened) bench command + full sets of the hot kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')" > /dev/null
timeout 840 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/launches_bench_small256_v2.csv python bench.py --steps 1 --warmup 3 --new-tokens 4 --no-cpu-baseline > gpurun_out/ncu_bench_final2.log 2>&1
tail -1 gpurun_out/ncu_bench_final2.log | cut -c1-200
# full sets: roofline kernel over the whole batch, one decoder layer of both micro-batches, first encoder layer
timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "xattn/" -c 2 -o gpurun_out/prof_xattn_b256_final2 -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof7.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "enc/" -s 2 -c 7 -o gpurun_out/prof_enc_small256_final2 -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof8.log 2>&1
tail -1 gpurun_out/prof7.log gpurun_out/prof8.log
