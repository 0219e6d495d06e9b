#!/bin/bash
# Evidence for profiles/: launch list of the bench command (timed region only, shortened decode) + full sets of the hot kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small')"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file gpurun_out/launches_bench_small256.csv python bench.py --steps 1 --warmup 3 --new-tokens 28 --no-cpu-baseline > gpurun_out/ncu_bench_final.log 2>&1
tail -2 gpurun_out/ncu_bench_final.log | cut -c1-300
# full set: one decoder layer of a late step (B=256: kernels of both micro-batches) and one encoder layer
KPS=$((2 + 12*22 + 2*3 + 1))
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec/" -s $((KPS*8 + 2)) -c 22 -o gpurun_out/prof_dec_small256_final -f python scripts/profile_kernels.py small 256 10 > gpurun_out/prof5.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "enc/" -s 2 -c 7 -o gpurun_out/prof_enc_small256_final -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "mel/" -c 3 -o gpurun_out/prof_mel_small256_final -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof6.log 2>&1
tail -1 gpurun_out/prof4.log gpurun_out/prof5.log gpurun_out/prof6.log
