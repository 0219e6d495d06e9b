#!/bin/bash
# ncu evidence for profiles/ (round 2): launch list of the bench command's timed region, full sets of the log-mel kernel and of
# one decoder layer at a mid-size batch.  Reports stay on the box (64 MiB limit of gpurun_out/); summaries / CSV pages come back.
cd "$(dirname "$0")/.."
O=gpurun_out/profile; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('small'); util.model_root('base')" > /dev/null
timeout 840 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/launches_bench_small256.csv python bench.py --steps 1 --warmup 3 --new-tokens 4 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
python scripts/agg_launches.py $O/launches_bench_small256.csv | tee $O/launches_bench_small256.txt | head -5
timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "mel/" -c 2 -o /tmp/prof_mel -f python scripts/profile_kernels.py small 256 6 > $O/prof_mel.log 2>&1
python scripts/ncu_summary.py /tmp/prof_mel.ncu-rep | tee $O/ncu_mel_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec/" -s 430 -c 26 -o /tmp/dec_base64 -f python scripts/profile_kernels.py base 64 8 > $O/prof_dec.log 2>&1
python scripts/ncu_summary.py /tmp/dec_base64.ncu-rep | tee $O/ncu_decode_base64_summary.txt | tail -3
ncu -i /tmp/dec_base64.ncu-rep --page source --csv --kernel-name regex:gemm_tcgen05_kernel > $O/dec_base64_source_gemm.csv 2>/dev/null
