#!/bin/bash
# round 2, GPU run 13: ncu --set full with source counters for the decoder-step kernels at a mid-size batch (base, B=64)
cd "$(dirname "$0")/.."
O=gpurun_out/run13; mkdir -p $O
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('base')" > /dev/null
# one decoder step of base B=64 = ~140 kernels; skip 3 steps, capture one layer of both streams (+ a bit)
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec/" -s 430 -c 26 -o /tmp/dec_base64 -f python scripts/profile_kernels.py base 64 8 > $O/prof.log 2>&1
tail -2 $O/prof.log
ncu -i /tmp/dec_base64.ncu-rep --page raw --csv > $O/dec_base64_raw.csv 2>/dev/null
ncu -i /tmp/dec_base64.ncu-rep --page details --csv > $O/dec_base64_details.csv 2>/dev/null
ncu -i /tmp/dec_base64.ncu-rep --page source --csv --kernel-name regex:gemm_tcgen05_kernel > $O/dec_base64_source_gemm.csv 2>/dev/null
ls -la /tmp/dec_base64.ncu-rep $O
sz=$(stat -c %s /tmp/dec_base64.ncu-rep); if [ $sz -lt 40000000 ]; then cp /tmp/dec_base64.ncu-rep $O/; fi
