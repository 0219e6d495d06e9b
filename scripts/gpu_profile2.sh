#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ARCH=${ARCH:-small}; B=${B:-64}
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('$ARCH')"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "enc/" --csv --log-file gpurun_out/launches_enc_${ARCH}${B}_r4.csv python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof1.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "enc/" -s 2 -c 7 -o gpurun_out/prof_enc_${ARCH}${B}_r4 -f python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof4.log 2>&1
tail -2 gpurun_out/prof1.log gpurun_out/prof4.log
