#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ARCH=${ARCH:-small}; B=${B:-256}
python -c "import sys; sys.path.insert(0,'tests'); import util; util.model_root('$ARCH')"
# per-launch durations of one full pass (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "enc/" --csv --log-file gpurun_out/launches_enc_${ARCH}${B}.csv python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "dec/" --csv --log-file gpurun_out/launches_dec_${ARCH}${B}.csv python scripts/profile_kernels.py $ARCH $B 10 > gpurun_out/prof2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "mel/" --csv --log-file gpurun_out/launches_mel_${ARCH}${B}.csv python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof3.log 2>&1
# full sets: first encoder layer (LN, QKV, attention, out-proj, LN, fc1, fc2) and one decoder layer of a late step
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "enc/" -s 2 -c 7 -o gpurun_out/prof_enc_${ARCH}${B} -f python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof4.log 2>&1
KPS=$((1 + 12*11 + 2 + 2)); [ "$ARCH" = "base" ] && KPS=$((1 + 6*11 + 2 + 2))
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec/" -s $((KPS*8 + 1)) -c 11 -o gpurun_out/prof_dec_${ARCH}${B} -f python scripts/profile_kernels.py $ARCH $B 10 > gpurun_out/prof5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "mel/" -c 3 -o gpurun_out/prof_mel_${ARCH}${B} -f python scripts/profile_kernels.py $ARCH $B 6 > gpurun_out/prof6.log 2>&1
tail -2 gpurun_out/prof*.log
ls -la gpurun_out
