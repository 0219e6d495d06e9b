#!/bin/bash
# Run on the GPU box (via gpurun): each test file in its own process under a timeout so a hung kernel cannot take
# the whole call down.  Logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
FILES=${@:-"tests/test_gpu_gemm.py tests/test_gpu_mel.py tests/test_gpu_encoder.py tests/test_gpu_decoder.py"}
rc=0
for f in $FILES; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/bringup.log
  timeout -k 10 ${T:-420} python -m pytest $f -m gpu -q -s --timeout 180 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  r=$?
  tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/bringup.log
  echo "exit $r" | tee -a gpurun_out/bringup.log
  [ $r -ne 0 ] && rc=1
done
exit $rc
