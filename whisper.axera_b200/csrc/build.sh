#!/bin/bash
# Builds libax_whisper.so (sm_100a only), whisper_cli and whisper_srv in-tree. nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
OUT=${OUT:-..}
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xcompiler -Wall,-Wno-unused-function"
OBJ=${OBJ:-obj}   # OBJ / OUT elsewhere = a scratch build that leaves the in-tree library untouched (e.g. while a gpurun call is queued)
mkdir -p $OBJ
pids=()
for f in logmel.cu gemm_tcgen05.cu gemm2cta_tcgen05.cu encoder_ops.cu attention_tcgen05.cu decode_ops.cu engine.cu model_abi.cu; do
  if [ ! -f $OBJ/${f%.cu}.o ] || [ $f -nt $OBJ/${f%.cu}.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' -o -name '*.inc' \) -newer $OBJ/${f%.cu}.o)" ] || [ -n "$(find ../../include -name '*.h' -newer $OBJ/${f%.cu}.o)" ]; then
    $NVCC $FLAGS -c $f -o $OBJ/${f%.cu}.o &
    pids+=($!)
  fi
done
for f in ax_whisper_api.cpp host_utils.cpp; do
  $NVCC $FLAGS -x cu -c $f -o $OBJ/${f%.cpp}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT/libax_whisper.so $OBJ/*.o -lcudart_static -ldl -lpthread -lrt -Xcompiler -static-libstdc++,-static-libgcc -Xlinker --exclude-libs=ALL
g++ -O2 -std=c++17 whisper_cli.cpp host_utils.cpp -o $OUT/whisper_cli -L$OUT -lax_whisper -Wl,-rpath,'$ORIGIN' -static-libstdc++ -static-libgcc
g++ -O2 -std=c++17 -pthread whisper_srv.cpp -o $OUT/whisper_srv -L$OUT -lax_whisper -Wl,-rpath,'$ORIGIN' -static-libstdc++ -static-libgcc
echo "built $OUT/libax_whisper.so $OUT/whisper_cli $OUT/whisper_srv"
