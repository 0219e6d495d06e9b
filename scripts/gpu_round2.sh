#!/bin/bash
# round check: full gpu suite, smoke, default bench, other configs, ncu traffic of the streaming cross-attention kernel
cd "$(dirname "$0")/.."
bash scripts/gpu_round.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,3) for k,v in d['stages'].items()}, round(d['roofline']['frac'],3))" $1 | tee -a gpurun_out/fulltests.log; }
timeout 900 python bench.py --arch turbo --batch 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_turbo128_r2.json 2> gpurun_out/bench_turbo128_r2.err; summ gpurun_out/bench_turbo128_r2.json
timeout 600 python bench.py --arch base --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base64_r2.json 2> gpurun_out/bench_base64_r2.err; summ gpurun_out/bench_base64_r2.json
timeout 600 python bench.py --arch base --batch 256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base256_r2.json 2> gpurun_out/bench_base256_r2.err; summ gpurun_out/bench_base256_r2.json
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dec/" -k regex:cross_attention_stream -s 30 -c 2 -o gpurun_out/prof_xattn_stream -f python scripts/profile_kernels.py small 256 6 > gpurun_out/prof_xattn_stream.log 2>&1; tail -1 gpurun_out/prof_xattn_stream.log
