#!/bin/bash
# round 2, GPU run 15: single-launch log-mel (tests + timing), ncu launch list of the bench command's timed region, full set of the mel kernel
cd "$(dirname "$0")/.."
O=gpurun_out/run15; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_mel.py tests/test_gpu_encoder.py tests/test_gpu_turbo.py tests/test_gpu_api.py tests/test_gpu_small256.py -m gpu -q -x 2>&1 | grep -v Warning | tail -4 | tee $O/tests.log
python bench.py --no-extras --no-cpu-baseline --steps 3 --warmup 3 2>>$O/err.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('small256 value', round(d['value']), 'mel_ms', d['stages']['mel_ms'], 'mel_frac', d['stages']['mel_frac_hbm'], 'enc_ms', d['stages']['encoder_ms'], 'dec_ms', d['stages']['decode_ms'], 'clk', d['clocks']['sm_mhz'])" | tee $O/mel.txt
timeout 840 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" --csv --log-file $O/launches_bench_small256_r02.csv python bench.py --steps 1 --warmup 3 --new-tokens 4 --no-cpu-baseline --no-extras > $O/ncu_bench.log 2>&1
tail -1 $O/ncu_bench.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "mel/" -c 2 -o /tmp/prof_mel -f python scripts/profile_kernels.py small 256 6 > $O/prof_mel.log 2>&1
python scripts/ncu_summary.py /tmp/prof_mel.ncu-rep | tee $O/ncu_mel_summary.txt
