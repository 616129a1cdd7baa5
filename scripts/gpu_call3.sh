#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== attention unit tests"
( time timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -8 ) 2>&1
for T in 512 1024 4096; do python scripts/attn_probe.py $T 1 32 8 5; done
python scripts/attn_probe.py 4096 1 8 1 5
echo "=================== ncu: tcgen05 prefill attention, T = 4096"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_prefill_tc -c 1 -f -o gpurun_out/r2_attn_prefill_tc python scripts/attn_probe.py 4096 1 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
echo "=================== sampling + engine + full-depth parity"
( time timeout 1500 python -m pytest tests/test_sampling_gpu.py tests/test_fulldepth_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q -s 2>&1 | grep -E "^E  |passed|failed|Error|full-depth|max \|logit" | cut -c1-700 | head -60 ) 2>&1
echo "=================== bench.py (default = config 2)"
( time timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 ) 2>&1
