#!/usr/bin/env bash
# Round 2, GPU call 2: tcgen05 prefill attention unit tests -> engine suite on it -> timings; full-depth
# noise probe across w_std; the tiny-g8 failure of call 1 in detail.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== attention unit tests (tcgen05 vs mma.sync vs numpy fp32)"
( time timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 ) 2>&1
echo "=================== tiny-g8 detail, old prefill kernel"
ACP_ATTN_PREFILL_TC=0 timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "tiny-g8" 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-600 | head -40
echo "=================== engine suite (small presets) on the tcgen05 prefill kernel"
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "not llama-3-8b-l2" 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-600 | head -40
echo "=================== config 1 / config 2 timings: tcgen05 vs mma.sync prefill attention"
for sw in ACP_ATTN_PREFILL_TC=1 ACP_ATTN_PREFILL_TC=0; do
  echo "--- $sw"
  env "$sw" REPS=2 timeout 300 python scripts/engine_probe.py llama-3-8b 64 512 64 2>&1 | grep '"rep": 1' | cut -c1-420
  env "$sw" REPS=2 timeout 400 python scripts/config2_probe.py 2>&1 | tail -1 | cut -c1-420
done
echo "=================== full-depth noise probe"
timeout 900 python scripts/fulldepth_probe.py 0.02 0.01 0.005 0.002 2>&1 | grep -E "w_std|Error|error" | head
LAYERS=2 timeout 300 python scripts/fulldepth_probe.py 0.02 2>&1 | grep -E "w_std|Error|error" | head
LAYERS=8 timeout 300 python scripts/fulldepth_probe.py 0.02 2>&1 | grep -E "w_std|Error|error" | head
