#!/usr/bin/env bash
# A/B of the experimental switches written after round 1's GPU budget was spent.
# Run on ONE GPU box in ONE call so that the numbers share a chip and its clocks:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/ab_switches.sh > gpurun_out/ab_switches.log 2>&1'
# Each switch must first pass the parity suite bit for bit; only then is its timing meaningful.
set -u
cd "$(dirname "$0")/.."
for sw in X=0 ACP_ATTN_PT_PREFETCH=1 ACP_ATTN_HALF=1 ACP_ATTN_PREFILL_3CTA=1 ACP_GEMM_PERSISTENT_DECODE=1; do
  echo "=================== $sw"
  env "$sw" timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
  env "$sw" REPS=3 timeout 300 python scripts/engine_probe.py llama-3-8b 64 512 64 2>&1 | grep '"rep": [12]' | cut -c1-420
done
echo "=================== B = 256 decode (wide GEMM tiles, attention)"
for sw in X=0 ACP_ATTN_HALF=1 ACP_ATTN_PT_PREFETCH=1; do
  echo "--- $sw"; env "$sw" REPS=2 timeout 300 python scripts/engine_probe.py llama-3-8b 256 512 24 2>&1 | grep '"rep": 1' | cut -c1-420
done
echo "=================== config 2 (B = 512): persistent split-K decode GEMMs"
for sw in X=0 ACP_GEMM_PERSISTENT_DECODE=1; do
  echo "--- $sw"; env "$sw" REPS=2 timeout 400 python scripts/config2_probe.py 2>&1 | tail -1 | cut -c1-420
done
echo "=================== tests written without a GPU"
ACP_UNVALIDATED_TESTS=1 timeout 300 python -m pytest tests/test_checkpoint_gpu.py -m gpu -x -q 2>&1 | tail -3
