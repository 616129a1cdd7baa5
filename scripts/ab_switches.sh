#!/usr/bin/env bash
# Round 2, first GPU call: (1) the new parity tests at BASELINE shapes on the default kernels,
# (2) validate-or-delete A/B of the four experimental switches left by round 1 (parity first, then
# timing, all on ONE box so the numbers share a chip).  Output: gpurun_out/ab_switches.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== parity at BASELINE shapes (default kernels)"
( time timeout 1500 python -m pytest tests/test_fulldepth_gpu.py tests/test_engine_gpu.py -m gpu -q -s -x 2>&1 | grep -v "^$" | tail -25 ) 2>&1
for sw in X=0 ACP_ATTN_PT_PREFETCH=1 ACP_ATTN_HALF=1 ACP_ATTN_PREFILL_3CTA=1 ACP_GEMM_PERSISTENT_DECODE=1; do
  echo "=================== $sw"
  if [ "$sw" != "X=0" ]; then
    ( time env "$sw" timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "not llama-3-8b-l2" 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys"
  fi
  env "$sw" REPS=3 timeout 300 python scripts/engine_probe.py llama-3-8b 64 512 64 2>&1 | grep '"rep": [12]' | cut -c1-420
done
echo "=================== B = 256 decode (wide GEMM tiles, attention)"
for sw in X=0 ACP_ATTN_HALF=1 ACP_ATTN_PT_PREFETCH=1; do
  echo "--- $sw"; env "$sw" REPS=2 timeout 300 python scripts/engine_probe.py llama-3-8b 256 512 24 2>&1 | grep '"rep": 1' | cut -c1-420
done
echo "=================== config 2 (B = 512, 128..4096-token windows)"
for sw in X=0 ACP_GEMM_PERSISTENT_DECODE=1 ACP_ATTN_PREFILL_3CTA=1 ACP_ATTN_HALF=1; do
  echo "--- $sw"; env "$sw" REPS=2 timeout 400 python scripts/config2_probe.py 2>&1 | tail -1 | cut -c1-420
done
echo "=================== per-kernel profile at config 2 (default)"
ACP_PROFILE=1 REPS=1 timeout 400 python scripts/config2_probe.py 2>&1 | tail -40
echo "=================== tests written without a GPU"
ACP_UNVALIDATED_TESTS=1 timeout 300 python -m pytest tests/test_checkpoint_gpu.py -m gpu -x -q 2>&1 | tail -3
