#!/usr/bin/env bash
# Round 2, GPU call 1 (HISTORICAL: the four switches it A/B-ed were measured with it and then REMOVED from the
# tree — profiles/r2_ab_switches.md; the env vars below are no-ops now).  Kept as the provenance of those numbers.
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
