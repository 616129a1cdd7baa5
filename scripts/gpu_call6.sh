#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== sampling + llmclient + MoE (own processes)"
( timeout 900 python -m pytest tests/test_sampling_gpu.py tests/test_llmclient_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-400 | head -20 ) 2>&1
( timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "tiny-moe" 2>&1 | grep -E "^E  |passed|failed|Error|acp_infer" | cut -c1-400 | head -20 ) 2>&1
echo "=================== bench --config 4 flow at tiny scale"
timeout 600 python bench.py --config 4 --model tiny-moe --tp 1 --steps 1 --warmup 1 2>&1 | tail -12 | cut -c1-1200
echo "=================== config 2 per-kernel decode profile (N > 256 split change in this build)"
ACP_PROFILE=1 REPS=1 timeout 400 python scripts/config2_probe.py 2>&1 | grep -E "^\{|decode profile|gemm_|attn_decode|add_rmsnorm|swiglu" | cut -c1-300 | head -16
REPS=2 timeout 400 python scripts/config2_probe.py 2>&1 | tail -1 | cut -c1-420
echo "=================== ncu --set full: DRAM traffic of the decode-step kernels at config 2 (for roofline.traffic)"
REPS=1 timeout 900 ncu --set full --clock-control none -k regex:"attn_decode_kernel|attn_merge_kernel|swiglu_kernel|add_rmsnorm_kernel|gemm_wx_kernel|gemm_wx_persistent" --launch-skip 21000 -c 24 -f -o gpurun_out/r2_config2_decode python scripts/config2_probe.py > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log | cut -c1-200
echo "=================== mixtral-8x7b-l2 (real width, 2 layers) decode / prefill timing"
REPS=2 timeout 300 python scripts/engine_probe.py mixtral-8x7b-l2 64 512 32 2>&1 | grep '"rep": 1' | cut -c1-420
