#!/usr/bin/env bash
# Round 2, GPU call 10: delegation-chain test, config-4 flow at Mixtral width on one GPU (diagnostics), and the
# `ncu --set full` pair for the 1-CTA / cta_group::2 gate-up GEMM at a 4096-row prefill step.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== delegation chain test (tiny, tiny-moe)"
timeout -k 10 600 python -m pytest tests/test_llmclient_gpu.py -m gpu -q -x -k "delegation" 2>&1 | tail -15
echo "=================== config 4 flow, mixtral-8x7b-l2, one GPU"
timeout -k 10 600 python bench.py --config 4 --model mixtral-8x7b-l2 --tp 1 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4_dev.json 2> gpurun_out/bench_r2_config4_dev.err
echo "rc=$?"; tail -c 1200 gpurun_out/bench_r2_config4_dev.err; tail -c 900 gpurun_out/bench_r2_config4_dev.json
echo "=================== ncu --set full: gate/up GEMM, 4096 rows, 1-CTA then 2-CTA"
ONLY=gate_up ITERS=0 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:persistent -c 2 \
  -o gpurun_out/r2_gemm_gateup_1cta_2cta -f python scripts/gemm2cta_probe.py 4096 > gpurun_out/ncu_gemm2cta.log 2>&1
tail -4 gpurun_out/ncu_gemm2cta.log
