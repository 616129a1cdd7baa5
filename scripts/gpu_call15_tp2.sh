#!/usr/bin/env bash
# Round 2, 2-GPU call: tensor-parallel shards back on the kernel set of the validated 8-GPU runs (default), then the
# set that hung at TP=8 (cta_group::2 + persistent fp32-plane epilogue) with the bounded waits naming the stuck role.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== TP=2 tests (default kernel set for shards)"
( time timeout -k 10 600 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "2-p2p" 2>&1 | grep -E "^E  |passed|failed|Error|skipped" | cut -c1-400 | head -12 ) 2>&1
echo "=================== config-3 flow, llama-3-8b TP=2, default kernel set"
( time timeout -k 10 600 python bench.py --config 3 --model llama-3-8b --tp 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_tp2_default.json 2> gpurun_out/bench_r2_tp2_default.err; tail -c 400 gpurun_out/bench_r2_tp2_default.err; tail -1 gpurun_out/bench_r2_tp2_default.json | cut -c1-300 ) 2>&1
echo "=================== same, ACP_TP_GEMM_2CTA=1 ACP_TP_GEMM_PERSISTENT_F32=1 (the set that hung at TP=8)"
( time ACP_TP_GEMM_2CTA=1 ACP_TP_GEMM_PERSISTENT_F32=1 timeout -k 10 300 python bench.py --config 3 --model llama-3-8b --tp 2 --steps 3 --warmup 1 > gpurun_out/bench_r2_tp2_fast.json 2> gpurun_out/bench_r2_tp2_fast.err; tail -c 400 gpurun_out/bench_r2_tp2_fast.err; sort gpurun_out/bench_r2_tp2_fast.json | uniq -c | cut -c1-300 | tail -8 ) 2>&1
