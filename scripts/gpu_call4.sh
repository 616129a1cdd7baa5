#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== attention unit + sampling + engine + fullsize parity"
( time timeout 1500 python -m pytest tests/test_attention_gpu.py tests/test_sampling_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py tests/test_llmclient_gpu.py tests/test_checkpoint_gpu.py tests/test_gemm_gpu.py -m gpu -q -k "not tiny-moe and not mixtral" 2>&1 | grep -E "^E  |passed|failed|Error" | cut -c1-500 | head -40 ) 2>&1
echo "=================== mixture of experts (own process: a trap here must not poison the suites above)"
( time timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "tiny-moe" 2>&1 | grep -E "^E  |passed|failed|Error|acp_infer" | cut -c1-500 | head -30 ) 2>&1
( time timeout 600 python -m pytest tests/test_checkpoint_gpu.py -m gpu -q -k mixtral 2>&1 | tail -3 ) 2>&1
echo "--- bench --config 4 flow at tiny scale (dev): Poisson arrivals + delegation chains on the MoE engine"
timeout 600 python bench.py --config 4 --model tiny-moe --tp 1 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-900
echo "=================== compute-sanitizer memcheck over the tiny-model workload"
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_probe.py 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|ok|Error|=========     at" | head -30 ) 2>&1
echo "=================== timings (decode attention K/V prefetch before the grid-dependency wait is in this build)"
REPS=3 timeout 300 python scripts/engine_probe.py llama-3-8b 64 512 64 2>&1 | grep '"rep": [12]' | cut -c1-420
REPS=2 timeout 300 python scripts/engine_probe.py llama-3-8b 256 512 24 2>&1 | grep '"rep": 1' | cut -c1-420
REPS=2 timeout 400 python scripts/config2_probe.py 2>&1 | tail -1 | cut -c1-420
echo "=================== launch list of one config-1 bench step (ncu, gpu__time_duration)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv python scripts/engine_probe.py llama-3-8b 64 512 3 > gpurun_out/launches_r2.log 2>&1; tail -1 gpurun_out/launches_r2.log | cut -c1-200
echo "=================== open-loop arrivals: prefill-first vs latency-bound scheduling"
timeout 600 python scripts/arrivals_probe.py 30 15 0 1 2>&1 | grep decode_interleave | cut -c1-700
