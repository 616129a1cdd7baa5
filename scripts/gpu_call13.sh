#!/usr/bin/env bash
# 8-GPU call (second run, after the cta_group::2 default, the split-K workspace fix and the KV sizing fix of config 4): what the driver's scaling run executes at N = 8 (DP line + embedded 70B TP=8 line), BASELINE
# config 4 (Mixtral EP=8, Poisson arrivals, delegation chains), and the TP=4/8 + EP=8 correctness tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== torchrun bench.py --gpus 8 (DP x8 + tp8_70b)"
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; tail -1 gpurun_out/bench_r2_n8.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print({k: j.get(k) for k in ('value','n_gpus','ms_per_step','decode_tokens_per_s','p50_decode_step_ms')})
print('e2e', j['e2e']['value']); print('roofline', j['roofline']['frac'], 'prefill', j['roofline_prefill']['frac'])
print('tp8_70b', json.dumps(j.get('tp8_70b'))[:1200])
" ) 2>&1
tail -3 gpurun_out/bench_r2_n8.err | cut -c1-300
echo "=================== bench.py --config 4 (Mixtral-8x7B EP=8)"
( time timeout 1200 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4.json 2> gpurun_out/bench_r2_config4.err; tail -1 gpurun_out/bench_r2_config4.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print({k: j.get(k) for k in ('value','n_gpus','ms_per_step','decode_tokens_per_s','p50_decode_step_ms','p99_decode_step_ms','task_ms_p50','task_ms_p99','gpu_launches')})
print('e2e', j['e2e']['value']); print('roofline', j['roofline']['frac'], 'prefill', j['roofline_prefill']['frac'])
print('config', json.dumps(j['config'])[:900])
" ) 2>&1
tail -3 gpurun_out/bench_r2_config4.err | cut -c1-300
echo "=================== TP = 4 / 8 and EP = 8 correctness"
( time timeout 1500 python -m pytest tests/test_tp_gpu.py -m gpu -q -k "tp_matches and (4-p2p or 8-p2p)" 2>&1 | grep -E "^E  |passed|failed|Error|skipped" | cut -c1-400 | head -20 ) 2>&1
