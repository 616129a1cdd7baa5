#!/usr/bin/env bash
# Round 2, last 8-GPU call: BASELINE config 4 (Mixtral-8x7B EP=8, Poisson arrivals, delegation chains) with the KV
# sizing fix (every Task must end FinalAnswer) and the shards on the validated kernel set.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout -k 10 240 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4.json 2> gpurun_out/bench_r2_config4.err; tail -c 600 gpurun_out/bench_r2_config4.err; tail -1 gpurun_out/bench_r2_config4.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print({k: j.get(k) for k in ('value','n_gpus','ms_per_step','decode_tokens_per_s','p50_decode_step_ms','p99_decode_step_ms','task_ms_p50','task_ms_p99','gpu_launches')})
print('e2e', j['e2e']['value']); print('roofline', j['roofline']['frac'], 'prefill', j['roofline_prefill']['frac'])
print('config', json.dumps(j['config'])[:900])
" ) 2>&1
