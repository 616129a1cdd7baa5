#!/usr/bin/env bash
# 2-GPU call: tensor parallel + expert parallel correctness, TP sampling, and the bench flow that embeds the
# TP line at N = 8 (exercised here with a 2-GPU stand-in).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== tensor / expert parallel tests (2 GPUs)"
( time timeout 1200 python -m pytest tests/test_tp_gpu.py tests/test_checkpoint_gpu.py -m gpu -q -k "tp or tensor_parallel or expert" 2>&1 | grep -E "^E  |passed|failed|Error|skipped" | cut -c1-500 | head -30 ) 2>&1
echo "=================== bench flow: DP line at N=2 + embedded TP line (stand-in: llama-3-8b tp=2)"
( time ACP_BENCH_TP_WORLD=2 ACP_BENCH_TP_ARGS="--model llama-3-8b --tp 2" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print({k: j.get(k) for k in ('value','n_gpus','ms_per_step','decode_tokens_per_s')})
print('e2e', j['e2e']); print('roofline', j['roofline']['frac'], 'prefill', j['roofline_prefill']['frac'])
print('tp8_70b', json.dumps(j.get('tp8_70b'))[:900])
" ) 2>&1
