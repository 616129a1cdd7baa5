#!/usr/bin/env bash
# What the driver runs at round end, on one B200: the whole GPU test suite, smoke(), both bench arms.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== pytest -m gpu (whole suite)"
( time timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | grep -E "^E  |passed|failed|Error|full-depth" | cut -c1-500 | head -40 ) 2>&1
echo "=================== smoke()"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=================== bench.py --impl reference"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700 ) 2>&1
echo "=================== bench.py (default)"
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; tail -1 gpurun_out/bench_r2_default.json | cut -c1-400 ) 2>&1
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r2_default.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','decode_tokens_per_s','prefill_tokens_per_s_rank0','p50_decode_step_ms','p99_decode_step_ms','gpu_launches','clocks'):
    print(k, j.get(k))
print('e2e', j['e2e']); print('roofline', {k: j['roofline'][k] for k in ('achieved','frac','traffic','share_of_device_time')})
print('roofline_prefill', {k: j['roofline_prefill'][k] for k in ('achieved','frac','share_of_device_time')})
print('config1', j.get('config1')); print('cpu_baseline', j.get('cpu_baseline'))
PY
