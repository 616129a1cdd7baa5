#!/usr/bin/env bash
# Round 2, GPU call 12 (one B200): config-4 flow at Mixtral width (split-K workspace fix), then what the driver
# runs at round end with the cta_group::2 GEMM as the default, an 8192-row-step comparison, and memcheck.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== config 4 flow, mixtral-8x7b-l2, one GPU"
timeout -k 10 600 python bench.py --config 4 --model mixtral-8x7b-l2 --tp 1 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4_dev.json 2> gpurun_out/bench_r2_config4_dev.err
echo "rc=$?"; head -c 1500 gpurun_out/bench_r2_config4_dev.err; tail -c 700 gpurun_out/bench_r2_config4_dev.json
bash scripts/gpu_final_check.sh
echo "=================== bench.py --max-tokens-per-step 8192"
timeout -k 10 900 python bench.py --steps 3 --warmup 2 --max-tokens-per-step 8192 > gpurun_out/bench_r2_8192.json 2> gpurun_out/bench_r2_8192.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_8192.json").read().strip().splitlines()[-1])
    print("8192-row steps:", d["value"], d["e2e"]["value"], d["roofline_prefill"]["frac"], d["roofline"]["frac"], d.get("clocks"))
except Exception as e:
    print("unreadable", e)
PY
echo "=================== compute-sanitizer memcheck (tiny, tiny-g8, tiny-moe engines + hooks incl. both persistent GEMMs)"
timeout -k 10 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_probe.py 2>&1 | tail -12
