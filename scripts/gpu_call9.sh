#!/usr/bin/env bash
# Round 2, GPU call 9: validate the cta_group::2 persistent prefill GEMM (bit-identical to the 1-CTA kernel),
# time both at the prefill shapes, then the engine tests and the bench with it switched on.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== config 4 flow at 2 layers on one GPU (every Task must end FinalAnswer)"
timeout -k 10 600 python bench.py --config 4 --model mixtral-8x7b-l2 --tp 1 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4_dev.json 2> gpurun_out/bench_r2_config4_dev.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_r2_config4_dev.err; tail -c 700 gpurun_out/bench_r2_config4_dev.json
echo "=================== 2-CTA GEMM parity"
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "two_cta" 2>&1 | tail -8
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "2-CTA parity FAILED rc=$rc: stopping"; exit 0; fi
echo "=================== 1-CTA vs 2-CTA timing"
timeout -k 10 600 python scripts/gemm2cta_probe.py 8192 2>&1 | tail -20
echo "=================== engine tests with ACP_GEMM_2CTA=1"
ACP_GEMM_2CTA=1 timeout -k 10 900 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py tests/test_fulldepth_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "=================== bench default (1-CTA) then ACP_GEMM_2CTA=1"
timeout -k 10 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r2_1cta.json 2> gpurun_out/bench_r2_1cta.err; tail -c 1500 gpurun_out/bench_r2_1cta.json | cut -c1-600
ACP_GEMM_2CTA=1 timeout -k 10 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r2_2cta.json 2> gpurun_out/bench_r2_2cta.err; tail -c 1500 gpurun_out/bench_r2_2cta.json | cut -c1-600
python - <<'PY'
import json
for n in ("1cta", "2cta"):
    try:
        d = json.loads(open(f"gpurun_out/bench_r2_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d.get("roofline_prefill"), d.get("clocks"))
    except Exception as e:
        print(n, "unreadable", e)
PY
