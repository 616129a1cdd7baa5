#!/usr/bin/env bash
# Round 2, GPU call 11: (call 9 repeated with the persistent kernels actually enabled — a stray edit had turned
# ACP_GEMM_PERSISTENT off in the library calls 9 and 10 ran) cta_group::2 GEMM parity + A/B + ncu, and the
# config-4 flow at Mixtral width on one GPU with every failing call named on stderr.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=================== config 4 flow, mixtral-8x7b-l2, one GPU"
timeout -k 10 600 python bench.py --config 4 --model mixtral-8x7b-l2 --tp 1 --steps 1 --warmup 1 > gpurun_out/bench_r2_config4_dev.json 2> gpurun_out/bench_r2_config4_dev.err
echo "rc=$?"; head -c 1500 gpurun_out/bench_r2_config4_dev.err; tail -c 900 gpurun_out/bench_r2_config4_dev.json
echo "=================== 2-CTA GEMM parity"
timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -8
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "GEMM tests FAILED rc=$rc: stopping"; exit 0; fi
echo "=================== 1-CTA vs 2-CTA timing"
timeout -k 10 600 python scripts/gemm2cta_probe.py 8192 2>&1 | tail -14
timeout -k 10 600 python scripts/gemm2cta_probe.py 4096 2>&1 | tail -14
echo "=================== ncu --set full: gate/up GEMM, 4096 rows, 1-CTA then 2-CTA"
ONLY=gate_up ITERS=0 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:persistent -c 2 \
  -o gpurun_out/r2_gemm_gateup_1cta_2cta -f python scripts/gemm2cta_probe.py 4096 > gpurun_out/ncu_gemm2cta.log 2>&1
tail -4 gpurun_out/ncu_gemm2cta.log
echo "=================== engine tests with ACP_GEMM_2CTA=1"
ACP_GEMM_2CTA=1 timeout -k 10 900 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5
echo "=================== bench default (1-CTA) then ACP_GEMM_2CTA=1"
timeout -k 10 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r2_1cta.json 2> gpurun_out/bench_r2_1cta.err
ACP_GEMM_2CTA=1 timeout -k 10 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_r2_2cta.json 2> gpurun_out/bench_r2_2cta.err
python - <<'PY'
import json
for n in ("1cta", "2cta"):
    try:
        d = json.loads(open(f"gpurun_out/bench_r2_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["e2e"]["value"], d["roofline_prefill"]["frac"], d["roofline"]["frac"], d.get("clocks"))
    except Exception as e:
        print(n, "unreadable", e)
PY
