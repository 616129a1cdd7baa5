"""Open-loop arrivals probe: Poisson arrivals of Tasks whose windows are
log-uniform on [128, 4096] tokens (BASELINE config 2's distribution) at `rate` Tasks/s for `seconds`;
reports end-to-end latency, time to first token and the inter-token gap of decoding sequences under
the two scheduling policies ("decode_interleave": 0 = prefill first, k = one decode step per k chunks).
usage: python scripts/arrivals_probe.py [rate] [seconds] [interleave ...]"""
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agentcontrolplane_b200.engine import Engine  # noqa: E402

rate = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
policies = [int(a) for a in sys.argv[3:]] or [0, 1]
MODEL, MAX_NEW = "llama-3-8b", 64
rng = np.random.default_rng(0xA221)
n = int(rate * seconds)
gaps = rng.exponential(1.0 / rate, size=n)
lens = np.exp(rng.uniform(np.log(128), np.log(4096), size=n)).astype(int)
prompts = [[128000] + [int(t) for t in np.random.default_rng(i).integers(0, 256, size=int(l) - 1)] for i, l in enumerate(lens)]
for k in policies:
    cfg = {"model": MODEL, "max_batch": 512, "kv_pages": int(sum((l + MAX_NEW) // 32 + 2 for l in lens)) + 64,
           "max_tokens_per_step": int(os.environ.get("MTPS", "2048")), "max_pages_per_seq": (4096 + MAX_NEW) // 32 + 2,
           "prefix_cache": False, "decode_interleave": k}
    with Engine(cfg) as eng:
        warm = eng.submit({"model": MODEL, "max_tokens": 4, "acp": {"prompt_token_ids": prompts[0][:256]}})
        eng.wait(warm, -1); eng.result(warm); eng.stats_reset()
        t_sub, t_done, bodies = {}, {}, {}

        def collector():
            while len(t_done) < n:
                for t in eng.poll(256, 50):
                    t_done[t] = time.perf_counter()

        th = threading.Thread(target=collector)
        th.start()
        t0 = time.perf_counter()
        due = t0
        for i in range(n):
            due += gaps[i]
            while time.perf_counter() < due:
                time.sleep(0.0002)
            t = eng.submit({"model": MODEL, "max_tokens": MAX_NEW, "acp": {"prompt_token_ids": prompts[i]}})
            t_sub[t] = time.perf_counter()
        th.join()
        wall = time.perf_counter() - t0
        ttft, per_tok = [], []
        for t in t_sub:
            st, body = eng.result(t)
            assert st == 200, body
            ext = body.get("acp", {})
            ttft.append(ext.get("queue_ms", 0) + ext.get("prefill_ms", 0))
            per_tok.append(ext.get("decode_ms", 0) / max(1, len(ext.get("token_ids", [])) - 1))
        lat = np.array([t_done[t] - t_sub[t] for t in t_sub]) * 1e3
        s = eng.stats()
        q = lambda a, p: float(np.percentile(a, p)) if len(a) else None
        print(json.dumps({"decode_interleave": k, "rate_per_s": rate, "tasks": n, "wall_s": round(wall, 2),
                          "completed_per_s": round(n / wall, 2), "latency_ms_p50": q(lat, 50), "latency_ms_p99": q(lat, 99),
                          "ttft_ms_p50": q(ttft, 50), "ttft_ms_p99": q(ttft, 99),
                          "ms_per_generated_token_p50": q(per_tok, 50), "ms_per_generated_token_p99": q(per_tok, 99),
                          "decode_step_ms_p50": s.get("decode_step_ms_p50"), "decode_step_ms_p99": s.get("decode_step_ms_p99"),
                          "decode_steps": s["decode_steps"], "prefill_steps": s["prefill_steps"],
                          "mean_decode_batch": round(s["decode_tokens"] / max(1, s["decode_steps"]), 1)}), flush=True)
