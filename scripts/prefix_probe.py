import sys, os
sys.path.insert(0, os.getcwd())
from agentcontrolplane_b200 import host
from agentcontrolplane_b200.engine import Engine
e = Engine({"model": "tiny", "max_batch": 64, "kv_pages": 4096, "max_tokens_per_step": 4096, "max_pages_per_seq": 40, "prefix_cache": True})
for step in range(3):
    s0 = e.stats()
    r = host.hostsim_run({"tasks": 32, "workers": 32, "provider": "local", "model": "tiny", "max_tokens": 64, "prompt_tokens": 512, "tools": 2, "tool_loop": True, "seed": step + 1}, e)
    s1 = e.stats()
    print(step, r["reconciles"], r["final_phases"], "hits", s1["prefix_hits"] - s0["prefix_hits"], "reused", s1["prefix_tokens_reused"] - s0["prefix_tokens_reused"],
          "prefill_tokens", s1["prefill_tokens"] - s0["prefill_tokens"], "cache pages", s1["prefix_cache_pages"], flush=True)
e.close()
