"""Dev probe (GPU): how far do the engine's logits sit from (a) the bf16-rounding torch oracle and (b)
HuggingFace fp32 at FULL depth, as a function of the synthetic weight scale?  Prints one line per w_std.
usage: python scripts/fulldepth_probe.py [w_std ...]"""
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from agentcontrolplane_b200.engine import Engine  # noqa: E402
from oracle.llama_oracle import PRESETS  # noqa: E402
from torch_oracle import TorchLlamaOracle, TorchWeights, hf_model_from_weights  # noqa: E402

MODEL = os.environ.get("MODEL", "llama-3-8b")
LAYERS = int(os.environ.get("LAYERS", "0"))
stds = [float(a) for a in sys.argv[1:]] or [0.02, 0.01, 0.005, 0.002]
rng = np.random.default_rng(20260921)
prompts = [[128000] + [int(t) for t in rng.integers(0, 256, size=511)] for _ in range(4)]
N_NEW = 4
for w_std in stds:
    cfg = dataclasses.replace(PRESETS[MODEL], w_std=w_std)
    ecfg = {"model": MODEL, "max_batch": 8, "kv_pages": 4 * 18 + 8, "max_tokens_per_step": 4096,
            "max_pages_per_seq": 32, "prefix_cache": False, "w_std": w_std}
    if LAYERS:
        cfg = dataclasses.replace(cfg, layers=LAYERS)
        ecfg["layers"] = LAYERS
    with Engine(ecfg) as eng:
        ts = [eng.submit({"model": MODEL, "max_tokens": N_NEW, "acp": {"prompt_token_ids": p, "return_logits": N_NEW}}) for p in prompts]
        outs = []
        for t in ts:
            assert eng.wait(t, 600000)
            lg = eng.logits(t, N_NEW, 128256)
            st, body = eng.result(t)
            assert st == 200, body
            outs.append((body["acp"]["token_ids"], lg))
    w = TorchWeights(cfg, 0xACB200, device="cuda:0")
    worst, tok_ok, tok_n, stds_l, margins_l = 0.0, 0, 0, [], []
    for p, (got, lg) in zip(prompts, outs):
        want, margins, ref = TorchLlamaOracle(w).greedy(p, N_NEW, eos=(128001, 128008, 128009))
        for j, (g, x) in enumerate(zip(got, want)):
            worst = max(worst, float(np.max(np.abs(lg[j] - ref[j]))))
            stds_l.append(float(np.std(ref[j])))
            margins_l.append(margins[j])
            tok_n += 1
            tok_ok += int(g == x)
            if g != x:
                break
    hf = hf_model_from_weights(w, "cuda:0")
    rels = []
    with torch.no_grad():
        for p, (got, lg) in zip(prompts, outs):
            ref = hf(torch.tensor([p], device="cuda:0")).logits[0, -1].float().cpu().numpy()
            rels.append(float(np.sqrt(np.mean((lg[0] - ref) ** 2)) / np.std(ref)))
    del hf, w
    torch.cuda.empty_cache()
    print(f"w_std={w_std} layers={cfg.layers}: max|engine - bf16 oracle|={worst:.5f}  logit std={np.mean(stds_l):.4f}  "
          f"rel={worst / np.mean(stds_l):.5f}  tokens equal {tok_ok}/{tok_n}  mean margin={np.mean(margins_l):.4f}  "
          f"engine vs HF fp32 rel RMS={max(rels):.5f}", flush=True)
