"""Pins oracle/llama_oracle.py (fp32 mode) against HuggingFace transformers at the REAL Llama-3-8B
width (hidden 4096, 32 query / 8 KV heads, ffn 14336, vocab 128256), truncated to 2 layers so that
it runs on a CPU: tests/golden/llama_8b_l2_golden.npz.  Needs ~12 GB of RAM and a few minutes.

    python tests/golden/make_llama_8b_l2_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_llama_golden import COLS, SEED, hf_model  # noqa: E402
from oracle.llama_oracle import PRESETS, LlamaOracle, Weights  # noqa: E402


def main():
    import torch
    torch.set_num_threads(8)
    cfg = PRESETS["llama-3-8b-l2"]
    t0 = time.time()
    w = Weights(cfg, SEED)
    model = hf_model(cfg, w)
    print("HF model built in", round(time.time() - t0, 1), "s")
    rng = np.random.default_rng(44)
    prompt = np.concatenate([[128000], rng.integers(0, 256, size=45), [128006, 128007]])
    with torch.no_grad():
        hf_logits = model(torch.from_numpy(prompt)[None]).logits[0].numpy()
    orc = LlamaOracle(cfg, SEED, mode="fp32", weights=w)
    o_logits = orc.forward(prompt, all_logits=True)
    err = float(np.max(np.abs(o_logits - hf_logits)))
    print(f"llama-3-8b-l2: max |oracle_fp32 - HF_fp32| over {len(prompt)} positions x {cfg.vocab} logits = {err:.3e}")
    assert err < 5e-4, err
    path = os.path.join(HERE, "llama_8b_l2_golden.npz")
    np.savez_compressed(path, prompt=prompt, prompt_logits=hf_logits[:, COLS].astype(np.float32),
                        greedy_next=np.argmax(hf_logits, axis=-1), max_abs_err_at_generation=np.float64(err))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
