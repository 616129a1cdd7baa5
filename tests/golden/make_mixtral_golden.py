"""Generates tests/golden/mixtral_tiny_golden.npz: fp32 logits of HuggingFace `MixtralForCausalLM` on the
oracle's seeded synthetic weights (tiny-moe preset: 8 experts, top-2) — pins oracle/llama_oracle.py's
mixture-of-experts block (router softmax, top-2, renormalised weights, SwiGLU experts) to transformers.
Run in the build container:  python tests/golden/make_mixtral_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.llama_oracle import PRESETS, Weights  # noqa: E402

SEED = 0xACB200


def build(cfg):
    from transformers import MixtralConfig, MixtralForCausalLM
    hf = MixtralConfig(hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                       num_key_value_heads=cfg.kv_heads, intermediate_size=cfg.ffn, vocab_size=cfg.vocab, head_dim=cfg.head_dim,
                       rope_theta=cfg.rope_theta, rms_norm_eps=cfg.eps, max_position_embeddings=cfg.max_pos,
                       num_local_experts=cfg.experts, num_experts_per_tok=2, sliding_window=None, tie_word_embeddings=False,
                       attn_implementation="eager", router_jitter_noise=0.0)
    model = MixtralForCausalLM(hf).float().eval()
    w = Weights(cfg, SEED)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    sd = model.state_dict()
    with torch.no_grad():
        sd["model.embed_tokens.weight"].copy_(t(w.embed_rows(np.arange(cfg.vocab))))
        sd["lm_head.weight"].copy_(t(w.lm_head()))
        sd["model.norm.weight"].copy_(t(w.final_norm()))
        q, kv = cfg.q_dim, cfg.kv_dim
        for l in range(cfg.layers):
            p = f"model.layers.{l}."
            qkv = w.wqkv(l)
            sd[p + "self_attn.q_proj.weight"].copy_(t(qkv[:q]))
            sd[p + "self_attn.k_proj.weight"].copy_(t(qkv[q:q + kv]))
            sd[p + "self_attn.v_proj.weight"].copy_(t(qkv[q + kv:]))
            sd[p + "self_attn.o_proj.weight"].copy_(t(w.wo(l)))
            sd[p + "input_layernorm.weight"].copy_(t(w.attn_norm(l)))
            sd[p + "post_attention_layernorm.weight"].copy_(t(w.ffn_norm(l)))
            moe = [k for k in sd if k.startswith(p) and ("experts" in k or "gate" in k)]
            gate_key = [k for k in moe if k.endswith("gate.weight")][0]
            sd[gate_key].copy_(t(w.router(l)))
            gu_key = [k for k in moe if k.endswith("gate_up_proj")][0]
            dn_key = [k for k in moe if k.endswith("down_proj")][0]
            sd[gu_key].copy_(torch.stack([t(w.expert_gu(l, e)) for e in range(cfg.experts)]))
            sd[dn_key].copy_(torch.stack([t(w.expert_down(l, e)) for e in range(cfg.experts)]))
    return model


if __name__ == "__main__":
    cfg = PRESETS["tiny-moe"]
    model = build(cfg)
    rng = np.random.default_rng(11)
    prompt = [128000] + [int(x) for x in rng.integers(0, 256, size=47)]
    with torch.no_grad():
        logits = model(torch.tensor([prompt])).logits[0].float().numpy()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mixtral_tiny_golden.npz")
    # keep the fixture small: every position's logits at 512 vocabulary slots + the full last position's top-16
    cols = np.linspace(0, cfg.vocab - 1, 512).astype(np.int64)
    np.savez_compressed(out, prompt=np.array(prompt), cols=cols, logits=logits[:, cols],
                        last_top=np.argsort(-logits[-1])[:16], last_top_vals=np.sort(logits[-1])[::-1][:16].copy())
    print("wrote", out, logits.shape)
