"""Pins oracle/llama_oracle.py (fp32 mode) against HuggingFace transformers' LlamaForCausalLM on
the synthetic `tiny` and `tiny-g2` configs and writes tests/golden/llama_tiny_golden.npz.

Run here (CPU container, transformers 5.5 installed):  python tests/golden/make_llama_golden.py
The fixture holds: the prompt, HF's fp32 logits at selected vocabulary columns for every prompt
position and 8 greedy steps, and HF's greedy token ids.  tests/test_oracle_numerics.py re-checks
the oracle against it without needing transformers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from oracle.bf16 import bits_to_f32  # noqa: E402
from oracle.llama_oracle import PRESETS, LlamaOracle, Weights  # noqa: E402

SEED = 0xACB200
COLS = np.concatenate([np.arange(0, 256), np.arange(127990, 128256), np.arange(1000, 128000, 997)])


def hf_model(cfg, w: Weights):
    import torch
    from transformers import LlamaConfig as HFConfig, LlamaForCausalLM
    hc = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn,
                  num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                  num_key_value_heads=cfg.kv_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.eps,
                  rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_pos,
                  tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
                  hidden_act="silu", torch_dtype="float32")
    if hasattr(hc, "rope_parameters"):
        hc.rope_parameters = {"rope_type": "default", "rope_theta": cfg.rope_theta}
    m = LlamaForCausalLM(hc).eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    sd = {}
    sd["model.embed_tokens.weight"] = t(bits_to_f32(synth.synth_matrix(SEED, synth.TID_EMBED, cfg.vocab, cfg.hidden, cfg.w_std)))
    sd["lm_head.weight"] = t(w.lm_head())
    sd["model.norm.weight"] = t(w.final_norm())
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        qkv = w.wqkv(l)
        sd[p + "self_attn.q_proj.weight"] = t(qkv[:cfg.q_dim])
        sd[p + "self_attn.k_proj.weight"] = t(qkv[cfg.q_dim:cfg.q_dim + cfg.kv_dim])
        sd[p + "self_attn.v_proj.weight"] = t(qkv[cfg.q_dim + cfg.kv_dim:])
        sd[p + "self_attn.o_proj.weight"] = t(w.wo(l))
        gu = w.wgu(l)
        sd[p + "mlp.gate_proj.weight"] = t(gu[:cfg.ffn])
        sd[p + "mlp.up_proj.weight"] = t(gu[cfg.ffn:])
        sd[p + "mlp.down_proj.weight"] = t(w.wdown(l))
        sd[p + "input_layernorm.weight"] = t(w.attn_norm(l))
        sd[p + "post_attention_layernorm.weight"] = t(w.ffn_norm(l))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "rotary" not in k], missing
    assert not unexpected, unexpected
    return m


def main():
    import torch
    torch.set_num_threads(8)
    out = {"seed": np.int64(SEED), "cols": COLS}
    for name in ("tiny", "tiny-g2"):
        cfg = PRESETS[name]
        w = Weights(cfg, SEED)
        model = hf_model(cfg, w)
        rng = np.random.default_rng(42)
        prompt = np.concatenate([[128000], rng.integers(0, 256, size=37), [128006, 128007, 128009]])
        steps = 8
        with torch.no_grad():
            ids = torch.from_numpy(prompt)[None]
            hf_logits = model(ids).logits[0].numpy()          # [T][V]
            toks, step_logits = [], []
            cur = ids
            for _ in range(steps):
                lg = model(cur).logits[0, -1].numpy()
                step_logits.append(lg[COLS])
                nxt = int(np.argmax(lg))
                toks.append(nxt)
                cur = torch.cat([cur, torch.tensor([[nxt]])], dim=1)
        orc = LlamaOracle(cfg, SEED, mode="fp32", weights=w)
        o_logits = orc.forward(prompt, all_logits=True)
        err = float(np.max(np.abs(o_logits - hf_logits)))
        print(f"{name}: max |oracle_fp32 - HF_fp32| over prompt logits = {err:.3e}")
        assert err < 2e-4, err
        out[f"{name}.prompt"] = prompt
        out[f"{name}.prompt_logits"] = hf_logits[:, COLS].astype(np.float32)
        out[f"{name}.greedy"] = np.array(toks, np.int64)
        out[f"{name}.step_logits"] = np.array(step_logits, np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "llama_tiny_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
