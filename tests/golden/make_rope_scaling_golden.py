"""Pins the oracle's Llama-3.1 ("llama3") RoPE frequency scaling against HuggingFace transformers:
fp32 LlamaForCausalLM with rope_scaling on the synthetic `tiny` weights -> prompt logits + greedy
tokens in tests/golden/llama_rope31_golden.npz (tests/test_oracle_numerics.py re-checks the oracle
against it without needing transformers).

    python tests/golden/make_rope_scaling_golden.py
"""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_llama_golden import COLS, SEED  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.bf16 import bits_to_f32  # noqa: E402
from oracle.llama_oracle import PRESETS, LlamaOracle, Weights  # noqa: E402

# original_max_position_embeddings is set to 64 so that a 100-token test already covers all three
# bands (untouched high frequencies, interpolated middle, slowed low frequencies)
SCALING = (8.0, 1.0, 4.0, 64)


def main():
    import torch
    from transformers import LlamaConfig as HFConfig, LlamaForCausalLM
    torch.set_num_threads(8)
    cfg = dataclasses.replace(PRESETS["tiny"], name="tiny-rope31", rope_scaling=SCALING)
    w = Weights(cfg, SEED)
    rs = {"rope_type": "llama3", "factor": SCALING[0], "low_freq_factor": SCALING[1], "high_freq_factor": SCALING[2],
          "original_max_position_embeddings": SCALING[3]}
    hc = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn, num_hidden_layers=cfg.layers,
                  num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, head_dim=cfg.head_dim,
                  rms_norm_eps=cfg.eps, rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_pos,
                  tie_word_embeddings=False, attention_bias=False, mlp_bias=False, hidden_act="silu",
                  torch_dtype="float32", rope_scaling=dict(rs))
    if hasattr(hc, "rope_parameters"):
        hc.rope_parameters = dict(rs, rope_theta=cfg.rope_theta)
    m = LlamaForCausalLM(hc).eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    sd = {"model.embed_tokens.weight": t(bits_to_f32(synth.synth_matrix(SEED, synth.TID_EMBED, cfg.vocab, cfg.hidden, cfg.w_std))),
          "lm_head.weight": t(w.lm_head()), "model.norm.weight": t(w.final_norm())}
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        qkv, gu = w.wqkv(l), w.wgu(l)
        sd[p + "self_attn.q_proj.weight"] = t(qkv[:cfg.q_dim])
        sd[p + "self_attn.k_proj.weight"] = t(qkv[cfg.q_dim:cfg.q_dim + cfg.kv_dim])
        sd[p + "self_attn.v_proj.weight"] = t(qkv[cfg.q_dim + cfg.kv_dim:])
        sd[p + "self_attn.o_proj.weight"] = t(w.wo(l))
        sd[p + "mlp.gate_proj.weight"] = t(gu[:cfg.ffn])
        sd[p + "mlp.up_proj.weight"] = t(gu[cfg.ffn:])
        sd[p + "mlp.down_proj.weight"] = t(w.wdown(l))
        sd[p + "input_layernorm.weight"] = t(w.attn_norm(l))
        sd[p + "post_attention_layernorm.weight"] = t(w.ffn_norm(l))
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected, (missing, unexpected)
    inv_hf = m.model.rotary_emb.inv_freq.double().numpy()
    rng = np.random.default_rng(43)
    prompt = np.concatenate([[128000], rng.integers(0, 256, size=99)])
    with torch.no_grad():
        hf_logits = m(torch.from_numpy(prompt)[None]).logits[0].numpy()
    orc = LlamaOracle(cfg, SEED, mode="fp32", weights=w)
    o_logits = orc.forward(prompt, all_logits=True)
    err = float(np.max(np.abs(o_logits - hf_logits)))
    plain = LlamaOracle(PRESETS["tiny"], SEED, mode="fp32", weights=Weights(PRESETS["tiny"], SEED)).forward(prompt, all_logits=True)
    print(f"max |oracle - HF| = {err:.3e}; the scaling moves logits by up to {float(np.max(np.abs(plain - hf_logits))):.3e}")
    assert err < 2e-4, err
    path = os.path.join(HERE, "llama_rope31_golden.npz")
    np.savez_compressed(path, scaling=np.array(SCALING, np.float64), prompt=prompt, inv_freq=inv_hf,
                        prompt_logits=hf_logits[:, COLS].astype(np.float32), greedy_next=np.argmax(hf_logits, axis=-1))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
