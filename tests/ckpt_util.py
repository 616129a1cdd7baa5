"""Writes HuggingFace-format Llama checkpoints (config.json + safetensors, via the real
`safetensors` library) from the oracle's seeded synthetic weights, so that the engine's checkpoint
loader can be compared bit-for-bit with its synthetic-weight path."""
import json
import os

import numpy as np
import torch
from safetensors.torch import save_file

from oracle.llama_oracle import LlamaConfig, Weights


def _bf16(a: np.ndarray) -> torch.Tensor:
    """fp32 array holding bf16-representable values -> torch.bfloat16 (exact)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16)


def hf_config(cfg: LlamaConfig, **extra) -> dict:
    c = {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": cfg.hidden,
         "num_hidden_layers": cfg.layers, "num_attention_heads": cfg.heads, "num_key_value_heads": cfg.kv_heads,
         "intermediate_size": cfg.ffn, "vocab_size": cfg.vocab, "rope_theta": cfg.rope_theta,
         "rms_norm_eps": cfg.eps, "max_position_embeddings": 8192, "hidden_act": "silu",
         "tie_word_embeddings": False, "torch_dtype": "bfloat16", "attention_bias": False, "mlp_bias": False}
    if cfg.experts:
        c.update({"architectures": ["MixtralForCausalLM"], "model_type": "mixtral", "num_local_experts": cfg.experts,
                  "num_experts_per_tok": 2, "sliding_window": None})
    c.update(extra)
    return c


def tensors_from_oracle(cfg: LlamaConfig, seed: int, dtype=torch.bfloat16) -> dict:
    w = Weights(cfg, seed)
    t = {"model.embed_tokens.weight": _bf16(w.embed_rows(np.arange(cfg.vocab))),
         "lm_head.weight": _bf16(w.lm_head()), "model.norm.weight": _bf16(w.final_norm())}
    q, kv = cfg.q_dim, cfg.kv_dim
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        qkv = w.wqkv(l)
        gu = None if cfg.experts else w.wgu(l)
        t[p + "self_attn.q_proj.weight"] = _bf16(qkv[:q])
        t[p + "self_attn.k_proj.weight"] = _bf16(qkv[q:q + kv])
        t[p + "self_attn.v_proj.weight"] = _bf16(qkv[q + kv:])
        t[p + "self_attn.o_proj.weight"] = _bf16(w.wo(l))
        if cfg.experts:   # hub layout of Mixtral-8x7B: block_sparse_moe.gate + experts.<e>.{w1 = gate, w3 = up, w2 = down}
            t[p + "block_sparse_moe.gate.weight"] = _bf16(w.router(l))
            for e in range(cfg.experts):
                egu = w.expert_gu(l, e)
                t[p + f"block_sparse_moe.experts.{e}.w1.weight"] = _bf16(egu[:cfg.ffn])
                t[p + f"block_sparse_moe.experts.{e}.w3.weight"] = _bf16(egu[cfg.ffn:])
                t[p + f"block_sparse_moe.experts.{e}.w2.weight"] = _bf16(w.expert_down(l, e))
        else:
            t[p + "mlp.gate_proj.weight"] = _bf16(gu[:cfg.ffn])
            t[p + "mlp.up_proj.weight"] = _bf16(gu[cfg.ffn:])
            t[p + "mlp.down_proj.weight"] = _bf16(w.wdown(l))
        t[p + "input_layernorm.weight"] = _bf16(w.attn_norm(l))
        t[p + "post_attention_layernorm.weight"] = _bf16(w.ffn_norm(l))
    if dtype != torch.bfloat16:
        t = {k: v.to(dtype) for k, v in t.items()}
    return t


def write_checkpoint(path: str, cfg: LlamaConfig, seed: int, shards: int = 1, dtype=torch.bfloat16, **cfg_extra) -> dict:
    os.makedirs(path, exist_ok=True)
    t = tensors_from_oracle(cfg, seed, dtype)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config(cfg, **cfg_extra), f)
    names = sorted(t)
    if shards <= 1:
        save_file({k: t[k].contiguous() for k in names}, os.path.join(path, "model.safetensors"))
    else:
        weight_map = {}
        for s in range(shards):
            part = names[s::shards]
            fname = f"model-{s + 1:05d}-of-{shards:05d}.safetensors"
            save_file({k: t[k].contiguous() for k in part}, os.path.join(path, fname))
            weight_map.update({k: fname for k in part})
        with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
            json.dump({"metadata": {}, "weight_map": weight_map}, f)
    return t
