"""Torch mirror of oracle/llama_oracle.py (bf16 mode) and oracle/synth.py — TEST INFRASTRUCTURE.

Why it exists: the numpy oracle cannot run the full-depth Llama-3-8B (32 layers, 64 x 512-token
windows = 4.9e14 FLOP) in seconds; on the GPU box the same restatement runs as fp32 torch matmuls
(TF32 off) on cuda:0 next to the engine.  It is a CHECKER: nothing in the product imports it, and it
is itself pinned against the numpy oracle (tests/test_torch_oracle_cpu.py, bit-equal weights,
logits within 1e-4 at the small presets) which is pinned against HuggingFace transformers.

Same rounding points as oracle/llama_oracle.py (every `_r` is a round-to-nearest-even to bf16).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from oracle import synth
from oracle.llama_oracle import LlamaConfig, rope_tables

_M64 = (1 << 64) - 1


def _i64(v: int) -> int:
    """two's-complement int64 view of a uint64 constant (torch has no uint64 arithmetic)"""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z: torch.Tensor, n: int) -> torch.Tensor:
    """logical shift right of int64 bit patterns"""
    return (z >> n) & ((1 << (64 - n)) - 1)


def synth_tensor(seed: int, tid: int, n: int, std: float, plus_one: bool = False, start: int = 0,
                 device="cpu", chunk: int = 1 << 26) -> torch.Tensor:
    """bf16 tensor of elements [start, start+n) of synthetic tensor `tid` — oracle/synth.py
    `synth_bits`, restated with int64 wrap-around arithmetic."""
    out = torch.empty(n, dtype=torch.bfloat16, device=device)
    base = _i64(seed + tid * 0x9E3779B97F4A7C15)
    scale = torch.tensor(np.float32(std / synth.IH_STD).item(), dtype=torch.float32, device=device)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        idx = torch.arange(start + c0, start + c1, dtype=torch.int64, device=device)
        z = idx * _i64(0xD1B54A32D192ED03) + base
        z = z ^ _lsr(z, 30)
        z = z * _i64(0xBF58476D1CE4E5B9)
        z = z ^ _lsr(z, 27)
        z = z * _i64(0x94D049BB133111EB)
        z = z ^ _lsr(z, 31)
        s = (z & 0xFFFF) + (_lsr(z, 16) & 0xFFFF) + (_lsr(z, 32) & 0xFFFF) + _lsr(z, 48)
        w = (s - 131070).to(torch.float32) * scale
        if plus_one:
            w = w + 1.0
        out[c0:c1] = w.to(torch.bfloat16)   # RNE, like bf16_round_to_bits
    return out


class TorchWeights:
    """bf16 weights of the seeded synthetic model on `device`, logical (un-tiled, un-sharded) layout."""

    def __init__(self, cfg: LlamaConfig, seed: int, device="cpu"):
        self.cfg, self.seed, self.device = cfg, seed, device
        c = cfg
        mat = lambda tid, r, k: synth_tensor(seed, tid, r * k, c.w_std, device=device).view(r, k)
        gain = lambda tid: synth_tensor(seed, tid, c.hidden, 0.1, plus_one=True, device=device)
        self.embed = mat(synth.TID_EMBED, c.vocab, c.hidden)
        self.lm_head = mat(synth.TID_LM_HEAD, c.vocab, c.hidden)
        self.final_norm = gain(synth.TID_FINAL_NORM)
        self.layers = []
        for l in range(c.layers):
            t = lambda which: synth.layer_tid(l, which)
            self.layers.append({
                "wqkv": mat(t(synth.TID_WQKV), c.q_dim + 2 * c.kv_dim, c.hidden),
                "wo": mat(t(synth.TID_WO), c.hidden, c.q_dim),
                "wgu": mat(t(synth.TID_WGU), 2 * c.ffn, c.hidden),
                "wdown": mat(t(synth.TID_WDOWN), c.hidden, c.ffn),
                "attn_norm": gain(t(synth.TID_ATTN_NORM)),
                "ffn_norm": gain(t(synth.TID_FFN_NORM)),
            })

    def hf_state_dict(self) -> dict:
        """The same tensors under HuggingFace LlamaForCausalLM names (what tests/ckpt_util.py writes)."""
        c = self.cfg
        sd = {"model.embed_tokens.weight": self.embed, "lm_head.weight": self.lm_head,
              "model.norm.weight": self.final_norm}
        q, kv = c.q_dim, c.kv_dim
        for l, L in enumerate(self.layers):
            p = f"model.layers.{l}."
            sd[p + "self_attn.q_proj.weight"] = L["wqkv"][:q]
            sd[p + "self_attn.k_proj.weight"] = L["wqkv"][q:q + kv]
            sd[p + "self_attn.v_proj.weight"] = L["wqkv"][q + kv:]
            sd[p + "self_attn.o_proj.weight"] = L["wo"]
            sd[p + "mlp.gate_proj.weight"] = L["wgu"][:c.ffn]
            sd[p + "mlp.up_proj.weight"] = L["wgu"][c.ffn:]
            sd[p + "mlp.down_proj.weight"] = L["wdown"]
            sd[p + "input_layernorm.weight"] = L["attn_norm"]
            sd[p + "post_attention_layernorm.weight"] = L["ffn_norm"]
        return sd


def _bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _mm(x: torch.Tensor, w_bf16: torch.Tensor) -> torch.Tensor:
    """x[T][K] fp32 @ W[M][K]^T with fp32 accumulation (bf16 weights are exact in fp32)."""
    return x @ w_bf16.to(torch.float32).t()


class TorchLlamaOracle:
    """One sequence with a KV cache; forward() mirrors LlamaOracle.forward() in bf16 mode."""

    def __init__(self, weights: TorchWeights, mode: str = "bf16"):
        assert mode in ("bf16", "fp32")   # fp32 = no rounding points: the mode that is compared with HF
        self._r = _bf16_round if mode == "bf16" else (lambda x: x)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.w, self.cfg = weights, weights.cfg
        dev = weights.device
        cos, sin = rope_tables(self.cfg, self.cfg.max_pos)
        self.cos = torch.from_numpy(cos).to(dev)
        self.sin = torch.from_numpy(sin).to(dev)
        self.k = [None] * self.cfg.layers
        self.v = [None] * self.cfg.layers
        self.pos = 0

    def _rmsnorm(self, x, g):
        var = (x * x).mean(dim=-1, keepdim=True)
        rstd = 1.0 / torch.sqrt(var + self.cfg.eps)
        return self._r(g.to(torch.float32) * self._r(x * rstd))

    def _rope(self, x, pos):
        half = self.cfg.head_dim // 2
        c = self.cos[pos][:, None, :]
        s = self.sin[pos][:, None, :]
        x1, x2 = x[..., :half], x[..., half:]
        return self._r(torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1))

    @torch.no_grad()
    def forward(self, tokens, all_logits: bool = False) -> torch.Tensor:
        c, dev, _r = self.cfg, self.w.device, self._r
        toks = torch.as_tensor(list(tokens), dtype=torch.int64, device=dev)
        T = toks.numel()
        pos = torch.arange(self.pos, self.pos + T, device=dev)
        x = self.w.embed[toks].to(torch.float32)
        scale = 1.0 / math.sqrt(c.head_dim)
        group = c.heads // c.kv_heads
        for l, L in enumerate(self.w.layers):
            xn = self._rmsnorm(x, L["attn_norm"])
            qkv = _r(_mm(xn, L["wqkv"]))
            q = self._rope(qkv[:, :c.q_dim].reshape(T, c.heads, c.head_dim), pos)
            k = self._rope(qkv[:, c.q_dim:c.q_dim + c.kv_dim].reshape(T, c.kv_heads, c.head_dim), pos)
            v = qkv[:, c.q_dim + c.kv_dim:].reshape(T, c.kv_heads, c.head_dim)
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], dim=0)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], dim=0)
            K, V = self.k[l], self.v[l]
            S = K.shape[0]
            causal = torch.arange(S, device=dev)[None, :] <= pos[:, None]              # [T][S]
            Kh = K.permute(1, 0, 2).repeat_interleave(group, dim=0)                    # [H][S][d]
            Vh = V.permute(1, 0, 2).repeat_interleave(group, dim=0)
            s = torch.matmul(q.permute(1, 0, 2), Kh.transpose(1, 2)) * scale           # [H][T][S]
            s = s.masked_fill(~causal[None], float("-inf"))
            p = torch.exp(s - s.max(dim=-1, keepdim=True).values)
            a = torch.matmul(p, Vh) / p.sum(dim=-1, keepdim=True)                      # [H][T][d]
            attn = _r(a.permute(1, 0, 2).reshape(T, c.q_dim))
            x = _r(x + _r(_mm(attn, L["wo"])))
            xn2 = self._rmsnorm(x, L["ffn_norm"])
            gu = _r(_mm(xn2, L["wgu"]))
            g, u = gu[:, :c.ffn], gu[:, c.ffn:]
            act = _r(g / (1.0 + torch.exp(-g)))
            x = _r(x + _r(_mm(_r(act * u), L["wdown"])))
        self.pos += T
        xf = self._rmsnorm(x if all_logits else x[-1:], self.w.final_norm)
        return _mm(xf, self.w.lm_head)

    def greedy(self, prompt, max_new: int, eos=()):
        """(token ids, top1-top2 margins, fp32 logits of every sampled position as numpy)"""
        out, margins, all_lg = [], [], []
        logits = self.forward(prompt)[-1]
        for step in range(max_new):
            top2 = torch.topk(logits, 2).values
            margins.append(float(top2[0] - top2[1]))
            mx = logits.max()
            tok = int(torch.nonzero(logits == mx)[0, 0])     # lowest index on ties
            all_lg.append(logits.cpu().numpy())
            out.append(tok)
            if tok in eos or step == max_new - 1:
                break
            logits = self.forward([tok])[-1]
        return out, margins, all_lg


def hf_model_from_weights(weights: TorchWeights, device, dtype=torch.float32):
    """HuggingFace `LlamaForCausalLM` holding exactly `weights` (fp32 compute on bf16-valued
    parameters) — "the reference's llmclient path pointed at the same weights" (SURVEY.md §8c(iii)):
    what a provider serving this checkpoint computes, up to its own rounding."""
    from transformers import LlamaConfig as HFConfig
    from transformers import LlamaForCausalLM
    c = weights.cfg
    hf = HFConfig(hidden_size=c.hidden, num_hidden_layers=c.layers, num_attention_heads=c.heads,
                  num_key_value_heads=c.kv_heads, intermediate_size=c.ffn, vocab_size=c.vocab, head_dim=c.head_dim,
                  rope_theta=c.rope_theta, rms_norm_eps=c.eps, max_position_embeddings=c.max_pos, hidden_act="silu",
                  tie_word_embeddings=False, attention_bias=False, mlp_bias=False, attn_implementation="eager")
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(hf)
    finally:
        torch.set_default_dtype(prev)
    sd = weights.hf_state_dict()
    with torch.no_grad():
        own = model.state_dict()
        missing = [k for k in own if k not in sd and "rotary" not in k]
        assert not missing, missing
        for k, v in sd.items():
            own[k].copy_(v.to(dtype))
    return model.eval()
