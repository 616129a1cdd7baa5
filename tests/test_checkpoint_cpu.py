"""Checkpoint reader (csrc/safetensors.cc, model_config_from_hf) against files written by the real
`safetensors` library — no GPU: only the container format, config mapping and error behaviour."""
import json
import os

import numpy as np
import pytest
import torch

from agentcontrolplane_b200 import host
from oracle.llama_oracle import LlamaConfig, rope_tables
from ckpt_util import write_checkpoint

CFG = LlamaConfig("ckpt-test", hidden=256, layers=2, heads=2, kv_heads=1, ffn=512, vocab=1024)
SEED = 77


def test_index_matches_what_safetensors_wrote(tmp_path):
    d = str(tmp_path / "m")
    t = write_checkpoint(d, CFG, SEED)
    idx = host.checkpoint_index(d)
    assert set(idx["tensors"]) == set(t)
    for name, ten in t.items():
        e = idx["tensors"][name]
        assert e["dtype"] == "BF16" and e["shape"] == list(ten.shape) and e["nbytes"] == ten.numel() * 2
    m = idx["model"]
    assert (m["hidden"], m["layers"], m["heads"], m["kv_heads"], m["ffn"], m["vocab"]) == (256, 2, 2, 1, 512, 1024)
    assert m["rope_theta"] == 500000.0 and abs(m["eps"] - 1e-5) < 1e-12 and m["tied_embeddings"] is False
    # a single file path works too (config.json beside it)
    assert host.checkpoint_index(os.path.join(d, "model.safetensors"))["model"] == m


def test_sharded_checkpoint_and_other_dtypes(tmp_path):
    d = str(tmp_path / "sharded")
    t = write_checkpoint(d, CFG, SEED, shards=3)
    idx = host.checkpoint_index(d)
    assert set(idx["tensors"]) == set(t) and len(os.listdir(d)) == 5
    d32 = str(tmp_path / "f32")
    write_checkpoint(d32, CFG, SEED, dtype=torch.float32)
    assert {e["dtype"] for e in host.checkpoint_index(d32)["tensors"].values()} == {"F32"}


def test_rope_frequencies_match_oracle_and_hf(tmp_path):
    d = str(tmp_path / "rope")
    write_checkpoint(d, CFG, SEED)
    inv = np.array(host.checkpoint_index(d)["model"]["rope_inv_freq"], dtype=np.float32)
    cos, sin = rope_tables(CFG, 64)
    ang = (np.arange(64, dtype=np.float32)[:, None] * inv[None, :]).astype(np.float32)
    assert np.array_equal(np.cos(ang.astype(np.float64)).astype(np.float32), cos)
    # Llama-3.1 "llama3" frequency scaling: same numbers as transformers' implementation
    scaling = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
               "original_max_position_embeddings": 8192}
    d2 = str(tmp_path / "rope31")
    write_checkpoint(d2, CFG, SEED, rope_scaling=scaling, max_position_embeddings=131072)
    inv31 = np.array(host.checkpoint_index(d2)["model"]["rope_inv_freq"], dtype=np.float64)
    from transformers import LlamaConfig as HFConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    hf = HFConfig(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, rope_theta=500000.0,
                  max_position_embeddings=131072, rope_scaling=scaling)
    want, _ = ROPE_INIT_FUNCTIONS["llama3"](hf, "cpu")
    np.testing.assert_allclose(inv31, want.double().numpy(), rtol=3e-7)
    assert inv31[-1] < inv[-1] / 7.9          # the lowest frequency really is slowed 8x
    # engine (C++) and oracle (numpy) build the SAME fp32 table, bit for bit, scaled or not
    import dataclasses
    from oracle.llama_oracle import rope_inv_freq
    got31 = np.array(host.checkpoint_index(d2)["model"]["rope_inv_freq"], dtype=np.float32)
    assert np.array_equal(got31, rope_inv_freq(dataclasses.replace(CFG, rope_scaling=(8.0, 1.0, 4.0, 8192))))
    assert np.array_equal(inv, rope_inv_freq(CFG))


@pytest.mark.parametrize("breakage,needle", [
    ("no_config", "config.json"), ("head_dim", "head_dim"), ("truncated", "header"), ("vocab", "multiples"),
    ("model_type", "not supported"), ("missing_dir", "does not exist")])
def test_unusable_checkpoints_are_rejected_with_a_reason(tmp_path, breakage, needle):
    d = str(tmp_path / "bad")
    if breakage == "missing_dir":
        with pytest.raises(ValueError, match=needle):
            host.checkpoint_index(d)
        return
    extra = {"head_dim": {"head_dim": 64}, "vocab": {"vocab_size": 1000}, "model_type": {"model_type": "gpt_neox"}}.get(breakage, {})
    write_checkpoint(d, CFG, SEED, **extra)
    if breakage == "no_config":
        os.remove(os.path.join(d, "config.json"))
    if breakage == "truncated":
        with open(os.path.join(d, "model.safetensors"), "r+b") as f:
            f.truncate(64)
    with pytest.raises(ValueError, match=needle):
        host.checkpoint_index(d)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_dtype_conversion_is_round_to_nearest_even(tmp_path, dtype):
    """F32 / F16 checkpoints are converted on the host exactly as torch converts to bfloat16
    (round to nearest even; subnormals, infinities and NaN included); BF16 passes through."""
    from safetensors.torch import save_file
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.standard_normal(4096).astype(np.float32) * np.float32(0.02),
                           np.array([0.0, -0.0, 1.0, -1.0, 1.00390625, 1.0078125, 1.01171875, 3.3895314e38, 65504.0, 6.1e-5, 5.96e-8,
                                     1e-40, -1e-40, np.inf, -np.inf, np.nan, 2.0 ** -126, 2.0 ** -133], dtype=np.float32)])
    with np.errstate(over="ignore"):
        t = torch.from_numpy(vals).to(dtype)
    f = str(tmp_path / "t.safetensors")
    save_file({"x": t}, f)
    got = host.checkpoint_tensor_bf16(f, "x")
    want = t.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    nan = np.isnan(t.float().numpy())
    assert np.array_equal(got[~nan], want[~nan])
    assert ((got[nan] & 0x7F80) == 0x7F80).all() and ((got[nan] & 0x007F) != 0).all()      # NaN stays NaN
