"""Pins tests/torch_oracle.py (the checker the full-depth GPU parity test uses) to the numpy oracle:
bit-equal synthetic weights, logits within 1e-4, identical greedy tokens — on CPU, small presets."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle.bf16 import bits_to_f32
from oracle.llama_oracle import PRESETS, LlamaOracle
from torch_oracle import TorchLlamaOracle, TorchWeights, hf_model_from_weights, synth_tensor

SEED = 0xACB200


def test_synth_bits_identical():
    for tid, n, std, plus, start in [(1, 5000, 0.02, False, 0), (3, 512, 0.1, True, 0), (2, 4097, 0.02, False, 123456789),
                                     (synth.layer_tid(31, synth.TID_WDOWN), 70000, 0.02, False, 4096 * 14336 - 1000)]:
        want = bits_to_f32(synth.synth_bits(SEED, tid, n, std, plus_one=plus, start=start))
        got = synth_tensor(SEED, tid, n, std, plus_one=plus, start=start, chunk=4099).to(torch.float32).numpy()
        assert np.array_equal(want, got), (tid, n)


@pytest.mark.parametrize("name", ["tiny", "tiny-g8"])
def test_forward_matches_numpy_oracle(name):
    cfg = PRESETS[name]
    rng = np.random.default_rng(5)
    prompt = [128000] + [int(t) for t in rng.integers(0, 256, size=70)]
    w = TorchWeights(cfg, SEED)
    t_or = TorchLlamaOracle(w)
    n_or = LlamaOracle(cfg, SEED, mode="bf16")
    got_t, m_t, lg_t = t_or.greedy(prompt, 5, eos=(128001, 128008, 128009))
    want, margins = n_or.greedy(prompt, 5, eos=(128001, 128008, 128009))
    ref0 = LlamaOracle(cfg, SEED, mode="bf16").forward(prompt)[-1]
    # both round at the same points; fp32 summation order differs (BLAS vs torch): a flipped bf16
    # rounding moves a logit by ~1e-3 at these sizes
    assert np.max(np.abs(lg_t[0] - ref0)) < 2e-2, float(np.max(np.abs(lg_t[0] - ref0)))
    for i, (g, wnt) in enumerate(zip(got_t, want)):
        if g != wnt:
            assert margins[i] < 4e-2, (i, got_t, want, margins)
            break


def test_fp32_mode_matches_huggingface():
    """The torch restatement without rounding points == transformers.LlamaForCausalLM (fp32) on the same
    synthetic weights: the checker of the full-depth GPU test is pinned to HF directly, not only via numpy."""
    cfg = PRESETS["tiny-g2"]
    w = TorchWeights(cfg, SEED)
    rng = np.random.default_rng(6)
    prompt = [128000] + [int(t) for t in rng.integers(0, 256, size=50)]
    ours = TorchLlamaOracle(w, mode="fp32").forward(prompt, all_logits=True).numpy()
    model = hf_model_from_weights(w, "cpu")
    with torch.no_grad():
        ref = model(torch.tensor([prompt])).logits[0].float().numpy()
    assert np.max(np.abs(ours - ref)) < 1e-4, float(np.max(np.abs(ours - ref)))
