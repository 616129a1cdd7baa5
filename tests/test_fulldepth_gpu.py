"""Parity of the CUDA engine at the FULL BASELINE model — Llama-3-8B shapes, all 32 layers — on ONE GPU,
against two independent checkers that run as fp32 torch on the same GPU (test side only):

  (a) tests/torch_oracle.py `TorchLlamaOracle` (bf16 mode): the numpy oracle's restatement with the
      engine's bf16 rounding points, pinned to the numpy oracle and to HuggingFace on CPU
      (tests/test_torch_oracle_cpu.py).  Greedy token ids are compared bit-exactly under the stated
      near-tie policy, logits within LOGIT_ATOL_32L.
  (b) HuggingFace `transformers.LlamaForCausalLM` holding the same weights, fp32 compute — "the
      reference's llmclient path pointed at the same weights" (SURVEY.md §8c(iii);
      acp/internal/llmclient/langchaingo_client.go:102 is the call a provider serving this checkpoint
      answers).  HF has no bf16 rounding points, so this comparison is statistical: relative RMS
      logit error, and arg-max agreement wherever HF's own top-1/top-2 margin clears the noise.

Workload = BASELINE config 1's batch: 64 concurrent requests x 512-token windows (all 64 run through
the engine together; the checkers run a sample, one forward pass of the 8B model each).
"""
import numpy as np
import pytest

from agentcontrolplane_b200.engine import Engine
from oracle.llama_oracle import PRESETS

pytestmark = pytest.mark.gpu

MODEL = "llama-3-8b"
SEED = 0xACB200
N_NEW = 8
# |engine - torch oracle (bf16 mode)| on fp32 logits at 32 layers.  Both round to bf16 at the same
# points; fp32 summation order (tcgen05 tensor core vs cuBLAS) flips single bf16 roundings, and 32
# layers of residual stream carry ~4x the flips of the 2-layer presets (3e-2 there).
LOGIT_ATOL_32L = 1.2e-1
# engine (bf16 rounding points) vs HuggingFace fp32: relative RMS error of the logit vector
HF_REL_RMS = 3e-2


@pytest.fixture(scope="module")
def run():
    import torch
    from torch_oracle import TorchWeights
    rng = np.random.default_rng(20260921)
    prompts = [[128000] + [int(t) for t in rng.integers(0, 256, size=511)] for _ in range(64)]
    sample = [0, 9, 22, 37, 50, 63]
    with Engine({"model": MODEL, "max_batch": 64, "kv_pages": 64 * 18 + 8, "max_tokens_per_step": 4096,
                 "max_pages_per_seq": 32, "prefix_cache": False}) as eng:
        ts = [eng.submit({"model": MODEL, "max_tokens": N_NEW,
                          "acp": {"prompt_token_ids": p, "return_logits": N_NEW if i in sample else 0}})
              for i, p in enumerate(prompts)]
        outs = {}
        for i, t in enumerate(ts):
            assert eng.wait(t, 600000)
            lg = eng.logits(t, N_NEW, 128256) if i in sample else None
            st, body = eng.result(t)
            assert st == 200, body
            outs[i] = (body["acp"]["token_ids"], lg)
        stats = eng.stats()
    assert stats["layers"] == 32 and stats["prefill_tokens"] == 64 * 512
    weights = TorchWeights(PRESETS[MODEL], SEED, device="cuda:0")
    yield prompts, sample, outs, weights
    del weights
    torch.cuda.empty_cache()


def test_full_depth_tokens_and_logits_match_the_bf16_oracle(run):
    from torch_oracle import TorchLlamaOracle
    prompts, sample, outs, weights = run
    worst = 0.0
    for i in sample:
        got, lg = outs[i]
        want, margins, ref_lg = TorchLlamaOracle(weights).greedy(prompts[i], N_NEW, eos=(128001, 128008, 128009))
        n_cmp = 0
        for j, (g, w) in enumerate(zip(got, want)):
            d = float(np.max(np.abs(lg[j] - ref_lg[j])))
            worst = max(worst, d)
            assert d < LOGIT_ATOL_32L, (i, j, d)
            n_cmp += 1
            if g != w:
                # near-tie policy: either candidate is legitimate when the oracle's own
                # margin is inside the logit tolerance; the continuations differ from here on
                assert margins[j] < 2 * LOGIT_ATOL_32L, (i, j, got, want, margins)
                break
        else:
            assert len(got) == len(want)
        assert n_cmp >= 1
    print(f"full-depth 8B: max |logit diff| vs bf16 torch oracle over {len(sample)} x <= {N_NEW} positions: {worst:.4f}")


def test_full_depth_logits_match_huggingface_fp32(run):
    import torch
    from torch_oracle import hf_model_from_weights
    prompts, sample, outs, weights = run
    model = hf_model_from_weights(weights, "cuda:0")
    rels, agree, decided = [], 0, 0
    with torch.no_grad():
        for i in sample:
            got, lg = outs[i]
            ref = model(torch.tensor([prompts[i]], device="cuda:0")).logits[0, -1].float().cpu().numpy()
            diff = lg[0] - ref
            rel = float(np.sqrt(np.mean(diff ** 2)) / np.std(ref))
            rels.append(rel)
            top2 = np.partition(ref, -2)[-2:]
            if top2[1] - top2[0] > 6 * float(np.sqrt(np.mean(diff ** 2))):   # HF's choice is clear of the bf16 noise
                decided += 1
                agree += int(got[0] == int(np.argmax(ref)))
            assert np.corrcoef(lg[0], ref)[0, 1] > 0.999
    del model
    torch.cuda.empty_cache()
    print(f"full-depth 8B vs HF fp32: relative RMS logit error {max(rels):.4f} (max over {len(rels)}); "
          f"arg-max agreement {agree}/{decided} where HF's margin > 6 sigma")
    assert max(rels) < HF_REL_RMS, rels
    assert agree == decided
