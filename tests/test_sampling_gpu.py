"""Sampling parity: the engine's sampler (csrc/kernels.cu sample_kernel: temperature / top-k / top-p,
counter RNG) against oracle/sampling_oracle.py evaluated on the engine's OWN returned logits — so the
comparison isolates the sampler — plus a distribution test over many seeds."""
import numpy as np
import pytest

from agentcontrolplane_b200.engine import Engine
from oracle import sampling_oracle as S

pytestmark = pytest.mark.gpu
MODEL = "tiny"


@pytest.fixture(scope="module")
def eng():
    with Engine({"model": MODEL, "max_batch": 64, "kv_pages": 512, "max_tokens_per_step": 1024, "prefix_cache": False}) as e:
        yield e


def _run(eng, prompt, n_new, **sampling):
    t = eng.submit(dict({"model": MODEL, "max_tokens": n_new, "acp": {"prompt_token_ids": prompt, "return_logits": n_new}}, **sampling))
    assert eng.wait(t, 120000)
    lg = eng.logits(t, n_new, 128256)
    st, body = eng.result(t)
    assert st == 200, body
    return t, body["acp"]["token_ids"], lg


CASES = [dict(temperature=1.0), dict(temperature=0.7), dict(temperature=1.3, top_k=50), dict(temperature=0.8, top_k=5),
         dict(temperature=1.0, top_p=0.9), dict(temperature=1.5, top_p=0.5), dict(temperature=0.9, top_k=40, top_p=0.8),
         dict(temperature=2.0, top_k=1), dict(temperature=1.0, top_p=0.05)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_sampled_tokens_match_the_oracle_on_the_engines_logits(eng, case):
    rng = np.random.default_rng(hash(str(case)) % (1 << 31))
    exact = total = 0
    for rep in range(6):
        prompt = [128000] + [int(t) for t in rng.integers(0, 256, size=int(rng.integers(3, 60)))]
        seed = int(rng.integers(1, 1 << 62))
        _, toks, lg = _run(eng, prompt, 5, seed=seed, **case)
        for step, tok in enumerate(toks):
            T, k, p = case["temperature"], case.get("top_k", 0), case.get("top_p", 1.0)
            want = S.sample(lg[step], T, k, p, seed, step)
            allowed = S.candidates(lg[step], T, k, p, seed, step)
            assert tok in allowed, (case, rep, step, tok, want, sorted(allowed))
            assert S.kept_set(lg[step], T, k, p, slack=2e-5)[tok] or S.kept_set(lg[step], T, k, p, slack=-2e-5)[tok]
            exact += int(tok == want)
            total += 1
            if tok in (128001, 128008, 128009):
                break
    assert exact >= 0.9 * total, (exact, total)      # the eps band is rarely needed


def test_default_seed_is_a_per_ticket_stream(eng):
    prompt = [128000, 72, 105]
    t, toks, lg = _run(eng, prompt, 3, temperature=1.0)
    for step, tok in enumerate(toks):
        assert tok in S.candidates(lg[step], 1.0, 0, 1.0, S.default_seed(t), step)


def test_temperature_zero_rows_in_a_sampling_batch_stay_greedy(eng):
    """a greedy request batched with sampling requests takes the sampler's arg-max path: same tokens as alone"""
    rng = np.random.default_rng(3)
    prompt = [128000] + [int(t) for t in rng.integers(0, 256, size=30)]
    _, alone, _ = _run(eng, prompt, 4)
    ts = [eng.submit({"model": MODEL, "max_tokens": 4, "temperature": 1.0, "seed": 5 + i, "acp": {"prompt_token_ids": prompt}}) for i in range(3)]
    tg = eng.submit({"model": MODEL, "max_tokens": 4, "acp": {"prompt_token_ids": prompt}})
    for t in ts + [tg]:
        assert eng.wait(t, 120000)
    st, body = eng.result(tg)
    assert st == 200 and body["acp"]["token_ids"] == alone
    for t in ts:
        eng.result(t)


def test_sampling_distribution_over_seeds(eng):
    """2000 seeds on ONE logit vector (top_k 8): empirical frequencies vs softmax(l / T) on the kept set."""
    prompt = [128000, 84, 104, 101, 32, 99, 97, 116]
    _, _, lg = _run(eng, prompt, 1)
    T, k, n = 1.2, 8, 2000
    ts = [eng.submit({"model": MODEL, "max_tokens": 1, "temperature": T, "top_k": k, "seed": 1000 + i,
                      "acp": {"prompt_token_ids": prompt}}) for i in range(n)]
    counts = {}
    for t in ts:
        assert eng.wait(t, 300000)
        st, body = eng.result(t)
        assert st == 200
        tok = body["acp"]["token_ids"][0]
        counts[tok] = counts.get(tok, 0) + 1
    keep = S.kept_set(lg[0], T, k, 1.0)
    assert set(counts) <= set(np.nonzero(keep)[0].tolist())
    w = np.where(keep, np.exp((lg[0].astype(np.float64) - lg[0].max()) / T), 0.0)
    prob = w / w.sum()
    chi2 = sum((counts.get(int(i), 0) - n * prob[i]) ** 2 / (n * prob[i]) for i in np.nonzero(keep)[0])
    assert chi2 < 30.0, (chi2, counts)     # 7 degrees of freedom: P(chi2 > 30) ~ 1e-4
