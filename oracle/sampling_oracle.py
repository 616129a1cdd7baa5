"""Numpy restatement of the local provider's sampler (agentcontrolplane_b200/csrc/kernels.cu
`sample_kernel`): temperature, top-k, top-p and the counter-based uniform.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference never samples locally — `LangchainClient.SendRequest` sends the window to a provider
(acp/internal/llmclient/langchaingo_client.go:102) and ACP sets no sampling option, so the wire says
temperature 0.  `LLM.spec.parameters` still carries temperature / topP / topK
(acp/api/v1alpha1/llm_types.go:41-71), and the local provider honours them.  "parity unpinned" against
the reference (it has no sampler); pinned to the definition below:

    w_i   = exp((l_i - max l) / T)
    top-k : keep the k largest logits
    top-p : keep the smallest set of largest-weight tokens whose mass >= p * Z (Z over the top-k set)
    u     = (splitmix64(seed * 0x9E3779B97F4A7C15 + step + 1) >> 40) / 2^24        (24-bit uniform)
    token = first kept index (in VOCABULARY order) whose cumulative weight exceeds u * Z_kept

The kernel evaluates the sums in fp32 in a fixed tree order; this restatement uses float64, so a
target that falls within `eps` of a CDF step (or a top-p mass within eps of the cut) legitimately
resolves either way: `candidates()` returns every token a correct implementation may emit.
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1


def splitmix64(z: int) -> int:
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def uniform(seed: int, step: int) -> float:
    """the kernel's u in [0, 1): top 24 bits of splitmix64(seed * golden + step + 1)"""
    r = splitmix64((seed * 0x9E3779B97F4A7C15 + step + 1) & M64)
    return (r >> 40) / 16777216.0


def default_seed(ticket: int) -> int:
    """engine.cc step(): a request without `seed` gets a per-ticket stream"""
    return (ticket * 0x9E3779B97F4A7C15) & M64


def kept_set(logits: np.ndarray, temperature: float, top_k: int, top_p: float, slack: float = 0.0) -> np.ndarray:
    """boolean mask of the tokens that survive top-k then top-p (mass cut moved by `slack` of Z)"""
    l = logits.astype(np.float64)
    V = l.shape[0]
    keep = np.ones(V, bool)
    if 0 < top_k < V:
        kth = np.partition(l, V - top_k)[V - top_k]
        keep &= l >= kth
    w = np.where(keep, np.exp((l - l.max()) / temperature), 0.0)
    if top_p < 1.0:
        order = np.argsort(-w, kind="stable")
        csum = np.cumsum(w[order])
        need = (top_p + slack) * csum[-1]
        n_keep = int(np.searchsorted(csum, need, side="left")) + 1
        cut = w[order[min(n_keep, V) - 1]]
        keep &= w >= cut      # ties at the cut stay together, like the kernel's threshold
    return keep


def sample(logits: np.ndarray, temperature: float, top_k: int, top_p: float, seed: int, step: int,
           slack: float = 0.0, du: float = 0.0) -> int:
    if temperature <= 0.0:
        return int(np.argmax(logits))
    keep = kept_set(logits, temperature, top_k, top_p, slack)
    l = logits.astype(np.float64)
    w = np.where(keep, np.exp((l - l.max()) / temperature), 0.0)
    target = min(max(uniform(seed, step) + du, 0.0), 1.0) * w.sum()
    csum = np.cumsum(w)
    idx = int(np.searchsorted(csum, target, side="right"))
    idx = min(idx, len(w) - 1)
    while not keep[idx]:       # target == Z corner
        idx -= 1
    return idx


def candidates(logits: np.ndarray, temperature: float, top_k: int, top_p: float, seed: int, step: int,
               eps: float = 5e-4) -> set[int]:
    """Every token a correct fp32 implementation may emit: the kernel sums up to 128256 fp32 weights in a
    fixed tree order (relative error ~1e-4 of Z) where this restatement uses float64, so the draw may land
    anywhere in the CDF band [u*Z - eps*Z, u*Z + eps*Z], and the top-p cut may move by eps of Z."""
    if temperature <= 0.0:
        return {int(np.argmax(logits))}
    out = set()
    l = logits.astype(np.float64)
    for slack in (0.0, -eps, eps):
        keep = kept_set(logits, temperature, top_k, top_p, slack)
        w = np.where(keep, np.exp((l - l.max()) / temperature), 0.0)
        z = w.sum()
        csum = np.cumsum(w)
        target = uniform(seed, step) * z
        lo = int(np.searchsorted(csum, max(target - eps * z, 0.0), side="right"))
        hi = int(np.searchsorted(csum, min(target + eps * z, z), side="right"))
        idx = np.nonzero(keep[lo:min(hi, len(w) - 1) + 1])[0] + lo
        out.update(int(i) for i in idx)
        out.add(sample(logits, temperature, top_k, top_p, seed, step, slack, 0.0))
    return out
