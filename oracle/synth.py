"""Seeded synthetic weights — the numpy statement of the generator the engine runs on the GPU
(agentcontrolplane_b200/csrc/kernels.cu, `synth_weight_kernel`).  Integer-only hashing plus ONE
fp32 multiply, so CPU and GPU produce bit-identical bf16 tensors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

There are no model weights on disk and no network, so every parity and bench
run uses these tensors at the named shapes.  Element i of tensor `tid` under `seed`:

    z  = seed + tid * 0x9E3779B97F4A7C15 + i * 0xD1B54A32D192ED03          (mod 2^64)
    z  = splitmix64_finalise(z)
    s  = sum of the four 16-bit fields of z  - 131070                      (Irwin-Hall, ~normal)
    w  = bf16_rne( float32(s) * float32(std / 37837.2267) [+ 1.0 for norm gains] )
"""
from __future__ import annotations

import numpy as np

from .bf16 import bf16_round_to_bits

MASK = np.uint64(0xFFFFFFFFFFFFFFFF)
IH_STD = 37837.22671196048  # sqrt(4 * (65536**2 - 1) / 12)

# tensor ids (shared with csrc/model.h)
TID_EMBED, TID_LM_HEAD, TID_FINAL_NORM = 1, 2, 3
TID_LAYER_BASE, TID_LAYER_STRIDE = 16, 16
TID_WQKV, TID_WO, TID_WGU, TID_WDOWN, TID_ATTN_NORM, TID_FFN_NORM, TID_ROUTER = 0, 1, 2, 3, 4, 5, 6
# mixture-of-experts layers (Mixtral): expert e of layer l owns two tensors, [gate; up] and down
TID_MOE_BASE, TID_MOE_LAYER_STRIDE = 1 << 20, 256
TID_EXPERT_GU, TID_EXPERT_DOWN = 0, 1


def layer_tid(layer: int, which: int) -> int:
    return TID_LAYER_BASE + layer * TID_LAYER_STRIDE + which


def expert_tid(layer: int, expert: int, which: int) -> int:
    return TID_MOE_BASE + layer * TID_MOE_LAYER_STRIDE + expert * 2 + which


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.copy()
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def _synth_chunk(out: np.ndarray, base: np.uint64, start: int, c0: int, c1: int, scale: np.float32,
                 plus_one: bool) -> None:
    with np.errstate(over="ignore"):
        idx = np.arange(start + c0, start + c1, dtype=np.uint64)
        z = _mix(base + idx * np.uint64(0xD1B54A32D192ED03))
        s = ((z & np.uint64(0xFFFF)) + ((z >> np.uint64(16)) & np.uint64(0xFFFF)) +
             ((z >> np.uint64(32)) & np.uint64(0xFFFF)) + (z >> np.uint64(48))).astype(np.int64)
        s -= 131070
        w = s.astype(np.float32) * scale
        if plus_one:
            w = w + np.float32(1.0)
        out[c0:c1] = bf16_round_to_bits(w)


def synth_bits(seed: int, tid: int, n: int, std: float, plus_one: bool = False,
               start: int = 0) -> np.ndarray:
    """bf16 bit patterns of elements [start, start+n) of tensor `tid`.  Large tensors are generated
    in chunks on a thread pool (numpy releases the GIL inside these element-wise kernels); every
    element depends only on its own index, so the result does not depend on the chunking."""
    scale = np.float32(std / IH_STD)
    out = np.empty(n, np.uint16)
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) + np.uint64(tid) * np.uint64(0x9E3779B97F4A7C15)) & MASK
    chunk = 1 << 22
    spans = [(c0, min(n, c0 + chunk)) for c0 in range(0, n, chunk)]
    if len(spans) <= 2:
        for c0, c1 in spans:
            _synth_chunk(out, base, start, c0, c1, scale, plus_one)
        return out
    import os
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        list(pool.map(lambda sp: _synth_chunk(out, base, start, sp[0], sp[1], scale, plus_one), spans))
    return out


def synth_matrix(seed: int, tid: int, rows: int, cols: int, std: float) -> np.ndarray:
    return synth_bits(seed, tid, rows * cols, std).reshape(rows, cols)


def synth_rows(seed: int, tid: int, row_ids: np.ndarray, cols: int, std: float) -> np.ndarray:
    """Selected rows of a [rows][cols] tensor without materialising it (embedding lookups)."""
    out = np.empty((len(row_ids), cols), np.uint16)
    for j, r in enumerate(row_ids):
        out[j] = synth_bits(seed, tid, cols, std, start=int(r) * cols)
    return out
