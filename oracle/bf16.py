"""bf16 <-> fp32 helpers for the oracle (numpy, round-to-nearest-even like __float2bfloat16_rn).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np


def bf16_round_to_bits(x: np.ndarray) -> np.ndarray:
    """fp32 array -> uint16 bf16 bit patterns, RNE (NaN kept quiet)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    lsb = (u >> 16) & 1
    rounded = (u + 0x7FFF + lsb) >> 16
    nan = np.isnan(x)
    out = rounded.astype(np.uint16)
    if np.any(nan):
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 value, returned as fp32."""
    return bits_to_f32(bf16_round_to_bits(x))
