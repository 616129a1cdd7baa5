"""Dev probe: one prefill-attention launch through the unit-test hook (for ncu).
usage: python scripts/attn_probe.py [T] [impl] [heads] [kv_heads] [iters]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agentcontrolplane_b200 import _lib  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 32
kvh = int(sys.argv[4]) if len(sys.argv) > 4 else 8
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rng = np.random.default_rng(1)
bits = lambda shape: ((rng.standard_normal(shape).astype(np.float32).view(np.uint32) + 0x8000) >> 16).astype(np.uint16)
q, k, v = bits((T, heads, 128)), bits((T, kvh, 128)), bits((T, kvh, 128))
out = np.zeros((T, heads, 128), np.uint16)
u16p = ctypes.POINTER(ctypes.c_uint16)
lib = _lib.load()
lib.acp_kernel_attn_prefill.argtypes = [u16p, u16p, u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, u16p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
ms = ctypes.c_float(0)
rc = lib.acp_kernel_attn_prefill(q.ctypes.data_as(u16p), k.ctypes.data_as(u16p), v.ctypes.data_as(u16p), heads, kvh, T, T,
                                 impl, out.ctypes.data_as(u16p), iters, ctypes.byref(ms))
assert rc == 0, rc
if iters:
    flops = 4.0 * heads * 128 * (T * (T + 1) / 2)
    print(f"attn_prefill impl={impl} T={T} heads={heads}/{kvh}: {ms.value * 1e3:.1f} us  {flops / ms.value / 1e9:.1f} TFLOP/s")
