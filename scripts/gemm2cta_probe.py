"""A/B of the two persistent prefill GEMM kernels (1-CTA vs cta_group::2) at the Llama-3-8B prefill shapes
(N = tokens of one prefill chunk) through the C-ABI hook acp_kernel_gemm (bn -1 / -2).  L2 is flushed
between the timed launches.  Run under gpurun; writes gpurun_out/gemm2cta_probe.json.  Dev tool only."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agentcontrolplane_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    u16p = ctypes.POINTER(ctypes.c_uint16)
    rng = np.random.default_rng(0)
    shapes = [("qkv", 6144, 4096, 0), ("o", 4096, 4096, 0), ("gate_up", 28672, 4096, 3), ("down", 4096, 14336, 0)]
    Ns = [int(a) for a in sys.argv[1:]] or [8192]
    only, iters = os.environ.get("ONLY"), int(os.environ.get("ITERS", "10"))   # ONLY=gate_up ITERS=0: one launch each (ncu)
    shapes = [s for s in shapes if not only or s[0] == only]
    res = []
    for name, M, K, epi in shapes:
        w = rng.integers(0x3000, 0x3C00, size=(M, K), dtype=np.uint16)
        w ^= (rng.integers(0, 2, size=(M, K), dtype=np.uint16) << 15)      # random signs
        for N in Ns:
            x = rng.integers(0x3000, 0x3F80, size=(N, K), dtype=np.uint16)
            outs = {}
            for mode in (-1, -2):
                out = np.zeros((N, M // 2 if epi == 3 else M), np.uint16)
                ms = ctypes.c_float(0)
                rc = lib.acp_kernel_gemm(w.ctypes.data_as(u16p), x.ctypes.data_as(u16p), M, N, K, 1, epi, mode,
                                         out.ctypes.data_as(ctypes.c_void_p), None, None, iters, ctypes.byref(ms))
                outs[mode] = out
                tf = 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0
                r = dict(gemm=name, M=M, K=K, N=N, kernel="1cta" if mode == -1 else "2cta", rc=rc, ms=ms.value, tflops=tf)
                print(json.dumps(r), flush=True)
                res.append(r)
            print(json.dumps(dict(gemm=name, N=N, bit_identical=bool(np.array_equal(outs[-1], outs[-2])))), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gemm2cta_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
