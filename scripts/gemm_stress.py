"""Back-to-back launch stress of the persistent GEMM kernels at the shapes a tensor-parallel shard runs (short K:
the accumulator hand-off between the MMA issuer and the epilogue warps is the critical path, not the operand
stream).  A pipeline deadlock ends in the kernels' bounded waits (8 s) with block / thread / barrier printed.
usage: python scripts/gemm_stress.py [launches_per_case]      Run under gpurun.  Dev tool only."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agentcontrolplane_b200 import _lib  # noqa: E402

lib = _lib.load()
u16p = ctypes.POINTER(ctypes.c_uint16)
rng = np.random.default_rng(1)
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
# (M, N, K, epi): Mixtral TP=8 O-proj (K=512), 70B TP=8 O-proj (K=1024) and QKV (M=1280), 8B dense for reference
cases = [(4096, 4096, 512, 1), (8192, 4096, 1024, 1), (768, 4096, 4096, 0), (1280, 4096, 8192, 0), (4096, 1024, 512, 1),
         (4096, 4096, 4096, 0), (7168, 4096, 8192, 3)]
for M, N, K, epi in cases:
    w = rng.integers(0x3000, 0x3C00, size=(M, K), dtype=np.uint16)
    x = rng.integers(0x3000, 0x3F80, size=(N, K), dtype=np.uint16)
    for mode, name in ((-2, "2cta"), (-1, "1cta")):
        flops = 2.0 * M * N * K
        n = max(2000, min(n_launch, int(n_launch * 6.9e10 / flops)))      # ~equal time per case
        out = np.zeros((1, N, M), np.float32) if epi == 1 else np.zeros((N, M // 2 if epi == 3 else M), np.uint16)
        ms = ctypes.c_float(0)
        t0 = time.time()
        rc = lib.acp_kernel_gemm(w.ctypes.data_as(u16p), x.ctypes.data_as(u16p), M, N, K, 1, epi, mode,
                                 out.ctypes.data_as(ctypes.c_void_p), None, None, -n, ctypes.byref(ms))
        print(json.dumps(dict(M=M, N=N, K=K, epi=epi, kernel=name, launches=n, rc=rc, us_per_launch=round(ms.value * 1e3, 2),
                              tflops=round(flops / (ms.value * 1e-3) / 1e12, 1) if ms.value > 0 else 0, wall_s=round(time.time() - t0, 1))), flush=True)
        if rc != 0:
            sys.exit(1)
