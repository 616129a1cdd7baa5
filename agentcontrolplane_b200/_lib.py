"""ctypes loader for libacp_infer.so (the C-ABI boundary, include/acp_infer.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C agentcontrolplane_b200/csrc``).
There is no fallback: if the shared object is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ACP_INFER_LIB: load a library built elsewhere (e.g. a scratch build while the in-tree one is in use)
LIB_PATH = os.environ.get("ACP_INFER_LIB") or os.path.join(_HERE, "lib", "libacp_infer.so")
HOST_LIB_PATH = os.environ.get("ACP_HOST_LIB") or os.path.join(_HERE, "lib", "libacp_host.so")
_lib = None
_host = None


def load_host() -> ctypes.CDLL:
    """libacp_host.so: the C++ mirror of the reference's Go host code (include/acp_host.h).  Pure C++,
    no CUDA, NOT linked against libacp_infer.so — the reference arm of bench.py maps only this one."""
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _host = ctypes.CDLL(HOST_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    return _host


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA engine)")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    return _lib
