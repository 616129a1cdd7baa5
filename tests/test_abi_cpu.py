"""The C-ABI library loads on a CPU-only box, exports every symbol the headers declare, and the
CUDA engine fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

from agentcontrolplane_b200 import _lib
from agentcontrolplane_b200.engine import Engine, EngineError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acp_[a-z_0-9]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["acp_infer.h", "acp_infer_kernels.h", "acp_host.h"])
def test_every_declared_symbol_is_exported(header):
    lib = _lib.load()
    names = _declared(header)
    assert names, header
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"


def test_expected_boundary_entry_points_present():
    assert set(_declared("acp_infer.h")) >= {
        "acp_infer_init", "acp_infer_submit", "acp_infer_wait", "acp_infer_poll", "acp_infer_result",
        "acp_infer_cancel", "acp_infer_stats", "acp_infer_free", "acp_infer_shutdown"}


def test_version_and_no_device_behaviour():
    lib = _lib.load()
    lib.acp_infer_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.acp_infer_version()
    if lib.acp_kernel_device_count() == 0:
        with pytest.raises(EngineError) as ei:
            Engine({"model": "tiny"})
        assert ei.value.code == -5          # ACP_ERR_CUDA: loud failure, never a CPU path


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "agentcontrolplane_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
