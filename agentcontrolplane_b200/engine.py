"""ctypes binding of the C-ABI boundary (include/acp_infer.h) — plumbing for tests, bench.py and
the Python mirror of the reference's llmclient interface.  No compute happens in Python."""
from __future__ import annotations

import ctypes
import json
from typing import Any

import numpy as np

from . import _lib

ACP_OK = 0
ACP_ERR_TIMEOUT = -3
ACP_ERR_PENDING = -7


class EngineError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what} failed with ACP error {code}")
        self.code = code


def _bind(lib):
    vp, u64 = ctypes.c_void_p, ctypes.c_uint64
    lib.acp_infer_init.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.acp_infer_init.restype = ctypes.c_int
    lib.acp_infer_submit.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(u64)]
    lib.acp_infer_submit.restype = ctypes.c_int
    lib.acp_infer_wait.argtypes = [vp, u64, ctypes.c_int]
    lib.acp_infer_wait.restype = ctypes.c_int
    lib.acp_infer_poll.argtypes = [vp, ctypes.POINTER(u64), ctypes.c_int, ctypes.c_int]
    lib.acp_infer_poll.restype = ctypes.c_int
    lib.acp_infer_result.argtypes = [vp, u64, ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)]
    lib.acp_infer_result.restype = ctypes.c_int
    lib.acp_infer_result_logits.argtypes = [vp, u64, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.acp_infer_result_logits.restype = ctypes.c_int
    lib.acp_infer_cancel.argtypes = [vp, u64]
    lib.acp_infer_cancel.restype = None
    lib.acp_infer_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p)]
    lib.acp_infer_stats.restype = ctypes.c_int
    lib.acp_infer_stats_reset.argtypes = [vp]
    lib.acp_infer_stats_reset.restype = None
    lib.acp_infer_free.argtypes = [ctypes.c_void_p]
    lib.acp_infer_free.restype = None
    lib.acp_infer_shutdown.argtypes = [vp]
    lib.acp_infer_shutdown.restype = None
    lib.acp_infer_version.restype = ctypes.c_char_p
    return lib


class Engine:
    """Process-wide engine handle (acp_infer_init .. acp_infer_shutdown)."""

    def __init__(self, config: dict[str, Any] | None = None):
        self._lib = _bind(_lib.load())
        self._h = ctypes.c_void_p()
        self.config = dict(config or {})
        rc = self._lib.acp_infer_init(json.dumps(self.config).encode(), ctypes.byref(self._h))
        if rc != ACP_OK:
            raise EngineError(rc, "acp_infer_init")

    def submit(self, request: dict[str, Any] | bytes) -> int:
        body = request if isinstance(request, (bytes, bytearray)) else json.dumps(request).encode()
        t = ctypes.c_uint64(0)
        rc = self._lib.acp_infer_submit(self._h, body, len(body), ctypes.byref(t))
        if rc != ACP_OK:
            raise EngineError(rc, "acp_infer_submit")
        return t.value

    def wait(self, ticket: int, timeout_ms: int = -1) -> bool:
        rc = self._lib.acp_infer_wait(self._h, ticket, timeout_ms)
        if rc == ACP_ERR_TIMEOUT:
            return False
        if rc != ACP_OK:
            raise EngineError(rc, "acp_infer_wait")
        return True

    def poll(self, max_tickets: int = 256, timeout_ms: int = 0) -> list[int]:
        arr = (ctypes.c_uint64 * max_tickets)()
        n = self._lib.acp_infer_poll(self._h, arr, max_tickets, timeout_ms)
        if n < 0:
            raise EngineError(n, "acp_infer_poll")
        return [arr[i] for i in range(n)]

    def logits(self, ticket: int, positions: int, vocab: int) -> np.ndarray:
        out = np.zeros((positions, vocab), np.float32)
        n = self._lib.acp_infer_result_logits(
            self._h, ticket, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), positions)
        if n < 0:
            raise EngineError(n, "acp_infer_result_logits")
        return out[:n]

    def result(self, ticket: int) -> tuple[int, dict[str, Any]]:
        buf, ln, st = ctypes.c_void_p(), ctypes.c_size_t(0), ctypes.c_int(0)
        rc = self._lib.acp_infer_result(self._h, ticket, ctypes.byref(buf), ctypes.byref(ln),
                                        ctypes.byref(st))
        if rc != ACP_OK:
            raise EngineError(rc, "acp_infer_result")
        try:
            body = ctypes.string_at(buf, ln.value)
        finally:
            self._lib.acp_infer_free(buf)
        return st.value, json.loads(body)

    def complete(self, request: dict[str, Any], timeout_ms: int = -1) -> tuple[int, dict[str, Any]]:
        t = self.submit(request)
        if not self.wait(t, timeout_ms):
            self.cancel(t)
            self.wait(t, -1)
        return self.result(t)

    def cancel(self, ticket: int) -> None:
        self._lib.acp_infer_cancel(self._h, ticket)

    def stats(self) -> dict[str, Any]:
        buf = ctypes.c_void_p()
        rc = self._lib.acp_infer_stats(self._h, ctypes.byref(buf))
        if rc != ACP_OK:
            raise EngineError(rc, "acp_infer_stats")
        try:
            return json.loads(ctypes.string_at(buf))
        finally:
            self._lib.acp_infer_free(buf)

    def stats_reset(self) -> None:
        self._lib.acp_infer_stats_reset(self._h)

    def close(self) -> None:
        if self._h:
            self._lib.acp_infer_shutdown(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
