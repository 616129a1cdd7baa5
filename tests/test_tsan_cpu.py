"""Host side under ThreadSanitizer (the reference runs `go test` without -race; the
engine boundary here is multi-producer, so the C++ mirror of the reconcile loop is built with
-fsanitize=thread in a host-only binary — tests/tsan/: real host sources + a stand-in for the CUDA
engine behind the same C ABI — and driven with many concurrent reconcile workers over HTTP
keep-alive, through LocalClient, and through the tool loop)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "agentcontrolplane_b200", "csrc")
SOURCES = [os.path.join(ROOT, "tests", "tsan", "main.cc"), os.path.join(ROOT, "tests", "tsan", "engine_stub.cc")] + [
    os.path.join(CSRC, f) for f in ("host/hostsim.cc", "host/llmclient.cc", "host/task.cc", "chat.cc", "tokenizer.cc", "safetensors.cc")]


def test_host_side_is_race_free_under_tsan(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "acp_host_tsan")
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
           "-I" + cuda_inc, *SOURCES, "-o", exe, "-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and ("tsan" in build.stderr.lower() or "sanitize" in build.stderr.lower()):
        pytest.skip("ThreadSanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="exitcode=66"))
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
    assert run.returncode == 0, (run.returncode, run.stdout[-2000:], run.stderr[-2000:])
    for leg in ("openai/http:", "local/abi:", "local/tool-loop:"):
        assert leg in run.stdout
