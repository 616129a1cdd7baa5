"""Host side under ThreadSanitizer (the reference runs `go test` without -race; the
engine boundary here is multi-producer, so the C++ mirror of the reconcile loop is built with
-fsanitize=thread in a host-only binary — tests/sanitizers/: real host sources + a stand-in for the CUDA
engine behind the same C ABI — and driven with many concurrent reconcile workers over HTTP
keep-alive, through LocalClient, and through the tool loop)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "agentcontrolplane_b200", "csrc")
SOURCES = [os.path.join(ROOT, "tests", "sanitizers", "main.cc"), os.path.join(ROOT, "tests", "sanitizers", "engine_stub.cc")] + [
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


def test_real_scheduler_over_a_fake_model_under_tsan(tmp_path):
    """csrc/engine.cc (scheduler, paged-KV allocator, shared prefix cache, cancellation, wait/poll)
    and c_api.cc are compiled UNCHANGED against tests/sanitizers/fake_model.cc — a stand-in Model
    whose emitted tokens are a hash chain over the K/V slots reached THROUGH the page tables the
    scheduler builds — and driven by 24 producer threads with shared prefixes, a KV pool that forces
    queueing and eviction, and cancellations.  Every response must equal the cache-free reference,
    no page may leak, and ThreadSanitizer must stay silent."""
    if not shutil.which("g++"):
        pytest.skip("no g++")
    san = os.path.join(ROOT, "tests", "sanitizers")
    exe = str(tmp_path / "acp_engine_sim")
    srcs = [os.path.join(san, "engine_sim_main.cc"), os.path.join(san, "fake_model.cc")] + [
        os.path.join(CSRC, f) for f in ("engine.cc", "c_api.cc", "chat.cc", "tokenizer.cc", "safetensors.cc")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-I" + san, "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
           "-I/usr/local/cuda/include", *srcs, "-o", exe, "-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and ("tsan" in build.stderr.lower() or "sanitize" in build.stderr.lower()):
        pytest.skip("ThreadSanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-3000:]
    for _ in range(2):
        run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="exitcode=66"))
        assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
        assert run.returncode == 0 and "bad=0" in run.stdout, (run.returncode, run.stdout[-1500:], run.stderr[-3000:])


def test_untrusted_input_parsers_under_asan_ubsan(tmp_path):
    """Request bodies, completion text, tokenizer input and checkpoint / tokenizer.json files are
    untrusted bytes: a deterministic mutation fuzzer (tests/sanitizers/fuzz_main.cc) runs them through the
    real parsers under AddressSanitizer + UBSan."""
    if not shutil.which("g++"):
        pytest.skip("no g++")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ckpt_util import write_checkpoint
    from oracle.llama_oracle import LlamaConfig
    exe = str(tmp_path / "acp_host_fuzz")
    srcs = [os.path.join(ROOT, "tests", "sanitizers", "fuzz_main.cc")] + [os.path.join(CSRC, f) for f in ("chat.cc", "tokenizer.cc", "safetensors.cc")]
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I/usr/local/cuda/include", *srcs, "-o", exe]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "sanitize" in build.stderr.lower():
        pytest.skip("sanitizer runtimes not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-3000:]
    ck = str(tmp_path / "ck")
    write_checkpoint(ck, LlamaConfig("f", hidden=128, layers=1, heads=1, kv_heads=1, ffn=128, vocab=256), 1)
    tok = os.path.join(ROOT, "tests", "golden", "llama3_style_tokenizer.json")
    for args in (["30000", tok, os.path.join(ck, "model.safetensors")], ["15000", ""]):
        run = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0 and "fuzz ok" in run.stdout, (run.returncode, run.stdout[-1000:], run.stderr[-6000:])
        assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-6000:]
