"""Host side under ThreadSanitizer (the reference runs `go test` without -race; the
engine boundary here is multi-producer, so the C++ mirror of the reconcile loop is built with
-fsanitize=thread in a host-only binary — tests/sanitizers/: real host sources + a stand-in for the CUDA
engine behind the same C ABI — and driven with many concurrent reconcile workers over HTTP
keep-alive, through LocalClient, and through the tool loop)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "agentcontrolplane_b200", "csrc")
SOURCES = [os.path.join(ROOT, "tests", "sanitizers", "main.cc"), os.path.join(ROOT, "tests", "sanitizers", "engine_stub.cc")] + [
    os.path.join(CSRC, f) for f in ("host/hostsim.cc", "host/llmclient.cc", "host/task.cc", "chat.cc", "tokenizer.cc", "safetensors.cc", "model_config.cc")]


SAN = os.path.join(ROOT, "tests", "sanitizers")
INC = ["-I" + SAN, "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I/usr/local/cuda/include"]
BUILDS = {
    "host_tsan": (["-fsanitize=thread"], SOURCES, ["-lpthread", "-ldl", "-rdynamic"]),
    "engine_sim": (["-fsanitize=thread"], [os.path.join(SAN, "engine_sim_main.cc"), os.path.join(SAN, "fake_model.cc")] + [
        os.path.join(CSRC, f) for f in ("engine.cc", "c_api.cc", "chat.cc", "tokenizer.cc", "safetensors.cc", "model_config.cc")], ["-lpthread"]),
    "edge_sim": (["-fsanitize=thread"], [os.path.join(SAN, "edge_sim_main.cc"), os.path.join(SAN, "fake_model.cc")] + [
        os.path.join(CSRC, f) for f in ("engine.cc", "c_api.cc", "chat.cc", "tokenizer.cc", "safetensors.cc", "model_config.cc")], ["-lpthread"]),
    "stack_sim": (["-fsanitize=thread"], [os.path.join(SAN, "stack_sim_main.cc"), os.path.join(SAN, "fake_model.cc")] + [
        os.path.join(CSRC, f) for f in ("engine.cc", "c_api.cc", "chat.cc", "tokenizer.cc", "safetensors.cc", "model_config.cc", "host/hostsim.cc",
                                        "host/llmclient.cc", "host/task.cc")], ["-lpthread", "-ldl", "-rdynamic"]),
    "fuzz": (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], [os.path.join(SAN, "fuzz_main.cc")] + [
        os.path.join(CSRC, f) for f in ("chat.cc", "tokenizer.cc", "safetensors.cc", "model_config.cc")], []),
}


@pytest.fixture(scope="module")
def binaries(tmp_path_factory):
    """The three sanitizer binaries, compiled concurrently (one g++ per harness)."""
    if not shutil.which("g++"):
        pytest.skip("no g++")
    out = tmp_path_factory.mktemp("sanitizers")
    procs = {}
    for name, (flags, srcs, libs) in BUILDS.items():
        exe = str(out / name)
        procs[name] = (exe, subprocess.Popen(["g++", "-std=c++17", "-O1", "-g", *flags, *INC, *srcs, "-o", exe, *libs],
                                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    built = {}
    for name, (exe, p) in procs.items():
        _, err = p.communicate(timeout=900)
        built[name] = (exe, p.returncode, err)
    return built


def _exe(binaries, name):
    exe, rc, err = binaries[name]
    # only a MISSING sanitizer runtime skips; any other build / link error of the harness is a failure
    if rc != 0 and re.search(r"cannot find -l(tsan|asan|ubsan)|lib(tsan|asan|ubsan)[^\n]*(not found|No such file)", err):
        pytest.skip("sanitizer runtime not available: " + err[-300:])
    assert rc == 0, err[-3000:]
    return exe


def test_host_side_is_race_free_under_tsan(binaries):
    exe = _exe(binaries, "host_tsan")
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="exitcode=66"))
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
    assert run.returncode == 0, (run.returncode, run.stdout[-2000:], run.stderr[-2000:])
    for leg in ("openai/http:", "local/abi:", "local/tool-loop:"):
        assert leg in run.stdout


def test_real_scheduler_over_a_fake_model_under_tsan(binaries):
    """csrc/engine.cc (scheduler, paged-KV allocator, shared prefix cache, cancellation, wait/poll)
    and c_api.cc are compiled UNCHANGED against tests/sanitizers/fake_model.cc — a stand-in Model
    whose emitted tokens are a hash chain over the K/V slots reached THROUGH the page tables the
    scheduler builds — and driven by 24 producer threads with shared prefixes, a KV pool that forces
    queueing and eviction, and cancellations.  Every response must equal the cache-free reference,
    no page may leak, and ThreadSanitizer must stay silent."""
    exe = _exe(binaries, "engine_sim")
    for extra in ({}, {"ACP_SIM_NO_CACHE": "1"}, {"ACP_SIM_REPLICAS": "1"}):
        run = subprocess.run([exe], capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, TSAN_OPTIONS="exitcode=66", **extra))
        assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
        assert run.returncode == 0 and "bad=0" in run.stdout, (run.returncode, run.stdout[-1500:], run.stderr[-3000:])


def test_scheduler_edge_cases_over_the_fake_model(binaries):
    """Page-boundary prompt lengths (1 ... 300 tokens, each a prefix of the next), prefill chunked into
    48-row steps with pages published chunk by chunk, sextuplets submitted together, max_tokens 1, a
    request that exceeds the context limit (400) and one that exactly fits, on a 63-page pool."""
    exe = _exe(binaries, "edge_sim")
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="exitcode=66"))
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
    assert run.returncode == 0 and "bad=0" in run.stdout, (run.returncode, run.stdout[-1500:], run.stderr[-3000:])


def test_whole_host_stack_tool_loop_under_tsan(binaries):
    """Reconcile workers -> LocalClient -> C ABI -> real scheduler -> fake Model, BASELINE config 3's
    shape (two tool schemas, scripted tool call, ToolCall CRs, fold-back, second LLM step): every
    FinalAnswer took two LLM steps, warm rounds are served from the shared prefix cache, no page
    leaks, ThreadSanitizer silent."""
    exe = _exe(binaries, "stack_sim")
    for extra in ({}, {"ACP_SIM_REPLICAS": "1"}):     # one engine; 4 data-parallel replicas behind one handle (sticky routing)
        run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, TSAN_OPTIONS="exitcode=66", **extra))
        assert "ThreadSanitizer" not in run.stderr, run.stderr[-6000:]
        assert run.returncode == 0 and run.stdout.count("round ") == 3, (run.returncode, run.stdout[-1500:], run.stderr[-3000:])


def test_untrusted_input_parsers_under_asan_ubsan(binaries, tmp_path):
    """Request bodies, completion text, tokenizer input and checkpoint / tokenizer.json files are
    untrusted bytes: a deterministic mutation fuzzer (tests/sanitizers/fuzz_main.cc) runs them through the
    real parsers under AddressSanitizer + UBSan."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ckpt_util import write_checkpoint
    from oracle.llama_oracle import LlamaConfig
    exe = _exe(binaries, "fuzz")
    ck = str(tmp_path / "ck")
    write_checkpoint(ck, LlamaConfig("f", hidden=128, layers=1, heads=1, kv_heads=1, ffn=128, vocab=256), 1)
    tok = os.path.join(ROOT, "tests", "golden", "llama3_style_tokenizer.json")
    for args in (["20000", tok, os.path.join(ck, "model.safetensors")], ["8000", ""]):
        run = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0 and "fuzz ok" in run.stdout, (run.returncode, run.stdout[-1000:], run.stderr[-6000:])
        assert "ERROR: AddressSanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-6000:]
