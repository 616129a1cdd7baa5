// Host-only stand-in for csrc/model.cu and the CUDA runtime, linked with the REAL csrc/engine.cc,
// c_api.cc, chat.cc, tokenizer.cc and host/*.cc in tests/test_sanitizers_cpu.py.  The scheduler,
// the paged-KV allocator, the shared prefix cache, cancellation and the C-ABI hand-off then run on
// a CPU under ThreadSanitizer, and — because this "model" reads K/V through the page tables the
// scheduler hands it — a wrong page mapping or a premature eviction changes the emitted tokens.
// TEST INFRASTRUCTURE: never linked into libacp_infer.so.
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "fake_model.h"
#include "model.h"
#include "safetensors.h"

// ---- the slice of the CUDA runtime engine.cc touches ----
extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "fake cuda"; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaSuccess; }
struct FakeEvent { std::chrono::steady_clock::time_point t; };
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new FakeEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (FakeEvent*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((FakeEvent*)e)->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(((FakeEvent*)b)->t - ((FakeEvent*)a)->t).count();
  return cudaSuccess;
}
}

namespace acp {

// ---- free functions engine.cc / hostsim.cc link against ----
// model_preset / model_config_from_hf / rope_inv_freq come from csrc/model_config.cc (pure C++, linked as is)
int attn_prefill_block_tokens(int heads, int kv_heads) { return (16 / (heads / kv_heads)) * 4; }
int attn_decode_chunks(int ctx_len) { return (ctx_len + 1023) / 1024; }
const NcclApi& nccl_api() { static NcclApi a; return a; }

// ---- Model ----
namespace {
struct FakeState { std::vector<uint64_t> kv; };   // [num_pages][KV_PAGE]
std::mutex g_mu;
std::unordered_map<const Model*, FakeState> g_state;
FakeState& state_of(const Model* m) { std::lock_guard<std::mutex> lk(g_mu); return g_state[m]; }
}  // namespace

Model::~Model() {
  free(h_ints_); free(h_sparams_); free(h_tokens_); free(h_logits_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_state.erase(this);
}

int Model::init(const ModelConfig& cfg, const ModelLimits& lim, int device, int tp_rank, int tp_size, NcclComm, Model*,
                const Checkpoint*) {
  cfg_ = cfg; lim_ = lim; device_ = device; tp_rank_ = tp_rank; tp_size_ = tp_size;
  ints_cap_ = (size_t)5 * lim_.max_tokens + (size_t)6 * lim_.max_batch + (size_t)lim_.max_batch * lim_.max_pages_per_seq + 64;
  h_ints_ = (int*)calloc(ints_cap_, sizeof(int));
  h_sparams_ = (SampleParams*)calloc((size_t)lim_.max_batch, sizeof(SampleParams));
  h_tokens_ = (int*)calloc((size_t)lim_.max_batch, sizeof(int));
  h_logits_ = (float*)calloc(16, sizeof(float));
  state_of(this).kv.assign((size_t)lim_.num_pages * KV_PAGE, 0xDEADDEADDEADDEADull);   // poison: never-written slots are visible
  return 0;
}

static inline size_t align4(size_t v) { return (v + 3) & ~(size_t)3; }
StepInput& Model::stage_begin(int T, int B, int n_blocks) {   // same carving as csrc/model.cu
  StepInput& s = stage_;
  s = StepInput();
  s.T = T; s.B = B; s.n_blocks = n_blocks;
  size_t off = 0;
  auto carve = [&](size_t n) { int* p = h_ints_ + off; off += align4(n); return p; };
  s.tok = carve(T); s.pos = carve(T); s.seq_of_row = carve(T);
  s.q_start = carve(B); s.q_len = carve(B); s.ctx_len = carve(B); s.sample_rows = carve(B);
  s.blk_seq = carve(n_blocks); s.blk_tok0 = carve(n_blocks);
  s.page_table = carve((size_t)B * lim_.max_pages_per_seq);
  s.tile_cum = carve((size_t)B + 1);
  ints_used_ = off;
  s.sample_params = h_sparams_;
  return s;
}

int Model::forward(const StepInput& in) {
  if (in.T <= 0 || in.T > lim_.max_tokens || in.B > lim_.max_batch || in.n_sample > lim_.max_batch) return -1;
  FakeState& st = state_of(this);
  // "append K/V": every row writes its slot through ITS sequence's page table
  for (int b = 0; b < in.B; ++b) {
    const int* pt = in.page_table + (size_t)b * lim_.max_pages_per_seq;
    for (int i = 0; i < in.q_len[b]; ++i) {
      const int row = in.q_start[b] + i, pos = in.pos[row];
      const int page = pt[pos / KV_PAGE];
      if (page <= 0 || page >= lim_.num_pages) return -1;                     // page 0 is reserved
      st.kv[(size_t)page * KV_PAGE + pos % KV_PAGE] = fakemodel::kv_value(in.tok[row], pos);
    }
  }
  // "attention + LM head": read the whole context back through the page table
  std::vector<uint64_t> ctx;
  for (int s = 0; s < in.n_sample; ++s) {
    const int row = in.sample_rows[s], b = in.seq_of_row[row];
    const int* pt = in.page_table + (size_t)b * lim_.max_pages_per_seq;
    const int n = in.pos[row] + 1;
    ctx.resize((size_t)n);
    for (int p = 0; p < n; ++p) ctx[(size_t)p] = st.kv[(size_t)pt[p / KV_PAGE] * KV_PAGE + p % KV_PAGE];
    h_tokens_[s] = fakemodel::next_token(ctx.data(), n);
  }
  launches_ += 8;
  // leave room for producers to race the step; a step of >= 400 rows is the checker's "blocker": it
  // holds the scheduler long enough for a whole burst to queue up behind it
  std::this_thread::sleep_for(std::chrono::microseconds(in.decode ? 150 : (in.T >= 400 ? 40000 : 400)));
  return 0;
}

int Model::sync() { return 0; }
std::string Model::profile_json() { return "{\"prefill\":{},\"decode\":{}}"; }
int Model::bench_exchange(int, int, float* us, int) { *us = 0.f; return -1; }

}  // namespace acp
