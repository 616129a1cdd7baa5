// Shared by fake_model.cc (the stand-in Model) and engine_sim_main.cc (the checker): the "model"
// is a hash chain over the K/V slots a sequence's page table points at, so a token it emits is
// correct iff every page mapping, shared prefix page and eviction decision of the REAL scheduler
// (csrc/engine.cc) was correct.
#pragma once
#include <stdint.h>
#include <vector>

namespace fakemodel {

inline uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// what the "attention" of a token at `pos` stores in its K/V slot
inline uint64_t kv_value(int tok, int pos) { return mix(((uint64_t)(uint32_t)tok << 32) | (uint32_t)pos); }
// next token given the K/V values of positions 0..ctx-1 (order sensitive)
inline int next_token(const uint64_t* kv, int ctx) {
  uint64_t h = 0x243F6A8885A308D3ull;
  for (int p = 0; p < ctx; ++p) h = mix(h ^ kv[p]);
  if (h % 23 == 0) return 128009;            // <|eot_id|>: sequences end at different lengths
  return 32 + (int)(h % 90);                 // printable ASCII byte tokens
}
// reference: the whole generation for a prompt, computed without any cache
inline std::vector<int> generate(const std::vector<int>& prompt, int max_tokens) {
  std::vector<uint64_t> kv;
  std::vector<int> toks = prompt, out;
  for (int p = 0; p < (int)prompt.size(); ++p) kv.push_back(kv_value(prompt[(size_t)p], p));
  for (int i = 0; i < max_tokens; ++i) {
    const int t = next_token(kv.data(), (int)kv.size());
    out.push_back(t);
    if (t == 128009 || t == 128001 || t == 128008) break;
    kv.push_back(kv_value(t, (int)toks.size()));
    toks.push_back(t);
  }
  return out;
}

}  // namespace fakemodel
