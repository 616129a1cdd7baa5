// attention.h — host interface of the paged-KV attention kernels (attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kernels.cuh"

namespace acp {

// Decode attention, flat-scheduled: the (sequence, kv head, 64-token tile) space is flattened
// and cut into `n_ctas` equal contiguous chunks, one per persistent CTA, so every SM streams the
// same number of K/V bytes whatever the mix of context lengths.  A (sequence, kv head) item cut
// by a chunk boundary is finished by attn_merge_kernel (fixed piece order => deterministic).
constexpr int ATTN_MAX_PIECES = 3;
struct AttnDecodeArgs {
  const __nv_bfloat16* q;   // [B][heads*128] (row b = sequence b)
  __nv_bfloat16* out;       // [B][heads*128]
  const int* ctx_len;       // [B] keys visible to the query of sequence b (incl. its own token)
  const int* tile_cum;      // [B+1] prefix sum of ceil(ctx_len/64)
  const int* page_table;    // [B][max_pages]
  int max_pages;
  int heads, kv_heads;
  int num_seqs;
  int total_tiles;          // tile_cum[B] * kv_heads
  int n_ctas;               // grid size (flat schedule)
  int per_item;             // 1: one CTA per (sequence, kv head), no workspace / merge (short, uniform items)
  float scale;              // 1/sqrt(128)
  float* ws;                // partials [(item*3 + piece)*4 + warp][G rows][130] fp32
};

struct AttnPrefillArgs {
  const __nv_bfloat16* q;
  __nv_bfloat16* out;
  const int* blk_seq;       // [num_blocks] sequence of each query block
  const int* blk_tok0;      // [num_blocks] first new-token index of the block
  const int* q_start;       // [B] first row of the sequence's new tokens
  const int* q_len;         // [B] number of new tokens
  const int* ctx_len;       // [B] total tokens after this step (cached + new)
  const int* page_table;
  int max_pages;
  int heads, kv_heads;
  float scale;
};

int attn_setup_attributes();
int attn_make_kv_map(CUtensorMap* out, const void* base, uint64_t num_pages, int kv_heads);
// picks n_ctas (<= 2 per SM, chunk >= half the longest item so an item spans <= 3 CTAs)
int attn_decode_plan(int total_tiles, int max_item_tiles);
size_t attn_decode_ws_floats(int max_batch, int heads, int kv_heads);
int launch_attn_decode(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnDecodeArgs& a,
                       cudaStream_t s);
int attn_prefill_block_tokens(int heads, int kv_heads);
int launch_attn_prefill(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnPrefillArgs& a,
                        int num_blocks, cudaStream_t s);

}  // namespace acp
