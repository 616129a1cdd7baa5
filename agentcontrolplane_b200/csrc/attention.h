// attention.h — host interface of the paged-KV attention kernels (attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kernels.cuh"

namespace acp {

struct AttnDecodeArgs {
  const __nv_bfloat16* q;   // [rows][heads*128]
  __nv_bfloat16* out;       // [rows][heads*128]
  const int* ctx_len;       // [B] keys visible to the query of sequence b (incl. its own token)
  const int* q_rows;        // [B] row of sequence b's query in q/out (nullptr => b)
  const int* page_table;    // [B][max_pages]
  int max_pages;
  int heads, kv_heads;
  float scale;              // 1/sqrt(128)
  int split_tokens;         // KV tokens per split (multiple of 64)
  int max_splits;           // capacity of the split workspace
  float* ws_o;              // [B][heads][max_splits][128]
  float* ws_m;              // [B][heads][max_splits]
  float* ws_l;
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
int launch_attn_decode(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnDecodeArgs& a,
                       int num_seqs, int max_ctx, cudaStream_t s);
int attn_prefill_block_tokens(int heads, int kv_heads);
int launch_attn_prefill(const CUtensorMap& tm_k, const CUtensorMap& tm_v, const AttnPrefillArgs& a,
                        int num_blocks, cudaStream_t s);

}  // namespace acp
