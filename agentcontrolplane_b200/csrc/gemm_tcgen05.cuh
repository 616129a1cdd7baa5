// gemm_tcgen05.cuh — the one dense-contraction kernel of the decode engine.
//
//   out[n][m] = sum_k W[m][k] * X[n][k]          (W: [M][K] bf16 stored as contiguous 128x64 tiles,
//                                                 X: [N][K] bf16 row-major; both K-major operands)
//
// "Swap-AB" orientation: the WEIGHT matrix is the UMMA A operand (M = 128 output features
// per tile = 128 TMEM lanes) and the ACTIVATIONS are the B operand (N = sequences/tokens,
// 16..256 per tile = TMEM columns).  Decode batches (N = 16..256 live sequences) therefore
// fill a full-rate M=128 MMA while streaming every weight byte exactly once from HBM; prefill
// uses the same kernel with N tiles of 256 tokens.
//
// Pipeline (warp-specialised, one CTA per (m-tile, n-tile, k-split)):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d W-tile {64 x 128} + X-tile {64 x BN},
//               128-byte swizzle, into a STAGES-deep shared-memory ring (full/empty mbarriers)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=16 per 64-wide k-block),
//               tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld 32x32b (lane quarter = warp_id % 4) -> registers ->
//               bf16 / fp32-partial / fused arg-max store
//
// Split-K: grid.z splits the k-block range; partials are written as fp32 [split][N][M] and are
// summed IN FIXED ORDER by the consumer kernel (rope / add+rmsnorm / swiglu), which keeps the
// result bit-reproducible run to run (no atomics).
//
// Reference boundary: this replaces the remote model behind
// acp/internal/llmclient/langchaingo_client.go:102 (GenerateContent); see DESIGN.md.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace acp {

constexpr int GEMM_BM = 128;  // UMMA M (weight rows per tile)
constexpr int GEMM_BK = 64;   // bf16 elements per k-block (= one 128-byte swizzle row)
constexpr int GEMM_THREADS = 192;
// grouped mode: layout of the device-side `group_ranges` array (written by moe_dispatch_kernel):
//   [2g], [2g+1]                    first row and row count of group g            (g < GEMM_GROUP_MAX)
//   [GEMM_GROUP_TILES]              number of entries of the flat tile list
//   [GEMM_GROUP_TILES + 2 + 2i ..]  tile i = {group, first row inside the group}
constexpr int GEMM_GROUP_MAX = 16;
constexpr int GEMM_GROUP_TILES = 2 * GEMM_GROUP_MAX;

enum GemmEpi : int {
  EPI_BF16 = 0,     // out_bf16[n*ld + m]
  EPI_F32 = 1,      // out_f32[(split*n_cap + n)*ld + m]      (split-K partials, logits)
  EPI_ARGMAX = 2,   // per (n, m-tile): max value + lowest index; optional fp32 logits
  EPI_SWIGLU = 3,   // rows interleaved (gate_j, up_j): out_bf16[n*ld + m/2] = bf16(bf16(silu(g))*u)
};

struct GemmArgs {
  int M;         // rows of W (output features)
  int N;         // valid activation rows
  int K;
  int splits;    // == gridDim.z
  int ld;        // output leading dimension (elements)
  int n_cap;     // row capacity of one split-K partial plane
  void* out;     // bf16* or float* (may be null for EPI_ARGMAX)
  float* amax_val;   // EPI_ARGMAX: [N][m_tiles]
  int* amax_idx;     // EPI_ARGMAX: [N][m_tiles]
  const int* n_dev;  // optional device scalar overriding N (CUDA-graph replay with varying batch)
  // GROUPED mode (mixture of experts): group g's weights are m-tiles [g*gridDim.y, (g+1)*gridDim.y) of the
  // (concatenated) weight tensor, its activation / output rows are [ranges[2g], ranges[2g] + ranges[2g+1]).
  // blockIdx.x indexes a flat tile list {group, first row inside the group} (layout: GEMM_GROUP_TILES above)
  // — all written on the device by moe_dispatch_kernel, so row counts never visit the host and
  // no CTA is launched for rows that do not exist (beyond the <= E slack of the grid bound).  splits must be 1.
  const int* group_ranges;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN <= 64) ? 4 : (BN == 128 ? 5 : 4);
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;  // 16 KiB
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kShallowStages = BN == 256 ? 2 : (BN == 128 ? 3 : kStages);   // two CTAs per SM
  static constexpr int smem_bytes(int stages) { return stages * kStageBytes + 1024 + 256; }
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

// SwiGLU epilogue of 16 accumulator columns.  W rows are interleaved (gate_j, up_j), so the pair
// of one output sits in adjacent TMEM lanes = adjacent threads.  The two threads split the 16
// tokens: the even lane (gate) finishes the even tokens, the odd lane (up) the odd ones, so one
// exchange serves two outputs and no lane idles through expf.
__device__ __forceinline__ void swiglu_store16(__nv_bfloat16* out, const uint32_t (&r)[16], int n_base, int n_valid,
                                               int m, int M, int ld, int lane) {
  const bool odd = lane & 1;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float mine0 = bf16_round(__uint_as_float(r[j]));       // the GEMM's bf16 rounding point
    const float mine1 = bf16_round(__uint_as_float(r[j + 1]));
    const float got = __shfl_xor_sync(0xffffffffu, odd ? mine0 : mine1, 1);
    const float g = odd ? got : mine0;     // odd lane: partner's gate of token j+1
    const float u = odd ? mine1 : got;     // even lane: partner's up of token j
    const int n = n_base + j + (odd ? 1 : 0);
    if (n < n_valid && m < M) {
      const float act = bf16_round(g / (1.0f + expf(-g)));
      out[(size_t)n * ld + (m >> 1)] = __float2bfloat16_rn(act * u);
    }
  }
}

// STAGES: depth of the TMA ring.  The wide tiles (BN 128 / 256) come in two flavours: a deep ring
// with one CTA per SM, or a shallow ring (3 / 2 stages, 96 KiB) with TWO CTAs per SM — the same
// bytes in flight per SM, but 296 CTA slots, so a split-K grid of 150-296 CTAs runs as ONE wave.
template <int BN, int EPI, int STAGES_ = GemmCfg<BN>::kStages>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_wx_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
               GemmArgs args) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = STAGES_;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atom (8 rows x 128 B).
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // blockIdx.x = n-tile (fastest): the CTAs that run together share ONE weight tile and sweep the
  // activation rows, so W streams from HBM once and X (<= 67 MB at 8192 x 4096) is served by L2.
  int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int m_tiles = gridDim.y;
  const bool grouped = args.group_ranges != nullptr;
  int m_tile = blockIdx.y;            // weight m-tile (global index over all groups)
  int split = blockIdx.z;
  int row_off = 0, group_rows = 0;
  pdl_launch_dependents();
  if (grouped) {
    pdl_wait();                        // the ranges come from the previous kernel (moe_dispatch)
    if ((int)blockIdx.x >= args.group_ranges[GEMM_GROUP_TILES]) return;   // whole CTA, before any barrier / TMEM allocation
    const int group = args.group_ranges[GEMM_GROUP_TILES + 2 + 2 * blockIdx.x];
    n0 = args.group_ranges[GEMM_GROUP_TILES + 2 + 2 * blockIdx.x + 1];
    row_off = args.group_ranges[2 * group];
    group_rows = args.group_ranges[2 * group + 1];
    m_tile = group * m_tiles + (int)blockIdx.y;
    split = 0;
  }
  const int nkb_total = (args.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_begin = (int)(((long long)nkb_total * split) / args.splits);
  const int kb_end = (int)(((long long)nkb_total * (split + 1)) / args.splits);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      // Phase A: the first ring-full of WEIGHT tiles does not depend on the previous kernel, so
      // it is issued before the grid-dependency wait (hides the HBM ramp under that kernel's tail).
      const int pre = nkb < STAGES ? nkb : STAGES;
      for (int kb = 0; kb < pre; ++kb) {
        mbar_arrive_expect_tx(&full_bar[kb], Cfg::kStageBytes);
        tma_load_2d(smem + kb * Cfg::kStageBytes, &tmap_w, &full_bar[kb], 0,
                    (m_tile * nkb_total + kb_begin + kb) * GEMM_BM, kEvictFirst);  // one contiguous 16 KiB tile
      }
      pdl_wait();  // activations come from the previous kernel
      for (int kb = 0; kb < pre; ++kb)
        tma_load_2d(smem + kb * Cfg::kStageBytes + Cfg::kABytes, &tmap_x, &full_bar[kb],
                    (kb_begin + kb) * GEMM_BK, row_off + n0, kEvictLast);
      int s = pre % STAGES;
      uint32_t ph = (pre == STAGES) ? 1u : 0u;
      for (int kb = pre; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + Cfg::kABytes;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
        const int kcoord = (kb_begin + kb) * GEMM_BK;
        // weights are read once per step: evict-first; activations are re-read by every m-tile
        tma_load_2d(a_dst, &tmap_w, &full_bar[s], 0, (m_tile * nkb_total + kb_begin + kb) * GEMM_BM, kEvictFirst);
        tma_load_2d(b_dst, &tmap_x, &full_bar[s], kcoord, row_off + n0, kEvictLast);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    pdl_wait();
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
        const uint64_t a_desc = umma_desc_k_sw128(a_addr);
        const uint64_t b_desc = umma_desc_k_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advancing 16 bf16 (32 bytes) along K inside the swizzle atom = +2 in the
          // (addr >> 4) start-address field
          umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                    (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // slot reusable once these MMAs have read smem
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    pdl_wait();  // outputs may still be read by the previous kernels until they complete
    int n_valid = grouped ? group_rows : (args.n_dev ? *args.n_dev : args.N);
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      tcgen05_fence_after();
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 16) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if constexpr (EPI == EPI_BF16) {
        __nv_bfloat16* out = (__nv_bfloat16*)args.out;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c + j;
          if (n < n_valid && m < args.M)
            out[(size_t)(row_off + n) * args.ld + m] = __float2bfloat16_rn(__uint_as_float(r[j]));
        }
      } else if constexpr (EPI == EPI_F32) {
        float* out = (float*)args.out;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c + j;
          if (n < n_valid && m < args.M)
            out[((size_t)split * args.n_cap + n) * args.ld + m] = __uint_as_float(r[j]);
        }
      } else if constexpr (EPI == EPI_SWIGLU) {
        // W rows are stored interleaved: even row = gate_j, odd row = up_j (j = row / 2), so the
        // pair sits in adjacent TMEM lanes = adjacent threads of this warp.
        swiglu_store16((__nv_bfloat16*)args.out + (size_t)row_off * args.ld, r, n0 + c, n_valid, m, args.M, args.ld, lane);
      } else {  // EPI_ARGMAX
        float* out = (float*)args.out;
        float* red_v = (float*)(smem);  // ring buffers are idle now: reuse [4][16] floats + ints
        int* red_i = (int*)(smem + 4 * 16 * sizeof(float));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c + j;
          float v = (m < args.M) ? __uint_as_float(r[j]) : -INFINITY;
          int idx = m;
          if (out != nullptr && n < n_valid && m < args.M) out[(size_t)n * args.ld + m] = v;
          // warp arg-max with lowest-index tie-break (deterministic)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
          }
          if (lane == 0) { red_v[q * 16 + j] = v; red_i[q * 16 + j] = idx; }
        }
        // named barrier over the 4 epilogue warps (128 threads), id 1
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q == 2 && lane < 16) {  // warp 2 finishes: 4 candidates per column
          float v = red_v[lane];
          int idx = red_i[lane];
#pragma unroll
          for (int w = 1; w < 4; ++w) {
            float ov = red_v[w * 16 + lane];
            int oi = red_i[w * 16 + lane];
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
          }
          const int n = n0 + c + lane;
          if (n < n_valid) {
            args.amax_val[(size_t)n * m_tiles + m_tile] = v;
            args.amax_idx[(size_t)n * m_tiles + m_tile] = idx;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace acp
