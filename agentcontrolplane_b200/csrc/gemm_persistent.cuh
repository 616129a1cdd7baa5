// gemm_persistent.cuh — persistent variant of the tcgen05 GEMM for PREFILL steps (N > 256 token
// rows): one CTA per SM loops over (m-tile, 256-token n-tile) tiles; the accumulator is
// DOUBLE-BUFFERED in TMEM (2 x 256 columns = all 512), so the epilogue warps drain tile i
// (tcgen05.ld -> bf16 -> global, or the fused SwiGLU) while the MMA warp already accumulates
// tile i+1 and the TMA producer streams tiles i+1, i+2 without ever draining the smem ring.
//
//   warp 0      TMA producer over all tiles of this CTA (ring of 4 x 48 KiB stages)
//   warp 1      TMEM alloc (512 columns) + tcgen05.mma issuer, ping-pongs the two accumulators
//   warps 2..5  epilogue, lane quarter = warp % 4
// Barriers: full/empty per smem stage, tmem_full/tmem_empty per accumulator.
// Tile order: n-tile fastest (CTAs running together share a weight tile; X stays in L2).
#pragma once
#include "gemm_tcgen05.cuh"

namespace acp {

constexpr int PGEMM_BN = 256;
constexpr int PGEMM_STAGES = 4;
constexpr int PGEMM_STAGE_BYTES = GEMM_BM * GEMM_BK * 2 + PGEMM_BN * GEMM_BK * 2;  // 48 KiB
constexpr int PGEMM_SMEM = PGEMM_STAGES * PGEMM_STAGE_BYTES + 1024 + 256;

struct PTile { int m_tile, n0; };
__device__ __forceinline__ PTile ptile_decode(int t, int n_tiles) {
  PTile p;
  p.m_tile = t / n_tiles; p.n0 = (t % n_tiles) * 256;
  return p;
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_wx_persistent_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                          GemmArgs args, int m_tiles, int n_tiles) {
  constexpr int STAGES = PGEMM_STAGES;
  constexpr int ABYTES = GEMM_BM * GEMM_BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * PGEMM_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = (args.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_tiles * n_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      pdl_wait();  // activations come from the previous kernel
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const PTile pt = ptile_decode(t, n_tiles);
        const int m_tile = pt.m_tile, n0 = pt.n0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * PGEMM_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], PGEMM_STAGE_BYTES);
          tma_load_2d(a_dst, &tmap_w, &full_bar[s], 0, (m_tile * nkb + kb) * GEMM_BM, kEvictNormal);
          tma_load_2d(a_dst + ABYTES, &tmap_x, &full_bar[s], kb * GEMM_BK, n0, kEvictLast);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, PGEMM_BN);
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
        const int buf = i & 1;
        mbar_wait(&tmem_empty[buf], (uint32_t)((i >> 1) & 1) ^ 1u);  // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * PGEMM_BN);
        const PTile pt = ptile_decode(t, n_tiles);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * PGEMM_STAGE_BYTES);
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + ABYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16(acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;
    pdl_wait();
    const int n_valid = args.n_dev ? *args.n_dev : args.N;
    int i = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      const PTile pt = ptile_decode(t, n_tiles);
      const int m_tile = pt.m_tile, n0 = pt.n0;
      const int m = m_tile * GEMM_BM + q * 32 + lane;
      mbar_wait(&tmem_full[buf], (uint32_t)((i >> 1) & 1));
      tcgen05_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * PGEMM_BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < PGEMM_BN; c += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(acc + (uint32_t)c, r);
        tmem_ld_wait();
        __nv_bfloat16* out = (__nv_bfloat16*)args.out;
        if constexpr (EPI == EPI_BF16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            if (n < n_valid && m < args.M)
              out[(size_t)n * args.ld + m] = __float2bfloat16_rn(__uint_as_float(r[j]));
          }
        } else {  // EPI_SWIGLU
          swiglu_store16(out, r, n0 + c, n_valid, m, args.M, args.ld, lane);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);  // 4 warps => accumulator free again
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace acp
