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
        } else if constexpr (EPI == EPI_F32) {   // one fp32 plane (row-parallel shards: rounded once AFTER the exchange)
          float* outf = (float*)args.out;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            if (n < n_valid && m < args.M) outf[(size_t)n * args.ld + m] = __uint_as_float(r[j]);
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


// =====================================================================================================
// 2-CTA variant (tcgen05 cta_group::2): a CLUSTER of two CTAs on the two SMs of a TPC computes a
// 256 (weight rows = two m-tiles) x 256 (token rows) tile with M = 256 UMMAs issued by the leader CTA.
// Each CTA stages its own 128 weight rows (A) and only HALF of the token rows (B: 128 of the 256):
// 32 KiB per k-block and SM instead of 48 KiB, which is what the 1-CTA kernel is limited by (every SM
// ingests 3 MB of operands per 268 MFLOP tile: ~61 B/clk at the power-capped clocks of this pool).
//
//   per CTA   warp 0  TMA producer (cp.async.bulk.tensor ... cta_group::2: bytes complete on the LEADER's
//                     full barrier), warp 1 TMEM alloc (cta_group::2, 512 columns, both CTAs) — and in the
//                     leader the MMA issuer —, warps 2..5 epilogue of this CTA's 128 accumulator rows
//   barriers  full[s]        leader only: 1 arrival (leader's expect_tx of BOTH CTAs' 64 KiB)
//             empty[s]       in each CTA: released by the leader's tcgen05.commit multicast (mask 0b11)
//             tmem_full[b]   in each CTA: leader's commit multicast when a tile's accumulator is complete
//             tmem_empty[b]  leader only: 8 arrivals = 4 epilogue warps of each CTA (the peer arrives remotely)
// Same arithmetic as the 1-CTA kernel (K accumulated in order in one fp32 accumulator per output element).
// =====================================================================================================
constexpr int P2_STAGES = 6;
constexpr int P2_STAGE_BYTES = 2 * GEMM_BM * GEMM_BK * 2;   // A 16 KiB + B half 16 KiB
constexpr int P2_SMEM = P2_STAGES * P2_STAGE_BYTES + 1024 + 256;

ACP_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
ACP_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ACP_DEVINL void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {  // whole warp, the same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ACP_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// executed by BOTH CTAs; the transaction bytes update the LEADER's mbarrier (peer bit of the address cleared)
ACP_DEVINL void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t hint) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(bar_leader), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
ACP_DEVINL void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs once the pair's previously issued MMAs are done
ACP_DEVINL void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
ACP_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {   // arrive on `bar` of CTA `cta` of this cluster
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_wx_persistent2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x128,
                           GemmArgs args, int m_pairs, int n_tiles) {
  constexpr int STAGES = P2_STAGES;
  constexpr int ABYTES = GEMM_BM * GEMM_BK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * P2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int nkb = (args.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_pairs * n_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x128);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers initialised and TMEM allocated before anything crosses the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      pdl_wait();
      int s = 0;
      uint32_t ph = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters) {
        const int m_tile = (t / n_tiles) * 2 + (int)rank, n0 = (t % n_tiles) * PGEMM_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * P2_STAGE_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * P2_STAGE_BYTES);
          tma_load_2d_2sm(a_dst, &tmap_w, &full_bar[s], 0, (m_tile * nkb + kb) * GEMM_BM, kEvictNormal);
          tma_load_2d_2sm(a_dst + ABYTES, &tmap_x128, &full_bar[s], kb * GEMM_BK, n0 + (int)rank * 128, kEvictLast);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, PGEMM_BN);
      int s = 0;
      uint32_t ph = 0;
      int i = 0;
      for (int t = cluster_id; t < num_tiles; t += n_clusters, ++i) {
        const int buf = i & 1;
        mbar_wait(&tmem_empty[buf], (uint32_t)((i >> 1) & 1) ^ 1u);   // both CTAs drained this accumulator
        tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * PGEMM_BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * P2_STAGE_BYTES);
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + ABYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            umma_bf16_2sm(acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[s]);     // slot free in BOTH CTAs
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2sm(&tmem_full[buf]);     // accumulator complete in BOTH CTAs
      }
    }
  } else {
    // ===== epilogue warps 2..5 (both CTAs: this CTA's 128 rows x 256 columns) =====
    const int q = warp & 3;
    pdl_wait();
    const int n_valid = args.n_dev ? *args.n_dev : args.N;
    int i = 0;
    for (int t = cluster_id; t < num_tiles; t += n_clusters, ++i) {
      const int buf = i & 1;
      const int m_tile = (t / n_tiles) * 2 + (int)rank, n0 = (t % n_tiles) * PGEMM_BN;
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
        } else if constexpr (EPI == EPI_F32) {   // one fp32 plane (row-parallel shards: rounded once AFTER the exchange)
          float* outf = (float*)args.out;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + c + j;
            if (n < n_valid && m < args.M) outf[(size_t)n * args.ld + m] = __uint_as_float(r[j]);
          }
        } else {  // EPI_SWIGLU
          swiglu_store16(out, r, n0 + c, n_valid, m, args.M, args.ld, lane);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[buf]);
        else mbar_arrive_cluster(&tmem_empty[buf], 0);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();              // re-converge the role branches before the .aligned cluster barrier
  cluster_sync_all();           // nobody leaves (or frees TMEM) while the peer can still touch this CTA
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace acp
