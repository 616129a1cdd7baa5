// common.cuh — sm_100a PTX wrappers shared by every kernel in libacp_infer.so.
//
// Everything here is inline PTX for Blackwell (mbarrier, TMA bulk-tensor copies,
// tcgen05 MMA / TMEM alloc / TMEM load).  No CUTLASS, no torch.  The bit layouts of
// the UMMA shared-memory descriptor and instruction descriptor follow the PTX ISA
// (cross-checked against cute/arch/mma_sm100_desc.hpp field comments).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define ACP_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------
// error handling (host)
// ---------------------------------------------------------------------------------
#define ACP_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      fprintf(stderr, "[acp_infer] CUDA error %s at %s:%d: %s\n", #expr, __FILE__,    \
              __LINE__, cudaGetErrorString(_e));                                      \
      return -5; /* ACP_ERR_CUDA */                                                   \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------------------------
// launch helper: cudaLaunchKernelEx with the programmatic-stream-serialization attribute
// (ACP_PDL=0 in the environment turns the attribute off for A/B measurements)
// ---------------------------------------------------------------------------------
#include <stdlib.h>
#include <utility>
static inline bool acp_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ACP_PDL"); v = (e && *e == '0') ? 0 : 1; }
  return v == 1;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t acp_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = acp_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------------------------
// bf16 helpers.  Rounding is round-to-nearest-even everywhere, which is what the
// oracle's bf16_round() does (oracle/llama_oracle.py).
// ---------------------------------------------------------------------------------
ACP_DEVINL float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
ACP_DEVINL float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
ACP_DEVINL float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }
ACP_DEVINL uint16_t f32_to_bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
ACP_DEVINL float bf16_round(float f) { return bf16_bits_to_f32(f32_to_bf16_bits(f)); }
ACP_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  return (uint32_t)f32_to_bf16_bits(lo) | ((uint32_t)f32_to_bf16_bits(hi) << 16);
}

// ---------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------
ACP_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ACP_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
ACP_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
ACP_DEVINL uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of a step calls pdl_launch_dependents() at
// its top (the next kernel's CTAs may then be scheduled as soon as SM resources free up, so its
// prologue — barrier init, TMEM alloc, descriptor prefetch, and for the GEMM the first weight
// tiles — overlaps this kernel's tail) and pdl_wait() before touching anything a previous kernel
// wrote.  Because every kernel waits, "my predecessor completed" holds transitively.
// ---------------------------------------------------------------------------------
ACP_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
ACP_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------
ACP_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
ACP_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ACP_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
ACP_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
ACP_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
ACP_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
ACP_DEVINL uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a
// hung GPU.  The bound is wall time: 20 s — far above any legitimate wait (a GEMM's MMA thread sits here while its
// producer waits for the previous kernel, which on a tensor-parallel shard can be an exchange waiting for a peer
// that is still loading its kernels on the first step) and below the 40 s of the cross-GPU waits in tp_comm.cu,
// so a kernel stuck on ONE GPU names itself before its peers give up on it: block, thread (= role), barrier.
ACP_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFu) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) {
        printf("[acp_infer] mbarrier wait timeout block=(%d,%d,%d) thread=%d bar=0x%x parity=%u\n", blockIdx.x,
               blockIdx.y, blockIdx.z, threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D tiled load, completes on an mbarrier.
// ---------------------------------------------------------------------------------
ACP_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// L2 cache-hint policies (same encodings CUTLASS uses for TMA::CacheHintSm90)
static constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
static constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
static constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

ACP_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// 1D bulk copy global -> shared (no tensor map), completes on an mbarrier.
ACP_DEVINL void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM load
// ---------------------------------------------------------------------------------
ACP_DEVINL void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ACP_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
ACP_DEVINL void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
ACP_DEVINL void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, one CTA.
ACP_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread is done.
ACP_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}
ACP_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets TMEM lane
// (lane_base + i), registers r[0..15] = columns col..col+15.
ACP_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, bf16.
//   tile in smem: rows of 128 bytes (64 bf16 of K), 8-row groups of 1024 bytes,
//   exactly what a TMA box {64, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes.
//   bits [0,14)  start address >> 4
//   bits [16,30) leading byte offset >> 4 (unused for swizzled K-major; canonical value 1)
//   bits [32,46) stride byte offset  >> 4 (1024 B between 8-row groups)
//   bits [46,48) descriptor version = 1 on sm_100
//   bits [61,64) layout type: 2 = SWIZZLE_128B
ACP_DEVINL uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
//   [4,6) c_format = 1 (F32); [7,10) a_format = 1 (BF16); [10,13) b_format = 1 (BF16)
//   [15] a_major = 0 (K); [16] b_major = 0 (K); [17,23) N >> 3; [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
