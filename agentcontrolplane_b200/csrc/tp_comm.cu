// tp_comm.cu — see tp_comm.h.
#include "tp_comm.h"
#include "common.cuh"

namespace acp {

namespace {

ACP_DEVINL void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
ACP_DEVINL int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

ACP_DEVINL void tp_signal(const TpPeers& P, int lane, int epoch) {  // lanes 0..size-1 of one warp
  // st.release.sys is itself a system-scope fence followed by the store (one MEMBAR.SYS, not two)
  if (lane < P.size) st_release_sys(P.flags[lane] + P.rank, epoch);
}
ACP_DEVINL void tp_wait_all(const TpPeers& P, int lane, int epoch) {  // lanes 0..size-1 of one warp
  if (lane < P.size) {
    unsigned spins = 0;
    uint64_t t0 = 0;
    while (ld_acquire_sys(P.flags[P.rank] + lane) < epoch) {
      if ((++spins & 0x3FFu) == 0) {   // wall-time bound (40 s): a stuck peer traps its own kernel first (20 s, common.cuh)
        const uint64_t now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 40000000000ull) {
          printf("[acp_infer] tp wait timeout rank=%d waiting for %d epoch=%d\n", P.rank, lane, epoch);
          __trap();
        }
      }
    }
  }
}

ACP_DEVINL float4 ld_f4_volatile(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
ACP_DEVINL float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One kernel: [signal "my partial sums are complete"] -> [wait for every rank] -> pull + reduce +
// residual + RMSNorm + push for the rows this rank owns -> [last CTA signals "my rows are pushed"]
// -> [CTA 0 waits for every rank's "pushed"], so grid completion == exchange complete.
// Epochs: `epoch` = ready, `epoch + 1` = done; flags only grow.
// NP = number of ranks (compile time: no redundant pulls, every loop fully unrolled)
template <int NP>
__global__ void __launch_bounds__(1024)
tp_reduce_norm_kernel(TpPeers P, int T, int hidden, const __nv_bfloat16* __restrict__ gain, float eps,
                      int epoch, int* done_ctr, int push_x, int diag) {
  extern __shared__ float row[];
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();  // this rank's GEMM (+ split-K reduce) is complete
  if (blockIdx.x == 0 && threadIdx.x < 32) tp_signal(P, threadIdx.x, epoch);
  if (threadIdx.x < 32) tp_wait_all(P, threadIdx.x, epoch);
  __syncthreads();
  const int rows_per = (T + P.size - 1) / P.size;
  const int t = P.rank * rows_per + blockIdx.x;
  if (t < T && (int)blockIdx.x < rows_per) {
    const size_t off = (size_t)t * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x * 4; i < hidden; i += blockDim.x * 4) {
      // pull the partial sums of every rank (all loads in flight together), add in rank order
      float4 q[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int src = (diag & 4) ? P.rank : p;   // diag 4: no remote pulls
        q[p] = (diag & 1) ? ld_f4_volatile(P.ar[src] + off + i) : ld_f4(P.ar[src] + off + i);
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NP; ++p) { acc.x += q[p].x; acc.y += q[p].y; acc.z += q[p].z; acc.w += q[p].w; }
      const uint2 raw = *reinterpret_cast<const uint2*>(P.x[P.rank] + off + i);
      float v[4] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y)};
      v[0] = bf16_round(v[0] + bf16_round(acc.x));
      v[1] = bf16_round(v[1] + bf16_round(acc.y));
      v[2] = bf16_round(v[2] + bf16_round(acc.z));
      v[3] = bf16_round(v[3] + bf16_round(acc.w));
      uint2 packed;
      packed.x = pack_bf16x2(v[0], v[1]);
      packed.y = pack_bf16x2(v[2], v[3]);
      // the residual row is only ever read again by its owner (the next exchange), so it stays
      // local; the last layer pushes it everywhere because the sampler gathers arbitrary rows
      if (push_x) {
#pragma unroll
        for (int p = 0; p < NP; ++p) *reinterpret_cast<uint2*>(P.x[p] + off + i) = packed;
      } else {
        *reinterpret_cast<uint2*>(P.x[P.rank] + off + i) = packed;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { row[i + j] = v[j]; ss += v[j] * v[j]; }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    const float rstd = 1.0f / sqrtf(tot / (float)hidden + eps);
    for (int i = threadIdx.x * 4; i < hidden; i += blockDim.x * 4) {
      const uint2 graw = *reinterpret_cast<const uint2*>(gain + i);
      const float g[4] = {bf16_lo(graw.x), bf16_hi(graw.x), bf16_lo(graw.y), bf16_hi(graw.y)};
      uint2 packed;
      packed.x = pack_bf16x2(g[0] * bf16_round(row[i] * rstd), g[1] * bf16_round(row[i + 1] * rstd));
      packed.y = pack_bf16x2(g[2] * bf16_round(row[i + 2] * rstd), g[3] * bf16_round(row[i + 3] * rstd));
#pragma unroll
      for (int p = 0; p < NP; ++p)
        if (!(diag & 2) || p == P.rank) *reinterpret_cast<uint2*>(P.xn[p] + off + i) = packed;  // diag 2: no remote pushes
    }
  }
  // The last CTA of this rank to finish its pushes tells every rank "my rows have landed".  Each
  // CTA publishes its stores at DEVICE scope (fence + counter atomic); only the last one pays for
  // the system-scope release — a MEMBAR.SYS per CTA serialises in L2 and made the exchange cost
  // 0.1 us per row.  Causality: pushes -> fence.gpu/atomic -> last CTA's atomic -> st.release.sys
  // -> peer's ld.acquire.sys, every link morally strong.
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence();
    const int old = atomicAdd(done_ctr, 1);
    is_last = (old == (int)gridDim.x - 1);
    if (is_last) { __threadfence(); *done_ctr = 0; }
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) tp_signal(P, threadIdx.x, epoch + 1);
  // CTA 0 closes the exchange: the grid (and with it the dependent GEMM's griddepcontrol.wait)
  // completes only when every rank's rows have landed in this rank's buffers.
  if (blockIdx.x == 0 && threadIdx.x < 32) tp_wait_all(P, threadIdx.x, epoch + 1);
}

}  // namespace

int launch_tp_reduce_norm(const TpPeers& p, int T, int hidden, const __nv_bfloat16* gain, float eps,
                          int epoch, int* done_ctr, bool push_x, cudaStream_t s, int diag) {
  if (T <= 0) return 0;
  const int rows_per = (T + p.size - 1) / p.size;
  int threads = ((hidden / 4 + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  cudaError_t e;
  const size_t smem = hidden * sizeof(float);
  const int px = push_x ? 1 : 0;
  switch (p.size) {
    case 2: e = acp_launch(tp_reduce_norm_kernel<2>, dim3(rows_per), dim3(threads), smem, s, p, T, hidden, gain, eps, epoch, done_ctr, px, diag); break;
    case 4: e = acp_launch(tp_reduce_norm_kernel<4>, dim3(rows_per), dim3(threads), smem, s, p, T, hidden, gain, eps, epoch, done_ctr, px, diag); break;
    case 8: e = acp_launch(tp_reduce_norm_kernel<8>, dim3(rows_per), dim3(threads), smem, s, p, T, hidden, gain, eps, epoch, done_ctr, px, diag); break;
    default: fprintf(stderr, "[acp_infer] tp exchange supports 2, 4 or 8 ranks (got %d)\n", p.size); return -1;
  }
  if (e != cudaSuccess) { fprintf(stderr, "[acp_infer] tp_reduce_norm launch: %s\n", cudaGetErrorString(e)); return -5; }
  return 0;
}

}  // namespace acp
