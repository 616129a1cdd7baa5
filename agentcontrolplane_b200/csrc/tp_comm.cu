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
  if (lane < P.size) {
    __threadfence_system();
    st_release_sys(P.flags[lane] + P.rank, epoch);
  }
}
ACP_DEVINL void tp_wait_all(const TpPeers& P, int lane, int epoch) {  // lanes 0..size-1 of one warp
  if (lane < P.size) {
    unsigned spins = 0;
    while (ld_acquire_sys(P.flags[P.rank] + lane) < epoch) {
      if (++spins > (1u << 27)) {
        printf("[acp_infer] tp wait timeout rank=%d waiting for %d epoch=%d\n", P.rank, lane, epoch);
        __trap();
      }
    }
  }
}

// wait-only kernel: every rank's rows have landed in this rank's x / xn
__global__ void __launch_bounds__(32)
tp_wait_kernel(TpPeers P, int epoch) {
  pdl_launch_dependents();
  pdl_wait();
  tp_wait_all(P, threadIdx.x, epoch);
}

ACP_DEVINL float4 ld_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// One kernel: [signal "my partial sums are complete"] -> [wait for every rank] -> pull + reduce +
// residual + RMSNorm + push for the rows this rank owns -> [last CTA signals "my rows are pushed"].
// Epochs: `epoch` = ready, `epoch + 1` = done; flags only grow.
__global__ void __launch_bounds__(1024)
tp_reduce_norm_kernel(TpPeers P, int T, int hidden, const __nv_bfloat16* __restrict__ gain, float eps,
                      int epoch, int* done_ctr, int push_x) {
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
      float4 q[TP_MAX];
#pragma unroll
      for (int p = 0; p < TP_MAX; ++p) q[p] = ld_f4(P.ar[p < P.size ? p : P.size - 1] + off + i);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int p = 0; p < TP_MAX; ++p)
        if (p < P.size) { acc.x += q[p].x; acc.y += q[p].y; acc.z += q[p].z; acc.w += q[p].w; }
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
        for (int p = 0; p < TP_MAX; ++p)
          if (p < P.size) *reinterpret_cast<uint2*>(P.x[p] + off + i) = packed;
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
      for (int p = 0; p < TP_MAX; ++p)
        if (p < P.size) *reinterpret_cast<uint2*>(P.xn[p] + off + i) = packed;
    }
  }
  // the last CTA of this rank to finish its pushes tells every rank "my rows have landed"
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    const int old = atomicAdd(done_ctr, 1);
    is_last = (old == (int)gridDim.x - 1);
    if (is_last) *done_ctr = 0;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) tp_signal(P, threadIdx.x, epoch + 1);
}

}  // namespace

int launch_tp_wait(const TpPeers& p, int epoch, cudaStream_t s) {
  cudaError_t e = acp_launch(tp_wait_kernel, dim3(1), dim3(32), 0, s, p, epoch);
  if (e != cudaSuccess) { fprintf(stderr, "[acp_infer] tp_wait launch: %s\n", cudaGetErrorString(e)); return -5; }
  return 0;
}

int launch_tp_reduce_norm(const TpPeers& p, int T, int hidden, const __nv_bfloat16* gain, float eps,
                          int epoch, int* done_ctr, bool push_x, cudaStream_t s) {
  if (T <= 0) return 0;
  const int rows_per = (T + p.size - 1) / p.size;
  int threads = ((hidden / 4 + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  cudaError_t e = acp_launch(tp_reduce_norm_kernel, dim3(rows_per), dim3(threads), hidden * sizeof(float), s, p,
                             T, hidden, gain, eps, epoch, done_ctr, push_x ? 1 : 0);
  if (e != cudaSuccess) { fprintf(stderr, "[acp_infer] tp_reduce_norm launch: %s\n", cudaGetErrorString(e)); return -5; }
  return 0;
}

}  // namespace acp
