// tp_comm.h — tensor-parallel all-reduce FUSED with the residual add and RMSNorm, over NVLink
// peer memory (single process, cudaDeviceEnablePeerAccess: every shard can load/store every
// other shard's buffers directly).
//
//   x   = bf16( x + bf16( sum_ranks partial ) )          (row-parallel O / down projection)
//   xn  = g * bf16( x * rsqrt(mean(x^2) + eps) )
//
// Rank r owns a contiguous block of token rows.  For its rows it LOADS the fp32 partial sums of
// all ranks straight from their HBM over NVLink (reduce-scatter by pull, fixed rank order =>
// deterministic and bit-identical on every rank), applies residual + RMSNorm in registers, and
// STORES the normed rows into every rank's buffers (all-gather by push); the new bf16 residual
// stays with its owner except after the last layer (push_x), when the sampler needs any row.
// The reduced fp32 activation never exists in HBM.  The flag handshakes (release/acquire at system
// scope) live INSIDE the kernel: CTA 0 signals "partials ready", every CTA waits for all ranks, the
// last CTA to finish signals "rows pushed" and CTA 0 waits for every rank's "rows pushed" before the grid completes.  Replaces ncclAllReduce + add_rmsnorm of the NCCL baseline path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace acp {

constexpr int TP_MAX = 8;
struct TpPeers {
  int rank = 0, size = 1;
  const float* ar[TP_MAX];        // [T][H] fp32 partial sums of each rank
  __nv_bfloat16* x[TP_MAX];       // residual stream replica of each rank
  __nv_bfloat16* xn[TP_MAX];      // normed activations replica of each rank
  int* flags[TP_MAX];             // flags[p][r]: rank r's arrival counter as seen by rank p
};

// epoch = "partials ready", epoch + 1 = "rows pushed"; done_ctr: one zeroed int per rank
int launch_tp_reduce_norm(const TpPeers& p, int T, int hidden, const __nv_bfloat16* gain, float eps,
                          int epoch, int* done_ctr, bool push_x, cudaStream_t s, int diag = 0);  // diag: bench only

}  // namespace acp
