// gemm_out.cuh — how a consumer kernel reads a GEMM result: bf16 [T][ld], or `splits` fp32 split-K
// planes summed in index order and rounded to bf16 once (the deterministic split-K reduction).
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace acp {

// Reading 4 consecutive GEMM outputs (columns m..m+3 of row t) as bf16-rounded fp32.
struct GemmOutDev {
  const void* ptr;
  int splits, n_cap, ld;
};
static inline GemmOutDev to_dev(const GemmOut& g) { return GemmOutDev{g.ptr, g.splits, g.n_cap, g.ld}; }

ACP_DEVINL float4 ld_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
ACP_DEVINL void gemm_out_load4(const GemmOutDev& g, int t, int m, float (&v)[4]) {
  if (g.splits == 0) {
    const uint2 raw = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)g.ptr + (size_t)t * g.ld + m);
    v[0] = bf16_lo(raw.x); v[1] = bf16_hi(raw.x); v[2] = bf16_lo(raw.y); v[3] = bf16_hi(raw.y);
  } else {
    const float* p = (const float*)g.ptr + (size_t)t * g.ld + m;
    const size_t plane = (size_t)g.n_cap * g.ld;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // 8 planes per batch: the loads are UNCONDITIONAL (index clamped) so that all of them are in
    // flight together; only the adds are predicated.  Sum order is plane 0,1,2,... => deterministic.
    for (int s0 = 0; s0 < g.splits; s0 += 8) {
      float4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sj = (s0 + j < g.splits) ? s0 + j : g.splits - 1;
        q[j] = ld_nc_f4(p + (size_t)sj * plane);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (s0 + j < g.splits) { acc.x += q[j].x; acc.y += q[j].y; acc.z += q[j].z; acc.w += q[j].w; }
      }
    }
    v[0] = bf16_round(acc.x); v[1] = bf16_round(acc.y); v[2] = bf16_round(acc.z); v[3] = bf16_round(acc.w);
  }
}


}  // namespace acp
