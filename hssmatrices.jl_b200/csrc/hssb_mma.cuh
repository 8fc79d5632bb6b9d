// hssb_mma.cuh — FP64 tensor-core primitive.  sm_100a has no tcgen05 kind for
// FP64; the FP64 tensor path is the warp-level mma.sync m8n8k4, which ptxas
// lowers to DMMA.8x8x4 (checked with cuobjdump -sass).
//
// Fragment layout (PTX ISA, mma.m8n8k4 .f64), lane = 4*g + t, g = 0..7, t = 0..3:
//   A (8x4, row):  lane holds A[g][t]
//   B (4x8, col):  lane holds B[t][g]
//   C/D (8x8):     lane holds C[g][2t], C[g][2t+1]
#pragma once

namespace hssb {

__device__ __forceinline__ void mma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace hssb
