// hssb_kernels_generic.cuh — any-shape tile kernel for one level of the HSS
// product.  Handles ragged leaves (62/63 rows), variable and zero ranks,
// unbalanced trees, nrhs not a multiple of anything.  One CTA = one 64x64 tile
// of one task's output; K is walked in 16-wide slabs staged through shared
// memory with coalesced (unit-stride) global reads for both A layouts.
#pragma once

#include "hssb_internal.h"

namespace hssb {

constexpr int G_TM = 64, G_TN = 64, G_TK = 16, G_THREADS = 256;

__device__ __forceinline__ const double* operand_b(const CallParams& p, int src, int64_t row, int32_t ldw,
                                                   int64_t& ld) {
  switch (src) {
    case SRC_X: ld = p.ldx; return p.X + row;
    case SRC_Z: ld = ldw; return p.Z + row * (int64_t)p.nrhs;
    case SRC_F: ld = ldw; return p.F + row * (int64_t)p.nrhs;
    default: ld = p.ldy; return p.Y + row;
  }
}

__global__ void __launch_bounds__(G_THREADS)
generic_level_kernel(const GTask* __restrict__ tasks, CallParams p) {
  const GTask t = tasks[blockIdx.x];
  const int m0 = blockIdx.y * G_TM;
  if (m0 >= t.M) return;
  const int n0 = blockIdx.z * G_TN;
  const int N = p.nrhs;

  __shared__ double As[G_TK][G_TM + 1];
  __shared__ double Bs[G_TK][G_TN + 1];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const int K = s ? t.K1 : t.K0;
    if (K <= 0) continue;
    const double* A = p.pool + (s ? t.a1 : t.a0);
    const int64_t lda = s ? t.lda1 : t.lda0;
    const bool ta = s ? t.ta1 : t.ta0;
    int64_t ldb;
    const double* B = operand_b(p, s ? t.sb1 : t.sb0, s ? t.b1 : t.b0, s ? t.ldb1 : t.ldb0, ldb);

    for (int k0 = 0; k0 < K; k0 += G_TK) {
      // ---- stage A slab: As[kk][mm] = op(A)(m0+mm, k0+kk)
      if (!ta) {
        const int mm = tid & 63;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int kk = (tid >> 6) + 4 * r;
          double v = 0.0;
          if (m0 + mm < t.M && k0 + kk < K) v = A[(int64_t)(k0 + kk) * lda + (m0 + mm)];
          As[kk][mm] = v;
        }
      } else {
        const int kk = tid & 15;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int mm = (tid >> 4) + 16 * r;
          double v = 0.0;
          if (m0 + mm < t.M && k0 + kk < K) v = A[(int64_t)(m0 + mm) * lda + (k0 + kk)];
          As[kk][mm] = v;
        }
      }
      // ---- stage B slab: Bs[kk][nn] = B(k0+kk, n0+nn)
      {
        const int kk = tid & 15;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int nn = (tid >> 4) + 16 * r;
          double v = 0.0;
          if (n0 + nn < N && k0 + kk < K) v = B[(int64_t)(n0 + nn) * ldb + (k0 + kk)];
          Bs[kk][nn] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < G_TK; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue
  int64_t ldc;
  double* C;
  switch (t.sc) {
    case SRC_Z: ldc = t.ldc; C = p.Z + t.c * (int64_t)N; break;
    case SRC_F: ldc = t.ldc; C = p.F + t.c * (int64_t)N; break;
    default: ldc = p.ldy; C = p.Y + t.c; break;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + ty + 16 * j;
    if (col >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + tx + 16 * i;
      if (row >= t.M) continue;
      double* dst = C + (int64_t)col * ldc + row;
      double v = acc[i][j];
      if (t.epilogue) {
        v *= p.alpha;
        if (p.beta != 0.0) v += p.beta * (*dst);  // beta == 0 never reads Y (matmul.jl:13)
      }
      *dst = v;
    }
  }
}

}  // namespace hssb
