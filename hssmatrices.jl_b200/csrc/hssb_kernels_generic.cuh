// hssb_kernels_generic.cuh — any-shape tile kernel for one level of the HSS
// product (and of the ULV solve).  Handles ragged leaves (62/63 rows), variable and
// zero ranks, unbalanced trees, nrhs not a multiple of anything.  One CTA = one 64x64
// tile of one task's output; K is walked in 16-wide slabs staged through shared
// memory with coalesced (unit-stride) global reads for both A layouts (masked, zero
// filled); the next slab is prefetched into registers while the current one is
// consumed with FP64 tensor-core DMMA m8n8k4 from conflict-free shared-memory images.
#pragma once

#include "hssb_internal.h"
#include "hssb_mma.cuh"

namespace hssb {
// programmatic dependent launch (HSSB_OPT_PDL, see hssb_fast.cuh); no-ops for kernels launched without the attribute
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
}  // namespace hssb

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

// Shared-memory images of one K slab, laid out so that both the staging stores and the DMMA fragment
// loads are bank-conflict free (16 doubles per bank row):
//   A applied as stored ("N", column-major M x K): k-major  As[kk * G_SA + m],  G_SA = 68 = 4 mod 16
//   A applied transposed (stored K x M):           m-major  As[m * G_SB + kk],  G_SB = 20 = 4 mod 16
//   B (always read K-fastest from global):         n-major  Bs[n * G_SB + kk]
// A fragment lane (g, t) reads A[m0 + g][k0 + t]: offsets 4t + g resp. 4g + t mod 16, all distinct
// within a half warp.
constexpr int G_SA = 68, G_SB = 20, G_SMEM = G_TM * G_SB;  // 1280 >= 16 * 68
static_assert(G_TK * G_SA <= G_SMEM, "k-major image must fit");

// One 64 x 64 output tile (rows m0.., right-hand sides n0..) of one task.  WS_CG: operands from the Z / F
// workspaces are read with ld.global.cg (L2 only) -- needed when the producer of the block ran in the SAME launch
// on another SM (hssb_flow.cuh), where a stale L1 line would be a wrong answer.
template <bool WS_CG, bool PDL = false>
__device__ __forceinline__ void generic_tile(const GTask& t, const int m0, const int n0, const CallParams& p, double* As, double* Bs) {
  const int N = p.nrhs;
  const int tid = threadIdx.x;
  // FP64 tensor path: 8 warps tile the 64 x 64 output as 2 (M) x 4 (N); a warp owns 32 x 16 =
  // 4 x 2 m8n8 accumulators (DMMA m8n8k4), lane = 4 g + q.
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  // Warps 0-3 (one per scheduler) own rows 0-31, warps 4-7 rows 32-63: when a task's last row tile
  // holds 32 rows or fewer (M = 32, 96, ... are the common ranks) the upper warps skip the tensor
  // work and every scheduler issues half as many DMMAs instead of two schedulers idling.
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
  const bool active = m0 + wm < t.M && n0 + wn < N;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // The K loop runs over the slabs of both products back to back; the global loads of slab i+1
  // are issued into registers before slab i is consumed from shared memory, so a level of small
  // tasks (a few slabs each) pays the global-memory latency once instead of once per slab.
  const int K0 = t.K0 > 0 ? t.K0 : 0, K1 = t.K1 > 0 ? t.K1 : 0;
  const int nslab0 = (K0 + G_TK - 1) / G_TK, nslab = nslab0 + (K1 + G_TK - 1) / G_TK;
  int64_t ldb0 = 0, ldb1 = 0;
  const double* B0 = operand_b(p, t.sb0, t.b0, t.ldb0, ldb0);
  const double* B1 = operand_b(p, t.sb1, t.b1, t.ldb1, ldb1);
  double ra[4], rb[4];
  auto fetch_a = [&](int slab) {
    const bool s1 = slab >= nslab0;
    const int K = s1 ? K1 : K0, k0 = (s1 ? slab - nslab0 : slab) * G_TK;
    const double* A = p.pool + (s1 ? t.a1 : t.a0);
    const int64_t lda = s1 ? t.lda1 : t.lda0;
    const bool ta = s1 ? t.ta1 : t.ta0;
    if (!ta) {
      const int mm = tid & 63;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int kk = (tid >> 6) + 4 * r;
        ra[r] = (m0 + mm < t.M && k0 + kk < K) ? A[(int64_t)(k0 + kk) * lda + (m0 + mm)] : 0.0;
      }
    } else {
      const int kk = tid & 15;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int mm = (tid >> 4) + 16 * r;
        ra[r] = (m0 + mm < t.M && k0 + kk < K) ? A[(int64_t)(m0 + mm) * lda + (k0 + kk)] : 0.0;
      }
    }
  };
  auto fetch_b = [&](int slab) {
    const bool s1 = slab >= nslab0;
    const int K = s1 ? K1 : K0, k0 = (s1 ? slab - nslab0 : slab) * G_TK;
    const double* B = s1 ? B1 : B0;
    const int64_t ldb = s1 ? ldb1 : ldb0;
    const bool ws = WS_CG && (s1 ? t.sb1 : t.sb0) != SRC_X;
    const int kk = tid & 15;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int nn = (tid >> 4) + 16 * r;
      const bool in = n0 + nn < N && k0 + kk < K;
      const double* src = B + (int64_t)(n0 + nn) * ldb + (k0 + kk);
      rb[r] = in ? (ws ? __ldcg(src) : *src) : 0.0;
    }
  };
  auto fetch = [&](int slab) { fetch_a(slab); fetch_b(slab); };
  auto stage = [&](int slab) {  // registers -> shared memory, same element mapping as fetch()
    const bool ta = slab >= nslab0 ? t.ta1 : t.ta0;
    if (!ta) {
#pragma unroll
      for (int r = 0; r < 4; ++r) As[((tid >> 6) + 4 * r) * G_SA + (tid & 63)] = ra[r];
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) As[((tid >> 4) + 16 * r) * G_SB + (tid & 15)] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) Bs[((tid >> 4) + 16 * r) * G_SB + (tid & 15)] = rb[r];
  };

  // HSSB_OPT_PDL (hssb_fast.cuh): launched with the attribute, the CTA is placed while the previous level still runs; the first
  // slab of the generator block -- which nobody produces -- travels before the wait, everything else after it
  if (nslab > 0) fetch_a(0);
  if (PDL) pdl_wait();
  if (nslab > 0) fetch_b(0);
  for (int slab = 0; slab < nslab; ++slab) {
    stage(slab);
    __syncthreads();
    if (slab + 1 < nslab) fetch(slab + 1);
    const bool ta = slab >= nslab0 ? t.ta1 : t.ta0;
    const double* ap = ta ? As + (wm + g) * G_SB + q : As + q * G_SA + wm + g;
    const int a_tile = ta ? 8 * G_SB : 8, a_step = ta ? 4 : 4 * G_SA;  // next m8 tile / next k4 step
    const double* bp = Bs + (wn + g) * G_SB + q;
    // k-steps past the end of the operand are zero fill: skipping them leaves the result unchanged (x + 0 * 0) and
    // shortens the chain of dependent DMMAs (~500 cycles each for a warp with nothing else in flight)
    const int ks_end = ((slab >= nslab0 ? K1 - (slab - nslab0) * G_TK : K0 - slab * G_TK) + 3) >> 2;
    if (active) {
#pragma unroll
    for (int ks = 0; ks < G_TK / 4; ++ks) {
      if (ks >= ks_end) break;
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = ap[ks * a_step + i * a_tile];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = bp[j * 8 * G_SB + ks * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    }
    __syncthreads();
  }

  // ---- epilogue: lane (g, q) of accumulator (i, j) holds C[wm + 8 i + g][wn + 8 j + 2 q + {0, 1}]
  int64_t ldc;
  double* C;
  switch (t.sc) {
    case SRC_Z: ldc = t.ldc; C = p.Z + t.c * (int64_t)N; break;
    case SRC_F: ldc = t.ldc; C = p.F + t.c * (int64_t)N; break;
    default: ldc = p.ldy; C = p.Y + t.c; break;
  }
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = n0 + wn + 8 * j + 2 * q + e;
      if (col >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm + 8 * i + g;
        if (row >= t.M) continue;
        double* dst = C + (int64_t)col * ldc + row;
        double v = acc[i][j][e];
        if (t.epilogue) {
          v *= p.alpha;
          if (p.beta != 0.0) v += p.beta * (*dst);  // beta == 0 never reads Y (matmul.jl:13)
        }
        *dst = v;
      }
    }
}

__global__ void __launch_bounds__(G_THREADS, 3)
generic_level_kernel(const GTask* __restrict__ tasks, CallParams p) {
  __shared__ double As[G_SMEM];
  __shared__ double Bs[G_SMEM];
  pdl_launch_dependents();
  const GTask t = tasks[blockIdx.x];
  const int m0 = blockIdx.y * G_TM;
  if (m0 >= t.M) return;
  generic_tile<false, true>(t, m0, blockIdx.z * G_TN, p, As, Bs);
}

}  // namespace hssb
