// hssb_leafx.cuh — the "X once" leaf variant of BASELINE.json's north star (2): ONE pass over X forms both
// D X and V' X.  Opt-in (HSSB_OPT_LEAF_FUSION = 1); the shipped default is the two-pass scheme of hssb_leaf2.cuh.
//
//   leafx_up_kernel    per leaf:  Z = V' X   and   Y = alpha D X + beta Y     (matmul.jl:34 and :46 in one sweep)
//                      [D ; V'] is streamed as ONE stacked operand: a ring stage holds KC columns of D, the same
//                      KC columns of V' and the matching KC rows of X; a warp's B fragments serve its 4 x TN tiles
//                      of D X and its TMV x TN tiles of V' X, so V' X costs no extra shared-memory reads of X.
//   (merges, translates: unchanged)
//   leafx_down_kernel  per leaf:  Y += alpha U F                              (matmul.jl:47)
//                      a stage holds U and the leaf's F tile; the old Y values are fetched into registers while
//                      the stage is in flight.  Memory bound: Y is read and written once more.
//
// Traffic per leaf against the two-pass scheme (m = 128, r = 32, k = 64, doubles): X is read once instead of
// twice (-64 KB) but Y is written twice and read once (+128 KB): +64 KB.  With the product FP64-bound the second
// read of X already hides under the DMMA work of the leaf-down kernel, so this variant is expected to lose;
// DESIGN.md §4 carries the measured numbers.
#pragma once

namespace hssb {

template <int M, int R, int NT_, int KC_>
struct LeafXUpCfg {
  static constexpr int NT = NT_, KC = KC_;
  static constexpr int NWARPS = 8, WR = M / 32, WC = NWARPS / WR;
  static constexpr int TM = 4, TN = NT / WC / 8, TMV = R / (8 * WR);
  static constexpr int KSTEPS = KC / 4, NCH = M / KC;
  static constexpr int LDA = M + 4, LDV = R + 4;
  static constexpr int XS_BYTES = KC * NT * 8, D_BYTES = KC * LDA * 8, V_BYTES = KC * LDV * 8;
  static constexpr int STAGE_BYTES = (XS_BYTES + D_BYTES + V_BYTES + 1023) / 1024 * 1024;
  static constexpr int BAR_BYTES = 256;
  static constexpr int FIT = (232448 - BAR_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = FIT > 12 ? 12 : FIT;
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + BAR_BYTES;
  static_assert(WR * WC == NWARPS && TN >= 1 && TMV >= 1 && R % (8 * WR) == 0, "warp tiling (needs rank % 32 == 0)");
  static_assert(KC % 16 == 0 && M % KC == 0 && NSTAGE >= 3 && 2 * NSTAGE <= BAR_BYTES / 8, "ring");
};

// up[i] / down[i]: the leaf-up and leaf-down tasks of the same leaf (both phases list the leaves left to right)
template <int M, int R, int NT_, int KC_>
__global__ void __launch_bounds__(288, 1)
leafx_up_kernel(const GTask* __restrict__ up, const GTask* __restrict__ down, int ntasks, int ntiles, CallParams p,
                const __grid_constant__ CUtensorMap xmap) {
  using C = LeafXUpCfg<M, R, NT_, KC_>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;  // [NSTAGE][X slab | D chunk | V' chunk]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + C::NSTAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  const int nrhs = p.nrhs;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], C::NWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (my <= 0) return;

  if (warp == C::NWARPS) {  // ====================== producer warp ======================
    int g = 0;
    for (int item = 0; item < my; ++item) {
      const int ti = (first + item) / ntiles, tile = (first + item) % ntiles;
      const int64_t d0 = down[ti].a0, v0 = up[ti].a0;
      const int x0 = (int)up[ti].b0;
      for (int c = 0; c < C::NCH; ++c, ++g) {
        const int st = g % C::NSTAGE;
        unsigned char* sp = stages + (size_t)st * C::STAGE_BYTES;
        mbar_wait(&a_empty[st], ((g / C::NSTAGE) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&a_full[st], (uint32_t)(C::XS_BYTES + C::D_BYTES + C::V_BYTES));
          bulk_g2s(sp + C::XS_BYTES, p.pool + d0 + (int64_t)c * C::KC * C::LDA, C::D_BYTES, &a_full[st]);
          bulk_g2s(sp + C::XS_BYTES + C::D_BYTES, p.pool + v0 + (int64_t)c * C::KC * C::LDV, C::V_BYTES, &a_full[st]);
        }
        __syncwarp();
        if (lane < C::KC / 16) tma_load_2d(sp + lane * (C::NT * 16 * 8), &xmap, x0 + c * C::KC + lane * 16, tile * C::NT, &a_full[st]);
      }
    }
    return;
  }

  // ====================== consumer warps ======================
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % C::WR, wc = warp / C::WR;
  const int pg = perm8(gq);
  double acc[C::TM][C::TN][2], accv[C::TMV][C::TN][2];
#pragma unroll
  for (int j = 0; j < C::TN; ++j) {
#pragma unroll
    for (int i = 0; i < C::TM; ++i) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int i = 0; i < C::TMV; ++i) accv[i][j][0] = accv[i][j][1] = 0.0;
  }
  int sw[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) sw[q] = ((q * 2 + (t >> 1)) ^ pg) * 2 + (t & 1);
  const int d_off = C::XS_BYTES / 8 + wr * 32 + gq + t * C::LDA;
  const int v_off = (C::XS_BYTES + C::D_BYTES) / 8 + wr * (C::TMV * 8) + gq + t * C::LDV;
  const int x_off = (wc * (C::TN * 8) + pg) * 16;
  int st = 0;
  uint32_t ph = 0;
  mbar_wait(&a_full[0], 0);
  for (int item = 0; item < my; ++item) {
    const bool more = item + 1 < my;
    const int ti = (first + item) / ntiles, tile = (first + item) - ti * ntiles;
    const int64_t y_row = down[ti].c, z_row = up[ti].c;
#pragma unroll 1
    for (int c = 0; c < C::NCH; ++c) {
      const int nst = (st + 1 == C::NSTAGE) ? 0 : st + 1;
      const uint32_t nph = (st + 1 == C::NSTAGE) ? ph ^ 1 : ph;
      const bool wait_next = c + 1 < C::NCH || more;
      const double* S = reinterpret_cast<const double*>(stages + (size_t)st * C::STAGE_BYTES);
      const double *A = S + d_off, *V = S + v_off, *B = S + x_off;
      bool ok = true;
#pragma unroll
      for (int kk = 0; kk < C::KSTEPS; ++kk) {
        double a[C::TM], av[C::TMV], b[C::TN];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) a[i] = A[kk * 4 * C::LDA + i * 8];
#pragma unroll
        for (int i = 0; i < C::TMV; ++i) av[i] = V[kk * 4 * C::LDV + i * 8];
#pragma unroll
        for (int j = 0; j < C::TN; ++j) b[j] = B[(kk >> 2) * (C::NT * 16) + j * 8 * 16 + sw[kk & 3]];
        if (kk == C::KSTEPS - 2 && wait_next) ok = mbar_try_wait(&a_full[nst], nph);
#pragma unroll
        for (int j = 0; j < C::TN; ++j) {
#pragma unroll
          for (int i = 0; i < C::TM; ++i) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
#pragma unroll
          for (int i = 0; i < C::TMV; ++i) mma_m8n8k4(accv[i][j][0], accv[i][j][1], av[i], b[j]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_empty[st]);
      if (!ok) mbar_wait(&a_full[nst], nph);
      st = nst; ph = nph;
    }
    // ---- epilogue: Y = alpha D X + beta Y, Z = V' X
    const int ncols = min(C::NT, nrhs - tile * C::NT);
    const int colb = wc * (C::TN * 8);
    const int pc[2] = {perm8(2 * t), perm8(2 * t + 1)};
    double* Y = p.Y + y_row + (int64_t)tile * C::NT * p.ldy + (int64_t)colb * p.ldy + wr * 32 + gq;
    double* Z = p.Z + z_row * (int64_t)nrhs + (int64_t)tile * C::NT * C::LDV + (int64_t)colb * C::LDV + wr * (C::TMV * 8) + gq;
#pragma unroll
    for (int j = 0; j < C::TN; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = colb + j * 8 + pc[e] < ncols;
        double* ycol = Y + (int64_t)(j * 8 + pc[e]) * p.ldy;
        double* zcol = Z + (int64_t)(j * 8 + pc[e]) * C::LDV;
#pragma unroll
        for (int i = 0; i < C::TM; ++i) {
          if (live) {
            double v = acc[i][j][e] * p.alpha;
            if (p.beta != 0.0) v += p.beta * ycol[i * 8];  // beta == 0 never reads Y (matmul.jl:13)
            ycol[i * 8] = v;
          }
          acc[i][j][e] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < C::TMV; ++i) {
          if (live) zcol[i * 8] = accv[i][j][e];
          accv[i][j][e] = 0.0;
        }
      }
    }
  }
}

template <int M, int R, int NT_>
struct LeafXDownCfg {
  static constexpr int NT = NT_;
  static constexpr int NWARPS = 8, WR = M / 32, WC = NWARPS / WR;
  static constexpr int TM = 4, TN = NT / WC / 8;
  static constexpr int KSTEPS = R / 4;
  static constexpr int LDA = M + 4, LDF = R + 4;
  static constexpr int U_BYTES = R * LDA * 8, F_BYTES = NT * LDF * 8;
  static constexpr int STAGE_BYTES = (U_BYTES + F_BYTES + 127) / 128 * 128;
  static constexpr int BAR_BYTES = 256;
  static constexpr int FIT = (232448 - BAR_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = FIT > 6 ? 6 : FIT;
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + BAR_BYTES;
  static_assert(WR * WC == NWARPS && TN >= 1 && NSTAGE >= 2, "configuration");
};

template <int M, int R, int NT_>
__global__ void __launch_bounds__(288, 1)
leafx_down_kernel(const GTask* __restrict__ down, int ntasks, int ntiles, CallParams p) {
  using C = LeafXDownCfg<M, R, NT_>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;  // [NSTAGE][U | F tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + C::NSTAGE;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  const int nrhs = p.nrhs;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], C::NWARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (my <= 0) return;

  if (warp == C::NWARPS) {  // ====================== producer warp ======================
    if (lane != 0) return;
    for (int item = 0; item < my; ++item) {
      const int ti = (first + item) / ntiles, tile = (first + item) % ntiles;
      const GTask& tk = down[ti];
      const int st = item % C::NSTAGE;
      unsigned char* sp = stages + (size_t)st * C::STAGE_BYTES;
      const uint32_t fbytes = (uint32_t)(min(C::NT, nrhs - tile * C::NT) * C::LDF * 8);
      mbar_wait(&a_empty[st], ((item / C::NSTAGE) & 1) ^ 1);
      mbar_expect_tx(&a_full[st], (uint32_t)C::U_BYTES + fbytes);
      bulk_g2s(sp, p.pool + tk.a1, C::U_BYTES, &a_full[st]);
      bulk_g2s(sp + C::U_BYTES, p.F + tk.b1 * (int64_t)nrhs + (int64_t)tile * C::NT * C::LDF, fbytes, &a_full[st]);
    }
    return;
  }

  // ====================== consumer warps ======================
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % C::WR, wc = warp / C::WR;
  const int pg = perm8(gq);
  const int colb = wc * (C::TN * 8);
  const int pc[2] = {perm8(2 * t), perm8(2 * t + 1)};
  for (int item = 0; item < my; ++item) {
    const int ti = (first + item) / ntiles, tile = (first + item) - ti * ntiles;
    const int st = item % C::NSTAGE;
    const int ncols = min(C::NT, nrhs - tile * C::NT);
    double* Y = p.Y + down[ti].c + (int64_t)tile * C::NT * p.ldy + (int64_t)colb * p.ldy + wr * 32 + gq;
    // the old Y values travel while the stage does
    double yold[C::TM][C::TN][2];
#pragma unroll
    for (int j = 0; j < C::TN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = colb + j * 8 + pc[e] < ncols;
#pragma unroll
        for (int i = 0; i < C::TM; ++i) yold[i][j][e] = live ? __ldcs(Y + (int64_t)(j * 8 + pc[e]) * p.ldy + i * 8) : 0.0;
      }
    double acc[C::TM][C::TN][2];
#pragma unroll
    for (int i = 0; i < C::TM; ++i)
#pragma unroll
      for (int j = 0; j < C::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    mbar_wait(&a_full[st], (item / C::NSTAGE) & 1);
    const double* S = reinterpret_cast<const double*>(stages + (size_t)st * C::STAGE_BYTES);
    const double* A = S + wr * 32 + gq + t * C::LDA;
    const double* B = S + C::U_BYTES / 8 + (colb + pg) * C::LDF + t;
#pragma unroll
    for (int kk = 0; kk < C::KSTEPS; ++kk) {
      double a[C::TM], b[C::TN];
#pragma unroll
      for (int i = 0; i < C::TM; ++i) a[i] = A[kk * 4 * C::LDA + i * 8];
#pragma unroll
      for (int j = 0; j < C::TN; ++j) b[j] = B[j * 8 * C::LDF + kk * 4];
#pragma unroll
      for (int i = 0; i < C::TM; ++i)
#pragma unroll
        for (int j = 0; j < C::TN; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&a_empty[st]);
#pragma unroll
    for (int j = 0; j < C::TN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (colb + j * 8 + pc[e] < ncols) {
#pragma unroll
          for (int i = 0; i < C::TM; ++i) Y[(int64_t)(j * 8 + pc[e]) * p.ldy + i * 8] = fma(p.alpha, acc[i][j][e], yold[i][j][e]);
        }
      }
  }
}

}  // namespace hssb
