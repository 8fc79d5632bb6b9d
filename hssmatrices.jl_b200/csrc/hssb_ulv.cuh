// hssb_ulv.cuh — implicit ULV factorisation of an HSS matrix (hssA \ B,
// src/ulvfactor.jl:10-107 of the reference).
//
// The reference factorises AND solves in one recursive pass on every call.  Here
// the two are split:
//   * hssb_ulv_factor  (this file): one CTA per tree node, level by level from the
//     leaves to the root, computes the node's orthogonal transformations with
//     Householder reflectors (the geqlf/ormql/gelqf/ormlq/trsm sequence of
//     _ulvreduce!, ulvfactor.jl:22-57) and FOLDS them into a handful of explicit
//     matrices per node, stored in a second level-ordered pool;
//   * hssb_solve: with those matrices every step of the solve is again
//     C = A0*B0 + A1*B1 on small blocks, i.e. the level-scheduled task table and
//     the kernels of the product path (hssb_api.cu: build_plan_ulv).
//
// Per node, with the incoming reduced block D (m x n), U (m x kr), V (n x kw) and the
// right-hand side rows b (m x nrhs), k = kr < m, mk = m - k:
//   Q  U   = [R; 0]                       Householder QR of U (the reference's QL: same
//                                         subspace split, rows ordered the other way round)
//   Dq     = Q D,  Qbot = Q[0:k,:], Qtop = Q[k:,:]
//   Dq[k:,:] P' = [L1 0]                  LQ of the mk x n block (QR of its transpose)
//   L2     = Dq[0:k,:] P'                 (ulvfactor.jl:47)
//   zloc   = L1^-1 Qtop b        =: T1 b  (ulvfactor.jl:48)
//   b_out  = Qbot b - L2[:, :mk] zloc = (Qbot - L2a T1) b =: T2 b      (ulvfactor.jl:49)
//   u      = (P V)[0:mk,:]' zloc = ((PV)top' T1) b        =: T3 b      (ulvfactor.jl:50-51)
//   handed up: D~ = L2[:, mk:], U~ = R, V~ = (P V)[mk:, :]            (ulvfactor.jl:53-54)
//   top-down: z[cols] = P' [zloc; z_trailing]                          (ulvfactor.jl:98-107)
// A branch first merges its children (ulvfactor.jl:74-79); its incoming right-hand side is
//   b = [b1 - U~1 B12 u2; b2 - U~2 B21 u1],  u += W1' u1 + W2' u2      (ulvfactor.jl:74, :89)
// which is linear in the children's c_s = [b_s; u_s], so T1..T3 fold into one matrix per child.
//
// Every routine is __host__ __device__ and written against a "team" (thread id, thread
// count, barrier): on the device a team is one CTA, on the host it is a single thread.  The
// host instantiation exists for the CPU tests of plan-only handles (hssb_debug_ulv_factor_host),
// which run the very same code path through the numpy plan interpreter.
#pragma once

#include <math.h>

#include "hssb_internal.h"

namespace hssb {

#define HSSB_HD __host__ __device__ __forceinline__

struct Team {
  int tid, nt;
  HSSB_HD void sync() const {
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
  }
};

// Strided matrix view: element (i, j) = p[i * rs + j * cs].
struct Mat {
  double* p;
  int64_t rs, cs;
  HSSB_HD double& operator()(int64_t i, int64_t j) const { return p[i * rs + j * cs]; }
  HSSB_HD Mat sub(int64_t i0, int64_t j0) const { return Mat{p + i0 * rs + j0 * cs, rs, cs}; }
  HSSB_HD Mat t() const { return Mat{p, cs, rs}; }
};
HSSB_HD Mat colmajor(double* p, int64_t ld) { return Mat{p, 1, ld}; }
HSSB_HD Mat colmajor(const double* p, int64_t ld) { return Mat{const_cast<double*>(p), 1, ld}; }
HSSB_HD Mat rowmajor(double* p, int64_t ld) { return Mat{p, ld, 1}; }

// C (M x N) = alpha * A (M x K) * B (K x N) + beta * C   (beta == 0 never reads C)
// One thread per 4 x 2 block of C: six loads feed eight FMAs per k (one thread per element needed two loads per FMA, and
// the factorisation is bound by exactly that: ncu on the leaf level of config 3 shows 183 MB of L1 load traffic per node,
// 0.57 IPC per SM).  Eight independent accumulation chains per thread.
HSSB_HD void tm_gemm(const Team& tm, Mat C, Mat A, Mat B, int M, int N, int K, double alpha, double beta) {
  // C never overlaps A or B (callers pass distinct blocks): without __restrict__ every store to C
  // would order the following loads behind it
  const int nbi = (M + 3) / 4, nbj = (N + 1) / 2;
  const int64_t total = (int64_t)nbi * nbj;
  for (int64_t e = tm.tid; e < total; e += tm.nt) {
    const int i0 = (int)(e % nbi) * 4, j0 = (int)(e / nbi) * 2;
    const double* __restrict__ a[4];
    const double* __restrict__ b[2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) a[ii] = A.p + (i0 + ii < M ? i0 + ii : M - 1) * A.rs;   // rows past the edge: recomputed, never stored
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) b[jj] = B.p + (j0 + jj < N ? j0 + jj : N - 1) * B.cs;
    double s[4][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) s[ii][0] = s[ii][1] = 0.0;
    for (int kk = 0; kk < K; ++kk) {
      double av[4], bv[2];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) av[ii] = a[ii][kk * A.cs];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) bv[jj] = b[jj][kk * B.rs];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) s[ii][jj] = fma(av[ii], bv[jj], s[ii][jj]);
    }
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
        if (i0 + ii < M && j0 + jj < N) {
          double& c = C(i0 + ii, j0 + jj);
          c = beta == 0.0 ? alpha * s[ii][jj] : alpha * s[ii][jj] + beta * c;
        }
  }
  tm.sync();
}

HSSB_HD void tm_copy(const Team& tm, Mat dst, Mat src, int M, int N, double scale = 1.0) {
  const int64_t total = (int64_t)M * N;
  double* __restrict__ d = dst.p;  // dst and src never overlap
  const double* __restrict__ sp = src.p;
  for (int64_t e = tm.tid; e < total; e += tm.nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    d[i * dst.rs + j * dst.cs] = scale * sp[i * src.rs + j * src.cs];
  }
  tm.sync();
}

HSSB_HD void tm_add(const Team& tm, Mat dst, Mat src, int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = tm.tid; e < total; e += tm.nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    dst(i, j) += src(i, j);  // tiny blocks (kw x kw_s), no need to untangle the aliasing
  }
  tm.sync();
}

// dst (M x N) = diag * [i == j]
HSSB_HD void tm_eye(const Team& tm, Mat dst, int M, int N, double diag) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = tm.tid; e < total; e += tm.nt) {
    const int i = (int)(e % M), j = (int)(e / M);
    dst(i, j) = i == j ? diag : 0.0;
  }
  tm.sync();
}

// Householder QR of A (p x q) in place with the orthogonal factor applied to Qt (p x nq) as well:
// on return A = [R; 0] (upper triangular, exact zeros below the diagonal) and Qt <- H_s ... H_1 Qt,
// so Qt = Q' when it enters as the identity.  One thread per column of [A | Qt]: the reflector is
// rebuilt by every thread from column j (a broadcast read), no reductions across threads.
HSSB_HD void tm_qr(const Team& tm, Mat A, int p, int q, Mat Qt, int nq) {
  const int steps = q < p - 1 ? q : p - 1;
  for (int j = 0; j < steps; ++j) {
    double tail = 0.0;  // sum of squares below the diagonal
    for (int i = j + 1; i < p; ++i) tail = fma(A(i, j), A(i, j), tail);
    const double x0 = A(j, j);
    const bool act = tail > 0.0;      // nothing to annihilate: H = I
    const double nrm = sqrt(fma(x0, x0, tail));
    const double alpha = x0 >= 0.0 ? -nrm : nrm;
    const double v0 = x0 - alpha;
    const double beta = act ? 2.0 / (tail + v0 * v0) : 0.0;
    if (act) {
      const int na = q - j - 1, ncol = na + nq;
      // column j (the reflector) is only read during this step and never overlaps the column a
      // thread updates: __restrict__ lets the loads run ahead of the stores
      const double* __restrict__ v = A.p + j * A.cs;
      const int64_t vs = A.rs;
      for (int cc = tm.tid; cc < ncol; cc += tm.nt) {
        const Mat Mx = cc < na ? A.sub(0, j + 1 + cc) : Qt.sub(0, cc - na);
        double* __restrict__ x = Mx.p;
        const int64_t xs = Mx.rs;
        double w0 = v0 * x[j * xs], w1 = 0.0;
        int i = j + 1;
        for (; i + 1 < p; i += 2) {
          w0 = fma(v[i * vs], x[i * xs], w0);
          w1 = fma(v[(i + 1) * vs], x[(i + 1) * xs], w1);
        }
        if (i < p) w0 = fma(v[i * vs], x[i * xs], w0);
        const double w = (w0 + w1) * beta;
        x[j * xs] -= w * v0;
        for (i = j + 1; i < p; ++i) x[i * xs] = fma(-w, v[i * vs], x[i * xs]);
      }
    }
    tm.sync();  // column j is still intact up to here: everybody has read it
    if (act && tm.tid == 0) {
      A(j, j) = alpha;
      for (int i = j + 1; i < p; ++i) A(i, j) = 0.0;
    }
    // the next step reads columns > j only; the final barrier below publishes column j
  }
  tm.sync();
}

// X (n x ncols) <- L^-1 X, L lower triangular n x n (forward substitution, one thread per column)
HSSB_HD void tm_trsm_lower(const Team& tm, Mat L, int n, Mat X, int ncols) {
  for (int c = tm.tid; c < ncols; c += tm.nt) {
    double* __restrict__ x = X.p + c * X.cs;  // L and X never overlap
    const double* __restrict__ l = L.p;
    for (int i = 0; i < n; ++i) {
      double s0 = x[i * X.rs], s1 = 0.0;
      int j = 0;
      for (; j + 1 < i; j += 2) {
        s0 = fma(-l[i * L.rs + j * L.cs], x[j * X.rs], s0);
        s1 = fma(-l[i * L.rs + (j + 1) * L.cs], x[(j + 1) * X.rs], s1);
      }
      if (j < i) s0 = fma(-l[i * L.rs + j * L.cs], x[j * X.rs], s0);
      x[i * X.rs] = (s0 + s1) / l[i * L.rs + i * L.cs];
    }
  }
  tm.sync();
}

// X (n x ncols) <- R^-1 X, R upper triangular n x n (back substitution)
HSSB_HD void tm_trsm_upper(const Team& tm, Mat R, int n, Mat X, int ncols) {
  for (int c = tm.tid; c < ncols; c += tm.nt) {
    double* __restrict__ x = X.p + c * X.cs;  // R and X never overlap
    const double* __restrict__ r = R.p;
    for (int i = n - 1; i >= 0; --i) {
      double s0 = x[i * X.rs], s1 = 0.0;
      int j = i + 1;
      for (; j + 1 < n; j += 2) {
        s0 = fma(-r[i * R.rs + j * R.cs], x[j * X.rs], s0);
        s1 = fma(-r[i * R.rs + (j + 1) * R.cs], x[(j + 1) * X.rs], s1);
      }
      if (j < n) s0 = fma(-r[i * R.rs + j * R.cs], x[j * X.rs], s0);
      x[i * X.rs] = (s0 + s1) / r[i * R.rs + i * R.cs];
    }
  }
  tm.sync();
}

struct UlvCtx {
  const UlvNode* nodes;
  const double* pool;  // primary generator pool
  double* fpool;       // factor pool (output)
  double* red;         // reduced generators of every node (input from the children, output for the parent)
  int32_t MI, NI, KR, KW;  // maxima over the nodes of this launch (one tree level): scratch sizing
  double* pivmin;      // per node: smallest |pivot| the node divides by (+inf if it eliminates nothing); may be NULL
};

// The factorisation divides by the diagonal of a triangular factor twice: L1 (ulvfactor.jl:48, trsm) and the
// root block (ulvfactor.jl:83, D \ b).  A zero (or non-finite) pivot is recorded here and turned into
// HSSB_ERR_SINGULAR by the caller, like the SingularException the reference throws.
HSSB_HD void ulv_record_pivot(const Team& tm, const UlvCtx& cx, int node, Mat Tri, int n) {
  if (cx.pivmin && tm.tid == 0) {
    double mn = INFINITY;
    for (int i = 0; i < n; ++i) {
      double a = fabs(Tri(i, i));
      if (!(a >= 0.0) || a == INFINITY) a = 0.0;  // NaN / Inf count as a breakdown
      mn = a < mn ? a : mn;
    }
    cx.pivmin[node] = mn;
  }
}

// doubles of scratch one team needs
HSSB_HD int64_t ulv_scratch_len(int64_t MI, int64_t NI, int64_t KR, int64_t KW) {
  const int64_t KX = KR > KW ? KR : KW;
  return 3 * MI * NI      // Din, Dq, L2
         + MI * KR        // Uin
         + 2 * NI * KW    // Vin, Vq
         + MI * MI        // Q
         + NI * NI        // P
         + (MI + KW) * MI // T = [T1; T2; T3]
         + 2 * MI * KX    // G12, G21
         + 64;
}

// [A1 | A2] = T * [E1 | E2] for `rows` rows of T (rows x m_in, m_in = k1 + k2), where the node's
// incoming right-hand side is  b = E1 c1 + E2 c2,  c_s = [b_s (k_s); u_s (kw_s)]:
//   A1 = [T[:, :k1], -T[:, k1:] G21]   (rows x (k1 + kw1))
//   A2 = [T[:, k1:], -T[:, :k1] G12]   (rows x (k2 + kw2))
HSSB_HD void ulv_fold(const Team& tm, const UlvNode& u, Mat T, int rows, Mat G12, Mat G21, Mat A1, Mat A2) {
  tm_copy(tm, A1, T, rows, u.k1);
  tm_gemm(tm, A1.sub(0, u.k1), T.sub(0, u.k1), G21, rows, u.kw1, u.k2, -1.0, 0.0);
  tm_copy(tm, A2, T.sub(0, u.k1), rows, u.k2);
  tm_gemm(tm, A2.sub(0, u.k2), T, G12, rows, u.kw2, u.k1, -1.0, 0.0);
}

// Factorises one node.  `s` is the team's scratch (ulv_scratch_len doubles).  FF: the plan is in fast
// form (UlvNode::g, pta_c, ptb_c); the default form is a separate instantiation so that its code does not
// change with the experimental one.
template <bool FF>
HSSB_HD void ulv_factor_node(const Team& tm, const UlvCtx& cx, int node, double* s) {
  const UlvNode& u = cx.nodes[node];
  const int m = u.m_in, n = u.n_in, kr = u.kr, kw = u.kw, k = u.k, mk = u.mk, no = u.n_out;
  const int64_t MI = cx.MI, NI = cx.NI, KR = cx.KR, KW = cx.KW, KX = KR > KW ? KR : KW;
  double* q = s;
  auto take = [&](int64_t len) { double* r = q; q += len; return r; };
  const Mat Din = colmajor(take(MI * NI), m > 0 ? m : 1);
  const Mat Dq = colmajor(take(MI * NI), m > 0 ? m : 1);
  const Mat L2 = colmajor(take(MI * NI), k > 0 ? k : 1);
  // the matrices tm_qr updates with ONE THREAD PER COLUMN are stored row-major: consecutive threads then touch consecutive
  // addresses in every step of the reflector loop (column-major, a warp's load hit 32 different cache lines for 32 doubles)
  const Mat Uin = rowmajor(take(MI * KR), kr > 0 ? kr : 1);
  const Mat Vin = colmajor(take(NI * KW), n > 0 ? n : 1);
  const Mat Vq = colmajor(take(NI * KW), n > 0 ? n : 1);
  const Mat Q = rowmajor(take(MI * MI), m > 0 ? m : 1);
  const Mat P = rowmajor(take(NI * NI), n > 0 ? n : 1);
  const Mat T = colmajor(take((MI + KW) * MI), m + kw > 0 ? m + kw : 1);
  const Mat G12 = colmajor(take(MI * KX), u.k1 > 0 ? u.k1 : 1);
  const Mat G21 = colmajor(take(MI * KX), u.k2 > 0 ? u.k2 : 1);

  // ---- incoming block
  if (u.is_leaf) {
    tm_copy(tm, Din, colmajor(cx.pool + u.D, u.ldD), m, n);
    if (!u.is_root) {
      tm_copy(tm, Uin, colmajor(cx.pool + u.U, u.ldU), m, kr);
      tm_copy(tm, Vin, colmajor(cx.pool + u.V, u.ldV).t(), n, kw);  // the pool holds V'
    }
  } else {  // merge of the children (ulvfactor.jl:74-79)
    const UlvNode& c1 = cx.nodes[u.left];
    const UlvNode& c2 = cx.nodes[u.right];
    const Mat D1 = colmajor(cx.red + c1.rD, u.k1 > 0 ? u.k1 : 1), D2 = colmajor(cx.red + c2.rD, u.k2 > 0 ? u.k2 : 1);
    const Mat U1 = colmajor(cx.red + c1.rU, u.k1 > 0 ? u.k1 : 1), U2 = colmajor(cx.red + c2.rU, u.k2 > 0 ? u.k2 : 1);
    const Mat V1 = colmajor(cx.red + c1.rV, u.no1 > 0 ? u.no1 : 1), V2 = colmajor(cx.red + c2.rV, u.no2 > 0 ? u.no2 : 1);
    tm_gemm(tm, G12, U1, colmajor(cx.pool + u.B12, u.ldB12), u.k1, u.kw2, u.kr1, 1.0, 0.0);  // U~1 B12
    tm_gemm(tm, G21, U2, colmajor(cx.pool + u.B21, u.ldB21), u.k2, u.kw1, u.kr2, 1.0, 0.0);  // U~2 B21
    tm_copy(tm, Din, D1, u.k1, u.no1);
    tm_copy(tm, Din.sub(u.k1, u.no1), D2, u.k2, u.no2);
    tm_gemm(tm, Din.sub(0, u.no1), G12, V2.t(), u.k1, u.no2, u.kw2, 1.0, 0.0);
    tm_gemm(tm, Din.sub(u.k1, 0), G21, V1.t(), u.k2, u.no1, u.kw1, 1.0, 0.0);
    if (!u.is_root) {
      tm_gemm(tm, Uin, U1, colmajor(cx.pool + u.R1, u.ldR1), u.k1, kr, u.kr1, 1.0, 0.0);
      tm_gemm(tm, Uin.sub(u.k1, 0), U2, colmajor(cx.pool + u.R2, u.ldR2), u.k2, kr, u.kr2, 1.0, 0.0);
      tm_gemm(tm, Vin, V1, colmajor(cx.pool + u.W1, u.ldW1).t(), u.no1, kw, u.kw1, 1.0, 0.0);  // the pool holds W'
      tm_gemm(tm, Vin.sub(u.no1, 0), V2, colmajor(cx.pool + u.W2, u.ldW2).t(), u.no2, kw, u.kw2, 1.0, 0.0);
    }
  }

  // ---- root: z[cols] = D \ b (ulvfactor.jl:83, :12) as an explicit inverse D^-1 = R^-1 Q'
  if (u.is_root) {
    tm_eye(tm, Q, m, m, 1.0);
    tm_qr(tm, Din, m, n, Q, m);
    ulv_record_pivot(tm, cx, node, Din, n);
    tm_trsm_upper(tm, Din, n, Q, m);  // Q <- R^-1 Q' = D^-1 (n x m, m == n)
    if (u.is_leaf) {
      tm_copy(tm, colmajor(cx.fpool + u.ac[0], u.ld_ac), Q, n, m);
    } else {
      ulv_fold(tm, u, Q, n, G12, G21, colmajor(cx.fpool + u.ac[0], u.ld_ac), colmajor(cx.fpool + u.ac[1], u.ld_ac));
    }
    return;
  }

  const Mat T1 = T, T23 = T.sub(mk, 0);  // T1: mk rows, T2: k rows, T3: kw rows
  if (mk > 0) {  // compressible: k = kr < m
    tm_eye(tm, Q, m, m, 1.0);
    tm_qr(tm, Uin, m, kr, Q, m);                                 // Q U = [R; 0]
    tm_copy(tm, colmajor(cx.red + u.rU, k > 0 ? k : 1), Uin, k, kr);  // U~ = R
    tm_gemm(tm, Dq, Q, Din, m, n, m, 1.0, 0.0);                  // Dq = Q D
    const Mat Dtop = Dq.sub(k, 0), Dbot = Dq;                    // mk x n, k x n
    tm_eye(tm, P, n, n, 1.0);
    tm_qr(tm, Dtop.t(), n, mk, P, n);                            // P Dtop' = [R1; 0]  =>  Dtop = [L1 0] P, L1 = R1' in place
    tm_gemm(tm, L2, Dbot, P.t(), k, n, n, 1.0, 0.0);             // L2 = Dbot P'
    tm_gemm(tm, Vq, P, Vin, n, kw, n, 1.0, 0.0);                 // Vq = P V
    tm_copy(tm, T1, Q.sub(k, 0), mk, m);                         // Qtop
    ulv_record_pivot(tm, cx, node, Dtop, mk);
    tm_trsm_lower(tm, Dtop, mk, T1, m);                          // T1 = L1^-1 Qtop
    tm_copy(tm, T23, Q, k, m);                                   // Qbot
    tm_gemm(tm, T23, L2, T1, k, m, mk, -1.0, 1.0);               // T2 = Qbot - L2a T1
    tm_gemm(tm, T23.sub(k, 0), Vq.t(), T1, kw, m, mk, 1.0, 0.0); // T3 = Vq_top' T1
    tm_copy(tm, colmajor(cx.red + u.rD, k > 0 ? k : 1), L2.sub(0, mk), k, no);
    tm_copy(tm, colmajor(cx.red + u.rV, no > 0 ? no : 1), Vq.sub(mk, 0), no, kw);
    if constexpr (!FF) {
      tm_copy(tm, colmajor(cx.fpool + u.pta, u.ld_pt), P.t(), n, mk);             // P'[:, :mk]
      tm_copy(tm, colmajor(cx.fpool + u.ptb, u.ld_pt), P.t().sub(0, mk), n, no);  // P'[:, mk:]
    } else {  // fast form (see UlvNode): the leaf output operator g = P'[:, :mk] T1, and P' split by child rows
      if (u.ptb >= 0) tm_copy(tm, colmajor(cx.fpool + u.ptb, u.ld_pt), P.t().sub(0, mk), n, no);
      if (u.g >= 0) tm_gemm(tm, colmajor(cx.fpool + u.g, u.ld_g), P.t(), T1, n, m, mk, 1.0, 0.0);
      for (int c = 0; c < 2; ++c) {
        const int r0 = c ? u.no1 : 0, rows = c ? u.no2 : u.no1;
        if (u.pta_c[c] >= 0) tm_copy(tm, colmajor(cx.fpool + u.pta_c[c], u.ld_ptc[c]), P.t().sub(r0, 0), rows, mk);
        if (u.ptb_c[c] >= 0) tm_copy(tm, colmajor(cx.fpool + u.ptb_c[c], u.ld_ptc[c]), P.t().sub(r0, mk), rows, no);
      }
    }
  } else {  // cannot be compressed (ulvfactor.jl:31-37): everything is handed to the parent
    ulv_record_pivot(tm, cx, node, Din, 0);
    tm_eye(tm, T23, k, m, 1.0);
    tm_eye(tm, T23.sub(k, 0), kw, m, 0.0);
    tm_copy(tm, colmajor(cx.red + u.rD, k > 0 ? k : 1), Din, k, no);
    tm_copy(tm, colmajor(cx.red + u.rU, k > 0 ? k : 1), Uin, k, kr);
    tm_copy(tm, colmajor(cx.red + u.rV, no > 0 ? no : 1), Vin, no, kw);
    tm_eye(tm, colmajor(cx.fpool + u.ptb, u.ld_pt), n, no, 1.0);
  }

  // ---- the matrices of the solve's upsweep
  if (u.is_leaf) {
    if (mk > 0 && (!FF || u.az[0] >= 0)) tm_copy(tm, colmajor(cx.fpool + u.az[0], u.ld_az), T1, mk, m);
    tm_copy(tm, colmajor(cx.fpool + u.ac[0], u.ld_ac), T23, k + kw, m);
  } else {
    if (mk > 0)
      ulv_fold(tm, u, T1, mk, G12, G21, colmajor(cx.fpool + u.az[0], u.ld_az), colmajor(cx.fpool + u.az[1], u.ld_az));
    const Mat A1 = colmajor(cx.fpool + u.ac[0], u.ld_ac), A2 = colmajor(cx.fpool + u.ac[1], u.ld_ac);
    ulv_fold(tm, u, T23, k + kw, G12, G21, A1, A2);
    // u += W1' u1 + W2' u2 (ulvfactor.jl:89); the pool holds W' (kw x kw_s)
    tm_add(tm, A1.sub(k, u.k1), colmajor(cx.pool + u.W1, u.ldW1), kw, u.kw1);
    tm_add(tm, A2.sub(k, u.k2), colmajor(cx.pool + u.W2, u.ldW2), kw, u.kw2);
  }
}

// One CTA per node of one level (grid-stride); scratch is per CTA.
__global__ void __launch_bounds__(256, 3)
ulv_factor_kernel(UlvCtx cx, const int32_t* __restrict__ level_nodes, int count, double* scratch, int64_t scratch_stride) {
  const Team tm{(int)threadIdx.x, (int)blockDim.x};
  double* s = scratch + (int64_t)blockIdx.x * scratch_stride;
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    ulv_factor_node<false>(tm, cx, level_nodes[i], s);
    __syncthreads();
  }
}

// the same for a plan in fast form (HSSB_OPT_ULV_FAST)
__global__ void __launch_bounds__(256, 4)
ulv_factor_kernel_ff(UlvCtx cx, const int32_t* __restrict__ level_nodes, int count, double* scratch, int64_t scratch_stride) {
  const Team tm{(int)threadIdx.x, (int)blockDim.x};
  double* s = scratch + (int64_t)blockIdx.x * scratch_stride;
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    ulv_factor_node<true>(tm, cx, level_nodes[i], s);
    __syncthreads();
  }
}

}  // namespace hssb
