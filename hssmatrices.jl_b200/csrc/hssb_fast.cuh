// hssb_fast.cuh — fixed-shape FP64 tensor-core (DMMA) kernels for uniform trees
// (square leaves of one size M, one rank R everywhere: BASELINE configs 3-5).
//
// sm_100a has no tcgen05 kind for FP64, so the tensor path is the warp-level
// mma.sync m8n8k4 (DMMA.8x8x4 in SASS).  All operands are staged in shared
// memory with a leading dimension == 4 (mod 16) doubles, which makes every
// fragment load of hssb_mma.cuh bank-conflict free:
//   "N" operand  A(i,k) at s[k*ld + i]   (D, U, B12, B21, R)        lane -> s[(k0+t)*ld + i0+g]
//   "T" operand  A(i,k) at s[i*ld + k]   (V', W', and every B(k,j) = s[j*ld + k])
// Global -> shared copies are 16-byte cp.async (LDGSTS), column by column, so
// arbitrary pool offsets work as long as columns are 16-byte aligned (the
// packer guarantees it).
//
// Kernels
//   leaf_up_kernel<M,R>    Z  = V' X          persistent, X resident, V streamed by column chunks
//   merge_kernel<R>        Z  = W1' Z1 + W2' Z2                    one CTA per (node, column tile)
//   translate_kernel<R>    F1 = B12 Z2 + R1 F ; F2 = B21 Z1 + R2 F  one CTA per (parent, column tile)
//   leaf_down_kernel<M,R>  Y  = a (D X + U F) + b Y   persistent, X resident, [D U] streamed by K chunks
#pragma once

#include "hssb_internal.h"
#include "hssb_mma.cuh"

namespace hssb {

enum FastKind : int { FAST_NONE = 0, FAST_LEAF_UP = 1, FAST_MERGE = 2, FAST_TRANSLATE = 3, FAST_LEAF_DOWN = 4 };

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Copies a rows x cols column-major block (rows even, columns 16-byte aligned)
// from global (leading dimension lds) to shared (leading dimension ldd).
template <int THREADS>
__device__ __forceinline__ void copy_block_async(double* dst, int ldd, const double* src, int64_t lds, int rows, int cols,
                                                 int tid) {
  const int h = rows >> 1;  // 16-byte chunks per column
  for (int idx = tid; idx < h * cols; idx += THREADS) {
    const int c = idx / h, r2 = (idx - c * h) << 1;
    cp_async16(dst + c * ldd + r2, src + (int64_t)c * lds + r2);
  }
}

// ------------------------------------------------------------- leaf shapes ---
template <int M, int R>
struct LeafCfg {
  static constexpr int WM = M / 32;       // warps along the leaf rows (leaf-down)
  static constexpr int WN = 8 / WM;       // warps along the right-hand sides
  static constexpr int NT = 32 * WN;      // right-hand sides per tile
  static constexpr int LDX = M + 4, LDF = R + 4, LDA = M + 4;
  // leaf-down: [D U] streamed in K chunks of KC columns (M x KC, contiguous in the pool)
  static constexpr int KC = (M >= 256) ? 8 : 16;
  static constexpr int NSTAGE = 3;
  static constexpr int NCH_D = M / KC, NCH_U = R / KC, NCH = NCH_D + NCH_U;
  static constexpr int XPIECE = (NT + (NCH - 2) - 1) / (NCH - 2);  // X columns prefetched per chunk group
  static constexpr size_t DOWN_SMEM = sizeof(double) * (2 * NT * LDX + NT * LDF + NSTAGE * KC * LDA);
  // leaf-up: V streamed in chunks of VC columns (M x VC, contiguous); one 8x8 output tile per warp and chunk
  static constexpr int VC = 64 / (NT / 8);  // (NT/8) * (VC/8) == 8 tiles per chunk
  static constexpr int NVC = R / VC;
  static constexpr int UP_STAGES = (M >= 256) ? 2 : (NVC + 1 < 4 ? NVC + 1 : 4);  // needs UP_STAGES - 1 <= NVC
  static constexpr size_t UP_SMEM = sizeof(double) * (2 * NT * LDX + UP_STAGES * VC * LDA);
  static_assert(M % 32 == 0 && WM * WN == 8 && R % KC == 0 && VC % 8 == 0 && R % VC == 0 && NCH_D > 2 &&
                    UP_STAGES >= 2 && UP_STAGES - 1 <= NVC, "unsupported leaf shape");
};

// =============================================================== leaf down ===
// Y[M x NT] = alpha * ([D U] * [X; F]) + beta * Y for one (leaf, column tile) item at a time.
// Persistent CTA (one per SM), 8 warps, warp tile 32 x 32 (4x4 DMMA tiles, 32 accumulators/lane).
// Pipeline of cp.async groups, one per K chunk: group (item, c) carries the A chunk c, the F block
// of the item (c == 2) and a slice of the NEXT item's X block (c >= 2), so the X block of item i+1
// is resident before item i ends and the FP64 pipe never waits at an item boundary.
template <int M, int R>
__global__ void __launch_bounds__(256, 1)
leaf_down_kernel(const GTask* __restrict__ tasks, int ntasks, int ntiles, CallParams p) {
  using C = LeafCfg<M, R>;
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;                       // [2][NT][LDX]
  double* Fs = Xs + 2 * C::NT * C::LDX;    // [NT][LDF]
  double* As = Fs + C::NT * C::LDF;        // [NSTAGE][KC][LDA]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp % C::WM, wn = warp / C::WM;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  if (my <= 0) return;
  const int G = my * C::NCH;
  const int nrhs = p.nrhs;

  auto item_cols = [&](int item) { return min(C::NT, nrhs - ((first + item) % ntiles) * C::NT); };
  auto load_x = [&](int item, int c0, int c1) {  // columns [c0, c1) of the item's X block
    const GTask& tk = tasks[(first + item) / ntiles];
    const int tile = (first + item) % ntiles;
    c1 = min(c1, item_cols(item));
    if (c0 >= c1) return;
    const double* src = p.X + tk.b0 + (int64_t)(tile * C::NT + c0) * p.ldx;
    copy_block_async<256>(Xs + (item & 1) * C::NT * C::LDX + c0 * C::LDX, C::LDX, src, p.ldx, M, c1 - c0, tid);
  };
  auto issue_group = [&](int gi) {
    if (gi < G) {
      const int item = gi / C::NCH, c = gi - item * C::NCH;
      const GTask& tk = tasks[(first + item) / ntiles];
      const double* asrc = p.pool + (c < C::NCH_D ? tk.a0 + (int64_t)c * C::KC * M : tk.a1 + (int64_t)(c - C::NCH_D) * C::KC * M);
      copy_block_async<256>(As + (gi % C::NSTAGE) * C::KC * C::LDA, C::LDA, asrc, M, M, C::KC, tid);
      if (c == 2) {
        const int tile = (first + item) % ntiles;
        const double* fsrc = p.F + tk.b1 * (int64_t)nrhs + (int64_t)tile * C::NT * R;
        copy_block_async<256>(Fs, C::LDF, fsrc, R, R, item_cols(item), tid);
      }
      if (c >= 2 && item + 1 < my) load_x(item + 1, (c - 2) * C::XPIECE, (c - 1) * C::XPIECE);
    }
    cp_async_commit();
  };

  // prologue: X of the first item, then the first NSTAGE-1 chunk groups
  load_x(0, 0, C::NT);
  cp_async_commit();
  issue_group(0);
  issue_group(1);

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int gi = 0; gi < G; ++gi) {
    cp_async_wait<1>();   // group gi (and everything older) has landed
    __syncthreads();      // ... for every thread; and chunk gi-1 has been consumed by all warps
    issue_group(gi + 2);  // refills the stage consumed at gi-1
    const int item = gi / C::NCH, c = gi - item * C::NCH;
    const double* A = As + (gi % C::NSTAGE) * C::KC * C::LDA + wm * 32 + g;
    const double* B;
    int ldb;
    if (c < C::NCH_D) {
      B = Xs + (item & 1) * C::NT * C::LDX + (wn * 32 + g) * C::LDX + c * C::KC + t;
      ldb = C::LDX;
    } else {
      B = Fs + (wn * 32 + g) * C::LDF + (c - C::NCH_D) * C::KC + t;
      ldb = C::LDF;
    }
#pragma unroll
    for (int kk = 0; kk < C::KC / 4; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = A[(kk * 4 + t) * C::LDA + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[j * 8 * ldb + kk * 4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (c == C::NCH - 1) {  // epilogue of the item: Y = alpha*acc + beta*Y (beta == 0 never reads Y)
      const GTask& tk = tasks[(first + item) / ntiles];
      const int tile = (first + item) % ntiles;
      const int ncols = item_cols(item);
      double* Y = p.Y + tk.c + (int64_t)tile * C::NT * p.ldy;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = wn * 32 + j * 8 + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = wm * 32 + i * 8 + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (col + e < ncols) {
              double* dst = Y + (int64_t)(col + e) * p.ldy + row;
              double v = p.alpha * acc[i][j][e];
              if (p.beta != 0.0) v += p.beta * (*dst);
              *dst = v;
            }
            acc[i][j][e] = 0.0;
          }
        }
      }
    }
  }
  cp_async_wait<0>();
}

// ================================================================= leaf up ===
// Z'[NT x R] = X'[NT x M] * V[M x R] for one (leaf, column tile) item at a time; X resident
// (double buffered across items), V streamed in contiguous chunks of VC columns.  Each chunk
// yields a complete NT x VC slab of Z' = (NT/8)*(VC/8) = 8 DMMA tiles, one per warp, K = M.
template <int M, int R>
__global__ void __launch_bounds__(256, 1)
leaf_up_kernel(const GTask* __restrict__ tasks, int ntasks, int ntiles, CallParams p) {
  using C = LeafCfg<M, R>;
  constexpr int S = C::UP_STAGES;
  extern __shared__ __align__(16) double smem[];
  double* Xs = smem;                     // [2][NT][LDX]
  double* Vs = Xs + 2 * C::NT * C::LDX;  // [S][VC][LDA]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int JT = C::NT / 8;          // column tiles of X per chunk
  const int jt = warp % JT, vt = warp / JT;  // this warp's tile: X columns jt*8.., V columns vt*8.. of the chunk
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  if (my <= 0) return;
  const int G = my * C::NVC;
  const int nrhs = p.nrhs;

  auto item_cols = [&](int item) { return min(C::NT, nrhs - ((first + item) % ntiles) * C::NT); };
  auto load_x = [&](int item) {
    const GTask& tk = tasks[(first + item) / ntiles];
    const int tile = (first + item) % ntiles;
    const double* src = p.X + tk.b0 + (int64_t)tile * C::NT * p.ldx;
    copy_block_async<256>(Xs + (item & 1) * C::NT * C::LDX, C::LDX, src, p.ldx, M, item_cols(item), tid);
  };
  auto issue_group = [&](int gi) {
    if (gi < G) {
      const int item = gi / C::NVC, c = gi - item * C::NVC;
      const GTask& tk = tasks[(first + item) / ntiles];
      copy_block_async<256>(Vs + (gi % S) * C::VC * C::LDA, C::LDA, p.pool + tk.a0 + (int64_t)c * C::VC * M, M, M, C::VC, tid);
    }
    cp_async_commit();
  };

  load_x(0);
  cp_async_commit();
#pragma unroll
  for (int s = 0; s < S - 1; ++s) issue_group(s);

  for (int gi = 0; gi < G; ++gi) {
    cp_async_wait<S - 2>();
    __syncthreads();
    const int item = gi / C::NVC, c = gi - item * C::NVC;
    // Prefetch the next item's X block at the first chunk of this item: the buffer it overwrites
    // was last read by item-1, which every warp left before the barrier above.  It joins the
    // cp.async group committed by issue_group() below, i.e. group gi+S-1 <= (item+1)*NVC, so it
    // has landed when item+1 starts (UP_STAGES - 1 <= NVC).
    if (c == 0 && item + 1 < my) load_x(item + 1);
    issue_group(gi + S - 1);
    const double* A = Xs + (item & 1) * C::NT * C::LDX + (jt * 8 + g) * C::LDX + t;   // X'(j, k)
    const double* B = Vs + (gi % S) * C::VC * C::LDA + (vt * 8 + g) * C::LDA + t;     // V(k, n)
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};  // two independent chains over K
#pragma unroll 8
    for (int kk = 0; kk < M / 4; kk += 2) {
      mma_m8n8k4(c0[0], c0[1], A[kk * 4], B[kk * 4]);
      mma_m8n8k4(c1[0], c1[1], A[kk * 4 + 4], B[kk * 4 + 4]);
    }
    // Z'(j, n) -> Z(n, j): lane holds n = 2t, 2t+1 (adjacent rows of Z) for column j = g
    const GTask& tk = tasks[(first + item) / ntiles];
    const int tile = (first + item) % ntiles;
    const int col = jt * 8 + g;
    if (col < item_cols(item)) {
      double* Z = p.Z + tk.c * (int64_t)nrhs + (int64_t)(tile * C::NT + col) * R + c * C::VC + vt * 8 + 2 * t;
      *reinterpret_cast<double2*>(Z) = make_double2(c0[0] + c1[0], c0[1] + c1[1]);
    }
  }
  cp_async_wait<0>();
}

// =================================================================== merge ===
// Z[R x NT] = W1'[R x R] Z1 + W2'[R x R] Z2 for one (node, column tile); 128 threads,
// warp w owns columns w*16 .. w*16+15 of the tile (R/8 x 2 DMMA tiles).
template <int R>
struct NodeCfg {
  static constexpr int NT = 64;
  static constexpr int LD = R + 4;
  static constexpr size_t MERGE_SMEM = sizeof(double) * (2 * R * LD + 2 * NT * LD);
  static constexpr size_t TRANS_SMEM = sizeof(double) * (4 * R * LD + 3 * NT * LD);
};

template <int R>
__global__ void __launch_bounds__(128)
merge_kernel(const GTask* __restrict__ tasks, CallParams p) {
  using C = NodeCfg<R>;
  extern __shared__ __align__(16) double smem[];
  double* Ws = smem;                   // [2][R][LD]   W(k, i) at Ws[i*LD + k]  ("T" operand)
  double* Zs = Ws + 2 * R * C::LD;     // [2][NT][LD]
  const GTask tk = tasks[blockIdx.x];
  const int tile = blockIdx.y, nrhs = p.nrhs;
  const int ncols = min(C::NT, nrhs - tile * C::NT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

  copy_block_async<128>(Ws, C::LD, p.pool + tk.a0, R, R, R, tid);
  copy_block_async<128>(Ws + R * C::LD, C::LD, p.pool + tk.a1, R, R, R, tid);
  copy_block_async<128>(Zs, C::LD, p.Z + tk.b0 * (int64_t)nrhs + (int64_t)tile * C::NT * R, R, R, ncols, tid);
  copy_block_async<128>(Zs + C::NT * C::LD, C::LD, p.Z + tk.b1 * (int64_t)nrhs + (int64_t)tile * C::NT * R, R, R, ncols, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  double acc[R / 8][2][2];
#pragma unroll
  for (int i = 0; i < R / 8; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const double* A = Ws + s * R * C::LD + g * C::LD + t;
    const double* B = Zs + s * C::NT * C::LD + (warp * 16 + g) * C::LD + t;
#pragma unroll
    for (int kk = 0; kk < R / 4; ++kk) {
      double b[2] = {B[kk * 4], B[8 * C::LD + kk * 4]};
#pragma unroll
      for (int i = 0; i < R / 8; ++i) {
        const double a = A[i * 8 * C::LD + kk * 4];
        mma_m8n8k4(acc[i][0][0], acc[i][0][1], a, b[0]);
        mma_m8n8k4(acc[i][1][0], acc[i][1][1], a, b[1]);
      }
    }
  }
  double* Z = p.Z + tk.c * (int64_t)nrhs + (int64_t)tile * C::NT * R;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = warp * 16 + j * 8 + 2 * t + e;
      if (col < ncols) {
#pragma unroll
        for (int i = 0; i < R / 8; ++i) Z[(int64_t)col * R + i * 8 + g] = acc[i][j][e];
      }
    }
}

// =============================================================== translate ===
// For one parent and column tile: F1 = B12 Z2 (+ R1 F), F2 = B21 Z1 (+ R2 F).  Tasks come in
// (left child, right child) pairs; 256 threads, warps 0-3 -> F1, warps 4-7 -> F2.
template <int R>
__global__ void __launch_bounds__(256)
translate_kernel(const GTask* __restrict__ tasks, CallParams p) {
  using C = NodeCfg<R>;
  extern __shared__ __align__(16) double smem[];
  double* Bs = smem;                   // [2][R][LD]  B12, B21   A(i,k) at [k*LD + i]  ("N" operand)
  double* Rs = Bs + 2 * R * C::LD;     // [2][R][LD]  R1, R2
  double* Zs = Rs + 2 * R * C::LD;     // [2][NT][LD] Z of the sibling of child 0 / child 1
  double* Fs = Zs + 2 * C::NT * C::LD; // [NT][LD]    F of the parent
  const GTask t0 = tasks[2 * blockIdx.x], t1 = tasks[2 * blockIdx.x + 1];
  const int tile = blockIdx.y, nrhs = p.nrhs;
  const int ncols = min(C::NT, nrhs - tile * C::NT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const bool has_f = t0.K1 > 0;
  const int64_t toff = (int64_t)tile * C::NT * R;

  copy_block_async<256>(Bs, C::LD, p.pool + t0.a0, R, R, R, tid);
  copy_block_async<256>(Bs + R * C::LD, C::LD, p.pool + t1.a0, R, R, R, tid);
  copy_block_async<256>(Zs, C::LD, p.Z + t0.b0 * (int64_t)nrhs + toff, R, R, ncols, tid);
  copy_block_async<256>(Zs + C::NT * C::LD, C::LD, p.Z + t1.b0 * (int64_t)nrhs + toff, R, R, ncols, tid);
  if (has_f) {
    copy_block_async<256>(Rs, C::LD, p.pool + t0.a1, R, R, R, tid);
    copy_block_async<256>(Rs + R * C::LD, C::LD, p.pool + t1.a1, R, R, R, tid);
    copy_block_async<256>(Fs, C::LD, p.F + t0.b1 * (int64_t)nrhs + toff, R, R, ncols, tid);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int child = warp >> 2, w = warp & 3;
  double acc[R / 8][2][2];
#pragma unroll
  for (int i = 0; i < R / 8; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nsrc = has_f ? 2 : 1;
  for (int s = 0; s < nsrc; ++s) {
    const double* A = (s ? Rs : Bs) + child * R * C::LD + g;
    const double* B = (s ? Fs : Zs + child * C::NT * C::LD) + (w * 16 + g) * C::LD + t;
#pragma unroll
    for (int kk = 0; kk < R / 4; ++kk) {
      double b[2] = {B[kk * 4], B[8 * C::LD + kk * 4]};
#pragma unroll
      for (int i = 0; i < R / 8; ++i) {
        const double a = A[(kk * 4 + t) * C::LD + i * 8];
        mma_m8n8k4(acc[i][0][0], acc[i][0][1], a, b[0]);
        mma_m8n8k4(acc[i][1][0], acc[i][1][1], a, b[1]);
      }
    }
  }
  double* F = p.F + (child ? t1.c : t0.c) * (int64_t)nrhs + toff;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = w * 16 + j * 8 + 2 * t + e;
      if (col < ncols) {
#pragma unroll
        for (int i = 0; i < R / 8; ++i) F[(int64_t)col * R + i * 8 + g] = acc[i][j][e];
      }
    }
}

// ================================================================ host side ===
struct FastState {
  int num_sms = 148;
  std::vector<const void*> configured;  // kernels whose dynamic shared-memory limit has been raised on this device
  int configure(const void* fn, size_t smem) {
    for (const void* f : configured)
      if (f == fn) return HSSB_OK;
    HSSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.push_back(fn);
    return HSSB_OK;
  }
};

template <int M, int R>
static int launch_leaf(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st, bool down) {
  using C = LeafCfg<M, R>;
  FastState* fs = (FastState*)H->fast_state;
  const int ntiles = (cp.nrhs + C::NT - 1) / C::NT;
  const int nitems = (int)ph.ntasks * ntiles;
  const int grid = std::min(nitems, fs->num_sms);
  if (down) {
    if (int rc = fs->configure((const void*)leaf_down_kernel<M, R>, C::DOWN_SMEM)) return rc;
    leaf_down_kernel<M, R><<<grid, 256, C::DOWN_SMEM, st>>>(H->tasks_dev + ph.task0, (int)ph.ntasks, ntiles, cp);
  } else {
    if (int rc = fs->configure((const void*)leaf_up_kernel<M, R>, C::UP_SMEM)) return rc;
    leaf_up_kernel<M, R><<<grid, 256, C::UP_SMEM, st>>>(H->tasks_dev + ph.task0, (int)ph.ntasks, ntiles, cp);
  }
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

template <int R>
static int launch_node(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  using C = NodeCfg<R>;
  FastState* fs = (FastState*)H->fast_state;
  const int ntiles = (cp.nrhs + C::NT - 1) / C::NT;
  if (ph.fast == FAST_MERGE) {
    if (int rc = fs->configure((const void*)merge_kernel<R>, C::MERGE_SMEM)) return rc;
    merge_kernel<R><<<dim3((unsigned)ph.ntasks, (unsigned)ntiles), 128, C::MERGE_SMEM, st>>>(H->tasks_dev + ph.task0, cp);
  } else {
    if (int rc = fs->configure((const void*)translate_kernel<R>, C::TRANS_SMEM)) return rc;
    translate_kernel<R><<<dim3((unsigned)(ph.ntasks / 2), (unsigned)ntiles), 256, C::TRANS_SMEM, st>>>(H->tasks_dev + ph.task0, cp);
  }
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

static bool fast_shape_supported(int64_t m, int64_t r) {
  return (m == 128 && (r == 16 || r == 32 || r == 64)) || (m == 256 && (r == 16 || r == 32 || r == 64));
}

// Tags the phases of a uniform tree that a fixed-shape kernel can run.
static void plan_fast_phases(hssb_matrix* H) {
  if (!H->uniform || !fast_shape_supported(H->uni_m, H->uni_r)) return;
  const int m = (int)H->uni_m, r = (int)H->uni_r;
  for (Phase& ph : H->phases) {
    if (ph.kind == PH_EXCHANGE || ph.ntasks == 0) continue;
    bool ok = true;
    const GTask* tk = H->tasks_host.data() + ph.task0;
    for (int64_t i = 0; i < ph.ntasks && ok; ++i) {
      const GTask& g = tk[i];
      switch (ph.kind) {
        case PH_LEAF_UP: ok = g.M == r && g.K0 == m && g.lda0 == m && g.ldc == r && g.a0 >= 0; break;
        case PH_MERGE: ok = g.M == r && g.K0 == r && g.K1 == r && g.lda0 == r && g.lda1 == r && g.ldb0 == r && g.ldb1 == r && g.ldc == r; break;
        case PH_TRANSLATE:
          ok = g.M == r && g.K0 == r && (g.K1 == r || g.K1 == 0) && g.lda0 == r && g.ldb0 == r && g.ldc == r && (g.K1 == 0 || (g.lda1 == r && g.ldb1 == r));
          if (ok && (i & 1)) ok = g.b1 == tk[i - 1].b1 && g.K1 == tk[i - 1].K1 && g.sb1 == tk[i - 1].sb1;  // sibling pair of one parent
          break;
        case PH_LEAF_DOWN: ok = g.M == m && g.K0 == m && g.K1 == r && g.lda0 == m && g.lda1 == m && g.ldb1 == r && g.a0 >= 0 && g.a1 >= 0; break;
        default: ok = false;
      }
    }
    if (ph.kind == PH_TRANSLATE && (ph.ntasks & 1)) ok = false;
    if (!ok) continue;
    ph.fast = ph.kind == PH_LEAF_UP ? FAST_LEAF_UP : ph.kind == PH_MERGE ? FAST_MERGE : ph.kind == PH_TRANSLATE ? FAST_TRANSLATE : FAST_LEAF_DOWN;
  }
}

static bool fast_phase_supported(const hssb_matrix* H, const Phase& ph, const CallParams& cp) {
  (void)H;
  if (ph.fast == FAST_LEAF_UP || ph.fast == FAST_LEAF_DOWN) {
    // 16-byte cp.async on X columns: base and leading dimension must be 16-byte aligned
    if (((uintptr_t)cp.X & 15) || (cp.ldx & 1)) return false;
  }
  return true;
}

static int launch_fast(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  if (!H->fast_state) {
    FastState* fs = new FastState();
    cudaDeviceGetAttribute(&fs->num_sms, cudaDevAttrMultiProcessorCount, H->device);
    H->fast_state = fs;
  }
  const int m = (int)H->uni_m, r = (int)H->uni_r;
  if (ph.fast == FAST_MERGE || ph.fast == FAST_TRANSLATE) {
    switch (r) {
      case 16: return launch_node<16>(H, ph, cp, st);
      case 32: return launch_node<32>(H, ph, cp, st);
      case 64: return launch_node<64>(H, ph, cp, st);
    }
  } else {
    const bool down = ph.fast == FAST_LEAF_DOWN;
#define HSSB_LEAF_CASE(MM, RR) if (m == MM && r == RR) return launch_leaf<MM, RR>(H, ph, cp, st, down);
    HSSB_LEAF_CASE(128, 16) HSSB_LEAF_CASE(128, 32) HSSB_LEAF_CASE(128, 64)
    HSSB_LEAF_CASE(256, 16) HSSB_LEAF_CASE(256, 32) HSSB_LEAF_CASE(256, 64)
#undef HSSB_LEAF_CASE
  }
  HSSB_FAIL(HSSB_ERR_STATE, "no fixed-shape kernel for leaf %d rank %d", m, r);
}

static void free_fast(hssb_matrix* H) {
  delete (FastState*)H->fast_state;
  H->fast_state = nullptr;
}

}  // namespace hssb
