// hssb_fast.cuh — fixed-shape FP64 tensor-core (DMMA) kernels for uniform trees
// (square leaves of one size M, one rank R everywhere: BASELINE configs 3-5).
//
// sm_100a has no tcgen05 kind for FP64, so the tensor path is the warp-level
// mma.sync m8n8k4 (DMMA.8x8x4 in SASS).  All operands are staged in shared
// memory with a leading dimension == 4 (mod 16) doubles, which makes every
// fragment load of hssb_mma.cuh bank-conflict free:
//   A operands  A(i,k) at s[k*ld + i]   (D, U, V', W', B12, B21, R)   lane -> s[(k0+t)*ld + i0+g]
//   B operands  B(k,j) at s[j*ld + k]   (Z, F tiles)                  lane -> s[(j0+g)*ld + k0+t]
// The pool and the workspaces already carry that padding, so global -> shared is ONE
// cp.async.bulk (UBLKCP) per block; the caller's X comes in through a tiled 2-D TMA tensor map
// with 128-byte swizzle (UTMALDG).  Completion is tracked in bytes on mbarriers.
//
// Kernels
//   stream_leaf_kernel<M,R,false>  Z = V' X               persistent, warp-specialised (TMA producer warp)
//   stream_leaf_kernel<M,R,true>   Y = a (D X + U F) + b Y   same kernel, [D U] streamed by K chunks
//   stream_node_kernel<R>          Z = W1' Z1 + W2' Z2  and  F1 = B12 Z2 + R1 F   persistent, warp-specialised
//   oneshot_node_kernel<R>         same, one CTA per item (small ranks, single column tile)
#pragma once

#include <cuda.h>
#include <type_traits>

#include "hssb_internal.h"
#include "hssb_mma.cuh"

namespace hssb {

enum FastKind : int { FAST_NONE = 0, FAST_LEAF_UP = 1, FAST_MERGE = 2, FAST_TRANSLATE = 3, FAST_LEAF_DOWN = 4 };

// ------------------------------------------------- TMA bulk copy + mbarrier ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug must trap (and surface as a CUDA error), never hang the GPU.  Before the trap the
// waiter says what it was waiting for in host-mapped memory (the context is gone afterwards, host memory is not);
// set_error() appends it to the CUDA error text.
enum TrapCode : unsigned long long { TRAP_MBAR = 1, TRAP_GRID = 2, TRAP_PEER_ACK = 3, TRAP_PEER_DATA = 4, TRAP_FLOW = 5 };
static __device__ unsigned long long* g_trap_slot = nullptr;
__device__ __noinline__ void trap_report(unsigned long long code, unsigned long long a, unsigned long long b) {
  unsigned long long* s = g_trap_slot;
  if (s) {
    s[1] = a; s[2] = b; s[3] = (unsigned long long)blockIdx.x | ((unsigned long long)threadIdx.x << 32);
    __threadfence_system();
    s[0] = code;
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) trap_report(TRAP_MBAR, smem_u32(bar), parity);
  }
}
// 1-D bulk copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled TMA load (UTMALDG) of a {16 rows x NT columns} box of the user's X through a tensor
// map with 128-byte swizzle: the box lands as NT rows of 128 bytes whose 16-byte chunks are XORed
// with (row & 7), which makes the B-fragment reads conflict free without padding.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// Programmatic dependent launch (HSSB_OPT_PDL): a kernel launched with the attribute may START while its predecessor in
// the stream is still running -- its CTAs are placed as soon as every CTA of the predecessor has passed
// launch_dependents or exited -- and blocks in griddepcontrol.wait until the predecessor has completed and its memory
// is visible.  Every kernel of the level schedule does both first thing, so the order of the schedule is unchanged; what
// goes away is the drain / launch gap between the 27 kernels of a product (and whatever a kernel does before its wait:
// barrier initialisation, the bulk copies of generator blocks, which nobody produces).  Without the attribute both
// instructions are no-ops.
// (pdl_launch_dependents / pdl_wait: hssb_kernels_generic.cuh)

// n-index of a DMMA tile -> right-hand side within its 8-column group (see stream_leaf_kernel)
__device__ __forceinline__ int perm8(int n) { return (0x74216530u >> (4 * n)) & 7; }

// ------------------------------------------------------------- leaf shapes ---
// Both leaf kernels are one template: a persistent, warp-specialised, streamed-A GEMM
//     OUT[MO x NT] = [A0 | A1] * [X ; F]           (K = K0 + K1)
//   leaf down:  MO = M, A0 = D (M x M),   A1 = U (M x R), OUT -> Y = alpha*OUT + beta*Y
//   leaf up:    MO = R, A0 = V' (R x M),  no A1,          OUT -> Z
// X (K0 x NT, the user's columns) is resident and double buffered across items; [A0 | A1] is
// streamed through an NSTAGE ring in chunks of KC columns (MO x KC, contiguous in the pool).
// Warp 8 is the producer: it issues one cp.async.bulk per column (so the shared-memory image can
// carry the conflict-free +4 padding) and tracks completion in bytes on mbarriers.  Warps 0-7
// only wait, load fragments and issue DMMAs; they hand buffers back through "empty" mbarriers.
#ifndef HSSB_ENABLE_DEBUG_MODES
#define HSSB_ENABLE_DEBUG_MODES 0
#endif
#ifndef HSSB_DOWN_WARPS
#define HSSB_DOWN_WARPS 8
#endif
template <int M, int R, bool DOWN, int NT_>
struct StreamCfg {
  static constexpr int MO = DOWN ? M : R;
  static constexpr int K0 = M, K1 = DOWN ? R : 0;
  static constexpr int NT = NT_;  // right-hand sides per tile: 8192 / M, or 32 for M = 128 when nrhs <= 32
  static constexpr int NWARPS = DOWN ? HSSB_DOWN_WARPS : 8;           // consumer warps (one extra warp produces)
  static constexpr int WR = DOWN ? M / 32 : ((R >= 32 ? 2 : 1) > 64 / NT ? (R >= 32 ? 2 : 1) : 64 / NT);  // warps along OUT rows
  static constexpr int WC = NWARPS / WR;                             // warps along right-hand sides
  static constexpr int TM = MO / WR / 8, TN = NT / WC / 8;           // DMMA tiles per warp
  static constexpr int KC = DOWN ? (M >= 256 ? 8 : 16) : (4096 / R > M ? M : 4096 / R);
  static constexpr int KSTEPS = KC / 4;
  static constexpr int NCH0 = K0 / KC, NCH1 = K1 / KC, NCH = NCH0 + NCH1;
  static constexpr int LDF = K1 + 4, LDA = MO + 4;
  static constexpr int XSLABS = K0 / 16;                 // X block = XSLABS boxes of {16 rows x NT cols}, 128B-swizzled
  static constexpr int XBUF = K0 * NT;                   // doubles per X buffer (dense)
  // X blocks in flight: double buffered; the (memory-bound) leaf-up kernel with narrow tiles has few
  // DMMAs per item, so it needs more items in flight to cover the load latency
  static constexpr int XBUFS = DOWN ? 2 : (NT <= 16 ? 4 : ((NT <= 32 && M <= 128) ? 3 : 2));
  static constexpr int BAR_BYTES = 256;
  static constexpr int FIXED_BYTES = BAR_BYTES + 8 * (XBUFS * XBUF + (K1 ? NT * LDF : 0));
  static constexpr int STAGE_BYTES = 8 * KC * LDA;
  static constexpr int FIT = (232448 - FIXED_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = DOWN ? (FIT > 6 ? 6 : FIT) : (FIT > XBUFS ? XBUFS : FIT);  // as deep a ring as shared memory allows
  // F(i) and the first slice of X(i+1) are issued after chunk CX of item i; by then every
  // consumer has left item i-1 (the producer can be at most NSTAGE chunks ahead), so the
  // waits on f_empty / x_empty never hold up the A ring.
  static constexpr int CX = (NSTAGE < NCH0 - 1) ? NSTAGE : NCH0 - 1;
  static constexpr int XPIECES = (NCH - CX) < XSLABS ? (NCH - CX) : XSLABS;  // X(i+1) is spread over this many chunks
  static constexpr size_t SMEM = FIXED_BYTES + (size_t)NSTAGE * STAGE_BYTES;
  static_assert(MO % (8 * WR) == 0 && NT % (8 * WC) == 0 && TM >= 1 && TN >= 1, "warp tiling");
  static_assert(K0 % KC == 0 && K1 % KC == 0 && KC % 4 == 0 && NCH0 >= 1 && NSTAGE >= 2 && XPIECES >= 1 && K0 % 16 == 0, "chunking");
  static_assert(KSTEPS == 2 || KSTEPS % 4 == 0, "k-steps per chunk must be 2 or a multiple of 4 (swizzle bookkeeping)");
  static_assert(SMEM <= 232448 && 2 * NSTAGE + 2 * XBUFS + 2 <= BAR_BYTES / 8, "shared memory budget");
};

template <int M, int R, bool DOWN, int NT_>
__global__ void __launch_bounds__(StreamCfg<M, R, DOWN, NT_>::NWARPS * 32 + 32, 1)
stream_leaf_kernel(const GTask* __restrict__ tasks, int ntasks, int ntiles, CallParams p, const __grid_constant__ CUtensorMap xmap) {
  using C = StreamCfg<M, R, DOWN, NT_>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // X buffers first: the 128-byte swizzle needs 1024-byte aligned boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;                    // [NSTAGE]
  uint64_t* a_empty = bars + C::NSTAGE;       // [NSTAGE]
  uint64_t* x_full = bars + 2 * C::NSTAGE;    // [XBUFS]
  uint64_t* x_empty = x_full + C::XBUFS;      // [XBUFS]
  uint64_t* f_full = x_empty + C::XBUFS;      // [1]
  uint64_t* f_empty = f_full + 1;             // [1]
  double* Xs = reinterpret_cast<double*>(smem_raw);                 // [2][XSLABS][NT][16], swizzled
  double* Fs = Xs + C::XBUFS * C::XBUF;                             // [NT][LDF]
  double* As = Fs + (C::K1 ? C::NT * C::LDF : 0);                   // [NSTAGE][KC][LDA]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  const int nrhs = p.nrhs;

  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], C::NWARPS); }
    for (int s = 0; s < C::XBUFS; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], C::NWARPS); }
    mbar_init(f_full, 1);
    mbar_init(f_empty, C::NWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (my <= 0) return;

  auto item_cols = [&](int item) { return min(C::NT, nrhs - ((first + item) % ntiles) * C::NT); };

#if HSSB_ENABLE_DEBUG_MODES  // measurement builds only (make DEBUG_MODES=1): tools/leaf_bounds.py
  const bool dbg_nowait = p.debug & 1, dbg_nomma = p.debug & 2, dbg_nostore = p.debug & 4;
#else
  constexpr bool dbg_nowait = false, dbg_nomma = false, dbg_nostore = false;
#endif
  if (warp == C::NWARPS) {
    if (dbg_nowait) return;
    // ====================== producer warp ======================
    // X keeps the user's layout in global memory: XSLABS tiled TMA loads per block, slice `piece`
    // of `npieces` per call.  Columns beyond nrhs are zero filled by the TMA unit and still
    // counted, so the expected byte count is always the full block.
    auto load_x = [&](int item, int piece, int npieces) {
      const GTask& tk = tasks[(first + item) / ntiles];
      const int tile = (first + item) % ntiles, buf = item % C::XBUFS;
      if (piece == 0) {
        mbar_wait(&x_empty[buf], ((item / C::XBUFS) & 1) ^ 1);
        if (lane == 0) mbar_expect_tx(&x_full[buf], (uint32_t)(C::XBUF * 8));
        __syncwarp();
      }
      const int per = (C::XSLABS + npieces - 1) / npieces;
      const int s0 = piece * per, s1 = min(C::XSLABS, s0 + per);
      for (int sl = s0 + lane; sl < s1; sl += 32)
        tma_load_2d(Xs + buf * C::XBUF + sl * (C::NT * 16), &xmap, (int)tk.b0 + sl * 16, tile * C::NT, &x_full[buf]);
    };
    auto load_f = [&](int item) {
      if (C::K1 == 0) return;
      const GTask& tk = tasks[(first + item) / ntiles];
      const int tile = (first + item) % ntiles, nc = item_cols(item);
      mbar_wait(f_empty, (item & 1) ^ 1);
      if (lane == 0) {  // the F workspace already carries the padded leading dimension: one copy per tile
        const uint32_t bytes = (uint32_t)(nc * C::LDF * 8);
        mbar_expect_tx(f_full, bytes);
        bulk_g2s(Fs, p.F + tk.b1 * (int64_t)nrhs + (int64_t)tile * C::NT * C::LDF, bytes, f_full);
      }
    };
    load_x(0, 0, 1);
    int g = 0;
    for (int item = 0; item < my; ++item) {
      const GTask& tk = tasks[(first + item) / ntiles];
      for (int c = 0; c < C::NCH; ++c, ++g) {
        const int st = g % C::NSTAGE;
        mbar_wait(&a_empty[st], ((g / C::NSTAGE) & 1) ^ 1);
        if (lane == 0) {  // KC columns of the padded pool image = one contiguous copy
          constexpr uint32_t bytes = C::STAGE_BYTES;
          const double* src = p.pool + (c < C::NCH0 ? tk.a0 + (int64_t)c * C::KC * C::LDA : tk.a1 + (int64_t)(c - C::NCH0) * C::KC * C::LDA);
          mbar_expect_tx(&a_full[st], bytes);
          bulk_g2s(As + st * C::KC * C::LDA, src, bytes, &a_full[st]);
        }
        __syncwarp();
        if (c == C::CX) load_f(item);
        if (c >= C::CX && c < C::CX + C::XPIECES && item + 1 < my) load_x(item + 1, c - C::CX, C::XPIECES);
      }
    }
    return;
  }

  // ====================== consumer warps ======================
  // Each warp owns a (TM*8) x (TN*8) tile of OUT.  Per chunk it issues KSTEPS x TM x TN DMMAs.  The
  // wait for the NEXT chunk's data is software-pipelined: the try_wait is issued two k-steps
  // before the end of the current chunk and only checked after its DMMAs have been issued, so the
  // ~100-cycle mbarrier round trip never sits between two DMMAs of this warp.
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % C::WR, wc = warp / C::WR;
  // Which right-hand side the n-index of a DMMA tile stands for is free: n -> column perm8(n) of the
  // 8-column group.  With the identity, the half-warp gq = 0..3 reads rows 0..3 of the swizzled X box,
  // whose 16-byte chunks (c ^ row) collide pairwise (rows 0/1 and 2/3 swap the same two chunks): a
  // 2-way bank conflict on every B fragment load.  Rows {0,3,5,6} / {1,2,4,7} differ in bits 1-2 (X box:
  // 8 distinct chunks per half-warp) and in row mod 4 (padded F block, 4*row + t mod 16): conflict free
  // on both operands.
  const int pg = perm8(gq);
  double acc[C::TM][C::TN][2];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  struct NextWait {
    uint64_t* b0; uint32_t p0;
    uint64_t* b1; uint32_t p1;
  };
  auto try_next = [&](const NextWait& w) -> bool {
    bool ok = true;
    if (w.b0) ok = mbar_try_wait(w.b0, w.p0);
    if (w.b1) ok = mbar_try_wait(w.b1, w.p1) && ok;
    return ok;
  };
  auto spin_next = [&](const NextWait& w) {
    if (w.b0) mbar_wait(w.b0, w.p0);
    if (w.b1) mbar_wait(w.b1, w.p1);
  };
  // One chunk.  A fragments: "N" operand from the ring stage.  B fragments: from the swizzled X
  // block (XPART: element (k, j) at slab k/16, row j, 16-byte chunk ((k%16)/2) ^ (j%8)) or from
  // the padded F block.  `kstep0` = index of the chunk's first k-step within the item.
  auto chunk = [&](const double* A, const double* B, auto xpart_tag, int kstep0, const NextWait& nw) -> bool {
    constexpr bool XPART = decltype(xpart_tag)::value;
    bool ok = true;
    int sw[4];
    if (XPART) {
      const int qb = kstep0 & 3;  // non-zero only when a chunk is shorter than a 16-row slab (KSTEPS == 2)
#pragma unroll
      for (int q = 0; q < 4; ++q) sw[q] = ((((qb + q) & 3) * 2 + (t >> 1)) ^ pg) * 2 + (t & 1);
      B += (kstep0 >> 2) * (C::NT * 16);
    }
#pragma unroll
    for (int kk = 0; kk < C::KSTEPS; ++kk) {
      double a[C::TM], b[C::TN];
#pragma unroll
      for (int i = 0; i < C::TM; ++i) a[i] = A[kk * 4 * C::LDA + i * 8];
#pragma unroll
      for (int j = 0; j < C::TN; ++j)
        b[j] = XPART ? B[(kk >> 2) * (C::NT * 16) + j * 8 * 16 + sw[kk & 3]] : B[j * 8 * C::LDF + kk * 4];
      if (kk == (C::KSTEPS >= 2 ? C::KSTEPS - 2 : 0)) ok = try_next(nw);
      if (!dbg_nomma) {
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
          for (int j = 0; j < C::TN; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    return ok;
  };

  int st = 0;
  uint32_t ph = 0;  // ring stage of the current chunk and the parity of its current use
  if (!dbg_nowait) {
    mbar_wait(&x_full[0], 0);
    mbar_wait(&a_full[0], 0);
  }
  const double* Abase = As + wr * (C::TM * 8) + gq + t * C::LDA;
  for (int item = 0; item < my; ++item) {
    const int buf = item % C::XBUFS, nbuf = (item + 1) % C::XBUFS;
    const uint32_t nxpar = ((item + 1) / C::XBUFS) & 1;
    const bool more = item + 1 < my;
    // issue the loads the epilogue needs now; their latency hides behind the whole item
    const int task_i = (first + item) / ntiles, tile = (first + item) - task_i * ntiles;
    const int64_t out_row = tasks[task_i].c;
    const double* Bx = Xs + buf * C::XBUF + (wc * (C::TN * 8) + pg) * 16;
#pragma unroll 1
    for (int c = 0; c < C::NCH0; ++c) {
      const int nst = (st + 1 == C::NSTAGE) ? 0 : st + 1;
      const uint32_t nph = (st + 1 == C::NSTAGE) ? ph ^ 1 : ph;
      NextWait nw{nullptr, 0, nullptr, 0};
      const bool last_x = c == C::NCH0 - 1;
      if (!dbg_nowait && (!last_x || C::K1 || more)) {
        nw.b0 = &a_full[nst]; nw.p0 = nph;
        if (last_x) {
          if (C::K1) { nw.b1 = f_full; nw.p1 = item & 1; }
          else { nw.b1 = &x_full[nbuf]; nw.p1 = nxpar; }
        }
      }
      const bool ok = chunk(Abase + st * C::KC * C::LDA, Bx, std::true_type{}, c * C::KSTEPS, nw);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_empty[st]);
        if (last_x) mbar_arrive(&x_empty[buf]);
      }
      if (!ok) spin_next(nw);
      st = nst; ph = nph;
    }
    if (C::K1) {
      const double* Bf = Fs + (wc * (C::TN * 8) + pg) * C::LDF + t;
#pragma unroll 1
      for (int c = 0; c < C::NCH1; ++c) {
        const int nst = (st + 1 == C::NSTAGE) ? 0 : st + 1;
        const uint32_t nph = (st + 1 == C::NSTAGE) ? ph ^ 1 : ph;
        NextWait nw{nullptr, 0, nullptr, 0};
        const bool last_f = c == C::NCH1 - 1;
        if (!dbg_nowait && (!last_f || more)) {
          nw.b0 = &a_full[nst]; nw.p0 = nph;
          if (last_f) { nw.b1 = &x_full[nbuf]; nw.p1 = nxpar; }
        }
        const bool ok = chunk(Abase + st * C::KC * C::LDA, Bf + c * C::KC, std::false_type{}, 0, nw);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_empty[st]);
          if (last_f) mbar_arrive(f_empty);
        }
        if (!ok) spin_next(nw);
        st = nst; ph = nph;
      }
    }
    // ---- epilogue of the item
    const int ncols = min(C::NT, nrhs - tile * C::NT);
    double* O;
    int64_t ldo;
    if (DOWN) { O = p.Y + out_row + (int64_t)tile * C::NT * p.ldy; ldo = p.ldy; }
    else { O = p.Z + out_row * (int64_t)nrhs + (int64_t)tile * C::NT * (R + 4); ldo = R + 4; }
    O += (int64_t)(wc * (C::TN * 8)) * ldo + wr * (C::TM * 8) + gq;
    const int colb = wc * (C::TN * 8);
    const int pc[2] = {perm8(2 * t), perm8(2 * t + 1)};  // columns of this lane's two accumulator elements
#pragma unroll
    for (int j = 0; j < C::TN; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = colb + j * 8 + pc[e] < ncols && !dbg_nostore;
        double* dcol = O + (int64_t)(j * 8 + pc[e]) * ldo;
#pragma unroll
        for (int i = 0; i < C::TM; ++i) {
          if (live) {
            double v = acc[i][j][e];
            if (DOWN) {
              v *= p.alpha;
              if (p.beta != 0.0) v += p.beta * dcol[i * 8];  // beta == 0 never reads Y (matmul.jl:13)
            }
            dcol[i * 8] = v;
          }
          acc[i][j][e] = 0.0;
        }
      }
    }
  }
}

// ============================================================ tree levels ===
// Merges and translates share one persistent, warp-specialised kernel.  An item is one
// (task, column tile):
//     OUT[R x NT] = A0[R x R] * B0[R x NT] (+ A1[R x R] * B1[R x NT])
//   merge      Z  = W1' Z1 + W2' Z2     (matmul.jl:39; the pool holds W')
//   translate  F1 = B12 Z2 (+ R1 F)     (matmul.jl:52-56)
// A0/A1 are padded pool blocks (one bulk copy each) streamed through a ring; B0/B1 are padded
// workspace tiles (one bulk copy each), double buffered per item.  Consecutive tiles of one task
// re-read A0/A1 from L2.  Same producer / consumer protocol as the leaf kernel.
template <int R, int NT_>
struct NodeCfg {
  static constexpr int NT = NT_;                        // 64, or 32 when nrhs <= 32 (no half-empty tiles)
  static constexpr int LD = R + 4;
  static constexpr int WR = (R >= 32 || NT < 64) ? 2 : 1, WC = 8 / WR;
  static constexpr int TM = R / WR / 8, TN = NT / WC / 8;
  static constexpr int KSTEPS = R / 4;
  static constexpr int TILE = NT * LD;                 // doubles per B tile
  static constexpr int STAGE = R * LD;                 // doubles per A block
  static constexpr int BAR_BYTES = 128;
  static constexpr int FIXED_BYTES = BAR_BYTES + 8 * (2 * 2 * TILE);
  static constexpr int FIT = (R >= 64 ? 232448 : 113000) - FIXED_BYTES;   // R < 64: leave room for 2 CTAs per SM
  static constexpr int NSTAGE = FIT / (8 * STAGE) > 4 ? 4 : FIT / (8 * STAGE);
  static constexpr size_t SMEM = FIXED_BYTES + (size_t)NSTAGE * 8 * STAGE;
  static constexpr int CTAS_PER_SM = R >= 64 ? 1 : 2;
  static_assert(NSTAGE >= 2 && TM >= 1 && TN >= 1 && 2 * NSTAGE + 4 <= BAR_BYTES / 8, "node kernel configuration");
};

template <int R, int NT_>
__global__ void __launch_bounds__(288, NodeCfg<R, NT_>::CTAS_PER_SM)
stream_node_kernel(const GTask* __restrict__ tasks, int ntasks, int ntiles, CallParams p) {
  pdl_launch_dependents();
  using C = NodeCfg<R, NT_>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* Bt = reinterpret_cast<double*>(smem_raw);            // [2 buffers][2 operands][NT][LD]
  double* As = Bt + 4 * C::TILE;                               // [NSTAGE][R][LD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + C::NSTAGE;
  uint64_t* b_full = bars + 2 * C::NSTAGE;   // [2]
  uint64_t* b_empty = b_full + 2;            // [2]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  const int nrhs = p.nrhs;
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 8); }
    for (int s = 0; s < 2; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (my <= 0) return;
  const bool two = tasks[0].K1 > 0;  // uniform over the launch (false only for the root translate)
  const int nch = two ? 2 : 1;
  auto ws = [&](int src) -> const double* { return src == SRC_F ? p.F : p.Z; };

  if (warp == 8) {
    // ====================== producer warp ======================
    if (lane != 0) return;
    auto load_b = [&](int item) {
      const GTask& tk = tasks[(first + item) / ntiles];
      const int tile = (first + item) % ntiles, buf = item & 1;
      const uint32_t bytes = (uint32_t)(min(C::NT, nrhs - tile * C::NT) * C::LD * 8);
      const int64_t toff = (int64_t)tile * C::TILE;
      mbar_wait(&b_empty[buf], ((item >> 1) & 1) ^ 1);
      mbar_expect_tx(&b_full[buf], two ? 2 * bytes : bytes);
      bulk_g2s(Bt + (buf * 2) * C::TILE, ws(tk.sb0) + tk.b0 * (int64_t)nrhs + toff, bytes, &b_full[buf]);
      if (two) bulk_g2s(Bt + (buf * 2 + 1) * C::TILE, ws(tk.sb1) + tk.b1 * (int64_t)nrhs + toff, bytes, &b_full[buf]);
    };
    int g = 0;
    auto load_a = [&](int item) {
      const GTask& tk = tasks[(first + item) / ntiles];
      for (int c = 0; c < nch; ++c, ++g) {
        const int st = g % C::NSTAGE;
        mbar_wait(&a_empty[st], ((g / C::NSTAGE) & 1) ^ 1);
        mbar_expect_tx(&a_full[st], C::STAGE * 8);
        bulk_g2s(As + st * C::STAGE, p.pool + (c ? tk.a1 : tk.a0), C::STAGE * 8, &a_full[st]);
      }
    };
    load_a(0);   // generator blocks: nobody produces them, they travel before the wait (HSSB_OPT_PDL)
    pdl_wait();  // the previous level is complete and visible from here on
    load_b(0);
    for (int item = 0; item < my; ++item) {
      if (item > 0) load_a(item);
      // in order of need: this item's A blocks first, then the next item's B tiles (their buffer
      // frees up when the consumers leave item-1)
      if (item + 1 < my) load_b(item + 1);
    }
    return;
  }

  // ====================== consumer warps ======================
  pdl_wait();  // they only read what the producer fetched after ITS wait; this one orders their stores as well
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % C::WR, wc = warp / C::WR;
  double acc[C::TM][C::TN][2];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* Abase = As + wr * (C::TM * 8) + gq + t * C::LD;
  const double* Bbase = Bt + (wc * (C::TN * 8) + gq) * C::LD + t;
  int st = 0;
  uint32_t ph = 0;
  for (int item = 0; item < my; ++item) {
    const int buf = item & 1;
    const int task_i = (first + item) / ntiles, tile = (first + item) - task_i * ntiles;
    const int64_t out_row = tasks[task_i].c;
    const int out_src = tasks[task_i].sc;
    mbar_wait(&b_full[buf], (item >> 1) & 1);
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&a_full[st], ph);
      const double* A = Abase + st * C::STAGE;
      const double* B = Bbase + (buf * 2 + c) * C::TILE;
#pragma unroll
      for (int kk = 0; kk < C::KSTEPS; ++kk) {
        double a[C::TM], b[C::TN];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) a[i] = A[kk * 4 * C::LD + i * 8];
#pragma unroll
        for (int j = 0; j < C::TN; ++j) b[j] = B[j * 8 * C::LD + kk * 4];
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
          for (int j = 0; j < C::TN; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_empty[st]);
        if (c == nch - 1) mbar_arrive(&b_empty[buf]);
      }
      if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
    }
    const int ncols = min(C::NT, nrhs - tile * C::NT);
    double* O = (out_src == SRC_F ? p.F : p.Z) + out_row * (int64_t)nrhs + (int64_t)tile * C::TILE;
    O += (int64_t)(wc * (C::TN * 8) + 2 * t) * C::LD + wr * (C::TM * 8) + gq;
    const int colb = wc * (C::TN * 8) + 2 * t;
#pragma unroll
    for (int j = 0; j < C::TN; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = colb + j * 8 + e < ncols;
#pragma unroll
        for (int i = 0; i < C::TM; ++i) {
          if (live) O[(int64_t)(j * 8 + e) * C::LD + i * 8] = acc[i][j][e];
          acc[i][j][e] = 0.0;
        }
      }
  }
}

// One-shot flavour of the node kernel: one CTA of 128 threads per (task, 64-column tile), four
// bulk copies on one mbarrier, no pipeline inside the CTA; latency is hidden by running four such
// CTAs per SM.  Slightly faster than the persistent kernel for small ranks and a single column
// tile (config 3), where an item is only ~0.5 us of DMMA work.
template <int R>
struct OneShotCfg {
  static constexpr int NT = 64, LD = R + 4, TN = 2;
  static constexpr size_t SMEM = 128 + sizeof(double) * (2 * R * LD + 2 * NT * LD);
};

template <int R>
__global__ void __launch_bounds__(128)
oneshot_node_kernel(const GTask* __restrict__ tasks, CallParams p) {
  using C = OneShotCfg<R>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* As = reinterpret_cast<double*>(smem_raw + 128);  // [2][R][LD]   A(i,k) at [k*LD + i]
  double* Bs = As + 2 * R * C::LD;                         // [2][NT][LD]  B(k,j) at [j*LD + k]
  const GTask tk = tasks[blockIdx.x];
  const int tile = blockIdx.y, nrhs = p.nrhs;
  const int ncols = min(C::NT, nrhs - tile * C::NT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const bool two = tk.K1 > 0;
  const int64_t toff = (int64_t)tile * C::NT * C::LD;
  pdl_launch_dependents();
  constexpr uint32_t ab = R * C::LD * 8;
  const uint32_t bb = (uint32_t)(ncols * C::LD * 8);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    mbar_expect_tx(bar, two ? 2 * (ab + bb) : ab + bb);
    bulk_g2s(As, p.pool + tk.a0, ab, bar);              // generators: nobody produces them, they travel before the wait
    if (two) bulk_g2s(As + R * C::LD, p.pool + tk.a1, ab, bar);
  }
  pdl_wait();                                           // the level before this one is complete and visible from here on
  if (tid == 0) {
    bulk_g2s(Bs, (tk.sb0 == SRC_F ? p.F : p.Z) + tk.b0 * (int64_t)nrhs + toff, bb, bar);
    if (two) bulk_g2s(Bs + C::NT * C::LD, (tk.sb1 == SRC_F ? p.F : p.Z) + tk.b1 * (int64_t)nrhs + toff, bb, bar);
  }
  __syncthreads();  // the barrier is initialised before anyone polls it
  mbar_wait(bar, 0);

  double acc[R / 8][C::TN][2];
#pragma unroll
  for (int i = 0; i < R / 8; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int col0 = warp * 16;
  const int nsrc = two ? 2 : 1;
  for (int s = 0; s < nsrc; ++s) {
    const double* A = As + s * R * C::LD + g + t * C::LD;
    const double* B = Bs + s * C::NT * C::LD + (col0 + g) * C::LD + t;
#pragma unroll
    for (int kk = 0; kk < R / 4; ++kk) {
      double b[C::TN];
#pragma unroll
      for (int j = 0; j < C::TN; ++j) b[j] = B[j * 8 * C::LD + kk * 4];
#pragma unroll
      for (int i = 0; i < R / 8; ++i) {
        const double a = A[kk * 4 * C::LD + i * 8];
#pragma unroll
        for (int j = 0; j < C::TN; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a, b[j]);
      }
    }
  }
  double* O = (tk.sc == SRC_F ? p.F : p.Z) + tk.c * (int64_t)nrhs + toff;
#pragma unroll
  for (int j = 0; j < C::TN; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = col0 + j * 8 + 2 * t + e;
      if (col < ncols) {
#pragma unroll
        for (int i = 0; i < R / 8; ++i) O[(int64_t)col * C::LD + i * 8 + g] = acc[i][j][e];
      }
    }
}

// ================================================================ host side ===
// Kernel launch with or without the programmatic-dependent-launch attribute.  HSSB_OPT_PDL bits: 1 = where it was measured
// to pay (tools/pdl_compare.py, profiles/pdl_r02.txt): the one-shot node kernels (config 3: -1.8 %) and the persistent node
// kernel at rank 64, one CTA per SM (config 5 shape -3 %, config 4 shape -0.8 %); 2 = every persistent node kernel (rank 32,
// two CTAs per SM: +1 %); 4 = the leaf kernels (+2 .. 6 %).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  memset(at, 0, sizeof(at));
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, (KArgs)args...);
}

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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of the caller's X (rows x nrhs, column-major, leading dimension ldx): box {16, NT}, 128B swizzle.
static int make_x_map(CUtensorMap* map, const double* X, int64_t rows, int64_t nrhs, int64_t ldx, int nt) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    HSSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) HSSB_FAIL(HSSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)nrhs};
  const cuuint64_t gstr[1] = {(cuuint64_t)ldx * 8};
  const cuuint32_t box[2] = {16, (cuuint32_t)nt};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) HSSB_FAIL(HSSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
  return HSSB_OK;
}

}  // namespace hssb
#include "hssb_leaf2.cuh"
#include "hssb_leafx.cuh"
namespace hssb {

// ---- the "X once" variant (HSSB_OPT_LEAF_FUSION): shapes it is instantiated for
static bool leafx_supported(int64_t m, int64_t r) { return m == 128 && (r == 32 || r == 64); }

template <int M, int R>
static int launch_leafx(hssb_matrix* H, const Phase& ph, const Phase& up, const Phase& down, const CallParams& cp, cudaStream_t st) {
  constexpr int NT = 8192 / M, KC = 16;
  FastState* fs = (FastState*)H->fast_state;
  const int ntiles = (cp.nrhs + NT - 1) / NT;
  const int grid = std::min((int)ph.ntasks * ntiles, fs->num_sms);
  if (ph.kind == PH_LEAF_UP) {
    using C = LeafXUpCfg<M, R, NT, KC>;
    CUtensorMap xmap;
    if (int rc = make_x_map(&xmap, cp.X, H->local_n, cp.nrhs, cp.ldx, NT)) return rc;
    if (int rc = fs->configure((const void*)leafx_up_kernel<M, R, NT, KC>, C::SMEM)) return rc;
    leafx_up_kernel<M, R, NT, KC><<<grid, 288, C::SMEM, st>>>(H->tasks_dev + up.task0, H->tasks_dev + down.task0, (int)ph.ntasks, ntiles, cp, xmap);
  } else {
    using C = LeafXDownCfg<M, R, NT>;
    if (int rc = fs->configure((const void*)leafx_down_kernel<M, R, NT>, C::SMEM)) return rc;
    leafx_down_kernel<M, R, NT><<<grid, 288, C::SMEM, st>>>(H->tasks_dev + down.task0, (int)ph.ntasks, ntiles, cp);
  }
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

// Both leaf phases of the product plan when the fused variant can run them, else nullptr.
static bool leafx_phases(const hssb_matrix* H, const CallParams& cp, const Phase*& up, const Phase*& down) {
  up = down = nullptr;
  if (!H->leaf_fusion || cp.trans != 0 || !H->uniform || !H->padded || !leafx_supported(H->uni_m, H->uni_r)) return false;
  for (const Phase& q : H->phases) {
    if (q.kind == PH_LEAF_UP && q.fast == FAST_LEAF_UP && !q.fast_m) up = &q;
    if (q.kind == PH_LEAF_DOWN && q.fast == FAST_LEAF_DOWN && !q.fast_m) down = &q;
  }
  return up && down && up->ntasks == down->ntasks;
}

template <int M, int R, bool DOWN, int NT, int KC>
static int launch_leaf2(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  using C = Leaf2Cfg<M, R, DOWN, NT, KC>;
  FastState* fs = (FastState*)H->fast_state;
  CUtensorMap xmap;
  const int ntiles = (cp.nrhs + C::NT - 1) / C::NT;
  const int grid = std::min((int)ph.ntasks * ntiles, fs->num_sms);
  if (int rc = make_x_map(&xmap, cp.X, H->local_n, cp.nrhs, cp.ldx, C::NT)) return rc;
  if (int rc = fs->configure((const void*)leaf2_kernel<M, R, DOWN, NT, KC>, C::SMEM)) return rc;
  HSSB_CUDA(launch_k((H->pdl & 4) != 0, leaf2_kernel<M, R, DOWN, NT, KC>, dim3(grid), dim3(C::NWARPS * 32 + 32), C::SMEM, st, (const GTask*)(H->tasks_dev + ph.task0), (int)ph.ntasks, ntiles, cp, xmap));
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

template <int M, int R, bool DOWN, int NT>
static int launch_leaf_nt(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  // second generation (self-contained ring stages, hssb_leaf2.cuh); HSSB_OPT_LEAF_KERNEL = 1 keeps the first
  if (H->leaf_kernel != 1) {
    if constexpr (DOWN) {
      // longer chunks = fewer chunk boundaries (each costs ~100 cycles without DMMAs): 32 columns of [D U] per
      // stage where the ring still holds 4 stages (measured: c3 0.687 -> 0.666 ms, c4 1.584 -> 1.534 ms)
      if constexpr (M == 128 && R % 32 == 0) {
        if (H->leaf_kernel != 3) return launch_leaf2<M, R, true, NT, 32>(H, ph, cp, st);
      }
      return launch_leaf2<M, R, true, NT, 16>(H, ph, cp, st);
    } else {
      if (H->leaf_kernel != 3) return launch_leaf2<M, R, false, NT, 64>(H, ph, cp, st);
      return launch_leaf2<M, R, false, NT, 32>(H, ph, cp, st);
    }
  }
  using C = StreamCfg<M, R, DOWN, NT>;
  FastState* fs = (FastState*)H->fast_state;
  CUtensorMap xmap;
  const int ntiles = (cp.nrhs + C::NT - 1) / C::NT;
  const int grid = std::min((int)ph.ntasks * ntiles, fs->num_sms);
  if (int rc = make_x_map(&xmap, cp.X, H->local_n, cp.nrhs, cp.ldx, C::NT)) return rc;
  if (int rc = fs->configure((const void*)stream_leaf_kernel<M, R, DOWN, NT>, C::SMEM)) return rc;
  stream_leaf_kernel<M, R, DOWN, NT><<<grid, C::NWARPS * 32 + 32, C::SMEM, st>>>(H->tasks_dev + ph.task0, (int)ph.ntasks, ntiles, cp, xmap);
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

template <int M, int R>
static int launch_leaf(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st, bool down) {
  constexpr int NTW = 8192 / M;  // widest tile: 64 right-hand sides for 128-row leaves, 32 for 256-row leaves
  // Few right-hand sides (e.g. nrhs = 20 in randcompress_adaptive, compression.jl:341): with wide
  // tiles most of the DMMA work would be spent on empty columns and the product, HBM-bound at small
  // nrhs, would become compute-bound.  Pick the narrowest tile that covers the last (or only) tile.
  const int rem = cp.nrhs % 64 == 0 ? 64 : cp.nrhs % 64;
  if constexpr (R >= 32) {
    if (cp.nrhs <= 16) return down ? launch_leaf_nt<M, R, true, 16>(H, ph, cp, st) : launch_leaf_nt<M, R, false, 16>(H, ph, cp, st);
  }
  const bool narrow = NTW > 32 && rem <= 32 && cp.nrhs < 128;
  if (narrow) return down ? launch_leaf_nt<M, R, true, 32>(H, ph, cp, st) : launch_leaf_nt<M, R, false, 32>(H, ph, cp, st);
  return down ? launch_leaf_nt<M, R, true, NTW>(H, ph, cp, st) : launch_leaf_nt<M, R, false, NTW>(H, ph, cp, st);
}

template <int R, int NT>
static int launch_node_nt(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  using C = NodeCfg<R, NT>;
  FastState* fs = (FastState*)H->fast_state;
  const int ntiles = (cp.nrhs + C::NT - 1) / C::NT;
  const int grid = (int)std::min<int64_t>(ph.ntasks * ntiles, (int64_t)fs->num_sms * C::CTAS_PER_SM);
  if (int rc = fs->configure((const void*)stream_node_kernel<R, NT>, C::SMEM)) return rc;
  HSSB_CUDA(launch_k((H->pdl & 2) != 0 || ((H->pdl & 1) != 0 && R >= 64), stream_node_kernel<R, NT>, dim3(grid), dim3(288), C::SMEM, st, (const GTask*)(H->tasks_dev + ph.task0), (int)ph.ntasks, ntiles, cp));
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

template <int R>
static int launch_node(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  if constexpr (R <= 32) {
    if (cp.nrhs > 32 && cp.nrhs <= 64) {  // one full-width tile of little work per item: one-shot CTAs, 4 per SM
      using C = OneShotCfg<R>;
      FastState* fs = (FastState*)H->fast_state;
      if (int rc = fs->configure((const void*)oneshot_node_kernel<R>, C::SMEM)) return rc;
      HSSB_CUDA(launch_k((H->pdl & 1) != 0, oneshot_node_kernel<R>, dim3((unsigned)ph.ntasks, 1), dim3(128), C::SMEM, st, (const GTask*)(H->tasks_dev + ph.task0), cp));
      H->launches++;
      HSSB_CUDA(cudaGetLastError());
      return HSSB_OK;
    }
  }
  // the last (or only) column tile is at most half full with 64-wide tiles: use 32-wide ones
  const int rem = cp.nrhs % 64;
  if (rem > 0 && rem <= 32 && cp.nrhs < 128) return launch_node_nt<R, 32>(H, ph, cp, st);
  return launch_node_nt<R, 64>(H, ph, cp, st);
}

static bool fast_shape_supported(int64_t m, int64_t r) {
  return (m == 128 && (r == 16 || r == 32 || r == 64)) || (m == 256 && (r == 16 || r == 32 || r == 64));
}

// Tags the phases of a uniform tree that a fixed-shape kernel can run.
static void plan_fast_phases(hssb_matrix* H) {
  if (!H->uniform || !H->padded || !fast_shape_supported(H->uni_m, H->uni_r)) return;
  const int m = (int)H->uni_m, r = (int)H->uni_r;
  for (Phase& ph : H->phases) {
    if (ph.kind == PH_EXCHANGE || ph.ntasks == 0) continue;
    bool ok = true;
    const GTask* tk = H->tasks_host.data() + ph.task0;
    for (int64_t i = 0; i < ph.ntasks && ok; ++i) {
      const GTask& g = tk[i];
      switch (ph.kind) {
        // the kernels rely on the padded (+4) leading dimensions of the pool and the workspaces
        case PH_LEAF_UP: ok = g.M == r && g.K0 == m && g.lda0 == r + 4 && g.ta0 == 0 && g.ldc == r + 4 && g.a0 >= 0; break;
        case PH_MERGE: ok = g.M == r && g.K0 == r && g.K1 == r && g.lda0 == r + 4 && g.lda1 == r + 4 && g.ldb0 == r + 4 && g.ldb1 == r + 4 && g.ldc == r + 4 && g.a0 >= 0 && g.a1 >= 0 && !g.ta0 && !g.ta1; break;
        case PH_TRANSLATE:
          ok = g.M == r && g.K0 == r && (g.K1 == r || g.K1 == 0) && g.lda0 == r + 4 && g.ldb0 == r + 4 && g.ldc == r + 4 && g.a0 >= 0 &&
               (g.K1 == 0 || (g.lda1 == r + 4 && g.ldb1 == r + 4 && g.a1 >= 0)) && g.K1 == tk[0].K1 && !g.ta0 && !g.ta1;
          break;
        case PH_LEAF_DOWN: ok = g.M == m && g.K0 == m && g.K1 == r && g.lda0 == m + 4 && g.lda1 == m + 4 && g.ldb1 == r + 4 && g.a0 >= 0 && g.a1 >= 0; break;
        default: ok = false;
      }
    }
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
  // the product's phases run at the tree's (leaf size, rank); phases of the ULV solve carry their own shape
  const int m = ph.fast_m ? ph.fast_m : (int)H->uni_m, r = ph.fast_r ? ph.fast_r : (int)H->uni_r;
  if (ph.fast == FAST_MERGE || ph.fast == FAST_TRANSLATE) {
    switch (r) {
      case 16: return launch_node<16>(H, ph, cp, st);
      case 32: return launch_node<32>(H, ph, cp, st);
      case 64: return launch_node<64>(H, ph, cp, st);
    }
  } else {
    const bool down = ph.fast == FAST_LEAF_DOWN;
    const Phase *xu, *xd;
    if (!ph.fast_m && leafx_phases(H, cp, xu, xd)) {  // opt-in: one pass over X for D X and V' X (north star (2))
      if (m == 128 && r == 32) return launch_leafx<128, 32>(H, ph, *xu, *xd, cp, st);
      if (m == 128 && r == 64) return launch_leafx<128, 64>(H, ph, *xu, *xd, cp, st);
    }
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
