// hssb_tree.cuh — ONE persistent kernel for every tree level between the two leaf kernels.
//
// The recursion of src/matmul.jl runs the merges Z = W1' Z1 + W2' Z2 (matmul.jl:39) bottom-up and the
// translates F1 = B12 Z2 + R1 F (matmul.jl:52-56) top-down: 2*depth - 1 dependent levels, most of them
// far too small to fill a B200 (config 3: 17 of 25 levels have <= 128 tasks of ~0.26 MFLOP).  One launch
// per level costs ~8 us of launch + ramp + drain each; here the whole sweep is one cooperative launch of
// one CTA per SM that walks a step table:
//
//   step = one level (merge by height / translate by depth), the NVLink exchange of the subtree-root
//          Z blocks (sharded handles), or its acknowledgement.
//   between steps: a grid barrier built from per-CTA arrival words in global memory (every CTA stores
//          its own word, every CTA polls all of them: no atomics, ~2 L2 round trips).
//   inside a level: the warp-specialised pipeline of stream_node_kernel (TMA producer warp, 8 DMMA
//          consumer warps, generator ring + operand buffers).  Generators do not depend on the previous
//          level, so the producer prefetches them BEFORE it waits on the barrier; only the Z / F operand
//          tiles are fetched after it.
//   small levels: an item is (task, column slice); the slice narrows from 64 to 32 to 16 columns until
//          the level has enough items for all SMs, so the dependent chain at the top of the tree costs
//          ~1 us per level (barrier + one L2 round trip + a few hundred DMMA cycles).
//
// Data written with st.global by one CTA is read by cp.async.bulk (async proxy) in another: writers fence
// (gpu scope + proxy fence) before they arrive, readers issue a proxy fence after the acquire.
#pragma once

#include "hssb_fast.cuh"

namespace hssb {

enum TreeStepKind : int { TS_LEVEL = 0, TS_XCHG = 1, TS_ACK = 2 };

struct TreeStep {
  int32_t kind;
  int32_t task0, ntasks;  // range in the task table
  int32_t two;            // tasks carry a second operand pair (false only for the root translate)
};

// Peer-memory exchange inside the tree kernel: XCHG_PARTS CTAs push one part of the slot to each peer.
constexpr int XCHG_PARTS = 4;
constexpr int XCHG_FLAG_WORDS = 256;  // [0,P) data (legacy kernels), [P,2P) ack, [2P] epoch, [2P+1] ticket, [64 + 4 r + j] data parts
constexpr int XCHG_PART0 = 64;

template <int R>
struct TreeCfg {
  static constexpr int LD = R + 4;
  static constexpr int CWMAX = 64;
  static constexpr int TILE = CWMAX * LD;  // doubles per operand buffer (one 64-column tile)
  static constexpr int STAGE = R * LD;     // doubles per generator block
  static constexpr int NBUF = R >= 64 ? 2 : (R >= 32 ? 3 : 4);    // items whose operands are in flight
  static constexpr int NSTAGE = R >= 64 ? 2 : (R >= 32 ? 6 : 8);  // generator ring depth
  static constexpr int BAR_BYTES = 256;
  static constexpr size_t SMEM = BAR_BYTES + sizeof(double) * ((size_t)NBUF * 2 * TILE + (size_t)NSTAGE * STAGE);
  static_assert(SMEM <= 232448 && 2 * NSTAGE + 2 * NBUF + 1 <= BAR_BYTES / 8, "tree kernel shared memory budget");
};

// Column-slice width of a level: narrow the slices until every SM has a few items.
__host__ __device__ __forceinline__ int tree_slice_width(int ntasks, int nrhs, int ncta) {
  int cw = nrhs <= 16 ? 16 : (nrhs <= 32 ? 32 : 64);
  while (cw > 16 && (long long)ntasks * ((nrhs + cw - 1) / cw) < 4ll * ncta) cw >>= 1;
  return cw;
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Whole-warp wait until every CTA of the grid has arrived `target` times.
__device__ __forceinline__ void grid_wait(const unsigned long long* arrive, int ncta, unsigned long long target, int lane) {
  const long long t0 = clock64();
  for (int base = 0; base < ncta; base += 128) {  // 4 independent polls per lane and round
    bool done;
    do {
      unsigned long long v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = base + u * 32 + lane;
        v[u] = j < ncta ? ld_relaxed_gpu(arrive + j) : target;
      }
      done = v[0] >= target && v[1] >= target && v[2] >= target && v[3] >= target;
      done = __all_sync(0xffffffffu, done);
      if (!done && clock64() - t0 > 8000000000ll) trap_report(TRAP_GRID, target, (unsigned long long)base);  // a lost CTA must surface as an error, not a hang
    } while (!done);
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");  // one acquire for all the relaxed polls
  fence_proxy_async();
}

// Levels too small to fill the machine take the DIRECT path: the consumer warps poll the grid barrier
// themselves and read their B fragments straight from L2 into registers (no TMA round trip, no shared
// memory hop), so a level of the dependent chain at the top of the tree costs one release store, one
// poll, one L2 read and a few hundred DMMA cycles.
__host__ __device__ __forceinline__ bool tree_level_direct(int ntasks, int nrhs, int ncta, int cw) {
  return cw == 16 && (long long)ntasks * ((nrhs + cw - 1) / cw) <= 8ll * ncta;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One level, consumer side, TMA-fed, for one slice width CW.  `itb` / `st` / `ph` are the CTA-wide running
// counters of TMA-fed items and of the generator ring (identical in every warp and in the producer).
template <int R, int CW>
__device__ __forceinline__ void tree_consume_level(const GTask* __restrict__ tasks, int first, int last, int nsl, int nch,
                                                   const CallParams& p, const double* As, const double* Bt, uint64_t* a_full,
                                                   uint64_t* a_empty, uint64_t* b_full, uint64_t* b_empty, int& itb, int& st,
                                                   uint32_t& ph, int warp, int lane) {
  using C = TreeCfg<R>;
  constexpr int WC = CW >= 32 ? 4 : 2;
  constexpr int WR = (8 / WC) < (R / 8) ? (8 / WC) : (R / 8);
  constexpr int TM = R / (8 * WR), TN = CW / (8 * WC);
  static_assert(TM >= 1 && TN >= 1 && WR * WC <= 8, "tree kernel warp tiling");
  const bool active = warp < WR * WC;
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % WR, wc = (warp / WR) % WC;
  const int nrhs = p.nrhs;
  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* Abase = As + wr * (TM * 8) + gq + t * C::LD;
  const double* Bbase = Bt + (wc * (TN * 8) + gq) * C::LD + t;
  for (int item = first; item < last; ++item, ++itb) {
    const int buf = itb % C::NBUF;
    const int task_i = item / nsl, sl = item - task_i * nsl;
    const GTask& tk = tasks[task_i];
    const int64_t out_row = tk.c;
    const int out_src = tk.sc;
    mbar_wait(&b_full[buf], (uint32_t)(itb / C::NBUF) & 1);
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&a_full[st], ph);
      if (active) {
        const double* A = Abase + st * C::STAGE;
        const double* B = Bbase + (buf * 2 + c) * C::TILE;
#pragma unroll
        for (int kk = 0; kk < R / 4; ++kk) {
          double a[TM], b[TN];
#pragma unroll
          for (int i = 0; i < TM; ++i) a[i] = A[kk * 4 * C::LD + i * 8];
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j] = B[j * 8 * C::LD + kk * 4];
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) mma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_empty[st]);
        if (c == nch - 1) mbar_arrive(&b_empty[buf]);
      }
      if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
    }
    if (active) {
      const int c0 = sl * CW, ncols = min(CW, nrhs - c0);
      double* O = (out_src == SRC_F ? p.F : p.Z) + out_row * (int64_t)nrhs + (int64_t)c0 * C::LD;
      O += (int64_t)(wc * (TN * 8) + 2 * t) * C::LD + wr * (TM * 8) + gq;
      const int colb = wc * (TN * 8) + 2 * t;
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool live = colb + j * 8 + e < ncols;
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            if (live) O[(int64_t)(j * 8 + e) * C::LD + i * 8] = acc[i][j][e];
            acc[i][j][e] = 0.0;
          }
        }
    }
  }
}

// One level, consumer side, DIRECT path (16-column slices).  Generators still come through the ring (the
// producer prefetched them before the barrier); the warp's B fragments of both operands are read from
// L2 with ld.global.cg, the next item's while the current one is multiplied.
template <int R>
__device__ __forceinline__ void tree_consume_direct(const GTask* __restrict__ tasks, int first, int last, int nsl, int nch,
                                                    const CallParams& p, const double* As, uint64_t* a_full, uint64_t* a_empty,
                                                    int& st, uint32_t& ph, int warp, int lane) {
  using C = TreeCfg<R>;
  constexpr int CW = 16, WC = 2;
  constexpr int WR = (8 / WC) < (R / 8) ? (8 / WC) : (R / 8);
  constexpr int TM = R / (8 * WR), KS = R / 4;
  const bool active = warp < WR * WC;
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % WR, wc = (warp / WR) % WC;
  const int nrhs = p.nrhs;
  const double* Abase = As + wr * (TM * 8) + gq + t * C::LD;
  constexpr bool PREFETCH = R <= 32;  // rank 64: 2 x 32 more registers per lane would spill
  double bcur[2][KS], bnxt[PREFETCH ? 2 : 1][PREFETCH ? KS : 1];
  auto fetch = [&](int item, auto& b) {
    const int task_i = item / nsl, sl = item - task_i * nsl;
    const GTask& tk = tasks[task_i];
    const int col = sl * CW + wc * 8 + gq;  // this lane's column of the slice (B fragment: row 4 kk + t)
    const bool live = active && col < nrhs;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (c < nch && live) {
        const int src = c ? tk.sb1 : tk.sb0;
        const double* B = (src == SRC_F ? p.F : p.Z) + (c ? tk.b1 : tk.b0) * (int64_t)nrhs + (int64_t)col * C::LD + t;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) b[c][kk] = __ldcg(B + kk * 4);
      } else {
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) b[c][kk] = 0.0;
      }
    }
  };
  if (PREFETCH && first < last) fetch(first, bcur);
  for (int item = first; item < last; ++item) {
    const int task_i = item / nsl, sl = item - task_i * nsl;
    const GTask& tk = tasks[task_i];
    if constexpr (PREFETCH) {
      if (item + 1 < last) fetch(item + 1, bnxt);
    } else {
      fetch(item, bcur);
    }
    double acc[TM][2];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int c = 0; c < nch; ++c) {
      mbar_wait(&a_full[st], ph);
      if (active) {
        const double* A = Abase + st * C::STAGE;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          const double b = c ? bcur[1][kk] : bcur[0][kk];
#pragma unroll
          for (int i = 0; i < TM; ++i) mma_m8n8k4(acc[i][0], acc[i][1], A[kk * 4 * C::LD + i * 8], b);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_empty[st]);
      if (++st == C::NSTAGE) { st = 0; ph ^= 1; }
    }
    if (active) {
      const int c0 = sl * CW, ncols = min(CW, nrhs - c0);
      double* O = (tk.sc == SRC_F ? p.F : p.Z) + tk.c * (int64_t)nrhs + (int64_t)c0 * C::LD;
      O += (int64_t)(wc * 8 + 2 * t) * C::LD + wr * (TM * 8) + gq;
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (wc * 8 + 2 * t + e < ncols) {
#pragma unroll
          for (int i = 0; i < TM; ++i) O[(int64_t)e * C::LD + i * 8] = acc[i][e];
        }
    }
    if constexpr (PREFETCH) {
      if (item + 1 < last) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) bcur[c][kk] = bnxt[c][kk];
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(288, 1)
tree_kernel(const GTask* __restrict__ tasks, const TreeStep* __restrict__ steps, int step0, int step1, CallParams p,
            unsigned long long* __restrict__ arrive, const __grid_constant__ XchgParams xq, long long* __restrict__ trace) {
  using C = TreeCfg<R>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double* Bt = reinterpret_cast<double*>(smem_raw);            // [NBUF][2 operands][64][LD]
  double* As = Bt + C::NBUF * 2 * C::TILE;                     // [NSTAGE][R][LD]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + C::NSTAGE;
  uint64_t* b_full = bars + 2 * C::NSTAGE;   // [NBUF]
  uint64_t* b_empty = b_full + C::NBUF;      // [NBUF]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncta = (int)gridDim.x, cta = (int)blockIdx.x;
  const int nrhs = p.nrhs;
  // Every CTA has completed the same number of barrier steps in earlier launches, so its own word is
  // the common base of this launch.  (Read before the CTA-wide barrier below: the word is only written
  // by consumer warp 0 of this CTA, after that barrier.)
  const unsigned long long base = ld_acquire_gpu(arrive + cta);
  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 8); }
    for (int s = 0; s < C::NBUF; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto ws = [&](int src) -> const double* { return src == SRC_F ? p.F : p.Z; };
  auto level_items = [&](const TreeStep& s, int& cw, int& nsl, int& first, int& last, bool& direct) {
    cw = tree_slice_width(s.ntasks, nrhs, ncta);
    direct = tree_level_direct(s.ntasks, nrhs, ncta, cw);
    nsl = (nrhs + cw - 1) / cw;
    const int64_t nitems = (int64_t)s.ntasks * nsl;
    first = (int)(((int64_t)cta * nitems) / ncta);
    last = (int)(((int64_t)(cta + 1) * nitems) / ncta);
  };

  if (warp == 8) {
    // ====================== producer warp ======================
    // Generators (ring) for every level; Z / F operand tiles for the TMA-fed levels, after the barrier.
    int itb = 0, g = 0;
    auto load_a = [&](const GTask& tk, int nch) {
      if (lane == 0) {
        for (int c = 0; c < nch; ++c, ++g) {
          const int stg = g % C::NSTAGE;
          mbar_wait(&a_empty[stg], ((g / C::NSTAGE) & 1) ^ 1);
          mbar_expect_tx(&a_full[stg], C::STAGE * 8);
          bulk_g2s(As + stg * C::STAGE, p.pool + (c ? tk.a1 : tk.a0), C::STAGE * 8, &a_full[stg]);
        }
      }
      g = __shfl_sync(0xffffffffu, g, 0);
    };
    auto load_b = [&](const GTask& tk, int sl, int cw, int nch) {
      if (lane == 0) {
        const int buf = itb % C::NBUF;
        const int c0 = sl * cw;
        const uint32_t bytes = (uint32_t)(min(cw, nrhs - c0) * C::LD * 8);
        const int64_t coff = (int64_t)c0 * C::LD;
        mbar_wait(&b_empty[buf], ((itb / C::NBUF) & 1) ^ 1);
        mbar_expect_tx(&b_full[buf], nch * bytes);
        bulk_g2s(Bt + (buf * 2) * C::TILE, ws(tk.sb0) + tk.b0 * (int64_t)nrhs + coff, bytes, &b_full[buf]);
        if (nch == 2) bulk_g2s(Bt + (buf * 2 + 1) * C::TILE, ws(tk.sb1) + tk.b1 * (int64_t)nrhs + coff, bytes, &b_full[buf]);
      }
      ++itb;
      __syncwarp();
    };
    for (int s = step0; s < step1; ++s) {
      const TreeStep stp = steps[s];
      const int j = s - step0;  // barrier steps completed by every CTA before this one
      if (trace && cta == 0) {  // diagnostics (hssb_debug_tree_trace): SM clock when step j may start
        grid_wait(arrive, ncta, base + j, lane);
        if (lane == 0) trace[j] = clock64();
      }
      if (stp.kind != TS_LEVEL) continue;  // exchange steps are run by consumer warp 0
      int cw, nsl, first, last;
      bool direct;
      level_items(stp, cw, nsl, first, last, direct);
      if (last <= first) continue;
      const int nch = stp.two ? 2 : 1;
      const GTask* tks = tasks + stp.task0;
      if (direct) {
        for (int item = first; item < last; ++item) load_a(tks[item / nsl], nch);
        continue;
      }
      // generators of the first item go out before the barrier: they do not depend on the previous level
      load_a(tks[first / nsl], nch);
      if (j > 0) grid_wait(arrive, ncta, base + j, lane);  // (the first step of a launch reads what earlier kernels wrote)
      load_b(tks[first / nsl], first % nsl, cw, nch);
      for (int item = first; item < last; ++item) {
        // operands of the NEXT item before the generators of this one's successor: with a ring that holds
        // one item's generators (rank 64) the operand fetch must not wait for the ring
        if (item + 1 < last) load_b(tks[(item + 1) / nsl], (item + 1) % nsl, cw, nch);
        if (item + 1 < last) load_a(tks[(item + 1) / nsl], nch);
      }
    }
    if (trace && cta == 0) {
      grid_wait(arrive, ncta, base + (unsigned long long)(step1 - step0), lane);
      if (lane == 0) trace[step1 - step0] = clock64();
    }
    return;
  }

  // ====================== consumer warps ======================
  int itb = 0, st = 0;
  uint32_t ph = 0;
  for (int s = step0; s < step1; ++s) {
    const TreeStep stp = steps[s];
    const int j = s - step0;
    if (stp.kind != TS_LEVEL) {
      if (warp != 0) continue;
      const int P = xq.nranks, me = xq.rank;
      if (P > 0) {
        unsigned long long* mine = xq.flags[me];
        // e = epoch of this product; the epoch word is only advanced by the ACK step of this launch
        const unsigned long long e = ld_volatile_sys(mine + 2 * P) + 1;
        if (stp.kind == TS_XCHG) {
          const int peer = cta / XCHG_PARTS, part = cta % XCHG_PARTS;
          const bool pusher = peer < P && peer != me;
          if (pusher || cta == 0) grid_wait(arrive, ncta, base + j, lane);  // the subtree-root Z block is complete
          if (pusher) {
            // the peer must have consumed the previous epoch before its copy of my slot is overwritten
            if (lane == 0) spin_until_ge(mine + P + peer, e - 1, TRAP_PEER_ACK);
            __syncwarp();
            const long long n2 = xq.slot_elems / 2;
            const long long i0 = n2 * part / XCHG_PARTS, i1 = n2 * (part + 1) / XCHG_PARTS;
            const double2* src = reinterpret_cast<const double2*>(xq.z[me] + xq.slot_off + (long long)me * xq.slot_elems);
            double2* dst = reinterpret_cast<double2*>(xq.z[peer] + xq.slot_off + (long long)me * xq.slot_elems);
            for (long long i = i0 + lane; i < i1; i += 32) dst[i] = __ldcg(src + i);
            __threadfence_system();
            __syncwarp();
            if (lane == 0)
              asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(xq.flags[peer] + XCHG_PART0 + me * XCHG_PARTS + part), "l"(e) : "memory");
          }
          if (cta == 0) {  // everybody else's slot has landed in my workspace
            for (int w = lane; w < P * XCHG_PARTS; w += 32)
              if (w / XCHG_PARTS != me) spin_until_ge(mine + XCHG_PART0 + w, e, TRAP_PEER_DATA);
            __threadfence_system();
            __syncwarp();
          }
        } else if (cta == 0) {  // TS_ACK: the gathered slots have been consumed, the peers may overwrite them
          grid_wait(arrive, ncta, base + j, lane);
          if (lane < P && lane != me)
            asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(xq.flags[lane] + P + me), "l"(e) : "memory");
          __syncwarp();
          if (lane == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(mine + 2 * P), "l"(e) : "memory");
        }
      }
      __syncwarp();
      if (lane == 0) st_release_gpu(arrive + cta, base + (unsigned long long)j + 1);
      continue;
    }
    int cw, nsl, first, last;
    bool direct;
    level_items(stp, cw, nsl, first, last, direct);
    if (last > first) {
      const int nch = stp.two ? 2 : 1;
      const GTask* tk = tasks + stp.task0;
      if (direct) {
        if (j > 0) grid_wait(arrive, ncta, base + j, lane);
        tree_consume_direct<R>(tk, first, last, nsl, nch, p, As, a_full, a_empty, st, ph, warp, lane);
      } else if (cw == 64) {
        tree_consume_level<R, 64>(tk, first, last, nsl, nch, p, As, Bt, a_full, a_empty, b_full, b_empty, itb, st, ph, warp, lane);
      } else if (cw == 32) {
        tree_consume_level<R, 32>(tk, first, last, nsl, nch, p, As, Bt, a_full, a_empty, b_full, b_empty, itb, st, ph, warp, lane);
      } else {
        tree_consume_level<R, 16>(tk, first, last, nsl, nch, p, As, Bt, a_full, a_empty, b_full, b_empty, itb, st, ph, warp, lane);
      }
      // all eight warps' stores of this level are ordered before warp 0's release below
      fence_proxy_async();
      named_bar_sync(1, 256);
    }
    if (warp == 0 && lane == 0) st_release_gpu(arrive + cta, base + (unsigned long long)j + 1);
  }
}

// ================================================================ host side ===
struct TreePlan {
  int phase0 = -1, phase1 = -1;     // phases [phase0, phase1) of H->phases are covered by the tree kernel
  std::vector<TreeStep> steps;      // one per covered phase, same order
  int xchg_step = -1;               // index of the TS_XCHG step (sharded handles)
  TreeStep* steps_dev = nullptr;
  unsigned long long* arrive = nullptr;  // grid-barrier words, one per CTA
  int grid = 0;
};

static void free_tree(hssb_matrix* H) {
  TreePlan* tp = (TreePlan*)H->tree_plan;
  if (!tp) return;
  cudaFree(tp->steps_dev);
  cudaFree(tp->arrive);
  delete tp;
  H->tree_plan = nullptr;
}

// Which phases of the forward plan the tree kernel can run: everything between the two leaf phases,
// provided every level is tagged for the fixed-shape kernels (uniform tree, rank 16 / 32 / 64).
static bool tree_plan_host(const hssb_matrix* H, TreePlan& tp) {
  tp.phase0 = tp.phase1 = -1;
  tp.steps.clear();
  tp.xchg_step = -1;
  if (!H->uniform || !H->padded) return false;
  if (H->uni_r != 16 && H->uni_r != 32 && H->uni_r != 64) return false;
  const auto& ph = H->phases;
  for (size_t i = 0; i < ph.size(); ++i) {
    const Phase& q = ph[i];
    if (q.kind == PH_LEAF_UP || q.kind == PH_LEAF_DOWN) continue;
    TreeStep s;
    memset(&s, 0, sizeof(s));
    if (q.kind == PH_EXCHANGE) { s.kind = TS_XCHG; tp.xchg_step = (int)tp.steps.size(); }
    else if (q.kind == PH_XCHG_ACK) s.kind = TS_ACK;
    else {
      if (q.fast != FAST_MERGE && q.fast != FAST_TRANSLATE) return false;
      if (q.fast_m || q.fast_r) return false;
      if (q.task0 > INT32_MAX || q.ntasks > INT32_MAX / 8) return false;
      s.kind = TS_LEVEL; s.task0 = (int32_t)q.task0; s.ntasks = (int32_t)q.ntasks;
      s.two = H->tasks_host[(size_t)q.task0].K1 > 0 ? 1 : 0;
    }
    if (tp.phase0 < 0) tp.phase0 = (int)i;
    if ((int)i != tp.phase0 + (int)tp.steps.size()) return false;  // covered phases must be contiguous
    tp.steps.push_back(s);
    tp.phase1 = (int)i + 1;
  }
  return !tp.steps.empty();
}

static int ensure_tree_plan(hssb_matrix* H) {
  if (H->tree_plan) return HSSB_OK;
  std::unique_ptr<TreePlan> tp(new (std::nothrow) TreePlan());
  if (!tp) HSSB_FAIL(HSSB_ERR_ALLOC, "tree plan: out of memory");
  if (tree_plan_host(H, *tp)) {
    int sms = 148;
    HSSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, H->device));
    tp->grid = sms;  // one CTA per SM (cooperative launch: all of them are resident)
    HSSB_CUDA(cudaMalloc(&tp->steps_dev, tp->steps.size() * sizeof(TreeStep)));
    HSSB_CUDA(cudaMemcpy(tp->steps_dev, tp->steps.data(), tp->steps.size() * sizeof(TreeStep), cudaMemcpyHostToDevice));
    HSSB_CUDA(cudaMalloc(&tp->arrive, (size_t)tp->grid * sizeof(unsigned long long)));
    HSSB_CUDA(cudaMemset(tp->arrive, 0, (size_t)tp->grid * sizeof(unsigned long long)));
  }
  H->tree_plan = tp.release();
  return HSSB_OK;
}

template <int R>
static int launch_tree_r(hssb_matrix* H, const TreePlan* tp, int s0, int s1, const CallParams& cp, const XchgParams& xq, cudaStream_t st,
                         long long* trace) {
  using C = TreeCfg<R>;
  if (s1 <= s0) return HSSB_OK;
  if (!H->fast_state) {
    FastState* fs = new FastState();
    cudaDeviceGetAttribute(&fs->num_sms, cudaDevAttrMultiProcessorCount, H->device);
    H->fast_state = fs;
  }
  FastState* fs = (FastState*)H->fast_state;
  if (int rc = fs->configure((const void*)tree_kernel<R>, C::SMEM)) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)tp->grid);
  cfg.blockDim = dim3(288);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // the grid barrier needs every CTA resident
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = H->tree_kernel == 2 ? 0 : 1;  // 2: plain launch (diagnostics; safe only on an otherwise idle device)
  HSSB_CUDA(cudaLaunchKernelEx(&cfg, tree_kernel<R>, (const GTask*)H->tasks_dev, (const TreeStep*)tp->steps_dev, s0, s1, cp,
                               tp->arrive, xq, trace));
  H->launches++;
  return HSSB_OK;
}

static int launch_tree(hssb_matrix* H, const TreePlan* tp, int s0, int s1, const CallParams& cp, const XchgParams& xq, cudaStream_t st,
                       long long* trace = nullptr) {
  switch ((int)H->uni_r) {
    case 16: return launch_tree_r<16>(H, tp, s0, s1, cp, xq, st, trace);
    case 32: return launch_tree_r<32>(H, tp, s0, s1, cp, xq, st, trace);
    case 64: return launch_tree_r<64>(H, tp, s0, s1, cp, xq, st, trace);
  }
  HSSB_FAIL(HSSB_ERR_STATE, "no tree kernel for rank %d", (int)H->uni_r);
}

}  // namespace hssb
