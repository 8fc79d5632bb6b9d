// hssb_leaf2.cuh — leaf kernels, second generation: every ring stage is SELF-CONTAINED.
//
// stream_leaf_kernel (hssb_fast.cuh) keeps the whole X block of an item (K0 x NT, 64 KB for config 3)
// resident and double buffered beside a ring of A chunks.  That costs 128 KB of shared memory, leaves
// room for only 6 (leaf down) or 2 (leaf up: a chunk is the whole V') ring stages, and makes the
// leaf-up kernel a whole-item double buffer: measured 0.19 ms where moving the data alone takes 0.18 ms
// and computing alone 0.16 ms -- load and compute barely overlap.
//
// Here a stage holds the KC columns of [D | U] (or V') AND the matching KC rows of X:
//     stage = [ X slab: KC rows x NT right-hand sides, 128B-swizzled TMA boxes | A chunk: MO x KC, padded ]
// so the consumers need nothing but the stage (plus F for the U chunks), X is never double buffered and
// all of shared memory is ring: 8-9 stages of 25 KB, i.e. two items of prefetch for the leaf-up
// kernel at chunk granularity.  Same arithmetic in the same order as the first generation (bit-identical
// results), same warp tiling, same conflict-free fragment addressing.
#pragma once

namespace hssb {

template <int M, int R, bool DOWN, int NT_, int KC_>
struct Leaf2Cfg {
  static constexpr int MO = DOWN ? M : R;
  static constexpr int K0 = M, K1 = DOWN ? R : 0;
  static constexpr int NT = NT_, KC = KC_;
  static constexpr int NWARPS = 8;
  static constexpr int WR = DOWN ? M / 32 : ((R >= 32 ? 2 : 1) > 64 / NT ? (R >= 32 ? 2 : 1) : 64 / NT);
  static constexpr int WC = NWARPS / WR;
  static constexpr int TM = MO / WR / 8, TN = NT / WC / 8;
  static constexpr int KSTEPS = KC / 4;
  static constexpr int NCH0 = K0 / KC, NCH1 = K1 / KC, NCH = NCH0 + NCH1;
  static constexpr int LDF = K1 + 4, LDA = MO + 4;
  static constexpr int XS_BYTES = KC * NT * 8;          // KC / 16 boxes of {16 rows x NT columns}
  static constexpr int A_BYTES = KC * LDA * 8;
  static constexpr int STAGE_BYTES = (XS_BYTES + A_BYTES + 1023) / 1024 * 1024;  // swizzled boxes need 1024-byte alignment
  static constexpr int BAR_BYTES = 256;
  static constexpr int F_BYTES = K1 ? NT * LDF * 8 : 0;
  static constexpr int FIT = (232448 - BAR_BYTES - F_BYTES) / STAGE_BYTES;
  static constexpr int NSTAGE = FIT > 12 ? 12 : FIT;
  // F(i) is requested after chunk CX of item i has been queued: by then the consumers have left item
  // i-1 (the producer is at most NSTAGE chunks ahead), so the wait on f_empty never holds up the ring
  static constexpr int CX = (NSTAGE - 1 < NCH0 - 1) ? NSTAGE - 1 : NCH0 - 1;
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + F_BYTES + BAR_BYTES;
  static_assert(MO % (8 * WR) == 0 && NT % (8 * WC) == 0 && TM >= 1 && TN >= 1, "warp tiling");
  static_assert(KC % 16 == 0 && K0 % KC == 0 && K1 % KC == 0 && NCH0 >= 1, "chunks are whole 16-row X slabs");
  static_assert(NSTAGE >= 3 && 2 * NSTAGE + 2 <= BAR_BYTES / 8 && SMEM <= 232448, "shared memory budget");
};

template <int M, int R, bool DOWN, int NT_, int KC_>
__global__ void __launch_bounds__(Leaf2Cfg<M, R, DOWN, NT_, KC_>::NWARPS * 32 + 32, 1)
leaf2_kernel(const GTask* __restrict__ tasks, int ntasks, int ntiles, CallParams p, const __grid_constant__ CUtensorMap xmap) {
  using C = Leaf2Cfg<M, R, DOWN, NT_, KC_>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = smem_raw;                                               // [NSTAGE][X slab | A chunk]
  double* Fs = reinterpret_cast<double*>(smem_raw + (size_t)C::NSTAGE * C::STAGE_BYTES);  // [NT][LDF]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + C::SMEM - C::BAR_BYTES);
  uint64_t* a_full = bars;               // [NSTAGE]
  uint64_t* a_empty = bars + C::NSTAGE;  // [NSTAGE]
  uint64_t* f_full = bars + 2 * C::NSTAGE;
  uint64_t* f_empty = f_full + 1;

  pdl_launch_dependents();  // HSSB_OPT_PDL (hssb_fast.cuh): the next kernel of the schedule may be placed while this one runs ...
  pdl_wait();               // ... and this one touches nothing before its predecessor is complete
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nitems = ntasks * ntiles;
  const int first = (int)(((int64_t)blockIdx.x * nitems) / gridDim.x);
  const int last = (int)(((int64_t)(blockIdx.x + 1) * nitems) / gridDim.x);
  const int my = last - first;
  const int nrhs = p.nrhs;

  if (tid == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], C::NWARPS); }
    mbar_init(f_full, 1);
    mbar_init(f_empty, C::NWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (my <= 0) return;

#if HSSB_ENABLE_DEBUG_MODES  // measurement builds only (make DEBUG_MODES=1): tools/leaf_bounds.py
  const bool dbg_nowait = p.debug & 1, dbg_nomma = p.debug & 2, dbg_nostore = p.debug & 4;
#else
  constexpr bool dbg_nowait = false, dbg_nomma = false, dbg_nostore = false;
#endif

  if (warp == C::NWARPS) {
    if (dbg_nowait) return;
    // ====================== producer warp ======================
    // Lane 0 queues the A chunk (one contiguous bulk copy of the padded pool image), lanes 0 .. KC/16-1
    // the X boxes (columns beyond nrhs are zero filled by the TMA unit and still counted).
    int g = 0;
    for (int item = 0; item < my; ++item) {
      const GTask& tk = tasks[(first + item) / ntiles];
      const int tile = (first + item) % ntiles;
      for (int c = 0; c < C::NCH; ++c, ++g) {
        const int st = g % C::NSTAGE;
        unsigned char* sp = stages + (size_t)st * C::STAGE_BYTES;
        mbar_wait(&a_empty[st], ((g / C::NSTAGE) & 1) ^ 1);
        const bool xpart = c < C::NCH0;
        if (lane == 0) {
          mbar_expect_tx(&a_full[st], (uint32_t)(C::A_BYTES + (xpart ? C::XS_BYTES : 0)));
          const double* src = p.pool + (xpart ? tk.a0 + (int64_t)c * C::KC * C::LDA : tk.a1 + (int64_t)(c - C::NCH0) * C::KC * C::LDA);
          bulk_g2s(sp + C::XS_BYTES, src, C::A_BYTES, &a_full[st]);
        }
        __syncwarp();
        if (xpart && lane < C::KC / 16)
          tma_load_2d(sp + lane * (C::NT * 16 * 8), &xmap, (int)tk.b0 + c * C::KC + lane * 16, tile * C::NT, &a_full[st]);
        if (C::K1 && c == C::CX) {
          const int nc = min(C::NT, nrhs - tile * C::NT);
          mbar_wait(f_empty, (item & 1) ^ 1);
          if (lane == 0) {  // the F workspace already carries the padded leading dimension: one copy per tile
            const uint32_t bytes = (uint32_t)(nc * C::LDF * 8);
            mbar_expect_tx(f_full, bytes);
            bulk_g2s(Fs, p.F + tk.b1 * (int64_t)nrhs + (int64_t)tile * C::NT * C::LDF, bytes, f_full);
          }
        }
      }
    }
    return;
  }

  // ====================== consumer warps ======================
  const int gq = lane >> 2, t = lane & 3;
  const int wr = warp % C::WR, wc = warp / C::WR;
  const int pg = perm8(gq);  // conflict-free column assignment, see stream_leaf_kernel
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
  int sw[4];  // swizzled offset of this lane's element in k-step q of a 16-row slab (row pg of the box)
#pragma unroll
  for (int q = 0; q < 4; ++q) sw[q] = ((q * 2 + (t >> 1)) ^ pg) * 2 + (t & 1);
  // One chunk: A fragments ("N" operand) from the stage, B fragments from the stage's X slab (XPART) or
  // from the padded F block.  The try_wait for the next chunk is issued two k-steps before the end and
  // only checked after this chunk's DMMAs are out.
  auto chunk = [&](const double* A, const double* B, auto xpart_tag, const NextWait& nw) -> bool {
    constexpr bool XPART = decltype(xpart_tag)::value;
    bool ok = true;
#pragma unroll
    for (int kk = 0; kk < C::KSTEPS; ++kk) {
      double a[C::TM], b[C::TN];
#pragma unroll
      for (int i = 0; i < C::TM; ++i) a[i] = A[kk * 4 * C::LDA + i * 8];
#pragma unroll
      for (int j = 0; j < C::TN; ++j)
        b[j] = XPART ? B[(kk >> 2) * (C::NT * 16) + j * 8 * 16 + sw[kk & 3]] : B[j * 8 * C::LDF + kk * 4];
      if (kk == C::KSTEPS - 2) ok = try_next(nw);
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
  uint32_t ph = 0;
  if (!dbg_nowait) mbar_wait(&a_full[0], 0);
  const int a_off = C::XS_BYTES / 8 + wr * (C::TM * 8) + gq + t * C::LDA;  // doubles from the stage base
  const int x_off = (wc * (C::TN * 8) + pg) * 16;
  for (int item = 0; item < my; ++item) {
    const bool more = item + 1 < my;
    const int task_i = (first + item) / ntiles, tile = (first + item) - task_i * ntiles;
    const int64_t out_row = tasks[task_i].c;
#pragma unroll 1
    for (int c = 0; c < C::NCH; ++c) {
      const int nst = (st + 1 == C::NSTAGE) ? 0 : st + 1;
      const uint32_t nph = (st + 1 == C::NSTAGE) ? ph ^ 1 : ph;
      const bool last_c = c == C::NCH - 1;
      NextWait nw{nullptr, 0, nullptr, 0};
      if (!dbg_nowait && (!last_c || more)) {
        nw.b0 = &a_full[nst]; nw.p0 = nph;
        if (C::K1 && c == C::NCH0 - 1) { nw.b1 = f_full; nw.p1 = item & 1; }
      }
      const double* S = reinterpret_cast<const double*>(stages + (size_t)st * C::STAGE_BYTES);
      bool ok;
      if (c < C::NCH0) ok = chunk(S + a_off, S + x_off, std::true_type{}, nw);
      else ok = chunk(S + a_off, Fs + (wc * (C::TN * 8) + pg) * C::LDF + t + (c - C::NCH0) * C::KC, std::false_type{}, nw);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_empty[st]);
        if (C::K1 && last_c) mbar_arrive(f_empty);
      }
      if (!ok) spin_next(nw);
      st = nst; ph = nph;
    }
    // ---- epilogue of the item
    const int ncols = min(C::NT, nrhs - tile * C::NT);
    double* O;
    int64_t ldo;
    if (DOWN) { O = p.Y + out_row + (int64_t)tile * C::NT * p.ldy; ldo = p.ldy; }
    else { O = p.Z + out_row * (int64_t)nrhs + (int64_t)tile * C::NT * (R + 4); ldo = R + 4; }
    O += (int64_t)(wc * (C::TN * 8)) * ldo + wr * (C::TM * 8) + gq;
    const int colb = wc * (C::TN * 8);
    const int pc[2] = {perm8(2 * t), perm8(2 * t + 1)};
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

}  // namespace hssb
