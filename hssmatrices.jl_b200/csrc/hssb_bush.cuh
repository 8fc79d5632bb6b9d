// hssb_bush.cuh — the merge / translate levels of a small any-shape tree (BASELINE configs 1-2: what a compression
// produces) in ONE launch with a SHORT dependent chain: the tree is cut into "bushes" of a few levels, one CTA runs a
// whole bush out of shared memory.
//
// The product of matmul.jl:32-62 is 2*depth + 2 dependent levels.  For a tree that fits the L2 (config 2: 1 GFLOP,
// 118 MB) every level between the two leaf phases costs a launch or, in the dataflow kernel of hssb_flow.cuh, a flag
// round trip through the L2 plus a staged 64 x 64 tile for a 17 x 17 block: 20 levels x 4-9 us.  Here
//   * the tree is cut at a few depths; the merges below one node down to the next cut (matmul.jl:39), or the
//     translates below it (:52-56), or both for the bush that holds the root, form ONE work item;
//   * everything an item reads -- its ops, the generator blocks, this column tile of the Z / F blocks other bushes
//     (or the leaf-up kernel) produced -- is brought into shared memory by bulk copies (cp.async.bulk, one
//     instruction per block, completion counted in bytes on an mbarrier): generators while the CTA still waits for
//     its producers, workspace tiles right after;
//   * the CTA then runs the item level by level with a __syncthreads in between; Z / F blocks that are produced and
//     consumed inside the bush never leave shared memory unless somebody outside needs them;
//   * only bushes talk through global flags: 7 dependent steps for config 2 instead of 20;
//   * the unit of work is a WARP: 8 rows x 16 right-hand sides of one task, operands read straight from the
//     shared-memory images -- no barrier inside a task, so a level of eight small blocks is 24 independent warp tasks.
//     They run on the FP64 FMA pipe, not on DMMA: a warp that works alone waits ~500 cycles for a dependent DMMA.
// The leaf phases stay on the tile kernel (they are throughput work on 64-row blocks): leaf-up launch, this kernel,
// leaf-down launch.  Items (bush, 16-column tile) are drawn in a topological order of the bushes from one atomic
// counter by the resident CTAs, as in hssb_flow.cuh: a CTA only waits for items drawn before its own, so there is no
// deadlock and no cooperative launch.
//
// Measured lesson (profiles/bush_kernel_r02.txt): at this granularity a warp is bound by the latency of its own
// instruction stream, so the per-op code is rolled and short, and nothing on an item's critical path touches global
// memory one word at a time.
#pragma once

#include <algorithm>
#include <functional>
#include <map>
#include <queue>
#include <set>
#include <unordered_map>

#include "hssb_flow.cuh"

namespace hssb {

constexpr int B_TN = 16, B_THREADS = 256, B_WARPS = B_THREADS / 32;

struct BushOp {            // one warp's work: rows [m0, m0 + mr) of one task, one 16-column tile
  int64_t a0, a1;          // pool offsets of the two A operands (whole block)
  int64_t b0, b1, c;       // global operands as in GTask: X / Y first row, or workspace row offset
  int32_t lda0, lda1, ldb0, ldb1, ldc;
  int32_t m0, mr;          // row chunk, mr <= 8
  int32_t K0, K1;
  int32_t s0, s1, sc_off;  // shared-memory image of B0 / B1 / the output block (offset in doubles), -1: none
  int32_t lds0, lds1, ldsc;
  int32_t sa0, sa1;        // shared-memory copy of the A blocks (same leading dimension), -1: read from the pool
  uint8_t ta0, ta1, sb0, sb1, sc, epilogue, to_global, pad;
  int32_t pad2[2];
};
static_assert(sizeof(BushOp) == 128, "BushOp is copied to shared memory in 16-byte pieces");

// One bulk copy into shared memory at the start of a bush.  ST_POOL: `count` contiguous doubles of the generator pool;
// ST_Z / ST_F: one 16-column tile of a workspace block (src = workspace row, count = ld * 16, contiguous); ST_OPS: the
// bush's ops.
enum BushStageKind : int { ST_POOL = 0, ST_Z = SRC_Z, ST_F = SRC_F, ST_OPS = 4 };
struct BushStage {
  int64_t src;
  int32_t dst, count;      // doubles
  int32_t ld, kind;
};

constexpr int B_MAXLEV = 16;
struct BushHdr {
  int32_t op0, nops, nlevels;
  int32_t dep0, ndeps;     // bushes whose flags this one waits for
  int32_t st0, nst_pre, nst_post;  // staged copies issued before the wait (ops, generators) and after it (workspace tiles)
  int32_t ops_dst;         // shared-memory copy of the ops (doubles), -1: read from global memory
  int32_t pre_bytes;       // bytes of the copies issued before the wait
  int32_t post_ld;         // sum of the leading dimensions of the workspace tiles (bytes = post_ld * valid columns * 8)
  int32_t lvl_end[B_MAXLEV];  // level l = ops [lvl_end[l - 1], lvl_end[l]) of the bush
};

struct BushParams {
  const BushOp* ops;
  const BushHdr* hdr;
  const BushStage* stages;
  const int32_t* deps;
  unsigned int* sync;      // [0] = next item, [1 + bush * ncol + coltile] = done
  int32_t nbush;
  unsigned long long* trace;  // diagnostics (hssb_debug_bush_trace): B_TRACE words per item, or NULL
  int32_t probe_item;         // ... and cycle stamps inside the ops of this item, after the per-item words
};
constexpr int B_PROBE_LEVELS = 8;

constexpr int B_TRACE = 12;   // item: [0] SM, [1] drawn, [2] dependencies met, [3 .. 3 + 7] end of level l, [11] flag published (ns, globaltimer)
__device__ __forceinline__ unsigned long long bush_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One warp, one op: 8 rows x 16 right-hand sides of one task, lane (g, q) = row g, columns q, q + 4, q + 8, q + 12
// (columns 4 q .. 4 q + 3 would put the four q of a B read on one bank for every even leading dimension), on the FP64
// FMA pipe -- NOT on DMMA.  Measured with cycle stamps inside the op (tools/bush_trace.py, profiles/bush_kernel_r02.txt):
// a warp that works alone waits ~500 cycles for a DMMA whose accumulator it needs again, so a rank-17 merge (10 k-steps,
// even spread over four accumulator sets) spent 4000 cycles in its k loops; a dependent DFMA costs a handful of cycles,
// and the two pipes have the same peak, which these blocks are nowhere near.  Two accumulator sets (even /
// odd k) halve the FMA chain.  Everything the op reads is normally in shared memory (ALLSM: 32-bit offsets, LDS); a
// workspace block that could not be staged is read with ld.global.cg -- it was written in this launch by another SM and
// must not come from a stale L1 line.
template <bool ALLSM>
__device__ __forceinline__ void bush_warp_op(const BushOp& o, const CallParams& p, double* sm, const int n0, const int lane, long long* probe) {
  if (probe && lane == 0) probe[1] = clock64();
  const int g = lane >> 2, q = lane & 3;
  const int N = p.nrhs;
  const int mr = o.mr, m0 = o.m0;
  const int row = g < mr ? g : mr - 1;         // lanes past the chunk recompute its last row and store nothing
  double acc[2][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[u][j] = 0.0;
  if (probe && lane == 0) probe[2] = clock64();
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const int K = s ? o.K1 : o.K0;
    if (K <= 0) continue;
    const bool ta = s ? o.ta1 : o.ta0;
    const int lda = s ? o.lda1 : o.lda0;
    const int sa = s ? o.sa1 : o.sa0;
    const int soff = s ? o.s1 : o.s0;
    if (ALLSM) {
      const int ldb = s ? o.lds1 : o.lds0;
      const int a_k = ta ? 1 : lda;
      int ia = sa + (ta ? (m0 + row) * lda : m0 + row);
      int ib = soff + q * ldb;
      int k = 0;
#pragma unroll 2
      for (; k + 1 < K; k += 2) {
        const double a0 = sm[ia], a1 = sm[ia + a_k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] = fma(a0, sm[ib + 4 * j * ldb], acc[0][j]);
          acc[1][j] = fma(a1, sm[ib + 4 * j * ldb + 1], acc[1][j]);
        }
        ia += 2 * a_k; ib += 2;
      }
      if (k < K) {
        const double a0 = sm[ia];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[0][j] = fma(a0, sm[ib + 4 * j * ldb], acc[0][j]);
      }
    } else {
      const double* pa = (sa >= 0 ? sm + sa : p.pool + (s ? o.a1 : o.a0)) + (ta ? (int64_t)(m0 + row) * lda : (int64_t)(m0 + row));
      const int64_t a_k = ta ? 1 : (int64_t)lda;
      const double* pb;
      int64_t ldb;
      const bool bsm = soff >= 0;
      if (bsm) {
        pb = sm + soff;
        ldb = s ? o.lds1 : o.lds0;
      } else {
        pb = operand_b(p, s ? o.sb1 : o.sb0, s ? o.b1 : o.b0, s ? o.ldb1 : o.ldb0, ldb) + (int64_t)n0 * ldb;
      }
      const bool cg = !bsm && (s ? o.sb1 : o.sb0) != SRC_X;
      bool cv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) cv[j] = bsm || n0 + q + 4 * j < N;   // global memory ends at column nrhs
      pb += (int64_t)q * ldb;
#pragma unroll 1
      for (int k = 0; k < K; ++k) {
        const double a0 = pa[k * a_k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double bv = !cv[j] ? 0.0 : (cg ? __ldcg(pb + 4 * j * ldb + k) : pb[4 * j * ldb + k]);
          acc[0][j] = fma(a0, bv, acc[0][j]);
        }
      }
    }
  }
  if (probe && lane == 0) probe[3] = clock64();
  // lane (g, q) holds C[m0 + g][n0 + q + 4 j]
  int64_t ldc = 0;
  double* C = nullptr;
  if (o.to_global) {
    switch (o.sc) {
      case SRC_Z: ldc = o.ldc; C = p.Z + o.c * (int64_t)N; break;
      case SRC_F: ldc = o.ldc; C = p.F + o.c * (int64_t)N; break;
      default: ldc = p.ldy; C = p.Y + o.c; break;
    }
    C += (int64_t)(n0 + q) * ldc + m0 + g;
  }
  const int sc_off = o.sc_off, ldsc = o.ldsc;
  const bool epi = o.epilogue;
  const double alpha = p.alpha, beta = p.beta;
  if (g < mr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double v = acc[0][j] + acc[1][j];
      if (probe && lane == 0 && j == 0) probe[4] = v == 12345.678 ? 0 : clock64();
      if (sc_off >= 0) sm[sc_off + (q + 4 * j) * ldsc + m0 + g] = v;  // columns past nrhs hold whatever the operands held there: never stored
      if (C && n0 + q + 4 * j < N) {
        double* dst = C + (int64_t)4 * j * ldc;
        if (epi) {
          v *= alpha;
          if (beta != 0.0) v += beta * (*dst);  // beta == 0 never reads Y (matmul.jl:13)
        }
        *dst = v;
      }
    }
  }
  if (probe && lane == 0) probe[5] = clock64();
}

__device__ __forceinline__ void bush_stage_issue(const BushStage& st, const BushParams& f, const CallParams& p, double* sm,
                                                 const int n0, const int ncv, const int op0, uint64_t* bar) {
  const double* src;
  int count = st.count;
  if (st.kind == ST_POOL) src = p.pool + st.src;
  else if (st.kind == ST_OPS) src = (const double*)(f.ops + op0);
  else {
    src = (st.kind == ST_Z ? p.Z : p.F) + st.src * (int64_t)p.nrhs + (int64_t)n0 * st.ld;
    count = st.ld * ncv;
  }
  bulk_g2s(sm + st.dst, src, (uint32_t)count * 8u, bar);
}

constexpr int B_MAXST = 160;   // stage descriptors held in shared memory (more are read from global memory)

__global__ void __launch_bounds__(B_THREADS, 2)
bush_kernel(BushParams f, CallParams p) {
  extern __shared__ __align__(128) double bush_sm[];
  __shared__ int s_idx[2];
  __shared__ BushHdr s_hdr;
  __shared__ BushStage s_st[B_MAXST];
  __shared__ __align__(8) uint64_t s_bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = p.nrhs;
  const int ncol = (N + B_TN - 1) / B_TN;
  const long long total = (long long)f.nbush * ncol;
  unsigned int* done = f.sync + 1;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    s_idx[0] = (int)atomicAdd(f.sync, 1u);
  }
  __syncthreads();
  uint32_t parity = 0;
  for (int it = 0;; ++it) {
    const int idx = s_idx[it & 1];
    if (idx >= total) break;
    if (tid == 0) s_idx[(it + 1) & 1] = (int)atomicAdd(f.sync, 1u);  // read after this item's barriers
    const int b = idx / ncol, j = idx - b * ncol;
    const int n0 = j * B_TN;
    const int ncv = N - n0 < B_TN ? N - n0 : B_TN;
    unsigned long long* tr = f.trace ? f.trace + (size_t)idx * B_TRACE : nullptr;
    if (tr && tid == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      tr[0] = smid; tr[1] = bush_now();
    }
    // the bush's descriptors: two round trips (header, then every stage descriptor at once), not one per stage
    if (tid < (int)(sizeof(BushHdr) / sizeof(int))) ((int*)&s_hdr)[tid] = ((const int*)(f.hdr + b))[tid];
    __syncthreads();
    const int op0 = s_hdr.op0, nlevels = s_hdr.nlevels, ndeps = s_hdr.ndeps, dep0 = s_hdr.dep0, st0 = s_hdr.st0;
    const int nst_pre = s_hdr.nst_pre, nst = s_hdr.nst_pre + s_hdr.nst_post, ops_dst = s_hdr.ops_dst;
    for (int w = tid; w < 3 * (nst < B_MAXST ? nst : B_MAXST); w += B_THREADS)
      ((unsigned long long*)s_st)[w] = ((const unsigned long long*)(f.stages + st0))[w];
    __syncthreads();
    // what nobody produces -- the ops and the generator blocks -- travels while the CTA waits for its producers
    if (warp == 0) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the previous item's accesses to these bytes come first
      if (lane == 0) mbar_expect_tx(&s_bar[0], (uint32_t)s_hdr.pre_bytes);
#pragma unroll 1
      for (int e = lane; e < nst_pre; e += 32) bush_stage_issue(e < B_MAXST ? s_st[e] : f.stages[st0 + e], f, p, bush_sm, n0, ncv, op0, &s_bar[0]);
    }
    for (int d = tid; d < ndeps; d += B_THREADS) {
      const int dep = f.deps[dep0 + d];
      const unsigned int* flag = done + (size_t)dep * ncol + j;
      if (ld_acquire_u32(flag) == 0u) {
        const long long t0 = clock64();
        while (ld_acquire_u32(flag) == 0u)
          if (clock64() - t0 > 8000000000ll) trap_report(TRAP_FLOW, (unsigned long long)dep, (unsigned long long)b);
      }
    }
    __syncthreads();
    if (tr && tid == 0) tr[2] = bush_now();
    // this column tile of the Z / F blocks other bushes (or the leaf-up kernel) produced: one round trip for all of them
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(&s_bar[1], (uint32_t)(s_hdr.post_ld * ncv * 8));
#pragma unroll 1
      for (int e = nst_pre + lane; e < nst; e += 32) bush_stage_issue(e < B_MAXST ? s_st[e] : f.stages[st0 + e], f, p, bush_sm, n0, ncv, op0, &s_bar[1]);
    }
    mbar_wait(&s_bar[0], parity);
    mbar_wait(&s_bar[1], parity);
    parity ^= 1u;
    const BushOp* ops = ops_dst >= 0 ? (const BushOp*)(bush_sm + ops_dst) : f.ops + op0;
    int o0 = 0;
    for (int l = 0; l < nlevels; ++l) {
      const int o1 = s_hdr.lvl_end[l];
      // diagnostics: cycle stamps of the probed item, per level and warp (level start, op start, k loops start / end,
      // first accumulator read, op end, after the barrier)
      long long* probe = (tr && idx == f.probe_item && l < B_PROBE_LEVELS) ? (long long*)(f.trace + (size_t)total * B_TRACE) + (l * B_WARPS + warp) * 8 : nullptr;
      if (probe && lane == 0) probe[0] = clock64();
      for (int o = o0 + warp; o < o1; o += B_WARPS) {
        const BushOp& op = ops[o];
        if ((op.K0 <= 0 || (op.sa0 >= 0 && op.s0 >= 0)) && (op.K1 <= 0 || (op.sa1 >= 0 && op.s1 >= 0))) bush_warp_op<true>(op, p, bush_sm, n0, lane, probe);
        else bush_warp_op<false>(op, p, bush_sm, n0, lane, probe);
      }
      o0 = o1;
      if (probe && lane == 0) probe[6] = clock64();
      __syncthreads();  // the level's blocks (shared memory, and workspace blocks re-read by this CTA) are complete
      if (probe && lane == 0) probe[7] = clock64();
      if (tr && tid == 0 && l < 8) tr[3 + l] = bush_now();
    }
    if (tid == 0) {
      __threadfence();
      atomicExch(done + (size_t)b * ncol + j, 1u);
      if (tr) tr[11] = bush_now();
    }
  }
}

// ================================================================ host side ===
struct BushPlan {
  bool usable = false;
  std::string why;
  int32_t nbush = 0;
  int64_t task0 = 0, ntasks = 0;
  int smem_doubles = 0;
  std::vector<BushOp> ops;
  std::vector<BushHdr> hdr;
  std::vector<BushStage> stages;
  std::vector<int32_t> deps;
  std::vector<int32_t> op_task, op_bush, op_level, stage_bush;  // debug export (tests/plan_interp.py)
  BushOp* ops_dev = nullptr;
  BushHdr* hdr_dev = nullptr;
  BushStage* stages_dev = nullptr;
  int32_t* deps_dev = nullptr;
  unsigned int* sync_dev = nullptr;
  unsigned long long* trace_dev = nullptr;   // hssb_debug_bush_trace: sized with sync_dev while H->bush_trace is set
  int64_t sync_cols = 0;
  int grid_cap = 148 * 2;
};

constexpr int BUSH_SMEM_BUDGET = 14000;  // doubles (109 KB): two CTAs per SM

static void free_bush(hssb_matrix* H) {
  for (void*& v : H->bush_plan) {
    BushPlan* bp = (BushPlan*)v;
    if (!bp) continue;
    if (bp->ops_dev) {  // plan-only handles (CPU tests) never touch the device
      cudaFree(bp->ops_dev); cudaFree(bp->hdr_dev); cudaFree(bp->deps_dev); cudaFree(bp->stages_dev);
    }
    if (bp->sync_dev) cudaFree(bp->sync_dev);
    if (bp->trace_dev) cudaFree(bp->trace_dev);
    delete bp;
    v = nullptr;
  }
}

// Cut the merge / translate levels of the plan `mode` (0: Y = A X, 1: Y = A' X on the any-shape task table) into bushes.
//   hb = levels per bush, ht = merge levels of the bush that holds the root.
// A merge (node at depth x) goes to the bush of its nearest ancestor-or-self at an up-cut depth {ht + 1 + i hb}, a
// translate (child at depth x) to the bush of its nearest PROPER ancestor at a down-cut depth {1 + i hb}; what reaches
// the root forms the one bush that turns around.  Leaf phases are not part of any bush (tbush = -1): their outputs are
// ready before the kernel starts, their inputs are written to the workspace.  Levels inside a bush and the edges
// between bushes come from the operands.
static void bush_plan_host(const hssb_matrix* H, int mode, int hb, int ht, int budget, BushPlan& bp) {
  const std::vector<Phase>& phases = mode == 1 ? H->phases_t : H->phases;
  const auto& nodes = H->nodes;
  bp.usable = false;
  bp.ops.clear(); bp.hdr.clear(); bp.stages.clear(); bp.deps.clear(); bp.op_task.clear(); bp.op_bush.clear(); bp.op_level.clear(); bp.stage_bush.clear();
  if (H->n_shards != 1) { bp.why = "sharded handle (the exchange sits between the levels)"; return; }
  if (phases.empty() || nodes.empty()) { bp.why = "empty plan"; return; }
  hb = std::max(hb, 1); ht = std::max(ht, 0);
  int64_t t0 = -1, t1 = -1;
  for (const Phase& ph : phases) {
    if (ph.kind == PH_EXCHANGE || ph.kind == PH_XCHG_ACK) { bp.why = "plan contains an exchange"; return; }
    if (ph.ntasks == 0) continue;
    if (t0 < 0) t0 = ph.task0;
    else if (ph.task0 != t1) { bp.why = "phases are not contiguous in the task table"; return; }
    t1 = ph.task0 + ph.ntasks;
  }
  if (t0 < 0 || t1 - t0 > INT32_MAX / 8) { bp.why = "no tasks"; return; }
  const int64_t nt = t1 - t0;
  bp.task0 = t0; bp.ntasks = nt;
  std::vector<uint8_t> kind((size_t)nt, 0);   // 0 leaf phase, 1 merge, 2 translate
  for (const Phase& ph : phases)
    for (int64_t i = 0; i < ph.ntasks; ++i) kind[(size_t)(ph.task0 - t0 + i)] = ph.kind == PH_MERGE ? 1 : ph.kind == PH_TRANSLATE ? 2 : 0;

  // ---- which node does a task belong to: by its output block
  std::unordered_map<int64_t, int32_t> zmap, fmap;
  for (size_t i = 1; i < nodes.size(); ++i) {
    if (nodes[i].kw > 0) zmap.emplace(nodes[i].zoff, (int32_t)i);
    if (nodes[i].kr > 0) fmap.emplace(nodes[i].foff, (int32_t)i);
  }
  std::vector<char> cut_up((size_t)H->depth + 2, 0), cut_down((size_t)H->depth + 2, 0);
  for (int64_t d = ht + 1; d <= H->depth; d += hb) cut_up[(size_t)d] = 1;
  for (int64_t d = 1; d <= H->depth; d += hb) cut_down[(size_t)d] = 1;
  std::vector<int32_t> tbush((size_t)nt, -1);
  std::map<int64_t, int32_t> key2bush;
  int32_t nb = 0;
  for (int64_t i = 0; i < nt; ++i) {
    if (!kind[(size_t)i]) continue;
    const GTask& g = H->tasks_host[(size_t)(t0 + i)];
    if (g.sc != SRC_Z && g.sc != SRC_F) { bp.why = "a merge / translate that does not write a workspace block"; return; }
    const auto& mp = g.sc == SRC_Z ? zmap : fmap;
    auto it = mp.find(g.c);
    if (it == mp.end()) { bp.why = "a task's output block does not belong to a node"; return; }
    int64_t a = it->second;
    const bool down = kind[(size_t)i] == 2;
    const std::vector<char>& cut = down ? cut_down : cut_up;
    if (down) a = nodes[(size_t)a].parent >= 0 ? nodes[(size_t)a].parent : 0;
    while (a != 0 && !cut[(size_t)nodes[(size_t)a].depth]) a = nodes[(size_t)a].parent;
    const int64_t key = a == 0 ? 0 : 2 * a + (down ? 1 : 0);
    auto kb = key2bush.find(key);
    if (kb == key2bush.end()) kb = key2bush.emplace(key, nb++).first;
    tbush[(size_t)i] = kb->second;
  }
  if (nb == 0) { bp.why = "no merge / translate levels (the root's children are leaves)"; return; }

  // ---- producers, levels inside a bush, edges between bushes
  struct Key { int src; int64_t row; bool operator<(const Key& o) const { return src != o.src ? src < o.src : row < o.row; } };
  std::map<Key, int32_t> producer;
  std::vector<int32_t> prod((size_t)(2 * nt), -1), level((size_t)nt, 0);
  std::vector<uint8_t> used_in((size_t)nt, 0), used_out((size_t)nt, 0);
  std::vector<std::set<int32_t>> bdeps((size_t)nb);
  for (int64_t i = 0; i < nt; ++i) {
    const GTask& g = H->tasks_host[(size_t)(t0 + i)];
    const int32_t bi = tbush[(size_t)i];
    for (int o = 0; o < 2; ++o) {
      const int K = o ? g.K1 : g.K0, src = o ? g.sb1 : g.sb0;
      if (K <= 0 || (src != SRC_Z && src != SRC_F)) continue;
      auto it = producer.find(Key{src, o ? g.b1 : g.b0});
      if (it == producer.end()) { bp.why = "a workspace operand is not the output block of an earlier task"; return; }
      const int32_t pt = it->second, bpt = tbush[(size_t)pt];
      prod[(size_t)(2 * i + o)] = pt;
      if (bi >= 0 && bpt == bi) {
        level[(size_t)i] = std::max(level[(size_t)i], level[(size_t)pt] + 1);
        used_in[(size_t)pt] = 1;
      } else {
        if (bi >= 0 && bpt >= 0) bdeps[(size_t)bi].insert(bpt);
        used_out[(size_t)pt] = 1;
      }
    }
    if (g.sc == SRC_Z || g.sc == SRC_F)
      if (!producer.emplace(Key{g.sc, g.c}, (int32_t)i).second) { bp.why = "a workspace block is written twice"; return; }
  }
  // a bush must not read what a LATER leaf phase writes, nor feed an EARLIER one: leaf-up first, leaf-down last
  for (int64_t i = 0; i < nt; ++i) {
    if (tbush[(size_t)i] < 0) continue;
    for (int o = 0; o < 2; ++o) {
      const int32_t pt = prod[(size_t)(2 * i + o)];
      if (pt >= 0 && tbush[(size_t)pt] < 0 && pt > i) { bp.why = "internal: operand produced after its reader"; return; }
    }
  }

  // ---- topological order of the bushes (Kahn; ties by first task, which keeps the up-sweep first)
  std::vector<int32_t> first((size_t)nb, INT32_MAX), indeg((size_t)nb, 0), order, newid((size_t)nb, -1);
  std::vector<std::vector<int32_t>> succ((size_t)nb);
  for (int64_t i = nt - 1; i >= 0; --i)
    if (tbush[(size_t)i] >= 0) first[(size_t)tbush[(size_t)i]] = (int32_t)i;
  for (int32_t b = 0; b < nb; ++b)
    for (int32_t d : bdeps[(size_t)b]) { succ[(size_t)d].push_back(b); indeg[(size_t)b]++; }
  typedef std::pair<int32_t, int32_t> PI;
  std::priority_queue<PI, std::vector<PI>, std::greater<PI>> ready;
  for (int32_t b = 0; b < nb; ++b)
    if (!indeg[(size_t)b]) ready.push(PI(first[(size_t)b], b));
  while (!ready.empty()) {
    const int32_t b = ready.top().second;
    ready.pop();
    newid[(size_t)b] = (int32_t)order.size();
    order.push_back(b);
    for (int32_t s : succ[(size_t)b])
      if (--indeg[(size_t)s] == 0) ready.push(PI(first[(size_t)s], s));
  }
  if ((int32_t)order.size() != nb) { bp.why = "the bushes depend on each other in a cycle"; return; }

  // ---- per bush: ops by level, then the shared-memory layout [ops | images | workspace tiles | generators]
  std::vector<std::vector<int32_t>> btasks((size_t)nb);
  for (int64_t i = 0; i < nt; ++i)
    if (tbush[(size_t)i] >= 0) btasks[(size_t)tbush[(size_t)i]].push_back((int32_t)i);
  std::vector<int32_t> slot((size_t)nt, -1), slot_ld((size_t)nt, 0);
  bp.smem_doubles = 0;
  for (int32_t ob = 0; ob < nb; ++ob) {
    const int32_t b = order[(size_t)ob];
    const auto& ts = btasks[(size_t)b];
    int nlev = 0;
    for (int32_t i : ts) nlev = std::max(nlev, level[(size_t)i] + 1);
    if (nlev > B_MAXLEV) { bp.why = "a bush has more levels than the kernel's header holds (HSSB_OPT_BUSH_LEVELS)"; return; }
    BushHdr h;
    memset(&h, 0, sizeof(h));
    h.op0 = (int32_t)bp.ops.size(); h.nlevels = nlev;
    h.st0 = (int32_t)bp.stages.size();
    h.dep0 = (int32_t)bp.deps.size(); h.ndeps = (int32_t)bdeps[(size_t)b].size();
    for (int32_t d : bdeps[(size_t)b]) bp.deps.push_back(newid[(size_t)d]);
    // row chunks: 8 rows per warp
    std::vector<int> chunk_of((size_t)nlev, 8);
    int nops = 0;
    for (int l = 0; l < nlev; ++l) {
      int chunk = 8, n = 0;
      for (;; chunk >>= 1) {
        n = 0;
        for (int32_t i : ts)
          if (level[(size_t)i] == l) n += (H->tasks_host[(size_t)(t0 + i)].M + chunk - 1) / chunk;
        if (n >= B_WARPS || chunk == 8) break;
      }
      chunk_of[(size_t)l] = chunk;
      nops += n;
    }
    int bump = 0;
    auto take = [&](int doubles) { const int at = bump; bump += (doubles + 1) / 2 * 2; return at; };
    std::vector<BushStage> pre, post;
    auto stage = [&](std::vector<BushStage>& v, int kind_, int64_t src, int dst, int count, int ld) {
      BushStage e;
      memset(&e, 0, sizeof(e));
      e.src = src; e.dst = dst; e.count = count; e.ld = ld; e.kind = kind_;
      v.push_back(e);
    };
    h.ops_dst = -1;
    if (nops * 16 <= budget / 4) { h.ops_dst = take(nops * 16); stage(pre, ST_OPS, 0, h.ops_dst, nops * 16, 0); }
    for (int32_t i : ts) {   // images of the blocks that are consumed inside the bush
      if (!used_in[(size_t)i]) continue;
      const int M = H->tasks_host[(size_t)(t0 + i)].M;
      const int ld = (M + 3) / 8 * 8 + 4;  // = 4 mod 8: the B-fragment reads of a half warp hit 16 different banks
      if (bump + ld * B_TN <= budget) { slot[(size_t)i] = take(ld * B_TN); slot_ld[(size_t)i] = ld; }
    }
    std::map<Key, std::pair<int32_t, int32_t>> ws_stage;   // workspace block -> (offset, ld)
    std::map<int64_t, int32_t> pool_stage;                  // pool offset -> offset
    for (int32_t i : ts) {   // this column tile of the Z / F blocks other bushes or the leaf-up phase produced
      const GTask& g = H->tasks_host[(size_t)(t0 + i)];
      for (int s = 0; s < 2; ++s) {
        const int32_t pt = prod[(size_t)(2 * i + s)];
        if (pt < 0 || tbush[(size_t)pt] == b) continue;
        const Key key{s ? g.sb1 : g.sb0, s ? g.b1 : g.b0};
        const int ld = s ? g.ldb1 : g.ldb0;
        if (ws_stage.count(key) || (key.row & 1) || (ld & 1) || bump + ld * B_TN > budget) continue;
        const int at = take(ld * B_TN);
        ws_stage[key] = std::make_pair((int32_t)at, (int32_t)ld);
        stage(post, key.src, key.row, at, ld * B_TN, ld);
        h.post_ld += ld;
      }
    }
    for (int l = 0; l < nlev; ++l)   // generator blocks, level by level as far as the budget goes
      for (int32_t i : ts) {
        const GTask& g = H->tasks_host[(size_t)(t0 + i)];
        if (level[(size_t)i] != l) continue;
        for (int s = 0; s < 2; ++s) {
          const int K = s ? g.K1 : g.K0;
          if (K <= 0) continue;
          const int64_t off = s ? g.a1 : g.a0;
          const int lda = s ? g.lda1 : g.lda0;
          const bool ta = s ? g.ta1 : g.ta0;
          const int rows = ta ? K : g.M, cols = ta ? g.M : K;
          const int64_t count = (int64_t)lda * (cols - 1) + (rows + 1) / 2 * 2;
          if (pool_stage.count(off) || (off & 1) || (lda & 1) || count > 8192 || bump + count > budget || off + count > H->pool_len) continue;
          pool_stage[off] = take((int)count);
          stage(pre, ST_POOL, off, pool_stage[off], (int)count, lda);
        }
      }
    for (const BushStage& e : pre) h.pre_bytes += e.count * 8;
    h.nst_pre = (int32_t)pre.size(); h.nst_post = (int32_t)post.size();
    bp.stages.insert(bp.stages.end(), pre.begin(), pre.end());
    bp.stages.insert(bp.stages.end(), post.begin(), post.end());
    bp.stage_bush.resize(bp.stages.size(), ob);
    bp.smem_doubles = std::max(bp.smem_doubles, bump);
    for (int l = 0; l < nlev; ++l) {
      const int chunk = chunk_of[(size_t)l];
      for (int32_t i : ts) {
        if (level[(size_t)i] != l) continue;
        const GTask& g = H->tasks_host[(size_t)(t0 + i)];
        for (int m0 = 0; m0 < g.M; m0 += chunk) {
          BushOp o;
          memset(&o, 0, sizeof(o));
          o.a0 = g.a0; o.a1 = g.a1; o.b0 = g.b0; o.b1 = g.b1; o.c = g.c;
          o.lda0 = g.lda0; o.lda1 = g.lda1; o.ldb0 = g.ldb0; o.ldb1 = g.ldb1; o.ldc = g.ldc;
          o.m0 = m0; o.mr = std::min(chunk, g.M - m0);
          o.K0 = g.K0 > 0 ? g.K0 : 0; o.K1 = g.K1 > 0 ? g.K1 : 0;
          o.ta0 = g.ta0; o.ta1 = g.ta1; o.sb0 = g.sb0; o.sb1 = g.sb1; o.sc = g.sc; o.epilogue = g.epilogue;
          o.s0 = o.s1 = o.sa0 = o.sa1 = -1;
          for (int s = 0; s < 2; ++s) {
            if ((s ? o.K1 : o.K0) <= 0) continue;
            const int32_t pt = prod[(size_t)(2 * i + s)];
            if (pt >= 0 && tbush[(size_t)pt] == b && slot[(size_t)pt] >= 0) {
              (s ? o.s1 : o.s0) = slot[(size_t)pt];
              (s ? o.lds1 : o.lds0) = slot_ld[(size_t)pt];
            } else if (pt >= 0 && tbush[(size_t)pt] != b) {
              auto ws = ws_stage.find(Key{s ? g.sb1 : g.sb0, s ? g.b1 : g.b0});
              if (ws != ws_stage.end()) { (s ? o.s1 : o.s0) = ws->second.first; (s ? o.lds1 : o.lds0) = ws->second.second; }
            }
            auto ps = pool_stage.find(s ? g.a1 : g.a0);
            if (ps != pool_stage.end()) (s ? o.sa1 : o.sa0) = ps->second;
          }
          o.sc_off = slot[(size_t)i]; o.ldsc = slot_ld[(size_t)i];
          // the workspace copy is skipped only when every reader sits in this bush and reads the shared-memory image
          o.to_global = !(used_in[(size_t)i] && !used_out[(size_t)i] && slot[(size_t)i] >= 0);
          bp.ops.push_back(o);
          bp.op_task.push_back(i); bp.op_bush.push_back(ob); bp.op_level.push_back(l);
        }
      }
      h.lvl_end[l] = (int32_t)bp.ops.size() - h.op0;
    }
    h.nops = (int32_t)bp.ops.size() - h.op0;
    bp.hdr.push_back(h);
  }
  bp.nbush = nb;
  bp.usable = nb > 0 && !bp.ops.empty();
  if (!bp.usable) bp.why = "no work";
}

// Small ranks only: a bush holds its generator blocks and Z / F tiles in shared memory, which pays for the rank-sized
// blocks of a compressed matrix; big uniform trees belong on the fixed-shape kernels, big ragged ones on hssb_flow.cuh.
static bool bush_eligible(const hssb_matrix* H) {
  return H->n_shards == 1 && H->max_leaf_m <= 64 && H->max_leaf_n <= 64 && H->max_rank <= 64 && (int64_t)H->leaves.size() <= 16384;
}

static int ensure_bush_plan(hssb_matrix* H, int mode, int64_t nrhs) {
  if (mode < 0 || mode > 1) return HSSB_OK;
  BushPlan* bp = (BushPlan*)H->bush_plan[mode];
  if (!bp) {
    std::unique_ptr<BushPlan> np(new (std::nothrow) BushPlan());
    if (!np) HSSB_FAIL(HSSB_ERR_ALLOC, "bush plan: out of memory");
    bush_plan_host(H, mode, H->bush_levels, H->bush_levels0, BUSH_SMEM_BUDGET, *np);
    if (np->usable) {
      int sms = 148;
      HSSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, H->device));
      np->grid_cap = sms * 2;
      HSSB_CUDA(cudaFuncSetAttribute(bush_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BUSH_SMEM_BUDGET * (int)sizeof(double)));
      HSSB_CUDA(cudaMalloc(&np->ops_dev, np->ops.size() * sizeof(BushOp)));
      HSSB_CUDA(cudaMalloc(&np->hdr_dev, np->hdr.size() * sizeof(BushHdr)));
      HSSB_CUDA(cudaMalloc(&np->deps_dev, std::max<size_t>(np->deps.size(), 1) * sizeof(int32_t)));
      HSSB_CUDA(cudaMemcpy(np->ops_dev, np->ops.data(), np->ops.size() * sizeof(BushOp), cudaMemcpyHostToDevice));
      HSSB_CUDA(cudaMemcpy(np->hdr_dev, np->hdr.data(), np->hdr.size() * sizeof(BushHdr), cudaMemcpyHostToDevice));
      HSSB_CUDA(cudaMalloc(&np->stages_dev, std::max<size_t>(np->stages.size(), 1) * sizeof(BushStage)));
      if (!np->stages.empty()) HSSB_CUDA(cudaMemcpy(np->stages_dev, np->stages.data(), np->stages.size() * sizeof(BushStage), cudaMemcpyHostToDevice));
      if (!np->deps.empty()) HSSB_CUDA(cudaMemcpy(np->deps_dev, np->deps.data(), np->deps.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    bp = np.release();
    H->bush_plan[mode] = bp;
  }
  if (!bp->usable) return HSSB_OK;
  const int64_t ncol = (nrhs + B_TN - 1) / B_TN;
  if (ncol > bp->sync_cols || (H->bush_trace && !bp->trace_dev)) {
    if (H->stream) HSSB_CUDA(cudaStreamSynchronize(H->stream));
    cudaFree(bp->sync_dev);
    bp->sync_dev = nullptr; bp->sync_cols = 0;
    if (bp->trace_dev) { cudaFree(bp->trace_dev); bp->trace_dev = nullptr; }
    if (H->bush_trace) {
      HSSB_CUDA(cudaMalloc(&bp->trace_dev, ((size_t)bp->nbush * ncol * B_TRACE + B_PROBE_LEVELS * B_WARPS * 8) * sizeof(unsigned long long)));
      HSSB_CUDA(cudaMemset(bp->trace_dev, 0, ((size_t)bp->nbush * ncol * B_TRACE + B_PROBE_LEVELS * B_WARPS * 8) * sizeof(unsigned long long)));
    }
    HSSB_CUDA(cudaMalloc(&bp->sync_dev, (size_t)(1 + (int64_t)bp->nbush * ncol) * sizeof(unsigned int)));
    bp->sync_cols = ncol;
    invalidate_graphs(H);
  }
  return HSSB_OK;
}

static bool bush_usable(const hssb_matrix* H, int mode) {
  if (mode < 0 || mode > 1 || !H->bush_kernel || H->profile) return false;
  if (H->bush_kernel == 1 && !bush_eligible(H)) return false;
  const BushPlan* bp = (const BushPlan*)H->bush_plan[mode];
  return bp && bp->usable && bp->sync_dev;
}

static int launch_bush(hssb_matrix* H, int mode, const CallParams& cp, cudaStream_t st) {
  const BushPlan* bp = (const BushPlan*)H->bush_plan[mode];
  const int64_t ncol = (cp.nrhs + B_TN - 1) / B_TN;
  if (ncol > bp->sync_cols) HSSB_FAIL(HSSB_ERR_STATE, "bush kernel: flags sized for %lld column tiles, call needs %lld", (long long)bp->sync_cols, (long long)ncol);
  HSSB_CUDA(cudaMemsetAsync(bp->sync_dev, 0, (size_t)(1 + (int64_t)bp->nbush * ncol) * sizeof(unsigned int), st));
  BushParams f;
  f.ops = bp->ops_dev; f.hdr = bp->hdr_dev; f.stages = bp->stages_dev; f.deps = bp->deps_dev;
  f.sync = bp->sync_dev; f.nbush = bp->nbush;
  f.trace = H->bush_trace && ncol == bp->sync_cols ? bp->trace_dev : nullptr;
  f.probe_item = H->bush_probe_item;
  const int grid = (int)std::min<int64_t>((int64_t)bp->nbush * ncol, bp->grid_cap);
  bush_kernel<<<grid, B_THREADS, (size_t)bp->smem_doubles * sizeof(double), st>>>(f, cp);
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

}  // namespace hssb
