// hssb_flow.cuh — the WHOLE product of an any-shape tree in ONE launch: a persistent dataflow kernel.
//
// Trees that come out of a real compression (README Cauchy example, BASELINE configs 1-2: ragged 62/63-row
// leaves, ranks 9-20 that differ from node to node) are small: config 2 is 1 GFLOP and 118 MB, config 1 a few
// MFLOP.  Their product is 2*depth + 2 DEPENDENT levels, and launched level by level (even replayed as a CUDA
// graph) it is pure launch / drain latency: 22 launches and 0.29 ms for config 2 against a roofline of 0.03 ms.
//
// Here the recursion of matmul.jl:32-62 runs as a task graph inside one kernel:
//   * the items -- (task, 64-row tile, 64-column tile) in the level order of the plan, which is a topological
//     order -- are handed out by one atomic counter to a grid of resident CTAs (2 per SM);
//   * every task knows the producer of each workspace operand (the merge / translate / leaf-up task that writes
//     that Z or F block); a CTA that draws an item waits until the producer's row tiles of ITS column tile are
//     done (one counter per (task, column tile), release / acquire at gpu scope), computes the tile with the
//     any-shape DMMA code of hssb_kernels_generic.cuh and bumps its own counter;
//   * a waiting CTA only ever waits for items with a smaller index, which were drawn earlier by CTAs that are
//     running: the first unfinished item can always proceed, so the kernel cannot deadlock and needs no
//     cooperative launch.
// Independent subtrees overlap freely (no level barrier); the dependent chain costs one flag round trip
// (~1 us) per level instead of a launch.  Workspace operands are read with ld.global.cg: their producer ran in
// the same launch on another SM.
#pragma once

#include "hssb_kernels_generic.cuh"

namespace hssb {

struct FlowParams {
  const GTask* tasks;          // task table of the plan (phase order = topological order)
  const int32_t* deps;         // [2 * ntasks]: producer task of operand B0 / B1, -1 if none (X, or absent)
  const int32_t* q_task;       // [nq]: task of the q-th (task, row tile)
  const int32_t* q_mtile;      // [nq]: its row tile
  unsigned int* sync;          // [0] = next item, [1 + task * ncol + coltile] = finished row tiles
  int32_t nq;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The dependent chain of the tree is a chain of flag round trips, so everything that does NOT depend on the
// producer is taken off it: the index of the next item is drawn while the current one is computed, and the
// generator slabs (A operands: pool data, always ready) are fetched BEFORE the CTA waits for its producers; only the
// workspace operand travels after the flag.  Up to F_PF K-slabs are in flight at once (a merge / translate of
// rank <= 32 is 2-4 slabs: one round trip instead of one per slab).
constexpr int F_PF = 4;

__global__ void __launch_bounds__(G_THREADS, 2)
flow_kernel(FlowParams f, CallParams p) {
  __shared__ double As[G_SMEM];
  __shared__ double Bs[G_SMEM];
  __shared__ int s_idx[2];
  const int tid = threadIdx.x;
  const int N = p.nrhs;
  const int ncol = (N + G_TN - 1) / G_TN;
  const long long total = (long long)f.nq * ncol;
  unsigned int* done = f.sync + 1;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;

  if (tid == 0) s_idx[0] = (int)atomicAdd(f.sync, 1u);
  __syncthreads();
  for (int it = 0;; ++it) {
    const int idx = s_idx[it & 1];
    if (idx >= total) break;
    if (tid == 0) s_idx[(it + 1) & 1] = (int)atomicAdd(f.sync, 1u);  // consumed after this item's barriers
    const int q = idx / ncol, j = idx - q * ncol;
    const int ti = f.q_task[q];
    const GTask t = f.tasks[ti];
    const int m0 = f.q_mtile[q] * G_TM, n0 = j * G_TN;
    const int K0 = t.K0 > 0 ? t.K0 : 0, K1 = t.K1 > 0 ? t.K1 : 0;
    const int nslab0 = (K0 + G_TK - 1) / G_TK, nslab = nslab0 + (K1 + G_TK - 1) / G_TK;
    int64_t ldb0 = 0, ldb1 = 0;
    const double* B0 = operand_b(p, t.sb0, t.b0, t.ldb0, ldb0);
    const double* B1 = operand_b(p, t.sb1, t.b1, t.ldb1, ldb1);
    double ra[F_PF][4], rb[F_PF][4];
    auto fetch_a = [&](int slab, double (&r)[4]) {
      const bool s1 = slab >= nslab0;
      const int K = s1 ? K1 : K0, k0 = (s1 ? slab - nslab0 : slab) * G_TK;
      const double* A = p.pool + (s1 ? t.a1 : t.a0);
      const int64_t lda = s1 ? t.lda1 : t.lda0;
      if (!(s1 ? t.ta1 : t.ta0)) {
        const int mm = tid & 63;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int kk = (tid >> 6) + 4 * r4;
          r[r4] = (m0 + mm < t.M && k0 + kk < K) ? A[(int64_t)(k0 + kk) * lda + (m0 + mm)] : 0.0;
        }
      } else {
        const int kk = tid & 15;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const int mm = (tid >> 4) + 16 * r4;
          r[r4] = (m0 + mm < t.M && k0 + kk < K) ? A[(int64_t)(m0 + mm) * lda + (k0 + kk)] : 0.0;
        }
      }
    };
    auto fetch_b = [&](int slab, double (&r)[4]) {
      const bool s1 = slab >= nslab0;
      const int K = s1 ? K1 : K0, k0 = (s1 ? slab - nslab0 : slab) * G_TK;
      const double* B = s1 ? B1 : B0;
      const int64_t ldb = s1 ? ldb1 : ldb0;
      const bool ws = (s1 ? t.sb1 : t.sb0) != SRC_X;  // produced in this launch on another SM: bypass L1
      const int kk = tid & 15;
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const int nn = (tid >> 4) + 16 * r4;
        const bool in = n0 + nn < N && k0 + kk < K;
        const double* src = B + (int64_t)(n0 + nn) * ldb + (k0 + kk);
        r[r4] = in ? (ws ? __ldcg(src) : *src) : 0.0;
      }
    };
    // generators first: they do not depend on anybody
#pragma unroll
    for (int u = 0; u < F_PF; ++u)
      if (u < nslab) fetch_a(u, ra[u]);
    if (tid < 2) {  // one thread per workspace operand waits for its producer
      const int d = f.deps[2 * ti + tid];
      if (d >= 0) {
        const unsigned int need = (unsigned int)((f.tasks[d].M + G_TM - 1) / G_TM);
        const unsigned int* flag = done + (size_t)d * ncol + j;
        if (ld_acquire_u32(flag) < need) {
          const long long t0 = clock64();
          while (ld_acquire_u32(flag) < need)
            if (clock64() - t0 > 8000000000ll) trap_report(TRAP_FLOW, (unsigned long long)d, (unsigned long long)ti);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < F_PF; ++u)
      if (u < nslab) fetch_b(u, rb[u]);

    const bool active = m0 + wm < t.M && n0 + wn < N;
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
    for (int base = 0; base < nslab; base += F_PF) {
#pragma unroll
      for (int u = 0; u < F_PF; ++u) {
        const int slab = base + u;
        if (slab < nslab) {
          const bool ta = slab >= nslab0 ? t.ta1 : t.ta0;
          if (!ta) {
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) As[((tid >> 6) + 4 * r4) * G_SA + (tid & 63)] = ra[u][r4];
          } else {
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) As[((tid >> 4) + 16 * r4) * G_SB + (tid & 15)] = ra[u][r4];
          }
#pragma unroll
          for (int r4 = 0; r4 < 4; ++r4) Bs[((tid >> 4) + 16 * r4) * G_SB + (tid & 15)] = rb[u][r4];
          __syncthreads();
          if (slab + F_PF < nslab) { fetch_a(slab + F_PF, ra[u]); fetch_b(slab + F_PF, rb[u]); }
          const double* ap = ta ? As + (wm + g) * G_SB + q4 : As + q4 * G_SA + wm + g;
          const int a_tile = ta ? 8 * G_SB : 8, a_step = ta ? 4 : 4 * G_SA;
          const double* bp = Bs + (wn + g) * G_SB + q4;
          // k-steps past the end of the operand are zero fill: skipping them changes nothing in the result (x + 0 * 0)
          // and takes them off the chain of DEPENDENT DMMAs, ~500 cycles each for a warp that has nothing else in
          // flight (profiles/bush_kernel_r02.txt) -- a rank-17 operand is 5 k-steps, not 8
          const int ks_end = ((slab >= nslab0 ? K1 - (slab - nslab0) * G_TK : K0 - slab * G_TK) + 3) >> 2;
          if (active) {
#pragma unroll
            for (int ks = 0; ks < G_TK / 4; ++ks) {
              if (ks >= ks_end) break;
              double a[4], b[2];
#pragma unroll
              for (int i = 0; i < 4; ++i) a[i] = ap[ks * a_step + i * a_tile];
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) b[jj] = bp[jj * 8 * G_SB + ks * 4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) mma_m8n8k4(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
            }
          }
          __syncthreads();
        }
      }
    }
    // ---- epilogue (same element mapping as generic_tile)
    int64_t ldc;
    double* C;
    switch (t.sc) {
      case SRC_Z: ldc = t.ldc; C = p.Z + t.c * (int64_t)N; break;
      case SRC_F: ldc = t.ldc; C = p.F + t.c * (int64_t)N; break;
      default: ldc = p.ldy; C = p.Y + t.c; break;
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n0 + wn + 8 * jj + 2 * q4 + e;
        if (col >= N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = m0 + wm + 8 * i + g;
          if (row >= t.M) continue;
          double* dst = C + (int64_t)col * ldc + row;
          double v = acc[i][jj][e];
          if (t.epilogue) {
            v *= p.alpha;
            if (p.beta != 0.0) v += p.beta * (*dst);  // beta == 0 never reads Y (matmul.jl:13)
          }
          *dst = v;
        }
      }
    __syncthreads();  // every thread's stores precede thread 0's release; s_idx[(it + 1) & 1] is visible
    if (tid == 0 && (t.sc == SRC_Z || t.sc == SRC_F)) {
      __threadfence();
      atomicAdd(done + (size_t)ti * ncol + j, 1u);
    }
  }
}

// ================================================================ host side ===
struct FlowPlan {
  bool usable = false;
  std::string why;
  int32_t nq = 0;
  int64_t ntasks = 0, task0 = 0;
  int32_t* deps_dev = nullptr;
  int32_t* q_task_dev = nullptr;
  int32_t* q_mtile_dev = nullptr;
  unsigned int* sync_dev = nullptr;
  int64_t sync_cols = 0;  // column tiles the sync block is sized for
  int grid_cap = 148 * 2;
};

static void free_flow(hssb_matrix* H) {
  for (void*& v : H->flow_plan) {
    FlowPlan* fp = (FlowPlan*)v;
    if (!fp) continue;
    cudaFree(fp->deps_dev); cudaFree(fp->q_task_dev); cudaFree(fp->q_mtile_dev); cudaFree(fp->sync_dev);
    delete fp;
    v = nullptr;
  }
}

// Producers of every workspace operand of the plan `mode` (0: Y = A X, 1: Y = A' X on the any-shape task table).
// The plan qualifies if its phases are one contiguous run of tasks and every Z / F operand is exactly the
// output block of an EARLIER task.
static void flow_plan_host(const hssb_matrix* H, int mode, FlowPlan& fp, std::vector<int32_t>& deps, std::vector<int32_t>& q_task,
                           std::vector<int32_t>& q_mtile) {
  const std::vector<Phase>& phases = mode == 1 ? H->phases_t : H->phases;
  fp.usable = false;
  if (H->n_shards != 1) { fp.why = "sharded handle (the exchange sits between the levels)"; return; }
  if (phases.empty()) { fp.why = "empty plan"; return; }
  int64_t t0 = -1, t1 = -1;
  for (const Phase& ph : phases) {
    if (ph.kind == PH_EXCHANGE || ph.kind == PH_XCHG_ACK) { fp.why = "plan contains an exchange"; return; }
    if (ph.ntasks == 0) continue;
    if (t0 < 0) t0 = ph.task0;
    else if (ph.task0 != t1) { fp.why = "phases are not contiguous in the task table"; return; }
    t1 = ph.task0 + ph.ntasks;
  }
  if (t0 < 0 || t1 - t0 > INT32_MAX / 4) { fp.why = "no tasks"; return; }
  fp.task0 = t0; fp.ntasks = t1 - t0;
  // output block -> producing task
  struct Key { int src; int64_t row; bool operator<(const Key& o) const { return src != o.src ? src < o.src : row < o.row; } };
  std::map<Key, int32_t> producer;
  deps.assign((size_t)(2 * fp.ntasks), -1);
  q_task.clear(); q_mtile.clear();
  for (int64_t i = 0; i < fp.ntasks; ++i) {
    const GTask& g = H->tasks_host[(size_t)(t0 + i)];
    for (int o = 0; o < 2; ++o) {
      const int K = o ? g.K1 : g.K0, src = o ? g.sb1 : g.sb0;
      const int64_t row = o ? g.b1 : g.b0;
      if (K <= 0 || (src != SRC_Z && src != SRC_F)) continue;
      auto it = producer.find(Key{src, row});
      if (it == producer.end()) { fp.why = "a workspace operand is not the output block of an earlier task"; return; }
      deps[(size_t)(2 * i + o)] = it->second;
    }
    if (g.sc == SRC_Z || g.sc == SRC_F) {
      if (!producer.emplace(Key{g.sc, g.c}, (int32_t)i).second) { fp.why = "a workspace block is written twice"; return; }
    }
    for (int mt = 0; mt * G_TM < std::max(g.M, 1); ++mt) {
      if (g.M <= 0) break;
      q_task.push_back((int32_t)i);
      q_mtile.push_back(mt);
    }
  }
  fp.nq = (int32_t)q_task.size();
  fp.usable = fp.nq > 0;
  if (!fp.usable) fp.why = "no work";
}

static int ensure_flow_plan(hssb_matrix* H, int mode, int64_t nrhs) {
  if (mode < 0 || mode > 1) return HSSB_OK;
  FlowPlan* fp = (FlowPlan*)H->flow_plan[mode];
  if (!fp) {
    std::unique_ptr<FlowPlan> np(new (std::nothrow) FlowPlan());
    if (!np) HSSB_FAIL(HSSB_ERR_ALLOC, "flow plan: out of memory");
    std::vector<int32_t> deps, q_task, q_mtile;
    flow_plan_host(H, mode, *np, deps, q_task, q_mtile);
    if (np->usable) {
      int sms = 148;
      HSSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, H->device));
      np->grid_cap = sms * 2;
      HSSB_CUDA(cudaMalloc(&np->deps_dev, deps.size() * sizeof(int32_t)));
      HSSB_CUDA(cudaMalloc(&np->q_task_dev, q_task.size() * sizeof(int32_t)));
      HSSB_CUDA(cudaMalloc(&np->q_mtile_dev, q_mtile.size() * sizeof(int32_t)));
      HSSB_CUDA(cudaMemcpy(np->deps_dev, deps.data(), deps.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      HSSB_CUDA(cudaMemcpy(np->q_task_dev, q_task.data(), q_task.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      HSSB_CUDA(cudaMemcpy(np->q_mtile_dev, q_mtile.data(), q_mtile.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    fp = np.release();
    H->flow_plan[mode] = fp;
  }
  if (!fp->usable) return HSSB_OK;
  const int64_t ncol = (nrhs + G_TN - 1) / G_TN;
  if (ncol > fp->sync_cols) {
    if (H->stream) HSSB_CUDA(cudaStreamSynchronize(H->stream));
    cudaFree(fp->sync_dev);
    fp->sync_dev = nullptr; fp->sync_cols = 0;
    HSSB_CUDA(cudaMalloc(&fp->sync_dev, (size_t)(1 + fp->ntasks * ncol) * sizeof(unsigned int)));
    fp->sync_cols = ncol;
    invalidate_graphs(H);
  }
  return HSSB_OK;
}

// The whole plan `mode` in one launch (plus the memset of its counters); false if the plan does not qualify.
static bool flow_usable(const hssb_matrix* H, int mode) {
  if (mode < 0 || mode > 1 || !H->flow_kernel || H->profile) return false;
  const FlowPlan* fp = (const FlowPlan*)H->flow_plan[mode];
  return fp && fp->usable && fp->sync_dev;
}

static int launch_flow(hssb_matrix* H, int mode, const CallParams& cp, cudaStream_t st) {
  const FlowPlan* fp = (const FlowPlan*)H->flow_plan[mode];
  const int64_t ncol = (cp.nrhs + G_TN - 1) / G_TN;
  if (ncol > fp->sync_cols) HSSB_FAIL(HSSB_ERR_STATE, "flow kernel: counters sized for %lld column tiles, call needs %lld", (long long)fp->sync_cols, (long long)ncol);
  HSSB_CUDA(cudaMemsetAsync(fp->sync_dev, 0, (size_t)(1 + fp->ntasks * ncol) * sizeof(unsigned int), st));
  FlowParams f;
  f.tasks = H->tasks_dev + fp->task0;
  f.deps = fp->deps_dev; f.q_task = fp->q_task_dev; f.q_mtile = fp->q_mtile_dev;
  f.sync = fp->sync_dev; f.nq = fp->nq;
  const int grid = (int)std::min<int64_t>((int64_t)fp->nq * ncol, fp->grid_cap);
  flow_kernel<<<grid, G_THREADS, 0, st>>>(f, cp);
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

}  // namespace hssb
