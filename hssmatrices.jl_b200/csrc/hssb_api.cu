// hssb_api.cu — C ABI, packer, level scheduler of the B200-native HSS x dense
// product (reference path: src/matmul.jl:13-62 of bonevbs/HssMatrices.jl).
//
// Data flow:  builder / synthetic description  ->  BFS node table
//             -> level-ordered generator pool in HBM (one allocation)
//             -> task table (one GTask per small GEMM of the recursion)
//             -> phases (leaf-up, merges by height, [exchange], translates by
//                depth, leaf-down) launched back to back on one stream.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <new>

#include "hssb_internal.h"
#include "hssb_kernels_generic.cuh"
#include "hssb_synth.cuh"
#include "hssb_fast.cuh"
#include "hssb_ulv.cuh"

namespace hssb {

static thread_local char g_err[1024] = "";

// Host-mapped words the kernels' bounded waits write to before they trap (hssb_fast.cuh: trap_report).
static unsigned long long* g_trap_host = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (g_trap_host && g_trap_host[0]) {  // a kernel gave up waiting: say for what
    static const char* what[] = {"?", "a shared-memory mbarrier (TMA data or a ring slot)", "the grid barrier of the tree kernel",
                                 "a peer's acknowledgement of the previous exchange", "a peer's subtree-root block (exchange)",
                                 "the producer task of an operand (dataflow kernel)"};
    const unsigned long long c = g_trap_host[0];
    const size_t len = strlen(g_err);
    snprintf(g_err + len, sizeof(g_err) - len, " [a kernel timed out waiting for %s: wanted %llu, saw %llu, block %llu thread %llu]",
             what[c < 6 ? c : 0], g_trap_host[1], g_trap_host[2], g_trap_host[3] & 0xffffffffull, g_trap_host[3] >> 32);
  }
}

static inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// One host-mapped block per process; every device's copy of g_trap_slot points at it.
static int ensure_trap_slot(int device) {
  static std::mutex mu;
  static bool done[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64 || done[device]) return HSSB_OK;
  DeviceGuard dg(device);
  if (!g_trap_host) {
    HSSB_CUDA(cudaHostAlloc((void**)&g_trap_host, 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(g_trap_host, 0, 64);
  }
  unsigned long long* dptr = nullptr;
  HSSB_CUDA(cudaHostGetDevicePointer((void**)&dptr, g_trap_host, 0));
  HSSB_CUDA(cudaMemcpyToSymbol(g_trap_slot, &dptr, sizeof(dptr)));
  done[device] = true;
  return HSSB_OK;
}

static int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    HSSB_FAIL(HSSB_ERR_CUDA, "no CUDA device available (%s); hssb200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n) HSSB_FAIL(HSSB_ERR_ARG, "device %d out of range (have %d)", device, n);
  cudaDeviceProp prop;
  HSSB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    HSSB_FAIL(HSSB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
  return ensure_trap_slot(device);
}

// ----------------------------------------------------------------- builder ---
struct BNode {
  bool leaf = false, remote = false, used = false;
  int64_t left = -1, right = -1;
  int64_t m = 0, n = 0, kr = 0, kw = 0;
  HostBlock blk[BK_COUNT];  // D,U,V (leaf) / B12,B21 (branch); R,W are filled by the parent
};

}  // namespace hssb

struct hssb_builder {
  std::vector<hssb::BNode> nodes;
};

namespace hssb {

static int copy_block(HostBlock& dst, const double* src, int64_t ld, int64_t rows, int64_t cols, const char* what) {
  dst.rows = rows;
  dst.cols = cols;
  dst.data.clear();
  if (rows == 0 || cols == 0) return HSSB_OK;
  if (!src) HSSB_FAIL(HSSB_ERR_ARG, "%s: NULL pointer for a %lld x %lld block", what, (long long)rows, (long long)cols);
  if (ld < rows) HSSB_FAIL(HSSB_ERR_DIM, "%s: leading dimension %lld < rows %lld", what, (long long)ld, (long long)rows);
  try {
    dst.data.resize((size_t)(rows * cols));
  } catch (const std::bad_alloc&) {
    HSSB_FAIL(HSSB_ERR_ALLOC, "%s: host allocation of %lld doubles failed", what, (long long)(rows * cols));
  }
  for (int64_t j = 0; j < cols; ++j) memcpy(dst.data.data() + j * rows, src + j * ld, (size_t)rows * sizeof(double));
  return HSSB_OK;
}

static void invalidate_graphs(hssb_matrix* H);
static int run_graph(hssb_matrix* H, const CallParams& cp, cudaStream_t st);

}  // namespace hssb
#include "hssb_plan.cuh"
#include "hssb_twin.cuh"
namespace hssb {

static void nodes_from_builder(const hssb_builder* b, int64_t root, hssb_matrix* H, std::vector<BlockSource>& src) {
  // BFS renumbering.  A builder that holds the WHOLE tree may be finalised once per shard: subtrees at the
  // shard cut (depth log2 n_shards) that belong to other shards are turned into size-only placeholders
  // here, exactly as if the caller had registered them with hssb_builder_add_remote.
  int p = 0;
  while ((1 << p) < H->n_shards) ++p;
  struct Item { int64_t id; int depth; int64_t pos; bool pruned; };
  std::vector<Item> order{{root, 0, 0, false}};
  for (size_t q = 0; q < order.size(); ++q) {
    const Item it = order[q];
    const BNode& bn = b->nodes[(size_t)it.id];
    if (H->n_shards > 1 && it.depth == p && it.pos != H->shard_rank && !bn.remote) { order[q].pruned = true; continue; }
    if (!bn.leaf && !bn.remote) {
      order.push_back({bn.left, it.depth + 1, 2 * it.pos, false});
      order.push_back({bn.right, it.depth + 1, 2 * it.pos + 1, false});
    }
  }
  std::vector<int64_t> newid(b->nodes.size(), -1);
  for (size_t q = 0; q < order.size(); ++q) newid[(size_t)order[q].id] = (int64_t)q;
  H->nodes.resize(order.size());
  src.resize(order.size());
  for (size_t q = 0; q < order.size(); ++q) {
    const BNode& bn = b->nodes[(size_t)order[q].id];
    Node& t = H->nodes[q];
    t.leaf = bn.leaf && !order[q].pruned; t.remote = bn.remote || order[q].pruned;
    t.m = bn.m; t.n = bn.n; t.kr = bn.kr; t.kw = bn.kw;
    if (!bn.leaf && !t.remote) { t.left = newid[(size_t)bn.left]; t.right = newid[(size_t)bn.right]; }
    for (int k = 0; k < BK_COUNT; ++k) src[q].blk[k] = &bn.blk[k];
  }
}

// BFS construction of the bisection tree (clustertree.jl:27-35): split while len > leafsize,
// left child gets ceil(len/2).  Subtrees at the shard cut owned by other ranks become placeholders.
static void nodes_synthetic(hssb_matrix* H, int64_t n, int64_t leafsize, int64_t rank) {
  int p = 0;
  while ((1 << p) < H->n_shards) ++p;
  struct Item { int64_t len; uint64_t heap; int depth; int64_t pos; };
  std::vector<Item> items{{n, 1, 0, 0}};
  H->nodes.resize(1);
  for (size_t q = 0; q < items.size(); ++q) {
    const Item it = items[q];
    Node& t = H->nodes[q];
    t.m = t.n = it.len;
    t.kr = t.kw = (q == 0) ? 0 : rank;
    t.heap_id = it.heap;
    const bool remote = (it.depth == p && H->n_shards > 1 && it.pos != H->shard_rank);
    if (remote) { t.remote = true; continue; }
    if (it.len > leafsize) {
      const int64_t nl = (it.len + 1) / 2;
      t.leaf = false;
      t.left = (int64_t)items.size();
      t.right = t.left + 1;
      items.push_back({nl, 2 * it.heap, it.depth + 1, 2 * it.pos});
      items.push_back({it.len - nl, 2 * it.heap + 1, it.depth + 1, 2 * it.pos + 1});
      H->nodes.resize(items.size());
    } else {
      H->nodes[q].leaf = true;
    }
  }
}

// Host image of the packed pool (plan-only handles used by the CPU tests).
static void fill_pool_host(hssb_matrix* H, const std::vector<BlockSource>* src) {
  H->pool_host.assign((size_t)H->pool_len, 0.0);
  const double tscale = H->synth_rank > 0 ? 1.0 / sqrt(2.0 * (double)H->synth_rank) : 1.0;
  for (size_t i = 0; i < H->nodes.size(); ++i) {
    const Node& t = H->nodes[i];
    for (int k = 0; k < BK_COUNT; ++k) {
      if (t.off[k] < 0) continue;
      double* dst = H->pool_host.data() + t.off[k];
      const bool tr = stored_transposed(k);  // stored(i, j) = logical(j, i); logical block is cols x rows
      if (src) {
        const HostBlock* hb = (*src)[i].blk[k];
        for (int64_t j = 0; j < t.cols[k]; ++j)
          for (int64_t r = 0; r < t.rows[k]; ++r)
            dst[j * t.ld[k] + r] = tr ? hb->data[(size_t)(r * t.cols[k] + j)] : hb->data[(size_t)(j * t.rows[k] + r)];
      } else {
        const uint64_t key = synth_key(H->seed, t.heap_id, k);
        const double c = IH4_SCALE * ((k == BK_R || k == BK_W) ? tscale : 1.0);
        for (int64_t j = 0; j < t.cols[k]; ++j)
          for (int64_t r = 0; r < t.rows[k]; ++r)
            dst[j * t.ld[k] + r] = synth_value(key, (uint64_t)(tr ? r * t.cols[k] + j : j * t.rows[k] + r), c);
      }
    }
  }
}

}  // namespace hssb
#include "hssb_xchg.cuh"
#include "hssb_tree.cuh"
#include "hssb_flow.cuh"
#include "hssb_bush.cuh"
#include "hssb_hostpipe.h"
namespace hssb {

// ------------------------------------------------------------------ launch ---
static int launch_generic(hssb_matrix* H, const Phase& ph, const CallParams& cp, cudaStream_t st) {
  if (ph.ntasks == 0) return HSSB_OK;
  dim3 grid((unsigned)ph.ntasks, (unsigned)((ph.maxM + G_TM - 1) / G_TM), (unsigned)((cp.nrhs + G_TN - 1) / G_TN));
  HSSB_CUDA(launch_k((H->pdl & 9) != 0, generic_level_kernel, grid, dim3(G_THREADS), 0, st, (const GTask*)(H->tasks_dev + ph.task0), cp));
  H->launches++;
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
}

// 0: Y = A X, 1: Y = A' X (any-shape transposed task table), 2: ULV solve
static const std::vector<Phase>& phase_list(const hssb_matrix* H, int mode) {
  return mode == 2 ? H->phases_u : mode == 1 ? H->phases_t : H->phases;
}

static bool plan_runs_generic(const hssb_matrix* H, const CallParams& cp) {
  for (const Phase& ph : phase_list(H, cp.trans)) {
    if (ph.kind == PH_EXCHANGE || ph.kind == PH_XCHG_ACK || ph.ntasks == 0) continue;
    if (ph.fast && !H->force_generic && fast_phase_supported(H, ph, cp)) return false;
  }
  return true;
}

static int run_phases(hssb_matrix* H, const CallParams& cp, cudaStream_t st) {
  const std::vector<Phase>& phases = phase_list(H, cp.trans);
  const bool prof = H->profile;
  // any-shape plans (no fixed-shape kernel applies): the whole product as one dataflow launch (hssb_flow.cuh)
  // ... small trees as one launch over whole bushes of the tree (hssb_bush.cuh)
  if (bush_usable(H, cp.trans) && plan_runs_generic(H, cp)) {  // leaf-up launch, every level in between as one launch, leaf-down launch
    bool tree_done = false;
    for (const Phase& ph : phases) {
      int rc = HSSB_OK;
      if (ph.kind == PH_LEAF_UP || ph.kind == PH_LEAF_DOWN) rc = launch_generic(H, ph, cp, st);
      else if (!tree_done) { rc = launch_bush(H, cp.trans, cp, st); tree_done = true; }
      if (rc) return rc;
    }
    return HSSB_OK;
  }
  // (automatic, HSSB_OPT_FLOW_KERNEL = 2: only for plain launches -- replayed as a CUDA graph, one launch per level with
  // programmatic dependent launch is faster: config-2 shape 170 us against 194 us, profiles/pdl_r02.txt)
  if (flow_usable(H, cp.trans) && (H->flow_kernel == 1 || !H->capturing) && plan_runs_generic(H, cp)) return launch_flow(H, cp.trans, cp, st);
  if (prof) {
    H->prof_mode = cp.trans;
    while (H->prof_events.size() < phases.size() + 1) {
      cudaEvent_t e;
      HSSB_CUDA(cudaEventCreate(&e));
      H->prof_events.push_back(e);
    }
    HSSB_CUDA(cudaEventRecord(H->prof_events[0], st));
    H->prof_nrhs = cp.nrhs;
  }
  // every level between the two leaf kernels in one persistent launch (hssb_tree.cuh)
  const TreePlan* tp = nullptr;
  if (cp.trans == 0 && H->tree_kernel && !H->force_generic) {
    tp = (const TreePlan*)H->tree_plan;  // built by matmul_dev_impl (outside graph capture)
    if (tp && tp->steps.empty()) tp = nullptr;
  }
  size_t pi = 0;
  for (const Phase& ph : phases) {
    ++pi;
    if (tp && (int)pi - 1 >= tp->phase0 && (int)pi - 1 < tp->phase1) {
      if ((int)pi - 1 > tp->phase0) continue;  // launched with the first covered phase
      const int ns = (int)tp->steps.size();
      XchgParams xq;
      memset(&xq, 0, sizeof(xq));
      int rc = HSSB_OK;
      if (tp->xchg_step < 0) {
        rc = launch_tree(H, tp, 0, ns, cp, xq, st);
      } else if (H->peer_xchg) {
        rc = launch_tree(H, tp, 0, ns, cp, xchg_params(H, cp), st);  // exchange + acknowledgement inside the kernel
      } else {
        // NCCL all-gather between two launches; the acknowledgement step is a no-op (no peer flags)
        if (!H->nccl_comm) HSSB_FAIL(HSSB_ERR_STATE, "sharded matrix: call hssb_comm_init (or hssb_xchg_import) before hssb_matmul");
        rc = launch_tree(H, tp, 0, tp->xchg_step, cp, xq, st);
        if (rc) return rc;
        double* buf = cp.Z + H->xchg_zoff * (int64_t)cp.nrhs;
        const size_t count = (size_t)H->xchg_slot_rows * (size_t)cp.nrhs;
        HSSB_NCCL(g_nccl.AllGather(buf + (size_t)H->shard_rank * count, buf, count, /*ncclDouble*/ 8, H->nccl_comm, st));
        rc = launch_tree(H, tp, tp->xchg_step + 1, ns, cp, xq, st);
      }
      if (rc) return rc;
      if (prof)  // the covered phases share one launch: its time is reported on the first of them
        for (int q = tp->phase0; q < tp->phase1; ++q) cudaEventRecord(H->prof_events[(size_t)q + 1], st);
      continue;
    }
    struct Rec {
      hssb_matrix* H; cudaStream_t st; size_t i; bool on;
      ~Rec() { if (on) cudaEventRecord(H->prof_events[i], st); }
    } rec{H, st, pi, prof};
    if (ph.kind == PH_XCHG_ACK) {
      if (H->peer_xchg) {
        xchg_ack_kernel<<<1, 32, 0, st>>>(xchg_params(H, cp));
        H->launches++;
        HSSB_CUDA(cudaGetLastError());
      }
      continue;
    }
    if (ph.kind == PH_EXCHANGE && H->peer_xchg) {
      const XchgParams q = xchg_params(H, cp);
      const int grid = (int)std::max<long long>(1, std::min<long long>(8, q.slot_elems / 2 / 256));
      xchg_push_kernel<<<grid, 256, 0, st>>>(q);
      H->launches++;
      HSSB_CUDA(cudaGetLastError());
      continue;
    }
    if (ph.kind == PH_EXCHANGE) {
      if (!H->nccl_comm) HSSB_FAIL(HSSB_ERR_STATE, "sharded matrix: call hssb_comm_init (or hssb_xchg_import) before hssb_matmul");
      double* buf = cp.Z + H->xchg_zoff * (int64_t)cp.nrhs;
      const size_t count = (size_t)H->xchg_slot_rows * (size_t)cp.nrhs;
      HSSB_NCCL(g_nccl.AllGather(buf + (size_t)H->shard_rank * count, buf, count, /*ncclDouble*/ 8, H->nccl_comm, st));
      continue;
    }
    int rc;
    if (ph.fast && !H->force_generic && fast_phase_supported(H, ph, cp))
      rc = launch_fast(H, ph, cp, st);
    else
      rc = launch_generic(H, ph, cp, st);
    if (rc) return rc;
  }
  return HSSB_OK;
}

// -------------------------------------------------------------- CUDA graphs ---
// The level schedule is 2*depth+2 dependent launches of a few microseconds
// each; replaying it as one graph removes the per-launch CPU cost.
static void invalidate_graphs(hssb_matrix* H) {
  for (auto& g : H->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  H->graphs.clear();
}

static bool same_call(const CallParams& a, const CallParams& b) {
  return a.pool == b.pool && a.X == b.X && a.Y == b.Y && a.Z == b.Z && a.F == b.F && a.ldx == b.ldx && a.ldy == b.ldy &&
         a.nrhs == b.nrhs && a.alpha == b.alpha && a.beta == b.beta && a.debug == b.debug && a.trans == b.trans;
}

static int run_graph(hssb_matrix* H, const CallParams& cp, cudaStream_t st) {
  for (auto& g : H->graphs)
    if (same_call(g.cp, cp)) {
      H->graph_miss_streak = 0;
      if (H->prepare_only) return HSSB_OK;
      HSSB_CUDA(cudaGraphLaunch(g.exec, st));
      H->launches += g.kernels;
      return HSSB_OK;
    }
  // A caller that hands over fresh device arrays on every call would pay a capture + instantiate (milliseconds) per
  // product: after two misses in a row the schedule is launched plainly, and a signature is only captured again
  // once it repeats (the host entry's staging pointers are stable and always captured).
  if (!H->in_host_call && !H->prepare_only) {
    const bool repeat = H->graph_plain_valid && same_call(H->graph_plain_cp, cp);
    if (++H->graph_miss_streak > 2 && !repeat) {
      H->graph_plain_cp = cp;
      H->graph_plain_valid = true;
      return run_phases(H, cp, st);
    }
  }
  if (H->graphs.size() >= 64) invalidate_graphs(H);
  cudaGraph_t graph = nullptr;
  const int64_t before = H->launches;
  // capture on the library's own stream (the legacy default stream cannot be captured),
  // replay on the caller's stream
  HSSB_CUDA(cudaStreamBeginCapture(H->stream, cudaStreamCaptureModeThreadLocal));
  H->capturing = true;
  int rc = run_phases(H, cp, H->stream);
  H->capturing = false;
  cudaError_t e = cudaStreamEndCapture(H->stream, &graph);
  const int64_t kernels = H->launches - before;
  H->launches = before;
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) { cudaGetLastError(); HSSB_FAIL(HSSB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e)); }
  hssb_matrix::GraphSlot slot;
  slot.cp = cp;
  slot.kernels = kernels;
  e = cudaGraphInstantiate(&slot.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); HSSB_FAIL(HSSB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
  H->graphs.push_back(slot);
  if (H->prepare_only) {  // upload now, so that the first replay has nothing left to set up
    HSSB_CUDA(cudaGraphUpload(slot.exec, st));
    return HSSB_OK;
  }
  HSSB_CUDA(cudaGraphLaunch(slot.exec, st));
  H->launches += kernels;
  return HSSB_OK;
}

}  // namespace hssb

using namespace hssb;

// =============================================================== C ABI =====
// No exception may cross the C ABI (std::vector growth, std::string, std::thread can throw).
template <class R, class F>
static R guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    set_error("out of host memory");
    return (R)HSSB_ERR_ALLOC;
  } catch (const std::exception& e) {
    set_error("internal error: %s", e.what());
    return (R)HSSB_ERR_STATE;
  }
}

extern "C" {

int hssb_version(void) { return HSSB_VERSION; }
const char* hssb_last_error(void) { return g_err; }

int hssb_device_count(void) {
  return guarded<int>([&]() -> int {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
  });
}

int hssb_builder_create(hssb_builder** out) {
  return guarded<int>([&]() -> int {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_builder_create: out is NULL");
  *out = new (std::nothrow) hssb_builder();
  if (!*out) HSSB_FAIL(HSSB_ERR_ALLOC, "hssb_builder_create: out of memory");
  return HSSB_OK;
  });
}

void hssb_builder_destroy(hssb_builder* b) { delete b; }

int64_t hssb_builder_add_leaf(hssb_builder* b, int64_t m, int64_t n, int64_t kr, int64_t kw, const double* D, int64_t ldd,
                              const double* U, int64_t ldu, const double* V, int64_t ldv) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!b) HSSB_FAIL(HSSB_ERR_ARG, "add_leaf: builder is NULL");
  if (m < 0 || n < 0 || kr < 0 || kw < 0) HSSB_FAIL(HSSB_ERR_ARG, "add_leaf: negative size");
  if (m > INT32_MAX || n > INT32_MAX || kr > INT32_MAX || kw > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "add_leaf: block too large");
  BNode nd;
  nd.leaf = true; nd.m = m; nd.n = n; nd.kr = kr; nd.kw = kw;
  int rc;
  if ((rc = copy_block(nd.blk[BK_D], D, ldd, m, n, "add_leaf D"))) return rc;
  if ((rc = copy_block(nd.blk[BK_U], U, ldu, m, kr, "add_leaf U"))) return rc;  // rows(U) == rows(D): hssmatrix.jl:41
  if ((rc = copy_block(nd.blk[BK_V], V, ldv, n, kw, "add_leaf V"))) return rc;  // rows(V) == cols(D): hssmatrix.jl:42
  b->nodes.push_back(std::move(nd));
  return (int64_t)b->nodes.size() - 1;
  });
}

int64_t hssb_builder_add_remote(hssb_builder* b, int64_t m, int64_t n, int64_t kr, int64_t kw) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!b) HSSB_FAIL(HSSB_ERR_ARG, "add_remote: builder is NULL");
  if (m < 0 || n < 0 || kr < 0 || kw < 0) HSSB_FAIL(HSSB_ERR_ARG, "add_remote: negative size");
  BNode nd;
  nd.remote = true; nd.m = m; nd.n = n; nd.kr = kr; nd.kw = kw;
  b->nodes.push_back(std::move(nd));
  return (int64_t)b->nodes.size() - 1;
  });
}

int64_t hssb_builder_add_branch(hssb_builder* b, int64_t left, int64_t right, int64_t kr, int64_t kw, const double* B12,
                                int64_t ldb12, const double* B21, int64_t ldb21, const double* R1, int64_t ldr1,
                                const double* W1, int64_t ldw1, const double* R2, int64_t ldr2, const double* W2,
                                int64_t ldw2) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!b) HSSB_FAIL(HSSB_ERR_ARG, "add_branch: builder is NULL");
  const int64_t nn = (int64_t)b->nodes.size();
  if (left < 0 || left >= nn || right < 0 || right >= nn || left == right)
    HSSB_FAIL(HSSB_ERR_ARG, "add_branch: invalid child ids %lld, %lld", (long long)left, (long long)right);
  if (b->nodes[(size_t)left].used || b->nodes[(size_t)right].used)
    HSSB_FAIL(HSSB_ERR_ARG, "add_branch: a child already has a parent");
  if (kr < 0 || kw < 0 || kr > INT32_MAX || kw > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "add_branch: bad gensize");
  BNode nd;
  nd.leaf = false; nd.left = left; nd.right = right; nd.kr = kr; nd.kw = kw;
  BNode& l = b->nodes[(size_t)left];
  BNode& r = b->nodes[(size_t)right];
  nd.m = l.m + r.m; nd.n = l.n + r.n;
  int rc;
  if ((rc = copy_block(nd.blk[BK_B12], B12, ldb12, l.kr, r.kw, "add_branch B12"))) return rc;
  if ((rc = copy_block(nd.blk[BK_B21], B21, ldb21, r.kr, l.kw, "add_branch B21"))) return rc;
  // translators of the children: R1 kr(left) x kr, W1 kw(left) x kw, ... (hssmatrix.jl:308-322)
  if ((rc = copy_block(l.blk[BK_R], R1, ldr1, l.kr, kr, "add_branch R1"))) return rc;
  if ((rc = copy_block(l.blk[BK_W], W1, ldw1, l.kw, kw, "add_branch W1"))) return rc;
  if ((rc = copy_block(r.blk[BK_R], R2, ldr2, r.kr, kr, "add_branch R2"))) return rc;
  if ((rc = copy_block(r.blk[BK_W], W2, ldw2, r.kw, kw, "add_branch W2"))) return rc;
  l.used = r.used = true;
  b->nodes.push_back(std::move(nd));
  return (int64_t)b->nodes.size() - 1;
  });
}

int hssb_builder_finalize(hssb_builder* b, int64_t root, int device, int shard_rank, int n_shards, hssb_matrix** out) {
  return guarded<int>([&]() -> int {
  if (!b || !out) HSSB_FAIL(HSSB_ERR_ARG, "finalize: NULL argument");
  *out = nullptr;
  if (root < 0 || root >= (int64_t)b->nodes.size()) HSSB_FAIL(HSSB_ERR_ARG, "finalize: invalid root id");
  int rc = check_device(device);
  if (rc) return rc;
  std::unique_ptr<hssb_matrix> H(new (std::nothrow) hssb_matrix());
  if (!H) HSSB_FAIL(HSSB_ERR_ALLOC, "finalize: out of memory");
  H->device = device; H->shard_rank = shard_rank; H->n_shards = n_shards;
  std::vector<BlockSource> src;
  nodes_from_builder(b, root, H.get(), src);
  rc = finish_matrix(H.get(), &src);
  if (rc) { hssb_destroy(H.release()); return rc; }
  *out = H.release();
  return HSSB_OK;
  });
}

int hssb_create_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, int device, int shard_rank,
                          int n_shards, hssb_matrix** out) {
  return guarded<int>([&]() -> int {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "create_synthetic: out is NULL");
  *out = nullptr;
  if (n <= 0 || leafsize <= 0 || rank < 0 || rank > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "create_synthetic: bad sizes");
  if (!is_pow2(n_shards)) HSSB_FAIL(HSSB_ERR_ARG, "n_shards must be a power of two");
  int rc = check_device(device);
  if (rc) return rc;
  std::unique_ptr<hssb_matrix> H(new (std::nothrow) hssb_matrix());
  if (!H) HSSB_FAIL(HSSB_ERR_ALLOC, "create_synthetic: out of memory");
  H->device = device; H->shard_rank = shard_rank; H->n_shards = n_shards;
  H->synthetic = true; H->seed = seed;
  nodes_synthetic(H.get(), n, leafsize, rank);
  H->synth_rank = rank;  // translator scale 1/sqrt(2 rank)
  rc = finish_matrix(H.get(), nullptr);
  if (rc) { hssb_destroy(H.release()); return rc; }
  *out = H.release();
  return HSSB_OK;
  });
}

int hssb_synthetic_rhs(uint64_t seed, int64_t n, int64_t nrhs, int64_t row0, int64_t rows, double* dX, int64_t ldx,
                       int device, void* stream) {
  return guarded<int>([&]() -> int {
  if (!dX || n <= 0 || nrhs < 0 || rows < 0 || row0 < 0 || row0 + rows > n || ldx < rows)
    HSSB_FAIL(HSSB_ERR_ARG, "synthetic_rhs: bad arguments");
  int rc = check_device(device);
  if (rc) return rc;
  DeviceGuard dg(device);
  if (rows * nrhs == 0) return HSSB_OK;
  const int64_t total = rows * nrhs;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  synth_rhs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(synth_key(seed, 0, KIND_X), n, nrhs, row0, rows, dX, ldx);
  HSSB_CUDA(cudaGetLastError());
  return HSSB_OK;
  });
}

int hssb_destroy(hssb_matrix* h) {
  return guarded<int>([&]() -> int {
  if (!h) return HSSB_OK;
  if (h->device < 0) { free_bush(h); delete h; return HSSB_OK; }
  DeviceGuard dg(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  invalidate_graphs(h);
  free_fast(h);
  free_tree(h);
  free_flow(h);
  free_bush(h);
  for (auto e : h->prof_events) cudaEventDestroy(e);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
  for (int r = 0; r < hssb_matrix::MAX_PEERS; ++r)
    if (h->peer_xchg && !h->peer_inprocess && r != h->shard_rank && r < h->n_shards) {
      if (h->peer_z[r]) cudaIpcCloseMemHandle(h->peer_z[r]);
      if (h->peer_flags[r]) cudaIpcCloseMemHandle(h->peer_flags[r]);
    }
  cudaFree(h->my_flags);
  cudaFree(h->pool_dev);
  cudaFree(h->pool_t_dev);
  cudaFree(h->ulv_pool_dev);
  cudaFree(h->ulv_pool_t_dev);
  cudaFree(h->tasks_dev);
  cudaFree(h->z_dev);
  cudaFree(h->f_dev);
  cudaFree(h->x_stage);
  cudaFree(h->y_stage);
  delete (Bounce*)h->bounce;
  if (h->copy_in) {
    cudaStreamDestroy(h->copy_in);
    cudaStreamDestroy(h->copy_out);
    for (int i = 0; i < hssb_matrix::MAX_BLOCKS; ++i) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_done[i]); }
  }
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return HSSB_OK;
  });
}

int hssb_info(const hssb_matrix* h, hssb_info_t* o) {
  return guarded<int>([&]() -> int {
  if (!h || !o) HSSB_FAIL(HSSB_ERR_ARG, "hssb_info: NULL argument");
  memset(o, 0, sizeof(*o));
  o->m = h->m; o->n = h->n; o->local_m = h->local_m; o->local_n = h->local_n;
  o->local_row0 = h->local_row0; o->local_col0 = h->local_col0;
  o->n_nodes = (int64_t)h->nodes.size(); o->n_leaves = (int64_t)h->leaves.size(); o->depth = h->depth;
  o->max_leaf_m = h->max_leaf_m; o->max_leaf_n = h->max_leaf_n; o->max_rank = h->max_rank;
  o->pool_bytes = h->pool_len * 8; o->gen_elems = h->gen_elems; o->flops_per_rhs = h->flops_per_rhs;
  o->z_rows = h->z_rows; o->f_rows = h->f_rows;
  o->shard_rank = h->shard_rank; o->n_shards = h->n_shards; o->device = h->device;
  o->uniform = h->uniform ? 1 : 0;
  return HSSB_OK;
  });
}

int hssb_node_info(const hssb_matrix* h, int64_t node, hssb_node_t* o) {
  return guarded<int>([&]() -> int {
  if (!h || !o) HSSB_FAIL(HSSB_ERR_ARG, "hssb_node_info: NULL argument");
  if (node < 0 || node >= (int64_t)h->nodes.size()) HSSB_FAIL(HSSB_ERR_ARG, "hssb_node_info: node id out of range");
  const Node& t = h->nodes[(size_t)node];
  o->left = t.left; o->right = t.right; o->parent = t.parent;
  o->depth = t.depth; o->is_leaf = t.leaf; o->is_remote = t.remote;
  o->row0 = t.row0; o->m = t.m; o->col0 = t.col0; o->n = t.n; o->kr = t.kr; o->kw = t.kw;
  return HSSB_OK;
  });
}

int hssb_get_block(const hssb_matrix* h, int64_t node, int kind, double* out, int64_t out_len) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_get_block: NULL handle");
  if (node < 0 || node >= (int64_t)h->nodes.size() || kind < 0 || kind >= BK_COUNT)
    HSSB_FAIL(HSSB_ERR_ARG, "hssb_get_block: bad node or kind");
  const Node& t = h->nodes[(size_t)node];
  const int64_t rows = t.rows[kind], cols = t.cols[kind];  // as stored
  if (out_len < rows * cols) HSSB_FAIL(HSSB_ERR_ARG, "hssb_get_block: need %lld doubles", (long long)(rows * cols));
  if (rows * cols == 0 || t.off[kind] < 0) return HSSB_OK;
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_get_block: out is NULL");
  std::vector<double> tmp((size_t)(rows * cols));
  if (h->device < 0) {
    for (int64_t j = 0; j < cols; ++j)
      memcpy(tmp.data() + j * rows, h->pool_host.data() + t.off[kind] + j * t.ld[kind], (size_t)rows * 8);
  } else {
    DeviceGuard dg(h->device);
    HSSB_CUDA(cudaMemcpy2D(tmp.data(), (size_t)rows * 8, h->pool_dev + t.off[kind], (size_t)t.ld[kind] * 8, (size_t)rows * 8,
                           (size_t)cols, cudaMemcpyDeviceToHost));
  }
  if (stored_transposed(kind)) {  // the pool holds V' / W'; hand back V (n x kw) / W (kw x kw(parent))
    for (int64_t j = 0; j < cols; ++j)
      for (int64_t i = 0; i < rows; ++i) out[i * cols + j] = tmp[(size_t)(j * rows + i)];
  } else {
    memcpy(out, tmp.data(), tmp.size() * 8);
  }
  return HSSB_OK;
  });
}

int hssb_reserve(hssb_matrix* h, int64_t max_nrhs) {
  return guarded<int>([&]() -> int {
  if (!h || max_nrhs < 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_reserve: bad argument");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle has no device");
  DeviceGuard dg(h->device);
  return ensure_workspace(h, max_nrhs);
  });
}

// How Y = A' X runs: 0 = forward plan over the adjoint twin pool, 1 = any-shape transposed task
// table over the primary pool (single shard only), < 0 = error.
static int select_adjoint(hssb_matrix* h) {
  const int tw = ensure_twin(h);
  if (tw < 0) return tw;
  if (tw == 1 && h->n_shards != 1)
    HSSB_FAIL(HSSB_ERR_STATE, "the transposed product of a sharded matrix needs the adjoint twin pool "
                              "(uniform tree, HSSB_OPT_ADJOINT_TWIN = 1 and room for a second pool on the device)");
  return tw;
}

// Mode 2 (hssb_solve): the factor pool must exist; the first solve factorises.
static int prepare_solve(hssb_matrix* h) {
  if (h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_solve: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle: hssb200 has no CPU fallback, the solve needs a B200");
  if (!h->ulv_factored || !h->ulv_pool_dev) return ulv_factor_device(h);
  return HSSB_OK;
}

// Mode 3 (hssb_solve_t: A' Z = B, i.e. `/(A, hssB)` of hssmatrix.jl:236 without building hssB'): on a uniform tree A'
// has the shapes of A, so the solve plan is shared and only the factors differ -- they are computed from the adjoint
// twin pool into a second factor pool.
static int prepare_solve_t(hssb_matrix* h) {
  if (h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_solve_t: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle: hssb200 has no CPU fallback, the solve needs a B200");
  const int tw = ensure_twin(h);
  if (tw < 0) return tw;
  if (tw != 0)
    HSSB_FAIL(HSSB_ERR_STATE, "hssb_solve_t needs the adjoint twin pool (uniform tree, HSSB_OPT_ADJOINT_TWIN = 1 and room for a second "
                              "pool on the device); for other trees pack the adjoint matrix and use hssb_solve");
  if (!h->ulv_t_factored || !h->ulv_pool_t_dev) return ulv_factor_device(h, true);
  return HSSB_OK;
}

static int matmul_dev_impl(hssb_matrix* h, int trans, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* dX, int64_t ldx,
                           double* dY, int64_t ldy, double alpha, double beta, void* stream) {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: NULL handle");
  if (trans >= 2 && h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_solve: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  // DimensionMismatch checks of matmul.jl:19-20 (for A' the roles of the two dimensions swap)
  const int64_t need_x = trans ? h->local_m : h->local_n, need_y = trans ? h->local_n : h->local_m;
  if (rows_x != need_x)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: first dimension of B (%lld) does not match second dimension of A (%lld)",
              (long long)rows_x, (long long)need_x);
  if (rows_y != need_y)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: dimensions of C (%lld rows) don't match up with A (%lld rows)",
              (long long)rows_y, (long long)need_y);
  if (nrhs < 0 || nrhs > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: bad nrhs");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle: hssb200 has no CPU fallback, the product needs a B200");
  if (nrhs == 0 || rows_y == 0) return HSSB_OK;
  if (ldx < std::max<int64_t>(rows_x, 1) || ldy < rows_y) HSSB_FAIL(HSSB_ERR_DIM, "hssb_matmul: leading dimension too small");
  if ((rows_x > 0 && !dX) || !dY) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: NULL matrix pointer");
  DeviceGuard dg(h->device);
  if (!dg.ok) HSSB_FAIL(HSSB_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
  int rc = ensure_workspace(h, nrhs, trans >= 2);
  if (rc) return rc;
  const double* pool = h->pool_dev;
  if (trans == 1) {
    rc = select_adjoint(h);
    if (rc < 0) return rc;
    if (rc == 0) { pool = h->pool_t_dev; trans = 0; }  // A' X = the forward plan over the twin pool
  } else if (trans == 2) {
    rc = prepare_solve(h);
    if (rc) return rc;
    pool = h->ulv_pool_dev;
  } else if (trans == 3) {
    rc = prepare_solve_t(h);
    if (rc) return rc;
    pool = h->ulv_pool_t_dev;
    trans = 2;  // same plan, other factors
  }
  if (trans == 0 && h->tree_kernel && !h->force_generic) {  // allocates: must happen outside graph capture
    rc = ensure_tree_plan(h);
    if (rc) return rc;
  }
  if (h->flow_kernel && trans <= 1 && h->n_shards == 1) {  // likewise
    rc = ensure_flow_plan(h, trans, nrhs);
    if (rc) return rc;
  }
  if (h->bush_kernel && trans <= 1 && h->n_shards == 1 && (h->bush_kernel >= 2 || bush_eligible(h))) {
    rc = ensure_bush_plan(h, trans, nrhs);
    if (rc) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  CallParams cp;
  cp.pool = pool; cp.X = dX; cp.Y = dY; cp.Z = h->z_dev; cp.F = h->f_dev;
  cp.ldx = ldx; cp.ldy = ldy; cp.nrhs = (int32_t)nrhs; cp.alpha = alpha; cp.beta = beta; cp.debug = h->debug_mode; cp.trans = trans;
  // graph replay: always for the host entry (its staging pointers are stable), on request for
  // caller-owned device pointers (a new pointer set costs a capture + instantiate)
  if ((h->use_graph || h->in_host_call) && !h->profile) return run_graph(h, cp, st);
  if (h->prepare_only) return HSSB_OK;  // plain launches have nothing to capture ahead of time
  return run_phases(h, cp, st);
}

// Device staging buffers, copy streams and events of the host-pointer entry (allocates, and cudaFree
// synchronises the device: callers that run several shards of one device concurrently do this up front).
static int ensure_host_entry(hssb_matrix* h, int64_t nrhs) {
  if (nrhs > h->stage_nrhs) {
    cudaFree(h->x_stage); cudaFree(h->y_stage);
    h->x_stage = h->y_stage = nullptr; h->stage_nrhs = 0;
    const size_t rmax = (size_t)std::max<int64_t>(std::max(h->local_n, h->local_m), 1);  // either stage may hold X or Y (A or A')
    const size_t xb = rmax * (size_t)nrhs * 8, yb = rmax * (size_t)nrhs * 8;
    if (cudaMalloc(&h->x_stage, xb) != cudaSuccess || cudaMalloc(&h->y_stage, yb) != cudaSuccess) {
      cudaGetLastError();
      HSSB_FAIL(HSSB_ERR_ALLOC, "device allocation of X/Y staging (%.3f GB) failed", (xb + yb) * 1e-9);
    }
    h->stage_nrhs = nrhs;
    invalidate_graphs(h);
  }
  if (!h->copy_in) {
    HSSB_CUDA(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
    HSSB_CUDA(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < hssb_matrix::MAX_BLOCKS; ++i) {
      HSSB_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      HSSB_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
  }
  return HSSB_OK;
}

static int ensure_bounce(hssb_matrix* h) {
  if (h->bounce) return HSSB_OK;
  Bounce* bn = new (std::nothrow) Bounce();
  if (!bn) HSSB_FAIL(HSSB_ERR_ALLOC, "hssb_matmul: out of memory");
  if (int rc = bn->init(h->device)) { delete bn; return rc; }
  h->bounce = bn;
  return HSSB_OK;
}

// prepare = true: everything the call needs is allocated, captured and instantiated, nothing is copied or launched
// (hssb_group: the shards of one call are prepared one after the other and only then run concurrently).
static int matmul_host_impl(hssb_matrix* h, int trans, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx,
                            double* Y, int64_t ldy, double alpha, double beta, bool prepare = false) {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: NULL handle");
  if (trans >= 2 && h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_solve: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  const int64_t need_x = trans ? h->local_m : h->local_n, need_y = trans ? h->local_n : h->local_m;
  if (rows_x != need_x)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: first dimension of B (%lld) does not match second dimension of A (%lld)",
              (long long)rows_x, (long long)need_x);
  if (rows_y != need_y)
    HSSB_FAIL(HSSB_ERR_DIM, "DimensionMismatch: dimensions of C (%lld rows) don't match up with A (%lld rows)",
              (long long)rows_y, (long long)need_y);
  if (nrhs < 0 || nrhs > INT32_MAX) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: bad nrhs");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle: hssb200 has no CPU fallback, the product needs a B200");
  if (nrhs == 0 || rows_y == 0) return HSSB_OK;
  if (ldx < std::max<int64_t>(rows_x, 1) || ldy < rows_y) HSSB_FAIL(HSSB_ERR_DIM, "hssb_matmul: leading dimension too small");
  if ((rows_x > 0 && !X) || !Y) HSSB_FAIL(HSSB_ERR_ARG, "hssb_matmul: NULL matrix pointer");
  DeviceGuard dg(h->device);
  if (!dg.ok) HSSB_FAIL(HSSB_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
  if (trans == 1) {  // fail before any copy is queued
    const int a = select_adjoint(h);
    if (a < 0) return a;
  } else if (trans == 2) {
    const int a = prepare_solve(h);
    if (a) return a;
  } else if (trans == 3) {
    const int a = prepare_solve_t(h);
    if (a) return a;
  }
  if (int rc = ensure_host_entry(h, nrhs)) return rc;
  const int64_t sx = std::max<int64_t>(rows_x, 1), sy = rows_y;
  // every shard must cut the call into the same blocks (each block contains the exchange step)
  const int64_t rows_eff = h->n_shards > 1 ? (h->m + h->n) / h->n_shards : rows_x + rows_y;
  int64_t cb = nrhs;
  if (h->pipeline_cols > 0) cb = std::min<int64_t>(nrhs, h->pipeline_cols);
  else if (nrhs >= 16 && rows_eff * nrhs * 8 >= ((int64_t)64 << 20)) cb = std::max<int64_t>(8, ((nrhs + 7) / 8 + 7) / 8 * 8);  // ~8 blocks: measured best on PCIe Gen5 (tools/e2e_blocks.py)
  int64_t nblk = (nrhs + cb - 1) / cb;
  if (nblk > hssb_matrix::MAX_BLOCKS) { cb = (nrhs + hssb_matrix::MAX_BLOCKS - 1) / hssb_matrix::MAX_BLOCKS; nblk = (nrhs + cb - 1) / cb; }
  // Pageable caller memory (an ordinary Julia Matrix) goes through the library's pinned slot rings and
  // worker threads (hssb_hostpipe.h); pinned / registered memory is handed to the copy engines directly.
  // Small calls are not worth the thread hand-offs.
  const bool big = (rows_x + rows_y) * nrhs * 8 >= ((int64_t)16 << 20);
  const bool bounce_in = rows_x > 0 && h->host_bounce && (h->host_bounce == 2 || (big && host_ptr_pageable(X)));
  const bool bounce_out = h->host_bounce && (h->host_bounce == 2 || (big && host_ptr_pageable(Y)));
  if (bounce_in || bounce_out) {
    if (int rc = ensure_bounce(h)) return rc;
  }
  Bounce* bn = (Bounce*)h->bounce;
  if (prepare) {
    if (h->profile) return HSSB_OK;  // profiled calls launch phase by phase: nothing to capture
    int rc = HSSB_OK;
    h->in_host_call = h->prepare_only = true;
    for (int64_t j = 0; j < nblk && !rc; ++j) {
      const int64_t c0 = j * cb, nc = std::min(cb, nrhs - c0);
      rc = matmul_dev_impl(h, trans, rows_y, rows_x, nc, h->x_stage + c0 * sx, sx, h->y_stage + c0 * sy, sy, alpha, beta, h->stream);
    }
    h->in_host_call = h->prepare_only = false;
    if (!rc) HSSB_CUDA(cudaStreamSynchronize(h->stream));
    return rc;
  }
  h->last_bounce = (bounce_in ? 1 : 0) | (bounce_out ? 2 : 0);
  std::vector<Latch> in_latch((size_t)(bounce_in ? nblk : 0));
  Latch out_latch;
  auto drain = [&]() {  // nothing may still reference the caller's memory or this frame when we return
    for (auto& l : in_latch) l.wait();
    out_latch.wait();
    cudaStreamSynchronize(h->copy_in); cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->copy_out);
  };
  // copies of this call must not overtake the previous call's use of the staging buffers
  HSSB_CUDA(cudaEventRecord(h->ev_done[0], h->stream));
  HSSB_CUDA(cudaStreamWaitEvent(h->copy_in, h->ev_done[0], 0));
  for (int64_t j = 0; j < nblk; ++j) {
    const int64_t c0 = j * cb, nc = std::min(cb, nrhs - c0);
    if (bounce_in) {
      Latch* l = &in_latch[(size_t)j];
      Bounce::for_pieces(rows_x, nc, ldx, sx, [&](size_t oh, size_t od, size_t bytes) {
        bn->submit_in((const char*)(X + c0 * ldx) + oh, (char*)(h->x_stage + c0 * sx) + od, bytes, h->copy_in, l);
      });
      if (beta != 0.0)
        Bounce::for_pieces(rows_y, nc, ldy, sy, [&](size_t oh, size_t od, size_t bytes) {
          bn->submit_in((const char*)(Y + c0 * ldy) + oh, (char*)(h->y_stage + c0 * sy) + od, bytes, h->copy_in, l);
        });
      continue;  // the block's event is recorded once its pieces have been queued (below)
    }
    cudaError_t e = cudaSuccess;
    if (rows_x > 0)
      e = cudaMemcpy2DAsync(h->x_stage + c0 * sx, (size_t)sx * 8, X + c0 * ldx, (size_t)ldx * 8, (size_t)rows_x * 8, (size_t)nc,
                            cudaMemcpyHostToDevice, h->copy_in);
    if (e == cudaSuccess && beta != 0.0)
      e = cudaMemcpy2DAsync(h->y_stage + c0 * sy, (size_t)sy * 8, Y + c0 * ldy, (size_t)ldy * 8, (size_t)rows_y * 8, (size_t)nc,
                            cudaMemcpyHostToDevice, h->copy_in);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_in[j], h->copy_in);
    if (e != cudaSuccess) { drain(); HSSB_CUDA(e); }
  }
  for (int64_t j = 0; j < nblk; ++j) {
    const int64_t c0 = j * cb, nc = std::min(cb, nrhs - c0);
    cudaError_t e = cudaSuccess;
    if (bounce_in) {
      e = in_latch[(size_t)j].wait();  // every piece of the block is queued on copy_in
      if (e == cudaSuccess) e = cudaEventRecord(h->ev_in[j], h->copy_in);
    }
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_in[j], 0);
    if (e != cudaSuccess) { drain(); HSSB_CUDA(e); }
    h->in_host_call = true;
    int rc = matmul_dev_impl(h, trans, rows_y, rows_x, nc, h->x_stage + c0 * sx, sx, h->y_stage + c0 * sy, sy, alpha, beta, h->stream);
    h->in_host_call = false;
    if (rc) { drain(); return rc; }
    e = cudaEventRecord(h->ev_done[j], h->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->copy_out, h->ev_done[j], 0);
    if (e == cudaSuccess) {
      if (bounce_out)
        Bounce::for_pieces(rows_y, nc, ldy, sy, [&](size_t oh, size_t od, size_t bytes) {
          bn->submit_out((const char*)(h->y_stage + c0 * sy) + od, (char*)(Y + c0 * ldy) + oh, bytes, h->copy_out, &out_latch);
        });
      else
        e = cudaMemcpy2DAsync(Y + c0 * ldy, (size_t)ldy * 8, h->y_stage + c0 * sy, (size_t)sy * 8, (size_t)rows_y * 8, (size_t)nc,
                              cudaMemcpyDeviceToHost, h->copy_out);
    }
    if (e != cudaSuccess) { drain(); HSSB_CUDA(e); }
  }
  {
    const cudaError_t e = out_latch.wait();
    if (e != cudaSuccess) { drain(); HSSB_CUDA(e); }
  }
  HSSB_CUDA(cudaStreamSynchronize(h->copy_out));
  HSSB_CUDA(cudaStreamSynchronize(h->stream));
  return HSSB_OK;
}

int hssb_matmul(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y,
                int64_t ldy, double alpha, double beta) {
  return matmul_host_impl(h, 0, rows_y, rows_x, nrhs, X, ldx, Y, ldy, alpha, beta);
}

int hssb_matmul_dev(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* dX, int64_t ldx, double* dY,
                    int64_t ldy, double alpha, double beta, void* stream) {
  return matmul_dev_impl(h, 0, rows_y, rows_x, nrhs, dX, ldx, dY, ldy, alpha, beta, stream);
}

int hssb_matmul_t(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y,
                  int64_t ldy, double alpha, double beta) {
  return matmul_host_impl(h, 1, rows_y, rows_x, nrhs, X, ldx, Y, ldy, alpha, beta);
}

int hssb_matmul_t_dev(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* dX, int64_t ldx, double* dY,
                      int64_t ldy, double alpha, double beta, void* stream) {
  return matmul_dev_impl(h, 1, rows_y, rows_x, nrhs, dX, ldx, dY, ldy, alpha, beta, stream);
}

// hssA \ B (hssmatrix.jl:234 -> ulvfactsolve, ulvfactor.jl:10-19)
int hssb_ulv_factor(hssb_matrix* h) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_ulv_factor: NULL handle");
  if (h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_ulv_factor: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle: hssb200 has no CPU fallback, the factorisation needs a B200");
  DeviceGuard dg(h->device);
  if (!dg.ok) HSSB_FAIL(HSSB_ERR_CUDA, "cudaSetDevice(%d) failed", h->device);
  invalidate_graphs(h);
  return ulv_factor_device(h);
  });
}

int hssb_solve(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* B, int64_t ldb, double* Z, int64_t ldz) {
  return matmul_host_impl(h, 2, rows, rows, nrhs, B, ldb, Z, ldz, 1.0, 0.0);
}

int hssb_solve_dev(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* dB, int64_t ldb, double* dZ, int64_t ldz,
                   void* stream) {
  return matmul_dev_impl(h, 2, rows, rows, nrhs, dB, ldb, dZ, ldz, 1.0, 0.0, stream);
}

// A' \ B: `/(A, hssB) = ulvfactsolve(hssB', collect(A'))'` (hssmatrix.jl:236) without the adjoint copy
int hssb_solve_t(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* B, int64_t ldb, double* Z, int64_t ldz) {
  return matmul_host_impl(h, 3, rows, rows, nrhs, B, ldb, Z, ldz, 1.0, 0.0);
}

int hssb_solve_t_dev(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* dB, int64_t ldb, double* dZ, int64_t ldz,
                     void* stream) {
  return matmul_dev_impl(h, 3, rows, rows, nrhs, dB, ldb, dZ, ldz, 1.0, 0.0, stream);
}

int hssb_ulv_info(const hssb_matrix* h, hssb_ulv_info_t* o) {
  return guarded<int>([&]() -> int {
  if (!h || !o) HSSB_FAIL(HSSB_ERR_ARG, "hssb_ulv_info: NULL argument");
  memset(o, 0, sizeof(*o));
  o->supported = h->ulv.empty() ? 0 : 1;
  o->factored = h->ulv_factored ? 1 : 0;
  if (!h->ulv.empty()) {
    o->pool_bytes = h->ulv_pool_len * 8;
    o->flops_per_rhs = h->ulv_flops_per_rhs;
    o->z_rows = h->ulv_z_rows; o->f_rows = h->ulv_f_rows;
  }
  return HSSB_OK;
  });
}

int hssb_sync(hssb_matrix* h) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_sync: NULL handle");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle has no device");
  DeviceGuard dg(h->device);
  HSSB_CUDA(cudaStreamSynchronize(h->stream));
  return HSSB_OK;
  });
}

// HSSB_OPT_ULV_FAST: the ULV plan (tasks after ulv_task0, factor-pool layout, workspace rows) is rebuilt in
// the other form; factors and graphs of the old form are dropped.
static int rebuild_ulv_plan(hssb_matrix* h) {
  if (h->ulv_task0 < 0) return HSSB_OK;
  if (h->device >= 0) {
    if (h->stream) cudaStreamSynchronize(h->stream);
    invalidate_graphs(h);
    cudaFree(h->ulv_pool_dev);
    cudaFree(h->ulv_pool_t_dev);
    h->ulv_pool_dev = h->ulv_pool_t_dev = nullptr;
  }
  h->ulv_pool_host.clear();
  h->ulv_factored = h->ulv_t_factored = false;
  h->ws_ulv = false;  // the solve's workspace rows change with the form: re-sized by the next solve
  h->tasks_host.resize((size_t)h->ulv_task0);
  build_plan_ulv(h);
  if (h->device >= 0 && !h->tasks_host.empty()) {
    GTask* fresh = nullptr;
    HSSB_CUDA(cudaMalloc(&fresh, h->tasks_host.size() * sizeof(GTask)));
    HSSB_CUDA(cudaMemcpy(fresh, h->tasks_host.data(), h->tasks_host.size() * sizeof(GTask), cudaMemcpyHostToDevice));
    cudaFree(h->tasks_dev);
    h->tasks_dev = fresh;
  }
  return HSSB_OK;
}

int hssb_set_option(hssb_matrix* h, int opt, int64_t value) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_set_option: NULL handle");
  switch (opt) {
    case HSSB_OPT_FORCE_GENERIC: h->force_generic = value != 0; break;
    case HSSB_OPT_USE_GRAPH: h->use_graph = value != 0; break;
    case HSSB_OPT_PROFILE: h->profile = value != 0; break;
    case HSSB_OPT_DEBUG: h->debug_mode = (int)value; break;
    case HSSB_OPT_PIPELINE_COLS: h->pipeline_cols = value; break;
    case HSSB_OPT_ADJOINT_TWIN: h->adjoint_twin = value != 0; break;
    case HSSB_OPT_TREE_KERNEL: h->tree_kernel = (int)value; break;
    case HSSB_OPT_LEAF_KERNEL:
      if (value < 1 || value > 3) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_LEAF_KERNEL: 1, 2 or 3");
      h->leaf_kernel = (int)value;
      break;
    case HSSB_OPT_LEAF_FUSION: h->leaf_fusion = value != 0; break;
    case HSSB_OPT_FLOW_KERNEL:
      if (value < 0 || value > 2) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_FLOW_KERNEL: 0, 1 or 2");
      h->flow_kernel = (int)value;
      break;
    case HSSB_OPT_PDL:
      if (value < 0 || value > 15) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_PDL: bits 0-3");
      h->pdl = (int)value;
      break;
    case HSSB_OPT_BUSH_KERNEL:
      if (value < 0 || value > 2) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_BUSH_KERNEL: 0, 1 or 2");
      h->bush_kernel = (int)value;
      break;
    case HSSB_OPT_BUSH_LEVELS: {
      const int hb = (int)(value / 16), hb0 = (int)(value % 16);
      if (value < 0 || hb < 1 || hb > 15) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_BUSH_LEVELS: levels per bush (1-15) * 16 + leaf-bush levels (0-15)");
      h->bush_levels = hb; h->bush_levels0 = hb0;
      if (h->device >= 0) {
        DeviceGuard dgb(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        free_bush(h);
      } else {
        free_bush(h);
      }
      break;
    }
    case HSSB_OPT_HOST_BOUNCE:
      if (value < 0 || value > 2) HSSB_FAIL(HSSB_ERR_ARG, "HSSB_OPT_HOST_BOUNCE: 0, 1 or 2");
      h->host_bounce = (int)value;
      return HSSB_OK;
    case HSSB_OPT_ULV_FAST: {
      if (h->ulv_fast_form == (value != 0)) return HSSB_OK;
      h->ulv_fast_form = value != 0;
      if (h->device < 0) return rebuild_ulv_plan(h);
      DeviceGuard dgu(h->device);
      return rebuild_ulv_plan(h);
    }
    default: HSSB_FAIL(HSSB_ERR_ARG, "hssb_set_option: unknown option %d", opt);
  }
  if (h->device < 0) return HSSB_OK;
  DeviceGuard dg(h->device);
  invalidate_graphs(h);
  if (opt == HSSB_OPT_ADJOINT_TWIN) {  // 0 releases the twin, 1 lets the next transposed product (re)build it
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_twin(h);
  }
  return HSSB_OK;
  });
}

int64_t hssb_get_option(const hssb_matrix* h, int opt) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!h) return -1;
  switch (opt) {
    case HSSB_OPT_FORCE_GENERIC: return h->force_generic;
    case HSSB_OPT_USE_GRAPH: return h->use_graph;
    case HSSB_OPT_PROFILE: return h->profile;
    case HSSB_OPT_DEBUG: return h->debug_mode;
    case HSSB_OPT_PIPELINE_COLS: return h->pipeline_cols;
    case HSSB_OPT_ADJOINT_TWIN: return !h->adjoint_twin ? 0 : (h->pool_t_dev ? 2 : 1);  // 2: built and in use
    case HSSB_OPT_ULV_FAST: return !h->ulv_fast_form ? 0 : (h->ulv_ff ? 2 : 1);          // 2: the plan is in fast form
    case HSSB_OPT_TREE_KERNEL: return h->tree_kernel;
    case HSSB_OPT_HOST_BOUNCE: return h->host_bounce;
    case HSSB_OPT_LEAF_KERNEL: return h->leaf_kernel;
    case HSSB_OPT_LEAF_FUSION: return h->leaf_fusion;
    case HSSB_OPT_FLOW_KERNEL: {
      if (!h->flow_kernel) return 0;
      const FlowPlan* fp = (const FlowPlan*)h->flow_plan[0];
      return fp && fp->usable ? 2 : 1;  // 2: the product plan qualifies and has been set up
    }
    case HSSB_OPT_BUSH_KERNEL: {
      if (!h->bush_kernel) return 0;
      const BushPlan* bp = (const BushPlan*)h->bush_plan[0];
      return bp && bp->usable && bp->sync_dev && (h->bush_kernel >= 2 || bush_eligible(h)) ? 3 : h->bush_kernel;
    }
    case HSSB_OPT_BUSH_LEVELS: return h->bush_levels * 16 + h->bush_levels0;
    case HSSB_OPT_PDL: return h->pdl;
    case HSSB_OPT_LAST_FACTOR_US: return h->ulv_last_factor_us;
    case HSSB_OPT_LAST_BOUNCE: return h->last_bounce;
    case HSSB_OPT_HOST_THREADS: return host_pool(0).size();
    default: return -1;
  }
  });
}

int64_t hssb_launch_count(const hssb_matrix* h) { return h ? h->launches : 0; }

int hssb_phase_count(const hssb_matrix* h) { return h ? (int)phase_list(h, h->prof_mode).size() : 0; }

int hssb_phase_time(hssb_matrix* h, int i, hssb_phase_time_t* o) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!h || !o || i < 0 || i >= (int)phase_list(h, h->prof_mode).size()) HSSB_FAIL(HSSB_ERR_ARG, "hssb_phase_time: bad argument");
  const Phase& ph = phase_list(h, h->prof_mode)[(size_t)i];
  memset(o, 0, sizeof(*o));
  o->kind = ph.kind; o->level = ph.level; o->top = ph.top; o->fast = ph.fast; o->ntasks = ph.ntasks;
  int64_t gen = 0, xrows = 0, yrows = 0, fl = 0;
  for (int64_t t = ph.task0; t < ph.task0 + ph.ntasks; ++t) {
    const GTask& g = h->tasks_host[(size_t)t];
    gen += (int64_t)g.M * g.K0 + (int64_t)g.M * g.K1;
    fl += 2ll * g.M * ((int64_t)g.K0 + g.K1);
    if (g.sb0 == SRC_X) xrows += g.K0;
    if (g.sc == SRC_Y) yrows += g.M;
  }
  o->flops_per_rhs = fl; o->gen_elems = gen; o->x_rows = xrows; o->y_rows = yrows;
  o->ms = -1.0;
  if (h->device >= 0 && h->prof_events.size() > (size_t)i + 1 && h->prof_nrhs > 0) {
    DeviceGuard dg(h->device);
    HSSB_CUDA(cudaEventSynchronize(h->prof_events[(size_t)i + 1]));
    float ms = 0;
    HSSB_CUDA(cudaEventElapsedTime(&ms, h->prof_events[(size_t)i], h->prof_events[(size_t)i + 1]));
    o->ms = ms;
  }
  return HSSB_OK;
  });
}

int hssb_comm_unique_id(void* id128) {
  return guarded<int>([&]() -> int {
  if (!id128) HSSB_FAIL(HSSB_ERR_ARG, "hssb_comm_unique_id: NULL buffer");
  int rc = load_nccl();
  if (rc) return rc;
  HSSB_NCCL(g_nccl.GetUniqueId(id128));
  return HSSB_OK;
  });
}

int hssb_comm_init(hssb_matrix* h, const void* id128, int rank, int n_ranks) {
  return guarded<int>([&]() -> int {
  if (!h || !id128) HSSB_FAIL(HSSB_ERR_ARG, "hssb_comm_init: NULL argument");
  if (n_ranks != h->n_shards || rank != h->shard_rank)
    HSSB_FAIL(HSSB_ERR_ARG, "hssb_comm_init: rank %d/%d does not match shard %d/%d", rank, n_ranks, h->shard_rank, h->n_shards);
  int rc = load_nccl();
  if (rc) return rc;
  DeviceGuard dg(h->device);
  Id128 id;
  memcpy(id.b, id128, 128);
  void* comm = nullptr;
  HSSB_NCCL(g_nccl.CommInitRank(&comm, n_ranks, id, rank));
  h->nccl_comm = comm;
  return HSSB_OK;
  });
}


int hssb_xchg_export(hssb_matrix* h, void* out128) {
  return guarded<int>([&]() -> int {
  if (!h || !out128) HSSB_FAIL(HSSB_ERR_ARG, "hssb_xchg_export: NULL argument");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle has no device");
  if (h->n_shards < 2 || h->n_shards > hssb_matrix::MAX_PEERS) HSSB_FAIL(HSSB_ERR_STATE, "hssb_xchg_export: needs 2..16 shards");
  if (!h->z_dev) HSSB_FAIL(HSSB_ERR_STATE, "hssb_xchg_export: call hssb_reserve(max_nrhs) first");
  DeviceGuard dg(h->device);
  if (!h->my_flags) {
    HSSB_CUDA(cudaMalloc(&h->my_flags, XCHG_FLAG_WORDS * sizeof(unsigned long long)));
    HSSB_CUDA(cudaMemset(h->my_flags, 0, XCHG_FLAG_WORDS * sizeof(unsigned long long)));
    HSSB_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t hz, hf;
  HSSB_CUDA(cudaIpcGetMemHandle(&hz, h->z_dev));
  HSSB_CUDA(cudaIpcGetMemHandle(&hf, h->my_flags));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out128, &hz, 64);
  memcpy((char*)out128 + 64, &hf, 64);
  h->xchg_exported = true;
  return HSSB_OK;
  });
}

int hssb_xchg_import(hssb_matrix* h, const void* all_handles, int n_ranks) {
  return guarded<int>([&]() -> int {
  if (!h || !all_handles) HSSB_FAIL(HSSB_ERR_ARG, "hssb_xchg_import: NULL argument");
  if (n_ranks != h->n_shards) HSSB_FAIL(HSSB_ERR_ARG, "hssb_xchg_import: %d handles for %d shards", n_ranks, h->n_shards);
  if (!h->xchg_exported) HSSB_FAIL(HSSB_ERR_STATE, "hssb_xchg_import: call hssb_xchg_export first");
  DeviceGuard dg(h->device);
  for (int r = 0; r < n_ranks; ++r) {
    if (r == h->shard_rank) { h->peer_z[r] = h->z_dev; h->peer_flags[r] = h->my_flags; continue; }
    cudaIpcMemHandle_t hz, hf;
    memcpy(&hz, (const char*)all_handles + 128 * r, 64);
    memcpy(&hf, (const char*)all_handles + 128 * r + 64, 64);
    void *pz = nullptr, *pf = nullptr;
    HSSB_CUDA(cudaIpcOpenMemHandle(&pz, hz, cudaIpcMemLazyEnablePeerAccess));
    HSSB_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    h->peer_z[r] = (double*)pz;
    h->peer_flags[r] = (unsigned long long*)pf;
  }
  h->peer_xchg = true;
  invalidate_graphs(h);
  return HSSB_OK;
  });
}

#include "hssb_file.h"

// ---- test hooks: host-only planning (no device required) --------------------
int hssb_plan_only(hssb_builder* b, int64_t root, int shard_rank, int n_shards, hssb_matrix** out) {
  return guarded<int>([&]() -> int {
  if (!b || !out) HSSB_FAIL(HSSB_ERR_ARG, "plan_only: NULL argument");
  *out = nullptr;
  if (root < 0 || root >= (int64_t)b->nodes.size()) HSSB_FAIL(HSSB_ERR_ARG, "plan_only: invalid root id");
  std::unique_ptr<hssb_matrix> H(new (std::nothrow) hssb_matrix());
  if (!H) HSSB_FAIL(HSSB_ERR_ALLOC, "plan_only: out of memory");
  H->device = -1; H->shard_rank = shard_rank; H->n_shards = n_shards;
  std::vector<BlockSource> src;
  nodes_from_builder(b, root, H.get(), src);
  int rc = plan_matrix(H.get());
  if (rc) return rc;
  fill_pool_host(H.get(), &src);
  *out = H.release();
  return HSSB_OK;
  });
}

int hssb_plan_only_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, int shard_rank, int n_shards,
                             hssb_matrix** out) {
  return guarded<int>([&]() -> int {
  if (!out) HSSB_FAIL(HSSB_ERR_ARG, "plan_only_synthetic: out is NULL");
  *out = nullptr;
  if (n <= 0 || leafsize <= 0 || rank < 0) HSSB_FAIL(HSSB_ERR_ARG, "plan_only_synthetic: bad sizes");
  if (!is_pow2(n_shards)) HSSB_FAIL(HSSB_ERR_ARG, "n_shards must be a power of two");
  std::unique_ptr<hssb_matrix> H(new (std::nothrow) hssb_matrix());
  if (!H) HSSB_FAIL(HSSB_ERR_ALLOC, "plan_only_synthetic: out of memory");
  H->device = -1; H->shard_rank = shard_rank; H->n_shards = n_shards;
  H->synthetic = true; H->seed = seed; H->synth_rank = rank;
  nodes_synthetic(H.get(), n, leafsize, rank);
  int rc = plan_matrix(H.get());
  if (rc) return rc;
  fill_pool_host(H.get(), nullptr);
  *out = H.release();
  return HSSB_OK;
  });
}

int hssb_debug_counts(const hssb_matrix* h, int64_t* n_tasks, int64_t* n_phases, int64_t* pool_len) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_counts: NULL handle");
  if (n_tasks) *n_tasks = (int64_t)h->tasks_host.size();
  if (n_phases) *n_phases = (int64_t)(h->phases.size() + h->phases_t.size() + h->phases_u.size());  // forward, transposed, ULV solve
  if (pool_len) *pool_len = h->pool_len;
  return HSSB_OK;
  });
}

int hssb_debug_task(const hssb_matrix* h, int64_t i, hssb_task_t* o) {
  return guarded<int>([&]() -> int {
  if (!h || !o || i < 0 || i >= (int64_t)h->tasks_host.size()) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_task: bad argument");
  const GTask& t = h->tasks_host[(size_t)i];
  o->a0 = t.a0; o->a1 = t.a1; o->b0 = t.b0; o->b1 = t.b1; o->c = t.c;
  o->lda0 = t.lda0; o->lda1 = t.lda1; o->ldb0 = t.ldb0; o->ldb1 = t.ldb1; o->ldc = t.ldc;
  o->M = t.M; o->K0 = t.K0; o->K1 = t.K1; o->ta0 = t.ta0; o->ta1 = t.ta1;
  o->sb0 = t.sb0; o->sb1 = t.sb1; o->sc = t.sc; o->epilogue = t.epilogue;
  return HSSB_OK;
  });
}

int hssb_debug_phase(const hssb_matrix* h, int64_t i, hssb_phase_t* o) {
  return guarded<int>([&]() -> int {
  if (!h || !o || i < 0 || i >= (int64_t)(h->phases.size() + h->phases_t.size() + h->phases_u.size()))
    HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_phase: bad argument");
  const size_t nf = h->phases.size(), nt = h->phases_t.size();
  const int which = (size_t)i < nf ? 0 : (size_t)i < nf + nt ? 1 : 2;
  const Phase& p = which == 0 ? h->phases[(size_t)i] : which == 1 ? h->phases_t[(size_t)i - nf] : h->phases_u[(size_t)i - nf - nt];
  o->kind = p.kind; o->task0 = p.task0; o->ntasks = p.ntasks; o->maxM = p.maxM; o->level = p.level;
  o->top = p.top; o->fast = p.fast; o->transposed = which;
  o->xchg_zoff = h->xchg_zoff; o->xchg_slot_rows = h->xchg_slot_rows;
  return HSSB_OK;
  });
}

// Per-step device time of the persistent tree kernel (single shard, diagnostics): launches the tree
// kernel alone on the current workspace contents with CTA 0 recording its SM clock at every grid
// barrier.  us_out[j] = microseconds of step j (j-th covered phase of the forward plan).
int hssb_debug_tree_trace(hssb_matrix* h, int64_t nrhs, double* us_out, int cap) {
  return guarded<int>([&]() -> int {
  if (!h || !us_out || nrhs <= 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_tree_trace: bad argument");
  if (h->device < 0) HSSB_FAIL(HSSB_ERR_CUDA, "plan-only handle has no device");
  if (h->n_shards != 1) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_tree_trace: single-shard handles only");
  DeviceGuard dg(h->device);
  int rc = ensure_workspace(h, nrhs);
  if (rc) return rc;
  rc = ensure_tree_plan(h);
  if (rc) return rc;
  const TreePlan* tp = (const TreePlan*)h->tree_plan;
  const int ns = (int)tp->steps.size();
  if (ns == 0) return 0;
  if (cap < ns) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_tree_trace: need room for %d steps", ns);
  long long* trace = nullptr;
  HSSB_CUDA(cudaMalloc(&trace, (size_t)(ns + 1) * sizeof(long long)));
  CallParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.pool = h->pool_dev; cp.Z = h->z_dev; cp.F = h->f_dev; cp.nrhs = (int32_t)nrhs; cp.alpha = 1.0;
  XchgParams xq;
  memset(&xq, 0, sizeof(xq));
  std::vector<long long> t((size_t)ns + 1);
  for (int rep = 0; rep < 3 && !rc; ++rep) {  // the last repetition is reported (warm instruction cache / L2)
    rc = launch_tree(h, tp, 0, ns, cp, xq, h->stream, trace);
    if (!rc && cudaStreamSynchronize(h->stream) != cudaSuccess) { set_error("hssb_debug_tree_trace: %s", cudaGetErrorString(cudaGetLastError())); rc = HSSB_ERR_CUDA; }
  }
  if (!rc && cudaMemcpy(t.data(), trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) rc = HSSB_ERR_CUDA;
  cudaFree(trace);
  if (rc) return rc;
  int khz = 1965000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
  for (int j = 0; j < ns; ++j) us_out[j] = (double)(t[(size_t)j + 1] - t[(size_t)j]) / (khz * 1e-3);
  return ns;
  });
}

int hssb_debug_pool(const hssb_matrix* h, double* out, int64_t len) {
  return guarded<int>([&]() -> int {
  if (!h || !out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_pool: NULL argument");
  if (len < h->pool_len) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_pool: need %lld doubles", (long long)h->pool_len);
  if (!h->pool_host.empty()) { memcpy(out, h->pool_host.data(), (size_t)h->pool_len * 8); return HSSB_OK; }
  if (h->device < 0 || !h->pool_dev) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_pool: no pool image");
  DeviceGuard dg(h->device);
  HSSB_CUDA(cudaMemcpy(out, h->pool_dev, (size_t)h->pool_len * 8, cudaMemcpyDeviceToHost));
  return HSSB_OK;
  });
}

// Host image of the adjoint twin pool (what ensure_twin builds on the device), for plan-only and
// device handles alike: CPU tests run the FORWARD plan over it and must obtain A' X.
static BushPlan* debug_bush_plan(hssb_matrix* h, int mode) {
  if (!h || mode < 0 || mode > 1) { set_error("hssb_debug_bush: bad argument"); return nullptr; }
  BushPlan* bp = (BushPlan*)h->bush_plan[mode];
  if (!bp) {  // host part only: works on plan-only handles
    bp = new (std::nothrow) BushPlan();
    if (!bp) { set_error("hssb_debug_bush: out of memory"); return nullptr; }
    bush_plan_host(h, mode, h->bush_levels, h->bush_levels0, BUSH_SMEM_BUDGET, *bp);
    if (h->device >= 0) { delete bp; bp = nullptr; }  // device handles build (and upload) theirs at the first product
    else h->bush_plan[mode] = bp;
    if (!bp) {
      DeviceGuard dg(h->device);
      if (ensure_bush_plan(h, mode, 1) != HSSB_OK) return nullptr;
      bp = (BushPlan*)h->bush_plan[mode];
    }
  }
  if (bp && !bp->usable) { set_error("hssb_debug_bush: no bush plan: %s", bp->why.c_str()); return nullptr; }
  return bp;
}

int hssb_debug_bush_counts(hssb_matrix* h, int mode, int64_t* n_bush, int64_t* n_ops, int64_t* n_deps, int64_t* smem_doubles) {
  return guarded<int>([&]() -> int {
  const BushPlan* bp = debug_bush_plan(h, mode);
  if (!bp) return HSSB_ERR_STATE;
  if (n_bush) *n_bush = bp->nbush;
  if (n_ops) *n_ops = (int64_t)bp->ops.size();
  if (n_deps) *n_deps = (int64_t)bp->deps.size();
  if (smem_doubles) *smem_doubles = bp->smem_doubles;
  return HSSB_OK;
  });
}

int hssb_debug_bush_op(hssb_matrix* h, int mode, int64_t i, hssb_bush_op_t* o) {
  return guarded<int>([&]() -> int {
  const BushPlan* bp = debug_bush_plan(h, mode);
  if (!bp) return HSSB_ERR_STATE;
  if (!o || i < 0 || i >= (int64_t)bp->ops.size()) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_bush_op: bad argument");
  const BushOp& b = bp->ops[(size_t)i];
  o->task = bp->task0 + bp->op_task[(size_t)i]; o->bush = bp->op_bush[(size_t)i]; o->level = bp->op_level[(size_t)i];
  o->m0 = b.m0; o->mr = b.mr; o->s0 = b.s0; o->s1 = b.s1; o->sc = b.sc_off;
  o->lds0 = b.lds0; o->lds1 = b.lds1; o->ldsc = b.ldsc; o->to_global = b.to_global;
  o->sa0 = b.sa0; o->sa1 = b.sa1;
  return HSSB_OK;
  });
}

int64_t hssb_debug_bush_trace(hssb_matrix* h, int mode, uint64_t* out, int64_t cap_items) {
  return guarded<int64_t>([&]() -> int64_t {
  if (!h || mode < 0 || mode > 1 || h->device < 0) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_bush_trace: bad argument");
  if (!out) { h->bush_trace = true; h->bush_probe_item = (int)cap_items; return 0; }   // cap_items doubles as the item to probe
  DeviceGuard dg(h->device);
  const BushPlan* bp = (const BushPlan*)h->bush_plan[mode];
  if (!bp || !bp->trace_dev) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_bush_trace: nothing recorded (switch it on, then run a product)");
  if (h->stream) HSSB_CUDA(cudaStreamSynchronize(h->stream));
  HSSB_CUDA(cudaDeviceSynchronize());
  const int64_t all = (int64_t)bp->nbush * bp->sync_cols;
  const int64_t items = std::min<int64_t>(cap_items, all);
  HSSB_CUDA(cudaMemcpy(out, bp->trace_dev, (size_t)items * B_TRACE * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (cap_items >= all + (B_PROBE_LEVELS * B_WARPS * 8 + B_TRACE - 1) / B_TRACE)   // room for the probe block behind the items
    HSSB_CUDA(cudaMemcpy(out + all * B_TRACE, bp->trace_dev + all * B_TRACE, (size_t)B_PROBE_LEVELS * B_WARPS * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return items;
  });
}

int64_t hssb_debug_bush_stage(hssb_matrix* h, int mode, int64_t i, hssb_bush_stage_t* o) {
  return guarded<int64_t>([&]() -> int64_t {
  const BushPlan* bp = debug_bush_plan(h, mode);
  if (!bp) return HSSB_ERR_STATE;
  if (!o) return (int64_t)bp->stages.size();
  if (i < 0 || i >= (int64_t)bp->stages.size()) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_bush_stage: bad argument");
  const BushStage& st = bp->stages[(size_t)i];
  o->bush = bp->stage_bush[(size_t)i]; o->kind = st.kind; o->src = st.src; o->dst = st.dst; o->count = st.count; o->ld = st.ld;
  o->level = 0;
  return (int64_t)bp->stages.size();
  });
}

int64_t hssb_debug_bush_deps(hssb_matrix* h, int mode, int64_t b, int64_t* out, int64_t cap) {
  return guarded<int64_t>([&]() -> int64_t {
  const BushPlan* bp = debug_bush_plan(h, mode);
  if (!bp) return HSSB_ERR_STATE;
  if (b < 0 || b >= bp->nbush) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_bush_deps: bad argument");
  const BushHdr& hd = bp->hdr[(size_t)b];
  for (int64_t d = 0; d < hd.ndeps && d < cap && out; ++d) out[d] = bp->deps[(size_t)(hd.dep0 + d)];
  return hd.ndeps;
  });
}

int hssb_debug_pool_t(const hssb_matrix* h, double* out, int64_t len) {
  return guarded<int>([&]() -> int {
  if (!h || !out) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_pool_t: NULL argument");
  if (len < h->pool_len) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_pool_t: need %lld doubles", (long long)h->pool_len);
  std::vector<TwinBlock> tb;
  if (!twin_blocks(h, tb)) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_pool_t: the tree is not uniform, there is no adjoint twin");
  if (h->device >= 0 && h->pool_t_dev) {
    DeviceGuard dg(h->device);
    HSSB_CUDA(cudaMemcpy(out, h->pool_t_dev, (size_t)h->pool_len * 8, cudaMemcpyDeviceToHost));
    return HSSB_OK;
  }
  if (h->pool_host.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_pool_t: no twin on the device yet and no host pool image");
  memset(out, 0, (size_t)h->pool_len * 8);
  for (const TwinBlock& b : tb)
    for (int64_t c = 0; c < b.cols; ++c)
      for (int64_t r = 0; r < b.rows; ++r) out[b.dst + r * b.ld_dst + c] = h->pool_host[(size_t)(b.src + c * b.ld_src + r)];
  return HSSB_OK;
  });
}

// ULV test hooks: factorise a plan-only handle on the HOST with the same node routine the device
// kernel runs (single-thread team), and expose the factor pool so that the numpy plan interpreter can
// run the solve's task table (phases with transposed == 2) over it.
int hssb_debug_ulv_factor_host(hssb_matrix* h) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_ulv_factor_host: NULL handle");
  if (h->ulv.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_ulv_factor_host: %s", h->ulv_why.empty() ? "no ULV plan" : h->ulv_why.c_str());
  if (h->pool_host.empty()) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_ulv_factor_host: plan-only handles only");
  if (int rc = ulv_factor_host(h)) return rc;
  return HSSB_OK;
  });
}

int hssb_debug_ulv_pool(const hssb_matrix* h, double* out, int64_t len) {
  return guarded<int>([&]() -> int {
  if (!h) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_ulv_pool: NULL handle");
  if (h->ulv.empty() || !h->ulv_factored) HSSB_FAIL(HSSB_ERR_STATE, "hssb_debug_ulv_pool: not factorised");
  if (!out || len < h->ulv_pool_len) HSSB_FAIL(HSSB_ERR_ARG, "hssb_debug_ulv_pool: need %lld doubles", (long long)h->ulv_pool_len);
  if (!h->ulv_pool_host.empty()) { memcpy(out, h->ulv_pool_host.data(), (size_t)h->ulv_pool_len * 8); return HSSB_OK; }
  DeviceGuard dg(h->device);
  HSSB_CUDA(cudaMemcpy(out, h->ulv_pool_dev, (size_t)h->ulv_pool_len * 8, cudaMemcpyDeviceToHost));
  return HSSB_OK;
  });
}

}  // extern "C"

#include "hssb_group.h"
