// hssb_internal.h — shared host/device declarations of the hssb200 library.
// Not part of the public ABI (that is include/hssb200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/hssb200.h"

#ifndef HSSB_PDL_DEFAULT
#define HSSB_PDL_DEFAULT 1
#endif

namespace hssb {

// ---------------------------------------------------------------- errors ---
void set_error(const char* fmt, ...);
#define HSSB_FAIL(code, ...)      \
  do {                            \
    ::hssb::set_error(__VA_ARGS__); \
    return (code);                \
  } while (0)
#define HSSB_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::hssb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                        __LINE__, cudaGetErrorString(_e));                           \
      return HSSB_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

// ------------------------------------------------- block kinds / sources ---
enum BlockKind : int { BK_D = 0, BK_U, BK_V, BK_B12, BK_B21, BK_R, BK_W, BK_COUNT };
enum OperandSrc : int { SRC_X = 0, SRC_Z = 1, SRC_F = 2, SRC_Y = 3 };

// One task of the generic (any-shape) kernel:
//   C[M x N] = alpha * (op(A0)[M x K0] * B0[K0 x N] + op(A1)[M x K1] * B1[K1 x N]) + beta * C
// A0/A1 live in the generator pool; B0/B1/C are the user's X / Y or a block of the
// Z / F workspace.  Every step of src/matmul.jl:32-62 has this form:
//   leaf up    Z = V' X                     (matmul.jl:34)
//   merge      Z = W1' Z1 + W2' Z2          (matmul.jl:39)
//   translate  F1 = B12 Z2 + R1 F           (matmul.jl:52-56)
//   leaf down  Y = a (D X + U F) + b Y      (matmul.jl:46-47)
struct GTask {
  int64_t a0, a1;     // pool offsets (doubles)
  int64_t b0, b1, c;  // X/Y: first row; Z/F: workspace row offset (element offset = row * nrhs)
  int32_t lda0, lda1;
  int32_t ldb0, ldb1, ldc;  // leading dimensions of workspace operands (ignored for X/Y)
  int32_t M, K0, K1;
  uint8_t ta0, ta1;         // 1: op(A) = A'
  uint8_t sb0, sb1, sc;     // OperandSrc
  uint8_t epilogue;         // 1: C = alpha*acc + beta*C, 0: C = acc
  uint8_t pad[2];
};

struct CallParams {
  const double* pool;
  const double* X;
  double* Y;
  double* Z;
  double* F;
  int64_t ldx, ldy;
  int32_t nrhs;
  double alpha, beta;
  int32_t trans;  // 0: Y = A X, 1: Y = A' X (transposed task table), 2: ULV solve Y = A \ X (phases_u over the factor pool)
  int32_t debug;  // HSSB_OPT_DEBUG bits: 1 = leaf kernels compute without waiting for data, 2 = move data without computing
};

// ------------------------------------------------------------- host tree ---
struct HostBlock {  // where a generator block comes from
  std::vector<double> data;  // compact copy, ld = rows (empty for synthetic / absent blocks)
  int64_t rows = 0, cols = 0;
};

struct Node {
  int64_t left = -1, right = -1, parent = -1;
  int32_t depth = 0, height = 0;
  bool leaf = false, remote = false, top = false;  // top: above the shard cut (replicated)
  bool local = true;                               // inside this shard's subtree
  int64_t row0 = 0, col0 = 0, m = 0, n = 0;
  int64_t kr = 0, kw = 0;        // gensize (hssmatrix.jl:254-262); 0 at the root
  uint64_t heap_id = 1;          // synthetic generator stream id
  // packed generator blocks (pool offsets in doubles, -1 = absent)
  int64_t off[BK_COUNT] = {-1, -1, -1, -1, -1, -1, -1};
  int32_t ld[BK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  int64_t rows[BK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  int64_t cols[BK_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  int64_t zoff = -1, foff = -1;  // workspace row offsets
  int32_t ldz = 0, ldf = 0;
};

// One block of the adjoint twin pool: twin[dst] (cols x rows, ld_dst) = pool[src] (rows x cols, ld_src) transposed.
struct TwinBlock {
  int64_t src, dst;
  int32_t rows, cols, ld_src, ld_dst;
};

// ULV solver (src/ulvfactor.jl): shapes and offsets of one node of the implicit factorisation.
// "Incoming" = the reduced problem that reaches the node: the leaf's own D/U/V, or the merge of
// the two children's reduced blocks (ulvfactor.jl:74-79).
struct UlvNode {
  int32_t m_in = 0, n_in = 0;   // rows / columns of the incoming diagonal block
  int32_t kr = 0, kw = 0;       // generator ranks (columns of U / V)
  int32_t k = 0;                // rows handed to the parent: kr if compressible, else m_in
  int32_t mk = 0;               // unknowns eliminated at this node (m_in - k, 0 if not compressible)
  int32_t n_out = 0;            // columns handed to the parent (n_in - mk)
  int32_t is_root = 0, is_leaf = 0;
  int32_t left = -1, right = -1;
  int32_t k1 = 0, kr1 = 0, kw1 = 0, no1 = 0, k2 = 0, kr2 = 0, kw2 = 0, no2 = 0;  // children
  // primary pool blocks (pool offsets / leading dimensions; V and W are stored transposed)
  int64_t D = -1, U = -1, V = -1, B12 = -1, B21 = -1, R1 = -1, R2 = -1, W1 = -1, W2 = -1;
  int32_t ldD = 0, ldU = 0, ldV = 0, ldB12 = 0, ldB21 = 0, ldR1 = 0, ldR2 = 0, ldW1 = 0, ldW2 = 0;
  // factor pool blocks, column-major:
  //   az[s]: mk rows, ac[s]: k + kw rows (root: n_in rows); leaf: one block of m_in columns (s = 0),
  //   branch: one block per child s of k_s + kw_s columns.  pta: n_in x mk, ptb: n_in x n_out.
  int64_t az[2] = {-1, -1}, ac[2] = {-1, -1}, pta = -1, ptb = -1;
  int32_t ld_az = 2, ld_ac = 2, ld_pt = 2;
  // reduced generators handed to the parent (scratch that lives during the factorisation only)
  int64_t rD = -1, rU = -1, rV = -1;  // k x n_out, k x kr, n_out x kw (leading dimension = rows)
  // solve workspaces (row offsets): Z space holds zloc (mk rows) and c = [b; u] (k + kw rows),
  // F space holds t (the n_out trailing unknowns, written by the parent)
  int64_t zloc = -1, c = -1, t = -1;
  int32_t ld_zloc = 2, ld_c = 2, ld_t = 2;
  // "fast form" (uniform trees, HSSB_OPT_ULV_FAST): blocks shaped and padded (+4) for the fixed-shape
  // kernels of the product.  Leaf: g = P'[:, :mk] T1 (n_in x m_in), so that the leaf output is
  // Z[cols] = g b + ptb t, the shape of the product's leaf-down step, and zloc is never formed.
  // Branch: P' split by rows into one pair of blocks per child (pta_c[s]: no_s x mk, ptb_c[s]: no_s x n_out).
  int64_t g = -1, pta_c[2] = {-1, -1}, ptb_c[2] = {-1, -1};
  int32_t ld_g = 2, ld_ptc[2] = {2, 2};
};

enum PhaseKind : int { PH_LEAF_UP = 0, PH_MERGE, PH_EXCHANGE, PH_TRANSLATE, PH_LEAF_DOWN, PH_XCHG_ACK };

struct Phase {
  int kind;
  int64_t task0 = 0, ntasks = 0;  // range in the task array
  int32_t maxM = 0;
  int32_t level = 0;              // height (merge) or depth (translate)
  bool top = false;               // replicated top-tree phase
  int fast = 0;                   // fixed-shape kernel id (0 = generic)
  int32_t fast_m = 0, fast_r = 0; // shape the fixed-shape kernel is instantiated for (0 = the tree's leaf size / rank)
};

}  // namespace hssb

namespace hssb {
// default of HSSB_OPT_PDL: the environment variable HSSB_PDL (0 / 1) overrides the built-in default for every new handle
inline int default_pdl() {
  static const int v = [] { const char* e = getenv("HSSB_PDL"); const int x = e ? atoi(e) : HSSB_PDL_DEFAULT; return x < 0 ? 0 : (x > 15 ? 15 : x); }();
  return v;
}
}  // namespace hssb

// The opaque handle of the public ABI.
struct hssb_matrix {
  int device = 0;
  int shard_rank = 0, n_shards = 1;
  std::vector<hssb::Node> nodes;  // BFS order, root = 0
  std::vector<int64_t> leaves;    // local leaves, left to right
  std::vector<hssb::Phase> phases;
  std::vector<hssb::Phase> phases_t;  // Y = A' X on the same generators (single shard)
  // ULV solver (hssb_solve): plan built with the product plan, factor pool filled by hssb_ulv_factor
  std::vector<hssb::Phase> phases_u;
  std::vector<hssb::UlvNode> ulv;     // parallel to `nodes`; empty if the solver does not apply
  std::string ulv_why;                // why it does not apply
  int64_t ulv_pool_len = 0, ulv_red_len = 0, ulv_z_rows = 0, ulv_f_rows = 0;
  int64_t ulv_flops_per_rhs = 0;
  int32_t ulv_MI = 0, ulv_NI = 0, ulv_KR = 0, ulv_KW = 0;  // scratch maxima
  double* ulv_pool_dev = nullptr;
  double* ulv_pool_t_dev = nullptr;   // factors of A' (from the adjoint twin pool): hssb_solve_t
  bool ulv_t_factored = false;
  std::vector<double> ulv_pool_host;  // plan-only handles: factorised on the host by the test hook
  bool ulv_factored = false;
  int64_t ulv_last_factor_us = 0;     // device time of the level launches of the last factorisation (HSSB_OPT_LAST_FACTOR_US)
  bool ulv_fast_form = true;          // HSSB_OPT_ULV_FAST (default on since it ran green on hardware: solve 2.60 -> 1.79 ms on config 3)
  bool ulv_ff = false;                // ... and the tree qualifies: the ULV plan is in fast form
  int64_t ulv_task0 = -1;             // first ULV task in tasks_host (the ULV plan can be rebuilt)
  std::vector<hssb::GTask> tasks_host;
  hssb::GTask* tasks_dev = nullptr;
  double* pool_dev = nullptr;
  std::vector<double> pool_host;  // plan-only handles (CPU tests): host image of the pool
  // Adjoint twin (uniform trees): a second pool of identical layout holding the generators of A'
  // (D', U <-> V, B12 <-> B21', R <-> W), built on the device at the first transposed product so
  // that A' X runs the forward plan and the fixed-shape kernels.  HSSB_OPT_ADJOINT_TWIN.
  double* pool_t_dev = nullptr;
  bool adjoint_twin = true;
  bool twin_unavailable = false;  // not eligible or did not fit: the any-shape transposed plan is used
  int64_t pool_len = 0;  // doubles
  int64_t gen_elems = 0, flops_per_rhs = 0;
  int64_t z_rows = 0, f_rows = 0;
  int64_t ws_nrhs = 0;  // workspace capacity in columns
  bool ws_ulv = false;  // workspaces sized for the ULV solve as well
  double* z_dev = nullptr;
  double* f_dev = nullptr;
  // staging for the host-pointer entry
  double* x_stage = nullptr;
  double* y_stage = nullptr;
  int64_t stage_nrhs = 0;
  cudaStream_t stream = nullptr;
  // host entry: column-block pipeline (H2D | product | D2H)
  static constexpr int MAX_BLOCKS = 16;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_in[MAX_BLOCKS] = {}, ev_done[MAX_BLOCKS] = {};
  int64_t pipeline_cols = 0;  // 0 = automatic
  // pageable caller memory: pinned slot rings + worker threads (hssb_hostpipe.h).  0 never, 1 automatic, 2 always
  int host_bounce = 1;
  int last_bounce = 0;  // what the last host call did: bit 0 = X went through the ring, bit 1 = Y
  void* bounce = nullptr;
  int64_t launches = 0;
  // global / local shape
  int64_t m = 0, n = 0, local_m = 0, local_n = 0, local_row0 = 0, local_col0 = 0;
  int64_t depth = 0, max_leaf_m = 0, max_leaf_n = 0, max_rank = 0;
  bool uniform = false;
  bool padded = false;  // pool / workspace blocks carry the +4 padded leading dimension (TMA-ready images)
  int64_t uni_m = 0, uni_r = 0;  // leaf size / rank when uniform
  // exchange
  int64_t xchg_zoff = -1, xchg_slot_rows = 0;  // all-gather buffer = P slots of slot_rows x nrhs
  void* nccl_comm = nullptr;
  // peer-memory exchange (hssb_xchg_export / hssb_xchg_import): every rank maps the Z workspace and
  // the flag block of every other rank (CUDA IPC) and pushes its subtree-root Z block over NVLink
  static constexpr int MAX_PEERS = 16;
  bool peer_xchg = false;
  bool xchg_exported = false;
  bool peer_inprocess = false;  // peer tables point at other handles of this process (hssb_group), not at IPC mappings
  double* peer_z[MAX_PEERS] = {};
  unsigned long long* peer_flags[MAX_PEERS] = {};  // [0,P): data flags, [P,2P): ack flags, [2P]: epoch, [2P+1]: ticket
  unsigned long long* my_flags = nullptr;
  // options
  bool force_generic = false, use_graph = false, profile = false;
  int debug_mode = 0;
  bool in_host_call = false;
  bool prepare_only = false;  // run_graph captures, instantiates and uploads but does not launch
  std::vector<cudaEvent_t> prof_events;  // HSSB_OPT_PROFILE: one event between consecutive phases
  int64_t prof_nrhs = 0;
  int prof_mode = 0;  // which plan the last profiled call ran (0 product, 1 transposed task table, 2 ULV solve)
  // synthetic
  bool synthetic = false;
  uint64_t seed = 0;
  int64_t synth_rank = 0;
  // CUDA-graph cache (HSSB_OPT_USE_GRAPH): one instantiated graph per distinct call signature
  struct GraphSlot {
    hssb::CallParams cp;
    cudaGraphExec_t exec = nullptr;
    int64_t kernels = 0;
  };
  std::vector<GraphSlot> graphs;
  int graph_miss_streak = 0;            // consecutive calls whose signature was not in the cache
  hssb::CallParams graph_plain_cp;      // last signature that was launched plainly instead of captured
  bool graph_plain_valid = false;
  // fixed-shape kernel state (hssb_fast.cuh)
  void* fast_state = nullptr;
  // persistent tree kernel (hssb_tree.cuh): 1 = all merge / translate levels (and the peer exchange) in one
  // cooperative launch, 0 = one launch per level, 2 = as 1 without the cooperative attribute (diagnostics)
  int tree_kernel = 0;
  // leaf kernels: 2 = second generation (hssb_leaf2.cuh, default), 1 = first generation (hssb_fast.cuh, cross-check),
  // 3 = second generation with longer chunks where instantiated (measurement)
  int leaf_kernel = 2;
  // HSSB_OPT_LEAF_FUSION: 1 = the "X once" variant (hssb_leafx.cuh): leaf-up forms D X and V' X in one pass over X and
  // parks alpha D X + beta Y in Y, leaf-down adds alpha U F.  Measured slower than the two-pass default.
  int leaf_fusion = 0;
  // HSSB_OPT_PDL (bits): 1 (default) the node kernels of the level schedule are launched with programmatic stream serialisation
  // where it pays (hssb_fast.cuh: launch_k), 2 every persistent node kernel, 4 the leaf kernels (measured slower), 0 plain stream order
  int pdl = hssb::default_pdl();
  // HSSB_OPT_FLOW_KERNEL: any-shape plans of single-shard handles run as ONE persistent dataflow kernel (hssb_flow.cuh)
  int flow_kernel = 2;   // 0 never, 1 always, 2 automatic: for plain launches; under CUDA-graph capture one launch per level (with HSSB_OPT_PDL) is faster
  bool capturing = false;  // run_graph is capturing the schedule
  void* flow_plan[2] = {nullptr, nullptr};  // [0] Y = A X, [1] Y = A' X on the any-shape task table
  // HSSB_OPT_BUSH_KERNEL: the merge / translate levels of small any-shape trees (64-row leaves, ranks <= 64: what a compression
  // produces) as ONE launch whose items are whole bushes of the tree (hssb_bush.cuh).  0 off (default: measured slower than the
  // dataflow kernel, profiles/bush_kernel_r02.txt), 1 eligible trees, 2 whenever a plan exists
  int bush_kernel = 0;
  int bush_levels = 2, bush_levels0 = 1;    // HSSB_OPT_BUSH_LEVELS: levels per bush / merge levels of the bush that holds the root
  void* bush_plan[2] = {nullptr, nullptr};
  int bush_probe_item = 0;                  // ... and cycle stamps inside the ops of this item
  bool bush_trace = false;                  // hssb_debug_bush_trace: the kernel records a timeline per item
  void* tree_plan = nullptr;
};
