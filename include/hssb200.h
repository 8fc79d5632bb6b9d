/* hssb200.h — C ABI of the B200-native HSS x dense product.
 *
 * Drop-in boundary for ONE path of bonevbs/HssMatrices.jl (v0.1.6):
 *     *(hssA::HssMatrix, B::AbstractMatrix)            src/matmul.jl:13
 *     *(hssA::HssMatrix, x::AbstractVector)            src/matmul.jl:15
 *     mul!(C, hssA::HssMatrix, B, alpha, beta)         src/matmul.jl:18-28
 *       _matmatup   (post-order upsweep)               src/matmul.jl:32-42
 *       _matmatdown! (pre-order downsweep)             src/matmul.jl:44-62
 * The reference has no FFI for this path (it is pure Julia on OpenBLAS), so the
 * entry points below are what a Julia `ccall` binding for it needs; the binding
 * itself is hssmatrices.jl_b200/julia/HssMatricesB200.jl and is walked through
 * in INTEGRATION.md.
 *
 * Conventions: extern "C"; plain pointers and sizes; all matrices Float64,
 * column-major with an explicit leading dimension (Julia `stride(A,2)`); every
 * function returns 0 on success or a negative hssb_status and never throws;
 * hssb_last_error() returns a thread-local message for the last failure.
 * There is NO CPU fallback: every compute entry point fails with
 * HSSB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef HSSB200_H
#define HSSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSSB_VERSION 100 /* 0.1.0 */

typedef enum hssb_status {
  HSSB_OK = 0,
  HSSB_ERR_ARG = -1,      /* bad pointer / size / id                                   */
  HSSB_ERR_DIM = -2,      /* Julia DimensionMismatch: matmul.jl:19-20, hssmatrix.jl:41-42,58-59 */
  HSSB_ERR_CUDA = -3,     /* CUDA runtime error or no usable device                    */
  HSSB_ERR_ALLOC = -4,    /* host or device allocation failed                          */
  HSSB_ERR_STATE = -5,    /* call not valid in this state (e.g. comm not initialised)  */
  HSSB_ERR_COMM = -6,     /* NCCL error / NCCL not loadable                            */
  HSSB_ERR_SINGULAR = -7  /* Julia SingularException: hssb_ulv_factor / hssb_solve met a zero pivot (ulvfactor.jl:48, :83) */
} hssb_status;

typedef struct hssb_builder hssb_builder; /* host-side packer front end */
typedef struct hssb_matrix hssb_matrix;   /* packed, device-resident HSS matrix */

/* ---- library ---------------------------------------------------------- */
int hssb_version(void);
const char* hssb_last_error(void);
/* Number of CUDA devices with compute capability 10.x (0 if none / no driver). */
int hssb_device_count(void);

/* ---- packer: flattens the recursive HssMatrix (src/hssmatrix.jl:11-68) --- */
/* The caller walks its tree in post-order and registers every node; block data
 * is copied out during the call, so Julia only needs GC.@preserve per call.
 * Node ids are returned (>= 0) or a negative hssb_status.                   */
int hssb_builder_create(hssb_builder** out);
void hssb_builder_destroy(hssb_builder* b);

/* Leaf node: HssMatrix(D, U, V) (hssmatrix.jl:40-44).  D is m x n, U is m x kr,
 * V is n x kw; kr/kw may be 0 (then U/V may be NULL), as for HssMatrix(D)
 * (hssmatrix.jl:36-39) and hss_blkdiag (compression.jl:447-475).             */
int64_t hssb_builder_add_leaf(hssb_builder* b, int64_t m, int64_t n, int64_t kr, int64_t kw,
                              const double* D, int64_t ldd, const double* U, int64_t ldu,
                              const double* V, int64_t ldv);

/* Branch node: HssMatrix(A11, A22, B12, B21, R1, W1, R2, W2) (hssmatrix.jl:56-67).
 * B12 is kr(left) x kw(right), B21 is kr(right) x kw(left); R1/R2 are
 * kr(child) x kr, W1/W2 are kw(child) x kw where (kr, kw) = gensize of this
 * node (hssmatrix.jl:254-262).  For a root-style node (hssmatrix.jl:46-55) pass
 * kr = kw = 0 and NULL translators.  Dimension violations return HSSB_ERR_DIM
 * (the checks of hssmatrix.jl:308-322).                                      */
int64_t hssb_builder_add_branch(hssb_builder* b, int64_t left, int64_t right, int64_t kr, int64_t kw,
                                const double* B12, int64_t ldb12, const double* B21, int64_t ldb21,
                                const double* R1, int64_t ldr1, const double* W1, int64_t ldw1,
                                const double* R2, int64_t ldr2, const double* W2, int64_t ldw2);

/* Placeholder for a subtree whose generators live on another GPU (multi-GPU
 * subtree sharding): only its size and gensize are known here. */
int64_t hssb_builder_add_remote(hssb_builder* b, int64_t m, int64_t n, int64_t kr, int64_t kw);

/* Pack the tree under `root` into level-ordered device arrays on `device`.
 * The node is treated as root exactly like rooted() (hssmatrix.jl:266; used at
 * matmul.jl:24): its own R/W are ignored.  `shard_rank`/`n_shards` describe
 * subtree sharding: with n_shards == 1 the tree must contain no remote nodes;
 * with n_shards = P > 1 the tree must contain exactly P-1 remote placeholders
 * and the local subtree must be the shard_rank-th of the P subtrees (left to
 * right).  The builder can be destroyed afterwards.                          */
int hssb_builder_finalize(hssb_builder* b, int64_t root, int device, int shard_rank, int n_shards,
                          hssb_matrix** out);

/* Synthetic random-generator HSS matrix generated ON THE DEVICE (BASELINE.json
 * configs 3-5): bisection tree (clustertree.jl:27-35) on n with `leafsize`,
 * every rank = `rank`, counter-based generator specified in
 * oracle/hss_oracle.py (synth_*), bit-identical to that host twin.  With
 * n_shards = P > 1 only the shard_rank-th depth-log2(P) subtree plus the
 * replicated top tree is generated.                                          */
int hssb_create_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, int device,
                          int shard_rank, int n_shards, hssb_matrix** out);
/* Fill rows [row0, row0+rows) of the n x nrhs synthetic right-hand side into
 * device memory dX (leading dimension ldx).                                   */
int hssb_synthetic_rhs(uint64_t seed, int64_t n, int64_t nrhs, int64_t row0, int64_t rows,
                       double* dX, int64_t ldx, int device, void* stream);

/* The packed format as a file (checkpoint / fixture exchange; the reference has no serialisation):
 * tree shape + level-ordered pool, versioned.  hssb_load with device < 0 gives a host-only handle
 * (inspection, hssb_get_block), with device >= 0 a ready-to-multiply device-resident matrix.       */
int hssb_save(const hssb_matrix* h, const char* path);
int hssb_load(const char* path, int device, hssb_matrix** out);

int hssb_destroy(hssb_matrix* h);

/* ---- queries ---------------------------------------------------------- */
typedef struct hssb_info_t {
  int64_t m, n;                 /* global size(hssA) (hssmatrix.jl:94)                   */
  int64_t local_m, local_n;     /* rows of Y / X owned by this shard                     */
  int64_t local_row0, local_col0;
  int64_t n_nodes, n_leaves, depth;
  int64_t max_leaf_m, max_leaf_n, max_rank;
  int64_t pool_bytes;           /* device bytes of packed generators (incl. padding)     */
  int64_t gen_elems;            /* local generator doubles, unpadded (algorithmic)       */
  int64_t flops_per_rhs;        /* local algorithmic flops per right-hand-side column    */
  int64_t z_rows, f_rows;       /* workspace rows (x nrhs x 8 bytes each)                */
  int32_t shard_rank, n_shards;
  int32_t device;
  int32_t uniform;              /* 1 if the fast fixed-shape kernels apply               */
} hssb_info_t;
int hssb_info(const hssb_matrix* h, hssb_info_t* out);

/* Read back one generator block (test/debug): kind 0..6 = D,U,V,B12,B21,R,W of
 * node `node` (ids in BFS order; see hssb_node_info).  out is rows x cols,
 * column-major, ld = rows.                                                    */
typedef struct hssb_node_t {
  int64_t left, right, parent;  /* -1 if none */
  int64_t depth, is_leaf, is_remote;
  int64_t row0, m, col0, n, kr, kw;
} hssb_node_t;
int hssb_node_info(const hssb_matrix* h, int64_t node, hssb_node_t* out);
int hssb_get_block(const hssb_matrix* h, int64_t node, int kind, double* out, int64_t out_len);

/* ---- the product ------------------------------------------------------ */
/* Pre-size the Z/F workspaces for up to max_nrhs columns (otherwise grown on
 * demand inside the first call).                                             */
int hssb_reserve(hssb_matrix* h, int64_t max_nrhs);

/* mul!(C, hssA, B, alpha, beta) with HOST pointers (matmul.jl:18): copies the
 * local rows of X to the device, runs the product, copies Y back, synchronous.
 * beta == 0 never reads Y (matmul.jl:13 passes uninitialised memory).
 * HSSB_ERR_DIM mirrors the DimensionMismatch of matmul.jl:19-20: rows_x must be
 * local_n and rows_y local_m.                                                */
int hssb_matmul(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs,
                const double* X, int64_t ldx, double* Y, int64_t ldy, double alpha, double beta);

/* Same with DEVICE pointers, asynchronous on `stream` (a cudaStream_t; NULL =
 * the CUDA default stream).  This is the timed entry.
 * CONTRACT for every *_dev entry (product, transposed product, solve): a handle owns ONE set of Z / F
 * workspaces, one task table and one graph cache, so
 *   - calls on one handle must be ordered: issue them on one stream, or make the next call's stream wait
 *     for the previous call (event) -- two calls in flight on different streams race on the workspaces;
 *   - a call with more right-hand sides than any before it (or the first transposed product / solve) may
 *     reallocate workspaces: let earlier asynchronous calls finish first (hssb_reserve up front avoids it);
 *   - X and Y (B and Z for the solve) must not overlap: the leaf phases read X while they write Y.
 * One handle = one owner thread at a time; different handles are independent.            */
int hssb_matmul_dev(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs,
                    const double* dX, int64_t ldx, double* dY, int64_t ldy, double alpha, double beta,
                    void* stream);
/* Transposed product Y = alpha * A' * X + beta * Y: what `*(A::AbstractMatrix, hssB)` (matmul.jl:14)
 * and `hssA' * X` need.  The reference builds a full copied adjoint HssMatrix (hssmatrix.jl:165-171)
 * on EVERY such call.  Here:
 *  - uniform trees (the fixed-shape kernel shapes): the first transposed product builds, on the
 *    device, an adjoint twin pool of identical layout (D', U <-> V, B12 <-> B21', R <-> W) and A' X
 *    then runs the forward plan and the DMMA kernels over it, sharded handles included.  Costs a
 *    second pool of device memory; HSSB_OPT_ADJOINT_TWIN = 0 turns it off, and it is skipped when
 *    the device cannot hold it;
 *  - otherwise: a second task table over the SAME pool on the any-shape kernel (single shard only).
 * rows_x must be size(A,1), rows_y size(A,2).                                                     */
int hssb_matmul_t(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs,
                  const double* X, int64_t ldx, double* Y, int64_t ldy, double alpha, double beta);
int hssb_matmul_t_dev(hssb_matrix* h, int64_t rows_y, int64_t rows_x, int64_t nrhs,
                      const double* dX, int64_t ldx, double* dY, int64_t ldy, double alpha, double beta,
                      void* stream);
int hssb_sync(hssb_matrix* h);

/* ---- the solver: hssA \ B --------------------------------------------- */
/* `\(hssA::HssMatrix, B::Matrix)` (hssmatrix.jl:234) = ulvfactsolve (ulvfactor.jl:10-19): implicit
 * ULV factorisation (Chandrasekaran, Gu, Pals 2006).  The reference factorises and solves in one
 * recursive pass on EVERY call; here hssb_ulv_factor runs the factorisation once on the device
 * (Householder QR / LQ per node, level by level, folded into explicit per-node matrices in a second
 * level-ordered pool) and hssb_solve applies it to any number of right-hand sides with the level
 * schedule and kernels of the product.  Square matrices, single shard.  hssb_solve factorises on
 * first use; call hssb_ulv_factor to do it ahead of time.  B is rows x nrhs, Z (the solution) too.  */
int hssb_ulv_factor(hssb_matrix* h);
int hssb_solve(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* B, int64_t ldb, double* Z, int64_t ldz);
int hssb_solve_dev(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* dB, int64_t ldb, double* dZ,
                   int64_t ldz, void* stream);
typedef struct hssb_ulv_info_t {
  int64_t supported;     /* 1 if hssb_solve applies to this handle                              */
  int64_t factored;      /* 1 once the factor pool exists                                       */
  int64_t pool_bytes;    /* device bytes of the factor pool                                     */
  int64_t flops_per_rhs; /* flops of one solve per right-hand-side column (factors x 8 B = bytes read / 2) */
  int64_t z_rows, f_rows; /* workspace rows of the solve                                        */
} hssb_ulv_info_t;
/* A' \ B, i.e. `/(A, hssB) = ulvfactsolve(hssB', collect(A'))'` (hssmatrix.jl:236) without building the adjoint
 * matrix: on a uniform tree A' has the shapes of A, so the solve plan is shared and only a second factor pool is
 * computed, from the adjoint twin pool (first call factorises; HSSB_ERR_STATE on trees without a twin pool: pack the
 * adjoint there and use hssb_solve).  Z = A' \ B, so A / hssB = (hssb_solve_t(hssB, A'))'.                      */
int hssb_solve_t(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* B, int64_t ldb, double* Z, int64_t ldz);
int hssb_solve_t_dev(hssb_matrix* h, int64_t rows, int64_t nrhs, const double* dB, int64_t ldb, double* dZ, int64_t ldz,
                     void* stream);
int hssb_ulv_info(const hssb_matrix* h, hssb_ulv_info_t* out);

/* Options. */
#define HSSB_OPT_FORCE_GENERIC 1 /* 1: never use the fixed-shape DMMA kernels (debug/parity)  */
#define HSSB_OPT_USE_GRAPH 2     /* 1: replay the level schedule as a CUDA graph              */
#define HSSB_OPT_PROFILE 4       /* 1: record a CUDA event between phases (hssb_phase_time)   */
#define HSSB_OPT_DEBUG 5         /* measurement only (library built with make DEBUG_MODES=1), WRONG RESULTS: bit 0 = leaf kernels compute on whatever is in
                                    shared memory without waiting for data, bit 1 = move data without computing */
#define HSSB_OPT_PIPELINE_COLS 6 /* host entry: right-hand sides per pipelined block (0 = automatic)  */
#define HSSB_OPT_ADJOINT_TWIN 7  /* 1 (default): hssb_matmul_t of a uniform tree keeps a transposed twin of the pool on the device;
                                    0: release it / never build it.  hssb_get_option returns 2 once the twin exists. */
#define HSSB_OPT_ULV_FAST 8      /* EXPERIMENTAL, default 0.  1: on uniform trees lay the ULV factors out in the shapes and padding of
                                    the product's blocks ("fast form": the leaf output becomes g b + ptb t like Y = D X + U F, zloc is
                                    not formed at the leaves, P' is split per child) so that the fixed-shape DMMA kernels run the solve's
                                    leaf phases, square merges and top-down steps.  Rebuilds the solve plan and drops existing factors.
                                    hssb_get_option returns 2 when the plan is in fast form.  Plan and factorisation are covered by the
                                    CPU tests; the kernels' use on this plan has not been run on a GPU yet.                            */
#define HSSB_OPT_TREE_KERNEL 9   /* 1 (default): on uniform trees every merge / translate level between the two leaf kernels (and, on
                                    sharded handles with the NVLink exchange, the exchange itself) runs inside ONE persistent cooperative
                                    kernel with grid barriers (csrc/hssb_tree.cuh) instead of one launch per level; 0: one launch per
                                    level (the round-1 schedule, kept as a cross-check); 2: as 1 without the cooperative-launch attribute */
#define HSSB_OPT_HOST_BOUNCE 10  /* host entry, PAGEABLE caller memory (an ordinary Julia Matrix, matmul.jl:13): 1 (default) = calls of
                                    16 MiB and more whose X / Y pointer is not pinned or registered are staged by the library through
                                    rings of pinned 2 MiB slots filled / drained by worker threads (HSSB_HOST_THREADS per direction,
                                    default min(8, cores / 2)), so that staging overlaps the DMA and the kernels; 0 = hand every
                                    pointer to cudaMemcpy2DAsync as it is; 2 = always stage (tests)                                 */
#define HSSB_OPT_LAST_BOUNCE 11  /* read-only: what the last host call staged through the rings (bit 0: X, bit 1: Y)                */
#define HSSB_OPT_HOST_THREADS 15 /* read-only: worker threads per direction of the pageable staging (HSSB_HOST_THREADS, else the cores the
                                    process may use / (2 x GPUs of the box), clamped to 2..8)                                        */
#define HSSB_OPT_LEAF_KERNEL 12  /* fixed-shape leaf kernels: 2 (default) = second generation, every ring stage holds a chunk of [D U] / V'
                                    together with the matching rows of X (csrc/hssb_leaf2.cuh); 1 = first generation (whole X block
                                    resident and double buffered, csrc/hssb_fast.cuh), kept as a bit-identical cross-check; 3 = second
                                    generation with longer chunks (measurement)                                                     */
#define HSSB_OPT_LEAF_FUSION 13  /* 0 (default): two passes over X (leaf-up V' X, leaf-down D X + U F).  1: the "X once" variant of
                                    csrc/hssb_leafx.cuh -- leaf-up reads each X block once for BOTH D X and V' X ([D ; V'] streamed as one
                                    stacked operand) and parks alpha D X + beta Y in Y, leaf-down adds alpha U F.  Uniform trees with
                                    (leaf, rank) in {(128,32), (128,64)}; other shapes keep the default.  Same results to
                                    rounding; measured slower (Y is written twice and read once more), kept as the measured alternative */
#define HSSB_OPT_FLOW_KERNEL 14  /* trees no fixed-shape kernel applies to (ragged leaves, variable ranks: every matrix that comes out of a
                                    compression) can run the WHOLE product as one persistent dataflow kernel: tasks are drawn from a
                                    queue in level order and wait on per-task counters for the producers of their operands, so the
                                    2*depth+2 dependent levels cost a flag round trip each instead of a launch (csrc/hssb_flow.cuh).
                                    Single-shard handles, product and transposed product.  2 (default): automatic -- the dataflow
                                    kernel for plain launches; when the schedule is replayed as a CUDA graph (the host entry, or
                                    HSSB_OPT_USE_GRAPH) one launch per level with programmatic dependent launch is faster (config-2
                                    shape 170 us against 194 us) and is used instead.  1: always.  0: never.  hssb_get_option returns
                                    2 once the product plan has been set up for it                                                  */
#define HSSB_OPT_BUSH_KERNEL 16  /* 0 (default).  1: SMALL any-shape trees (leaves of at most 64 rows / columns, ranks <= 64, at most 16384
                                    leaves: BASELINE configs 1-2) run every merge / translate level between the two leaf launches as ONE
                                    launch whose work items are BUSHES -- the tasks of a few consecutive levels below one node --
                                    executed by one CTA out of shared memory (bulk-copy staging, warp-sized tasks, flags between bushes
                                    only: 10 dependent steps for config 2 instead of 20 levels; csrc/hssb_bush.cuh).  2: any single-shard
                                    any-shape plan.  Takes precedence over HSSB_OPT_FLOW_KERNEL where it applies.  Parity-tested; measured
                                    SLOWER than the dataflow kernel (config-2 shape: 225 us against 196 us, profiles/bush_kernel_r02.txt),
                                    hence off by default.  hssb_get_option returns 3 once the product plan runs on it                   */
#define HSSB_OPT_BUSH_LEVELS 17  /* levels per bush * 16 + merge levels of the bush that holds the root (default 2 * 16 + 1); rebuilds the plan */
#define HSSB_OPT_PDL 18          /* bits; 1 (default; environment HSSB_PDL overrides): the node kernels of a uniform tree's level schedule are
                                    launched with programmatic dependent launch where that was measured to pay -- the one-shot kernels
                                    (rank <= 32, 32 < nrhs <= 64: the 25 merge / translate launches of config 3), the persistent
                                    node kernel at rank 64 and the any-shape tile kernel (one launch per level): a kernel's CTAs are placed while its predecessor still runs, fetch their
                                    generator blocks, and wait (griddepcontrol.wait) until the predecessor has completed -- no drain /
                                    launch gap between the levels; same order, bit-identical results (config 3 -1.8 %, config-5 shape
                                    -3 %, n = 2^16 -5 %, config-2 shape -10 %).  2: every persistent node kernel, 4: the leaf kernels (measured
                                    slower).  0: plain stream order                                                                */
#define HSSB_OPT_LAST_FACTOR_US 19 /* read-only: device time (microseconds, CUDA events) of the level launches of the last ULV factorisation;
                                    the wall time of hssb_ulv_factor also holds the allocation and clearing of the factor pool       */
int hssb_set_option(hssb_matrix* h, int opt, int64_t value);
int64_t hssb_get_option(const hssb_matrix* h, int opt);
/* Kernels launched by this handle since creation (for bench.py's gpu_launches). */
int64_t hssb_launch_count(const hssb_matrix* h);

/* Per-phase accounting of the level schedule and, after a call made with
 * HSSB_OPT_PROFILE = 1 (graph replay off), its device time measured with CUDA
 * events on the launching stream.  The phases are those of the plan the last
 * profiled call ran (product by default; hssb_matmul_t without the twin pool;
 * hssb_solve).  kind: 0 leaf-up, 1 merge, 2 exchange,
 * 3 translate, 4 leaf-down.  ms < 0 if no profiled call was made.             */
typedef struct hssb_phase_time_t {
  int64_t kind, level, top, fast, ntasks;
  int64_t flops_per_rhs; /* algorithmic flops per right-hand-side column        */
  int64_t gen_elems;     /* generator doubles the phase reads                   */
  int64_t x_rows, y_rows; /* rows of X read / rows of Y written (x nrhs x 8 B)  */
  double ms;
} hssb_phase_time_t;
int hssb_phase_count(const hssb_matrix* h);
int hssb_phase_time(hssb_matrix* h, int i, hssb_phase_time_t* out);

/* ---- multi-GPU exchange (one rank per GPU) ---------------------------- */
/* Rank 0 calls hssb_comm_unique_id and ships the 128 bytes to every rank by
 * any means (torch.distributed broadcast, MPI, a file); every rank then calls
 * hssb_comm_init.  The only collective on the path is one all-gather of the
 * subtree-root Z blocks per product.                                         */
int hssb_comm_unique_id(void* id128);
int hssb_comm_init(hssb_matrix* h, const void* id128, int rank, int n_ranks);

/* Peer-memory exchange (preferred on NVLink/NVSwitch boxes): instead of NCCL, every rank maps the
 * exchange buffers of all peers through CUDA IPC and the subtree-root Z blocks are pushed with
 * NVLink peer stores from inside the level schedule (one small kernel, graph-replayable).
 * Call hssb_reserve(max_nrhs) first, then hssb_xchg_export on every rank, all-gather the 128-byte
 * blobs by any means (rank order), then hssb_xchg_import on every rank.                          */
int hssb_xchg_export(hssb_matrix* h, void* handle128);
int hssb_xchg_import(hssb_matrix* h, const void* all_handles, int n_ranks);

/* ---- one process, one call, P devices ("a single ccall from one Julia thread drives all GPUs") ----
 * A group holds P sharded handles of one matrix, shard g (the g-th subtree at depth log2 P plus the
 * replicated top tree) on devices[g], wired to each other in process: the devices enable peer access
 * and the exchange kernels store into the peers' workspaces directly (no IPC, no NCCL, no MPI.jl).  This
 * is what the Julia drop-in `hssA * X` (matmul.jl:13) uses when more than one GPU is asked for.  P is a
 * power of two <= 16 and the tree must be at least log2 P deep.  Entries of `devices` may repeat, at most twice each (two
 * shards on one GPU: slower, but exercises the sharded path on a single-GPU box).
 *   hssb_group_finalize          from a builder that holds the WHOLE tree (every shard copies what it owns)
 *   hssb_group_create_synthetic  the benchmark matrices, generated per shard on its device
 *   hssb_group_matmul[_t]        X, Y = the whole host matrices (same contract as hssb_matmul[_t]); every shard
 *                                runs the pipelined host entry on its row block, all of them concurrently
 *   hssb_group_matmul_dev        dX[g] / dY[g] = device pointers ON devices[g] to shard g's row blocks (leading
 *                                dimensions ldx / ldy), asynchronous on streams[g] (NULL table: each shard's own
 *                                stream, wait with hssb_group_sync)
 *   hssb_group_shard             borrow shard g's handle (options, hssb_info, ...); owned by the group          */
typedef struct hssb_group hssb_group;
int hssb_group_finalize(hssb_builder* b, int64_t root, const int* devices, int n_devices, hssb_group** out);
int hssb_group_create_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, const int* devices, int n_devices,
                                hssb_group** out);
int hssb_group_destroy(hssb_group* g);
int hssb_group_size(const hssb_group* g);
hssb_matrix* hssb_group_shard(hssb_group* g, int i);
int hssb_group_reserve(hssb_group* g, int64_t max_nrhs);
int hssb_group_matmul(hssb_group* g, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y, int64_t ldy,
                      double alpha, double beta);
int hssb_group_matmul_t(hssb_group* g, int64_t rows_y, int64_t rows_x, int64_t nrhs, const double* X, int64_t ldx, double* Y,
                        int64_t ldy, double alpha, double beta);
int hssb_group_matmul_dev(hssb_group* g, int64_t nrhs, const double* const* dX, int64_t ldx, double* const* dY, int64_t ldy, double alpha,
                          double beta, void* const* streams);
int hssb_group_sync(hssb_group* g);

/* ---- measurement helpers (used by bench.py; not on the product path) --- */
/* kind 0: FP64 FMA (DFMA) register-resident peak, kind 1: FP64 tensor (DMMA
 * m8n8k4) peak, returns TFLOP/s.  kind 2: device copy bandwidth over `bytes`
 * bytes (read+write counted), returns GB/s.                                  */
int hssb_measure_peak(int device, int kind, int64_t bytes_or_iters, double* out);

/* ---- test hooks: host-only planning (no device needed) ----------------- */
/* Run the packer and the level scheduler without touching a GPU and expose the
 * resulting task table, so that CPU-only tests can check pool layout, task
 * wiring and shard plans with a numpy interpreter.  A plan-only handle cannot
 * multiply: hssb_matmul* fail with HSSB_ERR_CUDA (there is no CPU fallback). */
typedef struct hssb_task_t {
  int64_t a0, a1, b0, b1, c;
  int64_t lda0, lda1, ldb0, ldb1, ldc;
  int64_t M, K0, K1;
  int64_t ta0, ta1, sb0, sb1, sc, epilogue; /* sb/sc: 0 = X, 1 = Z, 2 = F, 3 = Y */
} hssb_task_t;
typedef struct hssb_phase_t {
  int64_t kind; /* 0 leaf-up, 1 merge, 2 exchange, 3 translate, 4 leaf-down */
  int64_t task0, ntasks, maxM, level, top, fast;
  int64_t xchg_zoff, xchg_slot_rows;
  int64_t transposed; /* 0: forward plan, 1: plan of hssb_matmul_t, 2: plan of hssb_solve over the ULV factor pool */
} hssb_phase_t;
int hssb_plan_only(hssb_builder* b, int64_t root, int shard_rank, int n_shards, hssb_matrix** out);
int hssb_plan_only_synthetic(int64_t n, int64_t leafsize, int64_t rank, uint64_t seed, int shard_rank,
                             int n_shards, hssb_matrix** out);
int hssb_debug_counts(const hssb_matrix* h, int64_t* n_tasks, int64_t* n_phases, int64_t* pool_len);
int hssb_debug_task(const hssb_matrix* h, int64_t i, hssb_task_t* out);
int hssb_debug_phase(const hssb_matrix* h, int64_t i, hssb_phase_t* out);
int hssb_debug_pool(const hssb_matrix* h, double* out, int64_t len);
/* Diagnostics: per-step microseconds of the persistent tree kernel (HSSB_OPT_TREE_KERNEL) launched alone on the
 * current workspace contents; returns the number of steps written (<= cap) or a negative status.               */
int hssb_debug_tree_trace(hssb_matrix* h, int64_t nrhs, double* us_out, int cap);
int hssb_debug_pool_t(const hssb_matrix* h, double* out, int64_t len); /* image of the adjoint twin pool */
/* The bush plan (csrc/hssb_bush.cuh) of plan-only or device handles, for the numpy interpreter: mode 0 = product,
 * 1 = transposed task table.  An op is one warp's work: rows [m0, m0 + mr) of task `task` (index into
 * hssb_debug_task), executed in level `level` of bush `bush`; s0 / s1 / sc >= 0: the operand / output block has a
 * shared-memory image at that offset (doubles) with leading dimension lds0 / lds1 / ldsc.                        */
typedef struct hssb_bush_op_t {
  int64_t task, bush, level, m0, mr;
  int64_t s0, s1, sc, lds0, lds1, ldsc, to_global;
  int64_t sa0, sa1; /* >= 0: the A block is read from a staged shared-memory copy of the generator block */
} hssb_bush_op_t;
/* A staged copy into shared memory at `dst` when the bush starts: kind 0 = `count` doubles of the pool from offset src;
 * 1 = Z, 2 = F: one 16-column tile of the workspace block at row src (count = ld * 16); 3 = rows [src, src + count) of
 * the column tile of X with leading dimension ld; 4 = the bush's ops.  Everything has landed before level 0.        */
typedef struct hssb_bush_stage_t {
  int64_t bush, kind, src, dst, count, ld, level;
} hssb_bush_stage_t;
int hssb_debug_bush_counts(hssb_matrix* h, int mode, int64_t* n_bush, int64_t* n_ops, int64_t* n_deps, int64_t* smem_doubles);
/* Diagnostics: timeline of the last bush-kernel launch.  hssb_debug_bush_trace(h, mode, NULL, 0) switches recording on (the
 * next call re-allocates the flags with a trace buffer); with `out` it copies 12 words per item (bush-major, then column
 * tile): SM, then globaltimer ns when the item was drawn, its dependencies were met, levels 0..7 ended, its flag was
 * published.  Returns the number of items written or a negative status.                                                */
int64_t hssb_debug_bush_trace(hssb_matrix* h, int mode, uint64_t* out, int64_t cap_items);
int64_t hssb_debug_bush_stage(hssb_matrix* h, int mode, int64_t i, hssb_bush_stage_t* out); /* returns the number of stages */
int hssb_debug_bush_op(hssb_matrix* h, int mode, int64_t i, hssb_bush_op_t* out);
/* dependencies of bush b: returns their count, writes at most cap bush indices */
int64_t hssb_debug_bush_deps(hssb_matrix* h, int mode, int64_t b, int64_t* out, int64_t cap);
/* ULV: factorise a plan-only handle on the host with the device's node routine (single-thread team),
 * and read the factor pool (either kind of handle) for the numpy plan interpreter.                   */
int hssb_debug_ulv_factor_host(hssb_matrix* h);
int hssb_debug_ulv_pool(const hssb_matrix* h, double* out, int64_t len);

#ifdef __cplusplus
}
#endif
#endif /* HSSB200_H */
