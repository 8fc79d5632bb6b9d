// hssb_plan.cuh — the packer's planning half: tree annotation (depth, offsets, shard cut), pool and workspace
// layout, the task tables of Y = A X and Y = A' X (one GTask per small GEMM of src/matmul.jl:32-62) and their phase
// lists, uniform-tree detection, upload.  Host code; included by hssb_api.cu.
#pragma once

namespace hssb {

// ------------------------------------------------------------ plan building ---
struct BlockSource {  // per node: either host copies or nothing (synthetic)
  const HostBlock* blk[BK_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// V and W are only ever applied transposed (matmul.jl:34, :39): the pool stores V' and W' so that
// every generator is a plain column-major "N" operand for the kernels.
static inline bool stored_transposed(int kind) { return kind == BK_V || kind == BK_W; }

static void block_shape(const std::vector<Node>& nodes, const Node& t, int kind, int64_t& rows, int64_t& cols) {
  rows = cols = 0;
  const Node* par = t.parent >= 0 ? &nodes[(size_t)t.parent] : nullptr;
  switch (kind) {
    case BK_D: if (t.leaf && !t.remote) { rows = t.m; cols = t.n; } break;
    case BK_U: if (t.leaf && !t.remote) { rows = t.m; cols = t.kr; } break;
    case BK_V: if (t.leaf && !t.remote) { rows = t.kw; cols = t.n; } break;  // stored TRANSPOSED (V' is kw x n): every A operand is then column-major "N"
    case BK_B12: if (!t.leaf && !t.remote) { rows = nodes[(size_t)t.left].kr; cols = nodes[(size_t)t.right].kw; } break;
    case BK_B21: if (!t.leaf && !t.remote) { rows = nodes[(size_t)t.right].kr; cols = nodes[(size_t)t.left].kw; } break;
    case BK_R: if (par) { rows = t.kr; cols = par->kr; } break;
    case BK_W: if (par) { rows = par->kw; cols = t.kw; } break;  // stored TRANSPOSED (W' is kw(parent) x kw)
    default: break;
  }
}

// Fills depth/height/row0/col0 (pre-order), validates the shard layout, marks
// top / local nodes.  `nodes` must be in BFS order with node 0 the root.
static int annotate_tree(hssb_matrix* H) {
  auto& nodes = H->nodes;
  const int P = H->n_shards;
  if (!is_pow2(P)) HSSB_FAIL(HSSB_ERR_ARG, "n_shards must be a power of two, got %d", P);
  int p = 0;
  while ((1 << p) < P) ++p;
  // the root acts as rooted(): no own translators (hssmatrix.jl:266)
  nodes[0].kr = nodes[0].kw = 0;
  nodes[0].parent = -1;
  nodes[0].depth = 0;
  nodes[0].row0 = nodes[0].col0 = 0;
  int64_t maxdepth = 0;
  for (size_t i = 0; i < nodes.size(); ++i) {  // BFS order: parents precede children
    Node& t = nodes[i];
    if (!t.leaf && !t.remote) {
      Node& l = nodes[(size_t)t.left];
      Node& r = nodes[(size_t)t.right];
      l.parent = r.parent = (int64_t)i;
      l.depth = r.depth = t.depth + 1;
      l.row0 = t.row0; l.col0 = t.col0;
      r.row0 = t.row0 + l.m; r.col0 = t.col0 + l.n;
      if (l.m + r.m != t.m || l.n + r.n != t.n)
        HSSB_FAIL(HSSB_ERR_DIM, "node %zu: children sizes do not add up", i);
    }
    maxdepth = std::max<int64_t>(maxdepth, t.depth);
  }
  H->depth = maxdepth;
  for (size_t i = nodes.size(); i-- > 0;) {
    Node& t = nodes[i];
    t.height = (t.leaf || t.remote) ? 0 : 1 + std::max(nodes[(size_t)t.left].height, nodes[(size_t)t.right].height);
  }
  // shard layout
  std::vector<int64_t> cut;  // nodes at depth p, left to right (BFS keeps that order)
  for (size_t i = 0; i < nodes.size(); ++i) {
    Node& t = nodes[i];
    t.top = t.depth < p;
    if (t.top && (t.leaf || t.remote))
      HSSB_FAIL(HSSB_ERR_ARG, "tree too shallow for %d shards: node %zu at depth %d is a leaf", P, i, t.depth);
    if (t.depth == p) cut.push_back((int64_t)i);
  }
  if ((int)cut.size() != P) HSSB_FAIL(HSSB_ERR_ARG, "expected %d subtrees at depth %d, found %zu", P, p, cut.size());
  if (H->shard_rank < 0 || H->shard_rank >= P) HSSB_FAIL(HSSB_ERR_ARG, "shard_rank %d out of range", H->shard_rank);
  for (int g = 0; g < P; ++g) {
    const Node& t = nodes[(size_t)cut[(size_t)g]];
    if (g == H->shard_rank ? t.remote : !t.remote)
      HSSB_FAIL(HSSB_ERR_ARG, "subtree %d at the shard cut must be %s on shard %d", g,
                g == H->shard_rank ? "local" : "a remote placeholder", H->shard_rank);
  }
  const Node& lr = nodes[(size_t)cut[(size_t)H->shard_rank]];
  for (size_t i = 0; i < nodes.size(); ++i) {
    Node& t = nodes[i];
    t.local = !t.top && !t.remote;
    if (t.remote && t.depth != p) HSSB_FAIL(HSSB_ERR_ARG, "remote placeholder %zu is not at the shard cut", i);
  }
  H->m = nodes[0].m; H->n = nodes[0].n;
  H->local_m = lr.m; H->local_n = lr.n;
  H->local_row0 = lr.row0; H->local_col0 = lr.col0;
  return HSSB_OK;
}

static bool on_root_path(const std::vector<Node>& nodes, int64_t node, int64_t local_root) {
  // true if `node` is an ancestor-or-self of local_root
  for (int64_t t = local_root; t >= 0; t = nodes[(size_t)t].parent)
    if (t == node) return true;
  return false;
}

// Assigns pool offsets in level order:
//   [leaf D][leaf U][leaf V] then per depth (deepest first) [B12][B21][R][W].
// Every block starts on a 128-byte boundary and has an even leading dimension
// so that any column is 16-byte aligned (vector loads / bulk copies).
static void layout_pool(hssb_matrix* H) {
  auto& nodes = H->nodes;
  int64_t off = 0, gen = 0;
  auto place = [&](Node& t, int kind) {
    int64_t rows, cols;
    block_shape(nodes, t, kind, rows, cols);
    t.rows[kind] = rows; t.cols[kind] = cols;
    // Uniform trees served by the fixed-shape kernels store every block with the +4 padded leading
    // dimension of its shared-memory image, so that a block (or a run of its columns) is ONE
    // contiguous TMA bulk copy that lands bank-conflict free.
    const int64_t ldp = H->padded ? rows + 4 : round_up(rows, 2);
    if (rows == 0 || cols == 0) { t.off[kind] = -1; t.ld[kind] = (int32_t)std::max<int64_t>(ldp, 2); return; }
    t.ld[kind] = (int32_t)ldp;
    t.off[kind] = off;
    off += round_up((int64_t)t.ld[kind] * cols, 16);
    gen += rows * cols;
  };
  const int leaf_kinds[3] = {BK_D, BK_U, BK_V};
  for (int kk = 0; kk < 3; ++kk)
    for (int64_t li : H->leaves) place(nodes[(size_t)li], leaf_kinds[kk]);
  const int lvl_kinds[4] = {BK_B12, BK_B21, BK_R, BK_W};
  for (int64_t d = H->depth; d >= 0; --d)
    for (int kk = 0; kk < 4; ++kk)
      for (auto& t : nodes)
        if (t.depth == d) place(t, lvl_kinds[kk]);
  H->pool_len = std::max<int64_t>(off, 16);
  H->gen_elems = gen;
}

static void layout_workspace(hssb_matrix* H) {
  auto& nodes = H->nodes;
  const int P = H->n_shards;
  int p = 0;
  while ((1 << p) < P) ++p;
  int64_t zo = 0, fo = 0;
  // exchange slots first (depth-p nodes in rank order, equal slot size)
  if (P > 1) {
    int64_t slot = 0;
    for (auto& t : nodes)
      if (t.depth == p) slot = std::max<int64_t>(slot, H->padded ? t.kw + 4 : round_up(t.kw, 2));
    H->xchg_zoff = 0;
    H->xchg_slot_rows = slot;
    for (auto& t : nodes)
      if (t.depth == p) { t.zoff = zo; t.ldz = (int32_t)std::max<int64_t>(H->padded ? t.kw + 4 : round_up(t.kw, 2), 2); zo += slot; }
  }
  for (size_t i = 1; i < nodes.size(); ++i) {  // BFS order keeps siblings adjacent
    Node& t = nodes[i];
    const int64_t lz = H->padded ? t.kw + 4 : round_up(t.kw, 2), lf = H->padded ? t.kr + 4 : round_up(t.kr, 2);
    if (t.zoff < 0) { t.zoff = zo; t.ldz = (int32_t)std::max<int64_t>(lz, 2); zo += lz; }
    t.foff = fo; t.ldf = (int32_t)std::max<int64_t>(lf, 2); fo += lf;
  }
  H->z_rows = std::max<int64_t>(zo, 2);
  H->f_rows = std::max<int64_t>(fo, 2);
}

static void add_phase(hssb_matrix* H, int kind, int level, bool top, std::vector<GTask>& batch,
                      std::vector<Phase>* dst = nullptr) {
  if (batch.empty()) return;
  Phase ph;
  ph.kind = kind; ph.level = level; ph.top = top;
  ph.task0 = (int64_t)H->tasks_host.size();
  ph.ntasks = (int64_t)batch.size();
  for (auto& t : batch) {
    ph.maxM = std::max(ph.maxM, t.M);
    if (!dst) H->flops_per_rhs += 2ll * t.M * ((int64_t)t.K0 + t.K1);
  }
  H->tasks_host.insert(H->tasks_host.end(), batch.begin(), batch.end());
  (dst ? *dst : H->phases).push_back(ph);
  batch.clear();
}

// Task table of Y = A' X on the SAME packed generators (`*(A, hssB)`,
// src/matmul.jl:14, which in the reference copies the whole adjoint tree, hssmatrix.jl:165-171,
// on every call).  The adjoint swaps roles: U <-> V, R <-> W, B12 <-> B21', D -> D'; every stored
// block is therefore applied transposed (ta = 1, the any-shape kernel), the "Z" blocks of the
// adjoint have kr rows and live in the F workspace, its "F" blocks have kw rows and live in Z.
// Single shard only.
static void build_plan_transposed(hssb_matrix* H) {
  auto& nodes = H->nodes;
  if (H->n_shards != 1) return;
  std::vector<GTask> batch;
  auto blank = []() { GTask t; memset(&t, 0, sizeof(t)); t.lda0 = t.lda1 = t.ldb0 = t.ldb1 = t.ldc = 2; return t; };
  auto& out = H->phases_t;
  // leaf up: Z' = U' X[rows]
  for (int64_t li : H->leaves) {
    const Node& t = nodes[(size_t)li];
    if (t.parent < 0 || t.kr == 0) continue;
    GTask g = blank();
    g.a0 = t.off[BK_U]; g.lda0 = t.ld[BK_U]; g.ta0 = 1; g.sb0 = SRC_X; g.b0 = t.row0; g.K0 = (int32_t)t.m;
    g.M = (int32_t)t.kr; g.sc = SRC_F; g.c = t.foff; g.ldc = t.ldf;
    batch.push_back(g);
  }
  add_phase(H, PH_LEAF_UP, 0, false, batch, &out);
  // merges: Z' = R1' Z1' + R2' Z2'
  for (int h = 1; h <= nodes[0].height; ++h) {
    for (auto& t : nodes) {
      if (t.leaf || t.height != h || t.parent < 0 || t.kr == 0) continue;
      const Node& l = nodes[(size_t)t.left];
      const Node& r = nodes[(size_t)t.right];
      GTask g = blank();
      g.a0 = l.off[BK_R]; g.lda0 = l.ld[BK_R]; g.ta0 = 1; g.sb0 = SRC_F; g.b0 = l.foff; g.ldb0 = l.ldf; g.K0 = (int32_t)l.kr;
      g.a1 = r.off[BK_R]; g.lda1 = r.ld[BK_R]; g.ta1 = 1; g.sb1 = SRC_F; g.b1 = r.foff; g.ldb1 = r.ldf; g.K1 = (int32_t)r.kr;
      if (g.a0 < 0) g.K0 = 0;
      if (g.a1 < 0) g.K1 = 0;
      g.M = (int32_t)t.kr; g.sc = SRC_F; g.c = t.foff; g.ldc = t.ldf;
      batch.push_back(g);
    }
    add_phase(H, PH_MERGE, h, false, batch, &out);
  }
  // translates: F1' = B21' Z2' (+ W1 F'), F2' = B12' Z1' (+ W2 F')   (the pool holds W', so W = (W')')
  for (int d = 0; d <= (int)H->depth; ++d) {
    for (auto& t : nodes) {
      if (t.leaf || t.depth != d) continue;
      const Node& l = nodes[(size_t)t.left];
      const Node& r = nodes[(size_t)t.right];
      const bool has_f = t.parent >= 0 && t.kw > 0;
      for (int side = 0; side < 2; ++side) {
        const Node& c = side ? r : l;
        const Node& sb = side ? l : r;
        if (c.kw == 0) continue;
        GTask g = blank();
        const int bk = side ? BK_B12 : BK_B21;  // B21 is kr(r) x kw(l): B21' maps Z'(r) to F'(l)
        g.a0 = t.off[bk]; g.lda0 = t.ld[bk]; g.ta0 = 1; g.sb0 = SRC_F; g.b0 = sb.foff; g.ldb0 = sb.ldf; g.K0 = (int32_t)sb.kr;
        if (has_f) { g.a1 = c.off[BK_W]; g.lda1 = c.ld[BK_W]; g.ta1 = 1; g.sb1 = SRC_Z; g.b1 = t.zoff; g.ldb1 = t.ldz; g.K1 = (int32_t)t.kw; }
        if (g.a0 < 0) g.K0 = 0;
        if (g.a1 < 0) g.K1 = 0;
        g.M = (int32_t)c.kw; g.sc = SRC_Z; g.c = c.zoff; g.ldc = c.ldz;
        batch.push_back(g);
      }
    }
    add_phase(H, PH_TRANSLATE, d, false, batch, &out);
  }
  // leaf down: Y[cols] = alpha (D' X[rows] + V F') + beta Y   (the pool holds V')
  for (int64_t li : H->leaves) {
    const Node& t = nodes[(size_t)li];
    GTask g = blank();
    g.a0 = t.off[BK_D]; g.lda0 = t.ld[BK_D]; g.ta0 = 1; g.sb0 = SRC_X; g.b0 = t.row0; g.K0 = (int32_t)t.m;
    if (t.parent >= 0 && t.kw > 0) { g.a1 = t.off[BK_V]; g.lda1 = t.ld[BK_V]; g.ta1 = 1; g.sb1 = SRC_Z; g.b1 = t.zoff; g.ldb1 = t.ldz; g.K1 = (int32_t)t.kw; }
    if (g.a0 < 0) g.K0 = 0;
    g.M = (int32_t)t.n; g.sc = SRC_Y; g.c = t.col0; g.epilogue = 1;
    if (g.M > 0) batch.push_back(g);
  }
  add_phase(H, PH_LEAF_DOWN, 0, false, batch, &out);
}

static void build_plan(hssb_matrix* H) {
  auto& nodes = H->nodes;
  const int P = H->n_shards;
  int p = 0;
  while ((1 << p) < P) ++p;
  int64_t local_root = 0;
  for (size_t i = 0; i < nodes.size(); ++i)
    if (nodes[i].depth == p && !nodes[i].remote) local_root = (int64_t)i;
  const int64_t r0 = H->local_row0, c0 = H->local_col0;
  std::vector<GTask> batch;
  auto blank = []() { GTask t; memset(&t, 0, sizeof(t)); t.lda0 = t.lda1 = t.ldb0 = t.ldb1 = t.ldc = 2; return t; };

  // ---- leaf up: Z = V' X (matmul.jl:34); skipped for a root leaf and for kw == 0
  for (int64_t li : H->leaves) {
    const Node& t = nodes[(size_t)li];
    if (t.parent < 0 || t.kw == 0) continue;
    GTask g = blank();
    g.a0 = t.off[BK_V]; g.lda0 = t.ld[BK_V]; g.ta0 = 0;  // the pool holds V' (kw x n)
    g.sb0 = SRC_X; g.b0 = t.col0 - c0;
    g.M = (int32_t)t.kw; g.K0 = (int32_t)t.n; g.K1 = 0;
    g.sc = SRC_Z; g.c = t.zoff; g.ldc = t.ldz;
    batch.push_back(g);
  }
  add_phase(H, PH_LEAF_UP, 0, false, batch);

  // ---- merges: Z = W1' Z1 + W2' Z2 (matmul.jl:39), never for the root (W is k x 0)
  auto merge_task = [&](const Node& t) {
    const Node& l = nodes[(size_t)t.left];
    const Node& r = nodes[(size_t)t.right];
    GTask g = blank();
    g.a0 = l.off[BK_W]; g.lda0 = l.ld[BK_W]; g.ta0 = 0; g.sb0 = SRC_Z; g.b0 = l.zoff; g.ldb0 = l.ldz; g.K0 = (int32_t)l.kw;
    g.a1 = r.off[BK_W]; g.lda1 = r.ld[BK_W]; g.ta1 = 0; g.sb1 = SRC_Z; g.b1 = r.zoff; g.ldb1 = r.ldz; g.K1 = (int32_t)r.kw;
    g.M = (int32_t)t.kw;
    g.sc = SRC_Z; g.c = t.zoff; g.ldc = t.ldz;
    return g;
  };
  const int max_h = nodes[0].height;
  for (int h = 1; h <= max_h; ++h) {
    for (auto& t : nodes)
      if (t.local && !t.leaf && t.height == h && t.parent >= 0 && t.kw > 0) batch.push_back(merge_task(t));
    add_phase(H, PH_MERGE, h, false, batch);
  }
  if (P > 1) {
    Phase ph; ph.kind = PH_EXCHANGE; H->phases.push_back(ph);
    for (int d = p - 1; d >= 1; --d) {  // top tree, bottom-up; only nodes OFF the root->local path feed a local F
      for (size_t i = 0; i < nodes.size(); ++i) {
        const Node& t = nodes[i];
        if (t.top && t.depth == d && t.kw > 0 && !on_root_path(nodes, (int64_t)i, local_root)) batch.push_back(merge_task(t));
      }
      add_phase(H, PH_MERGE, d, true, batch);
    }
  }

  // ---- translates: F1 = B12 Z2 (+ R1 F), F2 = B21 Z1 (+ R2 F) (matmul.jl:51-57)
  auto translate_tasks = [&](const Node& t, bool only_path) {
    const Node& l = nodes[(size_t)t.left];
    const Node& r = nodes[(size_t)t.right];
    const bool has_f = t.parent >= 0 && t.kr > 0;
    for (int side = 0; side < 2; ++side) {
      const Node& c = side ? r : l;   // child receiving F
      const Node& s = side ? l : r;   // sibling providing Z
      const int64_t ci = side ? t.right : t.left;
      if (c.kr == 0) continue;
      if (only_path && !on_root_path(nodes, ci, local_root)) continue;
      if (c.remote) continue;
      GTask g = blank();
      const int bk = side ? BK_B21 : BK_B12;
      g.a0 = t.off[bk]; g.lda0 = t.ld[bk]; g.ta0 = 0; g.sb0 = SRC_Z; g.b0 = s.zoff; g.ldb0 = s.ldz; g.K0 = (int32_t)s.kw;
      if (has_f) { g.a1 = c.off[BK_R]; g.lda1 = c.ld[BK_R]; g.ta1 = 0; g.sb1 = SRC_F; g.b1 = t.foff; g.ldb1 = t.ldf; g.K1 = (int32_t)t.kr; }
      g.M = (int32_t)c.kr;
      g.sc = SRC_F; g.c = c.foff; g.ldc = c.ldf;
      if (g.a0 < 0) g.K0 = 0;
      if (g.a1 < 0) g.K1 = 0;
      batch.push_back(g);
    }
  };
  for (int d = 0; d < p; ++d) {
    for (auto& t : nodes)
      if (t.top && t.depth == d) translate_tasks(t, true);
    add_phase(H, PH_TRANSLATE, d, true, batch);
  }
  if (P > 1) { Phase ph; ph.kind = PH_XCHG_ACK; H->phases.push_back(ph); }  // gathered Z blocks are consumed from here on
  for (int d = p; d <= (int)H->depth; ++d) {
    for (auto& t : nodes)
      if (t.local && !t.leaf && t.depth == d) translate_tasks(t, false);
    add_phase(H, PH_TRANSLATE, d, false, batch);
  }

  // ---- leaf down: Y = alpha (D X + U F) + beta Y (matmul.jl:46-47; :21-22 for a root leaf)
  for (int64_t li : H->leaves) {
    const Node& t = nodes[(size_t)li];
    GTask g = blank();
    g.a0 = t.off[BK_D]; g.lda0 = t.ld[BK_D]; g.sb0 = SRC_X; g.b0 = t.col0 - c0; g.K0 = (int32_t)t.n;
    if (t.parent >= 0 && t.kr > 0) { g.a1 = t.off[BK_U]; g.lda1 = t.ld[BK_U]; g.sb1 = SRC_F; g.b1 = t.foff; g.ldb1 = t.ldf; g.K1 = (int32_t)t.kr; }
    g.M = (int32_t)t.m;
    g.sc = SRC_Y; g.c = t.row0 - r0; g.epilogue = 1;
    if (g.a0 < 0) g.K0 = 0;
    if (g.M > 0) batch.push_back(g);
  }
  add_phase(H, PH_LEAF_DOWN, 0, false, batch);
}

}  // namespace hssb
#include "hssb_ulv_plan.cuh"
namespace hssb {

static void detect_uniform(hssb_matrix* H) {
  // Fast fixed-shape kernels need: a perfect local tree, square leaves of one
  // size, one rank everywhere (rows and columns).
  H->uniform = false;
  if (H->leaves.empty()) return;
  const Node& l0 = H->nodes[(size_t)H->leaves[0]];
  if (l0.parent < 0) return;
  const int64_t m = l0.m, r = l0.kr;
  if (m != l0.n || r != l0.kw || r <= 0) return;
  for (auto& t : H->nodes) {
    if (t.parent < 0) continue;
    if (t.kr != r || t.kw != r) return;
    if (t.leaf && !t.remote && (t.m != m || t.n != m || t.depth != H->depth)) return;
  }
  H->uniform = true;
  H->uni_m = m;
  H->uni_r = r;
}

// Common tail of finalize / create_synthetic once H->nodes is filled (BFS order).
// Host-only part: tree annotation, pool / workspace layout, task table, phases.
static int plan_matrix(hssb_matrix* H) {
  int rc = annotate_tree(H);
  if (rc) return rc;
  auto& nodes = H->nodes;
  // local leaves left to right = pre-order walk
  {
    std::vector<int64_t> stack{0};
    while (!stack.empty()) {
      const int64_t i = stack.back();
      stack.pop_back();
      const Node& t = nodes[(size_t)i];
      if (t.remote) continue;
      if (t.leaf) { H->leaves.push_back(i); continue; }
      stack.push_back(t.right);
      stack.push_back(t.left);
    }
  }
  for (auto& t : nodes) {
    if (t.leaf && !t.remote) {
      H->max_leaf_m = std::max(H->max_leaf_m, t.m);
      H->max_leaf_n = std::max(H->max_leaf_n, t.n);
    }
    H->max_rank = std::max(H->max_rank, std::max(t.kr, t.kw));
  }
  detect_uniform(H);
  H->padded = H->uniform && fast_shape_supported(H->uni_m, H->uni_r);
  layout_pool(H);
  layout_workspace(H);
  build_plan(H);
  plan_fast_phases(H);
  build_plan_transposed(H);
  H->ulv_task0 = (int64_t)H->tasks_host.size();
  build_plan_ulv(H);
  return HSSB_OK;
}

// Device part: allocate the pool, upload (or generate) the generators and the task table.
static int finish_matrix(hssb_matrix* H, const std::vector<BlockSource>* src, FILE* pool_file = nullptr) {
  int rc = plan_matrix(H);
  if (rc) return rc;
  auto& nodes = H->nodes;
  DeviceGuard dg(H->device);
  if (!dg.ok) HSSB_FAIL(HSSB_ERR_CUDA, "cudaSetDevice(%d) failed", H->device);
  HSSB_CUDA(cudaStreamCreateWithFlags(&H->stream, cudaStreamNonBlocking));
  if (cudaMalloc(&H->pool_dev, (size_t)H->pool_len * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    HSSB_FAIL(HSSB_ERR_ALLOC, "device allocation of the %.3f GB generator pool failed", H->pool_len * 8e-9);
  }
  HSSB_CUDA(cudaMemsetAsync(H->pool_dev, 0, (size_t)H->pool_len * sizeof(double), H->stream));
  if (!H->tasks_host.empty()) {
    HSSB_CUDA(cudaMalloc(&H->tasks_dev, H->tasks_host.size() * sizeof(GTask)));
    HSSB_CUDA(cudaMemcpyAsync(H->tasks_dev, H->tasks_host.data(), H->tasks_host.size() * sizeof(GTask),
                              cudaMemcpyHostToDevice, H->stream));
  }
  if (pool_file) {
    // packed pool image from a file written by hssb_save: stream it through a pinned buffer
    const size_t CH = (size_t)1 << 22;
    double* stage = nullptr;
    HSSB_CUDA(cudaMallocHost(&stage, CH * sizeof(double)));
    for (int64_t off = 0; off < H->pool_len; off += (int64_t)CH) {
      const size_t cnt = (size_t)std::min<int64_t>((int64_t)CH, H->pool_len - off);
      if (fread(stage, sizeof(double), cnt, pool_file) != cnt) {
        cudaFreeHost(stage);
        HSSB_FAIL(HSSB_ERR_ARG, "hssb_load: file is truncated");
      }
      HSSB_CUDA(cudaMemcpyAsync(H->pool_dev + off, stage, cnt * sizeof(double), cudaMemcpyHostToDevice, H->stream));
      HSSB_CUDA(cudaStreamSynchronize(H->stream));
    }
    cudaFreeHost(stage);
  } else if (src) {
    // host generators: assemble in pinned chunks and upload
    const size_t CH = (size_t)1 << 22;  // 32 MiB of doubles per chunk
    double* stage[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
      HSSB_CUDA(cudaMallocHost(&stage[i], CH * sizeof(double)));
      HSSB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    struct Piece { int64_t off; const HostBlock* hb; int32_t ld; bool tr; int64_t rows, cols; };  // rows/cols as stored
    std::vector<Piece> pieces;
    for (size_t i = 0; i < nodes.size(); ++i)
      for (int k = 0; k < BK_COUNT; ++k)
        if (nodes[i].off[k] >= 0)
          pieces.push_back({nodes[i].off[k], (*src)[i].blk[k], nodes[i].ld[k], stored_transposed(k), nodes[i].rows[k], nodes[i].cols[k]});
    auto put = [](double* dst, const Piece& pc) {  // dst has leading dimension pc.ld
      if (!pc.tr) {
        for (int64_t j = 0; j < pc.cols; ++j) memcpy(dst + j * pc.ld, pc.hb->data.data() + j * pc.rows, (size_t)pc.rows * sizeof(double));
      } else {  // stored(i, j) = host(j, i), host is cols x rows with leading dimension cols
        for (int64_t j = 0; j < pc.cols; ++j)
          for (int64_t i = 0; i < pc.rows; ++i) dst[j * pc.ld + i] = pc.hb->data[(size_t)(i * pc.cols + j)];
      }
    };
    std::sort(pieces.begin(), pieces.end(), [](const Piece& a, const Piece& b) { return a.off < b.off; });
    size_t pi = 0;
    int cur = 0;
    while (pi < pieces.size()) {
      const int64_t base = pieces[pi].off;
      HSSB_CUDA(cudaEventSynchronize(ev[cur]));
      size_t pj = pi;
      int64_t end = base;
      memset(stage[cur], 0, CH * sizeof(double));
      while (pj < pieces.size()) {
        const Piece& pc = pieces[pj];
        const int64_t pend = pc.off + (int64_t)pc.ld * pc.cols;
        if (pend - base > (int64_t)CH) break;
        put(stage[cur] + (pc.off - base), pc);
        end = pend;
        ++pj;
      }
      if (pj == pi) {  // single block larger than a chunk: upload it on its own
        const Piece& pc = pieces[pi];
        std::vector<double> tmp((size_t)pc.ld * (size_t)pc.cols, 0.0);
        put(tmp.data(), pc);
        HSSB_CUDA(cudaMemcpy(H->pool_dev + pc.off, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
        ++pi;
        continue;
      }
      HSSB_CUDA(cudaMemcpyAsync(H->pool_dev + base, stage[cur], (size_t)(end - base) * sizeof(double),
                                cudaMemcpyHostToDevice, H->stream));
      HSSB_CUDA(cudaEventRecord(ev[cur], H->stream));
      cur ^= 1;
      pi = pj;
    }
    HSSB_CUDA(cudaStreamSynchronize(H->stream));
    for (int i = 0; i < 2; ++i) { cudaFreeHost(stage[i]); cudaEventDestroy(ev[i]); }
  } else {
    // synthetic generators, produced on the device
    std::vector<SynthBlock> sb;
    const double tscale = H->synth_rank > 0 ? 1.0 / sqrt(2.0 * (double)H->synth_rank) : 1.0;
    for (auto& t : nodes)
      for (int k = 0; k < BK_COUNT; ++k)
        if (t.off[k] >= 0) {
          SynthBlock b;
          b.off = t.off[k];
          b.key = synth_key(H->seed, t.heap_id, k);
          b.rows = (int32_t)t.rows[k]; b.cols = (int32_t)t.cols[k]; b.ld = t.ld[k];
          b.transposed = stored_transposed(k);
          b.c = IH4_SCALE * ((k == BK_R || k == BK_W) ? tscale : 1.0);
          sb.push_back(b);
        }
    if (!sb.empty()) {
      SynthBlock* dsb = nullptr;
      HSSB_CUDA(cudaMalloc(&dsb, sb.size() * sizeof(SynthBlock)));
      HSSB_CUDA(cudaMemcpyAsync(dsb, sb.data(), sb.size() * sizeof(SynthBlock), cudaMemcpyHostToDevice, H->stream));
      const int grid = (int)std::min<size_t>(sb.size(), 148 * 16);
      synth_fill_kernel<<<grid, 256, 0, H->stream>>>(dsb, (int64_t)sb.size(), H->pool_dev);
      HSSB_CUDA(cudaGetLastError());
      HSSB_CUDA(cudaStreamSynchronize(H->stream));
      cudaFree(dsb);
    }
  }
  HSSB_CUDA(cudaStreamSynchronize(H->stream));
  return HSSB_OK;
}

// The ULV solve needs more workspace rows per node than the product (zloc + [b; u] against Z); the
// larger layout is only allocated once a solve asks for it.
static int ensure_workspace(hssb_matrix* H, int64_t nrhs, bool ulv = false) {
  if (nrhs <= H->ws_nrhs && (!ulv || H->ws_ulv)) return HSSB_OK;
  if (H->xchg_exported)
    HSSB_FAIL(HSSB_ERR_STATE, "the Z workspace is mapped by peer ranks: hssb_reserve(max_nrhs) before hssb_xchg_export (have %lld, need %lld)",
              (long long)H->ws_nrhs, (long long)nrhs);
  ulv = ulv || H->ws_ulv;
  nrhs = std::max(nrhs, H->ws_nrhs);
  if (H->z_dev) cudaFree(H->z_dev);
  if (H->f_dev) cudaFree(H->f_dev);
  H->z_dev = H->f_dev = nullptr;
  H->ws_nrhs = 0;
  H->ws_ulv = false;
  const size_t zb = (size_t)std::max(H->z_rows, ulv ? H->ulv_z_rows : 0) * (size_t)nrhs * sizeof(double);
  const size_t fb = (size_t)std::max(H->f_rows, ulv ? H->ulv_f_rows : 0) * (size_t)nrhs * sizeof(double);
  if (cudaMalloc(&H->z_dev, zb) != cudaSuccess || cudaMalloc(&H->f_dev, fb) != cudaSuccess) {
    cudaGetLastError();
    HSSB_FAIL(HSSB_ERR_ALLOC, "device allocation of the Z/F workspaces (%.3f GB) failed", (zb + fb) * 1e-9);
  }
  H->ws_nrhs = nrhs;
  H->ws_ulv = ulv;
  invalidate_graphs(H);
  return HSSB_OK;
}

}  // namespace hssb
