// hssb_ulv_plan.cuh — host side of the ULV solver: shapes, factor-pool layout, the solve's task
// table and the driver of the factorisation kernel.  Included by hssb_api.cu only (uses its
// static helpers round_up / add_phase); the device code is hssb_ulv.cuh.
#pragma once

#include "hssb_ulv.cuh"

namespace hssb {

// Shapes of the implicit ULV factorisation depend on sizes and ranks only, so the whole solve plan is
// built with the product plan; hssb_ulv_factor later fills the factor pool.
static void build_plan_ulv(hssb_matrix* H) {
  auto& nodes = H->nodes;
  H->ulv.clear();
  H->phases_u.clear();
  auto unsupported = [&](const char* why) { H->ulv.clear(); H->ulv_why = why; };
  if (H->n_shards != 1) return unsupported("the ULV solver runs on a single shard");
  if (H->m != H->n || H->m == 0) return unsupported("the ULV solver needs a square matrix");
  std::vector<UlvNode> uv(nodes.size());

  // ---- shapes, children before parents (BFS order reversed)
  for (size_t i = nodes.size(); i-- > 0;) {
    const Node& t = nodes[i];
    UlvNode& u = uv[i];
    u.is_root = t.parent < 0;
    u.is_leaf = t.leaf;
    u.kr = (int32_t)t.kr; u.kw = (int32_t)t.kw;
    if (t.leaf) {
      u.m_in = (int32_t)t.m; u.n_in = (int32_t)t.n;
      u.D = t.off[BK_D]; u.ldD = t.ld[BK_D];
      u.U = t.off[BK_U]; u.ldU = t.ld[BK_U];
      u.V = t.off[BK_V]; u.ldV = t.ld[BK_V];
    } else {
      const Node& l = nodes[(size_t)t.left];
      const Node& r = nodes[(size_t)t.right];
      const UlvNode& c1 = uv[(size_t)t.left];
      const UlvNode& c2 = uv[(size_t)t.right];
      u.left = (int32_t)t.left; u.right = (int32_t)t.right;
      u.k1 = c1.k; u.kr1 = c1.kr; u.kw1 = c1.kw; u.no1 = c1.n_out;
      u.k2 = c2.k; u.kr2 = c2.kr; u.kw2 = c2.kw; u.no2 = c2.n_out;
      u.m_in = u.k1 + u.k2; u.n_in = u.no1 + u.no2;
      u.B12 = t.off[BK_B12]; u.ldB12 = t.ld[BK_B12];
      u.B21 = t.off[BK_B21]; u.ldB21 = t.ld[BK_B21];
      u.R1 = l.off[BK_R]; u.ldR1 = l.ld[BK_R]; u.R2 = r.off[BK_R]; u.ldR2 = r.ld[BK_R];
      u.W1 = l.off[BK_W]; u.ldW1 = l.ld[BK_W]; u.W2 = r.off[BK_W]; u.ldW2 = r.ld[BK_W];
    }
    if (u.is_root) {
      if (u.m_in != u.n_in) return unsupported("the reduced root block of the ULV factorisation is not square");
      u.k = 0; u.mk = 0; u.n_out = 0;
    } else if (u.kr < u.m_in) {  // compressible (ulvfactor.jl:28-31)
      u.k = u.kr; u.mk = u.m_in - u.kr;
      if (u.mk > u.n_in) return unsupported("a node eliminates more rows than it has columns (ulvfactor.jl:29 nk < m-k)");
      u.n_out = u.n_in - u.mk;
    } else {                     // full-rank block, nothing is eliminated here (ulvfactor.jl:31-37)
      u.k = u.m_in; u.mk = 0; u.n_out = u.n_in;
    }
  }

  // ---- fast form (HSSB_OPT_ULV_FAST): on a uniform tree every block of the solve has the shape of a
  // block of the product (leaf size m, rank r: [T2; T3] is 2r x m like V', g is m x m like D, ptb m x r
  // like U, the c merges are 2r x 2r, the top-down blocks r x r), so with the padded (+4) layout the
  // fixed-shape kernels of hssb_fast.cuh can run the solve's big phases.
  bool ff = H->ulv_fast_form && H->padded && !uv[0].is_leaf;
  if (ff) {
    const int32_t m = (int32_t)H->uni_m, r = (int32_t)H->uni_r;
    for (const UlvNode& u : uv) {
      if (u.is_root) ff = ff && u.m_in == 2 * r && u.n_in == 2 * r;
      else if (u.is_leaf) ff = ff && u.m_in == m && u.n_in == m && u.kr == r && u.kw == r && u.mk == m - r && u.n_out == r;
      else ff = ff && u.m_in == 2 * r && u.n_in == 2 * r && u.kr == r && u.kw == r && u.mk == r && u.n_out == r;
    }
  }
  H->ulv_ff = ff;
  auto ld_of = [&](int64_t rows) { return (int32_t)std::max<int64_t>(ff ? rows + 4 : round_up(rows, 2), 2); };

  // ---- factor pool / reduced-generator scratch / workspaces
  int64_t off = 0, red = 0, zo = 0, fo = 0;
  auto place = [&](int64_t rows, int64_t cols, int32_t& ld) {
    ld = ld_of(rows);
    if (rows == 0 || cols == 0) return (int64_t)-1;
    const int64_t at = off;
    off += round_up((int64_t)ld * cols, 16);
    return at;
  };
  H->ulv_MI = H->ulv_NI = H->ulv_KR = H->ulv_KW = 1;
  for (size_t i = 0; i < nodes.size(); ++i) {
    UlvNode& u = uv[i];
    H->ulv_MI = std::max(H->ulv_MI, u.m_in); H->ulv_NI = std::max(H->ulv_NI, u.n_in);
    H->ulv_KR = std::max(H->ulv_KR, u.kr); H->ulv_KW = std::max(H->ulv_KW, u.kw);
    const int64_t rows_c = u.is_root ? u.n_in : u.k + u.kw;
    int32_t ld_tmp;
    if (u.is_leaf) {
      if (!ff) u.az[0] = place(u.mk, u.m_in, u.ld_az);
      u.ac[0] = place(rows_c, u.m_in, u.ld_ac);
      if (ff) u.g = place(u.n_in, u.m_in, u.ld_g);
    } else {
      u.az[0] = place(u.mk, u.k1 + u.kw1, u.ld_az);
      u.az[1] = place(u.mk, u.k2 + u.kw2, ld_tmp);
      u.ac[0] = place(rows_c, u.k1 + u.kw1, u.ld_ac);
      u.ac[1] = place(rows_c, u.k2 + u.kw2, ld_tmp);
    }
    if (!u.is_root) {
      if (ff && !u.is_leaf) {  // P' split by rows: one pair of r x r blocks per child
        u.pta_c[0] = place(u.no1, u.mk, u.ld_ptc[0]);
        u.ptb_c[0] = place(u.no1, u.n_out, ld_tmp);
        u.pta_c[1] = place(u.no2, u.mk, u.ld_ptc[1]);
        u.ptb_c[1] = place(u.no2, u.n_out, ld_tmp);
      } else {
        if (!ff) u.pta = place(u.n_in, u.mk, u.ld_pt);
        u.ptb = place(u.n_in, u.n_out, ff ? u.ld_pt : ld_tmp);
      }
      u.rD = red; red += (int64_t)u.k * u.n_out;
      u.rU = red; red += (int64_t)u.k * u.kr;
      u.rV = red; red += (int64_t)u.n_out * u.kw;
      u.ld_zloc = ld_of(u.mk);
      u.ld_c = ld_of(u.k + u.kw);
      u.ld_t = ld_of(u.n_out);
      if (!(ff && u.is_leaf)) { u.zloc = zo; zo += u.ld_zloc; }  // fast-form leaves never form zloc
      u.c = zo; zo += u.ld_c;
      u.t = fo; fo += u.ld_t;
    }
  }
  H->ulv_pool_len = std::max<int64_t>(off, 16);
  H->ulv_red_len = std::max<int64_t>(red, 16);
  H->ulv_z_rows = std::max<int64_t>(zo, 2);
  H->ulv_f_rows = std::max<int64_t>(fo, 2);

  // ---- the solve's task table: C = A0*B0 + A1*B1 over the factor pool
  std::vector<GTask> batch;
  auto& out = H->phases_u;
  int64_t flops = 0;
  auto blank = []() { GTask g; memset(&g, 0, sizeof(g)); g.lda0 = g.lda1 = g.ldb0 = g.ldb1 = g.ldc = 2; g.a0 = g.a1 = -1; return g; };
  auto push = [&](GTask g) {
    if (g.a0 < 0) g.K0 = 0;
    if (g.a1 < 0) g.K1 = 0;
    if (g.M <= 0) return;
    flops += 2ll * g.M * ((int64_t)g.K0 + g.K1);
    batch.push_back(g);
  };
  // rows [r0, r0 + M) of the two-block operator [A(:, child 1 columns) | A(:, child 2 columns)] applied to (c1, c2)
  auto merge_rows = [&](const UlvNode& u, const int64_t a[2], int32_t lda, int64_t r0, int32_t M) {
    const UlvNode& c1 = uv[(size_t)u.left];
    const UlvNode& c2 = uv[(size_t)u.right];
    GTask g = blank();
    g.M = M;
    if (a[0] >= 0) { g.a0 = a[0] + r0; g.lda0 = lda; g.sb0 = SRC_Z; g.b0 = c1.c; g.ldb0 = c1.ld_c; g.K0 = c1.k + c1.kw; }
    if (a[1] >= 0) { g.a1 = a[1] + r0; g.lda1 = lda; g.sb1 = SRC_Z; g.b1 = c2.c; g.ldb1 = c2.ld_c; g.K1 = c2.k + c2.kw; }
    return g;
  };

  if (uv[0].is_leaf) {  // hssA.D \ b (ulvfactor.jl:11-12)
    const UlvNode& u = uv[0];
    GTask g = blank();
    g.a0 = u.ac[0]; g.lda0 = u.ld_ac; g.sb0 = SRC_X; g.b0 = 0; g.K0 = u.m_in; g.M = u.n_in;
    g.sc = SRC_Y; g.c = 0;
    push(g);
    add_phase(H, PH_LEAF_DOWN, 0, false, batch, &out);
  } else {
    // upsweep, leaves: zloc = T1 b, c = [T2; T3] b
    for (int64_t li : H->leaves) {
      const UlvNode& u = uv[(size_t)li];
      const Node& t = nodes[(size_t)li];
      for (int part = ff ? 1 : 0; part < 2; ++part) {  // fast form: c only, zloc is folded into the leaf output
        GTask g = blank();
        g.a0 = part ? u.ac[0] : u.az[0]; g.lda0 = part ? u.ld_ac : u.ld_az;
        g.sb0 = SRC_X; g.b0 = t.row0; g.K0 = u.m_in;
        g.M = part ? u.k + u.kw : u.mk;
        g.sc = SRC_Z; g.c = part ? u.c : u.zloc; g.ldc = part ? u.ld_c : u.ld_zloc;
        push(g);
      }
    }
    add_phase(H, PH_LEAF_UP, 0, false, batch, &out);
    // upsweep, branches by height
    for (int h = 1; h < nodes[0].height; ++h) {
      for (size_t i = 0; i < nodes.size(); ++i) {
        const Node& t = nodes[i];
        if (t.leaf || t.height != h || t.parent < 0) continue;
        const UlvNode& u = uv[i];
        GTask g = merge_rows(u, u.az, u.ld_az, 0, u.mk);
        g.sc = SRC_Z; g.c = u.zloc; g.ldc = u.ld_zloc;
        if (!ff) push(g);
        g = merge_rows(u, u.ac, u.ld_ac, 0, u.k + u.kw);
        g.sc = SRC_Z; g.c = u.c; g.ldc = u.ld_c;
        push(g);
      }
      add_phase(H, PH_MERGE, h, false, batch, &out);
      if (ff) {  // the square 2r x 2r merges above form a fixed-shape phase of their own; zloc (r x 2r blocks) follows
        for (size_t i = 0; i < nodes.size(); ++i) {
          const Node& t = nodes[i];
          if (t.leaf || t.height != h || t.parent < 0) continue;
          const UlvNode& u = uv[i];
          GTask g = merge_rows(u, u.az, u.ld_az, 0, u.mk);
          g.sc = SRC_Z; g.c = u.zloc; g.ldc = u.ld_zloc;
          push(g);
        }
        add_phase(H, PH_MERGE, h, false, batch, &out);
      }
    }
    // root: [t1; t2] = D^-1 b (ulvfactor.jl:83), rows split between the children
    {
      const UlvNode& u = uv[0];
      const UlvNode& c1 = uv[(size_t)u.left];
      const UlvNode& c2 = uv[(size_t)u.right];
      GTask g = merge_rows(u, u.ac, u.ld_ac, 0, u.no1);
      g.sc = SRC_F; g.c = c1.t; g.ldc = c1.ld_t;
      push(g);
      g = merge_rows(u, u.ac, u.ld_ac, u.no1, u.no2);
      g.sc = SRC_F; g.c = c2.t; g.ldc = c2.ld_t;
      push(g);
      add_phase(H, PH_TRANSLATE, 0, false, batch, &out);
    }
    // top-down: z[cols] = P' [zloc; t] (ulvfactor.jl:98-107); a branch hands the result to its children
    auto down = [&](const UlvNode& u, int64_t r0, int32_t M) {
      GTask g = blank();
      g.M = M;
      if (u.pta >= 0) { g.a0 = u.pta + r0; g.lda0 = u.ld_pt; g.sb0 = SRC_Z; g.b0 = u.zloc; g.ldb0 = u.ld_zloc; g.K0 = u.mk; }
      if (u.ptb >= 0) { g.a1 = u.ptb + r0; g.lda1 = u.ld_pt; g.sb1 = SRC_F; g.b1 = u.t; g.ldb1 = u.ld_t; g.K1 = u.n_out; }
      return g;
    };
    auto down_child = [&](const UlvNode& u, int s) {  // fast form: child s's rows of P' are blocks of their own
      GTask g = blank();
      g.M = s ? u.no2 : u.no1;
      if (u.pta_c[s] >= 0) { g.a0 = u.pta_c[s]; g.lda0 = u.ld_ptc[s]; g.sb0 = SRC_Z; g.b0 = u.zloc; g.ldb0 = u.ld_zloc; g.K0 = u.mk; }
      if (u.ptb_c[s] >= 0) { g.a1 = u.ptb_c[s]; g.lda1 = u.ld_ptc[s]; g.sb1 = SRC_F; g.b1 = u.t; g.ldb1 = u.ld_t; g.K1 = u.n_out; }
      return g;
    };
    for (int d = 1; d <= (int)H->depth; ++d) {
      for (size_t i = 0; i < nodes.size(); ++i) {
        const Node& t = nodes[i];
        if (t.leaf || t.depth != d) continue;
        const UlvNode& u = uv[i];
        const UlvNode& c1 = uv[(size_t)u.left];
        const UlvNode& c2 = uv[(size_t)u.right];
        GTask g = ff ? down_child(u, 0) : down(u, 0, u.no1);
        g.sc = SRC_F; g.c = c1.t; g.ldc = c1.ld_t;
        push(g);
        g = ff ? down_child(u, 1) : down(u, u.no1, u.no2);
        g.sc = SRC_F; g.c = c2.t; g.ldc = c2.ld_t;
        push(g);
      }
      add_phase(H, PH_TRANSLATE, d, false, batch, &out);
    }
    for (int64_t li : H->leaves) {
      const UlvNode& u = uv[(size_t)li];
      GTask g = down(u, 0, u.n_in);
      if (ff) {  // Z[cols] = g b + ptb t: the shape of the product's leaf-down step (Y = D X + U F)
        g = blank();
        g.M = u.n_in;
        g.a0 = u.g; g.lda0 = u.ld_g; g.sb0 = SRC_X; g.b0 = nodes[(size_t)li].row0; g.K0 = u.m_in;
        g.a1 = u.ptb; g.lda1 = u.ld_pt; g.sb1 = SRC_F; g.b1 = u.t; g.ldb1 = u.ld_t; g.K1 = u.n_out;
        g.epilogue = 1;  // alpha = 1, beta = 0 in hssb_solve
      }
      g.sc = SRC_Y; g.c = nodes[(size_t)li].col0;
      push(g);
    }
    add_phase(H, PH_LEAF_DOWN, 0, false, batch, &out);
  }
  if (ff) {  // tag the phases a fixed-shape kernel of the product can run (same checks as plan_fast_phases)
    const int32_t m = (int32_t)H->uni_m, r = (int32_t)H->uni_r;
    auto node_r = [](int32_t R) { return R == 16 || R == 32 || R == 64; };
    for (Phase& ph : out) {
      const GTask* tk = H->tasks_host.data() + ph.task0;
      bool up = ph.kind == PH_LEAF_UP && fast_shape_supported(m, 2 * r), mc = ph.kind == PH_MERGE && node_r(2 * r);
      bool tr = ph.kind == PH_TRANSLATE && node_r(r), dn = ph.kind == PH_LEAF_DOWN && fast_shape_supported(m, r);
      for (int64_t i = 0; i < ph.ntasks; ++i) {
        const GTask& g = tk[i];
        const bool plain = !g.ta0 && !g.ta1 && g.a0 >= 0;
        up = up && plain && g.M == 2 * r && g.K0 == m && g.K1 == 0 && g.lda0 == 2 * r + 4 && g.ldc == 2 * r + 4 && g.sb0 == SRC_X && g.sc == SRC_Z;
        mc = mc && plain && g.a1 >= 0 && g.M == 2 * r && g.K0 == 2 * r && g.K1 == 2 * r && g.lda0 == 2 * r + 4 && g.lda1 == 2 * r + 4 &&
             g.ldb0 == 2 * r + 4 && g.ldb1 == 2 * r + 4 && g.ldc == 2 * r + 4 && g.sb0 == SRC_Z && g.sb1 == SRC_Z && g.sc == SRC_Z;
        tr = tr && plain && g.a1 >= 0 && g.M == r && g.K0 == r && g.K1 == r && g.lda0 == r + 4 && g.lda1 == r + 4 && g.ldb0 == r + 4 &&
             g.ldb1 == r + 4 && g.ldc == r + 4 && g.sb0 == SRC_Z && g.sb1 == SRC_F && g.sc == SRC_F;
        dn = dn && plain && g.a1 >= 0 && g.M == m && g.K0 == m && g.K1 == r && g.lda0 == m + 4 && g.lda1 == m + 4 && g.ldb1 == r + 4 &&
             g.sb0 == SRC_X && g.sb1 == SRC_F && g.sc == SRC_Y;
      }
      if (up) { ph.fast = FAST_LEAF_UP; ph.fast_m = m; ph.fast_r = 2 * r; }
      else if (mc) { ph.fast = FAST_MERGE; ph.fast_r = 2 * r; }
      else if (tr) { ph.fast = FAST_TRANSLATE; ph.fast_r = r; }
      else if (dn) { ph.fast = FAST_LEAF_DOWN; ph.fast_m = m; ph.fast_r = r; }
    }
  }
  H->ulv_flops_per_rhs = flops;
  H->ulv = std::move(uv);
  H->ulv_why.clear();
}

// Nodes grouped by height (children strictly below their parent).
static std::vector<std::vector<int32_t>> ulv_levels(const hssb_matrix* H) {
  std::vector<std::vector<int32_t>> lv((size_t)H->nodes[0].height + 1);
  for (size_t i = 0; i < H->nodes.size(); ++i) lv[(size_t)H->nodes[i].height].push_back((int32_t)i);
  return lv;
}

// Scratch maxima over the nodes of one level (ulv_scratch_len is sized per launch, not for the whole tree: in a
// tree whose upper nodes are much larger than its leaves -- ranks close to the block size -- a leaf level
// thousands of nodes wide must not be charged hundreds of copies of the root-sized scratch).
struct UlvLevelDims { int32_t MI = 1, NI = 1, KR = 1, KW = 1; };
static UlvLevelDims ulv_level_dims(const hssb_matrix* H, const std::vector<int32_t>& level) {
  UlvLevelDims d;
  for (int32_t i : level) {
    const UlvNode& u = H->ulv[(size_t)i];
    d.MI = std::max(d.MI, u.m_in); d.NI = std::max(d.NI, u.n_in);
    d.KR = std::max(d.KR, std::max(u.kr, std::max(u.kr1, u.kr2)));
    d.KW = std::max(d.KW, std::max(u.kw, std::max(u.kw1, u.kw2)));
  }
  return d;
}

// First node whose factorisation divided by a zero or non-finite pivot, -1 if none.
static int64_t ulv_first_breakdown(const std::vector<double>& pivmin) {
  for (size_t i = 0; i < pivmin.size(); ++i)
    if (!(pivmin[i] > 0.0)) return (int64_t)i;
  return -1;
}

// Host instantiation of the factorisation (single-thread team) for plan-only handles: CPU tests only.
static int ulv_factor_host(hssb_matrix* H) {
  H->ulv_pool_host.assign((size_t)H->ulv_pool_len, 0.0);
  H->ulv_factored = false;
  std::vector<double> red((size_t)H->ulv_red_len, 0.0);
  std::vector<double> pivmin(H->nodes.size(), INFINITY);
  const Team tm{0, 1};
  for (auto& level : ulv_levels(H)) {
    const UlvLevelDims d = ulv_level_dims(H, level);
    std::vector<double> scratch((size_t)ulv_scratch_len(d.MI, d.NI, d.KR, d.KW), 0.0);
    UlvCtx cx{H->ulv.data(), H->pool_host.data(), H->ulv_pool_host.data(), red.data(), d.MI, d.NI, d.KR, d.KW, pivmin.data()};
    for (int32_t node : level) {
      if (H->ulv_ff) ulv_factor_node<true>(tm, cx, node, scratch.data());
      else ulv_factor_node<false>(tm, cx, node, scratch.data());
    }
  }
  const int64_t bad = ulv_first_breakdown(pivmin);
  if (bad >= 0)
    HSSB_FAIL(HSSB_ERR_SINGULAR, "SingularException: the ULV factorisation met a zero pivot at node %lld (ulvfactor.jl:48 / :83)", (long long)bad);
  H->ulv_factored = true;
  return HSSB_OK;
}

// Device factorisation: one launch per tree level, one CTA per node.
// adjoint = true factorises A' from the adjoint twin pool (uniform trees: same shapes, hence the same solve plan)
// into a second factor pool: hssb_solve_t, i.e. `/(A, hssB)` (hssmatrix.jl:236).
static int ulv_factor_device(hssb_matrix* H, bool adjoint = false) {
  double*& fpool_dev = adjoint ? H->ulv_pool_t_dev : H->ulv_pool_dev;
  bool& factored = adjoint ? H->ulv_t_factored : H->ulv_factored;
  const double* gen_pool = adjoint ? H->pool_t_dev : H->pool_dev;
  if (!gen_pool) HSSB_FAIL(HSSB_ERR_STATE, "ULV factorisation: the generator pool is missing");
  if (fpool_dev) { cudaFree(fpool_dev); fpool_dev = nullptr; }
  factored = false;
  const size_t pool_b = (size_t)H->ulv_pool_len * sizeof(double);
  if (cudaMalloc(&fpool_dev, pool_b) != cudaSuccess) {
    cudaGetLastError();
    fpool_dev = nullptr;
    HSSB_FAIL(HSSB_ERR_ALLOC, "device allocation of the %.3f GB ULV factor pool failed", pool_b * 1e-9);
  }
  const auto levels = ulv_levels(H);
  // per level: scratch per CTA and CTAs in flight (three per SM: 80 registers x 256 threads; four for the fast form's
  // instantiation, 64 registers -- the kernel is latency-bound, ncu: 0.57 IPC per SM with 37 % of the warp slots in use;
  // fewer when the scratch of large nodes would not fit in 8 GiB)
  const size_t ctas_per_sm = H->ulv_ff ? 4 : 3;
  struct LevelRun { UlvLevelDims d; int64_t stride; int ctas; };
  std::vector<LevelRun> runs;
  size_t scratch_doubles = 1;
  for (auto& l : levels) {
    LevelRun r;
    r.d = ulv_level_dims(H, l);
    r.stride = round_up(ulv_scratch_len(r.d.MI, r.d.NI, r.d.KR, r.d.KW), 16);
    r.ctas = (int)std::min<size_t>(std::max<size_t>(l.size(), 1), 148 * ctas_per_sm);
    while (r.ctas > 1 && (size_t)r.ctas * (size_t)r.stride * 8 > ((size_t)8 << 30)) r.ctas /= 2;
    scratch_doubles = std::max(scratch_doubles, (size_t)r.ctas * (size_t)r.stride);
    runs.push_back(r);
  }
  UlvNode* d_nodes = nullptr;
  int32_t* d_list = nullptr;
  double *d_red = nullptr, *d_scratch = nullptr, *d_piv = nullptr;
  auto cleanup = [&]() { cudaFree(d_nodes); cudaFree(d_list); cudaFree(d_red); cudaFree(d_scratch); cudaFree(d_piv); };
  std::vector<double> pivmin(H->nodes.size(), INFINITY);
  cudaError_t e = cudaMalloc(&d_nodes, H->ulv.size() * sizeof(UlvNode));
  if (e == cudaSuccess) e = cudaMalloc(&d_list, H->nodes.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_red, (size_t)H->ulv_red_len * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_scratch, scratch_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_piv, pivmin.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(fpool_dev, 0, pool_b, H->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_nodes, H->ulv.data(), H->ulv.size() * sizeof(UlvNode), cudaMemcpyHostToDevice, H->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_piv, pivmin.data(), pivmin.size() * sizeof(double), cudaMemcpyHostToDevice, H->stream);
  std::vector<int32_t> flat;
  for (auto& l : levels) flat.insert(flat.end(), l.begin(), l.end());
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_list, flat.data(), flat.size() * sizeof(int32_t), cudaMemcpyHostToDevice, H->stream);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device time of the level launches alone (HSSB_OPT_LAST_FACTOR_US): the wall time of
  if (e == cudaSuccess) e = cudaEventCreate(&ev0);   // this call also holds the allocation and clearing of a multi-GB pool
  if (e == cudaSuccess) e = cudaEventCreate(&ev1);
  if (e == cudaSuccess) e = cudaEventRecord(ev0, H->stream);
  if (e == cudaSuccess) {
    size_t at = 0;
    for (size_t li = 0; li < levels.size(); ++li) {
      const auto& l = levels[li];
      const LevelRun& r = runs[li];
      if (!l.empty()) {
        UlvCtx cx{d_nodes, gen_pool, fpool_dev, d_red, r.d.MI, r.d.NI, r.d.KR, r.d.KW, d_piv};
        const int grid = (int)std::min<size_t>(l.size(), (size_t)r.ctas);
        if (H->ulv_ff) ulv_factor_kernel_ff<<<grid, 256, 0, H->stream>>>(cx, d_list + at, (int)l.size(), d_scratch, r.stride);
        else ulv_factor_kernel<<<grid, 256, 0, H->stream>>>(cx, d_list + at, (int)l.size(), d_scratch, r.stride);
        H->launches++;
      }
      at += l.size();
    }
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaEventRecord(ev1, H->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pivmin.data(), d_piv, pivmin.size() * sizeof(double), cudaMemcpyDeviceToHost, H->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(H->stream);
  if (e == cudaSuccess) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) H->ulv_last_factor_us = (int64_t)(ms * 1e3f);
  }
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  cleanup();
  if (e != cudaSuccess) {
    cudaFree(fpool_dev);
    fpool_dev = nullptr;
    HSSB_FAIL(HSSB_ERR_CUDA, "ULV factorisation failed: %s", cudaGetErrorString(e));
  }
  const int64_t bad = ulv_first_breakdown(pivmin);
  if (bad >= 0) {  // do not keep (or cache) factors full of Inf / NaN
    cudaFree(fpool_dev);
    fpool_dev = nullptr;
    HSSB_FAIL(HSSB_ERR_SINGULAR, "SingularException: the ULV factorisation met a zero pivot at node %lld (ulvfactor.jl:48 / :83)", (long long)bad);
  }
  factored = true;
  return HSSB_OK;
}

}  // namespace hssb
