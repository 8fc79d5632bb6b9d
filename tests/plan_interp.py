"""numpy interpreter of the task table the C++ packer/scheduler emits
(hssb_debug_*).  CPU tests use it to check pool layout, task wiring, level
ordering and the multi-shard plan WITHOUT a GPU.  It is test infrastructure: it
executes the library's plan, so a wrong plan gives a wrong answer here exactly
as it would on the device."""
import numpy as np

SRC_X, SRC_Z, SRC_F, SRC_Y = 0, 1, 2, 3
PH_LEAF_UP, PH_MERGE, PH_EXCHANGE, PH_TRANSLATE, PH_LEAF_DOWN = range(5)


class ShardState:
    def __init__(self, packed, X, Y, nrhs, pool=None):
        self.packed = packed
        self.tasks, self.phases, self.pool = packed.debug_plan()
        if pool is not None:  # the adjoint twin pool (same layout, same plan) or the ULV factor pool
            self.pool = pool
        self.nrhs = nrhs
        self.X = np.asfortranarray(X)
        self.Y = Y
        # poison the workspaces: reading an unwritten block must show up as NaN
        ui = packed.ulv_info   # the ULV solve uses more workspace rows per node than the product
        self.Z = np.full(max(packed.info.z_rows, ui.z_rows) * nrhs, np.nan)
        self.F = np.full(max(packed.info.f_rows, ui.f_rows) * nrhs, np.nan)

    def _a(self, off, ld, rows, cols, trans):
        if rows == 0 or cols == 0:
            return np.zeros((rows, cols))
        r, c = (cols, rows) if trans else (rows, cols)  # stored shape
        blk = np.lib.stride_tricks.as_strided(self.pool[off:], shape=(r, c), strides=(8, 8 * ld))
        return blk.T if trans else blk

    def _b(self, src, row, ld, K):
        N = self.nrhs
        if K == 0:
            return np.zeros((0, N))
        if src == SRC_X:
            return self.X[row:row + K, :]
        ws = self.Z if src == SRC_Z else self.F
        return np.lib.stride_tricks.as_strided(ws[row * N:], shape=(K, N), strides=(8, 8 * ld))

    def run_task(self, t, alpha, beta):
        N = self.nrhs
        acc = np.zeros((t.M, N))
        if t.K0 > 0:
            acc += self._a(t.a0, t.lda0, t.M, t.K0, t.ta0) @ self._b(t.sb0, t.b0, t.ldb0, t.K0)
        if t.K1 > 0:
            acc += self._a(t.a1, t.lda1, t.M, t.K1, t.ta1) @ self._b(t.sb1, t.b1, t.ldb1, t.K1)
        if t.sc == SRC_Y:
            dst = self.Y[t.c:t.c + t.M, :]
            if t.epilogue:
                dst[...] = alpha * acc + (beta * dst if beta != 0.0 else 0.0)
            else:
                dst[...] = acc
        else:
            ws = self.Z if t.sc == SRC_Z else self.F
            dst = np.lib.stride_tricks.as_strided(ws[t.c * N:], shape=(t.M, N), strides=(8, 8 * t.ldc))
            dst[...] = acc

    def run_phase(self, ph, alpha, beta):
        for i in range(ph.task0, ph.task0 + ph.ntasks):
            self.run_task(self.tasks[i], alpha, beta)


def run_plan(packed, X, Y, alpha=1.0, beta=0.0, trans=False, pool=None):
    """Single-shard plan: Y (in place) = alpha*op(A)*X + beta*Y, op(A) = A' if trans."""
    st = ShardState(packed, X, Y, X.shape[1], pool)
    mode = int(trans)   # 0 forward, 1 transposed task table, 2 ULV solve (pool = the factor pool)
    for ph in st.phases:
        if ph.transposed != mode:
            continue
        assert ph.kind != PH_EXCHANGE
        st.run_phase(ph, alpha, beta)
    return Y


def run_sharded(packs, Xs, Ys, alpha=1.0, beta=0.0, pools=None):
    """P plans in lockstep; the exchange phase is an all-gather of each shard's
    own slot of the exchange buffer."""
    nrhs = Xs[0].shape[1]
    pools = pools or [None] * len(packs)
    sts = [ShardState(p, x, y, nrhs, pl) for p, x, y, pl in zip(packs, Xs, Ys, pools)]
    for s in sts:
        s.phases = [ph for ph in s.phases if not ph.transposed]
    nph = len(sts[0].phases)
    assert all(len(s.phases) == nph for s in sts)
    for i in range(nph):
        kinds = {s.phases[i].kind for s in sts}
        assert len(kinds) == 1
        if PH_EXCHANGE in kinds:
            ph = sts[0].phases[i]
            cnt = ph.xchg_slot_rows * nrhs
            base = ph.xchg_zoff * nrhs
            slots = [s.Z[base + g * cnt: base + (g + 1) * cnt].copy() for g, s in enumerate(sts)]
            for s in sts:
                for g, sl in enumerate(slots):
                    s.Z[base + g * cnt: base + (g + 1) * cnt] = sl
        else:
            for s in sts:
                s.run_phase(s.phases[i], alpha, beta)
    return Ys


def run_bush_plan(packed, X, Y, alpha=1.0, beta=0.0, trans=False):
    """Execute the bush plan (csrc/hssb_bush.cuh: the tree cut into bushes of a few levels, one CTA per bush and
    16-column tile, intermediate blocks in shared memory) the way the kernel does and check its invariants:
    bushes are in topological order, a workspace block read from global memory was written by a declared
    dependency (or by an earlier level of the same bush), shared-memory images of one bush do not overlap and fit,
    the ops of one level never read each other's output, every row of every task is computed exactly once."""
    st = ShardState(packed, X, Y, X.shape[1])
    mode = int(trans)
    ops, deps, smem, stages = packed.debug_bush_plan(mode)
    N = st.nrhs
    TN = 16
    nb = len(deps)
    by_bush = [dict() for _ in range(nb)]
    for o in ops:
        by_bush[o.bush].setdefault(o.level, []).append(o)
    st_by_bush = [[] for _ in range(nb)]
    ST_OPS = 4
    for e in stages:
        assert e.dst % 2 == 0 and e.dst + e.count <= smem
        assert e.src % 2 == 0 and e.count % 2 == 0, "bulk copies move 16-byte pieces"
        st_by_bush[e.bush].append(e)
    writer = {}     # (src, workspace row) -> bush that wrote the block to global memory (-1: the leaf-up launch)
    rows_done = {}  # task -> rows computed
    done = set()
    # leaf-up launch first (the tile kernel), the bush kernel, then the leaf-down launch
    for ph in st.phases:
        if ph.transposed == mode and ph.kind == PH_LEAF_UP:
            for i in range(ph.task0, ph.task0 + ph.ntasks):
                st.run_task(st.tasks[i], alpha, beta)
                writer[(st.tasks[i].sc, st.tasks[i].c)] = -1
                rows_done[i] = st.tasks[i].M
    for b in range(nb):
        assert all(0 <= d < b for d in deps[b]), "bushes are not in topological order"
        assert all(d in done for d in deps[b])
        images = {}   # shared-memory offset -> (task, ld, array M x N)
        staged = {e.dst: e for e in st_by_bush[b]}
        regions = sorted((e.dst, e.dst + e.count) for e in st_by_bush[b])
        ops_here = sum(len(v) for v in by_bush[b].values())
        assert all(e.count == ops_here * 16 for e in st_by_bush[b] if e.kind == ST_OPS)
        assert all(regions[i][1] <= regions[i + 1][0] for i in range(len(regions) - 1)), "staged copies overlap"
        levels = by_bush[b]
        assert sorted(levels) == list(range(len(levels))), "a bush has an empty level"
        for l in range(len(levels)):
            pending = []
            for o in levels[l]:
                t = st.tasks[o.task]
                assert 0 <= o.m0 and 0 < o.mr <= 8 and o.m0 + o.mr <= t.M
                acc = np.zeros((o.mr, N))
                for s in range(2):
                    K = t.K1 if s else t.K0
                    if K <= 0:
                        continue
                    A = st._a(t.a1 if s else t.a0, t.lda1 if s else t.lda0, t.M, K, t.ta1 if s else t.ta0)[o.m0:o.m0 + o.mr]
                    soff, src, row = (o.s1, t.sb1, t.b1) if s else (o.s0, t.sb0, t.b0)
                    sa = o.sa1 if s else o.sa0
                    if sa >= 0:   # A from a staged copy of the generator block: same leading dimension, whole block present
                        e = staged[sa]
                        ta, lda, aoff = (t.ta1, t.lda1, t.a1) if s else (t.ta0, t.lda0, t.a0)
                        rows, cols = (K, t.M) if ta else (t.M, K)
                        assert e.kind == 0 and e.src == aoff and e.ld == lda and e.count >= lda * (cols - 1) + rows
                        assert e.src + e.count <= st.pool.size
                    if soff >= 0 and soff in staged:   # this column tile of a block another bush wrote, copied after the wait
                        e = staged[soff]
                        assert e.kind == src and e.src == row and e.ld == (t.ldb1 if s else t.ldb0) == (o.lds1 if s else o.lds0)
                        assert e.count == e.ld * TN and K <= e.ld
                        w = writer.get((src, row))
                        assert w is not None and w != b and (w == -1 or w in deps[b]), "staged operand without a dependency on its writer"
                        B = st._b(src, row, e.ld, K)
                    elif soff >= 0:
                        assert src in (SRC_Z, SRC_F) and soff in images, "shared-memory operand that nobody produced"
                        _, ld, img = images[soff]
                        assert ld == (o.lds1 if s else o.lds0) and img.shape[0] == K
                        B = img
                    elif src == SRC_X:
                        B = st.X[row:row + K, :]
                    else:
                        w = writer.get((src, row))
                        assert w is not None and (w == b or w == -1 or w in deps[b]), "global operand without a dependency on its writer"
                        B = st._b(src, row, t.ldb1 if s else t.ldb0, K)
                    acc += A @ B
                pending.append((o, t, acc))
            for o, t, acc in pending:   # the level's barrier
                if o.sc >= 0:
                    assert o.ldsc >= t.M and o.sc + o.ldsc * TN <= smem
                    if o.sc not in images:
                        for off, (tk, ld, _) in images.items():
                            assert o.sc >= off + ld * TN or off >= o.sc + o.ldsc * TN, "shared-memory images overlap"
                        for lo, hi in regions:
                            assert o.sc >= hi or lo >= o.sc + o.ldsc * TN, "an image overlaps a staged copy"
                        images[o.sc] = (o.task, o.ldsc, np.full((t.M, N), np.nan))
                    assert images[o.sc][0] == o.task
                    images[o.sc][2][o.m0:o.m0 + o.mr] = acc
                else:
                    assert o.to_global
                if o.to_global:
                    assert t.sc != SRC_Y
                    if True:
                        ws = st.Z if t.sc == SRC_Z else st.F
                        dst = np.lib.stride_tricks.as_strided(ws[t.c * N:], shape=(t.M, N), strides=(8, 8 * t.ldc))
                        dst[o.m0:o.m0 + o.mr] = acc
                        writer[(t.sc, t.c)] = b
                rows_done[o.task] = rows_done.get(o.task, 0) + o.mr
        done.add(b)
    for ph in st.phases:
        if ph.transposed == mode and ph.kind == PH_LEAF_DOWN:
            for i in range(ph.task0, ph.task0 + ph.ntasks):
                st.run_task(st.tasks[i], alpha, beta)
                rows_done[i] = st.tasks[i].M
    for ph in st.phases:
        if ph.transposed != mode or ph.kind == PH_EXCHANGE:
            continue
        for i in range(ph.task0, ph.task0 + ph.ntasks):
            assert rows_done.get(i, 0) == st.tasks[i].M, "a task is not covered exactly once"
    return Y, dict(bushes=nb, ops=len(ops), smem_doubles=smem, stages=len(stages), max_levels=max(len(x) for x in by_bush),
                   staged_a=sum(1 for o in ops if (o.sa0 >= 0 or st.tasks[o.task].K0 <= 0) and (o.sa1 >= 0 or st.tasks[o.task].K1 <= 0)),
                   staged_b=sum(1 for o in ops if (o.s0 >= 0 or st.tasks[o.task].K0 <= 0) and (o.s1 >= 0 or st.tasks[o.task].K1 <= 0)),
                   chain=_bush_chain(deps))


def _bush_chain(deps):
    depth = []
    for b, ds in enumerate(deps):
        depth.append(1 + max((depth[d] for d in ds), default=0))
    return max(depth, default=0)
