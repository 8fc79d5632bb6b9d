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
