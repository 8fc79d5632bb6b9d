"""CPU oracle for the HSS x dense product of bonevbs/HssMatrices.jl (v0.1.6).

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package (hssmatrices.jl_b200/)
may import this file; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and only as the checker.

PARITY UNPINNED: the reference holds no golden vectors, known-answer tests or
seeded fixtures for this path (its one assertion, test/runtests.jl:63-64, is a
5e-5 tolerance check against a dense product with an unseeded randn), and Julia
is not installed in this image, so the reference itself cannot be run here.
The restatement is instead anchored on (1) the mathematical identity
hssA*X == full(hssA)*X with full() restated from src/hssmatrix.jl:270-305,
(2) an independent plain-C twin (oracle/hss_oracle.c, no BLAS) and (3) the
reference's own assertion restated in tests/test_oracle.py.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  All matrices are numpy float64; generators are kept as
separate arrays per node exactly like the reference's pointer tree.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only needed by the compress restatement (fixture builder)
    import scipy.linalg as _sla
except Exception:  # pragma: no cover
    _sla = None


# --------------------------------------------------------------------------
# Trees: src/binarytree.jl:7-24, src/clustertree.jl:10-35
# --------------------------------------------------------------------------
class ClusterTree:
    """BinaryNode{UnitRange{Int}} (src/binarytree.jl:7-20).  `data` is a
    half-open Python range (lo, hi) with 0-based lo; length = hi - lo."""

    __slots__ = ("data", "left", "right")

    def __init__(self, data, left=None, right=None):
        self.data = data
        self.left = left
        self.right = right

    def isleaf(self):  # src/binarytree.jl:23
        return self.left is None and self.right is None

    def isbranch(self):  # src/binarytree.jl:24
        return self.left is not None and self.right is not None

    def __len__(self):
        return self.data[1] - self.data[0]


def bisection_cluster(n_or_range, leafsize=64):
    """src/clustertree.jl:14-35.  Split while length > leafsize; the left child
    gets the first ceil(len/2) indices (`_bisection_cluster`, :27-35)."""
    if isinstance(n_or_range, (int, np.integer)):
        lo, hi = 0, int(n_or_range)
    else:
        lo, hi = int(n_or_range[0]), int(n_or_range[1])
    if hi - lo <= 0:  # :18
        raise ValueError("Index range must be larger or equal to 0")
    if leafsize < 1:
        raise ValueError("leafsize must be >= 1")
    return _bisection_cluster(lo, hi, int(leafsize))


def _bisection_cluster(lo, hi, leafsize):
    node = ClusterTree((lo, hi))
    length = hi - lo
    if length > leafsize:
        n = -(-length // 2)  # ceil(len/2), :30
        node.left = _bisection_cluster(lo, lo + n, leafsize)
        node.right = _bisection_cluster(lo + n, hi, leafsize)
    return node


def nleaves_cl(cl):
    return 1 if cl.isleaf() else nleaves_cl(cl.left) + nleaves_cl(cl.right)


def leaves(cl):
    if cl.isleaf():
        return [cl.data]
    return leaves(cl.left) + leaves(cl.right)


# --------------------------------------------------------------------------
# HssMatrix: src/hssmatrix.jl:11-68
# --------------------------------------------------------------------------
class HssMatrix:
    """Mutable recursive struct of src/hssmatrix.jl:11-34 with the same field
    names.  Fields that the reference leaves #undef are None here."""

    def __init__(self):
        self.leafnode = False
        self.rootnode = False
        self.D = self.U = self.V = None
        self.A11 = self.A22 = None
        self.B12 = self.B21 = None
        self.sz1 = self.sz2 = None
        self.R1 = self.W1 = self.R2 = self.W2 = None


def _f64(a):
    return np.asarray(a, dtype=np.float64)


def hss_leaf(D, U=None, V=None, rootnode=None):
    """Leaf constructors src/hssmatrix.jl:36-44."""
    D = _f64(D)
    m, n = D.shape
    h = HssMatrix()
    h.leafnode = True
    if U is None and V is None:  # :36-39 (defaults rootnode=true)
        h.rootnode = True if rootnode is None else rootnode
        h.D, h.U, h.V = D, np.zeros((m, 0)), np.zeros((n, 0))
        return h
    U, V = _f64(U), _f64(V)
    if D.shape[0] != U.shape[0]:  # :41
        raise ValueError("D and U must have same number of rows")
    if D.shape[1] != V.shape[0]:  # :42
        raise ValueError("D and V must have same number of columns")
    h.rootnode = False if rootnode is None else rootnode
    h.D, h.U, h.V = D, U, V
    return h


def hss_branch(A11, A22, B12, B21, R1=None, W1=None, R2=None, W2=None, rootnode=None):
    """Branch constructors src/hssmatrix.jl:46-67."""
    h = HssMatrix()
    h.leafnode = False
    h.A11, h.A22 = A11, A22
    h.B12, h.B21 = _f64(B12), _f64(B21)
    h.sz1, h.sz2 = size(A11), size(A22)
    if R1 is None:  # :46-55 — translators become k x 0
        kr1, kw1 = gensize(A11)
        kr2, kw2 = gensize(A22)
        h.rootnode = True if rootnode is None else rootnode
        h.R1, h.W1 = np.zeros((kr1, 0)), np.zeros((kw1, 0))
        h.R2, h.W2 = np.zeros((kr2, 0)), np.zeros((kw2, 0))
        return h
    R1, W1, R2, W2 = _f64(R1), _f64(W1), _f64(R2), _f64(W2)
    if R1.shape[1] != R2.shape[1]:  # :58
        raise ValueError("R1 and R2 must have same number of columns")
    if W1.shape[1] != W2.shape[1]:  # :59
        raise ValueError("W1 and W2 must have same number of rows")
    h.rootnode = False if rootnode is None else rootnode
    h.R1, h.W1, h.R2, h.W2 = R1, W1, R2, W2
    return h


def isleaf(h):  # src/hssmatrix.jl:88
    return h.leafnode


def size(h):  # src/hssmatrix.jl:94
    if h.leafnode:
        return tuple(h.D.shape)
    return (h.sz1[0] + h.sz2[0], h.sz1[1] + h.sz2[1])


def gensize(h):  # src/hssmatrix.jl:254-262
    if h.leafnode:
        return h.U.shape[1], h.V.shape[1]
    kr = h.R1.shape[1]
    if kr != h.R2.shape[1]:
        raise ValueError("dimensions of column-translators do not match")
    kw = h.W1.shape[1]
    if kw != h.W2.shape[1]:
        raise ValueError("dimensions of row-translators do not match")
    return kr, kw


def rooted(h):  # src/hssmatrix.jl:266
    if h.leafnode:
        return hss_leaf(h.D, rootnode=True)
    return hss_branch(h.A11, h.A22, h.B12, h.B21, rootnode=True)


def hssrank(h):  # src/hssmatrix.jl:251
    if h.leafnode:
        return 0
    return max(hssrank(h.A11), hssrank(h.A22), *h.B12.shape, *h.B21.shape)


def checkdims(h):  # src/hssmatrix.jl:308-322
    if h.leafnode:
        return h.D.shape[0] == h.U.shape[0] and h.D.shape[1] == h.V.shape[0]
    c1, c2 = checkdims(h.A11), checkdims(h.A22)
    r1, w1 = gensize(h.A11)
    r2, w2 = gensize(h.A22)
    ok = r1 == h.R1.shape[0] and r2 == h.R2.shape[0] and w1 == h.W1.shape[0] and w2 == h.W2.shape[0]
    return ok and c1 and c2


def nleaves(h):
    return 1 if h.leafnode else nleaves(h.A11) + nleaves(h.A22)


def depth(h):
    return 0 if h.leafnode else 1 + max(depth(h.A11), depth(h.A22))


def full(h):
    """Dense expansion, src/hssmatrix.jl:270-305 (equivalently _hssleaf :338-346).
    Returns only the dense matrix of the node treated as root."""
    return _hssleaf(h)[0]


def _hssleaf(h):  # src/hssmatrix.jl:338-346
    if h.leafnode:
        return h.D, h.U, h.V
    A11, U1, V1 = _hssleaf(h.A11)
    A22, U2, V2 = _hssleaf(h.A22)
    A = np.block([[A11, U1 @ h.B12 @ V2.T], [U2 @ h.B21 @ V1.T, A22]])
    U = np.vstack([U1 @ h.R1, U2 @ h.R2])
    V = np.vstack([V1 @ h.W1, V2 @ h.W2])
    return A, U, V


def prune_leaves(h):  # src/hssmatrix.jl:325-335 (prune_leaves!)
    if h.leafnode:
        return h
    if h.A11.leafnode and h.A22.leafnode:
        D, U, V = _hssleaf(h)
        if h.rootnode:
            return hss_leaf(D, rootnode=True)
        return hss_leaf(D, U, V, rootnode=h.rootnode)
    h.A11 = prune_leaves(h.A11)
    h.A22 = prune_leaves(h.A22)
    h.sz1, h.sz2 = size(h.A11), size(h.A22)
    return h


def adjoint(h):  # src/hssmatrix.jl:165-171
    if h.leafnode:
        return hss_leaf(h.D.T.copy(), h.V.copy(), h.U.copy(), rootnode=h.rootnode)
    return hss_branch(adjoint(h.A11), adjoint(h.A22), h.B21.T.copy(), h.B12.T.copy(),
                      h.W1.copy(), h.R1.copy(), h.W2.copy(), h.R2.copy(), rootnode=h.rootnode)


# --------------------------------------------------------------------------
# THE HOT PATH: src/matmul.jl:13-62
# --------------------------------------------------------------------------
class DimensionMismatch(ValueError):
    pass


def matmul(hssA, B):
    """`*(hssA::HssMatrix, B::AbstractMatrix)`, src/matmul.jl:13, and the vector
    wrapper :15."""
    B = _f64(B)
    if B.ndim == 1:  # :15
        return matmul(hssA, B.reshape(-1, 1)).reshape(-1)
    C = np.empty((size(hssA)[0], B.shape[1]))  # similar(): uninitialised
    return mul(C, hssA, B, 1.0, 0.0)


def mul(C, hssA, B, alpha=1.0, beta=0.0, copy_slices=True):
    """`mul!(C, hssA, B, α, β)`, src/matmul.jl:18-28.  `copy_slices=True` makes
    the upsweep copy the row slices of B at every level as matmul.jl:37-38
    does (only matters for timing)."""
    if size(hssA)[1] != B.shape[0]:  # :19
        raise DimensionMismatch("First dimension of B does not match second dimension of A.")
    if C.shape != (size(hssA)[0], B.shape[1]):  # :20
        raise DimensionMismatch("Dimensions of C don't match up with A and B.")
    if hssA.leafnode:  # :21-22
        _gemm(C, hssA.D, B, alpha, beta)
        return C
    hssA = rooted(hssA)  # :24
    Z = _matmatup(hssA, B, copy_slices)  # :25
    _matmatdown(C, hssA, B, Z, None, alpha, beta)  # :26
    return C


def _gemm(C, A, B, alpha, beta):
    """BLAS dgemm semantics of LinearAlgebra.mul!(C, A, B, α, β): β == 0 never
    reads C (so NaNs in uninitialised C do not propagate)."""
    if beta == 0.0:
        np.matmul(A, B, out=C)
        if alpha != 1.0:
            C *= alpha
    else:
        C *= beta
        C += alpha * (A @ B)


def _matmatup(h, B, copy_slices=True):
    """Post-order upsweep, src/matmul.jl:32-42.  Returns (data, left, right)."""
    if h.leafnode:
        return (h.V.T @ B, None, None)  # :34
    n1 = h.sz1[1]
    if copy_slices:  # :37-38: B[1:n1,:] is a copying slice in Julia
        B1, B2 = B[:n1, :].copy(), B[n1:, :].copy()
    else:
        B1, B2 = B[:n1, :], B[n1:, :]
    Z1 = _matmatup(h.A11, B1, copy_slices)
    Z2 = _matmatup(h.A22, B2, copy_slices)
    return (h.W1.T @ Z1[0] + h.W2.T @ Z2[0], Z1, Z2)  # :39


def _matmatdown(C, h, B, Z, F, alpha, beta):
    """Pre-order downsweep, src/matmul.jl:44-62."""
    if h.leafnode:
        _gemm(C, h.D, B, alpha, beta)  # :46
        if F is not None:
            C += alpha * (h.U @ F)  # :47  mul!(C, U, F, α, 1.)
        return C
    m1, n1 = h.sz1
    if F is not None:  # :51-53
        F1 = h.B12 @ Z[2][0] + h.R1 @ F
        F2 = h.B21 @ Z[1][0] + h.R2 @ F
    else:  # :55-56
        F1 = h.B12 @ Z[2][0]
        F2 = h.B21 @ Z[1][0]
    _matmatdown(C[:m1, :], h.A11, B[:n1, :], Z[1], F1, alpha, beta)  # :58
    _matmatdown(C[m1:, :], h.A22, B[n1:, :], Z[2], F2, alpha, beta)  # :59
    return C


# --------------------------------------------------------------------------
# Accounting used by bench.py (SURVEY.md §8d general forms)
# --------------------------------------------------------------------------
def algorithmic_counts(h, nrhs, beta_nonzero=False):
    """(bytes, flops) of one product: every generator read once, X read once,
    Y written once (+ read once if beta != 0); Z/F workspaces excluded."""
    gen = [0]
    fl = [0]

    def rec(t, isroot):
        if t.leafnode:
            m, n = t.D.shape
            kr, kw = (0, 0) if isroot else (t.U.shape[1], t.V.shape[1])
            gen[0] += m * n + m * kr + n * kw
            fl[0] += 2 * nrhs * (m * n + n * kw + m * kr)
            return
        rec(t.A11, False)
        rec(t.A22, False)
        gen[0] += t.B12.size + t.B21.size
        fl[0] += 2 * nrhs * (t.B12.size + t.B21.size)
        if not isroot:
            gen[0] += t.R1.size + t.R2.size + t.W1.size + t.W2.size
            fl[0] += 2 * nrhs * (t.R1.size + t.R2.size + t.W1.size + t.W2.size)

    rec(h, True)
    m, n = size(h)
    byts = 8 * (gen[0] + n * nrhs + m * nrhs * (2 if beta_nonzero else 1))
    return byts, fl[0]


# --------------------------------------------------------------------------
# Synthetic random-generator HSS matrices (BASELINE configs 3-5).
# The generator is a counter-based hash so that the CUDA library can produce
# the SAME bits on the device (csrc/hssb_synth.cuh) without 64 GB crossing
# PCIe.  Integer arithmetic + one int->double conversion + one multiply by a
# constant: bit-exact between numpy, C and CUDA.
# --------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
KIND_D, KIND_U, KIND_V, KIND_B12, KIND_B21, KIND_R, KIND_W, KIND_X = range(8)
# 1/sqrt(Var) of the sum of four independent uniform integers on [0, 65535]
IH4_SCALE = float(np.sqrt(3.0 / (65536.0 * 65536.0 - 1.0)))


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_key(seed, heap_id, kind):
    """Stream key for block `kind` of the node with heap index `heap_id`
    (root = 1, children 2i and 2i+1)."""
    with np.errstate(over="ignore"):
        inner = _splitmix64(np.uint64(heap_id) * np.uint64(8) + np.uint64(kind))
        return _splitmix64(np.uint64(seed) ^ inner)


def synth_values(key, start, count, scale=1.0):
    """Elements [start, start+count) of stream `key`: zero-mean unit-variance
    Irwin-Hall(4) variates times `scale`."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + count, dtype=np.uint64)
        h = _splitmix64(np.uint64(key) + idx)
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + (h >> np.uint64(48))).astype(np.int64) - 131070
    c = np.float64(IH4_SCALE) * np.float64(scale)
    return s.astype(np.float64) * c


def synth_block(seed, heap_id, kind, rows, cols, scale=1.0):
    """rows x cols block, element (i, j) = stream element j*rows + i."""
    v = synth_values(synth_key(seed, heap_id, kind), 0, rows * cols, scale)
    return np.asfortranarray(v.reshape(cols, rows).T)


def synth_x(seed, n, nrhs, row0=0, rows=None):
    """Rows [row0, row0+rows) of the n x nrhs synthetic right-hand side;
    element (i, j) = stream element j*n + i of (heap_id 0, KIND_X)."""
    rows = n - row0 if rows is None else rows
    key = synth_key(seed, 0, KIND_X)
    X = np.empty((rows, nrhs), order="F")
    for j in range(nrhs):
        X[:, j] = synth_values(key, j * n + row0, rows)
    return X


def synthetic_hss(n, leafsize, rank, seed, heap_id=1, lo=0, hi=None, rootnode=True):
    """Perfect-or-bisected tree from `_bisection_cluster` on n with every rank
    equal to `rank`: D, U, V, B12, B21 unit variance;
    R, W scaled by 1/sqrt(2*rank).  `heap_id/lo/hi` select a subtree so a
    bounded sample of a huge matrix can be built on the host."""
    hi = n if hi is None else hi
    tscale = 1.0 / np.sqrt(2.0 * rank) if rank > 0 else 1.0

    def rec(lo, hi, hid, isroot):
        length = hi - lo
        if length <= leafsize:
            D = synth_block(seed, hid, KIND_D, length, length)
            if isroot:
                return hss_leaf(D, rootnode=True)
            U = synth_block(seed, hid, KIND_U, length, rank)
            V = synth_block(seed, hid, KIND_V, length, rank)
            return hss_leaf(D, U, V, rootnode=False)
        nl = -(-length // 2)
        A11 = rec(lo, lo + nl, 2 * hid, False)
        A22 = rec(lo + nl, hi, 2 * hid + 1, False)
        B12 = synth_block(seed, hid, KIND_B12, rank, rank)
        B21 = synth_block(seed, hid, KIND_B21, rank, rank)
        if isroot:
            return hss_branch(A11, A22, B12, B21, rootnode=True)
        # translators of the CHILDREN are stored in the parent (R1/W1, R2/W2);
        # their streams are keyed on the child's heap id.
        R1 = synth_block(seed, 2 * hid, KIND_R, rank, rank, tscale)
        W1 = synth_block(seed, 2 * hid, KIND_W, rank, rank, tscale)
        R2 = synth_block(seed, 2 * hid + 1, KIND_R, rank, rank, tscale)
        W2 = synth_block(seed, 2 * hid + 1, KIND_W, rank, rank, tscale)
        return hss_branch(A11, A22, B12, B21, R1, W1, R2, W2, rootnode=False)

    return rec(lo, hi, heap_id, rootnode)


def synthetic_counts(n, leafsize, rank, nrhs):
    """Closed forms of SURVEY §8d for a perfect tree: (bytes, flops)."""
    L = n // leafsize
    m, r, k = leafsize, rank, nrhs
    if L == 1:
        return 8 * (m * m + 2 * n * k), 2 * m * m * k
    byts = 8 * (L * (m * m + 2 * m * r) + (L - 2) * 6 * r * r + 2 * r * r) + 16 * n * k
    flops = L * 2 * m * k * (m + 2 * r) + (L - 2) * 12 * r * r * k + 4 * r * r * k
    return byts, flops


# --------------------------------------------------------------------------
# Fixture builders that PRODUCE HssMatrix inputs (not on the hot path).
# --------------------------------------------------------------------------
def _compress_block(A, atol, rtol):
    """src/compression.jl:22-30: Q, R[:, invperm(p)] of a column-pivoted QR
    truncated at tolerance.  LowRankApprox.pqrfact (compat 0.4/0.5, not vendored
    under /root/reference) decides the rank; restated as: keep |R_kk| >
    max(atol, rtol*|R_11|).  PARITY UNPINNED for ranks (see module header); the
    hot path is defined on a *given* set of generators so this does not matter."""
    m, n = A.shape
    if m == 0 or n == 0:
        return np.zeros((m, 0)), np.zeros((0, n))
    Q, R, p = _sla.qr(A, mode="economic", pivoting=True)
    d = np.abs(np.diag(R))
    tol = max(atol, rtol * d[0]) if d.size else 0.0
    rk = int(np.sum(d > tol))
    Rp = np.empty((rk, n))
    Rp[:, p] = R[:rk, :]
    return Q[:, :rk].copy(), Rp


def compress(A, rcl, ccl, atol=1e-9, rtol=1e-9):
    """Direct HSS compression, src/compression.jl:47-133."""
    A = _f64(A)
    m, n = len(rcl), len(ccl)
    if A.shape != (m, n):
        raise ValueError("size of row- and column-cluster-trees must match")
    Brow = np.zeros((m, 0))
    Bcol = np.zeros((0, n))
    h, _, _ = _compress(A, Brow, Bcol, rcl, ccl, atol, rtol, True)
    return h


def _compress(A, Brow, Bcol, rcl, ccl, atol, rtol, rootnode):
    if rcl.isleaf() and ccl.isleaf():  # leaf: src/compression.jl:66-74
        r0, r1 = rcl.data
        c0, c1 = ccl.data
        if rootnode:
            return hss_leaf(A[r0:r1, c0:c1].copy(), rootnode=True), Brow, Bcol
        U, Brow = _compress_block(Brow, atol, rtol)
        V, BcolT = _compress_block(Bcol.T, atol, rtol)
        return hss_leaf(A[r0:r1, c0:c1].copy(), U, V), Brow, BcolT.T.copy()
    if not (rcl.isbranch() and ccl.isbranch()):
        raise ValueError("row and column clusters are not compatible")
    # branch: src/compression.jl:77-133
    m1, m2 = len(rcl.left), len(rcl.right)
    n1, n2 = len(ccl.left), len(ccl.right)
    (ra, rb), (rc, rd) = rcl.left.data, rcl.right.data
    (ca, cb), (cc, cd) = ccl.left.data, ccl.right.data
    Brow1 = np.hstack([A[ra:rb, cc:cd], Brow[:m1, :]])  # :86
    Bcol1 = np.vstack([A[rc:rd, ca:cb], Bcol[:, :n1]])  # :87
    A11, Brow1, Bcol1 = _compress(A, Brow1, Bcol1, rcl.left, ccl.left, atol, rtol, False)
    Brow2 = np.hstack([Bcol1[:m2, :], Brow[m1:, :]])  # :97
    Bcol2 = np.vstack([Brow1[:, :n2], Bcol[:, n1:]])  # :98
    A22, Brow2, Bcol2 = _compress(A, Brow2, Bcol2, rcl.right, ccl.right, atol, rtol, False)
    rm1, rn1 = Brow1.shape[0], Bcol1.shape[1]  # :108-109
    B12 = Bcol2[:rm1, :].copy()  # :110
    B21 = Brow2[:, :rn1].copy()  # :111
    Brow = np.vstack([Brow1[:, n2:], Brow2[:, rn1:]])  # :114
    Bcol = np.hstack([Bcol1[m2:, :], Bcol2[rm1:, :]])  # :115
    if rootnode:
        return hss_branch(A11, A22, B12, B21, rootnode=True), Brow, Bcol
    R, Brow = _compress_block(Brow, atol, rtol)  # :119
    R1, R2 = R[:rm1, :].copy(), R[rm1:, :].copy()
    W, BcolT = _compress_block(Bcol.T.copy(), atol, rtol)  # :124
    W1, W2 = W[:rn1, :].copy(), W[rn1:, :].copy()
    return hss_branch(A11, A22, B12, B21, R1, W1, R2, W2, rootnode=False), Brow, BcolT.T.copy()


def hss(A, leafsize=64, atol=1e-9, rtol=1e-9):
    """Smart constructor for dense input, src/hssmatrix.jl:71-76."""
    A = _f64(A)
    return compress(A, bisection_cluster(A.shape[0], leafsize), bisection_cluster(A.shape[1], leafsize), atol, rtol)


def lowrank2hss(U, V, rcl, ccl):
    """src/constructors.jl:5-20: HSS form of U*V' with identity translators."""
    U, V = _f64(U), _f64(V)
    k = U.shape[1]
    if k != V.shape[1]:
        raise ValueError("second dimension of U and V must agree")
    eye = np.eye(k)

    def rec(rc, cc, root):
        if rc.isleaf() and cc.isleaf():
            Ur, Vc = U[rc.data[0]:rc.data[1], :], V[cc.data[0]:cc.data[1], :]
            return hss_leaf(Ur @ Vc.T, Ur.copy(), Vc.copy(), rootnode=root)
        if rc.isbranch() and cc.isbranch():
            return hss_branch(rec(rc.left, cc.left, False), rec(rc.right, cc.right, False),
                              eye.copy(), eye.copy(), eye.copy(), eye.copy(), eye.copy(), eye.copy(), rootnode=root)
        raise ValueError("row and column clusters are not compatible")

    return rec(rcl, ccl, True)


def hss_blkdiag(A, rcl, ccl, rootnode=True):
    """src/compression.jl:447-475: block diagonal of A as an HSS matrix of rank 0."""
    A = _f64(A)
    if rcl.isleaf():
        D = A[rcl.data[0]:rcl.data[1], ccl.data[0]:ccl.data[1]].copy()
        if rootnode:
            return hss_leaf(D, rootnode=True)
        return hss_leaf(D, np.zeros((D.shape[0], 0)), np.zeros((D.shape[1], 0)))
    A11 = hss_blkdiag(A, rcl.left, ccl.left, False)
    A22 = hss_blkdiag(A, rcl.right, ccl.right, False)
    z = np.zeros((0, 0))
    if rootnode:
        return hss_branch(A11, A22, z, z, rootnode=True)
    return hss_branch(A11, A22, z, z, z, z, z, z, rootnode=False)


def cauchy_matrix(n=2001, lo=-1.0, hi=1.0, diag=1.0):
    """README.md:17-18: K(x,y) = 1/(x-y), diagonal 1.0, x = lo:step:hi."""
    x = np.linspace(lo, hi, n)
    d = x[:, None] - x[None, :]
    with np.errstate(divide="ignore"):
        A = 1.0 / d
    A[np.arange(n), np.arange(n)] = diag
    return A


def random_hss(cl_rows, cl_cols, rng, rmin=1, rmax=6, rootnode=True):
    """Random HSS matrix with variable, independently drawn row/column ranks on
    arbitrary (compatible) cluster trees; used to fuzz edge shapes."""
    def rk():
        return int(rng.integers(rmin, rmax + 1))

    def rec(rc, cc, root):
        if rc.isleaf() != cc.isleaf():
            raise ValueError("row and column clusters are not compatible")
        m, n = len(rc), len(cc)
        if rc.isleaf():
            D = rng.standard_normal((m, n))
            if root:
                return hss_leaf(D, rootnode=True)
            return hss_leaf(D, rng.standard_normal((m, rk())), rng.standard_normal((n, rk())))
        A11 = rec(rc.left, cc.left, False)
        A22 = rec(rc.right, cc.right, False)
        kr1, kw1 = gensize(A11)
        kr2, kw2 = gensize(A22)
        B12 = rng.standard_normal((kr1, kw2))
        B21 = rng.standard_normal((kr2, kw1))
        if root:
            return hss_branch(A11, A22, B12, B21, rootnode=True)
        kr, kw = rk(), rk()
        s = 1.0 / np.sqrt(max(kr1 + kr2, 1))
        return hss_branch(A11, A22, B12, B21,
                          s * rng.standard_normal((kr1, kr)), s * rng.standard_normal((kw1, kw)),
                          s * rng.standard_normal((kr2, kr)), s * rng.standard_normal((kw2, kw)))

    return rec(cl_rows, cl_cols, rootnode)


# --------------------------------------------------------------------------
# Full-size parity without a full-size oracle run (BASELINE configs 3-5).
# For a right-hand side supported on a few leaves only, Z vanishes outside the
# support's ancestors, so rows of Y for any sampled leaf can be computed by
# generating O(|support| + depth) nodes on demand.  Same recursion as
# src/matmul.jl:32-62, evaluated lazily on the synthetic generator streams.
# --------------------------------------------------------------------------
class LazySyntheticHss:
    def __init__(self, n, leafsize, rank, seed):
        self.n, self.leafsize, self.rank, self.seed = n, leafsize, rank, seed
        self.tscale = 1.0 / np.sqrt(2.0 * rank) if rank > 0 else 1.0

    def blk(self, hid, kind, rows, cols):
        scale = self.tscale if kind in (KIND_R, KIND_W) else 1.0
        return synth_block(self.seed, hid, kind, rows, cols, scale)

    def _split(self, lo, hi):
        return lo + -(-(hi - lo) // 2)

    def zup(self, hid, lo, hi, s_lo, Xs, memo):
        """Z of node hid for X = zeros except rows [s_lo, s_lo+len(Xs)) = Xs; None if zero."""
        if hid in memo:
            return memo[hid]
        s_hi = s_lo + Xs.shape[0]
        if hi <= s_lo or lo >= s_hi:
            memo[hid] = None
            return None
        r = self.rank
        if hi - lo <= self.leafsize:  # matmul.jl:34
            a, b = max(lo, s_lo), min(hi, s_hi)
            V = self.blk(hid, KIND_V, hi - lo, r)
            z = V[a - lo:b - lo, :].T @ Xs[a - s_lo:b - s_lo, :]
        else:  # matmul.jl:39
            mid = self._split(lo, hi)
            z = np.zeros((r, Xs.shape[1]))
            for c, (clo, chi) in ((2 * hid, (lo, mid)), (2 * hid + 1, (mid, hi))):
                zc = self.zup(c, clo, chi, s_lo, Xs, memo)
                if zc is not None:
                    z += self.blk(c, KIND_W, r, r).T @ zc
        memo[hid] = z
        return z

    def rows(self, targets, s_lo, Xs):
        """{leaf_row0: Y rows of that leaf} for leaves whose first row is in `targets`."""
        out, memo = {}, {}
        k, r = Xs.shape[1], self.rank
        s_hi = s_lo + Xs.shape[0]

        def down(hid, lo, hi, F):
            if not any(lo <= t < hi for t in targets):
                return
            if hi - lo <= self.leafsize:  # matmul.jl:46-47
                y = np.zeros((hi - lo, k))
                a, b = max(lo, s_lo), min(hi, s_hi)
                if a < b:
                    D = self.blk(hid, KIND_D, hi - lo, hi - lo)
                    y += D[:, a - lo:b - lo] @ Xs[a - s_lo:b - s_lo, :]
                if F is not None:
                    y += self.blk(hid, KIND_U, hi - lo, r) @ F
                out[lo] = y
                return
            mid = self._split(lo, hi)
            zl = self.zup(2 * hid, lo, mid, s_lo, Xs, memo)
            zr = self.zup(2 * hid + 1, mid, hi, s_lo, Xs, memo)
            F1 = np.zeros((r, k))
            F2 = np.zeros((r, k))
            if zr is not None:
                F1 += self.blk(hid, KIND_B12, r, r) @ zr  # matmul.jl:52/55
            if zl is not None:
                F2 += self.blk(hid, KIND_B21, r, r) @ zl  # matmul.jl:53/56
            if F is not None:
                F1 += self.blk(2 * hid, KIND_R, r, r) @ F
                F2 += self.blk(2 * hid + 1, KIND_R, r, r) @ F
            down(2 * hid, lo, mid, F1)
            down(2 * hid + 1, mid, hi, F2)

        down(1, 0, self.n, None)
        return out


# --------------------------------------------------------------------------
# Randomized HSS compression (fixture builder for BASELINE config 2: the
# n = 2^16 Cauchy matrix cannot be compressed directly, the reference itself
# would route a matrix-free operator through randcompress, hssmatrix.jl:86).
# Restates src/compression.jl:278-294 (randcompress), :358-430 (_randcompress!)
# and :433-444 (_interpolate).  Not on the hot path.
# --------------------------------------------------------------------------
class KernelOperator:
    """Matrix-free A[i,j] = kernel(x_i, y_j) with a fixed diagonal; stands in for
    the reference's LinearMap (src/linearmap.jl).  Products are evaluated block
    by block (optionally on a torch device to make the fixture cheap)."""

    def __init__(self, x, y, kernel, diag=None, block=4096, device=None):
        self.x, self.y, self.kernel, self.diag, self.block, self.device = x, y, kernel, diag, block, device
        self.shape = (len(x), len(y))

    def getindex(self, I, J):
        I, J = np.asarray(I, dtype=np.int64), np.asarray(J, dtype=np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            B = self.kernel(self.x[I][:, None], self.y[J][None, :])
        if self.diag is not None:
            B = np.where(I[:, None] == J[None, :], self.diag, B)
        return B

    def _apply(self, Om, transpose):
        m, n = self.shape
        rows, cols = (n, m) if transpose else (m, n)
        out = np.empty((rows, Om.shape[1]))
        if self.device is not None:
            import torch
            xs = torch.as_tensor(self.y if transpose else self.x, device=self.device)
            ys = torch.as_tensor(self.x if transpose else self.y, device=self.device)
            Omt = torch.as_tensor(Om, device=self.device)
            for a in range(0, rows, self.block):
                b = min(rows, a + self.block)
                d = (xs[a:b, None] - ys[None, :]) if not transpose else (ys[None, :] - xs[a:b, None])
                Bk = 1.0 / d
                if self.diag is not None:
                    ii = torch.arange(a, b, device=self.device)
                    Bk[ii - a, ii] = self.diag
                out[a:b] = (Bk @ Omt).cpu().numpy()
            return out
        for a in range(0, rows, self.block):
            b = min(rows, a + self.block)
            I = np.arange(a, b)
            Bk = self.getindex(np.arange(cols), I).T if transpose else self.getindex(I, np.arange(cols))
            out[a:b] = Bk @ Om
        return out

    def matmat(self, Om):
        return self._apply(Om, False)

    def rmatmat(self, Om):
        return self._apply(Om, True)


def _interpolate(A, atol, rtol):
    """Interpolative decomposition, src/compression.jl:433-444: A[:, J] * X ~ A."""
    if A.shape[1] == 0:
        return np.zeros((0, 0)), np.zeros(0, dtype=np.int64)
    _, Rm, p = _sla.qr(A, mode="economic", pivoting=True)
    d = np.abs(np.diag(Rm))
    tol = min(atol, rtol * d[0]) if d.size else 0.0
    rk = int(np.sum(d > tol))
    J = p[:rk]
    Xp = _sla.solve_triangular(Rm[:rk, :rk], Rm[:rk, :], lower=False)
    X = np.empty_like(Xp)
    X[:, p] = Xp  # R[1:rk,1:rk] \ R[1:rk, invperm(p)]
    return X, J


def randcompress(Aop, rcl, ccl, kest, atol=1e-9, rtol=1e-9, noversampling=10, rng=None):
    """src/compression.jl:278-294."""
    rng = np.random.default_rng(0) if rng is None else rng
    m, n = Aop.shape
    Om_col = rng.standard_normal((n, kest + noversampling))
    Om_row = rng.standard_normal((m, kest + noversampling))
    Scol = Aop.matmat(Om_col)
    Srow = Aop.rmatmat(Om_row)

    def blkdiag(rc, cc, root):  # hss_blkdiag, :447-475
        if rc.isleaf():
            D = Aop.getindex(np.arange(*rc.data), np.arange(*cc.data))
            return hss_leaf(D, rootnode=True) if root else hss_leaf(D, np.zeros((D.shape[0], 0)), np.zeros((D.shape[1], 0)))
        z = np.zeros((0, 0))
        A11, A22 = blkdiag(rc.left, cc.left, False), blkdiag(rc.right, cc.right, False)
        return hss_branch(A11, A22, z, z, rootnode=True) if root else hss_branch(A11, A22, z, z, z, z, z, z, rootnode=False)

    h = blkdiag(rcl, ccl, True)
    return _randcompress(h, Aop, Scol, Srow, Om_col, Om_row, 0, 0, atol, rtol, True)[0]


def _randcompress(h, Aop, Scol, Srow, Om_col, Om_row, ro, co, atol, rtol, rootnode):
    """src/compression.jl:358-430."""
    if h.leafnode:
        Scol = Scol - h.D @ Om_col  # :360
        Srow = Srow - h.D.T @ Om_row  # :361
        Xcol, Jcol = _interpolate(Scol.T, atol, rtol)  # :366
        h.U = Xcol.T.copy()
        Scol = Scol[Jcol, :]
        U = h.U
        Jcol = ro + Jcol
        Xrow, Jrow = _interpolate(Srow.T, atol, rtol)  # :373
        h.V = Xrow.T.copy()
        Srow = Srow[Jrow, :]
        V = h.V
        Jrow = co + Jrow
        return h, Scol, Srow, Om_col, Om_row, Jcol, Jrow, U, V
    (m1, n1), (m2, n2) = h.sz1, h.sz2
    h.A11, Scol1, Srow1, Oc1, Or1, Jc1, Jr1, U1, V1 = _randcompress(
        h.A11, Aop, Scol[:m1], Srow[:n1], Om_col[:n1], Om_row[:m1], ro, co, atol, rtol, False)
    h.A22, Scol2, Srow2, Oc2, Or2, Jc2, Jr2, U2, V2 = _randcompress(
        h.A22, Aop, Scol[m1:], Srow[n1:], Om_col[n1:], Om_row[m1:], ro + m1, co + n1, atol, rtol, False)
    Oc2, Oc1 = V2.T @ Oc2, V1.T @ Oc1  # :385-386
    Or2, Or1 = U2.T @ Or2, U1.T @ Or1  # :387-388
    Jcol, Jrow = np.concatenate([Jc1, Jc2]), np.concatenate([Jr1, Jr2])
    Om_col, Om_row = np.vstack([Oc1, Oc2]), np.vstack([Or1, Or2])
    h.B12 = Aop.getindex(Jc1, Jr2)  # :395
    h.B21 = Aop.getindex(Jc2, Jr1)  # :396
    Scol = np.vstack([Scol1 - h.B12 @ Oc2, Scol2 - h.B21 @ Oc1])  # :398
    Srow = np.vstack([Srow1 - h.B21.T @ Or2, Srow2 - h.B12.T @ Or1])  # :399
    kr1, kw1 = gensize(h.A11)
    kr2, kw2 = gensize(h.A22)
    if rootnode:  # :401-412
        h.R1, h.R2 = np.zeros((kr1, 0)), np.zeros((kr2, 0))
        h.W1, h.W2 = np.zeros((kw1, 0)), np.zeros((kw2, 0))
        h.rootnode = True
        return h, Scol, Srow, Om_col, Om_row, Jcol, Jrow, np.zeros((kr1 + kr2, 0)), np.zeros((kw1 + kw2, 0))
    Xcol, Jcl = _interpolate(Scol.T, atol, rtol)  # :415
    h.R1, h.R2 = Xcol[:, :Scol1.shape[0]].T.copy(), Xcol[:, Scol1.shape[0]:].T.copy()
    Scol, Jcol = Scol[Jcl, :], Jcol[Jcl]
    U = np.vstack([h.R1, h.R2])
    Xrow, Jrl = _interpolate(Srow.T, atol, rtol)  # :423
    h.W1, h.W2 = Xrow[:, :Srow1.shape[0]].T.copy(), Xrow[:, Srow1.shape[0]:].T.copy()
    Srow, Jrow = Srow[Jrl, :], Jrow[Jrl]
    V = np.vstack([h.W1, h.W2])
    return h, Scol, Srow, Om_col, Om_row, Jcol, Jrow, U, V


def cauchy_operator(n, lo=-1.0, hi=1.0, diag=1.0, device=None):
    """Matrix-free README kernel K(x,y) = 1/(x-y), diagonal `diag`, on n points."""
    x = np.linspace(lo, hi, n)
    return KernelOperator(x, x, lambda a, b: 1.0 / (a - b), diag=diag, device=device)
