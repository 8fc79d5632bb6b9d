"""CPU oracle for `hssA \\ B` of bonevbs/HssMatrices.jl: the implicit ULV
factorisation solver (src/ulvfactor.jl:10-107; `\\` dispatches to it at
src/hssmatrix.jl:234).  SURVEY §8f rank 4.

TEST INFRASTRUCTURE ONLY (same rule as hss_oracle.py): only tests/, smoke() and
bench.py's CPU legs may import it.

PARITY UNPINNED: the reference has no stored vectors for the solver either; its
only check is test/runtests.jl's `norm(x0 - x)/norm(x0)` style residual on the
README matrix.  The restatement is anchored on the identity
full(hssA) @ ulvfactsolve(hssA, b) == b.

LAPACK calls of the reference and their restatement here:
  geqlf!/ormql!  (ulvfactor.jl:38-41)  -> QR of the row/column-reversed block
  gelqf!/ormlq!  (ulvfactor.jl:43-51)  -> QR of the transposed block
The orthogonal factors are unique only up to signs; the solution is not
affected.
"""
from __future__ import annotations

import numpy as np

import hss_oracle as o


def _ql(U):
    """U (m x k, m >= k) = Q @ [0; L] with Q m x m orthogonal, L k x k lower triangular (geqlf!)."""
    Qf, Rf = np.linalg.qr(U[::-1, ::-1], mode="complete")
    return Qf[::-1, ::-1], Rf[::-1, ::-1][U.shape[0] - U.shape[1]:, :]


def _lq(A):
    """A (p x n) = L @ Q with Q n x n orthogonal, L p x n lower trapezoidal (gelqf!)."""
    Qh, Rh = np.linalg.qr(A.T, mode="complete")
    return Rh.T, Qh.T


def _ulvreduce(D, U, V, b):  # src/ulvfactor.jl:22-57
    m, n = D.shape
    k = min(U.shape[1], m)
    nk = min(m - k, n)
    if k >= m:  # :33-37 cannot be compressed
        u = np.zeros((V.shape[1], b.shape[1]))
        zloc = np.zeros((0, b.shape[1]))
        return D, U, V, b, zloc, u, m - k, nk, None
    Q, L = _ql(U)                                # :39
    U = L                                        # :40
    D = Q.T @ D                                  # :41
    b = Q.T @ b                                  # :42
    Ll, Ql = _lq(D[:m - k, :])                   # :44
    L1 = Ll[:, :nk]                              # :45-46
    L2 = D[m - k:, :] @ Ql.T                     # :47
    zloc = np.linalg.solve(L1, b[:m - k, :])     # :48 (trsm)
    b = b[m - k:, :] - L2[:, :nk] @ zloc         # :49
    V = Ql @ V                                   # :50
    u = V[:m - k, :].T @ zloc                    # :51
    D = L2[:, nk:]                               # :53
    V = V[nk:, :]                                # :54
    return D, U, V, b, zloc, u, m - k, nk, Ql


class _QV:  # BinaryNode((cols, lqf...)), ulvfactor.jl:67,92
    __slots__ = ("cols", "Q", "left", "right")

    def __init__(self, cols=None, Q=None):
        self.cols, self.Q, self.left, self.right = cols, Q, None, None


def _ulvfactsolve(h, b, z, co, rootnode=False):  # src/ulvfactor.jl:60-96
    if h.leafnode:
        cols = co + np.arange(o.size(h)[1])
        D, U, V, b, zloc, u, mk, nk, Ql = _ulvreduce(h.D.copy(), h.U.copy(), h.V.copy(), b)
        z[cols[:mk], :] = zloc
        return b, u, D, U, V, cols, nk, _QV(cols, Ql)
    m1, n1 = h.sz1
    b1, u1, D1, U1, V1, cols1, nk1, QV1 = _ulvfactsolve(h.A11, b[:m1, :], z, co)
    b2, u2, D2, U2, V2, cols2, nk2, QV2 = _ulvfactsolve(h.A22, b[m1:, :], z, co + n1)
    b = np.vstack([b1, b2]) - np.vstack([U1 @ h.B12 @ u2, U2 @ h.B21 @ u1])          # :74
    D = np.block([[D1, U1 @ h.B12 @ V2.T], [U2 @ h.B21 @ V1.T, D2]])                  # :75
    cols = np.concatenate([cols1[nk1:], cols2[nk2:]])                                # :76
    U = np.vstack([U1 @ h.R1, U2 @ h.R2])                                            # :78
    V = np.vstack([V1 @ h.W1, V2 @ h.W2])                                            # :79
    if rootnode:
        z[cols, :] = np.linalg.solve(D, b)                                           # :83
        qv = _QV()
        u = np.zeros((0, b.shape[1]))
        nk = D.shape[1]
    else:
        D, U, V, b, zloc, u, mk, nk, Ql = _ulvreduce(D, U, V, b)                     # :88
        u = u + h.W1.T @ u1 + h.W2.T @ u2                                            # :89
        z[cols[:mk], :] = zloc                                                       # :90
        qv = _QV(cols, Ql)
    qv.left, qv.right = QV1, QV2
    return b, u, D, U, V, cols, nk, qv


def _topdown(qv, z):  # src/ulvfactor.jl:98-107
    if qv.Q is not None:
        z[qv.cols, :] = qv.Q.T @ z[qv.cols, :]
    if qv.left is not None:
        _topdown(qv.left, z)
    if qv.right is not None:
        _topdown(qv.right, z)
    return z


def ulvfactsolve(h, b):  # src/ulvfactor.jl:10-19
    b = np.array(b, dtype=np.float64)
    if b.ndim == 1:
        return ulvfactsolve(h, b.reshape(-1, 1)).reshape(-1)
    if h.leafnode:
        return np.linalg.solve(h.D, b)
    z = np.zeros((o.size(h)[1], b.shape[1]))
    *_, qv = _ulvfactsolve(h, b.copy(), z, 0, rootnode=True)
    return _topdown(qv, z)
