"""hssmatrices.jl_b200 — host side of the B200-native HSS x dense product.

Python mirror of the slice of HssMatrices.jl's interface that sits on the hot
path `hssA * X` / `mul!(C, hssA, X, alpha, beta)` (reference src/matmul.jl:13-62),
above the C ABI of include/hssb200.h (libhssb200.so, hand-written CUDA for
sm_100a).  Julia is not available in the build image, so this module plays the
role of the Julia binding (julia/HssMatricesB200.jl) for tests and benchmarks:
same names, same argument meaning, same error behaviour.

There is NO CPU fallback: every product goes through the CUDA library and
raises if the library or a B200 is missing.  (The directory name contains a
dot, so import it through the repo-root shim: `import hssb200`.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSSB200_LIB") or os.path.join(_HERE, "lib", "libhssb200.so")  # same override as the Julia binding

__all__ = [
    "HssMatrix", "PackedHss", "DimensionMismatch", "HssbError", "bisection_cluster", "ClusterTree",
    "isleaf", "isbranch", "size", "gensize", "rooted", "checkdims", "pack", "mul_", "ulvfactsolve", "synthetic",
    "lib", "device_count", "measure_peak", "load",
]


# --------------------------------------------------------------------------
# library loading
# --------------------------------------------------------------------------
class HssbError(RuntimeError):
    """Any failure reported by libhssb200 other than a dimension mismatch."""


class SingularException(ArithmeticError):
    """Julia's LinearAlgebra.SingularException (ulvfactor.jl:83 `D \\ b` on a singular reduced block): HSSB_ERR_SINGULAR."""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (src/matmul.jl:19-20, src/hssmatrix.jl:58-59)."""


class _Info(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "m", "n", "local_m", "local_n", "local_row0", "local_col0", "n_nodes", "n_leaves", "depth",
        "max_leaf_m", "max_leaf_n", "max_rank", "pool_bytes", "gen_elems", "flops_per_rhs", "z_rows", "f_rows")] + [
        (n, C.c_int32) for n in ("shard_rank", "n_shards", "device", "uniform")]


class _NodeT(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "left", "right", "parent", "depth", "is_leaf", "is_remote", "row0", "m", "col0", "n", "kr", "kw")]


class _TaskT(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "a0", "a1", "b0", "b1", "c", "lda0", "lda1", "ldb0", "ldb1", "ldc", "M", "K0", "K1",
        "ta0", "ta1", "sb0", "sb1", "sc", "epilogue")]


class _BushOpT(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "task", "bush", "level", "m0", "mr", "s0", "s1", "sc", "lds0", "lds1", "ldsc", "to_global", "sa0", "sa1")]


class _BushStageT(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("bush", "kind", "src", "dst", "count", "ld", "level")]


class _PhaseTime(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "kind", "level", "top", "fast", "ntasks", "flops_per_rhs", "gen_elems", "x_rows", "y_rows")] + [("ms", C.c_double)]


class _UlvInfo(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("supported", "factored", "pool_bytes", "flops_per_rhs", "z_rows", "f_rows")]


class _PhaseT(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "kind", "task0", "ntasks", "maxM", "level", "top", "fast", "xchg_zoff", "xchg_slot_rows", "transposed")]


_lib = None
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_i64 = C.c_int64

# every symbol include/hssb200.h declares: (restype, argtypes)
SIGNATURES = {
    "hssb_version": (C.c_int, []),
    "hssb_last_error": (C.c_char_p, []),
    "hssb_device_count": (C.c_int, []),
    "hssb_builder_create": (C.c_int, [C.POINTER(_P)]),
    "hssb_builder_destroy": (None, [_P]),
    "hssb_builder_add_leaf": (_i64, [_P, _i64, _i64, _i64, _i64, _P, _i64, _P, _i64, _P, _i64]),
    "hssb_builder_add_branch": (_i64, [_P, _i64, _i64, _i64, _i64] + [_P, _i64] * 6),
    "hssb_builder_add_remote": (_i64, [_P, _i64, _i64, _i64, _i64]),
    "hssb_builder_finalize": (C.c_int, [_P, _i64, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "hssb_create_synthetic": (C.c_int, [_i64, _i64, _i64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "hssb_synthetic_rhs": (C.c_int, [C.c_uint64, _i64, _i64, _i64, _i64, _P, _i64, C.c_int, _P]),
    "hssb_save": (C.c_int, [_P, C.c_char_p]),
    "hssb_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "hssb_destroy": (C.c_int, [_P]),
    "hssb_info": (C.c_int, [_P, C.POINTER(_Info)]),
    "hssb_node_info": (C.c_int, [_P, _i64, C.POINTER(_NodeT)]),
    "hssb_get_block": (C.c_int, [_P, _i64, C.c_int, _P, _i64]),
    "hssb_reserve": (C.c_int, [_P, _i64]),
    "hssb_matmul": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double]),
    "hssb_matmul_dev": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double, _P]),
    "hssb_matmul_t": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double]),
    "hssb_matmul_t_dev": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double, _P]),
    "hssb_sync": (C.c_int, [_P]),
    "hssb_set_option": (C.c_int, [_P, C.c_int, _i64]),
    "hssb_get_option": (_i64, [_P, C.c_int]),
    "hssb_launch_count": (_i64, [_P]),
    "hssb_phase_count": (C.c_int, [_P]),
    "hssb_phase_time": (C.c_int, [_P, C.c_int, C.POINTER(_PhaseTime)]),
    "hssb_comm_unique_id": (C.c_int, [_P]),
    "hssb_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "hssb_xchg_export": (C.c_int, [_P, _P]),
    "hssb_xchg_import": (C.c_int, [_P, _P, C.c_int]),
    "hssb_measure_peak": (C.c_int, [C.c_int, C.c_int, _i64, C.POINTER(C.c_double)]),
    "hssb_plan_only": (C.c_int, [_P, _i64, C.c_int, C.c_int, C.POINTER(_P)]),
    "hssb_plan_only_synthetic": (C.c_int, [_i64, _i64, _i64, C.c_uint64, C.c_int, C.c_int, C.POINTER(_P)]),
    "hssb_debug_counts": (C.c_int, [_P, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "hssb_debug_task": (C.c_int, [_P, _i64, C.POINTER(_TaskT)]),
    "hssb_debug_phase": (C.c_int, [_P, _i64, C.POINTER(_PhaseT)]),
    "hssb_debug_pool": (C.c_int, [_P, _P, _i64]),
    "hssb_debug_pool_t": (C.c_int, [_P, _P, _i64]),
    "hssb_debug_tree_trace": (C.c_int, [_P, _i64, _P, C.c_int]),
    "hssb_debug_bush_counts": (C.c_int, [_P, C.c_int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "hssb_debug_bush_op": (C.c_int, [_P, C.c_int, _i64, C.POINTER(_BushOpT)]),
    "hssb_debug_bush_deps": (_i64, [_P, C.c_int, _i64, _P, _i64]),
    "hssb_debug_bush_stage": (_i64, [_P, C.c_int, _i64, C.POINTER(_BushStageT)]),
    "hssb_debug_bush_trace": (_i64, [_P, C.c_int, _P, _i64]),
    "hssb_ulv_factor": (C.c_int, [_P]),
    "hssb_solve": (C.c_int, [_P, _i64, _i64, _P, _i64, _P, _i64]),
    "hssb_solve_dev": (C.c_int, [_P, _i64, _i64, _P, _i64, _P, _i64, _P]),
    "hssb_solve_t": (C.c_int, [_P, _i64, _i64, _P, _i64, _P, _i64]),
    "hssb_solve_t_dev": (C.c_int, [_P, _i64, _i64, _P, _i64, _P, _i64, _P]),
    "hssb_ulv_info": (C.c_int, [_P, _P]),
    "hssb_debug_ulv_factor_host": (C.c_int, [_P]),
    "hssb_debug_ulv_pool": (C.c_int, [_P, _P, _i64]),
    "hssb_group_finalize": (C.c_int, [_P, _i64, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "hssb_group_create_synthetic": (C.c_int, [_i64, _i64, _i64, C.c_uint64, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "hssb_group_destroy": (C.c_int, [_P]),
    "hssb_group_size": (C.c_int, [_P]),
    "hssb_group_shard": (_P, [_P, C.c_int]),
    "hssb_group_reserve": (C.c_int, [_P, _i64]),
    "hssb_group_matmul": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double]),
    "hssb_group_matmul_t": (C.c_int, [_P, _i64, _i64, _i64, _P, _i64, _P, _i64, C.c_double, C.c_double]),
    "hssb_group_matmul_dev": (C.c_int, [_P, _i64, C.POINTER(_P), _i64, C.POINTER(_P), _i64, C.c_double, C.c_double, C.POINTER(_P)]),
    "hssb_group_sync": (C.c_int, [_P]),
}

OPT_FORCE_GENERIC, OPT_USE_GRAPH, OPT_PROFILE, OPT_DEBUG, OPT_PIPELINE_COLS, OPT_ADJOINT_TWIN, OPT_ULV_FAST, OPT_TREE_KERNEL, OPT_HOST_BOUNCE, OPT_LAST_BOUNCE, OPT_LEAF_KERNEL, OPT_LEAF_FUSION, OPT_FLOW_KERNEL, OPT_HOST_THREADS, OPT_BUSH_KERNEL, OPT_BUSH_LEVELS, OPT_PDL, OPT_LAST_FACTOR_US = 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19
PHASE_NAMES = ("leaf_up", "merge", "exchange", "translate", "leaf_down", "exchange_ack")
KIND_NAMES = ("D", "U", "V", "B12", "B21", "R", "W")


def lib():
    """The loaded libhssb200.so (raises HssbError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HssbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(hssb200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc >= 0:
        return rc
    msg = lib().hssb_last_error().decode("utf-8", "replace")
    if rc == -2:
        raise DimensionMismatch(msg)
    if rc == -7:
        raise SingularException(msg)
    raise HssbError(f"hssb200 error {rc}: {msg}")


def device_count():
    return lib().hssb_device_count()


def measure_peak(kind, arg, device=0):
    """kind 0: DFMA TFLOP/s, 1: DMMA TFLOP/s, 2: copy GB/s over `arg` bytes."""
    out = C.c_double(0.0)
    _check(lib().hssb_measure_peak(device, kind, arg, C.byref(out)))
    return out.value


# --------------------------------------------------------------------------
# cluster trees: src/clustertree.jl:14-35
# --------------------------------------------------------------------------
class ClusterTree:
    """BinaryNode{UnitRange{Int}} (src/binarytree.jl:7-20); `data` = (lo, hi),
    0-based half-open."""

    __slots__ = ("data", "left", "right")

    def __init__(self, data, left=None, right=None):
        self.data, self.left, self.right = data, left, right

    def isleaf(self):
        return self.left is None and self.right is None

    def isbranch(self):
        return self.left is not None and self.right is not None

    def __len__(self):
        return self.data[1] - self.data[0]


def bisection_cluster(n, leafsize=64):
    """src/clustertree.jl:14-35: split while len > leafsize; left = ceil(len/2)."""
    lo, hi = (0, int(n)) if np.isscalar(n) else (int(n[0]), int(n[1]))
    if hi - lo <= 0:
        raise ValueError("Index range must be larger or equal to 0")
    if leafsize < 1:
        raise ValueError("leafsize must be >= 1")

    def rec(lo, hi):
        node = ClusterTree((lo, hi))
        if hi - lo > leafsize:
            nl = -(-(hi - lo) // 2)
            node.left, node.right = rec(lo, lo + nl), rec(lo + nl, hi)
        return node

    return rec(lo, hi)


# --------------------------------------------------------------------------
# HssMatrix: the input type (src/hssmatrix.jl:11-68), fields as in the reference
# --------------------------------------------------------------------------
def _f64(a):
    return np.asarray(a, dtype=np.float64)


class HssMatrix:
    """Recursive generator tree with the reference's field names: leaf `D, U, V`;
    branch `A11, A22, B12, B21, sz1, sz2, R1, W1, R2, W2`; flags `leafnode`,
    `rootnode`.  `hssA @ X` and `mul_(C, hssA, X, alpha, beta)` run on the GPU."""

    def __init__(self):
        self.leafnode = False
        self.rootnode = False
        self.D = self.U = self.V = None
        self.A11 = self.A22 = None
        self.B12 = self.B21 = None
        self.sz1 = self.sz2 = None
        self.R1 = self.W1 = self.R2 = self.W2 = None
        self._packed = None

    # constructors, src/hssmatrix.jl:36-67 -------------------------------
    @staticmethod
    def leaf(D, U=None, V=None, rootnode=None):
        D = _f64(D)
        h = HssMatrix()
        h.leafnode = True
        if U is None and V is None:  # :36-39
            h.rootnode = True if rootnode is None else rootnode
            h.D, h.U, h.V = D, np.zeros((D.shape[0], 0)), np.zeros((D.shape[1], 0))
            return h
        U, V = _f64(U), _f64(V)
        if D.shape[0] != U.shape[0]:  # :41
            raise ValueError("D and U must have same number of rows")
        if D.shape[1] != V.shape[0]:  # :42
            raise ValueError("D and V must have same number of columns")
        h.rootnode = False if rootnode is None else rootnode
        h.D, h.U, h.V = D, U, V
        return h

    @staticmethod
    def branch(A11, A22, B12, B21, R1=None, W1=None, R2=None, W2=None, rootnode=None):
        h = HssMatrix()
        h.A11, h.A22 = A11, A22
        h.B12, h.B21 = _f64(B12), _f64(B21)
        h.sz1, h.sz2 = size(A11), size(A22)
        if R1 is None:  # :46-55
            (kr1, kw1), (kr2, kw2) = gensize(A11), gensize(A22)
            h.rootnode = True if rootnode is None else rootnode
            h.R1, h.W1 = np.zeros((kr1, 0)), np.zeros((kw1, 0))
            h.R2, h.W2 = np.zeros((kr2, 0)), np.zeros((kw2, 0))
            return h
        R1, W1, R2, W2 = _f64(R1), _f64(W1), _f64(R2), _f64(W2)
        if R1.shape[1] != R2.shape[1]:  # :58
            raise DimensionMismatch("R1 and R2 must have same number of columns")
        if W1.shape[1] != W2.shape[1]:  # :59
            raise DimensionMismatch("W1 and W2 must have same number of rows")
        h.rootnode = False if rootnode is None else rootnode
        h.R1, h.W1, h.R2, h.W2 = R1, W1, R2, W2
        return h

    # AbstractMatrix surface used on the path ------------------------------
    @property
    def shape(self):
        return size(self)

    def repack(self, device=0):
        """(Re)build the device-resident packed copy.  HssMatrix is mutable
        (recompress!, prune_leaves!, field assignment as in test/runtests.jl:75),
        so the cache must be refreshed by hand after any mutation."""
        if self._packed is not None:
            self._packed.close()
        self._packed = pack(self, device=device)
        return self._packed

    def __rmatmul__(self, A):
        """`*(A::AbstractMatrix, hssB)` (src/matmul.jl:14) without the reference's adjoint copy."""
        if self._packed is None:
            self.repack()
        return self._packed.__rmatmul__(A)

    __array_priority__ = 1000  # let `ndarray @ HssMatrix` reach __rmatmul__

    def __matmul__(self, B):
        """`*(hssA, B)` (src/matmul.jl:13) and `*(hssA, x::Vector)` (:15)."""
        B = _f64(B)
        if B.ndim == 1:  # :15 reshape(x, length(x), 1) ... reshape back
            return (self @ B.reshape(-1, 1)).reshape(-1)
        Cm = np.empty((size(self)[0], B.shape[1]), order="F")  # similar(): uninitialised
        return mul_(Cm, self, B, 1.0, 0.0)

    def solve(self, B):
        """`hssA \\ B` (src/hssmatrix.jl:234): ULV factorisation once per packed copy, then solves."""
        return ulvfactsolve(self, B)


def isleaf(h):  # src/hssmatrix.jl:88
    return h.leafnode


def isbranch(h):  # src/hssmatrix.jl:89
    return not h.leafnode


def size(h, dim=None):  # src/hssmatrix.jl:94-95
    s = tuple(h.D.shape) if h.leafnode else (h.sz1[0] + h.sz2[0], h.sz1[1] + h.sz2[1])
    return s if dim is None else s[dim]


def gensize(h):  # src/hssmatrix.jl:254-262
    if h.leafnode:
        return h.U.shape[1], h.V.shape[1]
    kr = h.R1.shape[1]
    if kr != h.R2.shape[1]:
        raise DimensionMismatch("dimensions of column-translators do not match")
    kw = h.W1.shape[1]
    if kw != h.W2.shape[1]:
        raise DimensionMismatch("dimensions of row-translators do not match")
    return kr, kw


def rooted(h):  # src/hssmatrix.jl:266
    if h.leafnode:
        return HssMatrix.leaf(h.D, rootnode=True)
    return HssMatrix.branch(h.A11, h.A22, h.B12, h.B21, rootnode=True)


def checkdims(h):  # src/hssmatrix.jl:308-322
    if h.leafnode:
        return h.D.shape[0] == h.U.shape[0] and h.D.shape[1] == h.V.shape[0]
    c1, c2 = checkdims(h.A11), checkdims(h.A22)
    (r1, w1), (r2, w2) = gensize(h.A11), gensize(h.A22)
    ok = r1 == h.R1.shape[0] and r2 == h.R2.shape[0] and w1 == h.W1.shape[0] and w2 == h.W2.shape[0]
    return bool(ok and c1 and c2)


# --------------------------------------------------------------------------
# packer front end: walks the pointer tree, the C++ side flattens it
# --------------------------------------------------------------------------
def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a.size else None


def _fcol(a):
    a = np.asarray(a, dtype=np.float64)
    return a if a.flags.f_contiguous else np.asfortranarray(a)


def _build(b, hssA, shard_rank, n_shards):
    """Post-order registration of every node with the C++ builder."""
    L = lib()
    p = n_shards.bit_length() - 1
    cut = [0]

    def add(h, depth, isroot):
        if n_shards > 1 and depth == p:
            g = cut[0]
            cut[0] += 1
            if g != shard_rank:
                m, n = size(h)
                kr, kw = gensize(h)
                return _check(L.hssb_builder_add_remote(b, m, n, kr, kw))
        if h.leafnode:
            D, U, V = _fcol(h.D), _fcol(h.U), _fcol(h.V)
            m, n = D.shape
            kr, kw = (0, 0) if isroot else (U.shape[1], V.shape[1])
            if U.shape[0] != m or V.shape[0] != n:
                raise DimensionMismatch("leaf generators do not match D (hssmatrix.jl:41-42)")
            return _check(L.hssb_builder_add_leaf(b, m, n, kr, kw, _ptr(D), max(m, 1), _ptr(U), max(m, 1),
                                                  _ptr(V), max(n, 1)))
        left = add(h.A11, depth + 1, False)
        right = add(h.A22, depth + 1, False)
        (kr1, kw1), (kr2, kw2) = gensize(h.A11), gensize(h.A22)
        B12, B21 = _fcol(h.B12), _fcol(h.B21)
        if B12.shape != (kr1, kw2) or B21.shape != (kr2, kw1):
            raise DimensionMismatch("B12/B21 do not match the children's gensize")
        if isroot:
            return _check(L.hssb_builder_add_branch(b, left, right, 0, 0, _ptr(B12), max(kr1, 1), _ptr(B21), max(kr2, 1),
                                                    None, 1, None, 1, None, 1, None, 1))
        kr, kw = gensize(h)
        R1, W1, R2, W2 = _fcol(h.R1), _fcol(h.W1), _fcol(h.R2), _fcol(h.W2)
        if R1.shape[0] != kr1 or R2.shape[0] != kr2 or W1.shape[0] != kw1 or W2.shape[0] != kw2:
            raise DimensionMismatch("translators do not match the children's gensize (hssmatrix.jl:318)")
        return _check(L.hssb_builder_add_branch(
            b, left, right, kr, kw, _ptr(B12), max(kr1, 1), _ptr(B21), max(kr2, 1),
            _ptr(R1), max(kr1, 1), _ptr(W1), max(kw1, 1), _ptr(R2), max(kr2, 1), _ptr(W2), max(kw2, 1)))

    return add(hssA, 0, True)


def pack(hssA, device=0, shard_rank=0, n_shards=1, plan_only=False):
    """Flatten `hssA` (treated as root, like rooted(), src/matmul.jl:24) into the
    level-ordered device-resident format.  With n_shards = P > 1 only the
    shard_rank-th depth-log2(P) subtree and the replicated top tree are packed."""
    if n_shards < 1 or n_shards & (n_shards - 1):
        raise ValueError("n_shards must be a power of two")
    L = lib()
    b = C.c_void_p()
    _check(L.hssb_builder_create(C.byref(b)))
    try:
        root = _build(b, hssA, shard_rank, n_shards)
        h = C.c_void_p()
        if plan_only:
            _check(L.hssb_plan_only(b, root, shard_rank, n_shards, C.byref(h)))
        else:
            _check(L.hssb_builder_finalize(b, root, device, shard_rank, n_shards, C.byref(h)))
    finally:
        L.hssb_builder_destroy(b)
    return PackedHss(h)


def synthetic(n, leafsize, rank, seed, device=0, shard_rank=0, n_shards=1, plan_only=False):
    """Synthetic random-generator HSS matrix of BASELINE.json configs 3-5,
    generated on the device (bit-identical to oracle.synthetic_hss)."""
    h = C.c_void_p()
    if plan_only:
        _check(lib().hssb_plan_only_synthetic(n, leafsize, rank, seed, shard_rank, n_shards, C.byref(h)))
    else:
        _check(lib().hssb_create_synthetic(n, leafsize, rank, seed, device, shard_rank, n_shards, C.byref(h)))
    return PackedHss(h)


def _devs(devices):
    devices = list(devices)
    return (C.c_int * len(devices))(*devices), len(devices)


def pack_group(hssA, devices):
    """One process, several GPUs: shard `hssA` by subtree over `devices` (a power of two of them; entries may
    repeat) behind ONE handle whose `@` / `mul_` take the whole X and Y (hssb_group_*)."""
    L = lib()
    b = C.c_void_p()
    _check(L.hssb_builder_create(C.byref(b)))
    try:
        root = _build(b, hssA, 0, 1)       # the whole tree; every shard copies what it owns
        g = C.c_void_p()
        arr, nd = _devs(devices)
        _check(L.hssb_group_finalize(b, root, arr, nd, C.byref(g)))
    finally:
        L.hssb_builder_destroy(b)
    return PackedGroup(g)


def synthetic_group(n, leafsize, rank, seed, devices):
    """The synthetic benchmark matrix sharded over `devices` in one process."""
    g = C.c_void_p()
    arr, nd = _devs(devices)
    _check(lib().hssb_group_create_synthetic(n, leafsize, rank, seed, arr, nd, C.byref(g)))
    return PackedGroup(g)


class PackedGroup:
    """P sharded handles of one matrix in ONE process (hssb_group*): the Julia drop-in's multi-GPU form."""

    __array_priority__ = 1000

    def __init__(self, handle):
        self._g = handle
        L = lib()
        self.size = L.hssb_group_size(self._g)
        self.shards = []
        for i in range(self.size):
            sh = PackedHss.__new__(PackedHss)
            sh._h = C.c_void_p(L.hssb_group_shard(self._g, i))
            sh._borrowed = True
            sh.info = _Info()
            _check(L.hssb_info(sh._h, C.byref(sh.info)))
            self.shards.append(sh)
        self.m, self.n = self.shards[0].info.m, self.shards[0].info.n

    def close(self):
        if self._g:
            for sh in self.shards:
                sh._h = None
            lib().hssb_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reserve(self, max_nrhs):
        _check(lib().hssb_group_reserve(self._g, max_nrhs))

    def sync(self):
        _check(lib().hssb_group_sync(self._g))

    def set_option(self, opt, value):
        for sh in self.shards:
            sh.set_option(opt, value)

    def launch_count(self):
        return sum(sh.launch_count() for sh in self.shards)

    def mul_(self, Cm, B, alpha=1.0, beta=0.0, trans=False):
        """mul!(C, hssA, B, alpha, beta) on the whole host matrices; every shard works on its row block."""
        B = np.asarray(B, dtype=np.float64)
        if B.ndim != 2 or Cm.ndim != 2:
            raise DimensionMismatch("B and C must be matrices")
        if not (isinstance(Cm, np.ndarray) and Cm.dtype == np.float64 and Cm.flags.f_contiguous and Cm.flags.writeable):
            raise TypeError("C must be a writable column-major float64 array (Julia Matrix{Float64})")
        if Cm.shape[1] != B.shape[1]:
            raise DimensionMismatch("Dimensions of C don't match up with A and B.")
        Bf = _fcol(B)
        fn = lib().hssb_group_matmul_t if trans else lib().hssb_group_matmul
        _check(fn(self._g, Cm.shape[0], Bf.shape[0], Bf.shape[1], _ptr(Bf), max(Bf.shape[0], 1), _ptr(Cm), max(Cm.shape[0], 1),
                  float(alpha), float(beta)))
        return Cm

    def __matmul__(self, B):
        B = _f64(B)
        if B.ndim == 1:
            return (self @ B.reshape(-1, 1)).reshape(-1)
        return self.mul_(np.empty((self.m, B.shape[1]), order="F"), B, 1.0, 0.0)

    def tmatmul(self, B):
        B = _f64(B)
        if B.ndim == 1:
            return self.tmatmul(B.reshape(-1, 1)).reshape(-1)
        return self.mul_(np.empty((self.n, B.shape[1]), order="F"), B, 1.0, 0.0, trans=True)

    def __rmatmul__(self, A):
        A = _f64(A)
        return self.tmatmul(np.asfortranarray(A.T)).T

    def matmul_dev(self, dX, ldx, dY, ldy, nrhs, alpha=1.0, beta=0.0, streams=None):
        """Device pointers (one per shard, on that shard's device) to the local row blocks; asynchronous."""
        px = (C.c_void_p * self.size)(*dX)
        py = (C.c_void_p * self.size)(*dY)
        ps = (C.c_void_p * self.size)(*streams) if streams is not None else None
        _check(lib().hssb_group_matmul_dev(self._g, nrhs, px, ldx, py, ldy, float(alpha), float(beta), ps))


def load(path, device=0):
    """Load a packed matrix written by PackedHss.save(); device=-1 gives a host-only handle."""
    h = C.c_void_p()
    _check(lib().hssb_load(os.fsencode(path), device, C.byref(h)))
    return PackedHss(h)


class PackedHss:
    """Handle to a packed, device-resident HSS matrix (hssb_matrix*)."""

    __array_priority__ = 1000  # let `ndarray @ PackedHss` reach __rmatmul__

    def __init__(self, handle):
        self._h = handle
        self.info = _Info()
        _check(lib().hssb_info(self._h, C.byref(self.info)))

    # lifetime --------------------------------------------------------------
    def close(self):
        if self._h:
            if not getattr(self, "_borrowed", False):
                lib().hssb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def save(self, path):
        """Write the packed format (tree shape + level-ordered pool) to `path`."""
        _check(lib().hssb_save(self._h, os.fsencode(path)))

    # queries ---------------------------------------------------------------
    @property
    def shape(self):
        return (self.info.m, self.info.n)

    @property
    def local_shape(self):
        return (self.info.local_m, self.info.local_n)

    def node(self, i):
        nd = _NodeT()
        _check(lib().hssb_node_info(self._h, i, C.byref(nd)))
        return nd

    def block(self, node, kind):
        """Generator block `kind` (0..6 or 'D','U','V','B12','B21','R','W') of `node`."""
        if isinstance(kind, str):
            kind = KIND_NAMES.index(kind)
        nd = self.node(node)
        par = self.node(nd.parent) if nd.parent >= 0 else None
        leaf = bool(nd.is_leaf) and not nd.is_remote
        branch = not nd.is_leaf and not nd.is_remote
        if kind == 0:
            shp = (nd.m, nd.n) if leaf else (0, 0)
        elif kind == 1:
            shp = (nd.m, nd.kr) if leaf else (0, 0)
        elif kind == 2:
            shp = (nd.n, nd.kw) if leaf else (0, 0)
        elif kind == 3:
            shp = (self.node(nd.left).kr, self.node(nd.right).kw) if branch else (0, 0)
        elif kind == 4:
            shp = (self.node(nd.right).kr, self.node(nd.left).kw) if branch else (0, 0)
        elif kind == 5:
            shp = (nd.kr, par.kr if par else 0)
        else:
            shp = (nd.kw, par.kw if par else 0)
        out = np.zeros(shp, order="F")
        _check(lib().hssb_get_block(self._h, node, kind, _ptr(out), out.size))
        return out

    def launch_count(self):
        return lib().hssb_launch_count(self._h)

    def set_option(self, opt, value):
        _check(lib().hssb_set_option(self._h, opt, int(value)))

    def get_option(self, opt):
        return lib().hssb_get_option(self._h, opt)

    def phase_times(self, nrhs=None):
        """Per-phase accounting (+ device ms of the last profiled call, OPT_PROFILE)."""
        out = []
        for i in range(lib().hssb_phase_count(self._h)):
            t = _PhaseTime()
            _check(lib().hssb_phase_time(self._h, i, C.byref(t)))
            name = PHASE_NAMES[t.kind]
            if t.kind in (1, 3):
                name += ("_top" if t.top else "") + f"_L{t.level}"
            out.append({"name": name, "kind": t.kind, "level": t.level, "top": t.top, "fast": t.fast, "ntasks": t.ntasks,
                        "flops_per_rhs": t.flops_per_rhs, "gen_elems": t.gen_elems, "x_rows": t.x_rows, "y_rows": t.y_rows,
                        "ms": t.ms})
        return out

    def reserve(self, max_nrhs):
        _check(lib().hssb_reserve(self._h, max_nrhs))

    def sync(self):
        _check(lib().hssb_sync(self._h))

    def flops(self, nrhs):
        return self.info.flops_per_rhs * nrhs

    def algorithmic_bytes(self, nrhs, beta_nonzero=False):
        return 8 * (self.info.gen_elems + self.info.local_n * nrhs + self.info.local_m * nrhs * (2 if beta_nonzero else 1))

    # the product -------------------------------------------------------------
    def mul_(self, Cm, B, alpha=1.0, beta=0.0, trans=False):
        """mul!(C, hssA, B, alpha, beta) with host (numpy) arrays; C must be
        column-major (Julia layout) and is updated in place.  trans=True applies A'."""
        B = np.asarray(B, dtype=np.float64)
        if B.ndim != 2 or Cm.ndim != 2:
            raise DimensionMismatch("B and C must be matrices")
        if not (isinstance(Cm, np.ndarray) and Cm.dtype == np.float64 and Cm.flags.f_contiguous and Cm.flags.writeable):
            raise TypeError("C must be a writable column-major float64 array (Julia Matrix{Float64})")
        if Cm.shape[1] != B.shape[1]:  # matmul.jl:20
            raise DimensionMismatch("Dimensions of C don't match up with A and B.")
        Bf = _fcol(B)
        fn = lib().hssb_matmul_t if trans else lib().hssb_matmul
        _check(fn(self._h, Cm.shape[0], Bf.shape[0], Bf.shape[1], _ptr(Bf), max(Bf.shape[0], 1),
                  _ptr(Cm), max(Cm.shape[0], 1), float(alpha), float(beta)))
        return Cm

    def tmatmul(self, B):
        """A' * B on the same packed generators (no adjoint copy, cf. hssmatrix.jl:165-171)."""
        B = _f64(B)
        if B.ndim == 1:
            return self.tmatmul(B.reshape(-1, 1)).reshape(-1)
        Cm = np.empty((self.info.local_n, B.shape[1]), order="F")
        return self.mul_(Cm, B, 1.0, 0.0, trans=True)

    def __rmatmul__(self, A):
        """`*(A::AbstractMatrix, hssB)` (src/matmul.jl:14): A*hssB = (hssB' * A')'."""
        A = _f64(A)
        return self.tmatmul(np.asfortranarray(A.T)).T

    def __matmul__(self, B):
        B = _f64(B)
        if B.ndim == 1:
            return (self @ B.reshape(-1, 1)).reshape(-1)
        Cm = np.empty((self.info.local_m, B.shape[1]), order="F")
        return self.mul_(Cm, B, 1.0, 0.0)

    # the solver: hssA \ B (src/hssmatrix.jl:234 -> ulvfactsolve, src/ulvfactor.jl:10-19) ------------
    @property
    def ulv_info(self):
        o = _UlvInfo()
        _check(lib().hssb_ulv_info(self._h, C.byref(o)))
        return o

    def ulv_factor(self):
        """Implicit ULV factorisation on the device (done once; solve() does it on first use)."""
        _check(lib().hssb_ulv_factor(self._h))

    def solve(self, B):
        """`hssA \\ B` with host (numpy) arrays; returns a new column-major array."""
        B = _f64(B)
        if B.ndim == 1:
            return self.solve(B.reshape(-1, 1)).reshape(-1)
        Bf = _fcol(B)
        Z = np.empty((self.info.n, B.shape[1]), order="F")
        _check(lib().hssb_solve(self._h, Bf.shape[0], Bf.shape[1], _ptr(Bf), max(Bf.shape[0], 1), _ptr(Z), max(Z.shape[0], 1)))
        return Z

    ulvfactsolve = solve

    def solve_t(self, B):
        """`A' \\ B` on the same packed matrix (uniform trees: second factor pool from the adjoint twin pool)."""
        B = _f64(B)
        if B.ndim == 1:
            return self.solve_t(B.reshape(-1, 1)).reshape(-1)
        Bf = _fcol(B)
        Z = np.empty((self.info.m, B.shape[1]), order="F")
        _check(lib().hssb_solve_t(self._h, Bf.shape[0], Bf.shape[1], _ptr(Bf), max(Bf.shape[0], 1), _ptr(Z), max(Z.shape[0], 1)))
        return Z

    def __rtruediv__(self, A):
        """`/(A, hssB)` (src/hssmatrix.jl:236): A / hssB = (hssB' \\ A')'."""
        A = _f64(A)
        return self.solve_t(np.asfortranarray(A.T)).T

    def solve_dev(self, b_ptr, ldb, z_ptr, ldz, nrhs, stream=None):
        """Asynchronous solve on raw device pointers."""
        _check(lib().hssb_solve_dev(self._h, self.info.n, nrhs, b_ptr, ldb, z_ptr, ldz, stream))

    def debug_ulv_pool(self, factor_on_host=False):
        """Factor pool image (tests).  factor_on_host: plan-only handles run the node routine on the host first."""
        if factor_on_host:
            _check(lib().hssb_debug_ulv_factor_host(self._h))
        pool = np.zeros(self.ulv_info.pool_bytes // 8)
        _check(lib().hssb_debug_ulv_pool(self._h, _ptr(pool), pool.size))
        return pool

    def matmul_dev(self, x_ptr, ldx, y_ptr, ldy, nrhs, alpha=1.0, beta=0.0, stream=None, rows_x=None, rows_y=None,
                   trans=False):
        """Asynchronous product on raw device pointers (the timed entry); trans=True applies A'."""
        if rows_x is None:
            rows_x = self.info.local_m if trans else self.info.local_n
        if rows_y is None:
            rows_y = self.info.local_n if trans else self.info.local_m
        fn = lib().hssb_matmul_t_dev if trans else lib().hssb_matmul_dev
        _check(fn(self._h, rows_y, rows_x, nrhs, x_ptr, ldx, y_ptr, ldy, float(alpha), float(beta), stream))

    # multi-GPU ---------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        _check(lib().hssb_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        _check(lib().hssb_comm_init(self._h, buf, rank, n_ranks))

    def xchg_export(self):
        """128-byte blob (two CUDA IPC handles) of this rank's exchange buffers; reserve() first."""
        buf = (C.c_char * 128)()
        _check(lib().hssb_xchg_export(self._h, buf))
        return bytes(buf)

    def xchg_import(self, blobs):
        """`blobs`: the xchg_export() results of all ranks, in rank order."""
        raw = b"".join(blobs)
        buf = (C.c_char * len(raw)).from_buffer_copy(raw)
        _check(lib().hssb_xchg_import(self._h, buf, len(blobs)))

    def tree_trace(self, nrhs):
        """Microseconds per step (level / exchange) of the persistent tree kernel launched alone (diagnostics)."""
        buf = np.zeros(256)
        n = _check(lib().hssb_debug_tree_trace(self._h, nrhs, _ptr(buf), buf.size))
        return buf[:n].tolist()

    # plan export (tests) -------------------------------------------------------
    def debug_pool_t(self):
        """Image of the adjoint twin pool (uniform trees): the forward plan over it computes A' X."""
        pl = _i64()
        _check(lib().hssb_debug_counts(self._h, None, None, C.byref(pl)))
        pool = np.zeros(pl.value)
        _check(lib().hssb_debug_pool_t(self._h, _ptr(pool), pool.size))
        return pool

    def debug_bush_plan(self, mode=0):
        """The bush plan (csrc/hssb_bush.cuh) of the product (mode 0) or of the transposed task table (mode 1):
        (ops, deps per bush, shared-memory doubles per CTA, staged copies)."""
        nb, no, nd, sm = _i64(), _i64(), _i64(), _i64()
        _check(lib().hssb_debug_bush_counts(self._h, mode, C.byref(nb), C.byref(no), C.byref(nd), C.byref(sm)))
        ops = []
        for i in range(no.value):
            o = _BushOpT()
            _check(lib().hssb_debug_bush_op(self._h, mode, i, C.byref(o)))
            ops.append(o)
        deps = []
        for b in range(nb.value):
            buf = np.zeros(max(nd.value, 1), dtype=np.int64)
            n = _check(lib().hssb_debug_bush_deps(self._h, mode, b, _ptr(buf), buf.size))
            deps.append([int(x) for x in buf[:n]])
        stages = []
        for i in range(_check(lib().hssb_debug_bush_stage(self._h, mode, 0, None))):
            st = _BushStageT()
            _check(lib().hssb_debug_bush_stage(self._h, mode, i, C.byref(st)))
            stages.append(st)
        return ops, deps, sm.value, stages

    def debug_bush_trace(self, mode=0, items=None, probe_item=0):
        """Timeline of the last bush-kernel launch (diagnostics): None switches recording on; afterwards an
        (items, 12) array: SM, ns drawn, ns dependencies met, ns end of levels 0..7, ns flag published."""
        if items is None:
            _check(lib().hssb_debug_bush_trace(self._h, mode, None, probe_item))
            return None
        buf = np.zeros((items + 43, 12), dtype=np.uint64)   # + the probe block: 8 levels x 8 warps x 8 cycle stamps
        n = _check(lib().hssb_debug_bush_trace(self._h, mode, _ptr(buf), items + 43))
        self.bush_probe = buf[items:].reshape(-1)[:512].reshape(8, 8, 8).astype(np.int64)
        return buf[:min(n, items)]

    def debug_plan(self):
        nt, nph, pl = _i64(), _i64(), _i64()
        _check(lib().hssb_debug_counts(self._h, C.byref(nt), C.byref(nph), C.byref(pl)))
        tasks, phases = [], []
        for i in range(nt.value):
            t = _TaskT()
            _check(lib().hssb_debug_task(self._h, i, C.byref(t)))
            tasks.append(t)
        for i in range(nph.value):
            p = _PhaseT()
            _check(lib().hssb_debug_phase(self._h, i, C.byref(p)))
            phases.append(p)
        pool = np.zeros(pl.value)
        _check(lib().hssb_debug_pool(self._h, _ptr(pool), pool.size))
        return tasks, phases, pool


def ulvfactsolve(hssA, B):
    """`ulvfactsolve(hssA, b)` (src/ulvfactor.jl:10-19), i.e. `hssA \\ b`, for an HssMatrix or a PackedHss."""
    if isinstance(hssA, PackedHss):
        return hssA.solve(B)
    B = np.asarray(B, dtype=np.float64)
    if size(hssA, 0) != B.shape[0]:
        raise DimensionMismatch(f"First dimension of B ({B.shape[0]}) does not match first dimension of A ({size(hssA, 0)})")
    if hssA._packed is None:
        hssA.repack()
    return hssA._packed.solve(B)


def mul_(Cm, hssA, B, alpha=1.0, beta=0.0):
    """`mul!(C, hssA, B, alpha, beta)` (src/matmul.jl:18-28) for an HssMatrix or
    a PackedHss.  Dimension checks raise DimensionMismatch like :19-20."""
    if isinstance(hssA, PackedHss):
        return hssA.mul_(Cm, B, alpha, beta)
    B = np.asarray(B, dtype=np.float64)
    if size(hssA, 1) != B.shape[0]:  # :19
        raise DimensionMismatch(
            f"First dimension of B does not match second dimension of A. Expected {size(hssA, 1)}, got {B.shape[0]}")
    if tuple(Cm.shape) != (size(hssA, 0), B.shape[1]):  # :20
        raise DimensionMismatch("Dimensions of C don't match up with A and B.")
    if hssA._packed is None:
        hssA.repack()
    return hssA._packed.mul_(Cm, B, alpha, beta)
