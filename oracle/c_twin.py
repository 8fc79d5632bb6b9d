"""ctypes front end of oracle/hss_oracle.c (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhss_oracle.so")


class _Node(C.Structure):
    _fields_ = [("leaf", C.c_int), ("left", C.c_int), ("right", C.c_int), ("m", C.c_int), ("n", C.c_int),
                ("kr", C.c_int), ("kw", C.c_int)] + [(k, C.c_void_p) for k in
                                                      ("D", "U", "V", "B12", "B21", "R1", "W1", "R2", "W2")]


def build():
    src = os.path.join(_HERE, "hss_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hsso_mul.restype = C.c_int
        L.hsso_mul.argtypes = [C.POINTER(_Node), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long,
                               C.c_void_p, C.c_long, C.c_double, C.c_double]
        L.hsso_synth_values.restype = None
        L.hsso_synth_values.argtypes = [C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_ulonglong, C.c_long, C.c_double,
                                        C.c_void_p]
        _lib = L
    return _lib


def _flatten(h):
    """Post-order node array + keep-alive list of column-major copies."""
    nodes, keep = [], []

    def f(a):
        a = np.asfortranarray(np.asarray(a, dtype=np.float64))
        keep.append(a)
        return a.ctypes.data if a.size else None

    def rec(t):
        nd = _Node()
        if t.leafnode:
            nd.leaf = 1
            nd.left = nd.right = -1
            nd.m, nd.n = t.D.shape
            nd.kr, nd.kw = t.U.shape[1], t.V.shape[1]
            nd.D, nd.U, nd.V = f(t.D), f(t.U), f(t.V)
        else:
            l, r = rec(t.A11), rec(t.A22)
            nd.leaf = 0
            nd.left, nd.right = l, r
            nd.m = t.sz1[0] + t.sz2[0]
            nd.n = t.sz1[1] + t.sz2[1]
            nd.kr, nd.kw = t.R1.shape[1], t.W1.shape[1]
            nd.B12, nd.B21 = f(t.B12), f(t.B21)
            nd.R1, nd.W1, nd.R2, nd.W2 = f(t.R1), f(t.W1), f(t.R2), f(t.W2)
        nodes.append(nd)
        return len(nodes) - 1

    root = rec(h)
    return (_Node * len(nodes))(*nodes), root, keep


def mul(Cm, h, B, alpha=1.0, beta=0.0):
    """mul!(C, hssA, B, alpha, beta) through the plain-C twin; C column-major in place."""
    arr, root, keep = _flatten(h)
    Bf = np.asfortranarray(np.asarray(B, dtype=np.float64))
    assert Cm.flags.f_contiguous and Cm.dtype == np.float64
    rc = lib().hsso_mul(arr, len(arr), root, Cm.shape[0], Bf.shape[0], Bf.shape[1], Bf.ctypes.data, max(Bf.shape[0], 1),
                        Cm.ctypes.data, max(Cm.shape[0], 1), alpha, beta)
    if rc == -2:
        raise ValueError("DimensionMismatch")
    assert rc == 0
    return Cm


def synth_values(seed, heap_id, kind, start, count, c):
    out = np.empty(count)
    lib().hsso_synth_values(seed, heap_id, kind, start, count, c, out.ctypes.data)
    return out
