"""ctypes loader of oracle/libquant_oracle.so (oracle/quant_oracle.c) -- test infrastructure."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libquant_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_PATH)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def uniform_fwd(x, s, bucket, want_arg=True):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    n = x.size
    b = 0 if bucket is None else int(bucket)
    rows = 1 if (b == 0 or n < b) else -(-n // b)
    q = np.empty(n, np.float32)
    idx = np.empty(n, np.int64)
    alpha, beta = np.empty(rows, np.float32), np.empty(rows, np.float32)
    amin, amax = np.empty(rows, np.int64), np.empty(rows, np.int64)
    rc = lib().qo_uniform_fwd(_p(x), _p(q), _p(idx), _p(alpha), _p(beta), _p(amin), _p(amax), C.c_int64(n), C.c_int64(b), C.c_int(s))
    assert rc == 0
    return q, idx, dict(alpha=alpha, beta=beta, argmin=amin, argmax=amax)


def uniform_bwd_minmax(x, g, s, bucket):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
    out = np.empty_like(g)
    rc = lib().qo_uniform_bwd_minmax(_p(x), _p(g), _p(out), C.c_int64(x.size), C.c_int64(bucket), C.c_int(s))
    assert rc == 0
    return out


def uniform_bwd_minmax_ex(x, g, s, bucket):
    """(gout, per-row sum |v_j|, per-row r_b): the last two scale the summation-order tolerance."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
    n = x.size
    rows = 1 if n < bucket else -(-n // bucket)
    out = np.empty_like(g)
    ab, r = np.empty(rows, np.float64), np.empty(rows, np.float64)
    rc = lib().qo_uniform_bwd_minmax_ex(_p(x), _p(g), _p(out), _p(ab), _p(r), C.c_int64(n), C.c_int64(bucket), C.c_int(s))
    assert rc == 0
    return out, ab, r


def nonuniform_fwd(x, points, bucket, rule="nearest"):
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = x.size
    b = 0 if bucket is None else int(bucket)
    rows = 1 if (b == 0 or n < b) else -(-n // b)
    q, idx, alpha = np.empty(n, np.float32), np.empty(n, np.int64), np.empty(rows, np.float32)
    rc = lib().qo_nonuniform_fwd(_p(x), _p(pts), C.c_int(pts.size), C.c_int(1 if rule == "midpoint" else 0), _p(q), _p(idx),
                                 _p(alpha), C.c_int64(n), C.c_int64(b))
    assert rc == 0
    return q, idx, dict(alpha=alpha)


def nonuniform_bwd_points(g, idx, alpha, K, bucket):
    g = np.ascontiguousarray(g, dtype=np.float32).reshape(-1)
    idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(-1)
    alpha = np.ascontiguousarray(alpha, dtype=np.float32).reshape(-1)
    out = np.empty(K, np.float64)
    rc = lib().qo_nonuniform_bwd_points(_p(g), _p(idx), _p(alpha), C.c_int(K), _p(out), C.c_int64(g.size),
                                        C.c_int64(0 if bucket is None else int(bucket)))
    assert rc == 0
    return out
