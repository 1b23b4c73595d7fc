"""Three independent restatements of the reference arithmetic (NumPy closed forms, the torch
op chain, the C port) must agree bit for bit on arbitrary inputs, and obey the properties the
quantizer has by construction.  Hypothesis explores sizes / buckets / levels / value ranges
beyond the golden vectors (CPU only)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as CO
from oracle import quant_oracle as O
from oracle import torch_chain as T

sizes = st.integers(min_value=1, max_value=3000)
buckets = st.one_of(st.none(), st.integers(min_value=1, max_value=700))
levels = st.sampled_from([2, 3, 4, 8, 16, 17, 255, 256, 1000])
scales = st.sampled_from([1e-30, 1e-6, 0.05, 1.0, 1e4, 1e30])


def make(n, scale, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == 0:
        x = rng.standard_normal(n)
    elif kind == 1:
        x = rng.integers(-3, 4, n).astype(np.float64)            # many exact ties / duplicates
    else:
        x = rng.standard_normal(n) * np.exp(rng.uniform(-20, 5, n))
    return (x * scale).astype(np.float32)


@settings(max_examples=150, deadline=None, derandomize=True)
@given(n=sizes, bucket=buckets, s=levels, scale=scales, seed=st.integers(0, 2 ** 20), kind=st.integers(0, 2))
def test_uniform_three_way_agreement_and_properties(n, bucket, s, scale, seed, kind):
    x = make(n, scale, seed, kind)
    with np.errstate(all="ignore"):
        q, idx, stt = O.uniform_fwd(x, s, bucket)
    qc, idxc, stc = CO.uniform_fwd(x, s, bucket)
    qt, stch = T.uniform_fwd(torch.from_numpy(x.copy()), s, bucket)
    assert np.array_equal(q.view(np.uint32), qc.view(np.uint32))
    assert np.array_equal(q.view(np.uint32), qt.numpy().view(np.uint32))
    assert np.array_equal(idx, idxc) and np.array_equal(stt["argmin"], stc["argmin"]) and np.array_equal(stt["argmax"], stc["argmax"])
    assert np.array_equal(stt["alpha"], stch.alpha.reshape(-1).numpy())
    # properties: levels in range, at most s distinct values per bucket, bucket min reproduced exactly
    assert idx.min() >= 0 and idx.max() <= s - 1
    rows, row_len, _ = O.bucket_geometry(n, bucket)
    for r in range(min(rows, 4)):
        sl = slice(r * row_len, min((r + 1) * row_len, n))
        assert len(np.unique(q[sl])) <= s
        assert q[sl].min() == x[sl].min() or not np.isfinite(x[sl]).all()
        assert idx[sl][stt["argmin"][r]] == 0


@settings(max_examples=100, deadline=None, derandomize=True)
@given(n=sizes, bucket=buckets, K=st.integers(1, 40), seed=st.integers(0, 2 ** 20), kind=st.integers(0, 2),
       rule=st.sampled_from(["nearest", "midpoint"]))
def test_nonuniform_three_way_agreement(n, bucket, K, seed, kind, rule):
    x = make(n, 0.05, seed, kind)
    rng = np.random.default_rng(seed + 1)
    pts = np.sort(rng.random(K)).astype(np.float32)
    q, idx, stt = O.nonuniform_fwd(x, pts, bucket, rule=rule)
    qc, idxc, stc = CO.nonuniform_fwd(x, pts, bucket, rule=rule)
    qt, idxt, _ = T.nonuniform_fwd(torch.from_numpy(x.copy()), torch.from_numpy(pts.copy()), bucket, rule=rule)
    assert np.array_equal(idx, idxc) and np.array_equal(idx, idxt.numpy())
    assert np.array_equal(q.view(np.uint32), qc.view(np.uint32)) and np.array_equal(q.view(np.uint32), qt.numpy().view(np.uint32))
    assert idx.min() >= 0 and idx.max() <= K - 1
    # centroid gradient: sum over centroids equals the alpha-weighted gradient sum (partition of unity)
    g = rng.standard_normal(n).astype(np.float32)
    gp = O.nonuniform_bwd_points(g, idx, stt["alpha"], K, bucket)
    rows, row_len, _ = O.bucket_geometry(n, bucket)
    a = np.repeat(stt["alpha"], row_len)[:n]
    total = (g * a).astype(np.float32).astype(np.float64).sum()
    assert abs(gp.sum() - total) <= 1e-9 * max(1.0, np.abs(g * a).sum())
    assert np.allclose(gp, CO.nonuniform_bwd_points(g, idx, stt["alpha"], K, bucket), rtol=0, atol=1e-12 * max(1.0, np.abs(g * a).sum()))


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n=st.integers(1, 2000), bucket=st.integers(1, 600), s=st.sampled_from([2, 4, 16, 256]), seed=st.integers(0, 2 ** 20))
def test_minmax_backward_agreement_and_conservation(n, bucket, s, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) * 0.05).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    out, info = O.uniform_bwd_minmax(x, g, s, bucket)
    outc = CO.uniform_bwd_minmax(x, g, s, bucket)
    outt = T.uniform_bwd_minmax(torch.from_numpy(x.copy()), torch.from_numpy(g.copy()), s, bucket).numpy()
    tol = 1e-5 * max(1.0, np.abs(g).sum())
    assert np.abs(out.astype(np.float64) - outc).max() <= tol
    assert np.abs(out.astype(np.float64) - outt).max() <= tol
    # +r at argmax' and -r at argmin' cancel: every bucket's gradient sum is preserved
    rows, row_len, _ = O.bucket_geometry(n, bucket)
    for r in range(min(rows, 6)):
        sl = slice(r * row_len, min((r + 1) * row_len, n))
        assert abs(out[sl].astype(np.float64).sum() - g[sl].astype(np.float64).sum()) <= tol
