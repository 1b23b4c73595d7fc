"""Randomised GPU parity: hypothesis picks sizes, bucket sizes (all three execution paths, aligned and
unaligned), level counts and value ranges; the CUDA result must equal the C restatement of the
reference bit for bit."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import quantized_distillation_b200.quantization as Q
    return Q


def make(n, scale, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == 0:
        x = rng.standard_normal(n)
    elif kind == 1:
        x = rng.integers(-3, 4, n).astype(np.float64)
    else:
        x = rng.standard_normal(n) * np.exp(rng.uniform(-20, 5, n))
    return (x * scale).astype(np.float32)


buckets = st.one_of(st.none(), st.integers(1, 1024), st.integers(1025, 60000))


@settings(max_examples=120, deadline=None, derandomize=True)
@given(n=st.integers(1, 120000), bucket=buckets, s=st.sampled_from([2, 3, 4, 16, 255, 256, 1000]),
       scale=st.sampled_from([1e-30, 1e-6, 0.05, 1.0, 1e4, 1e30]), seed=st.integers(0, 2 ** 20), kind=st.integers(0, 2),
       offset=st.integers(0, 3))
def test_uniform_forward_random(Q, n, bucket, s, scale, seed, kind, offset):
    x = make(n + offset, scale, seed, kind)
    xd = torch.from_numpy(x).cuda()[offset:]                      # offset > 0: only 4-byte aligned
    q, sf = Q.uniformQuantization(xd, s, bucket_size=bucket)
    qc, idxc, stc = CO.uniform_fwd(x[offset:], s, bucket)
    got = q.cpu().numpy()
    nan = np.isnan(got) & np.isnan(qc)
    assert np.array_equal(np.where(nan, 0, got).view(np.uint32), np.where(nan, 0, qc).view(np.uint32)), (n, bucket, s, scale, kind, offset)
    assert np.array_equal(sf.alpha.view(-1).cpu().numpy().view(np.uint32), stc["alpha"].view(np.uint32))
    assert np.array_equal(sf.idx_min_rows.view(-1).cpu().numpy(), stc["argmin"])
    assert np.array_equal(sf.idx_max_rows.view(-1).cpu().numpy(), stc["argmax"])


@settings(max_examples=250, deadline=None, derandomize=True)
@given(n=st.integers(1, 120000), bucket=buckets, K=st.integers(1, 64), seed=st.integers(0, 2 ** 20), kind=st.integers(0, 2),
       rule=st.sampled_from(["nearest", "midpoint"]))
def test_nonuniform_forward_and_points_gradient_random(Q, n, bucket, K, seed, kind, rule):
    x = make(n, 0.05, seed, kind)
    rng = np.random.default_rng(seed + 7)
    pts = np.sort(rng.random(K)).astype(np.float32)
    xd, pd = torch.from_numpy(x).cuda(), torch.from_numpy(pts).cuda()
    qc, idxc, stc = CO.nonuniform_fwd(x, pts, bucket, rule=rule)
    if rule == "nearest":
        q, idx, sf = Q.nonUniformQuantization(xd, pd, bucket_size=bucket)
        alpha = sf.alpha
    else:
        f = Q.nonUniformQuantization_variable(bucket_size=bucket, pre_process_tensors=True, tensor=xd)
        q = f.forward(None, pd)
        idx = f.savedForBackward["indices"].to(torch.int64)
        alpha = f.savedForBackward["scalingFactor"]
        g = rng.standard_normal(n).astype(np.float32)
        _, gp = f.backward(torch.from_numpy(g).cuda())
        ref = CO.nonuniform_bwd_points(g, idxc, stc["alpha"], K, bucket)
        # tolerance relative to the sum of magnitudes (the per-centroid sums cancel heavily for random g)
        rows_, row_len_, _ = __import__("oracle.quant_oracle", fromlist=["x"]).bucket_geometry(n, bucket)
        mag = float(np.abs(g.astype(np.float64) * np.repeat(stc["alpha"].astype(np.float64), row_len_)[:n]).sum())
        err = np.abs(gp.cpu().numpy().astype(np.float64) - ref).max()
        assert err <= 1e-6 * mag + 1e-30, ("centroid gradient", n, bucket, K, err, mag)
    assert np.array_equal(idx.view(-1).cpu().numpy(), idxc), ("idx", n, bucket, K, rule, seed, kind)
    assert np.array_equal(q.view(-1).cpu().numpy().view(np.uint32), qc.view(np.uint32)), ("q", n, bucket, K, rule, seed, kind)
    assert np.array_equal(alpha.view(-1).cpu().numpy().view(np.uint32), stc["alpha"].view(np.uint32)), ("alpha", n, bucket, K)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(n=st.integers(1, 60000), bucket=st.one_of(st.integers(1, 1024), st.integers(1025, 20000)), s=st.sampled_from([2, 4, 16, 256]),
       seed=st.integers(0, 2 ** 20))
def test_minmax_backward_random(Q, n, bucket, s, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) * 0.05).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    f = Q.uniformQuantization_variable(s, bucket_size=bucket)
    f.forward(torch.from_numpy(x).cuda())
    out = f.backward(torch.from_numpy(g).cuda()).cpu().numpy()
    # a5 bar (see tests/test_gpu_parity.py::assert_minmax_gradient): every element within the summation-order tolerance of
    # its bucket -- 1e-6 * sum_j |v_j| + one ulp of r_b + one ulp of the result -- which is zero for untouched elements
    ref, abs_sum, r = CO.uniform_bwd_minmax_ex(x, g, s, bucket)
    row_len = bucket if n >= bucket else n
    row = np.arange(n) // row_len
    ulp = 2.0 ** -23
    tol = 1e-6 * abs_sum[row] + ulp * (np.abs(r[row]) + np.maximum(np.abs(ref), np.abs(g))) + 1e-37
    touched = (out != g) | (ref != g)
    tol = np.where(touched, tol, 0.0)
    err = np.abs(out.astype(np.float64) - ref.astype(np.float64))
    assert np.all(err <= tol), (n, bucket, s, float(err.max()))
    assert touched.sum() <= 2 * (n // row_len + 1)
