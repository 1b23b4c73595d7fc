"""Exhaustive / exact-arithmetic checks of the cheap sequences that replace the two IEEE
divisions of the uniform op on the GPU (quantized_distillation_b200/csrc/qd_rowops.cuh,
"fast, still exact, level").  Pure host arithmetic with rationals: no GPU needed."""
from fractions import Fraction as Fr

import numpy as np

f32 = np.float32


def rn(fr: Fr) -> np.float32:
    """Round a rational to the nearest float32, ties to even."""
    c = f32(float(fr))
    cands = [c, np.nextafter(c, f32(np.inf)), np.nextafter(c, f32(-np.inf))]
    return f32(min(cands, key=lambda v: (abs(Fr(float(v)) - fr), int(np.array(v).view(np.uint32)) & 1)))


def test_small_level_to_unit_is_correctly_rounded_for_every_pair():
    """y0 = RN(k*rS); e = fma(-S, y0, k); y = fma(e, rS, y0)  ==  RN(k/S)  for all 0 <= k <= S <= 255."""
    checked = 0
    for S in range(1, 256):
        rS = rn(Fr(1, S))
        assert rS == f32(1.0) / f32(S)                     # what the host puts in Params.rS
        for k in range(0, S + 1):
            y0 = rn(Fr(k) * Fr(float(rS)))
            e = rn(Fr(k) - Fr(S) * Fr(float(y0)))          # single rounding of the exact residual = FMA
            assert Fr(float(e)) == Fr(k) - Fr(S) * Fr(float(y0)), "residual must be exact"
            y = rn(Fr(float(y0)) + Fr(float(e)) * Fr(float(rS)))
            assert y == rn(Fr(k, S)), (S, k)
            assert y == f32(k) / f32(S)
            checked += 1
    assert checked == 32895


def test_fast_level_band_is_conservative():
    """Emulates the fast level (t = RN(a*c), c within 2 ulp of S/alpha) against the reference
    chain rint(RN(RN(a/alpha)*S)) on adversarial inputs placed around every half-integer:
    whenever the guard accepts the candidate, it equals the reference level."""
    rng = np.random.default_rng(0)
    for S in (1, 3, 15, 255):
        lim = f32(0.5) - f32(S) * f32(2.0 ** -20)
        for alpha in (f32(1.0), f32(0.0371), f32(3.3e-5), f32(7777.7), f32(2.0 ** -90), f32(2.0 ** 90)):
            # x_hat targets: exactly on, and a few ulps around, every rounding boundary
            ks = np.arange(0, S, dtype=np.float64) + 0.5
            base = (ks / S)[:, None] * (1 + np.arange(-40, 41)[None, :] * 2.0 ** -24)
            a = (base.reshape(-1) * float(alpha)).astype(f32)
            a = np.concatenate([a, (rng.random(20000) * float(alpha)).astype(f32)])
            a = a[(a >= 0) & (a <= alpha)]
            ref = np.rint(((a / alpha).astype(f32) * f32(S)).astype(f32))
            for ulps in (-2, -1, 0, 1, 2):                 # any c the approximate reciprocal may return
                c = f32(S) / alpha
                for _ in range(abs(ulps)):
                    c = np.nextafter(c, f32(np.inf) if ulps > 0 else f32(0))
                t = (a * c).astype(f32)
                k = np.rint(t)
                accepted = np.abs((t - k).astype(f32)) < lim
                assert np.array_equal(k[accepted], ref[accepted]), (S, alpha, ulps)
                assert accepted.mean() > 0.5                # the guard is not vacuous


# ---------------------------------------------------------------------------------------------
# centroid op: one threshold table for both index rules (qd_rowops.cuh, "centroids")
def _nearest_reference(k, v):
    """nonUniformQuantization's direct path on one value (quant_functions.py:267-273), float32 arithmetic."""
    K = len(k)
    i = int(np.searchsorted(k, v, side="left"))
    i = min(i, K - 1)
    if i > 0 and abs(f32(v) - k[i - 1]) < abs(f32(v) - k[i]):
        i -= 1
    return i


def _key(v):
    b = int(np.array(v, dtype=f32).view(np.uint32))
    return (~b) & 0xFFFFFFFF if b & 0x80000000 else b | 0x80000000


def _unkey(key):
    b = (key & 0x7FFFFFFF) if (key & 0x80000000) else (~key) & 0xFFFFFFFF
    return np.array(b, dtype=np.uint32).view(f32)[()]


def _nearest_threshold(k, j):
    """Host transcription of qd::nearest_threshold: smallest float whose nearest-rule index exceeds j;
    bisection over the ordered bit patterns, a +-16 ulp window around the float32 midpoint first."""
    lo, hi = _key(f32(-np.inf)), _key(f32(np.inf))
    c = f32(k[j] + f32(f32(k[j + 1] - k[j]) * f32(0.5)))
    ck = _key(c)
    if lo + 16 < ck < hi - 16:
        wl, wh = ck - 16, ck + 16
        if _nearest_reference(k, _unkey(wl)) <= j and _nearest_reference(k, _unkey(wh)) > j:
            lo, hi = wl + 1, wh
    while lo < hi:
        mid = lo + ((hi - lo) >> 1)
        if _nearest_reference(k, _unkey(mid)) > j:
            hi = mid
        else:
            lo = mid + 1
    return _unkey(lo)


def test_nearest_rule_thresholds():
    """The nearest-point rule is a monotone step function of x_hat, so K-1 thresholds reproduce it exactly:
    idx = #{ j : t_j <= x_hat }.  Random, evenly spaced, duplicated and nearly coincident point lists;
    probes on random values, every threshold +-4 ulp, every point +-3 ulp, zero / negative / > 1 inputs."""
    rng = np.random.default_rng(0)
    for trial in range(80):
        K = int(rng.choice([2, 3, 4, 5, 8, 16, 17]))
        kind = trial % 4
        if kind == 0:
            k = np.sort(rng.random(K).astype(f32))
        elif kind == 1:
            k = np.linspace(0, 1, K).astype(f32)
        elif kind == 2:
            k = np.sort(rng.random(K).astype(f32))
            k[K // 2] = k[K // 2 - 1]
            if K > 3:
                k[-1] = k[-2]
        else:
            k = np.sort((rng.random(K) * 1e-6 + 0.5).astype(f32))
        t = np.array([_nearest_threshold(k, j) for j in range(K - 1)], dtype=f32)
        assert np.all(np.diff(t) >= 0)
        probes = [rng.random(400).astype(f32), np.array([0, 1, -0.0, 1e-30, 2.0, -1.0], dtype=f32)]
        for v in list(t) + list(k):
            b = int(np.array([v], dtype=f32).view(np.uint32)[0])
            probes.append(np.array([max(b + d, 0) for d in range(-4, 5)], dtype=np.uint32).view(f32))
        xs = np.concatenate(probes)
        xs = xs[~np.isnan(xs)]
        for x in xs:
            assert _nearest_reference(k, x) == int((t <= x).sum()), (k, t, x)
