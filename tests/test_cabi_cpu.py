"""CPU-side checks of the product boundary: the C-ABI library loads, exports
every symbol include/gtscaffold_b200.h declares, fails loudly without a GPU,
and its host helper (ambiguous-order thresholds) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gtscaffold_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtsb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(pkg.api.EXPORTS) == names


def test_every_entry_point_has_ctypes_prototypes(pkg):
    """a pointer passed without argtypes is cut to a C int"""
    L = pkg.load_library()
    for n in pkg.api.EXPORTS:
        assert getattr(L, n).argtypes is not None, n


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must refuse, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pkg.ScaffoldGraphB200(device=0)


@pytest.mark.parametrize("cutoff", [0.01, 0.2, 0.3, 0.49, 0.5, 0.0, -0.5, 1e-9, 0.05])
def test_thresholds_reproduce_reference_decision(pkg, cutoff):
    """t_pos/t_neg must reproduce gt_scaffolder_graph_ambiguousorder (oracle
    restatement, algorithms.c:174-193) for intervals around the step."""
    rc, t_pos, t_neg, inf_true = pkg.api.ambig_thresholds(cutoff)
    assert rc == 0
    P = O.port_lib()
    rng = np.random.default_rng(1)
    # pairs with std chosen so that interval = -delta / sqrt(4 s^2) = -delta / (2 s)
    for _ in range(4000):
        s = np.float32(rng.uniform(0.5, 60))
        d1 = int(rng.integers(-400, 400))
        d2 = int(rng.integers(-400, 400))
        exp = P.ora_ambiguousorder(d1, s, d2, s, np.float32(cutoff))
        expval = np.float32(d1 - d2)
        var = np.float32(2) * (s * s + s * s)
        with np.errstate(all="ignore"):
            interval = np.float32(np.float64(np.float32(0) - expval) / np.sqrt(np.float64(var)))
        if interval >= 0:
            got = t_pos >= 0 and interval <= t_pos
        else:
            got = t_neg >= 0 and -interval <= t_neg
        assert bool(exp) == bool(got), (d1, d2, s, interval, t_pos, t_neg)


@pytest.mark.parametrize("cutoff", [0.01, 0.2, 0.3, 0.49, 0.4999999, 0.5, 0.0, -0.5, 1e-9, 1e-30, 0.05, 0.25,
                                    float(np.float32(0.01) + np.float32(1e-9)), 0.999, 7.0])
def test_thresholds_are_the_exact_step_of_the_reference_decision(pkg, cutoff):
    """The device decides `|interval| <= t` instead of evaluating erf (gtsb_common.cuh,
    ambiguous_order); t must be the LAST float for which algorithms.c:187-192 says
    "ambiguous" -- checked on every float within 5000 ulps of both thresholds, on two million
    random bit patterns (denormals, huge values, both zeros) and on inf / NaN."""
    rc, t_pos, t_neg, inf_true = pkg.api.ambig_thresholds(cutoff)
    assert rc == 0
    P = O.port_lib()
    rng = np.random.default_rng(11)
    parts = [rng.integers(0, 2**32, 2_000_000, dtype=np.uint64).astype(np.uint32).view(np.float32),
             np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, -3.4e38], np.float32)]
    for t, sign in ((t_pos, 1.0), (t_neg, -1.0)):
        if t >= 0 and np.isfinite(t):
            b = int(np.float32(t).view(np.uint32))
            near = np.arange(max(0, b - 5000), min(0x7F7FFFFF, b + 5000) + 1, dtype=np.int64)
            parts.append((near.astype(np.uint32).view(np.float32) * np.float32(sign)).astype(np.float32))
    x = np.ascontiguousarray(np.concatenate(parts), np.float32)
    exp = np.zeros(len(x), np.uint8)
    P.ora_ambiguous_intervals(x.ctypes.data, len(x), np.float32(cutoff), exp.ctypes.data)
    neg = np.signbit(x)
    with np.errstate(invalid="ignore"):
        got = np.where(neg, (t_neg >= 0) & (-x <= np.float32(t_neg)), (t_pos >= 0) & (x <= np.float32(t_pos)))
    got = np.where(np.isinf(x), bool(inf_true), got)
    got = np.where(np.isnan(x), False, got)
    bad = np.nonzero(got.astype(np.uint8) != exp)[0]
    assert len(bad) == 0, (cutoff, t_pos, t_neg, inf_true, x[bad[:5]], exp[bad[:5]])
    if 0 < cutoff < 0.5:
        assert exp.sum() > 1000 and (exp == 0).sum() > 1000


def device_ambiguous_order(d1, s1, d2, s2, t_pos, t_neg, inf_true):
    """gtsb_common.cuh ambiguous_order, operation by operation in numpy (every CUDA intrinsic
    there is a correctly rounded IEEE operation, so float32 / float64 numpy gives its bits)."""
    f32, f64 = np.float32, np.float64
    with np.errstate(all="ignore"):
        expval = (d1 - d2).astype(f32)
        variance = f32(2) * (s1 * s1 + s2 * s2)
        x = f32(0) - expval
        neg = x < 0
        t = np.where(neg, f32(t_neg), f32(t_pos)).astype(f32)
        c = (t * t).astype(f32)
        fast = (t < f32(3.0e38)) & (variance > f32(1.0e-30)) & (variance < f32(1.0e30)) & (np.abs(x) < f32(1.0e15))
        lhs = x * x
        rhs = c * variance
        band = rhs * f32(1.0e-5)
        interval = (x.astype(f64) / np.sqrt(variance.astype(f64))).astype(f32)
        exact = np.where(neg, -interval <= f32(t_neg), interval <= f32(t_pos))
        exact = np.where(np.isinf(interval), bool(inf_true), exact)
        exact = np.where(np.isnan(interval), False, exact)
        out = np.where(fast & (lhs < rhs - band), True, np.where(fast & (lhs > rhs + band), False, exact))
        return np.where(t >= 0, out, False)


@pytest.mark.parametrize("cutoff", [0.01, 0.2, 0.3, 0.49, 0.5, 0.0, 1e-9, 0.05])
def test_device_form_of_ambiguous_order_equals_the_reference(pkg, cutoff):
    """the whole device decision -- squared fast path with its guard band, exact division
    inside the band -- against algorithms.c:174-193 on pairs aimed at the threshold"""
    rc, t_pos, t_neg, inf_true = pkg.api.ambig_thresholds(cutoff)
    P = O.port_lib()
    rng = np.random.default_rng(5)
    n = 400_000
    d1 = rng.integers(-5000, 5000, n).astype(np.int64)
    s1 = rng.uniform(0.1, 80, n).astype(np.float32)
    s2 = np.where(rng.random(n) < 0.3, np.float32(0), rng.uniform(0.1, 80, n)).astype(np.float32)
    # d2 such that interval = -(d1 - d2) / sqrt(2 (s1^2 + s2^2)) lands within 1e-7 .. 1e-2 of +-t
    t = np.where(rng.random(n) < 0.5, t_pos, -t_neg)
    t = np.where(np.isfinite(t) & (np.abs(t) < 1e3), t, rng.normal(0, 2, n))
    eps = rng.choice([0, 1e-7, -1e-7, 1e-6, -1e-6, 2e-5, -2e-5, 1e-3, -1e-3, 1e-2, -1e-2], n)
    want = t * (1 + eps)
    d2 = d1 + np.rint(want * np.sqrt(2 * (s1.astype(np.float64) ** 2 + s2.astype(np.float64) ** 2))).astype(np.int64)
    # plus special values
    s1[:50] = np.array([0, np.inf, np.nan, 1e-30, 1e30, 1e-20, 3e38, -1.0, 1e19, 1e-19] * 5, np.float32)
    s2[:25] = 0
    d2[:10] = d1[:10]
    d1[10:20] = np.array([2**31 - 1, -2**31, 2**40, -2**40, 0, 1, -1, 2**24 + 1, 10**15, -10**15])
    exp = np.zeros(n, np.uint8)
    P.ora_ambiguousorders(d1.ctypes.data, s1.ctypes.data, d2.ctypes.data, s2.ctypes.data, n,
                          np.float32(cutoff), exp.ctypes.data)
    got = device_ambiguous_order(d1, s1, d2, s2, t_pos, t_neg, inf_true)
    bad = np.nonzero(got.astype(np.uint8) != exp)[0]
    assert len(bad) == 0, (cutoff, len(bad), d1[bad[:4]], d2[bad[:4]], s1[bad[:4]], s2[bad[:4]], exp[bad[:4]])
    if 0 < cutoff < 0.5:
        assert 1000 < exp.sum() < n - 1000


def test_low_copy_number_pruning_never_drops_a_proposing_pair():
    """k4_pairs_big (gtsb_filter.cu) evaluates only pairs with a member in the `low` list,
    !(c > cut/2 + |cut/2| * 1e-6 + 1e-30) in float32.  Whenever the reference's test
    fl(c1 + c2) < cut (algorithms.c:233) holds, the smaller of the two must be in that list."""
    f32 = np.float32
    rng = np.random.default_rng(3)
    n = 4_000_000
    cuts = np.concatenate([rng.uniform(-3, 6, n // 2), 10.0 ** rng.uniform(-44, 38, n // 4),
                           -(10.0 ** rng.uniform(-44, 38, n // 4))]).astype(f32)
    cuts[:8] = np.array([1.5, 2.5, 0.0, -0.0, np.inf, -np.inf, 3.4e38, 1e-45], f32)
    half = cuts * f32(0.5)
    with np.errstate(all="ignore"):
        bound = half + np.abs(half) * f32(1e-6) + f32(1e-30)
        # pairs aimed at the boundary: c1 just above / below cut/2, c2 >= c1 so that the sum sits at the cut
        ulps = rng.integers(-40, 41, n)
        c1 = (half.view(np.int32) + np.where(half >= 0, ulps, -ulps)).astype(np.int32).view(f32)
        c1 = np.where(rng.random(n) < 0.2, (half.astype(np.float64) * (1 + rng.normal(0, 3e-6, n))).astype(f32), c1)
        c2 = (cuts.astype(np.float64) - c1.astype(np.float64)).astype(f32)
        c2 = (c2.view(np.int32) + rng.integers(-3, 4, n)).astype(np.int32).view(f32)
        c2 = np.where(rng.random(n) < 0.1, c1, c2)
        special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, -1e-45, 3.4e38], f32)
        k = rng.random(n) < 0.02
        c2 = np.where(k, special[rng.integers(0, 8, n)], c2)
        lo = np.where(c1 < c2, c1, c2)                    # either order: the list holds slots, not roles
        lo = np.where(np.isnan(c1) | np.isnan(c2), np.nan, lo).astype(f32)
        proposes = (c1 + c2) < cuts
        in_low_1, in_low_2 = ~(c1 > bound), ~(c2 > bound)
    dropped = proposes & ~(in_low_1 | in_low_2)
    assert proposes.sum() > n // 10 and (~proposes).sum() > n // 10
    assert dropped.sum() == 0, (cuts[dropped][:4], c1[dropped][:4], c2[dropped][:4])
