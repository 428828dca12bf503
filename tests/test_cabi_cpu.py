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
