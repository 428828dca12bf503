"""The blocked bitonic sort of the general build's hub buckets (csrc/gtsb_sort_core.h, used by
k_resolve_large2) run on the host (tests/emul/sort_emul.cpp) in both thread orders: sizes below,
at and above the shared-memory chunk, both keys, tags moving with their entries, the padding
entry sorting last."""
import ctypes as C

import numpy as np
import pytest

import parse_emul

OTHER_MASK = (1 << 27) - 1


def _sort(ent, tag, mode, reverse, plain=False):
    L = parse_emul.lib()
    ent = np.ascontiguousarray(ent, np.uint32).copy()
    tg = None if tag is None else np.ascontiguousarray(tag, np.uint32).copy()
    (L.emul_plain_bitonic if plain else L.emul_blocked_bitonic)(ent.ctypes.data_as(C.c_void_p), None if tg is None else tg.ctypes.data_as(C.c_void_p),
                           C.c_uint32(ent.shape[0]), C.c_int(mode), C.c_int(reverse))
    return ent, tg


def _expected(ent, mode):
    if mode == 0:
        key = ((ent[:, 1] & OTHER_MASK).astype(np.uint64) << np.uint64(32)) | ent[:, 0].astype(np.uint64)
    else:
        key = ent[:, 0].astype(np.uint64)
    return np.argsort(key, kind="stable")


@pytest.mark.parametrize("logp", [0, 1, 2, 5, 6, 10, 11, 12, 13, 15])
@pytest.mark.parametrize("mode", [0, 1])
def test_blocked_bitonic_equals_a_key_sort(logp, mode):
    P = 1 << logp
    rng = np.random.default_rng(1000 * logp + mode)
    for reverse in (0, 1):
        n = P if logp < 3 else int(rng.integers(P // 2 + 1, P + 1))      # the bucket, padded to P
        ent = np.zeros((P, 4), np.uint32)
        ent[:n, 0] = rng.permutation(np.arange(3 * n, dtype=np.uint32))[:n]                 # unique record indices
        ent[:n, 1] = rng.integers(0, max(2, n // 3), n).astype(np.uint32) | (rng.integers(0, 32, n).astype(np.uint32) << 27)
        ent[:n, 2:] = rng.integers(0, 2 ** 32, (n, 2), dtype=np.uint64).astype(np.uint32)
        ent[n:] = (0xFFFFFFFF, 0xFFFFFFFF, 0, 0)                          # k_resolve_large's padding
        tag = rng.integers(0, 2 ** 32, P, dtype=np.uint64).astype(np.uint32)
        for tg, plain in ((None, False), (tag, False), (tag, True)):
            if plain and P > 128:                                         # the warp path holds up to 128 entries
                continue
            got, gtag = _sort(ent, tg, mode, reverse, plain)
            order = _expected(ent, mode)
            # keys are unique among the real entries; the padding entries are identical
            assert np.array_equal(got[:n], ent[order][:n]), (logp, mode, reverse)
            assert np.all(got[n:, 0] == 0xFFFFFFFF)
            if tg is not None:
                assert np.array_equal(gtag[:n], tag[order][:n])


def test_chunk_size_is_what_the_kernel_reserves():
    assert parse_emul.lib().emul_sort_chunk() == 2048
