"""`.scaf` text from scaffold records (gt_scaffolder_graph_write_scaffold,
gt_scaffolder_algorithms.c:1000-1042).

CPU: the "%f" routine of gtsb_format_core.h (integer arithmetic on the float's bits) against the
     C library's printf on the floats a `.de` file can hold, on random bit patterns, around every
     rounding boundary of the sixth decimal, on denormals / inf / nan; the record writer compiled for
     the host against the `.scaf` file the COMPILED REFERENCE writes (config 1, re-emitted from the
     parsed golden) and against a direct printf statement of the loop.
GPU: gtsb_scaf_lines_host against the same host build; the drop-in binary's `.scaf`
     (tests/test_dropin.py) goes through this path.
"""
import ctypes as C
import os

import numpy as np
import pytest

import parse_emul as PE

HERE = os.path.dirname(os.path.abspath(__file__))
C1 = os.path.join(HERE, "golden", "c1")
_libc = C.CDLL(None)
_libc.snprintf.restype = C.c_int


def printf_f(x32: np.float32) -> bytes:
    """what gt_file_xprintf("%f", e->std_dev) prints: the float promoted to double"""
    buf = C.create_string_buffer(400)
    n = _libc.snprintf(buf, C.c_size_t(400), b"%f", C.c_double(float(x32)))
    return buf.raw[:n]


def _check_bits(bits):
    for b in np.asarray(bits, np.uint32):
        x = np.array([b], np.uint32).view(np.float32)[0]
        assert PE.f6(int(b)) == printf_f(x), hex(int(b))


def test_f6_equals_printf_on_de_values():
    # `.de` standard deviations carry one decimal (parser.c:118); all of 0.0 .. 999.9 and some more
    vals = np.round(np.arange(0, 10000) / 10.0, 1).astype(np.float32)
    _check_bits(vals.view(np.uint32))
    _check_bits(np.array([0.05, 1e-7, 123456.789, 1e10, 3.4e38, 16777216.0, 0.5, 0.0000005, 0.0000015, 0.0000025],
                         np.float32).view(np.uint32))
    _check_bits((-vals[:200]).view(np.uint32))


def test_f6_equals_printf_on_random_bits_and_specials():
    rng = np.random.default_rng(5)
    _check_bits(rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32))
    # denormals, the smallest normals, +-0, +-inf, nans, the largest floats
    special = [0, 1, 2, 0x007FFFFF, 0x00800000, 0x00800001, 0x80000000, 0x80000001, 0x7F800000, 0xFF800000,
               0x7FC00000, 0xFFC00000, 0x7F800001, 0x7F7FFFFF, 0xFF7FFFFF, 0x3F800000, 0xBF800000]
    for b in special:
        x = np.array([b], np.uint32).view(np.float32)[0]
        got, exp = PE.f6(b), printf_f(x)
        if np.isnan(x):
            assert got.lstrip(b"-") == b"nan" and exp.lstrip(b"-") == b"nan"      # glibc prints the sign bit too
            assert got == exp
        else:
            assert got == exp, hex(b)


def test_f6_around_rounding_boundaries():
    # floats next to k + 0.5e-6 multiples (ties of the sixth decimal) in several binades
    rng = np.random.default_rng(6)
    bits = []
    for scale in (1.0, 0.001, 37.0, 1000.0, 0.25):
        k = rng.integers(0, 2_000_000, 300)
        t = ((k + 0.5) * 1e-6 * scale).astype(np.float32)
        b = t.view(np.uint32).astype(np.int64)
        for d in (-2, -1, 0, 1, 2):
            bits.extend((b + d).tolist())
    # exact ties exist only for dyadic values: x.5 at the sixth decimal needs 2^-7 .. multiples
    for e in range(1, 24):
        bits.append(int(np.array([2.0 ** -e], np.float32).view(np.uint32)[0]))
        bits.append(int(np.array([3 * 2.0 ** -e], np.float32).view(np.uint32)[0]))
        bits.append(int(np.array([1 + 2.0 ** -e], np.float32).view(np.uint32)[0]))
    _check_bits(np.array(bits, np.uint32))


def _direct(names, records):
    out = []
    for root, edges in records:
        out.append(names[root])
        for end, dist, std, sense, same in edges:
            out.append(b"\t" + names[end] + b"," + str(int(dist)).encode() + b"," + printf_f(np.float32(std))
                       + b"," + (b"1" if sense else b"0") + b"," + (b"1" if same else b"0") + b",")
        out.append(b"\n")
    return b"".join(out)


def random_records(rng, V, n, max_edges=12, wild=False):
    recs = []
    for _ in range(n):
        m = int(rng.integers(0, max_edges + 1))
        edges = []
        for _ in range(m):
            std = np.float32(np.round(rng.uniform(0, 80), 1))
            dist = int(rng.integers(-500, 30000))
            if wild:
                std = np.array([rng.integers(0, 2**32)], np.uint64).astype(np.uint32).view(np.float32)[0]
                dist = int(rng.integers(-2**63, 2**63 - 1, dtype=np.int64))
            edges.append((int(rng.integers(0, V)), dist, std, bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
        recs.append((int(rng.integers(0, V)), edges))
    return recs


@pytest.mark.parametrize("seed", range(4))
def test_host_build_equals_printf_loop(seed):
    rng = np.random.default_rng(100 + seed)
    V = 50 + 40 * seed
    names = [[b"contig-%d", b"c%d", b"%d_long_header_with_more_text"][v % 3] % v for v in range(V)]
    recs = random_records(rng, V, 30 + 50 * seed, wild=bool(seed % 2))
    recs += [(0, []), (V - 1, [])]                                   # records without edges
    assert PE.scaf_lines(names, recs) == _direct(names, recs)
    assert PE.scaf_lines(names, []) == b""


def parse_scaf(text, ids):
    recs = []
    for line in text.split(b"\n")[:-1]:
        parts = line.split(b"\t")
        edges = []
        for p in parts[1:]:
            f = p.split(b",")
            assert len(f) == 6 and f[5] == b""
            edges.append((ids[f[0]], int(f[1]), np.float32(float(f[2])), f[3] == b"1", f[4] == b"1"))
        recs.append((ids[parts[0]], edges))
    return recs


def test_config1_scaf_reemitted():
    """the `.scaf` file the compiled reference writes for config 1, parsed and written again"""
    golden = open(os.path.join(C1, "c1_expected.scaf"), "rb").read()
    heads = sorted({h for line in golden.split(b"\n")[:-1] for h in
                    [line.split(b"\t")[0]] + [p.split(b",")[0] for p in line.split(b"\t")[1:]]})
    ids = {h: i for i, h in enumerate(heads)}
    recs = parse_scaf(golden, ids)
    assert len(recs) > 0 and sum(len(e) for _, e in recs) > 0
    assert PE.scaf_lines(heads, recs) == golden


# ------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_device_equals_host_build(pkg, seed):
    rng = np.random.default_rng(200 + seed)
    V = 100 + 3000 * seed
    names = [[b"contig-%d", b"c%d", b"%d_long_header_with_more_text"][v % 3] % v for v in range(V)]
    recs = random_records(rng, V, 50 + 4000 * seed, wild=bool(seed % 2))
    recs += [(0, []), (V - 1, [])]
    g = pkg.ScaffoldGraphB200()
    g.set_vertex_names(names)
    a = PE.scaf_arrays(recs)
    cap = sum(len(names[r]) + 1 for r in a[0]) + sum(len(names[w]) + 80 for w in a[2])
    assert g.scaf_lines(*a, cap) == PE.scaf_lines(names, recs)
    # nothing; an id outside the names set; too little room
    assert g.scaf_lines([], [0], [], [], [], [], 0) == b""
    bad = list(a)
    bad[0] = a[0].copy()
    bad[0][0] = V
    with pytest.raises(RuntimeError, match="outside the names set"):
        g.scaf_lines(*bad, cap)
    with pytest.raises(RuntimeError, match="room for"):
        g.scaf_lines(*a, 10)


@pytest.mark.gpu
def test_device_scaf_at_size(pkg):
    import json
    rng = np.random.default_rng(9)
    V, n = 300_000, 40_000
    names = [b"contig-%d" % v for v in range(V)]
    order = rng.permutation(V)
    cuts = np.sort(rng.choice(np.arange(1, V), n - 1, replace=False))
    recs = []
    for part in np.split(order, cuts):
        edges = [(int(w), int(d), np.float32(s), bool(a), bool(b)) for w, d, s, a, b in
                 zip(part[1:], rng.integers(-99, 3000, len(part) - 1), np.round(rng.uniform(0.5, 60, len(part) - 1), 1),
                     rng.integers(0, 2, len(part) - 1), rng.integers(0, 2, len(part) - 1))]
        recs.append((int(part[0]), edges))
    g = pkg.ScaffoldGraphB200()
    g.set_vertex_names(names)
    a = PE.scaf_arrays(recs)
    cap = sum(len(names[r]) + 1 for r in a[0]) + sum(len(names[w]) + 80 for w in a[2])
    g.scaf_lines(*a, cap)
    g.set_profile(True)
    text = g.scaf_lines(*a, cap)
    prof = {k: round(v[0], 4) for k, v in g.profile().items()}
    assert text == PE.scaf_lines(names, recs)
    report = dict(records=n, edges=int(a[1][-1]), text_bytes=len(text), kernel_ms=prof)
    print("\n[scaf]", json.dumps(report))
    out = os.path.join(HERE, "..", "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "scaf_profile.json"), "w") as f:
            json.dump(report, f, indent=1)
