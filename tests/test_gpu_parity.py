"""GPU parity tests: the CUDA path through the C ABI against the oracles
(compiled reference when oracle/_ref is present, else the C restatement) and
against the committed golden vectors.  Bit-exact: integer/byte/index work and
IEEE-exact float decisions."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("src", "dst", "dist", "std_dev", "flags", "row_ptr", "adj_eid", "vstate", "estate")
DEFAULT = dict(cn_cut=0.3, a_cut=20.0, use_cn=True, pc=0.01, cnc=1.5, oc=400)


def _bits(a):
    """floats are compared bit for bit (NaN == NaN, -0.0 != 0.0)"""
    a = np.asarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def _cmp(got, exp, what=""):
    for k in KEYS:
        if not np.array_equal(_bits(got[k]), _bits(exp[k])):
            bad = np.nonzero(np.asarray(got[k]) != np.asarray(exp[k]))[0] if got[k].shape == exp[k].shape else []
            raise AssertionError(f"{what}: {k} differs at {len(bad)} positions, first {bad[:8]}; "
                                 f"got {np.asarray(got[k])[bad[:8]]} exp {np.asarray(exp[k])[bad[:8]]}")


def _run_both(pkg, inp, cn_cut, a_cut, use_cn, pc, cnc, oc, stagewise=True, force_general=False):
    g = pkg.ScaffoldGraphB200.new_from_records(inp, force_general=force_general)
    ref = O.best_oracle().build(inp)
    if stagewise:
        _cmp(g.result(), ref.result(), "build")
    g.mark_repeats(cn_cut, a_cut, use_cn)
    ref.mark_repeats(cn_cut, a_cut, use_copy_num=use_cn)
    if stagewise:
        _cmp(g.result(), ref.result(), "mark_repeats")
    g.filter(pc, cnc, oc)
    ref.filter(pc, cnc, oc)
    _cmp(g.result(), ref.result(), "filter")
    st = g.stats()
    g.close()
    ref.close()
    return st


PARAMS = [(0.3, 20.0, True, 0.01, 1.5, 400), (0.3, 20.0, True, 0.01, 1.5, 0),
          (0.3, 20.0, False, 0.01, 1.5, -1), (0.5, 19.5, True, 0.2, 2.5, 50),
          (0.3, 20.0, True, 0.5, 1.5, 400), (0.0, -1e9, True, -0.5, 9.0, 3000)]


@pytest.mark.parametrize("seed", range(60))
def test_tiny_adversarial(pkg, synth, seed):
    """Shuffled records / one-sided links: mostly the general build path."""
    inp = synth.tiny_dense(4 + seed % 13, 6 + 3 * (seed % 17), 3000 + seed, split_lines=bool(seed % 2))
    _run_both(pkg, inp, *PARAMS[seed % len(PARAMS)])


@pytest.mark.parametrize("seed", range(40))
def test_small_line_ordered(pkg, synth, seed):
    """Small .de-shaped inputs that stay on the line-ordered build: shuffled or
    id-ordered lines, repeated estimates on either line, links listed only on
    the earlier line, vertices without a line."""
    V = 6 + 7 * (seed % 9)
    inp = synth.generate("c2_bacterial", V=V, seed=900 + seed, mean_pairs=1.0 + (seed % 5),
                         line_order="id" if seed % 2 else "shuffled",
                         one_sided_frac=0.25 if seed % 3 == 0 else 0.0, one_sided_up=True,
                         mirror_diff_frac=0.3, dup_same_line_frac=0.2)
    st = _run_both(pkg, inp, *PARAMS[seed % len(PARAMS)])
    assert st["line_ordered_build"] == 1, st
    _run_both(pkg, inp, *PARAMS[seed % len(PARAMS)], force_general=True)


def test_fallback_reasons(pkg, synth, hub_V=20000):
    """Inputs outside the fast path's preconditions must be detected, not guessed."""
    # a link listed only on the later line
    inp = synth.generate("c2_bacterial", V=3000, one_sided_frac=0.3)
    st = _run_both(pkg, inp, *PARAMS[0], stagewise=False)
    assert st["line_ordered_build"] == 0 and st["fallback_reason"] & 8
    # hubs: lines longer than the per-thread scans accept
    inp = synth.generate("c4_repeat_hubs", V=hub_V, max_deg=500)
    st = _run_both(pkg, inp, *PARAMS[0], stagewise=False)
    assert st["line_ordered_build"] == 0 and st["fallback_reason"] & (2 | 4)
    # records not grouped by root
    inp = synth.tiny_dense(12, 60, 5)
    st = _run_both(pkg, inp, *PARAMS[0], stagewise=False)
    assert st["line_ordered_build"] == 0 and st["fallback_reason"] & 1


def test_empty_and_degenerate(pkg, synth):
    # vertices without any record; a single pair; records only one-sided
    inp = synth.tiny_dense(5, 1, 1)
    empty = synth.ScaffoldInput(inp.seq_len, inp.astat, inp.copy_num, inp.root[:0], inp.ctg[:0],
                                inp.dist[:0], inp.std_dev[:0], inp.num_pairs[:0], inp.flags[:0])
    g = pkg.ScaffoldGraphB200.new_from_records(empty)
    g.mark_repeats()
    g.filter()
    assert g.E == 0
    ref = O.PortGraph(empty)
    ref.mark_repeats(0.3, 20.0)
    assert np.array_equal(g.vstate(), ref.vstate())
    _run_both(pkg, inp, *PARAMS[0])


def test_invalid_records_are_refused(pkg, synth):
    inp = synth.tiny_dense(5, 4, 2)
    bad = synth.ScaffoldInput(inp.seq_len, inp.astat, inp.copy_num, inp.root.copy(), inp.ctg.copy(),
                              inp.dist, inp.std_dev, inp.num_pairs, inp.flags)
    bad.ctg[0] = 99
    with pytest.raises(RuntimeError):
        pkg.ScaffoldGraphB200.new_from_records(bad)
    bad.ctg[0] = bad.root[0]
    with pytest.raises(RuntimeError):
        pkg.ScaffoldGraphB200.new_from_records(bad)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "diff_*.npz"))))
def test_committed_differential_vectors(pkg, synth, path):
    z = np.load(path)
    inp = synth.ScaffoldInput(**{k: z[k] for k in ["seq_len", "astat", "copy_num", "root", "ctg",
                                                   "dist", "std_dev", "num_pairs", "flags"]})
    pc, cnc, oc, cn_cut, a_cut, use_cn = z["params"]
    g = pkg.ScaffoldGraphB200.new_from_records(inp)
    r = g.result()
    for k, zk in [("src", "e_src"), ("dst", "e_dst"), ("dist", "e_dist"), ("std_dev", "e_std"),
                  ("flags", "e_flags"), ("row_ptr", "row_ptr"), ("adj_eid", "adj_eid")]:
        assert np.array_equal(_bits(r[k]), _bits(z[zk])), k
    g.mark_repeats(float(cn_cut), float(a_cut), bool(use_cn))
    r = g.result()
    assert np.array_equal(r["vstate"], z["rep_vstate"]) and np.array_equal(r["estate"], z["rep_estate"])
    g.filter(float(pc), float(cnc), int(oc))
    r = g.result()
    assert np.array_equal(r["vstate"], z["fin_vstate"]) and np.array_equal(r["estate"], z["fin_estate"])


@pytest.mark.parametrize("name,V,kw", [
    ("c2_bacterial", None, {}),
    ("c2_bacterial", 30_000, dict(line_order="id", one_sided_frac=0.2)),
    ("c3_human", 400_000, {}),
    ("c4_repeat_hubs", 150_000, dict(max_deg=3000)),
])
def test_named_configs(pkg, synth, name, V, kw):
    inp = synth.generate(name, V=V, **kw)
    for force in (False, True):
        st = _run_both(pkg, inp, **{k: v for k, v in zip(
            ["cn_cut", "a_cut", "use_cn", "pc", "cnc", "oc"], PARAMS[0])}, stagewise=False,
            force_general=force)
        assert st["nof_edges"] > 0
        if not force and not kw and name != "c4_repeat_hubs":
            assert st["line_ordered_build"] == 1, st


def test_win_rec_points_at_the_winning_record(pkg, synth):
    inp = synth.tiny_dense(12, 60, 77)
    g = pkg.ScaffoldGraphB200.new_from_records(inp, want_win_rec=True)
    c = g.csr(win_rec=True)
    ref = O.PortGraph(inp)
    e = ref.edges()
    order = np.argsort(c["eid"])
    win = c["win_rec"][order]
    assert np.array_equal((win & 0x7FFFFFFF).astype(np.int64), e["win_rec"])
    # seeded <=> the edge carries the twin seed of a record of the other direction
    seeded = (win >> 31).astype(bool)
    rec_root = inp.root[(win & 0x7FFFFFFF)]
    assert np.array_equal(seeded, rec_root != e["src"])


def test_win_rec_on_the_line_ordered_build(pkg, synth):
    """The same on .de-shaped input, which stays on the line-ordered build (the drop-in binding
    needs the winning record of every edge for GtScaffolderGraphEdge.num_pairs)."""
    for seed in range(6):
        inp = synth.generate("c2_bacterial", V=300 + 500 * seed, seed=500 + seed, mean_pairs=2.0 + seed % 3,
                             mirror_diff_frac=0.4, dup_same_line_frac=0.3, one_sided_frac=0.2, one_sided_up=True)
        g = pkg.ScaffoldGraphB200.new_from_records(inp, want_win_rec=True)
        assert g.stats()["line_ordered_build"] == 1
        c = g.csr(win_rec=True)
        ref = O.PortGraph(inp)
        e = ref.edges()
        order = np.argsort(c["eid"])
        win = c["win_rec"][order]
        assert np.array_equal((win & 0x7FFFFFFF).astype(np.int64), e["win_rec"]), seed
        seeded = (win >> 31).astype(bool)
        assert np.array_equal(seeded, inp.root[(win & 0x7FFFFFFF)] != e["src"]), seed
        g.close()


def test_filter_on_uploaded_graph_with_arbitrary_states(pkg, synth):
    """gtsb_set_graph_host path (what the GtScaffolderGraph binding uses)."""
    rng = np.random.default_rng(11)
    for seed in range(12):
        inp = synth.tiny_dense(12, 50, 4000 + seed)
        ref = O.best_oracle().build(inp)
        res = ref.result()
        E, V = len(res["src"]), len(res["vstate"])
        vs = rng.choice([0, 0, 0, 1, 3, 7, 4], V).astype(np.uint8)
        es = rng.choice([0, 0, 0, 1, 2, 3, 7, 6], E).astype(np.uint8)
        # CSR in adjacency order with reverse-edge flags
        eids = res["adj_eid"].astype(np.int64)
        rev = {(int(s), int(d)): i for i, (s, d) in enumerate(zip(res["src"], res["dst"]))}
        flags = np.zeros(E, np.uint8)
        for slot, e in enumerate(eids):
            r = rev[(int(res["dst"][e]), int(res["src"][e]))]
            flags[slot] = res["flags"][e] | ((res["flags"][r] & 3) << 2)
        g = pkg.ScaffoldGraphB200()
        g.set_graph(res["row_ptr"].astype(np.uint32), res["dst"][eids], res["dist"][eids].astype(np.int32),
                    res["std_dev"][eids], flags, inp.seq_len, inp.astat, inp.copy_num, vs, es[eids])
        g.filter(0.01, 1.5, 400)
        ref.set_states(vs, es)
        ref.filter(0.01, 1.5, 400)
        c = g.csr(eid=False)
        assert np.array_equal(g.vstate(), ref.vstate()), seed
        assert np.array_equal(c["estate"], ref.estate()[eids]), seed


def test_line_shaped_input_and_states_by_eid(pkg, synth, V=20_000):
    """gtsb_set_record_lines_host / gtsb_get_edge_states: the same graph and marks as the
    flat record input, with 4 B/record less on the way in and 1 B/edge on the way out."""
    inp = synth.generate("c2_bacterial", V=V, seed=21, mirror_diff_frac=0.1, dup_same_line_frac=0.05)
    a = pkg.ScaffoldGraphB200.new_from_records(inp)
    a.mark_repeats()
    a.filter()
    ra = a.result()
    b = pkg.ScaffoldGraphB200()
    b.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
    line_root, line_start = pkg.api.lines_of(inp.root)
    b.set_record_lines(line_root, line_start, inp.ctg, inp.dist, inp.std_dev, inp.flags)
    b.pipeline()
    rb = b.result()
    _cmp(rb, ra, "line-shaped input")
    assert np.array_equal(b.edge_states(), ra["estate"])
    assert np.array_equal(a.edge_states(), ra["estate"])
    # records before vertices, twice in a row on one context (what bench.py's e2e leg does: the
    # late columns and the vertex attributes travel on the copy stream under the first kernels)
    for rep in range(2):
        b.set_record_lines(line_root, line_start, inp.ctg, inp.dist, inp.std_dev, inp.flags)
        b.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
        b.pipeline()
        assert np.array_equal(b.edge_states(), ra["estate"]) and np.array_equal(b.vstate(), ra["vstate"])
    # and a flat record upload right after a line-shaped one
    b.set_records(inp.root, inp.ctg, inp.dist, inp.std_dev, inp.flags)
    b.pipeline()
    _cmp(b.result(), ra, "flat after line-shaped")


def test_full_size_properties_c3(pkg, synth):
    """BASELINE.json's human-scale graph at full size (10^7 contigs, 8*10^7 edges), where the
    oracle is too slow: size-independent properties of the result, and the two independent
    build algorithms (line-ordered and general sort-based) against each other."""
    import torch
    t = synth.generate_torch("c3_human", device="cuda")
    V, R = int(t["seq_len"].shape[0]), int(t["root"].shape[0])

    def run(force_general):
        g = pkg.ScaffoldGraphB200(force_general=force_general)
        g.set_vertices_device(V, t["seq_len"].data_ptr(), t["astat"].data_ptr(), t["copy_num"].data_ptr())
        g.set_records_device(R, t["root"].data_ptr(), t["ctg"].data_ptr(), t["dist"].data_ptr(),
                             t["std_dev"].data_ptr(), t["flags"].data_ptr())
        g.pipeline()
        e, vs, st = g.edges(), g.vstate(), g.stats()
        again = None
        if not force_general:
            g.pipeline()                                   # idempotent on the same inputs
            again = (g.edge_states(), g.vstate())
        g.close()
        order = np.argsort(e["eid"], kind="stable")
        return {k: v[order] for k, v in e.items()}, vs, st, again

    e, vs, st, again = run(False)
    assert st["line_ordered_build"] == 1
    E = e["eid"].shape[0]
    assert E % 2 == 0 and np.array_equal(e["eid"], np.arange(E, dtype=np.uint32))
    assert np.array_equal(again[0], e["estate"]) and np.array_equal(again[1], vs)
    # edges come in mutually reverse pairs 2k / 2k+1 (parser.c:374-377)
    assert np.array_equal(e["src"][0::2], e["dst"][1::2]) and np.array_equal(e["dst"][0::2], e["src"][1::2])
    # mark_repeats closed form (algorithms.c:160-166, 61-87)
    astat, cn = t["astat"].cpu().numpy(), t["copy_num"].cpu().numpy()
    pred = (astat <= np.float32(20.0)) | (cn < np.float32(0.3))
    assert np.array_equal(vs == 3, pred)
    assert not np.any((vs == 1) & pred)
    touched = pred[e["src"]] | pred[e["dst"]]
    assert np.all(e["estate"][touched] != 0)
    assert np.all(touched[e["estate"] == 3])
    # every edge at a polymorphic contig is marked (mark_vertex, algorithms.c:76-87)
    poly = vs == 1
    at_poly = poly[e["src"]] | poly[e["dst"]]
    assert np.all(np.isin(e["estate"][at_poly], (1, 2)))
    assert set(np.unique(e["estate"]).tolist()) <= {0, 1, 2, 3}
    # the general build is a different algorithm; same graph, same marks
    e2, vs2, st2, _ = run(True)
    assert st2["line_ordered_build"] == 0
    for k in ("src", "dst", "dist", "std_dev", "flags", "estate"):
        assert np.array_equal(e[k], e2[k]), k
    assert np.array_equal(vs, vs2)
    del t
    torch.cuda.empty_cache()


def _full_size_vs_reference(pkg, synth, name, **kw):
    """One named config at FULL size: the CUDA path (device-resident inputs, gtsb_pipeline) against
    the compiled, unmodified reference run on the box's host (oracle/_ref; the C restatement if
    absent) on the same arrays: every edge attribute and state in graph->edges[] order, the
    adjacency order of every vertex, every vertex state."""
    import time
    import torch
    t = synth.generate_torch(name, device="cuda", **kw)
    V, R = int(t["seq_len"].shape[0]), int(t["root"].shape[0])
    inp = synth.torch_to_input(t)
    g = pkg.ScaffoldGraphB200()
    g.set_vertices_device(V, t["seq_len"].data_ptr(), t["astat"].data_ptr(), t["copy_num"].data_ptr())
    g.set_records_device(R, t["root"].data_ptr(), t["ctg"].data_ptr(), t["dist"].data_ptr(),
                         t["std_dev"].data_ptr(), t["flags"].data_ptr())
    g.pipeline(*[DEFAULT[k] for k in ("cn_cut", "a_cut", "use_cn", "pc", "cnc", "oc")])
    got, st = g.result(), g.stats()
    g.close()
    del t
    torch.cuda.empty_cache()
    t0 = time.time()
    ref = O.best_oracle().build(inp)
    ref.mark_repeats(DEFAULT["cn_cut"], DEFAULT["a_cut"], use_copy_num=DEFAULT["use_cn"])
    ref.filter(DEFAULT["pc"], DEFAULT["cnc"], DEFAULT["oc"])
    sec = time.time() - t0
    exp = ref.result()
    ref.close()
    print(f"{name}: V={V} R={R} E={len(exp['src'])} reference {sec:.1f} s on one host core; "
          f"device stats {st}")
    _cmp(got, exp, name + " at full size")
    # the filter did real work at this size
    assert (exp["estate"] == 1).sum() > 0 and (exp["estate"] == 2).sum() > 0 and (exp["vstate"] == 1).sum() > 0
    return st


def test_full_size_c3_vs_reference(pkg, synth):
    """BASELINE.json config 3 as named: 10^7 contigs, ~8*10^7 directed edges (algorithms.c:261-343
    and parser.c:357-379 run by the reference itself, ~25 s and ~5 GB on the host)."""
    st = _full_size_vs_reference(pkg, synth, "c3_human")
    assert st["nof_vertices"] == 10_000_000 and st["line_ordered_build"] == 1


def _full_size_vs_golden_digest(pkg, synth, name):
    """A named config at full size against tests/golden/full_size_<name>.json: sha256 of every
    result array of the compiled, unmodified reference, made in the build container by
    tests/golden/make_full_size_digest.py (the reference needs ~13 min for config 4).  The inputs
    are regenerated here by the same seeded numpy generator and checked by hash first."""
    import hashlib
    import json
    sys_path = os.path.join(HERE, "golden")
    gold = json.load(open(os.path.join(sys_path, "full_size_%s.json" % name)))
    inp = synth.generate(name, **gold["kwargs"])
    for k, h in gold["input_sha256"].items():
        assert hashlib.sha256(np.ascontiguousarray(getattr(inp, k)).tobytes()).hexdigest() == h, \
            f"generator drifted: input {k} is not the one the golden digest was made from"
    g = pkg.ScaffoldGraphB200()
    g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
    line_root, line_start = pkg.api.lines_of(inp.root)
    g.set_record_lines(line_root, line_start, inp.ctg, inp.dist, inp.std_dev, inp.flags)
    g.pipeline(*[DEFAULT[k] for k in ("cn_cut", "a_cut", "use_cn", "pc", "cnc", "oc")])
    got, st = g.result(), g.stats()
    g.close()
    dt = dict(src=np.uint32, dst=np.uint32, dist=np.int64, std_dev=np.float32, flags=np.uint8,
              row_ptr=np.uint64, adj_eid=np.uint32, vstate=np.uint8, estate=np.uint8)
    assert len(got["src"]) == gold["E"]
    assert np.bincount(got["estate"], minlength=8).tolist() == gold["states"]["estate"]
    assert np.bincount(got["vstate"], minlength=8).tolist() == gold["states"]["vstate"]
    for k in KEYS:
        a = np.ascontiguousarray(got[k]).astype(dt[k], copy=False)
        assert a.shape[0] == gold["digest"][k]["n"], k
        assert hashlib.sha256(a.tobytes()).hexdigest() == gold["digest"][k]["sha256"], \
            f"{name} at full size: {k} differs from the reference's result"
    return st, gold


def test_full_size_c4_hubs_vs_reference_digest(pkg, synth):
    """BASELINE.json config 4 as named: 5*10^6 contigs, power-law degrees with hubs of up to 10^4
    edges (the reference's O(d^2) paths: find_edge in parser.c:359, the pair loops of
    algorithms.c:283-320, mark_edge's twin scan :61-73 -- 88 s + 72 s + 544 s on one core)."""
    st, gold = _full_size_vs_golden_digest(pkg, synth, "c4_repeat_hubs")
    assert st["nof_vertices"] == 5_000_000 and st["max_degree"] == gold["max_degree"] >= 10_000


def test_full_size_c3_vs_reference_digest(pkg, synth):
    """Config 3 once more, through the host-buffer entry points and the numpy generator, against
    the digest (the direct comparison above uses the torch generator and device buffers)."""
    st, _ = _full_size_vs_golden_digest(pkg, synth, "c3_human")
    assert st["line_ordered_build"] == 1


@pytest.mark.skipif(os.environ.get("GTSB_FULL_REFERENCE") != "1",
                    reason="~13 min of reference CPU time on the box; set GTSB_FULL_REFERENCE=1 "
                           "(the digest test above pins the same result)")
def test_full_size_c4_hubs_vs_reference(pkg, synth):
    st = _full_size_vs_reference(pkg, synth, "c4_repeat_hubs", max_deg=10_000)
    assert st["nof_vertices"] == 5_000_000 and st["max_degree"] > 5_000


@pytest.mark.parametrize("seed", range(6))
def test_special_values(pkg, synth, seed):
    """NaN / inf / zero / denormal std_dev and copy numbers, NaN a-statistics, extreme
    distances and contig lengths: the float compares and the i64 interval arithmetic must
    fall the reference's way (algorithms.c:174-246, parser.c:362)."""
    bad = synth.special_values(seed)
    _run_both(pkg, bad, *PARAMS[seed % len(PARAMS)])
    _run_both(pkg, bad, *PARAMS[seed % len(PARAMS)], force_general=True)


@pytest.mark.gpu
def test_result_digest_equals_its_numpy_statement(pkg, synth):
    """gtsb_result_digest (what tools/c5_check.py compares between a partitioned and a single-device
    run of graphs too large to fetch) == the same sums over the fetched arrays; it moves when one
    state moves."""
    inp = synth.generate("c2_bacterial", V=30000, seed=3)
    g = pkg.ScaffoldGraphB200.new_from_records(inp)
    g.mark_repeats(0.3, 20.0, True)
    g.filter(0.01, 1.5, 400)
    e, vs = g.edges(), g.vstate()
    assert g.digest() == pkg.api.result_digest(e, vs)
    e2 = dict(e)
    e2["estate"] = e["estate"].copy()
    e2["estate"][len(e2["estate"]) // 2] ^= 1
    assert pkg.api.result_digest(e2, vs)[1] != g.digest()[1]
    # and equals the digest of the oracle's result on the same input
    ref = O.best_oracle().build(inp)
    ref.mark_repeats(0.3, 20.0, use_copy_num=True)
    ref.filter(0.01, 1.5, 400)
    r = ref.result()
    oe = dict(eid=np.arange(len(r["src"]), dtype=np.uint32), src=r["src"], dst=r["dst"], dist=r["dist"],
              std_dev=r["std_dev"], flags=e["flags"][np.argsort(e["eid"])], estate=r["estate"])
    assert pkg.api.result_digest(oe, r["vstate"]) == g.digest()
    g.close()


@pytest.mark.gpu
def test_more_proposals_than_the_list_was_sized_for(pkg, synth):
    """The proposal list holds E / 4 entries; with every pair ambiguous and a copy-number cutoff that
    nothing exceeds most slots are proposed, the pairs pass overflows and is run again with room
    for all of them."""
    inp = synth.generate("c3_human", V=120000, seed=21)
    st = _run_both(pkg, inp, 0.0, -1e9, True, -0.5, 9.0, 3000, stagewise=False)
    assert st["proposals"] > st["nof_edges"] // 4 + 65536, st


@pytest.mark.gpu
def test_stats_report_the_longest_row_on_both_build_paths(pkg, synth):
    inp = synth.generate("c2_bacterial", V=30000, seed=9)
    for force_general in (False, True):
        g = pkg.ScaffoldGraphB200.new_from_records(inp, force_general=force_general)
        st, r = g.stats(), g.result()
        assert st["line_ordered_build"] == (0 if force_general else 1)
        assert st["max_degree"] == int(np.diff(r["row_ptr"].astype(np.int64)).max()), st
        g.close()
