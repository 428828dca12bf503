"""Components and terminal vertices (gt_scaffolder_calc_cc_and_terminals and
gt_scaffolder_graph_isterminal, gt_scaffolder_algorithms.c:346-436).

The device computes label[v] = the smallest id that reaches v along unmarked edges between unmarked
vertices, and the terminal flags (gtsb_components); the binding turns them into the reference's
GtArray of GtArrays (singleton components directly, the others by the reference's own search
restricted to the component).

CPU: the label statement and the assembly, written out in numpy / Python below exactly as the
     kernels and the binding do them, against the COMPILED REFERENCE's own function -- on graphs
     after mark_repeats + filter, and with random (asymmetric) edge and vertex marks.
GPU: gtsb_components against the same numpy statement; the drop-in binary, whose removecycles and
     makescaffold stages call the binding's function, against the goldens (tests/test_dropin.py).
"""
import numpy as np
import pytest

import oracle_lib as O

NONE = 0xFFFFFFFF
V_MARKED = np.array([0, 1, 0, 1, 0, 0, 0, 1], bool)        # POLYMORPHIC, REPEAT, CYCLIC (algorithms.c:38-46)
E_MARKED = np.array([0, 1, 1, 1, 0, 0, 0, 1], bool)        # + INCONSISTENT (algorithms.c:49-58)
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not available")


@pytest.fixture(scope="module", autouse=True)
def _built():
    O.build_oracles()


def labels_and_terminals(src, dst, sense, estate, vstate):
    """the fixed point of k_cc_hook / k_cc_jump and the flags of k_cc_init, in numpy"""
    V = len(vstate)
    vm, em = V_MARKED[vstate], E_MARKED[estate]
    lab = np.where(vm, NONE, np.arange(V)).astype(np.int64)
    use = ~em & ~vm[src] & ~vm[dst]
    s, d = src[use].astype(np.int64), dst[use].astype(np.int64)
    while True:
        before = lab.copy()
        np.minimum.at(lab, d, lab[s])
        live = lab != NONE
        lab[live] = np.minimum(lab[live], lab[lab[live]])
        if np.array_equal(before, lab):
            break
    dirs = np.zeros(V, np.int64)
    np.bitwise_or.at(dirs, src[~em].astype(np.int64), np.where(sense[~em] != 0, 2, 1))
    return lab.astype(np.uint32), (dirs != 3).astype(np.uint8)


def assemble(lab, term, row_ptr, adj_dst, adj_estate):
    """what the binding builds from the device's answer: the reference's list of components"""
    V = len(lab)
    size = np.bincount(lab[lab != NONE].astype(np.int64), minlength=V)
    seen = np.zeros(V, bool)
    ccs = []
    for r in range(V):
        if lab[r] != r:
            continue
        if size[r] == 1:
            ccs.append([r] if term[r] else [])
            continue
        out, queue = [], [r]
        seen[r] = True
        while queue:                                   # algorithms.c:409-431 inside the component
            v = queue.pop(0)
            if term[v]:
                out.append(v)
            for k in range(int(row_ptr[v]), int(row_ptr[v + 1])):
                w = int(adj_dst[k])
                if not E_MARKED[adj_estate[k]] and lab[w] == r and not seen[w]:
                    seen[w] = True
                    queue.append(w)
        ccs.append(out)
    return ccs


def graph_arrays(g):
    e = g.edges()
    row_ptr, eids = g.adjacency()
    return e, row_ptr, e["dst"][eids], eids


def check_against_reference(g):
    e, row_ptr, adj_dst, eids = graph_arrays(g)
    vstate = g.vstate()
    lab, term = labels_and_terminals(e["src"], e["dst"], e["flags"] & 1, e["estate"], vstate)
    got = assemble(lab, term, row_ptr, adj_dst, e["estate"][eids])
    exp = g.calc_cc()
    assert got == exp
    # side effect of the reference's search: every unmarked vertex ends up GIS_VISITED (4)
    after = g.vstate()
    assert np.array_equal(after, np.where(V_MARKED[vstate], vstate, 4))
    return lab, term


@needs_ref
@pytest.mark.parametrize("seed", range(6))
def test_labels_and_assembly_equal_the_reference(seed, synth):
    rng = np.random.default_rng(seed)
    inp = synth.tiny_dense(30 + 40 * seed, 60 + 150 * seed, 9000 + seed) if seed < 3 else \
        synth.generate("c2_bacterial", V=1500 * seed, seed=40 + seed)
    g = O.RefGraph.build(inp)
    g.mark_repeats(0.3, 20.0)
    g.filter(0.01, 1.5, 400)
    check_against_reference(g)
    # random marks: asymmetric edge states, every state value on vertices and edges
    for p_edge, p_vertex in ((0.3, 0.1), (0.7, 0.0), (0.05, 0.4), (0.0, 0.0), (1.0, 0.0)):
        es = np.where(rng.random(g.E) < p_edge, rng.choice([1, 2, 3, 7], g.E), rng.choice([0, 4, 5, 6], g.E))
        vs = np.where(rng.random(g.V) < p_vertex, rng.choice([1, 3, 7], g.V), rng.choice([0, 2, 4, 5, 6], g.V))
        g.set_states(vs.astype(np.uint8), es.astype(np.uint8))
        check_against_reference(g)
    g.close()


@needs_ref
def test_chain_and_one_way_marks(synth):
    """a long path (pointer jumping), and one-way marks that split what an undirected search would join"""
    n = 400
    S = synth
    z = S.tiny_dense(n, 1, 3)
    u32 = lambda a: np.asarray(a, np.uint32)
    a = np.arange(n - 1)
    order = np.random.default_rng(1).permutation(n).astype(np.uint32)         # path through shuffled ids
    root = np.concatenate([order[a], order[a + 1]])
    ctg = np.concatenate([order[a + 1], order[a]])
    k = np.argsort(root, kind="stable")
    inp = S.ScaffoldInput(z.seq_len, z.astat, z.copy_num, u32(root[k]), u32(ctg[k]),
                          np.full(2 * (n - 1), 100, np.int32), np.full(2 * (n - 1), 5.0, np.float32),
                          np.full(2 * (n - 1), 20, np.uint32), np.full(2 * (n - 1), 3, np.uint8))
    g = O.RefGraph.build(inp)
    g.set_states(np.zeros(g.V, np.uint8), np.zeros(g.E, np.uint8))
    lab, _ = check_against_reference(g)
    assert (lab == 0).all()
    e = g.edges()
    es = np.zeros(g.E, np.uint8)
    es[(e["src"] > e["dst"]) & (np.arange(g.E) % 3 == 0)] = 2                   # some edges only one way
    g.set_states(np.zeros(g.V, np.uint8), es)
    check_against_reference(g)
    g.close()


# ------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_device_labels_equal_the_numpy_statement(pkg, synth, seed):
    rng = np.random.default_rng(50 + seed)
    inp = synth.generate("c2_bacterial", V=20000 * (seed + 1), seed=60 + seed)
    g = pkg.ScaffoldGraphB200.new_from_records(inp)
    g.mark_repeats(0.3, 20.0, True)
    g.filter(0.01, 1.5, 400)
    for trial in range(3):
        e, vs = g.edges(), g.vstate()
        lab, term = g.components()
        elab, eterm = labels_and_terminals(e["src"], e["dst"], e["flags"] & 1, e["estate"], vs)
        assert np.array_equal(lab, elab) and np.array_equal(term, eterm)
        # next trial: random asymmetric marks (by eid)
        E = len(e["eid"])
        es = np.where(rng.random(E) < 0.3 * (trial + 1), rng.choice([1, 2, 3, 7], E), 0).astype(np.uint8)
        nv = np.where(rng.random(len(vs)) < 0.1, rng.choice([1, 3, 7], len(vs)), 0).astype(np.uint8)
        g.set_states(nv, es)
    g.close()
