"""The product's CUDA sources RUN ON THE HOST (no GPU): gt-scaffold_b200/csrc/*.cu compiled by g++
against tests/emul/cusim/ -- a functional model of the CUDA runtime and device language in which
the threads of a block are fibers and barriers, shuffles, ballots and reductions are rendezvous
(tests/emul/cusim_build.py, TEST INFRASTRUCTURE).  What runs here is the kernels' own logic:
shared-memory staging, warp windows, counting sorts, the fix-point rounds (cooperative launch
included), the hub routes, the text kernels.  It is checked the way the GPU tests check the device:
against the compiled reference / the oracle, the committed golden vectors, the host builds of the
text stages, and -- through the reference's own driver, linked with the binding and pointed at the
emulated library -- against the golden `.dot` / `.scaf` files of config 1.

The functions called below ARE the GPU tests (tests/test_gpu_parity.py and friends, `-m gpu`): same
bodies, same assertions, a package whose ctypes loader was handed the emulated library.  Nothing in
the product knows the emulation exists; it is not a fallback (gtsb_create of the real library still
fails without a CUDA device)."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))

import cusim_build  # noqa: E402
import oracle_lib as O  # noqa: E402
import test_components as TC  # noqa: E402
import test_dropin as TD  # noqa: E402
import test_exhaustive as TE  # noqa: E402
import test_format as TF  # noqa: E402
import test_gpu_parity as G  # noqa: E402
import test_mle as TM  # noqa: E402
import test_parse as TP  # noqa: E402
import test_scaf as TS  # noqa: E402


@pytest.fixture(scope="module")
def sim(pkg):
    """The package with the emulated library behind its ctypes loader (restored afterwards)."""
    api = pkg.api
    saved = (api.LIB_PATH, api._lib)
    api.LIB_PATH, api._lib = cusim_build.build(), None
    api.load_library()
    yield pkg
    api.LIB_PATH, api._lib = saved


def test_the_real_library_still_needs_a_device(pkg):
    """The emulation is handed to the loader explicitly; the product's own library has no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert os.path.basename(pkg.api.LIB_PATH) == "libgtscaffold_b200.so"
    with pytest.raises(RuntimeError):
        pkg.ScaffoldGraphB200()


# ------------------------------------------------------------------ build + mark_repeats + filter

@pytest.mark.parametrize("seed", range(0, 60, 3))
def test_tiny_adversarial(sim, synth, seed):
    G.test_tiny_adversarial(sim, synth, seed)


@pytest.mark.parametrize("seed", range(0, 40, 5))
def test_small_line_ordered(sim, synth, seed):
    G.test_small_line_ordered(sim, synth, seed)


def test_fallbacks_degenerate_and_refused_inputs(sim, synth):
    G.test_fallback_reasons(sim, synth, hub_V=2500)
    G.test_empty_and_degenerate(sim, synth)
    G.test_invalid_records_are_refused(sim, synth)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "diff_*.npz"))))
def test_committed_differential_vectors(sim, synth, path):
    G.test_committed_differential_vectors(sim, synth, path)


@pytest.mark.parametrize("name,V,kw", [
    ("c2_bacterial", 2500, {}),
    ("c2_bacterial", 2000, dict(line_order="id", one_sided_frac=0.2)),
    ("c3_human", 3000, {}),
    # hubs: rows above 256 slots (split pairs pass), buckets above 2048 entries (sort in several chunks)
    ("c4_repeat_hubs", 4000, dict(max_deg=1500)),
])
def test_named_shapes(sim, synth, name, V, kw):
    inp = synth.generate(name, V=V, **kw)
    for force in (False, True):
        st = G._run_both(sim, inp, **G.DEFAULT, stagewise=False, force_general=force)
        assert st["nof_edges"] > 0
        if name == "c4_repeat_hubs":
            assert st["max_degree"] > 1024 and st["big_rows"] > 0 and st["large_buckets"] > 0
            assert st["fallback_reason"] == (0 if force else 2), st        # a long line: the hub routes
        elif not kw:
            assert st["line_ordered_build"] == (0 if force else 1), st


def test_the_other_entry_points(sim, synth):
    G.test_win_rec_points_at_the_winning_record(sim, synth)
    G.test_win_rec_on_the_line_ordered_build(sim, synth)
    G.test_filter_on_uploaded_graph_with_arbitrary_states(sim, synth)
    G.test_line_shaped_input_and_states_by_eid(sim, synth, V=2000)


@pytest.mark.parametrize("seed", range(6))
def test_special_values(sim, synth, seed):
    G.test_special_values(sim, synth, seed)


def test_pipeline_is_the_three_calls_and_the_digest_is_its_numpy_statement(sim, synth):
    inp = synth.generate("c2_bacterial", V=1500, seed=12)
    g = sim.ScaffoldGraphB200.new_from_records(inp)
    g.mark_repeats(0.3, 20.0, True)
    g.filter(0.01, 1.5, 400)
    a = g.result()
    e, vs = g.edges(), g.vstate()
    assert g.digest() == sim.api.result_digest(e, vs)
    g.close()
    g = sim.ScaffoldGraphB200()
    g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
    g.set_records(inp.root, inp.ctg, inp.dist, inp.std_dev, inp.flags)
    g.pipeline(0.3, 20.0, True, 0.01, 1.5, 400)
    b = g.result()
    g.close()
    for k in G.KEYS:
        assert np.array_equal(G._bits(a[k]), G._bits(b[k])), k


@pytest.mark.parametrize("kind,offset", [("lines", 0), ("runs", 5), ("any", 11)])
def test_a_sample_of_every_small_case(sim, synth, kind, offset):
    """tests/test_exhaustive.py (every record sequence of length <= 2 over three contigs, samples of
    length 3 and 4, six attribute settings, as the components of one graph) -- every 37th sequence,
    on both build paths and under both parameter sets."""
    TE.O.build_oracles()
    TE.test_cuda_path_equals_oracle_on_every_small_case(sim, synth, kind, stride=37, offset=offset)


@pytest.mark.parametrize("order", [1, 7, 12345])
def test_results_do_not_depend_on_the_order_of_the_threads(sim, synth, order):
    """The emulation gives the threads of a block their turns in ascending order by default; here
    descending (1) and pseudo-random per pass (seeded).  Between two rendezvous a thread may only
    touch what no other thread touches, so the graphs and marks must come out the same -- on the
    line-ordered build, the general build with its hub routes, and the filter with hub rows."""
    lib = sim.api.load_library()
    lib.cusim_set_order.argtypes = [__import__("ctypes").c_uint64]
    lib.cusim_set_order(order)
    try:
        for name, V, kw in (("c2_bacterial", 1500, {}), ("c4_repeat_hubs", 2500, dict(max_deg=1200)),
                            ("c2_bacterial", 1200, dict(line_order="id", one_sided_frac=0.2, mirror_diff_frac=0.2))):
            G._run_both(sim, synth.generate(name, V=V, seed=31, **kw), **G.DEFAULT, stagewise=False)
        G.test_tiny_adversarial(sim, synth, order % 60)
    finally:
        lib.cusim_set_order(0)


# ------------------------------------------------------------------ one graph over several ranks

def _partitioned(sim, inp, world, k=0, fail_at=None, reps=2):
    """gtsb_dist.cu with the ranks as THREADS of this process: one context per rank on the emulated
    device, NCCL replaced by a rendezvous stand-in (tests/emul/cusim/fake_nccl.cpp), peer memory by
    plain pointers.  Input handling per case as tests/dist_check.py does under torchrun."""
    import threading
    import dist_check as DC
    api = sim.api
    uid = api.dist_unique_id()
    out, err = [None] * world, [None] * world

    def rank_main(r):
        try:
            mine = api.shard_lines(inp, world, r)
            g = sim.ScaffoldGraphB200(device=0)
            g.dist_init(r, world, uid)
            if k % 3 != 2:
                g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
            if k % 2 == 1 and len(mine.root):
                g.set_record_lines(*api.lines_of(mine.root), mine.ctg, mine.dist, mine.std_dev, mine.flags)
            else:
                g.set_records(mine.root, mine.ctg, mine.dist, mine.std_dev, mine.flags)
            if k % 3 == 2:            # every rank uploads its slice of the contig attributes only
                V = inp.nof_vertices
                lo, hi = V * r // world, V * (r + 1) // world
                g._ck(g.L.gtsb_set_vertices_slice_host(
                    g.h, V, lo, hi - lo, api._ptr(np.ascontiguousarray(inp.seq_len[lo:hi], np.uint32)),
                    api._ptr(np.ascontiguousarray(inp.astat[lo:hi], np.float32)),
                    api._ptr(np.ascontiguousarray(inp.copy_num[lo:hi], np.float32))))
                g.synchronize()
                g.V = V
            outcomes = []
            for rep in range(reps):
                if fail_at is not None:
                    # the hook is read when a place is reached; every rank has left the call before it is cleared
                    barrier.wait()
                    if r == 0:
                        if rep == 0:
                            os.environ["GTSB_FAIL_AT"] = fail_at
                        else:
                            os.environ.pop("GTSB_FAIL_AT", None)
                    barrier.wait()
                try:
                    g.pipeline(**DC.PARAMS)
                    outcomes.append("ok")
                except RuntimeError as e:
                    outcomes.append("error: " + str(e))
            out[r] = (g.edges() if outcomes[-1] == "ok" else None, g.vstate() if outcomes[-1] == "ok" else None,
                      g.stats(), outcomes)
            g.close()
        except BaseException as e:      # noqa: BLE001 -- reported by the caller
            err[r] = e

    barrier = threading.Barrier(world)
    threads = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(240)
    os.environ.pop("GTSB_FAIL_AT", None)
    assert not any(t.is_alive() for t in threads), "a rank is still waiting in an exchange"
    assert not any(err), err
    return out


def _merged_equals_the_reference(out, inp):
    import dist_check as DC
    for o in out[1:]:
        assert np.array_equal(o[1], out[0][1]), "vertex states differ between ranks"
    got = DC.merged_result([o[0] for o in out], out[0][1])
    ref = O.best_oracle().build(inp)
    ref.mark_repeats(0.3, 20.0, use_copy_num=True)
    ref.filter(0.01, 1.5, 400)
    exp = ref.result()
    ref.close()
    for key in G.KEYS:
        assert np.array_equal(G._bits(got[key]), G._bits(exp[key])), key


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_partitioned_graph_equals_the_reference(sim, synth, world):
    """The rank-partitioned pipeline (line chunks, mail into the owners' receive buffers, the
    exchanges of vertex facts, proposals and fire rounds) at 2, 3, 4 and 8 ranks: every rank's edges
    merged by eid == the compiled reference, every attribute, state and the adjacency order."""
    import dist_check as DC
    names = ["one_record", "one_line", "tiny", "small_shuffled", "small_id_order", "mirror"]
    if world > 3:
        names = ["one_line", "tiny", "mirror"]
    for name in names:
        inp = DC.CASES[name](synth)
        out = _partitioned(sim, inp, world, k=list(DC.CASES).index(name))
        _merged_equals_the_reference(out, inp)
        assert all(o[2]["line_ordered_build"] == 1 for o in out)


@pytest.mark.parametrize("place,bad_rank", [("setup", 0), ("facts", 1), ("receive", 0), ("windows", 1),
                                            ("filter", 0), ("proposals", 1), ("fire", 1)])
def test_a_failure_on_one_rank_stops_every_rank(sim, synth, place, bad_rank):
    """tests/dist_fail_check.py on the emulated device: an allocation failure injected on one rank at
    a place where buffers grow comes back as an error from gtsb_pipeline on EVERY rank (nobody is left
    in a collective), and the next call on the same contexts succeeds and is correct."""
    inp = synth.generate("c2_bacterial", V=1500, seed=13, mirror_diff_frac=0.3, dup_same_line_frac=0.2)
    out = _partitioned(sim, inp, 2, k=0, fail_at=f"{bad_rank}:{place}")
    firsts = [o[3][0] for o in out]
    assert all(f.startswith("error") for f in firsts), firsts
    assert "injected failure" in firsts[bad_rank] and "rank %d failed" % bad_rank in firsts[1 - bad_rank], firsts
    assert all(o[3][1] == "ok" for o in out), [o[3] for o in out]
    _merged_equals_the_reference(out, inp)


def test_refused_inputs_come_back_from_every_rank(sim, synth):
    """What the partitioned build does not take (a line longer than 64 records,
    a contig heading several lines, a link listed only on the later line) and what no build takes
    (a self link, an unknown id): every rank returns the error -- twice in a row -- and nobody waits."""
    z = synth.generate("c2_bacterial", V=300, seed=8)

    def edited(value):
        ctg = z.ctg.copy()
        ctg[5] = value
        return synth.ScaffoldInput(z.seq_len, z.astat, z.copy_num, z.root, ctg, z.dist, z.std_dev, z.num_pairs, z.flags)

    cases = [(synth.generate("c4_repeat_hubs", V=1500, max_deg=300, seed=3), "reason mask 2"),
             (synth.tiny_dense(12, 60, 5), "heads more than one line"),
             (synth.generate("c2_bacterial", V=800, seed=4, one_sided_frac=0.3), "reason mask 8"),
             (edited(int(z.root[5])), "self link"), (edited(99999), "vertex id >= nof_vertices")]
    for world in (2, 3):
        for inp, what in cases:
            out = _partitioned(sim, inp, world, reps=2)
            for o in out:
                assert all(x.startswith("error") for x in o[3]), (what, o[3])
            assert any(what in o[3][0] for o in out), (what, [o[3][0] for o in out])


# ------------------------------------------------------------------ text kernels, components, MLE

@pytest.mark.parametrize("seed", [0, 1, 4])
def test_de_and_astat_tokenisers(sim, seed):
    TP.test_device_equals_emulation(sim, seed)
    TP.test_device_astat_equals_emulation(sim, seed)


def test_tokeniser_refusals(sim):
    TP.test_device_long_lines_and_degenerate(sim)
    TP.test_device_refuses_duplicate_headers(sim)


@pytest.mark.parametrize("seed", [0, 1])
def test_dot_and_scaf_writers(sim, synth, seed):
    TF.test_device_lines_equal_host_build(sim, synth, seed)
    TS.test_device_equals_host_build(sim, seed)


def test_components_and_terminals(sim, synth):
    rng = np.random.default_rng(50)
    inp = synth.generate("c2_bacterial", V=3000, seed=60)
    g = sim.ScaffoldGraphB200.new_from_records(inp)
    g.mark_repeats(0.3, 20.0, True)
    g.filter(0.01, 1.5, 400)
    e, vs = g.edges(), g.vstate()
    lab, term = g.components()
    elab, eterm = TC.labels_and_terminals(e["src"], e["dst"], e["flags"] & 1, e["estate"], vs)
    assert np.array_equal(lab, elab) and np.array_equal(term, eterm)
    assert len(np.unique(lab)) > 10 and rng is not None
    g.close()


def test_distance_mle(sim):
    kind, rf, seed = TM.CASES[0]
    TM.test_device_equals_host_build(sim, kind, rf, seed)


# ------------------------------------------------------------------ the reference's own driver

needs_bin = pytest.mark.skipif(not os.path.exists(TD.B200_TESTX), reason="integration/_build/test_b200.x not built")


@needs_bin
@pytest.mark.parametrize("tokeniser", [None, "host"])
def test_config1_goldens_through_the_binding(tmp_path, tokeniser, monkeypatch):
    """test.c of the reference, unmodified, linked with the binding; the library it calls is the
    emulated one (symbols resolved from it first): the four stage `.dot` files and the `.scaf` file of
    config 1 are byte-identical to the reference's goldens."""
    monkeypatch.setenv("LD_PRELOAD", cusim_build.build())
    TD.test_config1_goldens_through_the_binding(tmp_path, tokeniser)


@needs_bin
@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["tiny1", "tiny3_exponent", "c2_mirror"])
def test_binding_equals_reference_binary_on_text_inputs(case, tmp_path, synth, monkeypatch):
    monkeypatch.setenv("LD_PRELOAD", cusim_build.build())
    TD.test_binding_equals_reference_binary_on_text_inputs(case, tmp_path, synth)
