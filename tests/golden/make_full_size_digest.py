"""Golden digests of the NAMED configs at full size, made by the compiled, unmodified reference
(oracle/_ref, built from /root/reference) in the build container:

    python tests/golden/make_full_size_digest.py c4_repeat_hubs      # ~13 min, ~7 GB on one core
    python tests/golden/make_full_size_digest.py c3_human            # ~2 min

writes tests/golden/full_size_<config>.json: sha256 of every result array (edges in
graph->edges[] order, adjacency, states) after build + mark_repeats + filter with the
reference driver's constants (test.c:35-42), plus counts and the reference's stage times.
The inputs come from synth.generate (numpy PCG64, seeded): the GPU test regenerates them
on the box, so nothing under /root/reference is needed there."""
import hashlib
import importlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

KEYS = ("src", "dst", "dist", "std_dev", "flags", "row_ptr", "adj_eid", "vstate", "estate")
PARAMS = dict(copy_num_cutoff=0.3, astat_cutoff=20.0, pcutoff=0.01, cncutoff=1.5, ocutoff=400)
KW = {"c4_repeat_hubs": dict(max_deg=10_000), "c3_human": {}, "c2_bacterial": {}}


def digest(res):
    out = {}
    for k in KEYS:
        a = np.ascontiguousarray(res[k])
        out[k] = {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "dtype": str(a.dtype), "n": int(a.shape[0])}
    return out


def canonical(res):
    """One dtype per key, whoever produced the arrays."""
    dt = dict(src=np.uint32, dst=np.uint32, dist=np.int64, std_dev=np.float32, flags=np.uint8,
              row_ptr=np.uint64, adj_eid=np.uint32, vstate=np.uint8, estate=np.uint8)
    return {k: np.ascontiguousarray(res[k]).astype(dt[k], copy=False) for k in KEYS}


if __name__ == "__main__":
    import oracle_lib as O
    pkg = importlib.import_module("gt-scaffold_b200")
    name = sys.argv[1]
    V = int(sys.argv[2]) if len(sys.argv) > 2 else None
    assert O.have_ref(), "needs oracle/_ref (the compiled reference)"
    t0 = time.time()
    inp = pkg.synth.generate(name, V=V, **KW[name])
    t_gen = time.time() - t0
    g = O.RefGraph.build(inp)
    t_build = g.build_seconds
    t_rep = g.mark_repeats(PARAMS["copy_num_cutoff"], PARAMS["astat_cutoff"], use_copy_num=True)
    t_fil = g.filter(PARAMS["pcutoff"], PARAMS["cncutoff"], PARAMS["ocutoff"])
    res = canonical(g.result())
    deg = np.diff(res["row_ptr"].astype(np.int64))
    out = {"config": name, "generator": "synth.generate (numpy %s)" % np.__version__, "meta": inp.meta,
           "kwargs": KW[name], "params": PARAMS, "V": inp.nof_vertices, "R": inp.nof_records, "E": int(g.E),
           "max_degree": int(deg.max()), "rows_over_32": int((deg > 32).sum()),
           "states": {"vstate": np.bincount(res["vstate"], minlength=8).tolist(),
                      "estate": np.bincount(res["estate"], minlength=8).tolist()},
           "reference_seconds": {"build": t_build, "mark_repeats": t_rep, "filter": t_fil, "host": os.uname().nodename,
                                 "cores_used": 1},
           "input_sha256": {k: hashlib.sha256(np.ascontiguousarray(getattr(inp, k)).tobytes()).hexdigest()
                            for k in ("seq_len", "astat", "copy_num", "root", "ctg", "dist", "std_dev", "flags")},
           "digest": digest(res)}
    suffix = "" if V is None else "_V%d" % V
    path = os.path.join(HERE, "full_size_%s%s.json" % (name, suffix))
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, "gen %.0fs build %.0fs rep %.0fs filter %.0fs" % (t_gen, t_build, t_rep, t_fil))
