"""Parity of the rank-partitioned pipeline (one graph over N GPUs) against the
oracle on the same global input.  Run under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py [case ...]

Every rank generates the same global input, keeps its chunk of lines, runs
gtsb_pipeline, and rank 0 merges the ranks' edges by eid and compares vertex
states, every edge attribute, edge states and adjacency order bit for bit."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

PARAMS = dict(copy_num_cutoff=0.3, astat_cutoff=20.0, use_copy_num=True, pcutoff=0.01, cncutoff=1.5, ocutoff=400)

def _one_record(S):
    """One link listed on one line only: all but one rank hold no records at all."""
    z = S.tiny_dense(5, 1, 3)
    return S.ScaffoldInput(z.seq_len, z.astat, z.copy_num, z.root[:1], z.ctg[:1], z.dist[:1], z.std_dev[:1],
                           z.num_pairs[:1], z.flags[:1])


def _one_line(S):
    """One line with two links: rank 0 holds every record, the other ranks only receive mail
    (the neighbours have no line of their own: their rows live on the last rank)."""
    import numpy as np
    z = S.tiny_dense(5, 1, 3)
    u32 = lambda *a: np.array(a, np.uint32)
    return S.ScaffoldInput(z.seq_len, z.astat, z.copy_num, u32(2, 2), u32(0, 4), np.array([100, 120], np.int32),
                           np.array([5.0, 7.5], np.float32), u32(10, 12), np.array([3, 1], np.uint8))


CASES = {
    "one_record": _one_record,
    "one_line": _one_line,
    "tiny": lambda S: S.generate("c2_bacterial", V=40, seed=5, mean_pairs=3.0),
    "small_shuffled": lambda S: S.generate("c2_bacterial", V=3000, seed=11),
    "small_id_order": lambda S: S.generate("c2_bacterial", V=3000, seed=12, line_order="id"),
    "mirror": lambda S: S.generate("c2_bacterial", V=5000, seed=13, mirror_diff_frac=0.3, dup_same_line_frac=0.2,
                                   one_sided_frac=0.2, one_sided_up=True),
    "c2": lambda S: S.generate("c2_bacterial"),
    "c3_400k": lambda S: S.generate("c3_human", V=400_000),
    # too large for the oracle: checked against the single-device run of the same input
    "c3_2m_vs_single": lambda S: S.generate("c3_human", V=2_000_000),
}


def merged_result(parts, vstate):
    e = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    E = e["eid"].shape[0]
    order = np.argsort(e["eid"], kind="stable")
    assert np.array_equal(e["eid"][order], np.arange(E, dtype=np.uint32)), "eids of the ranks are not a permutation"
    src = e["src"][order]
    adj = np.lexsort((np.arange(E), src))          # by (src, eid): adjacency order = creation order
    V = vstate.shape[0]
    row_ptr = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=V))]).astype(np.uint64)
    return dict(vstate=vstate, row_ptr=row_ptr, adj_eid=adj.astype(np.uint32), src=src, dst=e["dst"][order],
                dist=e["dist"][order].astype(np.int64), std_dev=e["std_dev"][order],
                flags=(e["flags"][order] & 3).astype(np.uint8), estate=e["estate"][order])


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("gt-scaffold_b200")
    import oracle_lib as O
    cases = sys.argv[1:] or ["one_record", "one_line", "tiny", "small_shuffled", "small_id_order", "mirror", "c2"]
    ok = True
    for name in cases:
        inp = CASES[name](pkg.synth)
        mine = pkg.api.shard_lines(inp, world, rank)
        uid = [pkg.api.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g = pkg.ScaffoldGraphB200(device=local)
        g.dist_init(rank, world, uid[0])
        k = list(CASES).index(name)
        if k % 3 != 2:
            g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
        # both input shapes: a root per record, or (root, first record) per line; every other case also
        # uploads only this rank's slice of the vertex attributes (the pipeline gathers the rest)
        if k % 2 == 1 and len(mine.root):
            g.set_record_lines(*pkg.api.lines_of(mine.root), mine.ctg, mine.dist, mine.std_dev, mine.flags)
        else:
            g.set_records(mine.root, mine.ctg, mine.dist, mine.std_dev, mine.flags)
        if k % 3 == 2:
            V = inp.nof_vertices
            lo, hi = V * rank // world, V * (rank + 1) // world
            g._ck(g.L.gtsb_set_vertices_slice_host(g.h, V, lo, hi - lo, pkg.api._ptr(np.ascontiguousarray(inp.seq_len[lo:hi], np.uint32)),
                                                   pkg.api._ptr(np.ascontiguousarray(inp.astat[lo:hi], np.float32)),
                                                   pkg.api._ptr(np.ascontiguousarray(inp.copy_num[lo:hi], np.float32))))
            g.synchronize()                       # the slices are temporaries
            g.V = V
        for rep in range(2):                      # twice: buffers are reused between calls
            g.pipeline(**PARAMS)
        part, vstate, st = g.edges(), g.vstate(), g.stats()
        parts = [None] * world
        dist.all_gather_object(parts, part)
        vs = [None] * world
        dist.all_gather_object(vs, vstate)
        if rank == 0:
            for r in range(1, world):
                assert np.array_equal(vs[0], vs[r]), f"{name}: vstate differs between ranks 0 and {r}"
            got = merged_result(parts, vstate)
            if name.endswith("_vs_single"):
                one = pkg.ScaffoldGraphB200(device=local)
                one.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
                one.set_records(inp.root, inp.ctg, inp.dist, inp.std_dev, inp.flags)
                one.pipeline(**PARAMS)
                exp = one.result()
                one.close()
            else:
                ref = O.best_oracle().build(inp)
                ref.mark_repeats(PARAMS["copy_num_cutoff"], PARAMS["astat_cutoff"], use_copy_num=True)
                ref.filter(PARAMS["pcutoff"], PARAMS["cncutoff"], PARAMS["ocutoff"])
                exp = ref.result()
            bad = [k for k in ("src", "dst", "dist", "std_dev", "flags", "row_ptr", "adj_eid", "vstate", "estate")
                   if not np.array_equal(got[k], exp[k])]
            print(f"[dist_check] {name}: world={world} V={inp.nof_vertices} E={len(got['src'])} "
                  f"edges/rank={[len(p['eid']) for p in parts]} sweeps={st['poly_sweeps']} rounds={st['fire_rounds']} "
                  f"-> {'OK' if not bad else 'MISMATCH in ' + ','.join(bad)}", flush=True)
            ok &= not bad
        g.close()
        dist.barrier()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
