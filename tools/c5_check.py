"""One named config on N GPUs (torchrun) or one: device-timed steps, per-kernel table, and the
order-independent result digests of gtsb_result_digest -- edges and vertices of the whole graph --
so that a partitioned run and a single-device run of a graph too large to fetch (BASELINE.json
config 5: 10^8 contigs) can be compared from their JSON lines.

    python tools/c5_check.py [--workload c5_metagenome] [--vertices V] [--steps K] [--out file.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 tools/c5_check.py --out gpurun_out/c5_n8.json
    python tools/c5_check.py --compare a.json b.json        # exit 1 unless the digests agree
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PARAMS = (0.3, 20.0, True, 0.01, 1.5, 400)       # test.c:35-42


def compare(a, b):
    x, y = json.load(open(a)), json.load(open(b))
    keys = ("workload", "vertices", "edges", "digest_edges", "digest_vertices")
    bad = [k for k in keys if x[k] != y[k]]
    print(json.dumps({"compare": [a, b], "n_gpus": [x["n_gpus"], y["n_gpus"]],
                      "ms_per_step": [x["ms_per_step"], y["ms_per_step"]],
                      "speedup": x["ms_per_step"] / y["ms_per_step"] if y["ms_per_step"] else None,
                      "result": "EQUAL" if not bad else "MISMATCH in " + ",".join(bad),
                      **{k: [x[k], y[k]] for k in keys}}))
    return 0 if not bad else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5_metagenome")
    ap.add_argument("--vertices", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--line-order", default="shuffled")
    ap.add_argument("--out", default=None)
    ap.add_argument("--compare", nargs=2, default=None)
    args = ap.parse_args()
    if args.compare:
        sys.exit(compare(*args.compare))

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("gt-scaffold_b200")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    full = pkg.synth.generate_torch(args.workload, V=args.vertices, device=dev, line_order=args.line_order)
    Vn, Rg = int(full["seq_len"].shape[0]), int(full["root"].shape[0])
    if world > 1:
        fr = pkg.api.chunk_fractions(world)
        cuts = [0]
        for r in range(1, world):
            i = max(1, int(Rg * fr[r]))
            w = full["root"][i - 1:i + 65536].cpu()
            brk = (w[1:] != w[:-1]).nonzero().flatten()
            cuts.append(i + int(brk[0]) if len(brk) else Rg)
        cuts.append(Rg)
        lo, hi = cuts[rank], cuts[rank + 1]
        t = {k: full[k] for k in ("seq_len", "astat", "copy_num")}
        for k in ("root", "ctg", "dist", "std_dev", "flags"):
            t[k] = full[k][lo:hi].clone()
        del full
        torch.cuda.empty_cache()
    else:
        t = full
        torch.cuda.empty_cache()             # the generator's temporaries go back to the driver
    Rn = int(t["root"].shape[0])
    stream = torch.cuda.current_stream(dev)
    g = pkg.ScaffoldGraphB200(device=local, stream=stream.cuda_stream)
    if world > 1:
        box = [pkg.api.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        g.dist_init(rank, world, box[0])
    g.set_vertices_device(Vn, t["seq_len"].data_ptr(), t["astat"].data_ptr(), t["copy_num"].data_ptr())
    # records in the shape a .de tokeniser leaves them in: (root, first record) per line
    root = t["root"]
    brk = torch.nonzero(root[1:] != root[:-1]).flatten() + 1
    line_start = torch.cat([torch.zeros(1, dtype=brk.dtype, device=dev), brk,
                            torch.tensor([Rn], dtype=brk.dtype, device=dev)]).to(torch.int32).contiguous()
    line_root = root[line_start[:-1].long()].contiguous()
    del brk
    g.set_record_lines_device(int(line_root.shape[0]), line_root.data_ptr(), line_start.data_ptr(), Rn,
                              t["ctg"].data_ptr(), t["dist"].data_ptr(), t["std_dev"].data_ptr(), t["flags"].data_ptr())

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        g.pipeline(*PARAMS)
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        g.pipeline(*PARAMS)
        evs[k + 1].record(stream)
    barrier()
    per_step = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    # the mean is what bench.py reports; the per-step times show a one-off stall if there was one
    ms = torch.tensor([sum(per_step) / max(1, args.steps), sorted(per_step)[len(per_step) // 2]] + per_step,
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    g.set_profile(True)
    g.pipeline(*PARAMS)
    kern = {n: round(m, 4) for n, (m, c) in sorted(g.profile().items(), key=lambda kv: -kv[1][0])}
    g.set_profile(False)
    E, de, dv = g.digest()
    st = g.stats()
    parts = [(E, de, dv, Rn, torch.cuda.max_memory_allocated(dev))]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (E, de, dv, Rn, torch.cuda.max_memory_allocated(dev)))
    if rank == 0:
        mask = (1 << 64) - 1
        assert all(p[2] == parts[0][2] for p in parts), "ranks disagree on the vertex states"
        line = {"workload": args.workload, "vertices": Vn, "records": sum(p[3] for p in parts),
                "edges": sum(p[0] for p in parts), "n_gpus": world, "steps": args.steps,
                "ms_per_step": float(ms[1]), "ms_per_step_mean": float(ms[0]), "ms_each_step": [round(float(x), 3) for x in ms[2:]],
                "timing": "CUDA events around every step, max over ranks; ms_per_step = median step",
                "edges_per_s": sum(p[0] for p in parts) / float(ms[1]) * 1e3,
                "digest_edges": "%016x" % (sum(p[1] for p in parts) & mask), "digest_vertices": "%016x" % parts[0][2],
                "edges_per_rank": [p[0] for p in parts], "records_per_rank": [p[3] for p in parts],
                "stats": {k: st[k] for k in ("proposals", "poly_sweeps", "fire_rounds", "line_ordered_build",
                                             "fallback_reason", "big_rows", "max_degree")},
                "kernels_ms_rank0": kern}
        text = json.dumps(line)
        print(text, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(text + "\n")
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
