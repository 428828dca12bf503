"""A failure on ONE rank of the partitioned pipeline must stop ALL ranks at the same exchange
(no rank may be left waiting in a collective).  Run under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29519 tests/dist_fail_check.py

GTSB_FAIL_AT="<rank>:<place>" (a test hook of gtsb_dist.cu) makes that rank report an allocation
failure at one of the places where buffers grow; every rank must come back from gtsb_pipeline
with an error, and the next call (hook off) must succeed again on the same contexts."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

PLACES = ["setup", "facts", "receive", "corrections", "windows", "filter", "proposals", "fire"]


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("gt-scaffold_b200")
    inp = pkg.synth.generate("c2_bacterial", V=5000, seed=13, mirror_diff_frac=0.3, dup_same_line_frac=0.2)
    mine = pkg.api.shard_lines(inp, world, rank)
    ok = True
    for place in PLACES:
        for bad_rank in sorted({0, world - 1}):
            # a fresh context per case: every buffer has to grow, so every place is live
            uid = [pkg.api.dist_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            g = pkg.ScaffoldGraphB200(device=local)
            g.dist_init(rank, world, uid[0])
            g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
            g.set_records(mine.root, mine.ctg, mine.dist, mine.std_dev, mine.flags)
            os.environ["GTSB_FAIL_AT"] = f"{bad_rank}:{place}"
            try:
                g.pipeline()
                outcome = "returned ok"
            except RuntimeError as e:
                outcome = "error: " + str(e)[:80]
            del os.environ["GTSB_FAIL_AT"]
            outs = [None] * world
            dist.all_gather_object(outs, outcome)
            # and the same contexts recover
            try:
                g.pipeline()
                again = "ok"
            except RuntimeError as e:
                again = "error: " + str(e)[:80]
            agains = [None] * world
            dist.all_gather_object(agains, again)
            if rank == 0:
                all_failed = all(o.startswith("error") for o in outs)
                # "corrections" is only a growth step when the input has reverse-flag corrections
                # (rare); when the place is not reached every rank must return ok together
                not_reached = place == "corrections" and all(o == "returned ok" for o in outs)
                good = (all_failed or not_reached) and all(a == "ok" for a in agains)
                print(f"[dist_fail] place={place} failing rank={bad_rank}: every rank returned an error: {all_failed}; "
                      f"next call: {agains} -> {'OK' if good else 'BAD ' + str(outs)}", flush=True)
                ok &= good
            g.close()
            dist.barrier()
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
