"""Device-resident step timing + per-kernel table on one synthetic config (dev tool).
    python tools/probe.py [workload] [V] [steps] [key=value generator options]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("gt-scaffold_b200")

if __name__ == "__main__":
    import torch
    name = sys.argv[1] if len(sys.argv) > 1 else "c3_human"
    V = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else None
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    kw = {}
    for a in sys.argv[4:]:
        k, v = a.split("=")
        kw[k] = int(v) if v.lstrip("-").isdigit() else v
    t = pkg.synth.generate_torch(name, V=V, device="cuda", **kw)
    Vn, Rn = int(t["seq_len"].shape[0]), int(t["root"].shape[0])
    stream = torch.cuda.current_stream()
    g = pkg.ScaffoldGraphB200(device=0, stream=stream.cuda_stream)
    g.set_vertices_device(Vn, t["seq_len"].data_ptr(), t["astat"].data_ptr(), t["copy_num"].data_ptr())
    lines = kw.pop("lines", 1) if False else int(os.environ.get("PROBE_LINES", "1"))
    if lines:
        # what a .de tokeniser has: one (root, first record) per line instead of a root per record
        root = t["root"]
        brk = torch.nonzero(root[1:] != root[:-1]).flatten() + 1
        line_start = torch.cat([torch.zeros(1, dtype=brk.dtype, device=brk.device), brk,
                                torch.tensor([Rn], dtype=brk.dtype, device=brk.device)]).to(torch.int32)
        line_root = root[line_start[:-1].long()].contiguous()
        g.set_record_lines_device(int(line_root.shape[0]), line_root.data_ptr(), line_start.data_ptr(), Rn,
                                  t["ctg"].data_ptr(), t["dist"].data_ptr(), t["std_dev"].data_ptr(),
                                  t["flags"].data_ptr())
    else:
        g.set_records_device(Rn, t["root"].data_ptr(), t["ctg"].data_ptr(), t["dist"].data_ptr(),
                             t["std_dev"].data_ptr(), t["flags"].data_ptr())
    for _ in range(3):
        g.pipeline()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        g.pipeline()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = g.stats()
    g.set_profile(True)
    for _ in range(2):
        g.pipeline()
    kern = {n: round(m / 2, 4) for n, (m, c) in sorted(g.profile().items(), key=lambda kv: -kv[1][0])}
    g.set_profile(False)
    E = st["nof_edges"]
    _, de, dv = g.digest()          # order-independent digests of every edge and vertex state (gtsb_result_digest)
    print(json.dumps({"workload": name, "V": Vn, "R": Rn, "E": E, "ms_per_step": round(ms, 4),
                      "digest_edges": "%016x" % de, "digest_vertices": "%016x" % dv,
                      "edges_per_s": E / ms * 1e3, "env": {k: v for k, v in os.environ.items() if k.startswith("GTSB_")},
                      "stats": {k: st[k] for k in ("max_degree", "big_rows", "proposals", "poly_sweeps", "fire_rounds",
                                                   "line_ordered_build", "fallback_reason", "kernel_launches")},
                      "kernels_ms": kern, "sum_kernels_ms": round(sum(kern.values()), 4)}))
