"""Stage timings of the CUDA path on one synthetic config (dev tool)."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("gt-scaffold_b200")

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c3_human"
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    t = time.time()
    inp = pkg.synth.generate(name, V=V)
    print("generated", inp.meta, f"{time.time()-t:.1f}s", flush=True)
    g = pkg.ScaffoldGraphB200()
    g.set_vertices(inp.seq_len, inp.astat, inp.copy_num)
    g.set_records(inp.root, inp.ctg, inp.dist, inp.std_dev, inp.flags)
    for r in range(reps):
        t = time.time()
        g.build()
        g.mark_repeats()
        g.filter()
        g.synchronize()
        st = g.stats()
        print(f"rep {r}: wall {1e3*(time.time()-t):.2f} ms  build {st['ms_build']:.3f}  "
              f"repeats {st['ms_mark_repeats']:.3f}  filter {st['ms_filter']:.3f}  "
              f"E={st['nof_edges']} sweeps={st['poly_sweeps']} rounds={st['fire_rounds']} "
              f"launches={st['kernel_launches']}", flush=True)
