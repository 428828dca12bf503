"""First-contact debugging on the GPU box: run a few cases, print where the
CUDA path and the oracle diverge instead of just failing."""
import importlib
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("gt-scaffold_b200")
import oracle_lib as O  # noqa: E402

KEYS = ("src", "dst", "dist", "std_dev", "flags", "row_ptr", "adj_eid", "vstate", "estate")


def diff(got, exp, tag):
    ok = True
    for k in KEYS:
        if got[k].shape != exp[k].shape:
            print(f"  [{tag}] {k}: shape {got[k].shape} vs {exp[k].shape}")
            ok = False
        elif not np.array_equal(got[k], exp[k]):
            bad = np.nonzero(got[k] != exp[k])[0]
            print(f"  [{tag}] {k}: {len(bad)} differ, first {bad[:6]} got {got[k][bad[:6]]} exp {exp[k][bad[:6]]}")
            ok = False
    return ok


def case(inp, tag, params=(0.3, 20.0, True, 0.01, 1.5, 400), force_general=False):
    cn_cut, a_cut, use_cn, pc, cnc, oc = params
    t0 = time.time()
    try:
        g = pkg.ScaffoldGraphB200.new_from_records(inp, force_general=force_general)
        ref = O.best_oracle().build(inp)
        ok = diff(g.result(), ref.result(), tag + "/build")
        g.mark_repeats(cn_cut, a_cut, use_cn)
        ref.mark_repeats(cn_cut, a_cut, use_copy_num=use_cn)
        ok &= diff(g.result(), ref.result(), tag + "/repeats")
        g.filter(pc, cnc, oc)
        ref.filter(pc, cnc, oc)
        ok &= diff(g.result(), ref.result(), tag + "/filter")
        print(f"{tag}: {'OK' if ok else 'MISMATCH'} V={inp.nof_vertices} R={inp.nof_records} "
              f"{time.time()-t0:.2f}s stats={g.stats()}")
        g.close()
    except Exception:
        print(f"{tag}: EXCEPTION")
        traceback.print_exc()


if __name__ == "__main__":
    print("oracle:", O.best_oracle().__name__)
    for seed in range(6):
        case(pkg.synth.generate("c2_bacterial", V=6 + 5 * seed, seed=900 + seed, mean_pairs=1.0 + seed % 4,
                                line_order="id" if seed % 2 else "shuffled", one_sided_frac=0.25 * (seed % 2),
                                one_sided_up=True, mirror_diff_frac=0.3, dup_same_line_frac=0.2), f"line{seed}")
    for seed in range(3):
        case(pkg.synth.tiny_dense(4 + seed, 6 + 5 * seed, 3000 + seed), f"tiny{seed}")
    case(pkg.synth.generate("c2_bacterial", V=2000), "c2_2k")
    case(pkg.synth.generate("c2_bacterial", V=2000, line_order="id"), "c2_2k_id")
    case(pkg.synth.generate("c2_bacterial"), "c2_full")
    case(pkg.synth.generate("c4_repeat_hubs", V=50000, max_deg=2000), "c4_50k")
    case(pkg.synth.generate("c3_human", V=1_000_000), "c3_1M")
