"""Regenerate the committed golden fixtures (run in the build container, where
/root/reference exists and oracle/_ref has been built by `make -C oracle`).

c1/            the reference's own golden test:
               libPE.de, libPE.astat and the four expected stage .dot files are
               the reference's testdata, byte for byte; contigs.fa keeps every
               header line and sequence LENGTH of testdata/primary-contigs.fa
               but replaces the bases by 'N' (only header and length reach the
               graph: parser.c:420-494); c1_expected.scaf is what the compiled
               reference writes for these inputs (the reference ships no .scaf
               golden).
diff_*.npz     differential vectors: adversarial small graphs
               (synth.tiny_dense) pushed through the COMPILED REFERENCE
               (build -> mark_repeats -> filter); inputs and every output array.
               They pin the pairwise filter logic, which the reference's own
               goldens never exercise.  diff_g..i carry
               NaN / inf / denormal std_dev, NaN copy numbers and a-statistics
               and distances at the int32 limits (synth.special_values).
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

synth = importlib.import_module("gt-scaffold_b200.synth")
REF_TESTDATA = "/root/reference/testdata"
STAGES = ["mark_repeats", "filter", "removecycles", "makescaffold"]

DIFF_CASES = [
    # (name, V, pairs, seed, pcutoff, cncutoff, ocutoff, cn_cut, astat_cut, use_cn)
    ("diff_a", 12, 40, 101, 0.01, 1.5, 400, 0.3, 20.0, 1),
    ("diff_b", 30, 160, 102, 0.01, 1.5, 400, 0.3, 20.0, 1),
    ("diff_c", 9, 30, 103, 0.01, 1.5, 0, 0.3, 20.0, 1),
    ("diff_d", 14, 60, 104, 0.01, 1.5, -1, 0.3, 20.0, 0),
    ("diff_e", 40, 300, 105, 0.2, 2.5, 50, 0.5, 19.5, 1),
    ("diff_f", 200, 900, 106, 0.01, 1.5, 400, 0.3, 20.0, 1),
]
# the same through synth.special_values (NaN / inf / denormal std_dev, extreme distances)
SPECIAL_CASES = [
    ("diff_g", 14, 70, 0, 0.01, 1.5, 400, 0.3, 20.0, 1),
    ("diff_h", 24, 150, 1, 0.05, 2.0, 100, 0.5, 19.5, 1),
    ("diff_i", 16, 90, 2, 0.01, 1.5, -5, 0.3, 20.0, 0),
]


def make_c1():
    out = os.path.join(HERE, "c1")
    os.makedirs(out, exist_ok=True)
    for f in ["libPE.de", "libPE.astat", "wrong_libPE_1.de", "wrong_libPE_2.de",
              "gt_scaffolder_graph_test_expected.dot"] + \
            ["gt_scaffolder_algorithms_test_%s_expected.dot" % s for s in STAGES]:
        shutil.copyfile(os.path.join(REF_TESTDATA, f), os.path.join(out, f))
        os.chmod(os.path.join(out, f), 0o644)
    with open(os.path.join(REF_TESTDATA, "primary-contigs.fa")) as src, \
            open(os.path.join(out, "contigs.fa"), "w") as dst:
        n = 0
        for line in src:
            if line.startswith(">"):
                if n:
                    dst.write("N" * n + "\n")
                n = 0
                dst.write(line)
            else:
                n += len(line.strip())
        if n:
            dst.write("N" * n + "\n")
    # the reduced FASTA must reproduce the reference's goldens
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run([O.REF_TESTX, "scaffold", os.path.join(out, "contigs.fa"),
                        os.path.join(out, "libPE.de"), os.path.join(out, "libPE.astat"), "false"],
                       cwd=tmp, check=True, stderr=subprocess.DEVNULL)
        for s in STAGES:
            a = open(os.path.join(tmp, "gt_scaffolder_algorithms_test_%s.dot" % s)).read()
            b = open(os.path.join(out, "gt_scaffolder_algorithms_test_%s_expected.dot" % s)).read()
            assert a == b, s
        shutil.copyfile(os.path.join(tmp, "gt_scaffolder_new_write.scaf"),
                        os.path.join(out, "c1_expected.scaf"))
    print("c1 fixtures written and verified against the compiled reference")


def make_diff():
    cases = [(c, synth.tiny_dense(c[1], c[2], c[3])) for c in DIFF_CASES] + \
            [(c, synth.special_values(c[3], V=c[1], n_pairs=c[2])) for c in SPECIAL_CASES]
    for (name, V, pairs, seed, pc, cnc, oc, cn_cut, a_cut, use_cn), inp in cases:
        g = O.RefGraph.build(inp)
        built = g.result()
        g.mark_repeats(cn_cut, a_cut, use_copy_num=bool(use_cn))
        rep_v, rep_e = g.vstate(), g.estate()
        g.filter(pc, cnc, oc)
        fin = g.result()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            seq_len=inp.seq_len, astat=inp.astat, copy_num=inp.copy_num, root=inp.root,
            ctg=inp.ctg, dist=inp.dist, std_dev=inp.std_dev, num_pairs=inp.num_pairs,
            flags=inp.flags,
            params=np.array([pc, cnc, oc, cn_cut, a_cut, use_cn], np.float64),
            e_src=built["src"], e_dst=built["dst"], e_dist=built["dist"],
            e_std=built["std_dev"], e_np=built["num_pairs"], e_flags=built["flags"],
            row_ptr=built["row_ptr"], adj_eid=built["adj_eid"],
            rep_vstate=rep_v, rep_estate=rep_e,
            fin_vstate=fin["vstate"], fin_estate=fin["estate"])
        g.close()
        print(name, "V", V, "E", len(built["src"]),
              "poly", int((fin["vstate"] == 1).sum()),
              "incons", int((fin["estate"] == 2).sum()))


if __name__ == "__main__":
    O.build_oracles()
    make_c1()
    make_diff()
