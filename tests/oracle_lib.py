"""ctypes access to the two CPU oracles (TEST INFRASTRUCTURE).

* RefGraph  -- the compiled, unmodified reference (oracle/_ref/libgtscaf_ref.so,
               built by oracle/Makefile from /root/reference/src).
* PortGraph -- the array-level C restatement (oracle/liboracle_port.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgtscaf_ref.so")
REF_TESTX = os.path.join(ORACLE_DIR, "_ref", "test.x")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle_port.so")

_p = lambda a, t: a.ctypes.data_as(C.POINTER(t))


def build_oracles(quiet: bool = True) -> None:
    """make -C oracle (port always; _ref only where /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _load(path):
    if not os.path.exists(path):
        build_oracles()
    return C.CDLL(path)


_ref = None
_port = None


def ref_lib():
    global _ref
    if _ref is None:
        L = _load(REF_SO)
        L.refdrv_build.restype = C.c_void_p
        L.refdrv_build.argtypes = [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint64] + \
            [C.c_void_p] * 6 + [C.c_int, C.POINTER(C.c_double)]
        L.refdrv_nof_edges.restype = C.c_uint64
        L.refdrv_nof_edges.argtypes = [C.c_void_p]
        L.refdrv_nof_vertices.restype = C.c_uint64
        L.refdrv_nof_vertices.argtypes = [C.c_void_p]
        L.refdrv_get_vertices.argtypes = [C.c_void_p] * 5
        L.refdrv_get_edges.argtypes = [C.c_void_p] * 8
        L.refdrv_get_adjacency.argtypes = [C.c_void_p] * 3
        L.refdrv_set_states.argtypes = [C.c_void_p] * 3
        L.refdrv_mark_repeats.restype = C.c_int
        L.refdrv_mark_repeats.argtypes = [C.c_void_p, C.c_char_p, C.c_float, C.c_float,
                                          C.POINTER(C.c_double)]
        L.refdrv_filter.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int64,
                                    C.POINTER(C.c_double)]
        L.refdrv_new_from_file.restype = C.c_void_p
        L.refdrv_new_from_file.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.refdrv_print.argtypes = [C.c_void_p, C.c_char_p]
        if hasattr(L, "refdrv_print_scaffold"):
            L.refdrv_print_scaffold.argtypes = [C.c_void_p, C.c_char_p]
        L.refdrv_removecycles.argtypes = [C.c_void_p]
        L.refdrv_makescaffold.argtypes = [C.c_void_p]
        L.refdrv_write_scaffold.argtypes = [C.c_void_p, C.c_char_p]
        L.refdrv_delete.argtypes = [C.c_void_p]
        _ref = L
    return _ref


class _OraGraph(C.Structure):
    _fields_ = [("nof_vertices", C.c_uint64), ("seq_len", C.c_void_p),
                ("astat", C.c_void_p), ("copy_num", C.c_void_p),
                ("vstate", C.POINTER(C.c_uint8)), ("nof_vedges", C.c_void_p),
                ("vedges", C.c_void_p), ("nof_edges", C.c_uint64),
                ("max_nof_edges", C.c_uint64), ("src", C.POINTER(C.c_uint32)),
                ("dst", C.POINTER(C.c_uint32)), ("dist", C.POINTER(C.c_int64)),
                ("std_dev", C.POINTER(C.c_float)), ("num_pairs", C.POINTER(C.c_uint64)),
                ("flags", C.POINTER(C.c_uint8)), ("estate", C.POINTER(C.c_uint8)),
                ("win_rec", C.POINTER(C.c_int64))]


def port_lib():
    global _port
    if _port is None:
        L = _load(PORT_SO)
        L.ora_build.restype = C.POINTER(_OraGraph)
        L.ora_build.argtypes = [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint64] + [C.c_void_p] * 6
        L.ora_mark_repeats.argtypes = [C.POINTER(_OraGraph), C.c_int, C.c_float, C.c_float]
        L.ora_filter.argtypes = [C.POINTER(_OraGraph), C.c_float, C.c_float, C.c_int64]
        L.ora_ambiguousorders.restype = None
        L.ora_ambiguousorders.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_float, C.c_void_p]
        L.ora_ambiguous_intervals.restype = None
        L.ora_ambiguous_intervals.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
        L.ora_ambiguousorder.restype = C.c_int
        L.ora_ambiguousorder.argtypes = [C.c_int64, C.c_float, C.c_int64, C.c_float, C.c_float]
        L.ora_overlap.restype = C.c_int64
        L.ora_overlap.argtypes = [C.c_int64, C.c_uint64, C.c_int64, C.c_uint64]
        L.ora_get_adjacency.argtypes = [C.POINTER(_OraGraph), C.c_void_p, C.c_void_p]
        L.ora_delete.argtypes = [C.POINTER(_OraGraph)]
        _port = L
    return _port


def _inputs(inp):
    """ScaffoldInput -> the widened arrays both oracles take."""
    return dict(
        seq_len=np.ascontiguousarray(inp.seq_len, np.uint64),
        astat=np.ascontiguousarray(inp.astat, np.float32),
        copy_num=np.ascontiguousarray(inp.copy_num, np.float32),
        root=np.ascontiguousarray(inp.root, np.uint32),
        ctg=np.ascontiguousarray(inp.ctg, np.uint32),
        dist=np.ascontiguousarray(inp.dist, np.int64),
        std_dev=np.ascontiguousarray(inp.std_dev, np.float32),
        num_pairs=np.ascontiguousarray(inp.num_pairs, np.uint64),
        flags=np.ascontiguousarray(inp.flags, np.uint8))


class _Common:
    """Shared result accessors: edges() / adjacency() / vstate() / estate()."""

    def result(self):
        e = self.edges()
        row_ptr, eids = self.adjacency()
        return dict(vstate=self.vstate(), row_ptr=row_ptr, adj_eid=eids, **e)


class RefGraph(_Common):
    """A GtScaffolderGraph owned by the compiled reference."""

    def __init__(self, handle, seconds=0.0):
        self.L = ref_lib()
        self.h = handle
        self.build_seconds = seconds
        self.V = int(self.L.refdrv_nof_vertices(self.h))

    @classmethod
    def build(cls, inp, with_headers=False):
        L = ref_lib()
        a = _inputs(inp)
        sec = C.c_double(0)
        h = L.refdrv_build(a["seq_len"].shape[0], a["seq_len"].ctypes.data, a["astat"].ctypes.data,
                           a["copy_num"].ctypes.data, a["root"].shape[0], a["root"].ctypes.data,
                           a["ctg"].ctypes.data, a["dist"].ctypes.data, a["std_dev"].ctypes.data,
                           a["num_pairs"].ctypes.data, a["flags"].ctypes.data,
                           int(with_headers), C.byref(sec))
        return cls(h, sec.value)

    @classmethod
    def from_files(cls, fasta, de, min_ctg_len=200):
        h = ref_lib().refdrv_new_from_file(fasta.encode(), min_ctg_len, de.encode())
        if not h:
            raise RuntimeError("reference gt_scaffolder_graph_new_from_file failed")
        return cls(h)

    @property
    def E(self):
        return int(self.L.refdrv_nof_edges(self.h))

    def mark_repeats(self, copy_num_cutoff, astat_cutoff, use_copy_num=True, astat_file=None):
        sec = C.c_double(0)
        if astat_file is None and use_copy_num:
            # an EMPTY file switches the copy-number clause on without
            # touching the in-memory values (algorithms.c:108-164)
            with tempfile.NamedTemporaryFile(suffix=".astat") as f:
                rc = self.L.refdrv_mark_repeats(self.h, f.name.encode(), copy_num_cutoff,
                                                astat_cutoff, C.byref(sec))
        else:
            name = (astat_file or "").encode()
            rc = self.L.refdrv_mark_repeats(self.h, name, copy_num_cutoff, astat_cutoff,
                                            C.byref(sec))
        if rc != 0:
            raise RuntimeError("reference mark_repeats failed")
        return sec.value

    def filter(self, pcutoff, cncutoff, ocutoff):
        sec = C.c_double(0)
        self.L.refdrv_filter(self.h, pcutoff, cncutoff, int(ocutoff), C.byref(sec))
        return sec.value

    def set_states(self, vstate=None, estate=None):
        v = None if vstate is None else np.ascontiguousarray(vstate, np.uint8)
        e = None if estate is None else np.ascontiguousarray(estate, np.uint8)
        self.L.refdrv_set_states(self.h, None if v is None else v.ctypes.data,
                                 None if e is None else e.ctypes.data)

    def vertices(self):
        V = self.V
        out = dict(seq_len=np.zeros(V, np.uint64), astat=np.zeros(V, np.float32),
                   copy_num=np.zeros(V, np.float32), state=np.zeros(V, np.uint8))
        self.L.refdrv_get_vertices(self.h, out["seq_len"].ctypes.data, out["astat"].ctypes.data,
                                   out["copy_num"].ctypes.data, out["state"].ctypes.data)
        return out

    def vstate(self):
        return self.vertices()["state"]

    def edges(self):
        E = self.E
        o = dict(src=np.zeros(E, np.uint32), dst=np.zeros(E, np.uint32), dist=np.zeros(E, np.int64),
                 std_dev=np.zeros(E, np.float32), num_pairs=np.zeros(E, np.uint64),
                 flags=np.zeros(E, np.uint8), estate=np.zeros(E, np.uint8))
        self.L.refdrv_get_edges(self.h, o["src"].ctypes.data, o["dst"].ctypes.data,
                                o["dist"].ctypes.data, o["std_dev"].ctypes.data,
                                o["num_pairs"].ctypes.data, o["flags"].ctypes.data,
                                o["estate"].ctypes.data)
        return o

    def estate(self):
        return self.edges()["estate"]

    def adjacency(self):
        row_ptr = np.zeros(self.V + 1, np.uint64)
        eids = np.zeros(self.E, np.uint32)
        self.L.refdrv_get_adjacency(self.h, row_ptr.ctypes.data, eids.ctypes.data)
        return row_ptr, eids

    def print_dot(self, path):
        return self.L.refdrv_print(self.h, path.encode())

    def print_scaffold_dot(self, path):
        """gt_scaffolder_graph_print_scaffold (graph.c:310-343)"""
        return self.L.refdrv_print_scaffold(self.h, path.encode())

    def removecycles(self):
        self.L.refdrv_removecycles(self.h)

    def makescaffold(self):
        self.L.refdrv_makescaffold(self.h)

    def write_scaffold(self, path):
        return self.L.refdrv_write_scaffold(self.h, path.encode())

    def calc_cc(self):
        """gt_scaffolder_calc_cc_and_terminals: list of components, each the list of its terminal vertex
        ids in the order the reference's search stores them (vertex states change as the reference's do)"""
        off = np.zeros(self.V + 2, np.uint64)
        term = np.zeros(self.V + 1, np.uint32)
        self.L.refdrv_calc_cc.restype = C.c_int64
        n = self.L.refdrv_calc_cc(C.c_void_p(self.h), off.ctypes.data_as(C.c_void_p), C.c_uint64(len(off)),
                                  term.ctypes.data_as(C.c_void_p), C.c_uint64(len(term)))
        assert n >= 0
        return [term[int(off[i]):int(off[i + 1])].tolist() for i in range(int(n))]

    def close(self):
        if self.h:
            self.L.refdrv_delete(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PortGraph(_Common):
    """The C restatement (oracle/gtscaf_oracle.c)."""

    def __init__(self, inp):
        self.L = port_lib()
        a = _inputs(inp)
        self.g = self.L.ora_build(a["seq_len"].shape[0], a["seq_len"].ctypes.data,
                                  a["astat"].ctypes.data, a["copy_num"].ctypes.data,
                                  a["root"].shape[0], a["root"].ctypes.data, a["ctg"].ctypes.data,
                                  a["dist"].ctypes.data, a["std_dev"].ctypes.data,
                                  a["num_pairs"].ctypes.data, a["flags"].ctypes.data)
        self.V = int(self.g.contents.nof_vertices)

    build = classmethod(lambda cls, inp: cls(inp))

    @property
    def E(self):
        return int(self.g.contents.nof_edges)

    def mark_repeats(self, copy_num_cutoff, astat_cutoff, use_copy_num=True):
        self.L.ora_mark_repeats(self.g, int(use_copy_num), copy_num_cutoff, astat_cutoff)

    def filter(self, pcutoff, cncutoff, ocutoff):
        self.L.ora_filter(self.g, pcutoff, cncutoff, int(ocutoff))

    def set_states(self, vstate=None, estate=None):
        g = self.g.contents
        if vstate is not None:
            C.memmove(g.vstate, np.ascontiguousarray(vstate, np.uint8).ctypes.data, self.V)
        if estate is not None:
            C.memmove(g.estate, np.ascontiguousarray(estate, np.uint8).ctypes.data, self.E)

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    def vstate(self):
        return self._arr(self.g.contents.vstate, self.V, np.uint8)

    def estate(self):
        return self._arr(self.g.contents.estate, self.E, np.uint8)

    def edges(self):
        g, E = self.g.contents, self.E
        return dict(src=self._arr(g.src, E, np.uint32), dst=self._arr(g.dst, E, np.uint32),
                    dist=self._arr(g.dist, E, np.int64), std_dev=self._arr(g.std_dev, E, np.float32),
                    num_pairs=self._arr(g.num_pairs, E, np.uint64),
                    flags=self._arr(g.flags, E, np.uint8), estate=self.estate(),
                    win_rec=self._arr(g.win_rec, E, np.int64))

    def adjacency(self):
        row_ptr = np.zeros(self.V + 1, np.uint64)
        eids = np.zeros(self.E, np.uint32)
        self.L.ora_get_adjacency(self.g, row_ptr.ctypes.data, eids.ctypes.data)
        return row_ptr, eids

    def close(self):
        if self.g:
            self.L.ora_delete(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def best_oracle():
    """The compiled reference when present, else the port."""
    return RefGraph if have_ref() else PortGraph


# ---------------------------------------------------------------- text files

def write_text_inputs(inp, directory, write_seq=True):
    """ScaffoldInput -> (fasta, de, astat) files the reference's own parser can
    read.  Headers 'c%010d' sort in id order.  Lines longer than the
    reference's 1024-byte buffer are the caller's problem (parser.c:30,323).
    write_seq=False: the .de only."""
    V = inp.nof_vertices
    fa = os.path.join(directory, "contigs.fa")
    de = os.path.join(directory, "lib.de")
    astat = os.path.join(directory, "lib.astat")
    if write_seq:
        with open(fa, "w") as f:
            for v in range(V):
                f.write(">c%010d %d 0\n" % (v, inp.seq_len[v]))
                f.write("A" * int(inp.seq_len[v]) + "\n")
        with open(astat, "w") as f:
            for v in range(V):
                f.write("c%010d\t%d\t0\t0\t%.9g\t%.9g\n" % (v, inp.seq_len[v], inp.copy_num[v], inp.astat[v]))
    with open(de, "w") as f:
        i, R = 0, inp.nof_records
        while i < R:
            j = i
            r = int(inp.root[i])
            # a line = maximal run of one root whose sense flags are
            # non-increasing (sense block, then ';', then antisense block)
            while j < R and int(inp.root[j]) == r and not (
                    j > i and (inp.flags[j] & 1) and not (inp.flags[j - 1] & 1)):
                j += 1
            parts = ["c%010d" % r]
            switched = False
            for k in range(i, j):
                if not (inp.flags[k] & 1) and not switched:
                    parts.append(";")
                    switched = True
                parts.append("c%010d%s,%d,%d,%.9g" % (
                    inp.ctg[k], "+" if inp.flags[k] & 2 else "-", inp.dist[k],
                    inp.num_pairs[k], inp.std_dev[k]))
            if not switched:
                parts.append(";")
            f.write(" ".join(parts) + "\n")
            i = j
    return fa, de, astat


# ---- the reference's distance estimator (oracle/ref_bam_driver.c around gt_scaffolder_bamparser.c)

_refbam = None


def refbam_lib():
    global _refbam
    if _refbam is None:
        path = os.path.join(os.path.dirname(REF_SO), "libgtscaf_refbam.so")
        _refbam = C.CDLL(path)
    return _refbam


def have_refbam():
    return os.path.exists(os.path.join(os.path.dirname(REF_SO), "libgtscaf_refbam.so"))


def ref_estimate_dist(frag, ma, len_ref, len_mref, pmf, minp, rf, min_dist, max_dist):
    """estimate_dist_using_mle (bamparser.c:553-598) for one contig pair -> (dist, nof_pairs)"""
    fr = np.ascontiguousarray(np.asarray(frag, np.int64).reshape(-1, 2)).copy()       # sorted in place
    pm = np.ascontiguousarray(pmf, np.float64)
    d, n = C.c_int64(0), C.c_uint64(0)
    rc = refbam_lib().refbam_estimate_dist(fr.ctypes.data_as(C.c_void_p), C.c_uint64(len(fr)), C.c_uint64(ma),
                                           pm.ctypes.data_as(C.c_void_p), C.c_uint64(len(pm)), C.c_double(minp),
                                           C.c_int64(min_dist), C.c_int64(max_dist), C.c_uint64(len_ref),
                                           C.c_uint64(len_mref), C.c_int(int(rf)), C.byref(d), C.byref(n))
    assert rc == 0
    return d.value, n.value
