"""TEST INFRASTRUCTURE: the `.de` tokeniser's stage functions
(gt-scaffold_b200/csrc/gtsb_parse_core.h, the bodies of the kernels in
gtsb_parse.cu) compiled for the host and run as plain loops
(tests/emul/parse_emul.cpp).  Used by tests/test_parse.py to check the token
rules against the compiled reference where there is no GPU, and to generate
the `.de` texts both the emulation and the device path are tested on."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "gt-scaffold_b200", "csrc")
SRCS = [os.path.join(HERE, "emul", "parse_emul.cpp"), os.path.join(HERE, "emul", "format_emul.cpp"),
        os.path.join(HERE, "emul", "mle_emul.cpp"), os.path.join(HERE, "emul", "sort_emul.cpp")]
CORES = [os.path.join(CSRC, "gtsb_parse_core.h"), os.path.join(CSRC, "gtsb_format_core.h"),
         os.path.join(CSRC, "gtsb_mle_core.h"), os.path.join(CSRC, "gtsb_sort_core.h")]
OUT = os.path.join(HERE, "emul", "_build", "libparse_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        if (not os.path.exists(OUT)
                or os.path.getmtime(OUT) < max(os.path.getmtime(f) for f in SRCS + CORES)):
            # -ffp-contract=off: the float rule divides, it must not be fused with anything
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                            "-o", OUT] + SRCS, check=True)
        _lib = C.CDLL(OUT)
        _lib.emul_canonical_float.restype = C.c_uint32
    return _lib


def pack_names(names):
    """list of bytes (vertex id order) -> (blob, offsets) as the C ABI takes them"""
    off = np.zeros(len(names) + 1, np.uint64)
    if names:
        off[1:] = np.cumsum([len(x) for x in names], dtype=np.uint64)
    return b"".join(names), off


def parse(names, text, order=0):
    """-> (irregular bits, dict of record arrays or None)"""
    L = lib()
    blob, off = pack_names(names)
    R = C.c_uint64(0)
    irr = C.c_uint32(0)
    L.emul_parse_de(C.c_uint64(len(names)), blob, off.ctypes.data_as(C.c_void_p), text,
                    C.c_uint64(len(text)), C.c_int(order), C.byref(R), C.byref(irr))
    if irr.value:
        return irr.value, None
    n = R.value
    rec = dict(root=np.zeros(n, np.uint32), ctg=np.zeros(n, np.uint32), dist=np.zeros(n, np.int32),
               std_dev=np.zeros(n, np.float32), flags=np.zeros(n, np.uint8),
               num_pairs=np.zeros(n, np.uint32))
    L.emul_fetch(*[rec[k].ctypes.data_as(C.c_void_p)
                   for k in ("root", "ctg", "dist", "std_dev", "flags", "num_pairs")])
    return 0, rec


def parse_astat(names, text, astat, copy_num, order=0):
    """-> (irregular bits, astat, copy_num) -- copies of the inputs with the text applied"""
    L = lib()
    blob, off = pack_names(names)
    a = np.array(astat, np.float32)
    cn = np.array(copy_num, np.float32)
    irr = C.c_uint32(0)
    L.emul_parse_astat(C.c_uint64(len(names)), blob, off.ctypes.data_as(C.c_void_p), text,
                       C.c_uint64(len(text)), C.c_int(order), a.ctypes.data_as(C.c_void_p),
                       cn.ctypes.data_as(C.c_void_p), C.byref(irr))
    return irr.value, a, cn


def canonical_float(s: bytes):
    out = C.c_float(0)
    r = lib().emul_canonical_float(s, C.c_uint32(len(s)), C.byref(out))
    return r, out.value


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def dot_vertex_lines(names, vstate, first=0, scaffold_only=False):
    """.dot lines of vertices [first, first + len(vstate)) -> bytes (None: state out of range)"""
    blob, off = pack_names(names)
    vs = np.ascontiguousarray(vstate, np.uint8)
    cap = 64 * len(vs) + len(blob) + 16
    out = C.create_string_buffer(cap)
    n = C.c_uint64(0)
    rc = lib().emul_dot_vertex_lines(C.c_int(int(scaffold_only)), C.c_uint64(first), C.c_uint64(len(vs)), _vp(vs),
                                     blob, _vp(off), out, C.c_uint64(cap), C.byref(n))
    return None if rc else out.raw[:n.value]


def dot_edge_lines(src, dst, dist, estate, sense, scaffold_only=False):
    a = [np.ascontiguousarray(src, np.uint32), np.ascontiguousarray(dst, np.uint32),
         np.ascontiguousarray(dist, np.int32), np.ascontiguousarray(estate, np.uint8),
         np.ascontiguousarray(sense, np.uint8)]
    cap = 105 * len(a[0]) + 16
    out = C.create_string_buffer(cap)
    n = C.c_uint64(0)
    rc = lib().emul_dot_edge_lines(C.c_int(int(scaffold_only)), C.c_uint64(len(a[0])), *[_vp(x) for x in a],
                                   out, C.c_uint64(cap), C.byref(n))
    return None if rc else out.raw[:n.value]


def f6(bits: int) -> bytes:
    """printf("%f") of the float with these bits, by the integer routine of gtsb_format_core.h"""
    out = C.create_string_buffer(64)
    n = lib().emul_f6(C.c_uint32(bits), out)
    assert n != 0xFFFFFFFF, "f6_len and put_f6 disagree"
    return out.raw[:n]


def scaf_arrays(records):
    """records = [(root id, [(end id, dist, std_dev f32, sense, same), ...]), ...] -> the flat arrays of
    gtsb_scaf_lines_host"""
    root = np.array([r for r, _ in records], np.uint32)
    off = np.zeros(len(records) + 1, np.uint64)
    off[1:] = np.cumsum([len(e) for _, e in records])
    edges = [x for _, e in records for x in e]
    end = np.array([x[0] for x in edges], np.uint32)
    dist = np.array([x[1] for x in edges], np.int64)
    std = np.array([x[2] for x in edges], np.float32)
    flags = np.array([(1 if x[3] else 0) | (2 if x[4] else 0) for x in edges], np.uint8)
    return root, off, end, dist, std, flags


def scaf_lines(names, records):
    blob, off = pack_names(names)
    root, reo, end, dist, std, flags = scaf_arrays(records)
    cap = len(blob) * 0 + sum(len(names[r]) + 1 for r in root) + sum(len(names[w]) + 80 for w in end) + 16
    out = C.create_string_buffer(cap)
    n = C.c_uint64(0)
    rc = lib().emul_scaf_lines(C.c_uint64(len(root)), _vp(root), _vp(reo), _vp(end), _vp(dist), _vp(std.view(np.uint32)),
                               _vp(flags), blob, _vp(off), C.c_uint64(len(names)), out, C.c_uint64(cap), C.byref(n))
    return None if rc else out.raw[:n.value]


def mle_arrays(pairs):
    """pairs = [(frag (n, 2) int64 of (start, end), ma, len_ref, len_mref), ...] -> flat arrays of gtsb_mle_host"""
    off = np.zeros(len(pairs) + 1, np.uint64)
    off[1:] = np.cumsum([len(p[0]) for p in pairs])
    fr = np.concatenate([np.asarray(p[0], np.int64).reshape(-1, 2) for p in pairs]) if pairs else np.zeros((0, 2), np.int64)
    return (off, np.ascontiguousarray(fr[:, 0]), np.ascontiguousarray(fr[:, 1]),
            np.array([p[1] for p in pairs], np.uint64), np.array([p[2] for p in pairs], np.uint64),
            np.array([p[3] for p in pairs], np.uint64))


def mle(pairs, pmf, minp, rf, min_dist, max_dist, keep_all=False):
    """host build of gtsb_mle_host -> (dist[], pairs_used[], slots, candidates) or None"""
    off, fs, fe, ma, lr, lm = mle_arrays(pairs)
    pmf = np.ascontiguousarray(pmf, np.float64)
    dist, used = np.zeros(len(pairs), np.int64), np.zeros(len(pairs), np.uint64)
    slots, cand = C.c_uint64(0), C.c_uint64(0)
    rc = lib().emul_mle(C.c_uint64(len(pairs)), _vp(off), _vp(fs), _vp(fe), _vp(ma), _vp(lr), _vp(lm), _vp(pmf),
                        C.c_uint64(len(pmf)), C.c_double(minp), C.c_int(int(rf)), C.c_int64(min_dist),
                        C.c_int64(max_dist), C.c_int(int(keep_all)), _vp(dist), _vp(used), C.byref(slots), C.byref(cand))
    return None if rc else (dist, used, slots.value, cand.value)
