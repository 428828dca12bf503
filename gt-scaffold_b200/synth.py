"""Seeded synthetic scaffold graphs of the shapes BASELINE.json names.

Integer vertex ids and file-ordered integer distance records -- what the
reference's text front end (parser.c:295-394) hands to its insert/dedup loop
after header lookup.  Record layout mirrors a `.de` file: records are grouped
by root contig ("line"), sense records before antisense ones (the `;`
separator, parser.c:382-383), and every link is normally listed on both
contigs' lines (testdata/libPE.de:1,4).

Distributions follow SURVEY.md section 8(d).  Everything is numpy on the host;
the generator is input plumbing, not part of the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SENSE = 1
SAME = 2

_MASK = (1 << 64) - 1


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


@dataclass
class ScaffoldInput:
    """Vertices + file-ordered records (the hot path's input)."""
    seq_len: np.ndarray      # u32 [V]
    astat: np.ndarray        # f32 [V]
    copy_num: np.ndarray     # f32 [V]
    root: np.ndarray         # u32 [R]
    ctg: np.ndarray          # u32 [R]
    dist: np.ndarray         # i32 [R]
    std_dev: np.ndarray      # f32 [R]
    num_pairs: np.ndarray    # u32 [R]
    flags: np.ndarray        # u8  [R]  bit0 sense, bit1 same
    name: str = "custom"
    meta: dict = field(default_factory=dict)

    @property
    def nof_vertices(self) -> int:
        return int(self.seq_len.shape[0])

    @property
    def nof_records(self) -> int:
        return int(self.root.shape[0])


CONFIGS = {
    # name: (config_id, V, mean pairs per vertex, kind)
    "c2_bacterial": (2, 50_000, 3.0, "uniform"),
    "c3_human": (3, 10_000_000, 4.0, "uniform"),
    "c4_repeat_hubs": (4, 5_000_000, None, "powerlaw"),
    "c5_metagenome": (5, 100_000_000, 4.0, "uniform"),
}


def _vertices(rng: np.random.Generator, V: int):
    seq_len = np.clip(np.rint(rng.lognormal(7.6, 1.1, V)), 201, 200000).astype(np.uint32)
    u = rng.random(V)
    astat = rng.uniform(20.5, 6000.0, V)
    rep = u < 0.10
    # U(-400, 20.0]: mirror a [0,1) draw so that exactly 20.0 is reachable
    astat[rep] = 20.0 - 420.0 * rng.random(int(rep.sum()))
    astat[(u >= 0.10) & (u < 0.12)] = 0.0
    astat = astat.astype(np.float32)
    if V:
        # make "exactly at the cutoff" certain to occur
        astat[rng.integers(0, V, max(1, V // 1000))] = np.float32(20.0)
    u = rng.random(V)
    cn = rng.normal(1.0, 0.06, V)
    low = u < 0.06
    cn[low] = rng.uniform(0.05, 0.75, int(low.sum()))
    high = (u >= 0.06) & (u < 0.10)
    cn[high] = rng.uniform(1.6, 6.0, int(high.sum()))
    cn = cn.astype(np.float32)
    if V > 1:
        # 5 % duplicated from the previous vertex: forces the tie rule of
        # check_mark_polymorphic (algorithms.c:235-238)
        dup = np.nonzero(rng.random(V) < 0.05)[0]
        dup = dup[dup > 0]
        cn[dup] = cn[dup - 1]
    return seq_len, astat, cn


def _uniform_pairs(rng, V: int, mean_pairs: float):
    P = int(round(mean_pairs * V))
    if V < 2 or P == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    a = rng.integers(0, V, P, dtype=np.int64)
    b = rng.integers(0, V - 1, P, dtype=np.int64)
    b += b >= a                                   # uniform over the others
    return a, b


def _powerlaw_pairs(rng, V: int, alpha: float, max_deg: int, hubs_to_hubs: float):
    """Degree ~ Zipf(alpha) truncated at max_deg; stubs matched to uniform
    partners (so hubs mostly see low-degree neighbours)."""
    if V < 2:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    deg = np.minimum(rng.zipf(alpha, V), max_deg).astype(np.int64)
    # each link consumes one stub of its owner; partner uniform
    a = np.repeat(np.arange(V, dtype=np.int64), deg)
    b = rng.integers(0, V - 1, a.shape[0], dtype=np.int64)
    b += b >= a
    return a, b


def generate(name: str = "c2_bacterial", V: int | None = None, *,
             mean_pairs: float | None = None, seed: int | None = None,
             line_order: str = "shuffled", one_sided_frac: float = 0.0,
             mirror_diff_frac: float = 0.01, dup_same_line_frac: float = 0.005,
             max_deg: int = 10_000, zipf_alpha: float = 2.1,
             one_sided_up: bool = False) -> ScaffoldInput:
    """Generate one of the named configs (optionally at a reduced V).
    one_sided_up: links listed on one line only are listed on the EARLIER of the
    two lines (the only one-sided case the line-ordered build handles itself)."""
    config_id, V0, mp0, kind = CONFIGS[name]
    V = V0 if V is None else int(V)
    mp = mp0 if mean_pairs is None else mean_pairs
    s = splitmix64(0x5CAFF01D ^ config_id) if seed is None else splitmix64(seed)
    rng = np.random.Generator(np.random.PCG64(s))

    seq_len, astat, cn = _vertices(rng, V)
    if kind == "uniform":
        a, b = _uniform_pairs(rng, V, mp)
    else:
        a, b = _powerlaw_pairs(rng, V, zipf_alpha, max_deg, 0.0)

    # unordered-pair dedup, keeping the first occurrence's orientation
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    _, first = np.unique(lo * np.int64(V) + hi, return_index=True)
    first.sort()
    a, b = a[first], b[first]
    P = a.shape[0]

    if line_order == "shuffled":
        line_rank = rng.permutation(V).astype(np.int64)
    elif line_order == "id":
        line_rank = np.arange(V, dtype=np.int64)
    else:
        raise ValueError(line_order)
    keep_b = rng.random(P) >= one_sided_frac
    if one_sided_up:
        sw = (~keep_b) & (line_rank[a] > line_rank[b])
        a, b = np.where(sw, b, a), np.where(sw, a, b)

    sense_a = rng.random(P) < 0.5
    same = rng.random(P) < 0.5
    sense_b = np.where(same, ~sense_a, sense_a)       # parser.c:369-372
    dist = rng.integers(-99, 3001, P).astype(np.int64)
    std = np.round(rng.uniform(0.5, 60.0, P), 1)
    std[rng.random(P) < 0.005] = 0.0
    npairs = rng.integers(10, 801, P).astype(np.uint32)

    # 30 %: distance copied from a sibling link of endpoint a, +- U[0,8]
    if P > 1:
        order = np.argsort(a, kind="stable")
        prev = np.empty(P, np.int64)
        prev[order[1:]] = order[:-1]
        prev[order[0]] = order[0]
        has_sib = a[prev] == a
        has_sib[order[0]] = False
        pick = has_sib & (rng.random(P) < 0.30)
        jitter = rng.integers(-8, 9, P)
        base = dist.copy()
        dist[pick] = base[prev[pick]] + jitter[pick]

    # records: one on a's line, one on b's line
    r_root = [a, b[keep_b]]
    r_ctg = [b, a[keep_b]]
    r_sense = [sense_a, sense_b[keep_b]]
    r_same = [same, same[keep_b]]
    r_dist = [dist, dist[keep_b].copy()]
    r_std = [std, std[keep_b].copy()]
    r_np = [npairs, npairs[keep_b].copy()]
    # mirrored record with its own estimate (twin-seeded compare/replace)
    nb = int(keep_b.sum())
    md = rng.random(nb) < mirror_diff_frac
    r_dist[1][md] += rng.integers(-40, 41, int(md.sum()))
    r_std[1][md] = np.round(rng.uniform(0.5, 60.0, int(md.sum())), 1)
    # duplicate record on a's own line with a second estimate
    dup = np.nonzero(rng.random(P) < dup_same_line_frac)[0]
    if dup.size:
        r_root.append(a[dup]); r_ctg.append(b[dup])
        r_sense.append(sense_a[dup]); r_same.append(same[dup])
        r_dist.append(dist[dup] + rng.integers(-40, 41, dup.size))
        r_std.append(np.round(rng.uniform(0.5, 60.0, dup.size), 1))
        r_np.append(rng.integers(10, 801, dup.size).astype(np.uint32))

    root = np.concatenate(r_root)
    ctg = np.concatenate(r_ctg)
    sense = np.concatenate(r_sense)
    same_r = np.concatenate(r_same)
    dist_r = np.concatenate(r_dist)
    std_r = np.concatenate(r_std)
    np_r = np.concatenate(r_np)
    R = root.shape[0]

    # file order: lines (roots) in `line_order`, sense records first in a line,
    # then generation order
    key = line_rank[root] * 2 + (~sense).astype(np.int64)
    order = np.argsort(key, kind="stable")

    return ScaffoldInput(
        seq_len=seq_len, astat=astat, copy_num=cn,
        root=root[order].astype(np.uint32), ctg=ctg[order].astype(np.uint32),
        dist=dist_r[order].astype(np.int32),
        std_dev=std_r[order].astype(np.float32),
        num_pairs=np_r[order].astype(np.uint32),
        flags=(sense[order].astype(np.uint8) * SENSE
               | same_r[order].astype(np.uint8) * SAME),
        name=name,
        meta={"config": name, "V": V, "pairs": int(P), "records": int(R),
              "seed": int(s), "line_order": line_order},
    )


def tiny_dense(V: int, n_pairs: int, seed: int, *, repeats: float = 0.12,
               one_sided: float = 0.3, multi: float = 0.3,
               split_lines: bool = True) -> ScaffoldInput:
    """Small adversarial graphs for differential tests: dense, many ties,
    one-sided links, repeated estimates in either direction, roots that appear
    on several lines, zero std_dev, close distances, long contigs."""
    rng = np.random.Generator(np.random.PCG64(splitmix64(seed)))
    seq_len = rng.choice([201, 350, 800, 2000, 6000], V).astype(np.uint32)
    astat = np.where(rng.random(V) < repeats, rng.choice([20.0, 0.0, -3.5, 19.5], V),
                     rng.uniform(20.5, 500.0, V)).astype(np.float32)
    cn = rng.choice([0.2, 0.4, 0.4, 0.7, 0.75, 1.0, 1.0, 1.1, 2.5], V).astype(np.float32)
    a = rng.integers(0, V, n_pairs)
    b = rng.integers(0, V - 1, n_pairs)
    b += b >= a
    recs = []
    for i in range(n_pairs):
        sense = bool(rng.random() < 0.5)
        same = bool(rng.random() < 0.5)
        d = int(rng.choice([-50, 0, 10, 12, 100, 105, 400, 1000, 1003, 2500]))
        sd = float(rng.choice([0.0, 0.5, 1.4, 3.3, 10.0, 40.0]))
        recs.append((a[i], b[i], d, sd, sense, same))
        if rng.random() >= one_sided:
            tw = (not sense) if same else sense
            d2, sd2 = d, sd
            if rng.random() < multi:
                d2 = d + int(rng.integers(-30, 31))
                sd2 = float(rng.choice([0.0, 0.5, 1.4, 3.3, 10.0, 40.0]))
            recs.append((b[i], a[i], d2, sd2, tw, same))
        if rng.random() < multi:
            recs.append((a[i], b[i], d + int(rng.integers(-30, 31)),
                         float(rng.choice([0.5, 1.4, 3.3, 10.0, 40.0, 55.5])),
                         bool(rng.random() < 0.5), bool(rng.random() < 0.5)))
    recs = [recs[j] for j in rng.permutation(len(recs))]
    if not split_lines:
        recs.sort(key=lambda r: (r[0], not r[4]))
    arr = list(zip(*recs)) if recs else [[]] * 6
    R = len(recs)
    flags = (np.array(arr[4], bool).astype(np.uint8) * SENSE
             | np.array(arr[5], bool).astype(np.uint8) * SAME) if R else np.zeros(0, np.uint8)
    return ScaffoldInput(
        seq_len=seq_len, astat=astat, copy_num=cn,
        root=np.array(arr[0], np.uint32), ctg=np.array(arr[1], np.uint32),
        dist=np.array(arr[2], np.int32), std_dev=np.array(arr[3], np.float32),
        num_pairs=rng.integers(10, 801, R).astype(np.uint32), flags=flags,
        name="tiny_dense", meta={"V": V, "records": R, "seed": seed})


# --------------------------------------------------------------------------
# torch version of generate(): same distributions, runs on the GPU so that the
# 10^7..10^8-vertex configs materialise in a second instead of minutes.  The
# random streams differ from the numpy generator (and between devices); parity
# tests always feed ONE set of arrays to both the CUDA path and the oracle.

def special_values(seed: int, V: int = 14, n_pairs: int = 70) -> ScaffoldInput:
    """tiny_dense with the values the float compares and the i64 interval
    arithmetic are most likely to get wrong: NaN / inf / signed zero / denormal /
    negative std_dev, NaN and infinite copy numbers and a-statistics, distances
    at the int32 limits, contig lengths up to 2^31-1 (algorithms.c:174-246,
    parser.c:362)."""
    rng = np.random.default_rng(100 + seed)
    inp = tiny_dense(V, n_pairs, 6000 + seed)
    sd = inp.std_dev.copy()
    specials = np.array([np.nan, np.inf, 0.0, -0.0, 1e-40, 3.4e38, -1.0], np.float32)
    idx = rng.choice(len(sd), size=len(sd) // 3, replace=False)
    sd[idx] = specials[rng.integers(0, len(specials), len(idx))]
    dist = inp.dist.copy()
    idx = rng.choice(len(dist), size=len(dist) // 5, replace=False)
    dist[idx] = rng.choice(np.array([-2**31 + 1, 2**31 - 1, 0, -1, 2**30], np.int64),
                           len(idx)).astype(np.int32)
    cn = inp.copy_num.copy()
    cn[rng.integers(0, len(cn), 4)] = np.array([np.nan, np.inf, -1.0, 0.0], np.float32)
    astat = inp.astat.copy()
    astat[rng.integers(0, len(astat), 2)] = np.array([np.nan, -np.inf], np.float32)
    seq_len = inp.seq_len.copy()
    seq_len[rng.integers(0, len(seq_len), 3)] = np.array([2**31 - 1, 1, 2**30], np.uint32)
    return ScaffoldInput(seq_len, astat, cn, inp.root, inp.ctg, dist, sd, inp.num_pairs,
                         inp.flags, name="special_values",
                         meta={"V": V, "records": len(dist), "seed": seed})


def generate_torch(name: str = "c3_human", V: int | None = None, *, device="cuda",
                   mean_pairs: float | None = None, seed: int | None = None,
                   line_order: str = "shuffled", one_sided_frac: float = 0.0,
                   mirror_diff_frac: float = 0.01, dup_same_line_frac: float = 0.005,
                   max_deg: int = 10_000, zipf_alpha: float = 2.1):
    """Returns a dict of torch tensors on `device`: seq_len(i32) astat(f32)
    copy_num(f32) root(i32) ctg(i32) dist(i32) std_dev(f32) num_pairs(i32)
    flags(u8), plus 'meta'."""
    import torch
    config_id, V0, mp0, kind = CONFIGS[name]
    V = V0 if V is None else int(V)
    mp = mp0 if mean_pairs is None else mean_pairs
    s = splitmix64(0x5CAFF01D ^ config_id) if seed is None else splitmix64(seed)
    gen = torch.Generator(device=device)
    gen.manual_seed(s & ((1 << 63) - 1))
    dev = torch.device(device)

    def rand(n):
        return torch.rand(n, generator=gen, device=dev, dtype=torch.float64)

    def randint(lo, hi, n):
        return torch.randint(lo, hi, (n,), generator=gen, device=dev, dtype=torch.int64)

    def normal(n):
        return torch.randn(n, generator=gen, device=dev, dtype=torch.float64)

    # vertices
    seq_len = torch.clamp(torch.round(torch.exp(7.6 + 1.1 * normal(V))), 201, 200000).to(torch.int32)
    u = rand(V)
    astat = 20.5 + (6000.0 - 20.5) * rand(V)
    rep = u < 0.10
    astat = torch.where(rep, 20.0 - 420.0 * rand(V), astat)
    astat = torch.where((u >= 0.10) & (u < 0.12), torch.zeros_like(astat), astat).to(torch.float32)
    if V:
        astat[randint(0, V, max(1, V // 1000))] = 20.0
    u = rand(V)
    cn = 1.0 + 0.06 * normal(V)
    cn = torch.where(u < 0.06, 0.05 + 0.70 * rand(V), cn)
    cn = torch.where((u >= 0.06) & (u < 0.10), 1.6 + 4.4 * rand(V), cn).to(torch.float32)
    if V > 1:
        dup = torch.nonzero(rand(V) < 0.05).flatten()
        dup = dup[dup > 0]
        cn[dup] = cn[dup - 1]

    # links
    if kind == "uniform":
        P = int(round(mp * V))
        a = randint(0, V, P)
    else:
        # Zipf(alpha) degrees by inverse transform on a truncated support
        k = torch.arange(1, max_deg + 1, device=dev, dtype=torch.float64)
        cdf = torch.cumsum(k ** (-zipf_alpha), 0)
        cdf = cdf / cdf[-1]
        deg = torch.searchsorted(cdf, rand(V)).clamp(max=max_deg - 1) + 1
        a = torch.repeat_interleave(torch.arange(V, device=dev, dtype=torch.int64), deg)
        P = int(a.shape[0])
    if V < 2:
        P = 0
        a = a[:0]
    b = randint(0, max(V - 1, 1), P)
    b = b + (b >= a).to(torch.int64)
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    key = torch.unique(lo * V + hi)
    key = key[torch.randperm(key.shape[0], generator=gen, device=dev)]
    lo, hi = key // V, key % V
    swap = rand(key.shape[0]) < 0.5
    a, b = torch.where(swap, hi, lo), torch.where(swap, lo, hi)
    P = int(a.shape[0])

    sense_a = rand(P) < 0.5
    same = rand(P) < 0.5
    sense_b = torch.where(same, ~sense_a, sense_a)
    dist = randint(-99, 3001, P)
    std = torch.round((0.5 + 59.5 * rand(P)) * 10.0) / 10.0
    std = torch.where(rand(P) < 0.005, torch.zeros_like(std), std)
    npairs = randint(10, 801, P)
    if P > 1:
        order = torch.sort(a, stable=True).indices
        prev = torch.empty(P, dtype=torch.int64, device=dev)
        prev[order[1:]] = order[:-1]
        prev[order[0]] = order[0]
        has_sib = a[prev] == a
        has_sib[order[0]] = False
        pick = has_sib & (rand(P) < 0.30)
        dist = torch.where(pick, dist[prev] + randint(-8, 9, P), dist)

    keep_b = rand(P) >= one_sided_frac
    md = keep_b & (rand(P) < mirror_diff_frac)
    dist_b = torch.where(md, dist + randint(-40, 41, P), dist)
    std_b = torch.where(md, torch.round((0.5 + 59.5 * rand(P)) * 10.0) / 10.0, std)
    dupm = rand(P) < dup_same_line_frac
    nd = int(dupm.sum())
    root = torch.cat([a, b[keep_b], a[dupm]])
    ctg = torch.cat([b, a[keep_b], b[dupm]])
    sense = torch.cat([sense_a, sense_b[keep_b], sense_a[dupm]])
    same_r = torch.cat([same, same[keep_b], same[dupm]])
    dist_r = torch.cat([dist, dist_b[keep_b], dist[dupm] + randint(-40, 41, nd)])
    std_r = torch.cat([std, std_b[keep_b], torch.round((0.5 + 59.5 * rand(nd)) * 10.0) / 10.0])
    np_r = torch.cat([npairs, npairs[keep_b], randint(10, 801, nd)])

    if line_order == "shuffled":
        line_rank = torch.randperm(V, generator=gen, device=dev)
    elif line_order == "id":
        line_rank = torch.arange(V, device=dev, dtype=torch.int64)
    else:
        raise ValueError(line_order)
    okey = line_rank[root] * 2 + (~sense).to(torch.int64)
    order = torch.sort(okey, stable=True).indices
    out = dict(
        seq_len=seq_len, astat=astat, copy_num=cn,
        root=root[order].to(torch.int32).contiguous(), ctg=ctg[order].to(torch.int32).contiguous(),
        dist=dist_r[order].to(torch.int32).contiguous(),
        std_dev=std_r[order].to(torch.float32).contiguous(),
        num_pairs=np_r[order].to(torch.int32).contiguous(),
        flags=(sense[order].to(torch.uint8) * SENSE + same_r[order].to(torch.uint8) * SAME).contiguous(),
        meta={"config": name, "V": V, "pairs": P, "records": int(root.shape[0]), "seed": int(s),
              "line_order": line_order, "generator": "torch:" + str(dev)})
    return out


def torch_to_input(t) -> ScaffoldInput:
    """generate_torch() result -> host ScaffoldInput (numpy)."""
    g = lambda k, dt: t[k].detach().cpu().numpy().astype(dt, copy=False)
    return ScaffoldInput(seq_len=g("seq_len", np.uint32), astat=g("astat", np.float32),
                         copy_num=g("copy_num", np.float32), root=g("root", np.uint32),
                         ctg=g("ctg", np.uint32), dist=g("dist", np.int32),
                         std_dev=g("std_dev", np.float32), num_pairs=g("num_pairs", np.uint32),
                         flags=g("flags", np.uint8), name=t["meta"]["config"], meta=dict(t["meta"]))
