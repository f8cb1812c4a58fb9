"""SURVEY.md Appendix A.4, exhaustively on the reference's toy PRG (tests/cases/expected/dr.prg): every one of the
36 864 gid and 768 pncA full paths is spelled and sketched as a LINEAR sequence by an independent brute-force numpy
sketcher (hash64 restated here, window minima by definition), and

  * forward direction: every linear (w,k)-minimizer of every path is a k-mer node of the graph sketch — same PRG
    coordinates, same canonical hash, same strand — for the oracle's index AND the product's host index builder;
  * converse: every inner k-mer node of the graph sketch is a linear minimizer of at least one full path.

This pins the graph sketch (stage a1: node set, hashes, strands, coordinates) to the definition of a minimizer
without trusting either implementation of it."""
import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import lib
from helpers import TOY_PRG

CODE = np.full(256, 4, np.uint8)
for i, c in enumerate("ACGT"):
    CODE[ord(c)] = i
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash64(key, mask):
    """minimap2's invertible integer hash as pandora uses it, vectorised over uint64"""
    u = np.uint64
    key = (~key + (key << u(21))) & mask
    key = key ^ (key >> u(24))
    key = (key + (key << u(3)) + (key << u(8))) & mask
    key = key ^ (key >> u(14))
    key = (key + (key << u(2)) + (key << u(4))) & mask
    key = key ^ (key >> u(28))
    key = (key + (key << u(31))) & mask
    return key


def linear_minimizers(codes, w, k):
    """positions i whose canonical hash equals the minimum of some window of w consecutive k-mers containing i
    (all ties kept), by definition: O(n w)"""
    n = len(codes) - k + 1
    if n < w:
        return np.zeros(0, np.int64), None, None
    c = codes.astype(np.uint64)
    f = np.zeros(n, np.uint64)
    r = np.zeros(n, np.uint64)
    for j in range(k):
        f = (f << np.uint64(2)) | c[j:j + n]
        r = r | ((np.uint64(3) - c[j:j + n]) << np.uint64(2 * j))
    mask = np.uint64((1 << (2 * k)) - 1)
    hf, hr = hash64(f, mask), hash64(r, mask)
    h = np.minimum(hf, hr)
    win = np.lib.stride_tricks.sliding_window_view(h, w)          # (n-w+1, w)
    wmin = win.min(axis=1)
    is_min = np.zeros(n, bool)
    for j in range(w):                                            # window s covers positions s..s+w-1
        is_min[j:j + len(wmin)] |= h[j:j + len(wmin)] == wmin
    pos = np.nonzero(is_min)[0]
    return pos, h[pos], (hf <= hr)[pos]


def full_paths(lg):
    """every source-to-sink node path of a local graph (oracle layout: start, len, n_out, edges)"""
    eo = np.concatenate([[0], np.cumsum(lg["n_out"])]).astype(int)
    out = [lg["edges"][eo[i]:eo[i + 1]].tolist() for i in range(len(lg["start"]))]
    stack = [(0, (0,))]
    while stack:
        node, path = stack.pop()
        if not out[node]:
            yield path
            continue
        for nxt in out[node]:
            stack.append((nxt, path + (nxt,)))


def knode_table(kn, base, n):
    """(run-compressed base coordinates of the k-mer) -> (hash, strand) for the inner k-mer nodes of one locus"""
    io = np.concatenate([[0], np.cumsum(kn["n_iv"])]).astype(int)
    table = {}
    for g in range(base + 1, base + n - 1):  # first and last node of a locus are the null start / end
        ivs = tuple((int(s), int(l)) for s, l in zip(kn["iv_start"][io[g]:io[g + 1]], kn["iv_len"][io[g]:io[g + 1]]) if l > 0)
        assert ivs not in table, "two k-mer nodes over the same bases"
        table[ivs] = (int(kn["hash"][g]), int(kn["strand"][g]))
    return table


def runs(coords):
    cut = np.nonzero(np.diff(coords) != 1)[0] + 1
    return tuple((int(seg[0]), len(seg)) for seg in np.split(coords, cut))


def condense(rows):
    """unique (first, last, coordinate sum, sum of squares, hash, strand) signatures, one representative (path, position) each"""
    R = np.concatenate(rows)
    _, idx = np.unique(R[:, :6], axis=0, return_index=True)
    return R[idx]


@pytest.mark.parametrize("w,k", [(11, 15), (14, 15)])
def test_every_path_of_the_toy_prg_both_directions(w, k):
    ox = O.Index(TOY_PRG, w, k)
    gx = lib.Index(TOY_PRG, w, k, device=-1)
    text = open(TOY_PRG).read().splitlines()
    okn, gkn = ox.knodes(), gx.knodes()
    expected_paths = {"gid": 36864, "pncA": 768}
    for li, name in enumerate(ox.names):
        body = np.frombuffer(text[2 * li + 1].encode(), np.uint8)
        lg = ox.local_graph(li)
        node_coords = [np.arange(int(s), int(s + l), dtype=np.int64) for s, l in zip(lg["start"], lg["len"])]
        base, n = int(ox.knode_base[li]), int(ox.knode_base[li + 1] - ox.knode_base[li])
        tables = [knode_table(okn, base, n), knode_table(gkn, base, n)]
        assert tables[0] == tables[1]
        table = tables[0]
        # Every minimizer of every path is reduced to a signature first (all vectorised); each distinct signature is then
        # checked exactly, on the coordinates of a representative occurrence.
        paths, rows, uniq = [], [], None
        for path in full_paths(lg):
            coords = np.concatenate([node_coords[v] for v in path])
            codes = CODE[body[coords]]
            assert (codes < 4).all()
            pos, hs, st = linear_minimizers(codes, w, k)
            c1 = np.concatenate([[0], np.cumsum(coords)])
            c2 = np.concatenate([[0], np.cumsum(coords * coords)])
            R = np.empty((len(pos), 8), np.int64)
            R[:, 0], R[:, 1] = coords[pos], coords[pos + k - 1]
            R[:, 2], R[:, 3] = c1[pos + k] - c1[pos], c2[pos + k] - c2[pos]
            R[:, 4], R[:, 5] = hs.astype(np.int64), st
            R[:, 6], R[:, 7] = len(paths), pos
            rows.append(R)
            paths.append(path)
            if len(rows) >= 2048:
                uniq = condense(rows + ([uniq] if uniq is not None else []))
                rows = []
        uniq = condense(rows + ([uniq] if uniq is not None else []))
        assert len(paths) == expected_paths[name]
        seen = set()
        for first, last, s1, s2, h, s, pi, p in uniq.tolist():
            coords = np.concatenate([node_coords[v] for v in paths[pi]])
            key = runs(coords[p:p + k])
            # forward direction: the linear minimizer is a k-mer node with this hash and strand
            assert key in table, (name, paths[pi], p)
            assert table[key] == (h, s), (name, key, table[key], h, s)
            seen.add(key)
        # converse: every inner k-mer node is a linear minimizer of some full path
        missing = [key for key in table if key not in seen]
        assert not missing, (name, len(missing), len(table), missing[:5])
