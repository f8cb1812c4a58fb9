"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (numpy / plain Python, small inputs) of the mapping front half of
`pandora discover` that drprg launches at /root/reference/src/predict.rs:247-256 -> src/lib.rs:513-578, computed from
the oracle's own map results (oracle_py.MapRun / Genotype).  Upstream pandora functions restated (source not in the
reference tree: parity unpinned; names for when a source tree is available): denovo_discovery/candidate_region.cpp
find_candidate_regions_for_pan_node, get_covgs_along_localnode_path, identify_low_coverage_intervals,
get_read_overlap_coordinates / find_hits_inside_path; option defaults of discover_main.cpp (--covg-threshold 3, -l 1, -L 30,
-P 22, at least two hits of a read inside a region)."""
import numpy as np


def _locus_tables(ox, locus):
    lg = ox.local_graph(locus)
    kn = ox.knodes()
    base, n = int(ox.knode_base[locus]), int(ox.knode_base[locus + 1] - ox.knode_base[locus])
    io = np.concatenate([[0], np.cumsum(kn["n_iv"])]).astype(int)
    eo = np.concatenate([[0], np.cumsum(lg["n_out"])]).astype(int)
    out = [lg["edges"][eo[i]:eo[i + 1]].tolist() for i in range(len(lg["start"]))]
    starts, lens = lg["start"].astype(int), lg["len"].astype(int)

    def node_of(s, l):
        if l == 0:
            hit = [i for i in range(len(starts)) if starts[i] == s and lens[i] == 0]
        else:
            hit = [i for i in range(len(starts)) if lens[i] > 0 and starts[i] <= s < starts[i] + lens[i]]
        assert len(hit) == 1, (s, l, hit)
        return hit[0]

    kpaths = []
    for r in range(n):
        g = base + r
        if r == 0 or r == n - 1:  # the null start / end nodes of the k-mer graph carry no bases
            kpaths.append([])
            continue
        kpaths.append([(node_of(int(s), int(l)), int(s), int(l)) for s, l in zip(kn["iv_start"][io[g]:io[g + 1]], kn["iv_len"][io[g]:io[g + 1]])])
    return starts, lens, out, kpaths


def local_path(starts, lens, out, kpaths, ml):
    """localnode_path_from_kmernode_path: the nodes under the ML k-mers, extended to node 0 and to the sink"""
    lp = []
    for r in ml:
        kp = kpaths[r]
        if not kp or sum(l for _, _, l in kp) == 0 and len(kp) == 1:
            continue
        first = kp[0][0]
        while lp and out[lp[-1]] and first > out[lp[-1]][0] and first not in out[lp[-1]]:
            lp.append(out[lp[-1]][0])
        while lp and first <= lp[-1]:
            lp.pop()
        lp.extend(n for n, _, _ in kp)
    if not lp:
        lp = [0]
    if lp[0] != 0:
        target = lp[0]
        reaches = {target}
        for i in range(target - 1, -1, -1):
            if any(o <= target and o in reaches for o in out[i]):
                reaches.add(i)
        head, cur = [], 0
        while cur != target:
            head.append(cur)
            cur = next(o for o in out[cur] if o <= target and o in reaches)
        lp = head + lp
    while out[lp[-1]]:
        lp.append(out[lp[-1]][0])
    return lp


def discover(ox, mr, og, text_lines, covg_threshold=3, min_len=1, max_len=30, padding=22, min_hits=2):
    """-> ({locus: (consensus, coverage)}, [(locus, start, end, pad_start, pad_end, [(read, start, end, fwd)])])"""
    k = ox.k
    f, r = mr.coverage()
    hits = mr.hits()
    loci, regions = {}, []
    for locus in range(ox.n_loci):
        ml = og.mlpath(locus)
        if ml is None:
            continue
        starts, lens, out, kpaths = _locus_tables(ox, locus)
        lp = local_path(starts, lens, out, kpaths, ml.tolist())
        body = text_lines[2 * locus + 1]
        node_off, off = {}, 0
        for i, n in enumerate(lp):
            node_off[n] = (off, i)
            off += lens[n]
        cons = "".join(body[starts[n]:starts[n] + lens[n]] for n in lp)
        base = int(ox.knode_base[locus])
        cov = np.zeros(off, np.uint32)
        for rnk in ml.tolist():
            c = min(int(f[base + rnk]), 65535) + min(int(r[base + rnk]), 65535)
            for n, s, l in kpaths[rnk]:
                if l == 0 or n not in node_off:
                    continue
                a = node_off[n][0] + (s - starts[n])
                cov[a:a + l] = np.maximum(cov[a:a + l], c)
        loci[locus] = (cons, cov)
        # identify_low_coverage_intervals
        iv, cur, n = [], 0, len(cov)
        while cur < n:
            prev = cur
            while cur < n and cov[cur] < covg_threshold:
                cur += 1
            if min_len <= cur - prev <= max_len and cur > prev:
                iv.append((prev, cur))
            if cur == n:
                break
            cur += 1
        if not iv:
            continue
        # consensus start of every k-mer node lying on the local path
        cstart = {}
        for rnk, kp in enumerate(kpaths):
            if not kp or sum(l for _, _, l in kp) == 0:
                continue
            if any(nd not in node_off for nd, _, _ in kp):
                continue
            if any(node_off[kp[i][0]][1] != node_off[kp[i - 1][0]][1] + 1 for i in range(1, len(kp))):
                continue
            cstart[rnk] = node_off[kp[0][0]][0] + (kp[0][1] - starts[kp[0][0]])
        sel = (hits["prg"] == locus) & (hits["kept"] == 1)
        hr, hs, hk, hf = hits["read"][sel], hits["start"][sel], hits["knode"][sel], hits["fwd"][sel]
        for a, b in iv:
            pa, pb = max(0, a - padding), min(n, b + padding)
            per_read = {}
            for rd, st, kn_, fw in zip(hr.tolist(), hs.tolist(), hk.tolist(), hf.tolist()):  # pandora order within a read
                cs = cstart.get(kn_)
                if cs is None or cs < pa or cs + k > pb:
                    continue
                acc = per_read.setdefault(rd, [0, st, 0, fw])
                acc[0] += 1
                acc[1] = min(acc[1], st)
                acc[2] = max(acc[2], st + k)
            reads = sorted((rd, v[1], v[2], v[3]) for rd, v in per_read.items() if v[0] >= min_hits)
            regions.append((locus, a, b, pa, pb, reads))
    return loci, regions
