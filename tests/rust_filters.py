"""numpy-f32 restatement of the record statistics drprg computes from a pandora VCF record (test infrastructure):
Filterer::_covg_for_gt and the strand-bias ratio (/root/reference/src/filter.rs:212-301), VcfExt::fraction_read_support /
depth_proportions / has_no_depth (src/lib.rs:980-1011, 1165-1180), MinorAllele::check_for_minor_alternate
(src/minor.rs:70-127) and the nulling of calls without depth (src/predict.rs:440-444)."""
import numpy as np

F = np.float32
MAX_GAPS, MAX_CALLED_GAPS, MAX_GAPS_DIFF, MINOR_MIN_COVG, MINOR_MIN_SB = F(0.5), F(0.39), F(0.2), 3, F(0.01)


def effective_gt(fc, rc, gt, gt_conf):
    """src/predict.rs:440-444: a record with no depth and GT_CONF 0 gets a null call before the filters see it"""
    return -1 if (sum(fc) + sum(rc) == 0 and F(gt_conf) == F(0)) else gt


def covg_for_gt(fc, rc, gt):
    return sum(fc) + sum(rc) if gt < 0 else fc[gt] + rc[gt]


def fraction_read_support(fc, rc, gt):
    if len(fc) < 2:
        return F(1.0)
    if gt < 0:
        return None
    called = F(fc[gt] + rc[gt])
    other = 0
    if gt > 0:
        other = fc[0] + rc[0]
    else:
        for i, (f, r) in enumerate(zip(fc, rc)):
            if i != gt and f + r > other:
                other = f + r
    with np.errstate(invalid="ignore", divide="ignore"):
        v = called / (called + F(other))
    return None if np.isnan(v) else v


def strand_bias_ratio(fc, rc, gt):
    if gt < 0:
        tf, tr = F(sum(fc)), F(sum(rc))
        tot = tf + tr
        return None if tot == 0 else min(tf, tr) / tot
    f, r = F(fc[gt]), F(rc[gt])
    s = f + r
    return None if s == 0 else min(f, r) / s


def depth_proportions(fc, rc):
    d = [F(f + r) for f, r in zip(fc, rc)]
    tot = F(sum(d, F(0)))
    return None if tot == 0 else [x / tot for x in d]


def minor_alternate(fc, rc, gaps, gt, maf):
    props = depth_proportions(fc, rc)
    if len(fc) < 2 or props is None or gt < 0:
        return -1
    gaps = [F(g) for g in gaps]
    order = sorted(range(len(props)), key=lambda i: props[i])  # stable ascending, like sort_by(total_cmp)
    if gaps[gt] > MAX_CALLED_GAPS:
        return -1
    pick = None
    for i in reversed(order):
        if i == gt:
            continue
        if props[i] >= F(maf) and gaps[i] <= MAX_GAPS and gaps[i] - gaps[gt] <= MAX_GAPS_DIFF:
            pick = i
            break
    if pick is None:
        return -1
    s = F(fc[pick] + rc[pick])
    low = fc[pick] + rc[pick] < MINOR_MIN_COVG
    bias = True if s == 0 else F(min(fc[pick], rc[pick])) / s < MINOR_MIN_SB
    return -1 if (low or bias) else pick


def check_against(stats, rec_off, mean_fwd, mean_rev, gaps, gt, gt_conf, maf):
    """assert the device statistics (dict of arrays) equal the restatement for every record"""
    n = len(rec_off) - 1
    for r in range(n):
        b, e = int(rec_off[r]), int(rec_off[r + 1])
        fc, rc, g = [int(x) for x in mean_fwd[b:e]], [int(x) for x in mean_rev[b:e]], list(gaps[b:e])
        egt = effective_gt(fc, rc, int(gt[r]), gt_conf[r])
        assert int(stats["covg_gt"][r]) == covg_for_gt(fc, rc, egt), ("covg_gt", r)
        for name, want in (("frs", fraction_read_support(fc, rc, egt)), ("sb_ratio", strand_bias_ratio(fc, rc, egt))):
            got = stats[name][r]
            if want is None:
                assert np.isnan(got), (name, r, got)
            else:
                assert got == want, (name, r, got, want)
        props = depth_proportions(fc, rc)
        if props is None:
            assert np.isnan(stats["pdp"][b:e]).all(), ("pdp", r)
        else:
            assert (stats["pdp"][b:e] == np.array(props, F)).all(), ("pdp", r)
        assert int(stats["minor_gt"][r]) == minor_alternate(fc, rc, g, egt, maf), ("minor_gt", r, fc, rc, g, egt)
    return n
