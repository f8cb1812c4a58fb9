"""The CUDA genotype kernel against the vectors the REFERENCE holds: every data row of the seven pandora VCF
fixtures under /root/reference/tests/cases/predict (copied to tests/golden/) goes through genotype_kernel — the
kernel of the product path — by the C-ABI hook drprg_cuda_genotype_rows, and its LIKELIHOOD / GT / GT_CONF must
reproduce what pandora printed (6 significant digits).  Same rows, tolerances and hand-edited exclusions as the
oracle's pin in tests/test_oracle_golden.py."""
import math
import os

import numpy as np
import pytest

import rust_filters
from drprg_b200 import lib
from test_oracle_golden import FIXTURE_E, HAND_EDITED, vcf_rows

pytestmark = pytest.mark.gpu


def fixture_rows(golden, name):
    keys, rec_off, mf, mr, gaps, want, gts, confs = [], [0], [], [], [], [], [], []
    for f, v in vcf_rows(os.path.join(golden, name)):
        if (name, f[0], int(f[1])) in HAND_EDITED:
            continue
        a = [int(x) for x in v["MEAN_FWD_COVG"].split(",")]
        mf += a
        mr += [int(x) for x in v["MEAN_REV_COVG"].split(",")]
        gaps += [float(x) for x in v["GAPS"].split(",")]
        want += [float(x) for x in v["LIKELIHOOD"].split(",")]
        rec_off.append(rec_off[-1] + len(a))
        gts.append(v["GT"])
        confs.append(float(v["GT_CONF"]))
        keys.append((f[0], f[1]))
    return keys, rec_off, mf, mr, gaps, want, gts, confs


@pytest.mark.parametrize("name", sorted(FIXTURE_E))
def test_genotype_kernel_reproduces_reference_vcf_rows(golden, name):
    keys, rec_off, mf, mr, gaps, want, gts, confs = fixture_rows(golden, name)
    lik, gt, conf = lib.genotype_rows(rec_off, mf, mr, gaps, FIXTURE_E[name])
    n_alleles = 0
    for r, key in enumerate(keys):
        b, e = rec_off[r], rec_off[r + 1]
        for a in range(b, e):
            # 6 significant digits printed; GAPS is itself rounded to 6 digits and multiplied by E
            assert math.isclose(lik[a], want[a], rel_tol=2e-5, abs_tol=2e-3), (name, key, lik[b:e], want[b:e])
            n_alleles += 1
        if gts[r] == ".":
            continue
        srt = sorted(lik[b:e], reverse=True)
        if len(srt) > 1 and srt[0] - srt[1] > 1e-3:
            assert int(gts[r]) == int(gt[r]), (name, key)
        assert math.isclose(conf[r], confs[r], rel_tol=2e-4, abs_tol=5e-3), (name, key, conf[r], confs[r])
    assert n_alleles >= 12


def test_genotype_rows_zero_depth_and_min_conf(golden):
    # zero-depth site: every allele gets exactly -2E, GT 0, GT_CONF 0 (ERR4796933.pandora.vcf ethA 19: -144,-144)
    lik, gt, conf = lib.genotype_rows([0, 2], [0, 0], [0, 0], [1.0, 1.0], 72)
    assert lik.tolist() == [-144.0, -144.0] and gt.tolist() == [0] and conf.tolist() == [0.0]
    # --gt-conf above the confidence nulls the call
    lik, gt, conf = lib.genotype_rows([0, 2], [40, 0], [30, 1], [0.0, 1.0], 72, min_gt_conf=1e6)
    assert gt.tolist() == [-1] and conf[0] > 100


@pytest.mark.parametrize("name", sorted(FIXTURE_E))
@pytest.mark.parametrize("maf", [0.1, 1.0])
def test_fused_filter_statistics_on_reference_vcf_rows(golden, name, maf):
    """the statistics drprg's Filterer / MinorAllele derive from a pandora record, computed by the genotype
    kernel on the real pandora rows of the fixtures, against a numpy-f32 restatement of the Rust functions"""
    keys, rec_off, mf, mr, gaps, want, gts, confs = fixture_rows(golden, name)
    lik, gt, conf, stats = lib.genotype_rows(rec_off, mf, mr, gaps, FIXTURE_E[name], minor_af=maf, stats=True)
    n = rust_filters.check_against(stats, rec_off, mf, mr, gaps, gt, conf, maf)
    assert n == len(keys) and n >= 4
    if name == "in.vcf":
        assert np.isnan(stats["frs"]).sum() > 0  # null calls have no FRS


def test_fused_minor_allele_statistic_constructed_rows():
    """rows built to hit every branch of check_for_minor_alternate (src/minor.rs:70-127)"""
    rows = [
        ([40, 6], [38, 5], [0.0, 0.1]),        # alt holds 12 % of the depth, clean GAPS, both strands: minor allele 1
        ([40, 6], [38, 0], [0.0, 0.1]),        # alt only on one strand: strand bias -> none
        ([40, 1], [38, 1], [0.0, 0.1]),        # alt depth 2 < 3 -> none
        ([40, 6], [38, 5], [0.0, 0.6]),        # alt GAPS above 0.5 -> none
        ([40, 6], [38, 5], [0.4, 0.45]),       # called GAPS above 0.39 -> none
        ([40, 6], [38, 5], [0.1, 0.35]),       # GAPS difference above 0.2 -> none
        ([30, 5, 5], [30, 5, 5], [0.0, 0.0, 0.0]),  # two alts with equal proportions: the higher index wins
        ([0, 0], [0, 0], [1.0, 1.0]),          # no depth: null call, nothing
        ([5], [4], [0.0]),                     # single allele
    ]
    rec_off, mf, mr, gaps = [0], [], [], []
    for f, r, g in rows:
        mf += f; mr += r; gaps += g
        rec_off.append(rec_off[-1] + len(f))
    lik, gt, conf, stats = lib.genotype_rows(rec_off, mf, mr, gaps, 80, minor_af=0.1, stats=True)
    rust_filters.check_against(stats, rec_off, mf, mr, gaps, gt, conf, 0.1)
    assert stats["minor_gt"].tolist() == [1, -1, -1, -1, -1, -1, 2, -1, -1]
    assert stats["covg_gt"].tolist()[:2] == [78, 78] and stats["frs"][8] == 1.0 and np.isnan(stats["frs"][7])


def test_device_float_formatting_matches_printf_g():
    """the VCF record lines are formatted by a kernel: its "%g" must be byte-identical to printf wherever it answers, and it
    must decline (host fallback) exactly the cases it cannot print that way: exponent notation, rounding ties, -0, inf, nan"""
    import ctypes as C
    L = lib.lib()
    rng = np.random.default_rng(0)
    vals = [0.0, 1.0, -1.0, 0.5, 0.25, 0.125, 0.2, 1 / 3, 2 / 3, 0.666667, 1e-4, 9.99999e-5, 123456.5, 999999.4, 999999.5,
            999999.6, 1e6, -144.0, -0.0001, 100000.0, 99999.95, 0.000123456789, 5e-324, 1e300, float("inf"), float("nan"), -0.0,
            0.1 + 0.2, 1234565.0, 12.34565, 0.3333335, 2.5e-5, 1.0000005, 1.00000049999, 322.1215, -716.9895]
    vals += list(-np.exp(rng.uniform(-12, 14, 60000)))               # likelihood-like magnitudes
    vals += list(rng.uniform(0, 1000, 30000)) + list(rng.integers(0, 40, 5000) / rng.integers(1, 40, 5000))
    vals += [round(float(x), 5) + 5e-7 for x in rng.uniform(0, 100, 5000)]   # near decimal ties
    v = np.ascontiguousarray(np.array(vals, np.float64))
    n = len(v)
    out = np.zeros(48 * n, np.uint8)
    ln = np.zeros(n, np.uint8)
    ref = np.zeros(n, np.uint8)
    rc = L.drprg_cuda_format_g6_device(C.c_int(0), v.ctypes.data_as(C.c_void_p), C.c_uint32(n), out.ctypes.data_as(C.c_void_p),
                                       ln.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p))
    assert rc == 0, L.drprg_cuda_last_error()
    answered = 0
    for i in range(n):
        want = "%g" % float(v[i])
        if ref[i]:
            # declined: only what the fast path cannot print (exponent form, non-finite, -0, or within 1e-6 of a tie)
            a = abs(float(v[i]))
            near_tie = np.isfinite(a) and a > 0 and 1e-4 <= a < 999999.0 and \
                abs((a * 10 ** (5 - int(np.floor(np.log10(a))))) % 1 - 0.5) < 1e-4
            assert "e" in want or not np.isfinite(a) or want == "-0" or near_tie or a >= 999999.0, (v[i], want)
        else:
            got = bytes(out[48 * i:48 * i + int(ln[i])]).decode()
            assert got == want, (float(v[i]), got, want)
            answered += 1
    assert answered > 0.9 * n
