"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the
header declares, the host index builder agrees with the oracle, the packer round-trips, and compute
entry points refuse to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import lib, sim
from helpers import TOY_PRG, TOY_REFS, small_panel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "drprg_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(drprg_cuda_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(lib.SYMBOLS)
    L = lib.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.drprg_cuda_version() == 100


@pytest.mark.parametrize("w,k", [(11, 15), (14, 15), (5, 7), (20, 11), (1, 15), (11, 16)])
def test_host_index_builder_matches_oracle_toy(w, k):
    gx = lib.Index(TOY_PRG, w, k, device=-1)
    ox = O.Index(TOY_PRG, w, k)
    assert gx.names == ox.names and gx.total_knodes == ox.total_knodes
    assert (gx.knode_base == ox.knode_base).all()
    a, b = gx.knodes(), ox.knodes()
    for key in ("n_out", "n_iv", "edges", "iv_start", "iv_len"):
        assert (a[key] == b[key]).all(), key
    inner = b["hash"] != np.uint64(2 ** 64 - 1)
    assert (a["hash"][inner] == b["hash"][inner]).all() and (a["strand"][inner] == b["strand"][inner]).all()
    ra, rb = gx.records(), ox.records()
    for key in ra:
        assert (ra[key] == rb[key]).all(), key
    assert gx.min_path_lengths().tolist() == [ox.min_path_length(i) for i in range(ox.n_loci)]


def test_host_index_builder_matches_oracle_panel():
    p, prg, refs = small_panel()
    gx = lib.Index(prg, 11, 15, device=-1)
    ox = O.Index(prg, 11, 15)
    a, b = gx.knodes(), ox.knodes()
    for key in ("n_out", "n_iv", "edges", "iv_start", "iv_len"):
        assert (a[key] == b[key]).all(), key
    ra, rb = gx.records(), ox.records()
    for key in ra:
        assert (ra[key] == rb[key]).all(), key


def test_every_linear_minimizer_of_a_prg_path_is_indexed():
    """SURVEY A.4 property: sketching any full path of the PRG as a linear sequence only yields
    k-mers the graph sketch indexed (so reads from any haplotype of the panel find their k-mer nodes)."""
    p, prg, refs = small_panel()
    ox = O.Index(prg, 11, 15)
    rec = ox.records()
    for li, locus in enumerate(p.loci):
        have = set(rec["hash"][rec["prg"] == li].tolist())
        rng = np.random.default_rng(li)
        for trial in range(6):
            seq = locus.spell(lambda site: int(rng.integers(0, len(site.alleles))))
            h, s, d = O.sketch(seq, 11, 15)
            missing = [int(x) for x in h if int(x) not in have]
            assert not missing, (locus.name, trial, len(missing))


def test_packer_layout_and_flags():
    strs = ["ACGT" * 5, "T" * 33, "ACGNACGT", "", "acgtACGT"]
    data = np.frombuffer("".join(strs).encode(), np.uint8).copy()
    off = np.zeros(len(strs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in strs])
    words, woff, lens = lib.pack_reads(data, off)
    w2, o2, l2 = sim.pack_reads(data, off)
    assert (words == w2).all() and (woff == o2).all() and (lens == l2).all()
    assert lens.tolist() == [20, 33, 0, 0, 8]
    assert words[0] == int("00011011" * 4, 2)        # ACGT ACGT ACGT ACGT, first base in the top bits
    assert words[int(woff[1])] == 0xFFFFFFFF         # 16 T
    words, woff, lens = lib.pack_reads(data, off, stride_words=3)
    assert len(words) == 15 and words[3] == 0xFFFFFFFF and words[5] == 0xC0000000
    with pytest.raises(lib.DrprgCudaError):
        lib.pack_reads(data, off, stride_words=2)     # 33 bases need 3 words


def test_compute_refuses_without_gpu():
    gx = lib.Index(TOY_PRG, 11, 15, device=-1)
    with pytest.raises(lib.DrprgCudaError, match="no CPU fallback"):
        gx.sample_begin(lib.make_opts(), 150)
    with pytest.raises(lib.DrprgCudaError):
        gx.map_genotype("x.fq", TOY_REFS, ".")
    with pytest.raises(lib.DrprgCudaError, match="k > 16|no CUDA device"):
        lib.Index(TOY_PRG, 11, 17, device=0)


def test_vcf_float_formatting_matches_printf_g():
    """the VCF writer's fast "%g" must be byte-identical to printf (pandora prints floats with ostream defaults)"""
    import ctypes as C
    L = lib.lib()
    L.drprg_cuda_format_g6.argtypes = [C.c_double, C.c_char_p]
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(0)
    vals = [0.0, 1.0, -1.0, 0.5, 0.25, 0.125, 0.2, 1 / 3, 2 / 3, 0.666667, 1e-4, 9.99999e-5, 123456.5, 999999.4, 999999.5,
            999999.6, 1e6, -144.0, -0.0001, 100000.0, 99999.95, 0.000123456789, 5e-324, 1e300, float("inf"), 0.1 + 0.2,
            1234565.0, 12.34565, 0.3333335, 2.5e-5, 1.0000005, 1.00000049999, 322.1215, -716.9895]
    vals += list(-np.exp(rng.uniform(-12, 14, 60000)))               # likelihood-like magnitudes
    vals += list(rng.uniform(0, 1000, 30000)) + list(rng.integers(0, 40, 5000) / rng.integers(1, 40, 5000))
    vals += [round(float(x), 5) + 5e-7 for x in rng.uniform(0, 100, 5000)]   # near decimal ties
    for v in vals:
        n = L.drprg_cuda_format_g6(float(v), buf)
        assert buf.value[:n].decode() == "%g" % float(v), (v, buf.value, "%g" % float(v))


@pytest.mark.parametrize("fmt,gz,threads", [("fq", False, 1), ("fq", True, 4), ("fa", False, 3), ("fa", True, 1), ("fq", False, 8)])
def test_read_file_ingest_matches_packer(tmp_path, fmt, gz, threads):
    """fasta/fastq, plain or gzip (src/predict.rs:166-170), any host thread count: same packed reads as the in-memory packer"""
    import ctypes as C
    import gzip
    rng = np.random.default_rng(5)
    strs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(L))) for L in rng.integers(1, 400, size=3000)]
    strs[7] = strs[7][:10] + "N" + strs[7][11:]
    strs[100] = "@" + "ACGT" * 3 if False else strs[100]
    path = tmp_path / ("r." + fmt + (".gz" if gz else ""))
    op = gzip.open if gz else open
    with op(path, "wt") as f:
        for i, s in enumerate(strs):
            if fmt == "fq":
                q = "".join(chr(33 + int(x)) for x in rng.integers(0, 41, size=len(s)))  # qualities include '@' and '+'
                f.write(f"@r{i} desc\n{s}\n+\n{q}\n")
            else:
                f.write(f">r{i}\n" + "\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + "\n")
    L = lib.lib()
    words, woff, lens = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)()
    n, tb, fl = C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = L.drprg_cuda_read_fastx(str(path).encode(), threads, C.byref(words), C.byref(woff), C.byref(lens), C.byref(n), C.byref(tb), C.byref(fl))
    assert rc == 0, L.drprg_cuda_last_error()
    data = np.frombuffer("".join(strs).encode(), np.uint8)
    off = np.zeros(len(strs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in strs])
    w2, o2, l2 = lib.pack_reads(data, off)
    assert n.value == len(strs) and tb.value == int(off[-1]) and fl.value == len(strs[0])
    assert (np.ctypeslib.as_array(lens, (n.value,)) == l2).all()
    assert (np.ctypeslib.as_array(woff, (n.value + 1,)) == o2).all()
    assert (np.ctypeslib.as_array(words, (len(w2),)) == w2).all()
    for p in (words, woff, lens):
        L.drprg_cuda_host_free(p)


def test_wrapped_fastq_is_parsed_like_kseq(tmp_path):
    """multi-line (wrapped) FASTQ, which pandora's kseq-based reader accepts: sequence lines up to '+', quality lines
    until they are as long as the sequence — including quality lines that start with '@' or '+'"""
    rng = np.random.default_rng(11)
    strs = ["".join("ACGT"[i] for i in rng.integers(0, 4, size=int(L))) for L in rng.integers(1, 500, size=400)]
    strs[3] = strs[3][:5] + "N" + strs[3][6:]
    path = tmp_path / "wrapped.fq"
    with open(path, "w") as f:
        for i, s_ in enumerate(strs):
            q = "".join(chr(33 + int(x)) for x in rng.integers(0, 41, size=len(s_)))
            if i % 7 == 0:
                q = "@" + q[1:]
            if i % 11 == 0:
                q = "+" + q[1:]
            wrap = 60 if i % 3 else 10 ** 9   # every third record stays on one line
            f.write(f"@r{i}\n" + "\n".join(s_[j:j + wrap] for j in range(0, len(s_), wrap)) + "\n+\n")
            f.write("\n".join(q[j:j + wrap] for j in range(0, len(q), wrap)) + "\n")
    for threads in (1, 8):
        words, woff, lens, n, tb, fl = lib.read_fastx(path, threads=threads)
        data = np.frombuffer("".join(strs).encode(), np.uint8)
        off = np.zeros(len(strs) + 1, np.uint64)
        off[1:] = np.cumsum([len(s_) for s_ in strs])
        w2, o2, l2 = lib.pack_reads(data, off)
        assert n == len(strs) and tb == int(off[-1]) and fl == len(strs[0])
        assert (lens == l2).all() and (woff == o2).all() and (words == w2).all()


def test_hash64_is_inverted_exactly():
    """the k-mer screen is built from hash64's inverse: inverse(hash(x)) == x and hash(inverse(h)) == h for every k, and
    the host hash agrees with the oracle's"""
    import oracle_py as O
    L = lib.lib()
    rng = np.random.default_rng(5)
    for k in (3, 7, 11, 15, 16, 21, 31):
        mask = (1 << (2 * k)) - 1
        for x in [0, 1, mask, mask >> 1] + [int(v) & mask for v in rng.integers(0, 2 ** 62, size=300)]:
            h = L.drprg_cuda_hash64(x, k)
            assert h <= mask
            assert L.drprg_cuda_hash64_inverse(h, k) == x
            assert L.drprg_cuda_hash64(L.drprg_cuda_hash64_inverse(x, k), k) == x
    # against the oracle's sketch: a read of exactly w + k - 1 = k bases with w = 1 has one minimizer = min(hash(fwd), hash(rc))
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for s_ in ("ACGTACGTACGTACG", "TTTTTGGGGGCCCCC", "GATTACAGATTACAG"):
        fwd = 0
        for c in s_:
            fwd = (fwd << 2) | code[c]
        rc = 0
        for c in reversed(s_):
            rc = (rc << 2) | (3 - code[c])
        hs, st, sd = O.sketch(s_.encode(), 1, 15)
        assert len(hs) == 1 and int(hs[0]) == min(L.drprg_cuda_hash64(fwd, 15), L.drprg_cuda_hash64(rc, 15))
