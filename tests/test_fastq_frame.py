"""Host half of the file ingest (drprg_b200/csrc/fastq_frame.cpp) without a GPU: the framer that cuts a plain FASTQ into
slices at record starts and keeps only the sequence lines must return exactly the reads the general host parser
(drprg_cuda_read_fastx) returns, or decline.  The reads file is the one of /root/reference/src/predict.rs:166-170."""
import os

import numpy as np
import pytest

os.environ.setdefault("DRPRG_FRAME_SLICE", "4096")  # read once by the library: many slices even for small files

from drprg_b200 import lib, sim


def records(n, seed, lens=None, at_quals=True):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(lens[i]) if lens is not None else int(rng.integers(0, 400))
        seq = bytes(sim.BASES[rng.integers(0, 4, size=L)])
        qual = bytes(rng.choice(np.frombuffer(b"@+FI:#", np.uint8), size=L)) if at_quals else b"F" * L
        out.append((b"@r%d some comment" % i, seq, qual))
    return out


def write(path, recs, eol=b"\n", final_newline=True, plus_repeat=False):
    parts = []
    for h, s, q in recs:
        parts += [h, eol, s, eol, b"+" + (h[1:] if plus_repeat else b""), eol, q, eol]
    text = b"".join(parts)
    if not final_newline:
        text = text[:-len(eol)]
    with open(path, "wb") as f:
        f.write(text)


def check(path, recs, threads=8):
    got = lib.frame_fastq(path, threads=threads)
    assert got is not None
    a, l = got
    assert l.tolist() == [len(s) for _, s, _ in recs]
    assert bytes(a) == b"".join(s for _, s, _ in recs)
    # and the general host parser sees the same reads
    w, wo, hl, n, tb, fl = lib.read_fastx(path, threads=threads)
    assert n == len(recs) and tb == int(l.sum())
    data = np.frombuffer(bytes(a), np.uint8)
    off = np.concatenate([[0], np.cumsum(l.astype(np.uint64))]).astype(np.uint64)
    w2, o2, l2 = lib.pack_reads(data, off)
    assert (hl == l2).all() and (wo == o2).all() and (w == w2).all()


@pytest.mark.parametrize("eol,final_newline,plus_repeat", [(b"\n", True, False), (b"\r\n", True, False), (b"\n", False, False),
                                                            (b"\r\n", False, True)])
def test_framer_matches_host_parser(tmp_path, eol, final_newline, plus_repeat):
    """ragged lengths (incl. empty reads), quality lines starting with '@' and '+', CRLF, missing final newline,
    slice boundaries everywhere (4 KB slices)"""
    recs = records(3000, 11)
    p = tmp_path / "r.fq"
    write(p, recs, eol, final_newline, plus_repeat)
    check(p, recs)
    check(p, recs, threads=1)


def test_fixed_length_and_long_reads(tmp_path):
    recs = records(5000, 5, lens=[150] * 5000)
    p = tmp_path / "a.fq"
    write(p, recs)
    check(p, recs)
    rng = np.random.default_rng(2)
    recs = records(40, 6, lens=rng.integers(20_000, 900_000, size=40))  # lines longer than a slice and than the read chunk
    p = tmp_path / "long.fq"
    write(p, recs)
    check(p, recs)


def test_declines_what_is_not_strict_fastq(tmp_path):
    recs = records(500, 3, lens=[200] * 500, at_quals=False)
    wrapped = tmp_path / "wrapped.fq"
    with open(wrapped, "wb") as f:
        for h, s, q in recs:
            f.write(h + b"\n" + s[:100] + b"\n" + s[100:] + b"\n+\n" + q[:100] + b"\n" + q[100:] + b"\n")
    assert lib.frame_fastq(wrapped) is None
    blank = tmp_path / "blank.fq"
    write(blank, recs)
    text = blank.read_bytes()
    cut = text.index(b"\n@r250 ") + 1
    blank.write_bytes(text[:cut] + b"\n" + text[cut:])
    assert lib.frame_fastq(blank) is None
    damaged = tmp_path / "damaged.fq"
    damaged.write_bytes(text[:cut] + text[cut + 37:])   # a record loses the start of its header
    assert lib.frame_fastq(damaged) is None
    fasta = tmp_path / "r.fa"
    fasta.write_bytes(b">a\nACGT\n>b\nGGCC\n")
    assert lib.frame_fastq(fasta) is None
