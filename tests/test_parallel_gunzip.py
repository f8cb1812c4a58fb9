"""The hand-written parallel gzip inflater (drprg_b200/csrc/gzip_inflate.cpp) against zlib: it must either reproduce
the text byte for byte or decline (the reader then falls back to zlib), never return something else.  Exercised through
the public host parser drprg_cuda_read_fastx on FASTQ text compressed in different ways."""
import gzip
import os
import zlib

import numpy as np
import pytest

os.environ.setdefault("DRPRG_PARALLEL_GZIP_CHUNK", "65536")  # read once by the library: cut even small files into chunks
os.environ.setdefault("DRPRG_PARALLEL_GZIP_GROUP", "5")      # ... and decode them in groups of five (large files: 4 x threads)

from drprg_b200 import lib, sim


def fastq_text(n, L, seed, realistic_quals=True):
    rng = np.random.default_rng(seed)
    d = sim.BASES[rng.integers(0, 4, size=n * L)]
    o = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    return d, o


def check(path, d, o, threads=8):
    w, wo, l, n, tb, fl = lib.read_fastx(path, threads=threads)
    w2, o2, l2 = lib.pack_reads(d, o)
    assert n == len(o) - 1 and tb == int(o[-1])
    assert (l == l2).all() and (wo == o2).all() and (w == w2).all()


@pytest.mark.parametrize("level", [1, 6, 9])
def test_single_stream_levels(tmp_path, level):
    d, o = fastq_text(60_000, 150, level)
    raw = tmp_path / "r.fq"
    sim.write_fastq_fast(str(raw), d, o)
    text = raw.read_bytes()
    co = zlib.compressobj(level, zlib.DEFLATED, 31)
    gz = tmp_path / f"r{level}.fq.gz"
    gz.write_bytes(co.compress(text) + co.flush())
    check(gz, d, o)
    check(gz, d, o, threads=3)


def test_multi_member_and_header_fields(tmp_path):
    """concatenated members (cat a.gz b.gz) and optional header fields (file name): result identical to zlib's"""
    d, o = fastq_text(40_000, 150, 3)
    raw = tmp_path / "r.fq"
    sim.write_fastq_fast(str(raw), d, o)
    text = raw.read_bytes()
    half = text.rfind(b"\n@", 0, len(text) // 2) + 1
    gz = tmp_path / "multi.fq.gz"
    with open(gz, "wb") as f:
        f.write(gzip.compress(text[:half], 1))
        f.write(gzip.compress(text[half:], 1))
    check(gz, d, o)
    gz2 = tmp_path / "named.fq.gz"
    with gzip.GzipFile(filename="reads_with_a_name.fq", mode="wb", fileobj=open(gz2, "wb"), compresslevel=1, mtime=12345) as f:
        f.write(text)
    check(gz2, d, o)


def test_low_entropy_text_long_matches(tmp_path):
    """homopolymer reads and constant qualities: 258-byte matches, overlapping copies (distance < length), tiny alphabets"""
    n, L = 80_000, 150
    d = np.tile(np.frombuffer(b"ACGT" * 37 + b"AC", np.uint8), n).copy()
    d[::7] = ord("T")
    o = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    raw = tmp_path / "r.fq"
    sim.write_fastq(str(raw), d, o)       # qualities 'I' * L
    text = raw.read_bytes() * 4           # ~100 MB of highly repetitive text -> a few MB compressed
    d4, o4 = np.tile(d, 4), np.arange(4 * n + 1, dtype=np.uint64) * np.uint64(L)
    gz = tmp_path / "rep.fq.gz"
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    gz.write_bytes(co.compress(text) + co.flush())
    check(gz, d4, o4)


def test_corrupt_stream_is_an_error_not_garbage(tmp_path):
    d, o = fastq_text(60_000, 150, 5)
    raw = tmp_path / "r.fq"
    sim.write_fastq_fast(str(raw), d, o)
    co = zlib.compressobj(1, zlib.DEFLATED, 31)
    blob = bytearray(co.compress(raw.read_bytes()) + co.flush())
    blob[len(blob) // 2] ^= 0x5A
    bad = tmp_path / "bad.fq.gz"
    bad.write_bytes(bytes(blob))
    with pytest.raises(lib.DrprgCudaError):
        lib.read_fastx(bad, threads=8)


def test_bgzf_blocks_and_many_members(tmp_path):
    """bgzip-style input (every <= 64 KB of text its own gzip member, plus the empty EOF member) and `cat a.gz b.gz c.gz`
    of members of different sizes and levels: chunks start at member headers, members end inside chunks, every member's
    CRC-32 and length is verified"""
    d, o = fastq_text(70_000, 150, 9)
    raw = tmp_path / "r.fq"
    sim.write_fastq_fast(str(raw), d, o)
    text = raw.read_bytes()
    bg = tmp_path / "r.fq.bgz.gz"
    with open(bg, "wb") as f:
        for i in range(0, len(text), 65280):
            f.write(gzip.compress(text[i:i + 65280], 6, mtime=0))
        f.write(gzip.compress(b"", 6, mtime=0))
    check(bg, d, o)
    check(bg, d, o, threads=3)
    cuts = [0, len(text) // 7, len(text) // 7 + 100, len(text) // 2, len(text)]
    cat = tmp_path / "cat.fq.gz"
    with open(cat, "wb") as f:
        for i, (a, b) in enumerate(zip(cuts, cuts[1:])):
            f.write(gzip.compress(text[a:b], (1, 9, 6, 1)[i], mtime=i))
    check(cat, d, o)
    # one member's CRC damaged: never garbage
    blob = bytearray(cat.read_bytes())
    first_len = len(gzip.compress(text[cuts[0]:cuts[1]], 1, mtime=0))
    blob[first_len - 8] ^= 0x01  # CRC field of the first member
    bad = tmp_path / "badcrc.fq.gz"
    bad.write_bytes(bytes(blob))
    with pytest.raises(lib.DrprgCudaError):
        lib.read_fastx(bad, threads=8)
