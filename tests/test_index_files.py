"""the `pandora index` replacement.  drprg only checks that `<prg>.k{K}.w{W}.idx` (found by extension,
/root/reference/src/lib.rs:1222-1231) and `kmer_prgs/` (src/predict.rs:400-418) exist; the files are written in pandora's text
layout and must round-trip to exactly the index the loader holds (the same index the oracle builds)."""
import os
import re
import shutil

import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import lib
from helpers import TOY_PRG, small_panel

PATH_RE = re.compile(r"(\d+)\{((?:\[\d+, \d+\))*)\}")
IV_RE = re.compile(r"\[(\d+), (\d+)\)")


def parse_path(txt):
    m = PATH_RE.fullmatch(txt)
    assert m, txt
    ivs = [(int(a), int(b)) for a, b in IV_RE.findall(m.group(2))]
    assert len(ivs) == int(m.group(1))
    return ivs


@pytest.mark.parametrize("which,w", [("toy", 11), ("toy", 14), ("panel", 11)])
def test_written_index_files_round_trip(tmp_path, which, w):
    src = TOY_PRG if which == "toy" else small_panel()[1]
    prg = tmp_path / "dr.prg"
    shutil.copy(src, prg)
    gx = lib.Index(prg, w, 15, device=-1)
    gx.write_pandora_index(prg)
    idx = tmp_path / f"dr.prg.k15.w{w}.idx"
    assert idx.exists() and (tmp_path / "kmer_prgs").is_dir()
    # what drprg's find_prg_index_in does: any file with the extension "idx" in the index directory
    assert [f for f in os.listdir(tmp_path) if f.endswith(".idx")] == [idx.name]
    ox = O.Index(src, w, 15)
    okn, rec = ox.knodes(), ox.records()
    io = np.concatenate([[0], np.cumsum(okn["n_iv"])]).astype(int)
    base = ox.knode_base
    # ---- .idx: hash -> records (prg, path, knode, strand)
    lines = idx.read_text().splitlines()
    assert int(lines[0]) == len(np.unique(rec["hash"])) == len(lines) - 1
    got = []
    for line in lines[1:]:
        f = line.split("\t")
        assert int(f[1]) == len(f) - 2
        for r in f[2:]:
            m = re.fullmatch(r"\((\d+), (.+), (\d+), ([01])\)", r)
            prg_id, path, kn, strand = int(m.group(1)), parse_path(m.group(2)), int(m.group(3)), int(m.group(4))
            g = int(base[prg_id]) + kn
            want = [(int(s), int(s + l)) for s, l in zip(okn["iv_start"][io[g]:io[g + 1]], okn["iv_len"][io[g]:io[g + 1]])]
            assert path == want
            got.append((int(f[0]), prg_id, kn, strand))
    want = sorted(zip(rec["hash"].tolist(), rec["prg"].tolist(), rec["knode"].tolist(), rec["strand"].tolist()))
    assert sorted(got) == want
    # ---- kmer_prgs/01/<locus>.k15.w{w}.gfa: nodes with their paths, edges
    eo = np.concatenate([[0], np.cumsum(okn["n_out"])]).astype(int)
    for l, name in enumerate(ox.names):
        gfa = tmp_path / "kmer_prgs" / "01" / f"{name}.k15.w{w}.gfa"
        assert gfa.exists(), gfa
        rows = gfa.read_text().splitlines()
        assert rows[0].startswith("H\tVN:Z:1.0")
        n = int(base[l + 1] - base[l])
        seg = [r.split("\t") for r in rows if r.startswith("S\t")]
        assert [int(s[1]) for s in seg] == list(range(n))
        for s in seg[1:-1]:
            g = int(base[l]) + int(s[1])
            assert parse_path(s[2]) == [(int(a), int(a + b)) for a, b in zip(okn["iv_start"][io[g]:io[g + 1]], okn["iv_len"][io[g]:io[g + 1]])]
        assert parse_path(seg[0][2]) == [(0, 0)]
        edges = sorted((int(r.split("\t")[1]), int(r.split("\t")[3])) for r in rows if r.startswith("L\t"))
        want_e = sorted((r_, int(t) - int(base[l])) for r_ in range(n) for t in okn["edges"][eo[int(base[l]) + r_]:eo[int(base[l]) + r_ + 1]])
        assert edges == want_e


def test_pandora_cuda_index_subcommand(tmp_path):
    """`pandora_cuda index -t N -w W -k K <prg>` — the argv drprg builds at src/predict.rs:283 (and src/builder.rs:644-657
    through Pandora::index_with, src/lib.rs:479-510) — writes the same files as the library call, without a GPU"""
    import filecmp
    import subprocess
    a, b = tmp_path / "a", tmp_path / "b"
    for d in (a, b):
        d.mkdir()
        shutil.copy(TOY_PRG, d / "dr.prg")
    exe = os.path.join(os.path.dirname(lib.SO_PATH), "pandora_cuda")
    r = subprocess.run([exe, "index", "-t", "2", "-w", "11", "-k", "15", str(a / "dr.prg")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    gx = lib.Index(b / "dr.prg", 11, 15, device=-1)
    gx.write_pandora_index(b / "dr.prg")
    assert filecmp.cmp(a / "dr.prg.k15.w11.idx", b / "dr.prg.k15.w11.idx", shallow=False)
    for name in ("gid", "pncA"):
        assert filecmp.cmp(a / "kmer_prgs" / "01" / f"{name}.k15.w11.gfa", b / "kmer_prgs" / "01" / f"{name}.k15.w11.gfa", shallow=False)
    bad = subprocess.run([exe, "index", "-w", "11", "-k", "15", str(tmp_path / "missing.prg")], capture_output=True, text=True)
    assert bad.returncode != 0 and "cannot open" in bad.stderr


def test_pandora_mirror_index_with(tmp_path):
    """drprg_b200.pandora.Pandora.index_with mirrors Pandora::index_with (src/lib.rs:479-510) with the argv of src/predict.rs:283"""
    from drprg_b200.pandora import DependencyError, Pandora
    prg = tmp_path / "dr.prg"
    shutil.copy(TOY_PRG, prg)
    pan = Pandora(device=-1)
    pan.index_with(prg, ["-t", "2", "-w", "14", "-k", "15"])
    assert (tmp_path / "dr.prg.k15.w14.idx").exists() and (tmp_path / "kmer_prgs" / "01" / "gid.k15.w14.gfa").exists()
    with pytest.raises(DependencyError):
        pan.index_with(tmp_path / "missing.prg", ["-w", "14", "-k", "15"])
    with pytest.raises(DependencyError):
        pan.index_with(prg, ["--bogus"])
