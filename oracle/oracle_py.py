"""ORACLE ctypes binding — TEST INFRASTRUCTURE ONLY (see oracle/oracle.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  The product package (drprg_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Opts(C.Structure):
    """Mirror of drprg_map_opts (include/drprg_cuda.h)."""

    _fields_ = [
        ("threads", C.c_uint32),
        ("min_cluster_size", C.c_uint32),
        ("illumina", C.c_uint8),
        ("debug", C.c_uint8),
        ("genome_size", C.c_uint32),
        ("max_covg", C.c_uint32),
        ("gt_conf", C.c_double),
        ("genotyping_error_rate", C.c_double),
        ("max_diff", C.c_uint32),
        ("error_rate", C.c_double),
    ]


def make_opts(threads=1, min_cluster_size=10, illumina=False, genome_size=4411532, gt_conf=0.0,
              genotyping_error_rate=0.01, max_diff=0, error_rate=0.0):
    return Opts(threads, min_cluster_size, int(illumina), 0, genome_size, 0xFFFFFFFF, gt_conf,
                genotyping_error_rate, max_diff, error_rate)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("prg.cpp", "mapper.cpp", "vcf.cpp", "capi.cpp", "oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_last_error.restype = C.c_char_p
        for f in ("orc_index_build", "orc_index_build_text", "orc_map", "orc_genotype", "orc_map_from_coverage"):
            getattr(L, f).restype = C.c_void_p
        L.orc_locus_name.restype = C.c_char_p
        L.orc_gt_vcf.restype = C.c_char_p
        L.orc_num_records.restype = C.c_uint64
        L.orc_map_num_hits.restype = C.c_uint64
        L.orc_sketch.restype = C.c_int64
        L.orc_gt_mlpath.restype = C.c_int64
        L.orc_allele_likelihood.restype = C.c_double
        L.orc_allele_likelihood.argtypes = [C.c_double] * 5
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Index:
    def __init__(self, prg_path=None, w=11, k=15, text=None):
        L = lib()
        if text is not None:
            self.h = L.orc_index_build_text(text.encode(), w, k)
        else:
            self.h = L.orc_index_build(str(prg_path).encode(), w, k)
        if not self.h:
            raise RuntimeError(L.orc_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.w, self.k = w, k
        self.n_loci = L.orc_num_loci(self.h)
        self.names = [L.orc_locus_name(self.h, i).decode() for i in range(self.n_loci)]
        self.total_knodes = L.orc_total_knodes(self.h)
        self.knode_base = np.zeros(self.n_loci + 1, np.uint32)
        L.orc_knode_base(self.h, _p(self.knode_base))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    def knodes(self):
        L, n = lib(), self.total_knodes
        hsh = np.zeros(n, np.uint64); strand = np.zeros(n, np.uint8)
        n_out = np.zeros(n, np.uint32); n_iv = np.zeros(n, np.uint32)
        L.orc_knode_info(self.h, _p(hsh), _p(strand), _p(n_out), _p(n_iv))
        edges = np.zeros(int(n_out.sum()), np.uint32)
        L.orc_knode_edges(self.h, _p(edges))
        ivs = np.zeros(int(n_iv.sum()), np.uint32); ivl = np.zeros(int(n_iv.sum()), np.uint32)
        L.orc_knode_paths(self.h, _p(ivs), _p(ivl))
        return dict(hash=hsh, strand=strand, n_out=n_out, n_iv=n_iv, edges=edges, iv_start=ivs, iv_len=ivl)

    def records(self):
        L = lib()
        n = L.orc_num_records(self.h)
        hsh = np.zeros(n, np.uint64); prg = np.zeros(n, np.uint32); kn = np.zeros(n, np.uint32); st = np.zeros(n, np.uint8)
        L.orc_records(self.h, _p(hsh), _p(prg), _p(kn), _p(st))
        return dict(hash=hsh, prg=prg, knode=kn, strand=st)

    def min_path_length(self, locus):
        return lib().orc_min_path_length(self.h, locus)

    def local_graph(self, locus):
        L = lib()
        n = L.orc_num_local_nodes(self.h, locus)
        st = np.zeros(n, np.uint32); ln = np.zeros(n, np.uint32); no = np.zeros(n, np.uint32)
        L.orc_local_nodes(self.h, locus, _p(st), _p(ln), _p(no))
        e = np.zeros(int(no.sum()), np.uint32)
        L.orc_local_edges(self.h, locus, _p(e))
        return dict(start=st, len=ln, n_out=no, edges=e)


def sketch(seq, w, k):
    L = lib()
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    cap = max(16, len(b))
    hsh = np.zeros(cap, np.uint64); st = np.zeros(cap, np.uint32); sd = np.zeros(cap, np.uint8)
    n = L.orc_sketch(b, C.c_uint64(len(b)), w, k, _p(hsh), _p(st), _p(sd), C.c_uint64(cap))
    return hsh[:n].copy(), st[:n].copy(), sd[:n].copy()


class MapRun:
    """S1-S5 on ASCII reads: data = uint8 array of concatenated bases, off = uint64 offsets (n+1)."""

    def __init__(self, index, data=None, off=None, opts=None, first_len_hint=0, _h=None):
        L = lib()
        self.index = index
        if _h is not None:
            self.h = C.c_void_p(_h)
            return
        data = np.ascontiguousarray(data, np.uint8)
        off = np.ascontiguousarray(off, np.uint64)
        self.opts = opts or make_opts()
        h = L.orc_map(index.h, _p(data), _p(off), C.c_uint64(len(off) - 1), C.byref(self.opts), first_len_hint)
        if not h:
            raise RuntimeError(L.orc_last_error().decode())
        self.h = C.c_void_p(h)

    @classmethod
    def from_coverage(cls, index, fwd, rev, locus_reads, total_bases, n_reads):
        fwd = np.ascontiguousarray(fwd, np.uint32); rev = np.ascontiguousarray(rev, np.uint32)
        lr = np.ascontiguousarray(locus_reads, np.uint32)
        h = lib().orc_map_from_coverage(index.h, _p(fwd), _p(rev), _p(lr), C.c_uint64(int(total_bases)), C.c_uint64(int(n_reads)))
        return cls(index, _h=h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_free(self.h)
            self.h = None

    def hits(self):
        L = lib()
        n = L.orc_map_num_hits(self.h)
        a = {k: np.zeros(n, np.uint32) for k in ("read", "start", "prg", "knode")}
        fwd = np.zeros(n, np.uint8); kept = np.zeros(n, np.uint8)
        L.orc_map_hits(self.h, _p(a["read"]), _p(a["start"]), _p(a["prg"]), _p(a["knode"]), _p(fwd), _p(kept))
        a["fwd"], a["kept"] = fwd, kept
        return a

    def coverage(self):
        n = self.index.total_knodes
        f = np.zeros(n, np.uint32); r = np.zeros(n, np.uint32)
        lib().orc_map_coverage(self.h, _p(f), _p(r))
        return f, r

    def locus_reads(self):
        o = np.zeros(self.index.n_loci, np.uint32)
        lib().orc_map_locus_reads(self.h, _p(o))
        return o

    def scalars(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().orc_map_scalars(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(total_bases=a.value, n_minimizers=b.value, n_reads=c.value)


class Genotype:
    def __init__(self, index, maprun, opts=None, vcf_refs=None, sample="sample"):
        L = lib()
        self.index = index
        self.opts = opts or make_opts()
        h = L.orc_genotype(index.h, maprun.h, C.byref(self.opts), (str(vcf_refs).encode() if vcf_refs else None), sample.encode())
        if not h:
            raise RuntimeError(L.orc_last_error().decode())
        self.h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_gt_free(self.h)
            self.h = None

    def vcf(self):
        return lib().orc_gt_vcf(self.h).decode()

    def params(self):
        o = np.zeros(11, np.float64)
        lib().orc_gt_params(self.h, _p(o))
        keys = ["E", "bin", "nb_p", "nb_r", "e_rate", "thresh", "covg", "min_kmer_covg", "mean", "var", "num_reads"]
        return dict(zip(keys, o.tolist()))

    def mlpath(self, locus):
        cap = int(self.index.knode_base[locus + 1] - self.index.knode_base[locus]) + 1
        o = np.zeros(cap, np.uint32)
        n = lib().orc_gt_mlpath(self.h, locus, _p(o), C.c_uint64(cap))
        return None if n < 0 else o[:n].copy()

    def records(self):
        L = lib()
        n = L.orc_gt_num_records(self.h); na = L.orc_gt_num_alleles(self.h)
        locus = np.zeros(n, np.uint32); pos = np.zeros(n, np.uint32); nal = np.zeros(n, np.uint32)
        gt = np.zeros(n, np.int32); conf = np.zeros(n, np.float64)
        L.orc_gt_records(self.h, _p(locus), _p(pos), _p(nal), _p(gt), _p(conf))
        lik = np.zeros(na, np.float64); gaps = np.zeros(na, np.float64)
        u = {k: np.zeros(na, np.uint32) for k in ("mean_fwd", "mean_rev", "med_fwd", "med_rev", "sum_fwd", "sum_rev", "n_knodes")}
        L.orc_gt_alleles(self.h, _p(lik), _p(gaps), _p(u["mean_fwd"]), _p(u["mean_rev"]), _p(u["med_fwd"]), _p(u["med_rev"]),
                         _p(u["sum_fwd"]), _p(u["sum_rev"]), _p(u["n_knodes"]))
        kn = np.zeros(int(u["n_knodes"].sum()), np.uint32)
        L.orc_gt_allele_knodes(self.h, _p(kn))
        d = dict(locus=locus, pos=pos, n_alleles=nal, gt=gt, gt_conf=conf, lik=lik, gaps=gaps, allele_knodes=kn)
        d.update(u)
        return d


def allele_likelihood(E, c, o, gaps, err=0.01):
    return lib().orc_allele_likelihood(E, c, o, gaps, err)
