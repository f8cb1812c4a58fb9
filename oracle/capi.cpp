// ORACLE (test infrastructure, see oracle.hpp): flat C interface for ctypes (tests/, smoke(),
// bench.py cpu_baseline) and the `pandora`-argv-compatible CLI main (oracle/cli.cpp).
#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>

#include "oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;

struct OrcOpts {  // mirrors drprg_map_opts in include/drprg_cuda.h
    uint32_t threads, min_cluster_size;
    uint8_t illumina, debug;
    uint32_t genome_size, max_covg;
    double gt_conf, genotyping_error_rate;
    uint32_t max_diff;
    double error_rate;
};
MapOpts to_opts(const OrcOpts* o) {
    MapOpts m;
    if (!o) return m;
    m.threads = o->threads ? (int)o->threads : 1;
    m.min_cluster_size = o->min_cluster_size;
    m.illumina = o->illumina != 0;
    if (o->genome_size) m.genome_size = o->genome_size;
    m.gt_conf = o->gt_conf;
    if (o->genotyping_error_rate > 0) m.gt_error_rate = o->genotyping_error_rate;
    if (o->max_diff) m.max_diff = o->max_diff;
    if (o->error_rate > 0) m.e_rate = o->error_rate;
    return m;
}
struct MapHandle {
    MapResult mr;
};
struct GtHandle {
    GenotypeResult g;
    std::vector<uint32_t> rec_locus;
};
template <class F>
auto guard(F f) -> decltype(f()) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_err = e.what();
        return decltype(f())();
    }
}
}  // namespace

extern "C" {
const char* orc_last_error() { return g_err.c_str(); }

void* orc_index_build(const char* prg_path, uint32_t w, uint32_t k) {
    return guard([&]() -> void* { return build_index(prg_path, w, k).release(); });
}
void* orc_index_build_text(const char* text, uint32_t w, uint32_t k) {
    return guard([&]() -> void* { return build_index_from_text(text, w, k).release(); });
}
void orc_index_free(void* h) { delete (Index*)h; }
uint32_t orc_num_loci(void* h) { return (uint32_t)((Index*)h)->prgs.size(); }
const char* orc_locus_name(void* h, uint32_t i) { return ((Index*)h)->prgs[i].name.c_str(); }
uint32_t orc_total_knodes(void* h) { return ((Index*)h)->total_knodes; }
void orc_knode_base(void* h, uint32_t* out) {
    Index* x = (Index*)h;
    for (size_t i = 0; i < x->prgs.size(); ++i) out[i] = x->knode_base[i];
    out[x->prgs.size()] = x->total_knodes;
}
uint32_t orc_min_path_length(void* h, uint32_t l) { return ((Index*)h)->prgs[l].kg.min_path_length(); }
// per global knode (rank order): hash, strand, number of out edges, number of path intervals
void orc_knode_info(void* h, uint64_t* hash, uint8_t* strand, uint32_t* n_out, uint32_t* n_iv) {
    Index* x = (Index*)h;
    size_t g = 0;
    for (auto& p : x->prgs)
        for (uint32_t id : p.kg.sorted) {
            const KmerNode& kn = p.kg.nodes[id];
            hash[g] = kn.khash;
            strand[g] = kn.strand;
            n_out[g] = (uint32_t)kn.out.size();
            n_iv[g] = (uint32_t)kn.path.size();
            ++g;
        }
}
// out edges as global knode ids, ascending per node
void orc_knode_edges(void* h, uint32_t* edges) {
    Index* x = (Index*)h;
    size_t e = 0;
    for (size_t l = 0; l < x->prgs.size(); ++l) {
        auto& kg = x->prgs[l].kg;
        for (uint32_t id : kg.sorted) {
            std::vector<uint32_t> o;
            for (uint32_t t : kg.nodes[id].out) o.push_back(x->knode_base[l] + kg.rank[t]);
            std::sort(o.begin(), o.end());
            for (uint32_t t : o) edges[e++] = t;
        }
    }
}
void orc_knode_paths(void* h, uint32_t* iv_start, uint32_t* iv_len) {
    Index* x = (Index*)h;
    size_t e = 0;
    for (auto& p : x->prgs)
        for (uint32_t id : p.kg.sorted)
            for (auto& iv : p.kg.nodes[id].path) {
                iv_start[e] = iv.start;
                iv_len[e] = iv.length;
                ++e;
            }
}
uint64_t orc_num_records(void* h) {
    uint64_t n = 0;
    for (auto& kv : ((Index*)h)->minhash) n += kv.second.size();
    return n;
}
// index records sorted by (hash, prg, knode rank)
void orc_records(void* h, uint64_t* hash, uint32_t* prg, uint32_t* knode, uint8_t* strand) {
    Index* x = (Index*)h;
    struct R {
        uint64_t h;
        uint32_t p, k;
        uint8_t s;
    };
    std::vector<R> v;
    for (auto& kv : x->minhash)
        for (auto& r : kv.second) v.push_back({kv.first, r.prg_id, x->prgs[r.prg_id].kg.rank[r.knode_id], (uint8_t)r.strand});
    std::sort(v.begin(), v.end(), [](const R& a, const R& b) {
        if (a.h != b.h) return a.h < b.h;
        if (a.p != b.p) return a.p < b.p;
        return a.k < b.k;
    });
    for (size_t i = 0; i < v.size(); ++i) {
        hash[i] = v[i].h;
        prg[i] = v[i].p;
        knode[i] = v[i].k;
        strand[i] = v[i].s;
    }
}
uint32_t orc_num_local_nodes(void* h, uint32_t l) { return (uint32_t)((Index*)h)->prgs[l].nodes.size(); }
void orc_local_nodes(void* h, uint32_t l, uint32_t* start, uint32_t* len, uint32_t* n_out) {
    auto& p = ((Index*)h)->prgs[l];
    for (size_t i = 0; i < p.nodes.size(); ++i) {
        start[i] = p.nodes[i].pos.start;
        len[i] = p.nodes[i].pos.length;
        n_out[i] = (uint32_t)p.nodes[i].out.size();
    }
}
void orc_local_edges(void* h, uint32_t l, uint32_t* edges) {
    auto& p = ((Index*)h)->prgs[l];
    size_t e = 0;
    for (auto& n : p.nodes)
        for (uint32_t o : n.out) edges[e++] = o;
}

int64_t orc_sketch(const char* seq, uint64_t len, uint32_t w, uint32_t k, uint64_t* hash, uint32_t* start,
                   uint8_t* strand, uint64_t cap) {
    auto v = sketch_read(seq, len, w, k);
    // position order for easy comparison
    std::sort(v.begin(), v.end(), [](const Minimizer& a, const Minimizer& b) { return a.start < b.start; });
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
        hash[i] = v[i].hash;
        start[i] = v[i].start;
        strand[i] = v[i].strand;
    }
    return (int64_t)v.size();
}
void orc_hash_kmer(const char* s, uint32_t k, uint64_t* fwd, uint64_t* rev) {
    auto p = kmerhash(std::string(s, k), k);
    *fwd = p.first;
    *rev = p.second;
}

void* orc_map(void* h, const char* data, const uint64_t* off, uint64_t n, const void* opts, uint32_t first_len_hint) {
    return guard([&]() -> void* {
        auto* m = new MapHandle();
        ReadSet rs{data, off, n};
        map_reads(*(Index*)h, rs, to_opts((const OrcOpts*)opts), m->mr, first_len_hint);
        return m;
    });
}
void orc_map_free(void* m) { delete (MapHandle*)m; }
uint64_t orc_map_num_hits(void* m) { return ((MapHandle*)m)->mr.hits.size(); }
void orc_map_hits(void* m, uint32_t* read, uint32_t* start, uint32_t* prg, uint32_t* knode, uint8_t* fwd, uint8_t* kept) {
    auto& mr = ((MapHandle*)m)->mr;
    for (size_t i = 0; i < mr.hits.size(); ++i) {
        read[i] = mr.hits[i].read_id;
        start[i] = mr.hits[i].read_start;
        prg[i] = mr.hits[i].prg_id;
        knode[i] = mr.hits[i].knode;
        fwd[i] = mr.hits[i].forward;
        kept[i] = mr.kept[i];
    }
}
void orc_map_coverage(void* m, uint32_t* fwd, uint32_t* rev) {
    auto& mr = ((MapHandle*)m)->mr;
    memcpy(fwd, mr.cov_fwd.data(), mr.cov_fwd.size() * 4);
    memcpy(rev, mr.cov_rev.data(), mr.cov_rev.size() * 4);
}
void orc_map_locus_reads(void* m, uint32_t* out) {
    auto& mr = ((MapHandle*)m)->mr;
    memcpy(out, mr.locus_reads.data(), mr.locus_reads.size() * 4);
}
void orc_map_scalars(void* m, uint64_t* total_bases, uint64_t* n_minimizers, uint64_t* n_reads) {
    auto& mr = ((MapHandle*)m)->mr;
    *total_bases = mr.total_bases;
    *n_minimizers = mr.n_minimizers;
    *n_reads = mr.n_reads;
}
// build a MapResult from externally supplied coverage (e.g. summed over shards)
void* orc_map_from_coverage(void* h, const uint32_t* fwd, const uint32_t* rev, const uint32_t* locus_reads,
                            uint64_t total_bases, uint64_t n_reads) {
    Index* x = (Index*)h;
    auto* m = new MapHandle();
    m->mr.cov_fwd.assign(fwd, fwd + x->total_knodes);
    m->mr.cov_rev.assign(rev, rev + x->total_knodes);
    m->mr.locus_reads.assign(locus_reads, locus_reads + x->prgs.size());
    m->mr.total_bases = total_bases;
    m->mr.n_reads = n_reads;
    return m;
}

void* orc_genotype(void* h, void* m, const void* opts, const char* vcf_refs, const char* sample) {
    return guard([&]() -> void* {
        std::map<std::string, std::string> refs;
        if (vcf_refs && *vcf_refs) refs = read_fasta(vcf_refs);
        auto* g = new GtHandle();
        Index* x = (Index*)h;
        g->g = genotype(*x, ((MapHandle*)m)->mr, to_opts((const OrcOpts*)opts), refs, sample ? sample : "sample");
        std::map<std::string, uint32_t> byname;
        for (uint32_t l = 0; l < x->prgs.size(); ++l) byname[x->prgs[l].name] = l;
        for (auto& r : g->g.records) g->rec_locus.push_back(byname[r.chrom]);
        return g;
    });
}
void orc_gt_free(void* g) { delete (GtHandle*)g; }
const char* orc_gt_vcf(void* g) { return ((GtHandle*)g)->g.vcf_text.c_str(); }
// out[0..9] = E, bin, nb_p, nb_r, e_rate, thresh, covg, min_kmer_covg, mean, var ; out[10]=num_reads
void orc_gt_params(void* g, double* out) {
    const Params& P = ((GtHandle*)g)->g.params;
    out[0] = P.exp_depth_covg; out[1] = P.bin; out[2] = P.nb_p; out[3] = P.nb_r; out[4] = P.e_rate;
    out[5] = P.thresh; out[6] = P.covg; out[7] = P.min_kmer_covg; out[8] = P.mean; out[9] = P.var;
    out[10] = (double)P.num_reads;
}
// ML k-mer path of a locus as ranks within the locus; returns length, or -1 if the locus is absent
int64_t orc_gt_mlpath(void* g, uint32_t locus, uint32_t* out, uint64_t cap) {
    const MLPath& p = ((GtHandle*)g)->g.ml[locus];
    if (p.skipped) return -1;
    for (size_t i = 0; i < p.kpath.size() && i < cap; ++i) out[i] = p.kpath[i];
    return (int64_t)p.kpath.size();
}
uint32_t orc_gt_num_records(void* g) { return (uint32_t)((GtHandle*)g)->g.records.size(); }
uint32_t orc_gt_num_alleles(void* g) {
    uint32_t n = 0;
    for (auto& r : ((GtHandle*)g)->g.records) n += (uint32_t)r.lik.size();
    return n;
}
void orc_gt_records(void* g, uint32_t* locus, uint32_t* pos, uint32_t* n_alleles, int32_t* gt, double* gt_conf) {
    auto* G = (GtHandle*)g;
    for (size_t i = 0; i < G->g.records.size(); ++i) {
        auto& r = G->g.records[i];
        locus[i] = G->rec_locus[i];
        pos[i] = r.pos;
        n_alleles[i] = (uint32_t)r.lik.size();
        gt[i] = r.gt;
        gt_conf[i] = r.gt_conf;
    }
}
void orc_gt_alleles(void* g, double* lik, double* gaps, uint32_t* mean_fwd, uint32_t* mean_rev, uint32_t* med_fwd,
                    uint32_t* med_rev, uint32_t* sum_fwd, uint32_t* sum_rev, uint32_t* n_knodes) {
    size_t e = 0;
    for (auto& r : ((GtHandle*)g)->g.records)
        for (size_t a = 0; a < r.lik.size(); ++a) {
            lik[e] = r.lik[a]; gaps[e] = r.gaps[a];
            mean_fwd[e] = r.mean_fwd[a]; mean_rev[e] = r.mean_rev[a];
            med_fwd[e] = r.med_fwd[a]; med_rev[e] = r.med_rev[a];
            sum_fwd[e] = r.sum_fwd[a]; sum_rev[e] = r.sum_rev[a];
            n_knodes[e] = (uint32_t)r.allele_knodes[a].size();
            ++e;
        }
}
void orc_gt_allele_knodes(void* g, uint32_t* out) {
    size_t e = 0;
    for (auto& r : ((GtHandle*)g)->g.records)
        for (auto& v : r.allele_knodes)
            for (uint32_t x : v) out[e++] = x;
}
// likelihood arithmetic alone (known-answer tests against the reference's VCF fixtures)
double orc_allele_likelihood(double E, double c, double o, double gaps, double err) {
    return allele_likelihood(E, c, o, gaps, err);
}

// whole `pandora map --genotype --local` replacement; writes outdir/pandora_genotyped.vcf
int orc_run_map(const char* prg, const char* reads, const char* vcf_refs, const char* outdir, const void* opts,
                uint32_t w, uint32_t k) {
    try {
        auto idx = build_index(prg, w, k);
        std::string data;
        std::vector<uint64_t> off;
        read_fastx(reads, data, off);
        MapOpts o = to_opts((const OrcOpts*)opts);
        MapResult mr;
        ReadSet rs{data.data(), off.data(), off.size() - 1};
        map_reads(*idx, rs, o, mr);
        std::map<std::string, std::string> refs;
        if (vcf_refs && *vcf_refs) refs = read_fasta(vcf_refs);
        auto g = genotype(*idx, mr, o, refs, "sample");
        std::ofstream f(std::string(outdir) + "/pandora_genotyped.vcf");
        if (!f) throw std::runtime_error("cannot write VCF in " + std::string(outdir));
        f << g.vcf_text;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
}
}
