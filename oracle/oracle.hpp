// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by the
// product path (drprg_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may use it, and there only as the checker / CPU baseline.
//
// CPU restatement (C++17, no dependencies) of the `pandora map --genotype --local` step that
// /root/reference/src/lib.rs:580-642 (Pandora::genotype_with) launches from
// /root/reference/src/predict.rs:285-303.  pandora itself (v0.10.0-alpha.0.1, pinned at
// /root/reference/justfile:16-17) is an un-vendored binary: its source is NOT in the reference
// tree, so this file restates pandora's published algorithm (rmcolq/pandora: src/inthash.cpp,
// seq.cpp, localPRG.cpp, localgraph.cpp, kmergraph.cpp, kmergraphwithcoverage.cpp, utils.cpp,
// estimate_parameters.cpp, sampleinfo.cpp, vcf.cpp, vcfrecord.cpp, map_main.cpp).
//
// PARITY STATUS: sketch / lookup / cluster / coverage / ML-path are **parity unpinned** (the
// reference holds no test or fixture for them).  Pinned against the reference's
// own fixtures: the PRG text grammar (tests/cases/expected/dr.prg), site -> VCF record
// enumeration (tests/cases/predict/in.vcf, SRR6824468.vcf rows for gid/pncA) and the genotype
// likelihood / GT / GT_CONF arithmetic (every data row of the pandora VCF fixtures).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Interval / Path in PRG-*string* coordinates (digits and spaces of the PRG text count).
// pandora: src/interval.cpp, src/prg/path.cpp
struct Interval {
    uint32_t start = 0, length = 0;
    Interval() = default;
    Interval(uint32_t s, uint32_t e) : start(s), length(e - s) {}
    uint32_t end() const { return start + length; }
    bool operator==(const Interval& y) const { return start == y.start && length == y.length; }
    bool operator<(const Interval& y) const {
        if (start != y.start) return start < y.start;
        return length < y.length;
    }
};
using Path = std::vector<Interval>;
uint32_t path_length(const Path& p);
inline uint32_t path_start(const Path& p) { return p.front().start; }
inline uint32_t path_end(const Path& p) { return p.back().end(); }
bool path_less(const Path& a, const Path& b);
Path path_subpath(const Path& p, uint32_t start, uint32_t len);
bool path_is_branching(const Path& x, const Path& y);
Path path_union(const Path& x, const Path& y);
bool path_is_subpath(const Path& small, const Path& big);

struct PathLess {
    bool operator()(const Path& a, const Path& b) const { return path_less(a, b); }
};

// ---------------------------------------------------------------------------------------------
// hashing (pandora src/inthash.cpp)
uint64_t hash64(uint64_t key, uint64_t mask);
int nt4(uint8_t c);
// fwd & reverse-complement hash of a k-long string
std::pair<uint64_t, uint64_t> kmerhash(const std::string& s, uint32_t k);

struct Minimizer {
    uint64_t hash;
    uint32_t start;  // [start, start+k)
    bool strand;     // hf <= hr
    bool operator<(const Minimizer& y) const {
        if (hash != y.hash) return hash < y.hash;
        if (start != y.start) return start < y.start;
        return strand < y.strand;
    }
};
// pandora Seq::minimizer_sketch; result ordered (hash, start, strand) like pandora's std::set
std::vector<Minimizer> sketch_read(const char* seq, size_t len, uint32_t w, uint32_t k);

// ---------------------------------------------------------------------------------------------
struct LocalNode {
    uint32_t id;
    std::string seq;
    Interval pos;
    std::vector<uint32_t> out, in;
};

struct KmerNode {
    uint32_t id;  // creation order (pandora's KmerNode::id)
    Path path;
    uint64_t khash = UINT64_MAX;
    bool strand = true;  // strand of the index record (hf <= hr)
    uint32_t num_AT = 0;
    std::vector<uint32_t> out, in;  // creation-order ids
};

struct KmerGraph {
    std::vector<KmerNode> nodes;         // by id
    std::vector<uint32_t> sorted;        // ids ordered by path  (pandora sorted_nodes)
    std::vector<uint32_t> rank;          // id -> position in sorted
    std::map<Path, uint32_t, PathLess> by_path;
    uint32_t add_node(const Path& p);
    void add_edge(uint32_t from, uint32_t to);
    void finalize();                     // builds sorted / rank
    void remove_shortcut_edges();
    uint32_t min_path_length() const;    // fewest edges start -> end
};

struct MiniRecord {
    uint32_t prg_id;
    uint32_t knode_id;  // creation-order id within the locus
    bool strand;
};

struct LocalPRG {
    uint32_t id;
    std::string name;
    std::string seq;  // PRG text
    std::vector<LocalNode> nodes;
    std::map<uint32_t, uint32_t> start_to_node;  // pos.start -> node id (starts are unique)
    KmerGraph kg;

    void build_graph();
    std::vector<Path> walk(uint32_t node_id, uint32_t pos, uint32_t len) const;
    std::vector<uint32_t> nodes_along_path(const Path& p) const;
    std::string string_along_path(const Path& p) const;
    std::vector<Path> shift(const Path& p) const;
    void minimizer_sketch(std::unordered_map<uint64_t, std::vector<MiniRecord>>& index, uint32_t w, uint32_t k);
    uint32_t last_end() const { return nodes.back().pos.end(); }
    std::vector<uint32_t> top_path() const;
    // node path spelling `s` exactly from node 0 to the last node (empty if none)
    std::vector<uint32_t> path_spelling(const std::string& s) const;
    std::string string_along_nodes(const std::vector<uint32_t>& np) const;
};

struct Index {
    uint32_t w = 0, k = 0;
    std::vector<LocalPRG> prgs;
    std::unordered_map<uint64_t, std::vector<MiniRecord>> minhash;
    std::vector<uint32_t> knode_base;  // global knode id = knode_base[prg] + rank
    uint32_t total_knodes = 0;
};
std::unique_ptr<Index> build_index(const std::string& prg_path, uint32_t w, uint32_t k);
std::unique_ptr<Index> build_index_from_text(const std::string& text, uint32_t w, uint32_t k);

// ---------------------------------------------------------------------------------------------
struct MapOpts {
    uint32_t min_cluster_size = 10;  // -c
    bool illumina = false;           // -I
    uint32_t genome_size = 4411532;  // -g
    uint32_t max_diff = 250;         // -m
    double e_rate = 0.11;            // -e
    double gt_error_rate = 0.01;     // -E
    double gt_conf = 0.0;            // --gt-conf
    uint32_t max_kmers_to_average = 100;
    int threads = 1;
};
// applies pandora's -I adjustments (e_rate 0.001, max_diff 2k+1)
MapOpts effective_opts(const MapOpts& o, uint32_t k);

struct Hit {
    uint32_t read_id, read_start;
    uint32_t prg_id, knode;  // knode = rank within the locus (sorted order)
    uint8_t forward;         // read strand == prg strand
    bool operator<(const Hit& y) const;
    bool operator==(const Hit& y) const {
        return read_id == y.read_id && read_start == y.read_start && prg_id == y.prg_id && knode == y.knode && forward == y.forward;
    }
};

struct ReadSet {  // ASCII reads
    const char* data;
    const uint64_t* off;  // n+1 offsets
    uint64_t n;
};

struct MapResult {
    std::vector<Hit> hits;          // all hits, globally ordered (read, prg, fwd-first, start, knode)
    std::vector<uint8_t> kept;      // per hit: survives cluster filters
    std::vector<uint32_t> cluster;  // per hit: cluster ordinal within the run (diagnostic)
    std::vector<uint32_t> cov_fwd, cov_rev;  // per global knode (saturating at 65535)
    std::vector<uint32_t> locus_reads;       // per locus: kept clusters
    uint64_t total_bases = 0, n_reads = 0, n_minimizers = 0;
    uint32_t first_read_len = 0;
};
void map_reads(const Index& idx, const ReadSet& rs, const MapOpts& o, MapResult& out, uint32_t first_read_len_hint = 0);

struct Params {
    uint32_t exp_depth_covg = 0;
    bool bin = false;
    double nb_p = 0.015, nb_r = 2.0;
    double e_rate = 0.11;
    int thresh = -25;
    uint32_t covg = 0;        // sum(read_len)/genome_size
    uint32_t min_kmer_covg = 0;
    double mean = 0, var = 0;
    uint64_t num_reads = 0;
};
Params estimate_parameters(const Index& idx, const MapResult& mr, const MapOpts& o);
double knode_log_prob(const Params& P, uint32_t fwd, uint32_t rev, bool terminal);

struct MLPath {
    std::vector<uint32_t> kpath;  // ranks (excluding null start / end)
    std::vector<uint32_t> lpath;  // local node ids
    bool skipped = false;         // locus dropped (no reads / coverage filter)
};
MLPath find_max_path(const Index& idx, uint32_t prg, const MapResult& mr, const Params& P, const MapOpts& o);

struct VcfRecord {
    std::string chrom;
    uint32_t pos = 0;  // 0-based
    std::string ref;
    std::vector<std::string> alts;
    std::string svtype, graphtype;
    // per allele (ref first): knode ranks overlapping the allele
    std::vector<std::vector<uint32_t>> allele_knodes;
    // sample stats
    std::vector<uint32_t> mean_fwd, mean_rev, med_fwd, med_rev, sum_fwd, sum_rev;
    std::vector<double> gaps, lik;
    int gt = -1;
    double gt_conf = 0;
    int ml_gt = -1;
};
std::string infer_svtype(const std::string& ref, const std::string& alt);
// site enumeration (read independent)
std::vector<VcfRecord> build_vcf_records(const Index& idx, uint32_t prg, const std::vector<uint32_t>& ref_path);
// likelihood arithmetic (pinned by the reference's VCF fixtures)
double allele_likelihood(double E, double c, double o, double gaps, double err);
void genotype_record(VcfRecord& r, const Params& P, const MapOpts& o);

struct GenotypeResult {
    Params params;
    std::vector<MLPath> ml;            // per locus
    std::vector<VcfRecord> records;    // merged, sorted
    std::vector<std::string> contigs;  // loci present in the sample
    std::string vcf_text;
};
GenotypeResult genotype(const Index& idx, const MapResult& mr, const MapOpts& o,
                        const std::map<std::string, std::string>& vcf_refs, const std::string& sample);
std::string format_vcf(const GenotypeResult& g, const std::string& sample);
std::map<std::string, std::string> read_fasta(const std::string& path);
// reads fasta/fastq (plain or gz) into ASCII read set storage
void read_fastx(const std::string& path, std::string& data, std::vector<uint64_t>& off);

}  // namespace orc
