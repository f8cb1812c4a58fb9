// ORACLE (test infrastructure, see oracle.hpp).  Site enumeration, per-allele coverage statistics,
// genotype likelihoods and the pandora_genotyped.vcf writer.  Restates pandora src/localPRG.cpp
// (build_vcf, add_sample_gt_to_vcf, add_sample_covgs_to_vcf, kmernode_path_from_localnode_path),
// sampleinfo.cpp (get_gaps, compute_likelihood, genotype_from_coverage), vcfrecord.cpp
// (infer_SVTYPE), vcf.cpp (merge_multi_allelic, sort, save).
// Output contract: /root/reference/src/lib.rs:644-646 (file name), src/filter.rs:48-63 (FORMAT
// tags), src/lib.rs:973-1027 (VcfExt), src/consequence.rs:100-113 (REF must equal genes.fa).
// PINNED by the reference fixtures: likelihood/GT/GT_CONF arithmetic (tests/cases/predict/*.vcf),
// POS/REF/ALT/VC/GRAPHTYPE of the gid/pncA sites (in.vcf, SRR6824468.vcf vs expected/dr.prg).
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// pandora VCFRecord::infer_SVTYPE
std::string infer_svtype(const std::string& ref, const std::string& alt) {
    if (ref.empty() && alt.empty()) return ".";
    if (ref.empty() || alt.empty()) return "INDEL";
    if (ref.size() == 1 && alt.size() == 1) return "SNP";
    if (alt.size() == ref.size()) return "PH_SNPs";
    if (ref.size() < alt.size() && alt.compare(0, ref.size(), ref) == 0) return "INDEL";
    if (alt.size() < ref.size() && ref.compare(0, alt.size(), alt) == 0) return "INDEL";
    return "COMPLEX";
}

namespace {
// offset (in bases) of `small` inside `big`, or -1 if small is not a sub-path of big.
// Trailing zero-length intervals of `small` (terminus extension through empty nodes) are ignored.
int64_t subpath_offset(Path small, const Path& big) {
    while (small.size() > 1 && small.back().length == 0) small.pop_back();
    if (small.empty() || big.empty()) return -1;
    uint32_t ls = path_length(small), lb = path_length(big);
    if (lb < ls || path_start(big) > path_start(small) || path_end(big) < path_end(small)) return -1;
    uint32_t offset = 0;
    for (const auto& iv : big) {
        if (iv.end() >= path_start(small)) {
            if (path_start(small) < iv.start) return -1;
            offset += path_start(small) - iv.start;
            if (offset + ls > lb) return -1;
            Path sp = path_subpath(big, offset, ls);
            return sp == small ? (int64_t)offset : -1;
        }
        offset += iv.length;
    }
    return -1;
}

// knode ranks lying on the local node path `np` that overlap sequence interval [A,B) of it
std::vector<uint32_t> knodes_overlapping(const LocalPRG& L, uint32_t k, const std::vector<uint32_t>& np, uint32_t A,
                                         uint32_t B) {
    std::vector<uint32_t> res;
    const KmerGraph& kg = L.kg;
    std::vector<uint32_t> cum(np.size() + 1, 0);
    for (size_t i = 0; i < np.size(); ++i) cum[i + 1] = cum[i] + L.nodes[np[i]].pos.length;
    size_t i0 = 0, i1 = np.size() - 1;
    for (size_t i = 0; i < np.size(); ++i)
        if (cum[i] + k <= A) i0 = i;
    for (size_t i = np.size(); i-- > 0;)
        if (cum[i + 1] >= B + k) i1 = i;
    if (i1 < i0) i1 = i0;
    Path sub;
    for (size_t i = i0; i <= i1; ++i) sub.push_back(L.nodes[np[i]].pos);
    uint32_t lo_start = L.nodes[np[i0]].pos.start, hi_end = L.nodes[np[i1]].pos.end();
    // knodes are sorted by path => by first-interval start
    size_t n = kg.sorted.size();
    size_t lo = 0, hi = n;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (path_start(kg.nodes[kg.sorted[mid]].path) < lo_start) lo = mid + 1;
        else hi = mid;
    }
    for (size_t r = lo; r < n; ++r) {
        const KmerNode& kn = kg.nodes[kg.sorted[r]];
        if (path_start(kn.path) > hi_end) break;
        if (r == 0 || r == n - 1) continue;  // null terminals
        if (path_length(kn.path) == 0) continue;
        int64_t off = subpath_offset(kn.path, sub);
        if (off < 0) continue;
        uint32_t s = cum[i0] + (uint32_t)off, e = s + path_length(kn.path);
        bool ov = (A == B) ? (s < A && e > A) : (s < B && e > A);
        if (ov) res.push_back((uint32_t)r);
    }
    return res;
}

struct Bi {  // biallelic record before merging
    uint32_t pos;
    std::string ref, alt, svtype, graphtype;
    std::vector<uint32_t> ref_kn, alt_kn;
    int ml = -1;
};
}  // namespace

// pandora LocalPRG::build_vcf + add_sample_covgs_to_vcf's allele -> k-mer-node mapping.
// Returns biallelic records folded into VcfRecord (alts.size()==1), unsorted.
std::vector<VcfRecord> build_vcf_records(const Index& idx, uint32_t prg, const std::vector<uint32_t>& ref) {
    const LocalPRG& L = idx.prgs[prg];
    std::vector<VcfRecord> out;
    if (ref.size() <= 1) return out;
    std::vector<uint32_t> level_start;
    int level = 0;
    std::string vartype = "SIMPLE";
    std::set<std::tuple<uint32_t, std::string, std::string>> seen;
    for (uint32_t ref_i = 0; ref_i + 1 < ref.size(); ++ref_i) {
        const LocalNode& nd = L.nodes[ref[ref_i]];
        if (nd.out.size() > 1) {
            ++level;
            level_start.push_back(ref_i);
            if (level > 1) vartype = "NESTED";
            continue;
        }
        if (level_start.empty()) continue;  // malformed (linear chain); nothing to close
        --level;
        uint32_t ls = level_start.back();
        uint32_t pos = 0;
        for (uint32_t j = 0; j <= ls; ++j) pos += (uint32_t)L.nodes[ref[j]].seq.size();
        std::string ref_seq;
        for (uint32_t j = ls + 1; j <= ref_i; ++j) ref_seq += L.nodes[ref[j]].seq;
        const uint32_t post = ref[ref_i + 1];
        std::deque<std::vector<uint32_t>> paths;
        std::vector<std::vector<uint32_t>> alts;
        for (uint32_t o : L.nodes[ref[ls]].out)
            if (o != ref[ls + 1]) paths.push_back({o});
        while (!paths.empty()) {
            auto vp = paths.front();
            paths.pop_front();
            const LocalNode& b = L.nodes[vp.back()];
            if (!b.out.empty() && b.out[0] == post) {
                alts.push_back(vp);
            } else {
                for (uint32_t o : b.out) {
                    paths.push_back(vp);
                    paths.back().push_back(o);
                }
            }
        }
        // ref allele k-mer nodes (on the reference path)
        std::vector<uint32_t> ref_kn = knodes_overlapping(L, idx.k, ref, pos, pos + (uint32_t)ref_seq.size());
        for (auto& alt : alts) {
            std::string alt_seq = L.string_along_nodes(alt);
            if (alt_seq == ref_seq) continue;
            if (!seen.insert({pos, ref_seq, alt_seq}).second) continue;
            VcfRecord r;
            r.chrom = L.name;
            r.pos = pos;
            r.ref = ref_seq;
            r.alts = {alt_seq};
            r.svtype = infer_svtype(ref_seq, alt_seq);
            r.graphtype = vartype;
            std::vector<uint32_t> ap(ref.begin(), ref.begin() + ls + 1);
            ap.insert(ap.end(), alt.begin(), alt.end());
            ap.insert(ap.end(), ref.begin() + ref_i + 1, ref.end());
            r.allele_knodes.push_back(ref_kn);
            r.allele_knodes.push_back(knodes_overlapping(L, idx.k, ap, pos, pos + (uint32_t)alt_seq.size()));
            out.push_back(std::move(r));
        }
        level_start.pop_back();
        if (level == 0) vartype = "SIMPLE";
    }
    return out;
}

// pandora SampleInfo::compute_likelihood (min-coverage thresholds all 0, as drprg runs it).
// LIKELIHOOD[i] = -E + c_i ln E - lnGamma(c_i+1) + o_i ln(err) - E g_i + (1-g_i) ln(1-e^-E)
double allele_likelihood(double E, double c, double o, double gaps, double err) {
    return -E + c * std::log(E) - std::lgamma(c + 1.0) + o * std::log(err) - E * gaps +
           std::log(1.0 - std::exp(-E)) * (1.0 - gaps);
}

// statistics must already be filled (mean_fwd/mean_rev/gaps); computes lik, gt, gt_conf
void genotype_record(VcfRecord& r, const Params& P, const MapOpts& o) {
    size_t na = r.mean_fwd.size();
    r.lik.assign(na, 0.0);
    double total = 0;
    for (size_t a = 0; a < na; ++a) total += (double)r.mean_fwd[a] + (double)r.mean_rev[a];
    for (size_t a = 0; a < na; ++a) {
        double c = (double)r.mean_fwd[a] + (double)r.mean_rev[a];
        r.lik[a] = allele_likelihood((double)P.exp_depth_covg, c, total - c, r.gaps[a], o.gt_error_rate);
    }
    size_t best = 0;
    for (size_t a = 1; a < na; ++a)
        if (r.lik[a] > r.lik[best]) best = a;
    double second = -INFINITY;
    for (size_t a = 0; a < na; ++a)
        if (a != best && r.lik[a] > second) second = r.lik[a];
    r.gt_conf = (na > 1) ? std::fabs(r.lik[best] - second) : 0.0;
    r.gt = (r.gt_conf >= o.gt_conf) ? (int)best : -1;
}

namespace {
void fill_stats(VcfRecord& r, const MapResult& mr, uint32_t base, uint32_t T) {
    size_t na = r.allele_knodes.size();
    r.mean_fwd.assign(na, 0); r.mean_rev.assign(na, 0); r.med_fwd.assign(na, 0); r.med_rev.assign(na, 0);
    r.sum_fwd.assign(na, 0); r.sum_rev.assign(na, 0); r.gaps.assign(na, 0.0);
    auto median = [](std::vector<uint32_t> v) -> uint32_t {
        if (v.empty()) return 0;
        std::sort(v.begin(), v.end());
        size_t n = v.size();
        return (n % 2) ? v[n / 2] : (v[n / 2 - 1] + v[n / 2]) / 2;
    };
    for (size_t a = 0; a < na; ++a) {
        std::vector<uint32_t> f, v;
        uint32_t gaps = 0;
        for (uint32_t kr : r.allele_knodes[a]) {
            uint32_t cf = mr.cov_fwd[base + kr], cr = mr.cov_rev[base + kr];
            f.push_back(cf);
            v.push_back(cr);
            r.sum_fwd[a] += cf;
            r.sum_rev[a] += cr;
            if (cf + cr < T) ++gaps;
        }
        size_t n = f.size();
        if (n) {
            r.mean_fwd[a] = r.sum_fwd[a] / (uint32_t)n;
            r.mean_rev[a] = r.sum_rev[a] / (uint32_t)n;
            r.gaps[a] = (double)gaps / (double)n;
        }
        r.med_fwd[a] = median(f);
        r.med_rev[a] = median(v);
    }
}

std::string g6(double v) {
    char b[64];
    snprintf(b, sizeof b, "%g", v);
    return b;
}
}  // namespace

GenotypeResult genotype(const Index& idx, const MapResult& mr, const MapOpts& o_in,
                        const std::map<std::string, std::string>& vcf_refs, const std::string& sample) {
    MapOpts o = effective_opts(o_in, idx.k);
    GenotypeResult G;
    G.params = estimate_parameters(idx, mr, o_in);
    const Params& P = G.params;
    G.ml.resize(idx.prgs.size());
    for (uint32_t l = 0; l < idx.prgs.size(); ++l) {
        const LocalPRG& L = idx.prgs[l];
        MLPath& ml = G.ml[l];
        ml = find_max_path(idx, l, mr, P, o);
        if (ml.skipped) continue;
        const uint32_t base = idx.knode_base[l];
        if (ml.kpath.empty()) {
            ml.skipped = true;
            continue;
        }
        // add_consensus_path_to_fastaq coverage sanity filter: per-base coverage along the ML
        // path (max over covering ML k-mers); skip locus if its mode is far from the global covg
        {
            std::vector<std::vector<uint32_t>> cv;
            std::map<uint32_t, size_t> where;
            for (size_t i = 0; i < ml.lpath.size(); ++i) {
                where[ml.lpath[i]] = i;
                cv.emplace_back(L.nodes[ml.lpath[i]].pos.length, 0);
            }
            for (uint32_t r : ml.kpath) {
                const KmerNode& kn = L.kg.nodes[L.kg.sorted[r]];
                uint32_t c = mr.cov_fwd[base + r] + mr.cov_rev[base + r];
                for (auto& iv : kn.path) {
                    if (iv.length == 0) continue;
                    auto nn = L.nodes_along_path({iv});
                    if (nn.empty()) continue;
                    auto it = where.find(nn[0]);
                    if (it == where.end()) continue;
                    uint32_t s = iv.start - L.nodes[nn[0]].pos.start;
                    for (uint32_t x = s; x < s + iv.length; ++x) cv[it->second][x] = std::max(cv[it->second][x], c);
                }
            }
            std::vector<uint32_t> flat;
            for (auto& v : cv) flat.insert(flat.end(), v.begin(), v.end());
            if (!flat.empty()) {
                std::sort(flat.begin(), flat.end());
                uint32_t mode = flat[0], best = 0;
                for (size_t i = 0; i < flat.size();) {
                    size_t j = i;
                    while (j < flat.size() && flat[j] == flat[i]) ++j;
                    if (j - i > best) {
                        best = (uint32_t)(j - i);
                        mode = flat[i];
                    }
                    i = j;
                }
                if (P.covg > 20 && ((uint64_t)mode * 10 < P.covg || mode > 10ull * P.covg)) {
                    ml.skipped = true;
                    ml.kpath.clear();
                    continue;
                }
            }
        }
        G.contigs.push_back(L.name);
        // reference path: --vcf-refs sequence if it threads the graph, else the top path
        std::vector<uint32_t> ref;
        auto it = vcf_refs.find(L.name);
        if (it != vcf_refs.end()) ref = L.path_spelling(it->second);
        if (ref.empty()) ref = L.top_path();
        std::vector<VcfRecord> recs = build_vcf_records(idx, l, ref);
        // add_sample_gt_to_vcf: regions where the ML path leaves the reference path
        {
            const auto& sp = ml.lpath;
            std::vector<uint32_t> cum(ref.size() + 1, 0);
            for (size_t i = 0; i < ref.size(); ++i) cum[i + 1] = cum[i] + (uint32_t)L.nodes[ref[i]].seq.size();
            size_t ri = 0, si = 0;
            while (ri < ref.size() && si < sp.size()) {
                // advance to next common node pair after (ri, si)
                size_t rj = ri + 1, sj = si + 1;
                while (rj < ref.size() && sj < sp.size() && ref[rj] != sp[sj]) {
                    if (ref[rj] < sp[sj]) ++rj; else ++sj;
                }
                if (rj >= ref.size() || sj >= sp.size()) break;
                if (rj > ri + 1 || sj > si + 1) {
                    std::string rs, as;
                    for (size_t j = ri + 1; j < rj; ++j) rs += L.nodes[ref[j]].seq;
                    for (size_t j = si + 1; j < sj; ++j) as += L.nodes[sp[j]].seq;
                    uint32_t pos = cum[ri + 1];
                    if (!(rs.empty() && as.empty()) && rs != as) {
                        bool found = false;
                        for (auto& r : recs)
                            if (r.pos == pos && r.ref == rs && r.alts[0] == as) {
                                r.ml_gt = 1;
                                found = true;
                            }
                        if (!found) {
                            VcfRecord r;
                            r.chrom = L.name;
                            r.pos = pos;
                            r.ref = rs;
                            r.alts = {as};
                            r.svtype = "COMPLEX";
                            r.graphtype = "TOO_MANY_ALTS";
                            r.ml_gt = 1;
                            std::vector<uint32_t> ap(ref.begin(), ref.begin() + ri + 1);
                            ap.insert(ap.end(), sp.begin() + si + 1, sp.begin() + sj);
                            ap.insert(ap.end(), ref.begin() + rj, ref.end());
                            r.allele_knodes.push_back(knodes_overlapping(L, idx.k, ref, pos, pos + (uint32_t)rs.size()));
                            r.allele_knodes.push_back(knodes_overlapping(L, idx.k, ap, pos, pos + (uint32_t)as.size()));
                            recs.push_back(std::move(r));
                        }
                    }
                }
                ri = rj;
                si = sj;
            }
        }
        // merge_multi_allelic: sort by (pos, ref, alt); merge equal (pos, ref)
        std::sort(recs.begin(), recs.end(), [](const VcfRecord& a, const VcfRecord& b) {
            if (a.pos != b.pos) return a.pos < b.pos;
            if (a.ref != b.ref) return a.ref < b.ref;
            return a.alts < b.alts;
        });
        std::vector<VcfRecord> merged;
        for (auto& r : recs) {
            if (!merged.empty() && merged.back().pos == r.pos && merged.back().ref == r.ref &&
                merged.back().graphtype != "TOO_MANY_ALTS" && r.graphtype != "TOO_MANY_ALTS") {
                merged.back().alts.push_back(r.alts[0]);
                merged.back().allele_knodes.push_back(r.allele_knodes[1]);
            } else {
                merged.push_back(r);
            }
        }
        // correct_dot_alleles: give empty alleles an anchor base
        std::string refseq = L.string_along_nodes(ref);
        for (auto& r : merged) {
            bool any_empty = r.ref.empty();
            for (auto& a : r.alts) any_empty |= a.empty();
            if (!any_empty) continue;
            if (r.pos > 0) {
                char anchor = refseq[r.pos - 1];
                r.pos -= 1;
                r.ref = std::string(1, anchor) + r.ref;
                for (auto& a : r.alts) a = std::string(1, anchor) + a;
            } else if (r.pos + r.ref.size() < refseq.size()) {
                char anchor = refseq[r.pos + r.ref.size()];
                r.ref += anchor;
                for (auto& a : r.alts) a += anchor;
            }
        }
        for (auto& r : merged) {
            fill_stats(r, mr, base, P.min_kmer_covg);
            genotype_record(r, P, o);
            G.records.push_back(std::move(r));
        }
    }
    std::stable_sort(G.records.begin(), G.records.end(), [](const VcfRecord& a, const VcfRecord& b) {
        if (a.chrom != b.chrom) return a.chrom < b.chrom;
        if (a.pos != b.pos) return a.pos < b.pos;
        if (a.ref != b.ref) return a.ref < b.ref;
        return a.alts < b.alts;
    });
    std::sort(G.contigs.begin(), G.contigs.end());
    G.vcf_text = format_vcf(G, sample);
    return G;
}

std::string format_vcf(const GenotypeResult& G, const std::string& sample) {
    std::ostringstream os;
    char date[32];
    time_t t = time(nullptr);
    strftime(date, sizeof date, "%d/%m/%y", localtime(&t));
    os << "##fileformat=VCFv4.3\n"
          "##FILTER=<ID=PASS,Description=\"All filters passed\">\n"
          "##fileDate=="
       << date
       << "\n"
          "##ALT=<ID=SNP,Description=\"SNP\">\n"
          "##ALT=<ID=PH_SNPs,Description=\"Phased SNPs\">\n"
          "##ALT=<ID=INDEL,Description=\"Insertion-deletion\">\n"
          "##ALT=<ID=COMPLEX,Description=\"Complex variant, collection of SNPs and indels\">\n"
          "##INFO=<ID=VC,Number=1,Type=String,Description=\"Type (class) of variant\">\n"
          "##ALT=<ID=SIMPLE,Description=\"Graph bubble is simple\">\n"
          "##ALT=<ID=NESTED,Description=\"Variation site was a nested feature in the graph\">\n"
          "##ALT=<ID=TOO_MANY_ALTS,Description=\"Variation site was a multinested feature with too many alts to include all in the VCF\">\n"
          "##INFO=<ID=GRAPHTYPE,Number=1,Type=String,Description=\"Type of graph feature\">\n"
          "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
          "##FORMAT=<ID=MEAN_FWD_COVG,Number=R,Type=Integer,Description=\"Mean forward coverage\">\n"
          "##FORMAT=<ID=MEAN_REV_COVG,Number=R,Type=Integer,Description=\"Mean reverse coverage\">\n"
          "##FORMAT=<ID=MED_FWD_COVG,Number=R,Type=Integer,Description=\"Med forward coverage\">\n"
          "##FORMAT=<ID=MED_REV_COVG,Number=R,Type=Integer,Description=\"Med reverse coverage\">\n"
          "##FORMAT=<ID=SUM_FWD_COVG,Number=R,Type=Integer,Description=\"Sum forward coverage\">\n"
          "##FORMAT=<ID=SUM_REV_COVG,Number=R,Type=Integer,Description=\"Sum reverse coverage\">\n"
          "##FORMAT=<ID=GAPS,Number=R,Type=Float,Description=\"Number of gap bases\">\n"
          "##FORMAT=<ID=LIKELIHOOD,Number=R,Type=Float,Description=\"Likelihood\">\n"
          "##FORMAT=<ID=GT_CONF,Number=1,Type=Float,Description=\"Genotype confidence\">\n";
    for (auto& c : G.contigs) os << "##contig=<ID=" << c << ">\n";
    os << "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" << sample << "\n";
    auto joinu = [](const std::vector<uint32_t>& v) {
        std::string s;
        for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + std::to_string(v[i]);
        return s;
    };
    auto joind = [](const std::vector<double>& v) {
        std::string s;
        for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + g6(v[i]);
        return s;
    };
    for (auto& r : G.records) {
        os << r.chrom << '\t' << r.pos + 1 << "\t.\t" << (r.ref.empty() ? "." : r.ref) << '\t';
        for (size_t i = 0; i < r.alts.size(); ++i) os << (i ? "," : "") << (r.alts[i].empty() ? "." : r.alts[i]);
        os << "\t.\t.\tVC=" << r.svtype << ";GRAPHTYPE=" << r.graphtype
           << "\tGT:MEAN_FWD_COVG:MEAN_REV_COVG:MED_FWD_COVG:MED_REV_COVG:SUM_FWD_COVG:SUM_REV_COVG:GAPS:LIKELIHOOD:GT_CONF\t";
        os << (r.gt < 0 ? std::string(".") : std::to_string(r.gt)) << ':' << joinu(r.mean_fwd) << ':' << joinu(r.mean_rev)
           << ':' << joinu(r.med_fwd) << ':' << joinu(r.med_rev) << ':' << joinu(r.sum_fwd) << ':' << joinu(r.sum_rev)
           << ':' << joind(r.gaps) << ':' << joind(r.lik) << ':' << g6(r.gt_conf) << '\n';
    }
    return os.str();
}

// ----------------------------------------------------------------------------------- IO ---
std::map<std::string, std::string> read_fasta(const std::string& path) {
    std::map<std::string, std::string> m;
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string name, line;
    char buf[1 << 16];
    while (gzgets(f, buf, sizeof buf)) {
        line = buf;
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            name = line.substr(1);
            size_t sp = name.find_first_of(" \t");
            if (sp != std::string::npos) name = name.substr(0, sp);
            m[name] = "";
        } else if (!name.empty()) {
            m[name] += line;
        }
    }
    gzclose(f);
    return m;
}

void read_fastx(const std::string& path, std::string& data, std::vector<uint64_t>& off) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1 << 20);
    data.clear();
    off.assign(1, 0);
    std::vector<char> buf(1 << 22);
    auto getline = [&](std::string& s) -> bool {
        s.clear();
        while (gzgets(f, buf.data(), (int)buf.size())) {
            s += buf.data();
            if (!s.empty() && s.back() == '\n') break;
        }
        if (s.empty()) return false;
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        return true;
    };
    std::string line;
    bool have = getline(line);
    while (have) {
        if (line.empty()) {
            have = getline(line);
            continue;
        }
        if (line[0] == '>') {
            std::string seq;
            while ((have = getline(line)) && (line.empty() || line[0] != '>')) seq += line;
            data += seq;
            off.push_back(data.size());
        } else if (line[0] == '@') {
            std::string seq, plus, qual;
            getline(seq);
            getline(plus);
            getline(qual);
            data += seq;
            off.push_back(data.size());
            have = getline(line);
        } else {
            throw std::runtime_error("unrecognised read file format: " + path);
        }
    }
    gzclose(f);
}

}  // namespace orc
