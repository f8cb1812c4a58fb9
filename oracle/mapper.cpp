// ORACLE (test infrastructure, see oracle.hpp).  Reads -> minimizer hits -> clusters -> k-mer
// coverage -> model parameters -> max-likelihood k-mer path.  Restates pandora src/utils.cpp
// (add_read_hits, define_clusters, filter_clusters, filter_clusters2, pangraph_from_read_file),
// pangenome/pangraph.cpp (add_hits_to_kmergraphs), estimate_parameters.cpp and
// kmergraphwithcoverage.cpp (nbin_prob, bin_prob, find_max_path), as run by the argv that
// /root/reference/src/lib.rs:594-617 + /root/reference/src/predict.rs:288-294 build.
// parity unpinned: the reference has no fixture for these stages.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "oracle.hpp"
#include <thread>

namespace orc {

MapOpts effective_opts(const MapOpts& o, uint32_t k) {
    MapOpts e = o;
    if (o.illumina) {
        if (e.e_rate == 0.11) e.e_rate = 0.001;       // pandora map_main: -I sets the error rate
        if (e.max_diff > 200) e.max_diff = 2 * k + 1;  // and caps the intra-cluster gap
    }
    return e;
}

// pandora MinimizerHit::operator< (forward hits first)
bool Hit::operator<(const Hit& y) const {
    if (read_id != y.read_id) return read_id < y.read_id;
    if (prg_id != y.prg_id) return prg_id < y.prg_id;
    if (forward != y.forward) return forward > y.forward;
    if (read_start != y.read_start) return read_start < y.read_start;
    return knode < y.knode;
}

namespace {
struct Cluster {
    uint32_t b, e;  // hit index range in the read's sorted hits
    uint32_t size() const { return e - b; }
};

// one read: sketch, lookup, cluster, filter.  Appends hits/kept/cluster to the per-thread buffers.
struct ReadWorker {
    const Index& idx;
    const MapOpts& o;  // effective
    uint32_t expected_kmers_short;
    std::vector<uint32_t> min_path_len;
    double fraction;

    void run(uint32_t read_id, const char* s, size_t len, std::vector<Hit>& hits, std::vector<uint8_t>& kept,
             std::vector<uint32_t>& clus, uint64_t& n_mini) const {
        auto sk = sketch_read(s, len, idx.w, idx.k);
        n_mini += sk.size();
        std::vector<Hit> h;
        for (auto& m : sk) {
            auto it = idx.minhash.find(m.hash);
            if (it == idx.minhash.end()) continue;
            for (auto& r : it->second) {
                Hit x;
                x.read_id = read_id;
                x.read_start = m.start;
                x.prg_id = r.prg_id;
                x.knode = idx.prgs[r.prg_id].kg.rank[r.knode_id];
                x.forward = (m.strand == r.strand) ? 1 : 0;
                h.push_back(x);
            }
        }
        if (h.empty()) return;
        std::sort(h.begin(), h.end());
        h.erase(std::unique(h.begin(), h.end()), h.end());
        const size_t n = h.size();
        std::vector<uint8_t> k(n, 0);
        std::vector<uint32_t> cid(n, UINT32_MAX);

        // define_clusters
        std::vector<Cluster> cl;
        auto close = [&](uint32_t b, uint32_t e) {
            uint32_t lbt = (uint32_t)(std::min(min_path_len[h[b].prg_id], expected_kmers_short) * fraction);
            if (e - b > std::max(lbt, o.min_cluster_size)) cl.push_back({b, e});
        };
        uint32_t b = 0;
        for (uint32_t i = 1; i < n; ++i) {
            const Hit &p = h[i - 1], &c = h[i];
            int64_t d = (int64_t)c.read_start - (int64_t)p.read_start;
            if (d < 0) d = -d;
            if (c.prg_id != p.prg_id || c.forward != p.forward || d > (int64_t)o.max_diff) {
                close(b, i);
                b = i;
            }
        }
        close(b, (uint32_t)n);
        if (cl.empty()) {
            hits.insert(hits.end(), h.begin(), h.end());
            kept.insert(kept.end(), k.begin(), k.end());
            clus.insert(clus.end(), cid.begin(), cid.end());
            return;
        }
        // filter_clusters: order by clusterComp, compare adjacent pairs
        std::vector<uint32_t> ord(cl.size());
        for (uint32_t i = 0; i < cl.size(); ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) {
            const Hit &hx = h[cl[x].b], &hy = h[cl[y].b];
            if (hx.read_start != hy.read_start) return hx.read_start < hy.read_start;
            if (cl[x].size() != cl[y].size()) return cl[x].size() > cl[y].size();
            if (hx.prg_id != hy.prg_id) return hx.prg_id < hy.prg_id;
            return hx.forward < hy.forward;
        });
        std::vector<uint8_t> alive(cl.size(), 1);
        {
            uint32_t prev = ord[0];
            for (size_t t = 1; t < ord.size(); ++t) {
                uint32_t cur = ord[t];
                const Hit &pf = h[cl[prev].b], &cf = h[cl[cur].b];
                const Hit &pl = h[cl[prev].e - 1], &cl_last = h[cl[cur].e - 1];
                bool cond = (cf.prg_id == pf.prg_id && cf.forward != pf.forward) || (cl_last.read_start <= pl.read_start);
                if (cond) {
                    if (cl[prev].size() >= cl[cur].size()) {
                        alive[cur] = 0;
                        continue;  // prev stays
                    }
                    alive[prev] = 0;
                }
                prev = cur;
            }
        }
        // filter_clusters2: by decreasing size; drop clusters whose read span is already covered
        {
            std::vector<uint32_t> ord2;
            for (uint32_t i = 0; i < cl.size(); ++i)
                if (alive[i]) ord2.push_back(i);
            std::sort(ord2.begin(), ord2.end(), [&](uint32_t x, uint32_t y) {
                const Hit &hx = h[cl[x].b], &hy = h[cl[y].b];
                if (cl[x].size() != cl[y].size()) return cl[x].size() > cl[y].size();
                if (hx.read_start != hy.read_start) return hx.read_start < hy.read_start;
                if (hx.prg_id != hy.prg_id) return hx.prg_id < hy.prg_id;
                return hx.forward < hy.forward;
            });
            std::vector<std::pair<uint32_t, uint32_t>> spans;  // processed [first_start, last_start)
            for (size_t t = 0; t < ord2.size(); ++t) {
                uint32_t c = ord2[t];
                uint32_t a = h[cl[c].b].read_start, z = h[cl[c].e - 1].read_start;
                if (t > 0) {
                    uint32_t cur = a;
                    bool contained = true;
                    while (cur < z) {
                        uint32_t best = cur;
                        for (auto& sp : spans)
                            if (sp.first <= cur && cur < sp.second) best = std::max(best, sp.second);
                        if (best == cur) {
                            contained = false;
                            break;
                        }
                        cur = best;
                    }
                    if (contained) {
                        alive[c] = 0;
                        continue;
                    }
                }
                spans.push_back({a, z});
            }
        }
        for (uint32_t i = 0; i < cl.size(); ++i) {
            for (uint32_t j = cl[i].b; j < cl[i].e; ++j) {
                cid[j] = i;
                k[j] = alive[i];
            }
        }
        hits.insert(hits.end(), h.begin(), h.end());
        kept.insert(kept.end(), k.begin(), k.end());
        clus.insert(clus.end(), cid.begin(), cid.end());
    }
};
}  // namespace

void map_reads(const Index& idx, const ReadSet& rs, const MapOpts& o_in, MapResult& out, uint32_t first_read_len_hint) {
    MapOpts o = effective_opts(o_in, idx.k);
    out = MapResult();
    out.n_reads = rs.n;
    out.first_read_len = first_read_len_hint ? first_read_len_hint : (rs.n ? (uint32_t)(rs.off[1] - rs.off[0]) : 0);
    uint32_t expected = UINT32_MAX;
    if (o.illumina) expected = out.first_read_len * 2 / (idx.w + 1);
    ReadWorker wk{idx, o, expected, {}, 0.5 / std::exp(o.e_rate * idx.k)};
    for (auto& p : idx.prgs) wk.min_path_len.push_back(p.kg.min_path_length());

    int T = std::max(1, o.threads);
    std::vector<std::vector<Hit>> th(T);
    std::vector<std::vector<uint8_t>> tk(T);
    std::vector<std::vector<uint32_t>> tc(T);
    std::vector<uint64_t> tm(T, 0), tb(T, 0);
    // contiguous read blocks per thread keep the global (read-ordered) hit order deterministic
    auto body = [&](int t) {
        uint64_t lo = rs.n * (uint64_t)t / T, hi = rs.n * (uint64_t)(t + 1) / T;
        for (uint64_t r = lo; r < hi; ++r) {
            size_t len = rs.off[r + 1] - rs.off[r];
            tb[t] += len;
            wk.run((uint32_t)r, rs.data + rs.off[r], len, th[t], tk[t], tc[t], tm[t]);
        }
    };
    if (T == 1) {
        body(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t) pool.emplace_back(body, t);
        for (auto& p : pool) p.join();
    }
    for (int t = 0; t < T; ++t) {
        out.hits.insert(out.hits.end(), th[t].begin(), th[t].end());
        out.kept.insert(out.kept.end(), tk[t].begin(), tk[t].end());
        out.cluster.insert(out.cluster.end(), tc[t].begin(), tc[t].end());
        out.n_minimizers += tm[t];
        out.total_bases += tb[t];
    }
    // pangenome::Graph::add_hits_to_kmergraphs + add_hits_between_PRG_and_read
    out.cov_fwd.assign(idx.total_knodes, 0);
    out.cov_rev.assign(idx.total_knodes, 0);
    out.locus_reads.assign(idx.prgs.size(), 0);
    for (size_t i = 0; i < out.hits.size(); ++i) {
        if (!out.kept[i]) continue;
        const Hit& h = out.hits[i];
        uint32_t g = idx.knode_base[h.prg_id] + h.knode;
        uint32_t& c = h.forward ? out.cov_fwd[g] : out.cov_rev[g];
        if (c < 65535) ++c;  // saturating uint16 upstream
        bool first_of_cluster = (i == 0) || out.hits[i - 1].read_id != h.read_id || out.cluster[i - 1] != out.cluster[i];
        if (first_of_cluster) out.locus_reads[h.prg_id] += 1;
    }
}

// ------------------------------------------------------------------- estimate_parameters ---
namespace {
double fit_mean_covg(const std::vector<uint32_t>& d, uint32_t zero_thresh) {
    double sum = 0, total = 0;
    for (uint32_t i = zero_thresh; i < d.size(); ++i) {
        sum += (double)d[i] * i;
        total += d[i];
    }
    return total == 0 ? 0 : sum / total;
}
double fit_variance_covg(const std::vector<uint32_t>& d, double mean, uint32_t zero_thresh) {
    double acc = 0, total = 0;
    for (uint32_t i = zero_thresh; i < d.size(); ++i) {
        acc += ((double)i - mean) * ((double)i - mean) * d[i];
        total += d[i];
    }
    return total == 0 ? 0 : acc / total;
}
// position of the second peak of a bimodal k-mer coverage distribution
uint32_t find_mean_covg(const std::vector<uint32_t>& d) {
    bool first_peak = true;
    uint32_t max_covg = 0, noise_buffer = 0;
    for (uint32_t i = 1; i < d.size(); ++i) {
        if (d[i] <= d[i - 1]) continue;
        if (first_peak && noise_buffer < 3) {
            ++noise_buffer;
            continue;
        }
        if (first_peak) {
            first_peak = false;
            max_covg = i;
        } else if (d[i] > d[max_covg]) {
            max_covg = i;
        }
    }
    if (first_peak) max_covg = 0;
    return max_covg;
}
// valley between the error peak and the signal peak of the log-prob histogram (bins [-200,0))
int find_prob_thresh(const std::vector<uint32_t>& d) {
    if (d.empty()) return 0;
    int n = (int)d.size();
    int p1 = (int)(std::max_element(d.begin(), d.end()) - d.begin());
    int p2 = -1;
    for (int i = 0; i < n; ++i) {
        if (std::abs(i - p1) <= 10) continue;
        if (d[i] == 0) continue;
        if (p2 < 0 || d[i] > d[p2]) p2 = i;
    }
    if (p2 < 0) return p1 - 200 - 10 < -200 ? -200 : p1 - 200 - 10;
    int a = std::min(p1, p2), b = std::max(p1, p2);
    int t = (int)(std::min_element(d.begin() + a, d.begin() + b + 1) - d.begin());
    return t - 200;
}
}  // namespace

double knode_log_prob(const Params& P, uint32_t fwd, uint32_t rev, bool terminal) {
    const double FLOOR = (double)std::numeric_limits<float>::lowest() / 1000.0;
    (void)terminal;  // pandora's nbin_prob also scores the null terminals (coverage 0)
    double c = (double)fwd + (double)rev;
    double v = std::lgamma(c + P.nb_r) - std::lgamma(P.nb_r) - std::lgamma(c + 1.0) + P.nb_r * std::log(P.nb_p) +
               c * std::log(1.0 - P.nb_p);
    return std::max(v, FLOOR);
}

namespace {
double bin_log_prob(double p, uint32_t num, uint32_t fwd, uint32_t rev, bool terminal) {
    if (terminal) return 0.0;
    uint32_t s = fwd + rev;
    auto lnck2 = [](double n, double a, double b) {
        return std::lgamma(n + 1) - std::lgamma(a + 1) - std::lgamma(b + 1) - std::lgamma(n - a - b + 1);
    };
    if (s > num) return lnck2(s, fwd, rev) + s * std::log(p / 2);
    return lnck2(num, fwd, rev) + s * std::log(p / 2) + (num - s) * std::log(1 - p);
}
double node_prob(const Params& P, uint32_t k, uint32_t fwd, uint32_t rev, bool terminal) {
    if (P.bin) return bin_log_prob(1.0 / std::exp(P.e_rate * k), P.exp_depth_covg, fwd, rev, terminal);
    return knode_log_prob(P, fwd, rev, terminal);
}
}  // namespace

Params estimate_parameters(const Index& idx, const MapResult& mr, const MapOpts& o_in) {
    MapOpts o = effective_opts(o_in, idx.k);
    Params P;
    P.e_rate = o.e_rate;
    P.covg = (uint32_t)(mr.total_bases / std::max<uint32_t>(1, o.genome_size));
    P.exp_depth_covg = P.covg;
    std::vector<uint32_t> dist(1000, 0), pdist(200, 0);
    uint64_t num_reads = 0, n_present = 0;
    for (size_t l = 0; l < idx.prgs.size(); ++l) {
        if (mr.locus_reads[l] == 0) continue;
        ++n_present;
        num_reads += mr.locus_reads[l];
        const KmerGraph& kg = idx.prgs[l].kg;
        for (uint32_t r = 1; r + 1 < kg.sorted.size(); ++r) {
            uint32_t g = idx.knode_base[l] + r;
            uint32_t c = mr.cov_fwd[g] + mr.cov_rev[g];
            if (c < 1000) dist[c] += 1;
        }
    }
    if (n_present == 0) {
        P.min_kmer_covg = P.exp_depth_covg / 10;
        return P;
    }
    num_reads /= n_present;
    P.num_reads = num_reads;
    double mean = fit_mean_covg(dist, P.covg / 10);
    double var = fit_variance_covg(dist, mean, P.covg / 10);
    P.mean = mean;
    P.var = var;
    if (P.bin && num_reads > 30 && P.covg > 30) {
        uint32_t mc = find_mean_covg(dist);
        P.exp_depth_covg = mc;
        if (mc > 0 && mc < P.covg) P.e_rate = -std::log((double)mc / P.covg) / idx.k;
    } else if (!P.bin && num_reads > 30 && P.covg > 2 && mean < var && mean > 0) {
        // fit_negative_binomial, then KmerGraphWithCoverage::set_negative_binomial_parameters (+=)
        double p = mean / var;
        double r = (mean * p / (1 - p) + p * var / (1 - p)) / 2;
        P.nb_p = 0.015 + p;
        P.nb_r = 2.0 + r;
        P.exp_depth_covg = (uint32_t)mean;
    } else {
        P.exp_depth_covg = (uint32_t)fit_mean_covg(dist, P.covg / 10);
        P.exp_depth_covg = std::max<uint32_t>(P.exp_depth_covg, 1);
    }
    if (P.nb_p >= 1.0) P.nb_p = 0.999999;
    // probability threshold
    for (size_t l = 0; l < idx.prgs.size(); ++l) {
        if (mr.locus_reads[l] == 0) continue;
        const KmerGraph& kg = idx.prgs[l].kg;
        for (uint32_t r = 1; r + 1 < kg.sorted.size(); ++r) {
            uint32_t g = idx.knode_base[l] + r;
            double p = node_prob(P, idx.k, mr.cov_fwd[g], mr.cov_rev[g], false);
            if (p >= -200.0 && p < 0.0) {
                int j = (int)std::floor(p + 200.0);
                if (j >= 0 && j < 200) pdist[j] += 1;
            }
        }
    }
    P.thresh = find_prob_thresh(pdist);
    P.min_kmer_covg = P.exp_depth_covg / 10;
    return P;
}

// ------------------------------------------------------------------------ find_max_path ---
MLPath find_max_path(const Index& idx, uint32_t prg, const MapResult& mr, const Params& P, const MapOpts& o) {
    MLPath out;
    const LocalPRG& L = idx.prgs[prg];
    const KmerGraph& kg = L.kg;
    const uint32_t n = (uint32_t)kg.sorted.size();
    if (mr.locus_reads[prg] == 0 || n < 2) {
        out.skipped = true;
        return out;
    }
    const uint32_t base = idx.knode_base[prg];
    std::vector<double> prob(n), M(n, 0.0);
    std::vector<uint32_t> len(n, 0), prev(n, n - 1);
    for (uint32_t r = 0; r < n; ++r)
        prob[r] = node_prob(P, idx.k, mr.cov_fwd[base + r], mr.cov_rev[base + r], r == 0 || r == n - 1);
    const double tol = 0.000001;
    const double thresh = (double)P.thresh;
    const uint32_t W = o.max_kmers_to_average;
    for (uint32_t j = n - 1; j-- > 0;) {
        double max_mean = (double)std::numeric_limits<float>::lowest();
        uint32_t max_len = 0;
        const KmerNode& nd = kg.nodes[kg.sorted[j]];
        // out-neighbours visited in rank order (canonical; pandora uses insertion order)
        std::vector<uint32_t> outs;
        for (uint32_t oid : nd.out) outs.push_back(kg.rank[oid]);
        std::sort(outs.begin(), outs.end());
        for (uint32_t v : outs) {
            bool is_term = (v == n - 1);
            bool take;
            if (is_term) {
                take = thresh > max_mean + tol;
            } else {
                double mean_v = M[v] / len[v];
                take = (mean_v > max_mean + tol) || (max_mean - mean_v <= tol && len[v] > max_len);
            }
            if (!take) continue;
            M[j] = prob[j] + M[v];
            len[j] = 1 + len[v];
            prev[j] = v;
            if (len[j] > W) {
                uint32_t pn = prev[j];
                for (uint32_t step = 1; step < W; ++step) pn = prev[pn];
                M[j] -= prob[pn];
                len[j] -= 1;
            }
            if (!is_term) {
                max_mean = M[v] / len[v];
                max_len = len[v];
            } else {
                max_mean = thresh;
            }
        }
    }
    uint32_t p = prev[0];
    while (p < n - 1) {
        out.kpath.push_back(p);
        p = prev[p];
        if (out.kpath.size() > 1000000) break;
    }
    // localnode_path_from_kmernode_path: union of the local nodes under the k-mer path, gaps and
    // both ends completed along the top (first out-edge) path.
    std::vector<uint32_t>& lp = out.lpath;
    for (uint32_t r : out.kpath) {
        auto nn = L.nodes_along_path(kg.nodes[kg.sorted[r]].path);
        if (nn.empty()) continue;
        while (!lp.empty() && !L.nodes[lp.back()].out.empty() && nn[0] > L.nodes[lp.back()].out[0] &&
               std::find(L.nodes[lp.back()].out.begin(), L.nodes[lp.back()].out.end(), nn[0]) == L.nodes[lp.back()].out.end())
            lp.push_back(L.nodes[lp.back()].out[0]);
        while (!lp.empty() && nn[0] <= lp.back()) lp.pop_back();
        lp.insert(lp.end(), nn.begin(), nn.end());
    }
    if (lp.empty()) {
        lp = L.top_path();
    } else {
        if (lp.front() != 0) {  // extend to the start: any node path 0 -> lp.front(), first-edge preference
            std::vector<uint32_t> pre;
            std::vector<char> can(L.nodes.size(), 0);
            can[lp.front()] = 1;
            for (uint32_t i = lp.front(); i-- > 0;)
                for (uint32_t oo : L.nodes[i].out)
                    if (oo <= lp.front() && can[oo]) can[i] = 1;
            uint32_t cur = 0;
            while (cur != lp.front()) {
                pre.push_back(cur);
                uint32_t nxt = UINT32_MAX;
                for (uint32_t oo : L.nodes[cur].out)
                    if (oo <= lp.front() && can[oo]) {
                        nxt = oo;
                        break;
                    }
                if (nxt == UINT32_MAX) break;
                cur = nxt;
            }
            lp.insert(lp.begin(), pre.begin(), pre.end());
        }
        while (!L.nodes[lp.back()].out.empty()) lp.push_back(L.nodes[lp.back()].out[0]);
    }
    return out;
}

}  // namespace orc
