// See genotype_host.hpp.  Behaviour follows pandora's estimate_parameters.cpp, LocalPRG::build_vcf /
// add_sample_gt_to_vcf / add_sample_covgs_to_vcf, VCF::merge_multi_allelic and VCF::save as they
// run under the argv of /root/reference/src/lib.rs:594-609; the VCF schema is the one
// /root/reference/src/filter.rs:48-63 and src/lib.rs:935-1181 (VcfExt) consume.
#include "genotype_host.hpp"
#include "gzip_inflate.hpp"

#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace drprg {

namespace {
class WorkerPool {
  public:
    static WorkerPool& get() {
        static WorkerPool* p = new WorkerPool(15u, true);  // leaked on purpose: workers may outlive static destructors
        return *p;
    }
    // ingest / gzip inflate: coarse, long sections that want every core of the rank's share; its workers sleep between
    // sections instead of spinning (the genotype step's short sections keep the small spinning pool above)
    static WorkerPool& io() {
        static WorkerPool* p = new WorkerPool(63u, false);
        return *p;
    }
    void run(size_t n, const std::function<void(size_t)>& fn, size_t max_threads) {
        if (n == 0) return;
        const size_t helpers = std::min({workers_.size(), max_threads > 0 ? max_threads - 1 : 0, n - 1});
        if (helpers == 0) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        std::unique_lock<std::mutex> run_lock(run_mutex_);  // one parallel_for at a time
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = helpers;
            wanted_ = helpers;
            error_.clear();
            ++generation_;
            gen_hint_.store(generation_, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        if (!error_.empty()) throw std::runtime_error(error_);
    }

  private:
    WorkerPool(unsigned cap, bool spin) : spin_(spin) {
        // one process per GPU shares the host with its sibling ranks: the pool takes this rank's share of the cores
        // (DRPRG_THREADS overrides; LOCAL_WORLD_SIZE is what torchrun sets), otherwise 8 ranks x 16 spinning threads
        // would fight over the same cores
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned ranks = 1, budget = 0;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(e));
        if (const char* e = getenv("DRPRG_THREADS")) budget = (unsigned)std::max(1, atoi(e));
        if (!budget) budget = std::max(1u, hw / ranks);
        const unsigned n = std::min(cap, budget > 1 ? budget - 1 : 0);
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
        for (auto& w : workers_) w.detach();
    }
    void work() {
        try {
            for (size_t i = next_++; i < n_; i = next_++) (*fn_)(i);
        } catch (const std::exception& e) {
            std::lock_guard<std::mutex> g(m_);
            if (error_.empty()) error_ = e.what();
            next_.store(n_);
        }
    }
    void loop() {
        uint64_t seen = 0;
        while (true) {
            // the parallel sections of one sample follow each other within ~0.1 ms: spin briefly before sleeping, a futex
            // wake-up of 15 threads costs more than the work of a section
            static const int spin_max = [] {
                const char* e = getenv("DRPRG_SPIN");
                return e ? atoi(e) : 4000;
            }();
            for (int spin = 0; spin_ && spin < spin_max && gen_hint_.load(std::memory_order_acquire) == seen; ++spin) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_ != seen && wanted_ > 0; });
                seen = generation_;
                --wanted_;
            }
            work();
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    const bool spin_;
    std::vector<std::thread> workers_;
    std::mutex m_, run_mutex_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0, pending_ = 0, wanted_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t generation_ = 0;
    std::atomic<uint64_t> gen_hint_{0};
    std::string error_;
};
}  // namespace

void parallel_for(size_t n, const std::function<void(size_t)>& fn, size_t max_threads) {
    WorkerPool::get().run(n, fn, max_threads);
}
void parallel_for_io(size_t n, const std::function<void(size_t)>& fn, size_t max_threads) {
    WorkerPool::io().run(n, fn, max_threads);
}

static inline uint32_t sat16(int32_t c) { return c > 65535 ? 65535u : (uint32_t)(c < 0 ? 0 : c); }

// ----------------------------------------------------------------------------- S6 ---
double host_node_log_prob(const FitParams& P, uint32_t k, uint32_t f, uint32_t r, bool terminal) {
    if (P.bin) {
        if (terminal) return 0.0;
        const double p = 1.0 / std::exp(P.e_rate * k);
        const uint32_t s = f + r;
        const double n = (double)std::max(s, P.E);
        const double lnck2 = std::lgamma(n + 1.0) - std::lgamma(f + 1.0) - std::lgamma(r + 1.0) - std::lgamma(n - f - r + 1.0);
        if (s > P.E) return lnck2 + s * std::log(p / 2);
        return lnck2 + s * std::log(p / 2) + (P.E - s) * std::log(1 - p);
    }
    const double c = (double)f + (double)r;
    const double v = std::lgamma(c + P.nb_r) - std::lgamma(P.nb_r) - std::lgamma(c + 1.0) + P.nb_r * std::log(P.nb_p) +
                     c * std::log(1.0 - P.nb_p);
    return std::max(v, -(double)FLT_MAX / 1000.0);
}

FitParams fit_parameters(const HostIndex& H, const int32_t* cov, const int32_t* locus_reads, uint64_t total_bases,
                         const SampleOpts& o) {
    uint32_t hist[1000] = {0};
    for (size_t l = 0; l < H.loci.size(); ++l) {
        if (locus_reads[l] <= 0) continue;
        for (uint32_t g = H.knode_base[l] + 1; g + 1 < H.knode_base[l + 1]; ++g) {
            uint32_t c = sat16(cov[2 * g]) + sat16(cov[2 * g + 1]);
            if (c < 1000) ++hist[c];
        }
    }
    return fit_parameters_hist(H, hist, locus_reads, total_bases, o);
}

// the same fit from a ready histogram of per-node total coverage (built on the device by cov_hist_kernel)
FitParams fit_parameters_hist(const HostIndex& H, const uint32_t* hist, const int32_t* locus_reads, uint64_t total_bases,
                              const SampleOpts& o) {
    FitParams P;
    P.e_rate = o.e_rate;
    P.covg = (uint32_t)(total_bases / std::max<uint32_t>(1u, o.genome_size));
    P.E = P.covg;
    uint64_t reads = 0, present = 0;
    for (size_t l = 0; l < H.loci.size(); ++l) {
        if (locus_reads[l] <= 0) continue;
        ++present;
        reads += (uint64_t)locus_reads[l];
    }
    if (!present) {
        P.min_kmer_covg = P.E / 10;
        return P;
    }
    P.num_reads = reads / present;
    auto moments = [&](uint32_t from, double& mean, double& var) {
        double sum = 0, tot = 0;
        for (uint32_t i = from; i < 1000; ++i) {
            sum += (double)hist[i] * i;
            tot += hist[i];
        }
        mean = tot == 0 ? 0 : sum / tot;
        double acc = 0;
        for (uint32_t i = from; i < 1000; ++i) acc += ((double)i - mean) * ((double)i - mean) * hist[i];
        var = tot == 0 ? 0 : acc / tot;
    };
    moments(P.covg / 10, P.mean, P.var);
    if (P.bin && P.num_reads > 30 && P.covg > 30) {
        // second peak of the coverage histogram
        bool first_peak = true;
        uint32_t peak = 0, noise = 0;
        for (uint32_t i = 1; i < 1000; ++i) {
            if (hist[i] <= hist[i - 1]) continue;
            if (first_peak && noise < 3) { ++noise; continue; }
            if (first_peak) { first_peak = false; peak = i; }
            else if (hist[i] > hist[peak]) peak = i;
        }
        if (first_peak) peak = 0;
        P.E = peak;
        if (peak > 0 && peak < P.covg) P.e_rate = -std::log((double)peak / P.covg) / H.k;
    } else if (!P.bin && P.num_reads > 30 && P.covg > 2 && P.mean < P.var && P.mean > 0) {
        const double p = P.mean / P.var;
        const double r = (P.mean * p / (1 - p) + p * P.var / (1 - p)) / 2;
        P.nb_p = 0.015 + p;  // pandora adds the fit to its defaults (set_negative_binomial_parameters)
        P.nb_r = 2.0 + r;
        P.E = (uint32_t)P.mean;
    } else {
        double m, v;
        moments(P.covg / 10, m, v);
        P.E = std::max<uint32_t>((uint32_t)m, 1u);
    }
    if (P.nb_p >= 1.0) P.nb_p = 0.999999;
    P.min_kmer_covg = P.E / 10;
    return P;
}

// valley between the error peak and the signal peak of the log-probability histogram (bins [-200, 0));
// the histogram itself is built on the device (prob_hist_kernel)
int prob_threshold(const uint32_t* ph) {
    int p1 = (int)(std::max_element(ph, ph + 200) - ph), p2 = -1;
    for (int i = 0; i < 200; ++i) {
        if (std::abs(i - p1) <= 10 || ph[i] == 0) continue;
        if (p2 < 0 || ph[i] > ph[p2]) p2 = i;
    }
    if (p2 < 0) return std::max(-200, p1 - 200 - 10);
    int a = std::min(p1, p2), b = std::max(p1, p2);
    return (int)(std::min_element(ph + a, ph + b + 1) - ph) - 200;
}

// ---------------------------------------------------------------- reference path ---
std::vector<uint32_t> top_path(const Locus& L) {
    std::vector<uint32_t> p{0};
    while (!L.nodes[p.back()].out.empty()) p.push_back(L.nodes[p.back()].out[0]);
    return p;
}

std::vector<uint32_t> thread_sequence(const Locus& L, const std::string& seq) {
    auto fits = [&](uint32_t n, size_t off) {
        uint32_t len = L.node_len(n);
        if (off + len > seq.size()) return false;
        for (uint32_t i = 0; i < len; ++i)
            if (std::toupper((unsigned char)L.text[L.nodes[n].s + i]) != std::toupper((unsigned char)seq[off + i])) return false;
        return true;
    };
    std::vector<uint32_t> path, cursor;
    std::vector<size_t> offs;
    if (!fits(0, 0)) return {};
    path.push_back(0);
    cursor.push_back(0);
    offs.push_back(L.node_len(0));
    while (!path.empty()) {
        const LNode& nd = L.nodes[path.back()];
        if (nd.out.empty() && offs.back() == seq.size()) return path;
        if (cursor.back() >= nd.out.size()) {
            path.pop_back();
            cursor.pop_back();
            offs.pop_back();
            continue;
        }
        uint32_t c = nd.out[cursor.back()++];
        if (fits(c, offs.back())) {
            size_t o = offs.back() + L.node_len(c);
            path.push_back(c);
            cursor.push_back(0);
            offs.push_back(o);
        }
    }
    return {};
}

// ------------------------------------------------------------------- site records ---
namespace {
std::string classify(const std::string& ref, const std::string& alt) {
    if (ref.empty() && alt.empty()) return ".";
    if (ref.empty() || alt.empty()) return "INDEL";
    if (ref.size() == 1 && alt.size() == 1) return "SNP";
    if (ref.size() == alt.size()) return "PH_SNPs";
    const std::string& shorter = ref.size() < alt.size() ? ref : alt;
    const std::string& longer = ref.size() < alt.size() ? alt : ref;
    if (longer.compare(0, shorter.size(), shorter) == 0) return "INDEL";
    return "COMPLEX";
}

// k-mer nodes seen as runs of local nodes: a k-mer lies on a node path iff its node run is a
// contiguous piece of it (every node occurs at most once on a path of the DAG)
struct KnodeRuns {
    std::vector<std::vector<uint32_t>> nodes_of;  // per rank, trailing terminus padding removed
    std::vector<uint32_t> first_off, length;
    std::vector<std::vector<uint32_t>> starting_in;  // per local node: ranks whose run starts there
    explicit KnodeRuns(const Locus& L) {
        const uint32_t N = (uint32_t)L.kpath.size();
        nodes_of.resize(N);
        first_off.assign(N, 0);
        length.assign(N, 0);
        starting_in.resize(L.nodes.size());
        for (uint32_t r = 1; r + 1 < N; ++r) {
            KPath p = L.kpath[r];
            while (p.size() > 1 && p.back().s == p.back().e) p.pop_back();
            for (auto& sg : p) {
                nodes_of[r].push_back(sg.node);
                length[r] += sg.e - sg.s;
            }
            first_off[r] = p[0].s - L.nodes[p[0].node].s;
            if (length[r]) starting_in[p[0].node].push_back(r);
        }
    }
};

std::vector<uint32_t> kmers_over(const Locus& L, const KnodeRuns& KR, const std::vector<uint32_t>& np, uint32_t A, uint32_t B) {
    std::vector<uint32_t> res;
    std::vector<uint32_t> cum(np.size() + 1, 0);
    for (size_t j = 0; j < np.size(); ++j) cum[j + 1] = cum[j] + L.node_len(np[j]);
    for (size_t j = 0; j < np.size(); ++j) {
        if (cum[j] >= std::max(B, A + 1)) break;
        for (uint32_t r : KR.starting_in[np[j]]) {
            const uint32_t s = cum[j] + KR.first_off[r], e = s + KR.length[r];
            const bool over = (A == B) ? (s < A && e > A) : (s < B && e > A);
            if (!over) continue;
            const auto& run = KR.nodes_of[r];
            if (j + run.size() > np.size()) continue;
            bool same = true;
            for (size_t t = 1; t < run.size() && same; ++t) same = (np[j + t] == run[t]);
            if (same) res.push_back(r);
        }
    }
    std::sort(res.begin(), res.end());
    return res;
}
}  // namespace

std::vector<SiteRecord> enumerate_sites(const HostIndex& H, uint32_t locus, const std::vector<uint32_t>& ref) {
    const Locus& L = H.loci[locus];
    std::vector<SiteRecord> out;
    if (ref.size() < 2) return out;
    KnodeRuns KR(L);
    std::vector<uint32_t> cum(ref.size() + 1, 0);
    for (size_t i = 0; i < ref.size(); ++i) cum[i + 1] = cum[i] + L.node_len(ref[i]);
    std::vector<uint32_t> open;  // indices into ref of nodes that opened a site
    bool nested = false;
    std::set<std::tuple<uint32_t, std::string, std::string>> seen;
    for (uint32_t i = 0; i + 1 < ref.size(); ++i) {
        const LNode& nd = L.nodes[ref[i]];
        if (nd.out.size() > 1) {
            open.push_back(i);
            if (open.size() > 1) nested = true;
            continue;
        }
        if (open.empty()) continue;
        const uint32_t o = open.back();
        open.pop_back();
        const uint32_t pos = cum[o + 1];
        std::string ref_seq;
        for (uint32_t j = o + 1; j <= i; ++j) ref_seq += L.node_seq(ref[j]);
        const uint32_t join = ref[i + 1];
        // every alternative route from the opening node to the node where the reference re-joins
        std::deque<std::vector<uint32_t>> work;
        for (uint32_t a : L.nodes[ref[o]].out)
            if (a != ref[o + 1]) work.push_back({a});
        const std::vector<uint32_t> ref_kn = kmers_over(L, KR, ref, pos, pos + (uint32_t)ref_seq.size());
        while (!work.empty()) {
            std::vector<uint32_t> route = std::move(work.front());
            work.pop_front();
            const LNode& tail = L.nodes[route.back()];
            if (tail.out.empty()) continue;
            if (tail.out[0] != join) {
                for (uint32_t nx : tail.out) {
                    work.push_back(route);
                    work.back().push_back(nx);
                }
                continue;
            }
            std::string alt_seq;
            for (uint32_t n : route) alt_seq += L.node_seq(n);
            if (alt_seq == ref_seq) continue;
            if (!seen.insert({pos, ref_seq, alt_seq}).second) continue;
            SiteRecord r;
            r.locus = locus;
            r.pos = pos;
            r.ref = ref_seq;
            r.alts = {alt_seq};
            r.vc = classify(ref_seq, alt_seq);
            r.graphtype = nested ? "NESTED" : "SIMPLE";
            std::vector<uint32_t> ap(ref.begin(), ref.begin() + o + 1);
            ap.insert(ap.end(), route.begin(), route.end());
            ap.insert(ap.end(), ref.begin() + i + 1, ref.end());
            r.allele_kn.push_back(ref_kn);
            r.allele_kn.push_back(kmers_over(L, KR, ap, pos, pos + (uint32_t)alt_seq.size()));
            out.push_back(std::move(r));
        }
        if (open.empty()) nested = false;
    }
    return out;
}

std::vector<uint32_t> local_path_of(const Locus& L, const std::vector<uint32_t>& kpath) {
    std::vector<uint32_t> lp;
    lp.reserve(L.nodes.size());
    for (uint32_t r : kpath) {
        const KPath& kp = L.kpath[r];
        if (kp.empty()) continue;
        const uint32_t first = kp[0].node;
        while (!lp.empty()) {
            const auto& o = L.nodes[lp.back()].out;
            if (o.empty() || !(first > o[0]) || std::find(o.begin(), o.end(), first) != o.end()) break;
            lp.push_back(o[0]);
        }
        while (!lp.empty() && first <= lp.back()) lp.pop_back();
        for (auto& sg : kp) lp.push_back(sg.node);
    }
    if (lp.empty()) return top_path(L);
    if (lp.front() != 0) {
        const uint32_t target = lp.front();
        std::vector<char> reaches(L.nodes.size(), 0);
        reaches[target] = 1;
        for (uint32_t i = target; i-- > 0;)
            for (uint32_t o : L.nodes[i].out)
                if (o <= target && reaches[o]) reaches[i] = 1;
        std::vector<uint32_t> head;
        uint32_t cur = 0;
        while (cur != target) {
            head.push_back(cur);
            uint32_t nx = UINT32_MAX;
            for (uint32_t o : L.nodes[cur].out)
                if (o <= target && reaches[o]) {
                    nx = o;
                    break;
                }
            if (nx == UINT32_MAX) break;
            cur = nx;
        }
        lp.insert(lp.begin(), head.begin(), head.end());
    }
    while (!L.nodes[lp.back()].out.empty()) lp.push_back(L.nodes[lp.back()].out[0]);
    return lp;
}

void find_ml_path_records(const HostIndex& H, uint32_t locus, const std::vector<uint32_t>& ref,
                          const std::vector<uint32_t>& sp, const SiteKeySet& known, std::vector<SiteRecord>& extra) {
    const Locus& L = H.loci[locus];
    std::vector<uint32_t> cum;
    std::unique_ptr<KnodeRuns> KR;
    size_t ri = 0, si = 0;
    while (ri < ref.size() && si < sp.size()) {
        size_t rj = ri + 1, sj = si + 1;
        while (rj < ref.size() && sj < sp.size() && ref[rj] != sp[sj]) {
            if (ref[rj] < sp[sj]) ++rj;
            else ++sj;
        }
        if (rj >= ref.size() || sj >= sp.size()) break;
        if (rj > ri + 1 || sj > si + 1) {
            std::string rs, as;
            for (size_t j = ri + 1; j < rj; ++j) rs += L.node_seq(ref[j]);
            for (size_t j = si + 1; j < sj; ++j) as += L.node_seq(sp[j]);
            if (cum.empty()) {
                cum.assign(ref.size() + 1, 0);
                for (size_t i = 0; i < ref.size(); ++i) cum[i + 1] = cum[i] + L.node_len(ref[i]);
            }
            const uint32_t pos = cum[ri + 1];
            if (!(rs.empty() && as.empty()) && rs != as && !known.count(std::make_tuple(pos, rs, as))) {
                bool dup = false;
                for (auto& r : extra) dup = dup || (r.pos == pos && r.ref == rs && r.alts[0] == as);
                if (!dup) {
                    if (!KR) KR.reset(new KnodeRuns(L));
                    SiteRecord r;
                    r.locus = locus;
                    r.pos = pos;
                    r.ref = rs;
                    r.alts = {as};
                    r.vc = "COMPLEX";
                    r.graphtype = "TOO_MANY_ALTS";
                    std::vector<uint32_t> ap(ref.begin(), ref.begin() + ri + 1);
                    ap.insert(ap.end(), sp.begin() + si + 1, sp.begin() + sj);
                    ap.insert(ap.end(), ref.begin() + rj, ref.end());
                    r.allele_kn.push_back(kmers_over(L, *KR, ref, pos, pos + (uint32_t)rs.size()));
                    r.allele_kn.push_back(kmers_over(L, *KR, ap, pos, pos + (uint32_t)as.size()));
                    extra.push_back(std::move(r));
                }
            }
        }
        ri = rj;
        si = sj;
    }
}

std::vector<SiteRecord> merge_records(const Locus& L, const std::vector<uint32_t>& ref, std::vector<SiteRecord> recs) {
    std::sort(recs.begin(), recs.end(), [](const SiteRecord& a, const SiteRecord& b) {
        if (a.pos != b.pos) return a.pos < b.pos;
        if (a.ref != b.ref) return a.ref < b.ref;
        return a.alts < b.alts;
    });
    std::vector<SiteRecord> merged;
    for (auto& r : recs) {
        if (!merged.empty() && merged.back().pos == r.pos && merged.back().ref == r.ref &&
            merged.back().graphtype != "TOO_MANY_ALTS" && r.graphtype != "TOO_MANY_ALTS") {
            merged.back().alts.push_back(r.alts[0]);
            merged.back().allele_kn.push_back(r.allele_kn[1]);
        } else {
            merged.push_back(r);
        }
    }
    std::string refseq;
    for (uint32_t n : ref) refseq += L.node_seq(n);
    for (auto& r : merged) {
        bool any_empty = r.ref.empty();
        for (auto& a : r.alts) any_empty = any_empty || a.empty();
        if (!any_empty) continue;
        if (r.pos > 0) {
            const char anchor = refseq[r.pos - 1];
            r.pos -= 1;
            r.ref.insert(r.ref.begin(), anchor);
            for (auto& a : r.alts) a.insert(a.begin(), anchor);
        } else if (r.pos + r.ref.size() < refseq.size()) {
            const char anchor = refseq[r.pos + r.ref.size()];
            r.ref.push_back(anchor);
            for (auto& a : r.alts) a.push_back(anchor);
        }
    }
    return merged;
}

bool locus_coverage_outlier(const HostIndex& H, uint32_t locus, const std::vector<uint32_t>& kpath,
                            const std::vector<uint32_t>& lpath, const int32_t* cov, uint32_t global_covg) {
    if (global_covg <= 20) return false;  // pandora only applies the filter above 20x
    const Locus& L = H.loci[locus];
    const uint32_t base = H.knode_base[locus];
    // per-base coverage along the local path = max over the ML k-mers covering the base (one flat array)
    std::vector<int64_t> node_off(L.nodes.size(), -1);
    size_t nbases = 0;
    for (uint32_t n : lpath) {
        node_off[n] = (int64_t)nbases;
        nbases += L.node_len(n);
    }
    if (!nbases) return false;
    std::vector<uint32_t> per_base(nbases, 0);
    uint32_t top = 0;
    for (uint32_t r : kpath) {
        const uint32_t g = base + r;
        const uint32_t c = sat16(cov[2 * g]) + sat16(cov[2 * g + 1]);
        top = std::max(top, c);
        for (auto& sg : L.kpath[r]) {
            if (sg.s == sg.e || node_off[sg.node] < 0) continue;
            uint32_t* v = per_base.data() + node_off[sg.node];
            for (uint32_t x = sg.s - L.nodes[sg.node].s; x < sg.e - L.nodes[sg.node].s; ++x) v[x] = std::max(v[x], c);
        }
    }
    // mode (smallest value among ties), by counting
    std::vector<uint32_t> count(top + 1, 0);
    for (uint32_t x : per_base) ++count[x];
    uint32_t mode = 0;
    for (uint32_t x = 0; x <= top; ++x)
        if (count[x] > count[mode]) mode = x;
    return global_covg > 20 && ((uint64_t)mode * 10 < global_covg || mode > 10ull * global_covg);
}

// -------------------------------------------------------------------------- VCF ---
// "%g" (6 significant digits) like pandora's ostream output.  Fast path: scale to a 6-digit integer and print
// it in fixed notation; whenever the decimal rounding could be affected by the one floating-point rounding of the
// scaling (within 1e-6 of a tie), or the value needs exponent notation, fall back to std::to_chars(general, 6),
// which is specified to give the printf result.
size_t format_g6(double v, char* out) {
    auto slow = [&]() -> size_t {
        auto r = std::to_chars(out, out + 40, v, std::chars_format::general, 6);
        return (size_t)(r.ptr - out);
    };
    if (v == 0.0) {
        if (std::signbit(v)) return slow();
        out[0] = '0';
        return 1;
    }
    const double a = std::fabs(v);
    if (!(a >= 1e-4 && a < 999999.0)) return slow();  // exponent notation, inf, nan
    static const double P10[11] = {1e-4, 1e-3, 1e-2, 1e-1, 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6};
    int e10 = -4;
    while (a >= P10[e10 + 5]) ++e10;  // a in [10^e10, 10^(e10+1))
    static const double SC[10] = {1e9, 1e8, 1e7, 1e6, 1e5, 1e4, 1e3, 1e2, 1e1, 1e0};  // 10^(5-e10)
    const double x = a * SC[e10 + 4];
    const double fl = std::floor(x);
    const double frac = x - fl;
    if (std::fabs(frac - 0.5) < 1e-6) return slow();
    uint32_t n = (uint32_t)fl + (frac > 0.5 ? 1u : 0u);
    if (n >= 1000000u) {
        n = 100000u;
        ++e10;
        if (e10 > 5) return slow();
    }
    char d[6];
    for (int i = 5; i >= 0; --i) {
        d[i] = (char)('0' + n % 10);
        n /= 10;
    }
    int last = 5;
    while (last > 0 && d[last] == '0') --last;  // significant digits d[0..last]
    char* o = out;
    if (v < 0) *o++ = '-';
    if (e10 >= 0) {
        for (int i = 0; i <= e10; ++i) *o++ = d[i];  // integer part (zeros included)
        if (last > e10) {
            *o++ = '.';
            for (int i = e10 + 1; i <= last; ++i) *o++ = d[i];
        }
    } else {
        *o++ = '0';
        *o++ = '.';
        for (int i = 0; i < -e10 - 1; ++i) *o++ = '0';
        for (int i = 0; i <= last; ++i) *o++ = d[i];
    }
    return (size_t)(o - out);
}

void format_vcf_header(const std::vector<std::string>& contigs, const std::string& sample, std::string& s) {
    s.clear();  // keeps its capacity: a fresh ~1 MB string would be mmap'ed and page-faulted in on every sample
    char date[32];
    time_t t = time(nullptr);
    strftime(date, sizeof date, "%d/%m/%y", localtime(&t));
    s += "##fileformat=VCFv4.3\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n##fileDate==";
    s += date;
    s += "\n##ALT=<ID=SNP,Description=\"SNP\">\n##ALT=<ID=PH_SNPs,Description=\"Phased SNPs\">\n"
         "##ALT=<ID=INDEL,Description=\"Insertion-deletion\">\n"
         "##ALT=<ID=COMPLEX,Description=\"Complex variant, collection of SNPs and indels\">\n"
         "##INFO=<ID=VC,Number=1,Type=String,Description=\"Type (class) of variant\">\n"
         "##ALT=<ID=SIMPLE,Description=\"Graph bubble is simple\">\n"
         "##ALT=<ID=NESTED,Description=\"Variation site was a nested feature in the graph\">\n"
         "##ALT=<ID=TOO_MANY_ALTS,Description=\"Variation site was a multinested feature with too many alts to include all in the VCF\">\n"
         "##INFO=<ID=GRAPHTYPE,Number=1,Type=String,Description=\"Type of graph feature\">\n"
         "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n";
    static const char* tags[6] = {"MEAN_FWD_COVG", "MEAN_REV_COVG", "MED_FWD_COVG", "MED_REV_COVG", "SUM_FWD_COVG", "SUM_REV_COVG"};
    static const char* desc[6] = {"Mean forward coverage", "Mean reverse coverage", "Med forward coverage",
                                  "Med reverse coverage", "Sum forward coverage", "Sum reverse coverage"};
    for (int i = 0; i < 6; ++i)
        s += std::string("##FORMAT=<ID=") + tags[i] + ",Number=R,Type=Integer,Description=\"" + desc[i] + "\">\n";
    s += "##FORMAT=<ID=GAPS,Number=R,Type=Float,Description=\"Number of gap bases\">\n"
         "##FORMAT=<ID=LIKELIHOOD,Number=R,Type=Float,Description=\"Likelihood\">\n"
         "##FORMAT=<ID=GT_CONF,Number=1,Type=Float,Description=\"Genotype confidence\">\n";
    for (auto& c : contigs) s += "##contig=<ID=" + c + ">\n";
    s += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + sample + "\n";
}

// CHROM .. FORMAT columns of a record (they never change for a site: formatted once and cached)
const std::string& vcf_record_prefix(const HostIndex& H, const SiteRecord& r) {
    if (r.text_prefix.empty()) {
        std::string& t = r.text_prefix;
        t = H.loci[r.locus].name + "\t" + std::to_string(r.pos + 1) + "\t.\t" + (r.ref.empty() ? "." : r.ref) + "\t";
        for (size_t a = 0; a < r.alts.size(); ++a) t += (a ? "," : "") + (r.alts[a].empty() ? std::string(".") : r.alts[a]);
        t += "\t.\t.\tVC=" + r.vc + ";GRAPHTYPE=" + r.graphtype +
             "\tGT:MEAN_FWD_COVG:MEAN_REV_COVG:MED_FWD_COVG:MED_REV_COVG:SUM_FWD_COVG:SUM_REV_COVG:GAPS:LIKELIHOOD:GT_CONF\t";
    }
    return r.text_prefix;
}

void format_vcf(const HostIndex& H, const std::vector<const SiteRecord*>& recs, const GenotypeArrays& G,
                const std::vector<std::string>& contigs, const std::string& sample, std::string& s) {
    format_vcf_header(contigs, sample, s);
    s.reserve(4096 + recs.size() * 256);
    // Records are independent lines.  Each worker formats a contiguous range with raw pointer writes into its own
    // buffer (an upper bound on the line length is known: prefix + 12 bytes per integer + 26 per float), then the
    // ranges are copied into place in parallel.
    auto put_uint = [](char* o, uint32_t v) -> char* {
        char tmp[10];
        int n = 0;
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) *o++ = tmp[--n];
        return o;
    };
    auto line_bound = [&](size_t i) -> size_t {
        const SiteRecord& r = *recs[i];
        const size_t na = G.rec_off[i + 1] - G.rec_off[i];
        size_t prefix = r.text_prefix.size();
        if (!prefix) {
            prefix = H.loci[r.locus].name.size() + r.ref.size() + 160;
            for (auto& a : r.alts) prefix += a.size() + 1;
        }
        return prefix + 16 + na * (6 * 12 + 2 * 28) + 40;
    };
    auto format_range = [&](size_t lo, size_t hi, std::vector<char>& buf) -> size_t {
        size_t bound = 0;
        for (size_t i = lo; i < hi; ++i) bound += line_bound(i);
        buf.resize(bound);
        char* o = buf.data();
        const std::vector<uint32_t>* cols[6] = {&G.mean_fwd, &G.mean_rev, &G.med_fwd, &G.med_rev, &G.sum_fwd, &G.sum_rev};
        for (size_t i = lo; i < hi; ++i) {
            const SiteRecord& r = *recs[i];
            const uint32_t b = G.rec_off[i], e = G.rec_off[i + 1];
            vcf_record_prefix(H, r);
            memcpy(o, r.text_prefix.data(), r.text_prefix.size());
            o += r.text_prefix.size();
            if (G.gt[i] < 0) *o++ = '.';
            else o = put_uint(o, (uint32_t)G.gt[i]);
            for (auto* col : cols) {
                *o++ = ':';
                for (uint32_t a = b; a < e; ++a) {
                    if (a > b) *o++ = ',';
                    o = put_uint(o, (*col)[a]);
                }
            }
            *o++ = ':';
            for (uint32_t a = b; a < e; ++a) {
                if (a > b) *o++ = ',';
                o += format_g6(G.gaps[a], o);
            }
            *o++ = ':';
            for (uint32_t a = b; a < e; ++a) {
                if (a > b) *o++ = ',';
                o += format_g6(G.lik[a], o);
            }
            *o++ = ':';
            o += format_g6(G.gt_conf[i], o);
            *o++ = '\n';
        }
        return (size_t)(o - buf.data());
    };
    static const bool timing = getenv("DRPRG_TIMING") != nullptr;
    auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tq0 = timing ? now_us() : 0;
    const size_t parts_n = recs.size() >= 512 ? 16 : 1;
    static thread_local std::vector<std::vector<char>> parts_tls;  // reused across samples (same reason as `s`)
    std::vector<std::vector<char>>& parts = parts_tls;  // the workers must see THIS thread's buffers, not their own
    if (parts.size() < parts_n) parts.resize(parts_n);
    std::vector<size_t> used(parts_n, 0);
    parallel_for(parts_n, [&](size_t t) {
        used[t] = format_range(recs.size() * t / parts_n, recs.size() * (t + 1) / parts_n, parts[t]);
    });
    const double tq1 = timing ? now_us() : 0;
    // plain appends: resize() would zero-fill the text first and a second parallel section costs more than the copy
    for (size_t t = 0; t < parts_n; ++t) s.append(parts[t].data(), used[t]);
    if (timing) fprintf(stderr, "[drprg-cuda] vcf text: format %.1f us, gather %.1f us\n", tq1 - tq0, now_us() - tq1);
}

// ---------------------------------------------------------------------------- IO ---
namespace {
struct GzLines {
    gzFile f;
    std::vector<char> buf;
    explicit GzLines(const std::string& path) : buf(1 << 20) {
        f = gzopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("cannot open " + path);
        gzbuffer(f, 1 << 20);
    }
    ~GzLines() { gzclose(f); }
    bool next(std::string& s) {
        s.clear();
        while (gzgets(f, buf.data(), (int)buf.size())) {
            s += buf.data();
            if (!s.empty() && s.back() == '\n') break;
        }
        if (s.empty()) return false;
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        return true;
    }
};
}  // namespace

std::map<std::string, std::string> load_fasta(const std::string& path) {
    std::map<std::string, std::string> m;
    GzLines in(path);
    std::string line, name;
    while (in.next(line)) {
        if (line.empty()) continue;
        if (line[0] == '>') {
            name = line.substr(1, line.find_first_of(" \t") == std::string::npos ? std::string::npos : line.find_first_of(" \t") - 1);
            m[name];
        } else if (!name.empty()) {
            m[name] += line;
        }
    }
    return m;
}

static inline uint32_t code4(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return 4;
}

static void pack_one(const uint8_t* s, uint32_t len, uint32_t* w, uint32_t& out_len) {
    uint32_t bad = 0;
    const uint32_t nw = (len + 15) / 16;
    for (uint32_t j = 0; j < nw; ++j) {
        uint32_t word = 0;
        const uint32_t hi = std::min(len, (j + 1) * 16);
        for (uint32_t i = j * 16; i < hi; ++i) {
            uint32_t c = code4(s[i]);
            bad |= c >> 2;
            word |= (c & 3u) << (30 - 2 * (i & 15));
        }
        w[j] = word;
    }
    out_len = bad ? 0 : len;
}

int64_t pack_ascii(const uint8_t* ascii, const uint64_t* off, uint64_t n, uint32_t stride_words, uint32_t* words,
                   uint64_t words_cap, uint64_t* word_off, uint32_t* lens) {
    uint64_t cur = 0;
    for (uint64_t r = 0; r < n; ++r) {
        const uint32_t len = (uint32_t)(off[r + 1] - off[r]);
        const uint32_t nw = (len + 15) / 16;
        if (stride_words) {
            if (nw > stride_words) return -1;
            cur = r * (uint64_t)stride_words;
        }
        if (word_off) word_off[r] = cur;
        if (cur + std::max(nw, stride_words) > words_cap) return -2;
        if (stride_words) std::memset(words + cur, 0, (size_t)stride_words * 4);
        pack_one(ascii + off[r], len, words + cur, lens[r]);
        cur += stride_words ? stride_words : nw;
    }
    if (word_off) word_off[n] = stride_words ? n * (uint64_t)stride_words : cur;
    return (int64_t)(stride_words ? n * (uint64_t)stride_words : cur);
}

// Ingest: the whole file is brought into memory (gzip inflated on one thread — a gzip stream is serial), split
// at record boundaries into one piece per worker, parsed + 2-bit packed in parallel, then stitched together.
// Inputs: fasta/fastq, plain or gzip (/root/reference/src/predict.rs:166-170).
namespace {
struct RawText {  // whole input file in memory, no zero-initialisation, no growth copies for plain files
    char* p = nullptr;
    size_t n = 0;
    ~RawText() { free(p); }
    size_t size() const { return n; }
    const char* data() const { return p; }
    char operator[](size_t i) const { return p[i]; }
};

void slurp(const std::string& path, RawText& buf) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("cannot open " + path);
    unsigned char magic[2] = {0, 0};
    size_t got = fread(magic, 1, 2, fp);
    fseek(fp, 0, SEEK_END);
    const size_t fsize = (size_t)ftell(fp);
    fseek(fp, 0, SEEK_SET);
    const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (!gz) {
        buf.p = (char*)malloc(fsize + 1);
        if (!buf.p) { fclose(fp); throw std::runtime_error("out of memory reading " + path); }
        buf.n = fread(buf.p, 1, fsize, fp);
        fclose(fp);
        return;
    }
    {   // single-stream gzip: inflate on all host threads (gzip_inflate.cpp); anything it declines goes through zlib below
        static const bool par = [] {
            const char* e = getenv("DRPRG_PARALLEL_GZIP");
            return !e || atoi(e) != 0;
        }();
        static const size_t min_size = [] {
            const char* e = getenv("DRPRG_PARALLEL_GZIP_CHUNK");
            return e && atol(e) > 0 ? (size_t)atol(e) * 2 : (size_t)(4u << 20);
        }();
        if (par && fsize >= min_size) {
            uint8_t* z = (uint8_t*)malloc(fsize + 64);  // the bit reader loads 8 bytes at a time: zero padding behind the image
            bool got_all = false;
            if (z) {
                memset(z + fsize, 0, 64);
                // the compressed image is read by several threads at once (page-cache copies are memcpy-bound)
                const int fd = fileno(fp);
                const size_t RT = std::max<size_t>(1, std::min<size_t>(8, fsize / (8u << 20)));
                std::vector<char> okv(RT, 0);
                parallel_for(RT, [&](size_t t) {
                    const size_t lo = fsize * t / RT, hi = fsize * (t + 1) / RT;
                    size_t done = 0;
                    while (lo + done < hi) {
                        const ssize_t r = pread(fd, z + lo + done, hi - lo - done, (off_t)(lo + done));
                        if (r <= 0) break;
                        done += (size_t)r;
                    }
                    okv[t] = (lo + done == hi);
                }, RT);
                got_all = std::all_of(okv.begin(), okv.end(), [](char c) { return c != 0; });
            }
            if (z && got_all) {
                const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
                char* text = nullptr;
                size_t tn = 0;
                const bool ok = parallel_gunzip(z, fsize, hw, &text, &tn);
                free(z);
                if (ok) {
                    fclose(fp);
                    buf.p = text;
                    buf.n = tn;
                    return;
                }
            } else {
                free(z);
            }
        }
    }
    fclose(fp);
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1 << 20);
    size_t cap = std::max<size_t>(fsize * 5, 1 << 20);
    buf.p = (char*)malloc(cap);
    size_t n = 0;
    while (buf.p) {
        if (n == cap) {
            cap *= 2;
            char* q = (char*)realloc(buf.p, cap);
            if (!q) break;
            buf.p = q;
        }
        int r = gzread(f, buf.p + n, (unsigned)std::min<size_t>(cap - n, 1u << 30));
        if (r < 0) {
            gzclose(f);
            throw std::runtime_error("read error in " + path);
        }
        if (r == 0) break;
        n += (size_t)r;
    }
    gzclose(f);
    if (!buf.p) throw std::runtime_error("out of memory reading " + path);
    buf.n = n;
}

}  // namespace
// whole file inflated into one malloc'ed buffer (the batch driver runs this ahead of the GPU on spare threads)
bool inflate_file(const std::string& path, char** out, size_t* n) {
    RawText buf;
    slurp(path, buf);
    *out = buf.p;
    *n = buf.n;
    buf.p = nullptr;
    return true;
}
bool file_is_gzip(const std::string& path) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    unsigned char m[2] = {0, 0};
    const size_t got = fread(m, 1, 2, fp);
    fclose(fp);
    return got == 2 && m[0] == 0x1f && m[1] == 0x8b;
}
namespace {
struct Piece {
    std::vector<uint32_t> words, lens;
    std::vector<uint32_t> nwords;  // per read
    uint64_t total_bases = 0, n_dropped = 0;
    uint32_t first_raw_len = 0;  // length of the piece's first read before the non-ACGT drop
};

inline size_t line_end(const RawText& b, size_t p) {
    const void* q = memchr(b.data() + p, '\n', b.size() - p);
    return q ? (size_t)((const char*)q - b.data()) : b.size();
}

// first record start at or after p
size_t next_record(const RawText& b, size_t p, bool fastq) {
    if (p == 0) return 0;
    p = line_end(b, p - 1);  // end of the line containing p-1
    if (p >= b.size()) return b.size();
    ++p;
    while (p < b.size()) {
        if (!fastq) {
            if (b[p] == '>') return p;
        } else if (b[p] == '@') {  // a header line is followed, two lines later, by a '+' line
            size_t l1 = line_end(b, p);
            size_t l2 = l1 < b.size() ? line_end(b, l1 + 1) : b.size();
            if (l2 + 1 < b.size() && b[l2 + 1] == '+') return p;
        }
        p = line_end(b, p);
        if (p >= b.size()) return b.size();
        ++p;
    }
    return b.size();
}

void parse_piece(const RawText& b, size_t lo, size_t hi, bool fastq, Piece& out) {
    std::string seq;
    out.words.reserve((hi - lo) / 8 + 16);  // FASTQ: ~2 text bytes per base, 4 bases per byte
    out.lens.reserve((hi - lo) / 128 + 16);
    out.nwords.reserve((hi - lo) / 128 + 16);
    auto emit = [&](const char* s, uint32_t len) {
        const size_t at = out.words.size();
        const uint32_t nw = (len + 15) / 16;
        out.words.resize(at + nw);
        uint32_t l = 0;
        pack_one((const uint8_t*)s, len, out.words.data() + at, l);
        if (l == 0 && len > 0) ++out.n_dropped;
        if (out.lens.empty()) out.first_raw_len = len;
        out.lens.push_back(l);
        out.nwords.push_back(nw);
        out.total_bases += len;
    };
    size_t p = lo;
    while (p < hi) {
        size_t e = line_end(b, p);
        if (e == p) {  // blank line
            p = e + 1;
            continue;
        }
        if (fastq) {
            // kseq-style record (what pandora's reader accepts): '@' header, sequence lines up to the '+' line, then
            // quality lines until they are as long as the sequence.  Strict 4-line records take the first branch.
            if (b[p] != '@') throw std::runtime_error("malformed FASTQ record");
            size_t s0 = e + 1, s1 = s0 < b.size() ? line_end(b, s0) : b.size();
            size_t len = s1 > s0 ? s1 - s0 : 0;
            if (len && b[s0 + len - 1] == '\r') --len;
            size_t nxt = s1 + 1;  // start of the line after the first sequence line
            if (nxt >= b.size() || b[nxt] == '+') {
                emit(b.data() + std::min(s0, b.size()), (uint32_t)len);
            } else {  // wrapped sequence
                seq.assign(b.data() + s0, len);
                while (nxt < b.size() && b[nxt] != '+') {
                    size_t le = line_end(b, nxt);
                    size_t l2 = le - nxt;
                    if (l2 && b[nxt + l2 - 1] == '\r') --l2;
                    seq.append(b.data() + nxt, l2);
                    nxt = le + 1;
                }
                if (nxt >= b.size()) throw std::runtime_error("malformed FASTQ record: no '+' line");
                len = seq.size();
                emit(seq.data(), (uint32_t)len);
            }
            size_t q = nxt < b.size() ? line_end(b, nxt) + 1 : b.size();  // past the '+' line
            size_t qlen = 0;
            bool first_qual = true;
            while (q < b.size() && (first_qual || qlen < len)) {
                size_t le = line_end(b, q);
                size_t l2 = le - q;
                if (l2 && b[q + l2 - 1] == '\r') --l2;
                qlen += l2;
                q = le + 1;
                first_qual = false;
            }
            if (qlen > len && len) throw std::runtime_error("malformed FASTQ record: quality longer than sequence");
            p = q;
        } else {
            if (b[p] != '>') throw std::runtime_error("malformed FASTA record");
            p = e + 1;
            seq.clear();
            while (p < hi && b[p] != '>') {
                size_t le = line_end(b, p);
                size_t len = le - p;
                if (len && b[p + len - 1] == '\r') --len;
                seq.append(b.data() + p, len);
                p = le + 1;
            }
            emit(seq.data(), (uint32_t)seq.size());
        }
    }
}
}  // namespace

void load_reads_packed(const std::string& path, uint32_t threads, PackedReads& out) {
    out = PackedReads();
    RawText buf;
    slurp(path, buf);
    size_t first = 0;
    while (first < buf.size() && (buf[first] == '\n' || buf[first] == '\r')) ++first;
    out.word_off.assign(1, 0);
    if (first >= buf.size()) return;
    const bool fastq = buf[first] == '@';
    if (!fastq && buf[first] != '>') throw std::runtime_error("unrecognised read file format: " + path);
    size_t T = std::max<size_t>(1, std::min<size_t>({(size_t)std::max(1u, threads), (size_t)16, buf.size() / (1 << 20) + 1}));
    if (fastq && T > 1) {
        // the parallel cut finds record starts by the strict 4-line pattern ('@' line, '+' two lines later); a wrapped
        // (multi-line) FASTQ is parsed by one thread instead.  Probe the first records.
        size_t p = first;
        for (int rec = 0; rec < 64 && p < buf.size() && T > 1; ++rec) {
            size_t l1 = line_end(buf, p), l2 = l1 < buf.size() ? line_end(buf, l1 + 1) : buf.size();
            size_t l3 = l2 < buf.size() ? line_end(buf, l2 + 1) : buf.size(), l4 = l3 < buf.size() ? line_end(buf, l3 + 1) : buf.size();
            if (buf[p] != '@' || l2 + 1 >= buf.size() || buf[l2 + 1] != '+') T = 1;
            p = l4 + 1;
        }
    }
    std::vector<size_t> cut;
    std::vector<Piece> pieces;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cut.assign(T + 1, buf.size());
        cut[0] = first;
        for (size_t t = 1; t < T; ++t) cut[t] = std::max(cut[t - 1], next_record(buf, first + (buf.size() - first) * t / T, fastq));
        pieces.assign(T, Piece());
        try {
            parallel_for(T, [&](size_t t) { parse_piece(buf, cut[t], cut[t + 1], fastq, pieces[t]); }, T);
            break;
        } catch (const std::exception&) {
            if (T == 1 || attempt) throw;
            T = 1;  // records that only look irregular from a mid-file cut: one thread, from the top
        }
    }
    // stitch
    std::vector<uint64_t> rbase(T + 1, 0), wbase(T + 1, 0);
    for (size_t t = 0; t < T; ++t) {
        rbase[t + 1] = rbase[t] + pieces[t].lens.size();
        wbase[t + 1] = wbase[t] + pieces[t].words.size();
        out.total_bases += pieces[t].total_bases;
        out.n_dropped += pieces[t].n_dropped;
    }
    out.words.resize(wbase[T]);
    out.lens.resize(rbase[T]);
    out.word_off.resize(rbase[T] + 1);
    parallel_for(T, [&](size_t t) {
        const Piece& P = pieces[t];
        if (!P.words.empty()) memcpy(out.words.data() + wbase[t], P.words.data(), P.words.size() * 4);
        if (!P.lens.empty()) memcpy(out.lens.data() + rbase[t], P.lens.data(), P.lens.size() * 4);
        uint64_t w = wbase[t];
        for (size_t i = 0; i < P.nwords.size(); ++i) {
            out.word_off[rbase[t] + i] = w;
            w += P.nwords[i];
        }
    }, T);
    out.word_off[rbase[T]] = wbase[T];
    // pandora's short-read cluster threshold uses the length of the first read
    for (size_t t = 0; t < T; ++t)
        if (!pieces[t].nwords.empty()) {
            out.first_read_len = pieces[t].first_raw_len;
            break;
        }
}

}  // namespace drprg
