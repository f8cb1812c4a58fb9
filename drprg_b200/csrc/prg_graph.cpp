// PRG text -> local graph -> graph (w,k)-minimizers -> rank-ordered k-mer graphs + records.
// Behaviour follows pandora's LocalPRG (build_graph / minimizer_sketch / shift), LocalGraph::walk
// and KmerGraph (sorted_nodes, remove_shortcut_edges) — the index `pandora index` builds for
// drprg at /root/reference/src/lib.rs:479-510 and `pandora map` consumes at :580-642.
// PRG grammar: /root/reference/tests/cases/expected/dr.prg.
#include "prg_graph.hpp"

#include <algorithm>
#include <deque>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace drprg {

uint64_t hash64_host(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key ^= key >> 24;
    key = (key * 265) & mask;
    key ^= key >> 14;
    key = (key * 21) & mask;
    key ^= key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

// hash64 is a bijection on [0, mask]: every step (odd multiply, xor-shift, x -> x*(2^21-1) - 1) is invertible
// modulo mask + 1.  The device k-mer screen needs the k-mer behind every indexed hash.
uint64_t hash64_inverse_host(uint64_t key, uint64_t mask) {
    auto inv_odd = [](uint64_t a) {  // Newton iteration for a^-1 mod 2^64
        uint64_t x = a;
        for (int i = 0; i < 6; ++i) x *= 2 - a * x;
        return x;
    };
    auto unxorshift = [&](uint64_t v, int s) {
        uint64_t x = v;
        for (int i = 0; i * s < 64; ++i) x = v ^ (x >> s);
        return x;
    };
    key &= mask;
    key = (key * inv_odd((1ull << 31) + 1)) & mask;
    key = unxorshift(key, 28);
    key = (key * inv_odd(21)) & mask;
    key = unxorshift(key, 14);
    key = (key * inv_odd(265)) & mask;
    key = unxorshift(key, 24);
    key = ((key + 1) * inv_odd((1ull << 21) - 1)) & mask;
    return key;
}

std::string read_text_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// ------------------------------------------------------------------------------ parsing ---
namespace {
struct SiteFrame {
    uint32_t site;
    uint32_t pre;
    std::vector<uint32_t> ends;
};

void build_local_graph(Locus& L) {
    const std::string& t = L.text;
    std::vector<SiteFrame> stack;
    std::vector<uint32_t> from;
    bool expect_seq = true;
    size_t i = 0;
    const size_t n = t.size();
    auto fail = [&](const char* what) { throw std::runtime_error(std::string("PRG parse error in ") + L.name + ": " + what); };
    while (true) {
        size_t j = i;
        while (j < n && t[j] != ' ') ++j;
        bool is_marker = (j > i) && (t[i] >= '0' && t[i] <= '9');
        if (!is_marker) {
            if (!expect_seq) fail("two sequence tokens in a row");
            LNode nd{(uint32_t)i, (uint32_t)j, {}};
            uint32_t id = (uint32_t)L.nodes.size();
            L.nodes.push_back(nd);
            for (uint32_t f : from) L.nodes[f].out.push_back(id);
            from.assign(1, id);
            expect_seq = false;
        } else {
            if (expect_seq) fail("marker where a sequence token was expected");
            uint32_t m = (uint32_t)std::stoul(t.substr(i, j - i));
            if (m % 2 == 1) {
                if (!stack.empty() && stack.back().site == m) {  // closes site m
                    SiteFrame fr = stack.back();
                    stack.pop_back();
                    fr.ends.insert(fr.ends.end(), from.begin(), from.end());
                    from = fr.ends;
                } else {  // opens site m
                    if (from.size() != 1) fail("site must open after a sequence token");
                    stack.push_back(SiteFrame{m, from[0], {}});
                }
            } else {
                if (stack.empty() || stack.back().site + 1 != m) fail("allele separator outside its site");
                SiteFrame& fr = stack.back();
                fr.ends.insert(fr.ends.end(), from.begin(), from.end());
                from.assign(1, fr.pre);
            }
            expect_seq = true;
        }
        if (j >= n) break;
        i = j + 1;
    }
    if (!stack.empty()) fail("unclosed site");
    if (expect_seq) fail("PRG ends with a marker");
}

inline int code_of(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
    }
    return 4;
}

// ------------------------------------------------------------------- path primitives ---
uint32_t plen(const KPath& p) {
    uint32_t n = 0;
    for (auto& s : p) n += s.e - s.s;
    return n;
}

// bases [start, start+len) along p; zero-length stretches strictly inside or exactly at the
// start boundary are kept, none after the last base (pandora prg::Path::subpath)
KPath sub(const KPath& p, uint32_t start, uint32_t len) {
    KPath out;
    uint32_t covered = 0, added = 0;
    for (const Seg& sg : p) {
        uint32_t l = sg.e - sg.s;
        if (out.empty()) {
            bool holds_start = covered <= start && covered + l > start;
            bool empty_at_start = covered == start && l == 0;
            if (holds_start || empty_at_start) {
                uint32_t s = sg.s + (start - covered);
                uint32_t e = std::min(sg.e, s + (len - added));
                out.push_back(Seg{sg.node, s, e});
                added += e - s;
            }
        } else if (covered >= start && covered <= start + len) {
            uint32_t e = std::min(sg.e, sg.s + (len - added));
            out.push_back(Seg{sg.node, sg.s, e});
            added += e - sg.s;
        }
        covered += l;
        if (!out.empty() && added >= len) break;
    }
    return out;
}

bool kpath_less(const KPath& a, const KPath& b) {
    size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; ++i) {
        if (a[i].s != b[i].s) return a[i].s < b[i].s;
        uint32_t la = a[i].e - a[i].s, lb = b[i].e - b[i].s;
        if (la != lb) return la < lb;
    }
    return a.size() < b.size();
}
struct KPathLess {
    bool operator()(const KPath& a, const KPath& b) const { return kpath_less(a, b); }
};

struct Sketcher {
    Locus& L;
    uint32_t w, k;
    uint64_t mask;
    uint32_t END;

    // all walks of exactly `len` bases from (node, pos); iterative DFS, out-edge order
    std::vector<KPath> walk(uint32_t node, uint32_t pos, uint32_t len) const {
        std::vector<KPath> res;
        if (len == 0) return res;
        struct Fr {
            uint32_t node, pos, need;
            size_t child;
        };
        std::vector<Fr> st;
        KPath cur;
        st.push_back({node, pos, len, 0});
        while (!st.empty()) {
            Fr& f = st.back();
            const LNode& nd = L.nodes[f.node];
            if (f.child == 0) {
                if (f.pos + f.need <= nd.e) {
                    cur.push_back(Seg{f.node, f.pos, f.pos + f.need});
                    res.push_back(cur);
                    cur.pop_back();
                    st.pop_back();
                    continue;
                }
                cur.push_back(Seg{f.node, f.pos, nd.e});
            }
            if (f.child < nd.out.size()) {
                uint32_t o = nd.out[f.child++];
                uint32_t need = f.need - (nd.e - f.pos);
                st.push_back({o, L.nodes[o].s, need, 0});
            } else {
                cur.pop_back();
                st.pop_back();
            }
        }
        return res;
    }

    void hash_of(const KPath& p, uint64_t& h, uint8_t& strand) const {
        uint64_t f = 0, r = 0;
        const uint64_t sh = 2 * (k - 1);
        for (const Seg& sg : p)
            for (uint32_t x = sg.s; x < sg.e; ++x) {
                int c = code_of(L.text[x]);
                if (c > 3) continue;
                f = ((f << 2) | (uint64_t)c) & mask;
                r = (r >> 2) | ((uint64_t)(3 ^ c) << sh);
            }
        uint64_t hf = hash64_host(f, mask), hr = hash64_host(r, mask);
        h = std::min(hf, hr);
        strand = hf <= hr;
    }

    // every path of the same length one base further along the graph (pandora LocalPRG::shift)
    std::vector<KPath> shift(const KPath& p0) const {
        std::vector<KPath> res, grown;
        uint32_t len = plen(p0);
        if (len == 0) return res;
        std::deque<KPath> q;
        q.push_back(sub(p0, 1, len - 1));
        while (!q.empty()) {
            KPath p = std::move(q.front());
            q.pop_front();
            if (p.empty()) continue;
            const Seg last = p.back();
            const LNode& nd = L.nodes[last.node];
            if (last.e < nd.e) {
                p.back().e += 1;
                grown.push_back(std::move(p));
            } else if (last.e != END) {
                for (uint32_t o : nd.out) {
                    KPath e = p;
                    e.push_back(Seg{o, L.nodes[o].s, L.nodes[o].s});
                    q.push_back(std::move(e));
                }
            }
        }
        for (KPath& g : grown) {
            bool non_terminal = false;
            std::deque<KPath> t;
            t.push_back(g);
            while (!t.empty()) {
                KPath p = std::move(t.front());
                t.pop_front();
                const Seg last = p.back();
                const LNode& nd = L.nodes[last.node];
                if (nd.e == END) {
                    res.push_back(std::move(p));
                } else if (nd.e == last.e) {
                    for (uint32_t o : nd.out) {
                        if (L.nodes[o].s == L.nodes[o].e) {
                            KPath e = p;
                            e.push_back(Seg{o, L.nodes[o].s, L.nodes[o].e});
                            t.push_back(std::move(e));
                        } else {
                            non_terminal = true;
                        }
                    }
                } else {
                    non_terminal = true;
                }
            }
            if (non_terminal) res.push_back(std::move(g));
        }
        return res;
    }
};

// does `a` followed by `c` describe one consistent stretch of the graph?  If so return it.
KPath join(const KPath& a, const KPath& c) {
    KPath u;
    if (a.empty() || c.empty()) return u;
    if (a.back().e < c.front().s) return u;
    size_t i = 0;
    while (i < a.size() && !(a[i].node == c[0].node && a[i].s <= c[0].s && c[0].s <= a[i].e)) ++i;
    if (i == a.size()) return u;
    // the rest of a must run along the front of c
    size_t m = a.size() - i;
    if (m > c.size()) return u;
    for (size_t t = 1; t < m; ++t) {
        const Seg &x = a[i + t], &y = c[t];
        bool lastx = (i + t + 1 == a.size());
        if (x.node != y.node || x.s != y.s) return KPath();
        if (lastx ? (x.e > y.e) : (x.e != y.e)) return KPath();
    }
    if (m > 1 && a[i].e != c[0].e) return KPath();
    if (m == 1 && a[i].e > c[0].e && c.size() > 1) return KPath();
    u.assign(a.begin(), a.begin() + i);
    u.push_back(Seg{a[i].node, a[i].s, std::max(a[i].e, c[0].e)});
    u.insert(u.end(), c.begin() + 1, c.end());
    return u;
}

// is b a contiguous stretch of u?
bool lies_on(const KPath& b, const KPath& u) {
    if (b.empty() || u.empty()) return false;
    for (size_t j = 0; j < u.size(); ++j) {
        if (u[j].node != b[0].node) continue;
        if (b[0].s < u[j].s || b[0].e > u[j].e) return false;
        if (b.size() == 1) return true;
        if (b[0].e != u[j].e) return false;
        if (j + b.size() > u.size()) return false;
        for (size_t t = 1; t < b.size(); ++t) {
            const Seg &x = b[t], &y = u[j + t];
            bool last = (t + 1 == b.size());
            if (x.node != y.node || x.s != y.s) return false;
            if (last ? (x.e > y.e) : (x.e != y.e)) return false;
        }
        return true;
    }
    return false;
}
}  // namespace

void parse_prg_text(const std::string& text, std::vector<Locus>& loci) {
    loci.clear();
    size_t pos = 0;
    Locus cur;
    bool open = false;
    while (pos <= text.size()) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) nl = text.size();
        std::string line = text.substr(pos, nl - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            if (open) loci.push_back(std::move(cur));
            cur = Locus();
            cur.name = line.substr(1, line.find_first_of(" \t", 1) == std::string::npos ? std::string::npos : line.find_first_of(" \t", 1) - 1);
            open = true;
        } else if (open) {
            cur.text += line;
        }
        if (nl == text.size()) break;
        pos = nl + 1;
    }
    if (open) loci.push_back(std::move(cur));
    for (Locus& L : loci) build_local_graph(L);
}

void sketch_locus(Locus& L, uint32_t prg_id, uint32_t w, uint32_t k, std::vector<Record>& records) {
    Sketcher S{L, w, k, (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1), L.end_coord()};
    // creation-order store
    std::vector<KPath> paths;
    std::vector<uint64_t> hashes;
    std::vector<uint8_t> strands;
    std::vector<std::vector<uint32_t>> outs;
    std::map<KPath, uint32_t, KPathLess> ids;
    auto find_or_add = [&](const KPath& p, uint64_t h, uint8_t st, bool& created) {
        auto it = ids.find(p);
        created = (it == ids.end());
        if (!created) return it->second;
        uint32_t id = (uint32_t)paths.size();
        ids.emplace(p, id);
        paths.push_back(p);
        hashes.push_back(h);
        strands.push_back(st);
        outs.emplace_back();
        return id;
    };
    auto link = [&](uint32_t a, uint32_t b) {
        if (a == b) return;
        auto& o = outs[a];
        if (std::find(o.begin(), o.end(), b) == o.end()) o.push_back(b);
    };
    bool created;
    find_or_add(KPath{Seg{0, 0, 0}}, UINT64_MAX, 1, created);  // null start

    std::deque<uint32_t> leaves;
    std::vector<uint32_t> end_leaves;
    std::vector<char> queued;
    auto enqueue = [&](uint32_t id, const KPath& window_last) {
        if (window_last.back().e == S.END) {
            end_leaves.push_back(id);
            return;
        }
        if (queued.size() < paths.size()) queued.resize(paths.size(), 0);
        if (!queued[id]) {
            queued[id] = 1;
            leaves.push_back(id);
        }
    };

    const bool trivial = L.nodes.size() == 1 && L.node_len(0) < k;
    if (!trivial) {
        for (const KPath& wp : S.walk(0, 0, w + k - 1)) {
            std::vector<KPath> km(w);
            std::vector<uint64_t> hs(w);
            std::vector<uint8_t> st(w);
            uint64_t best = UINT64_MAX;
            for (uint32_t j = 0; j < w; ++j) {
                km[j] = sub(wp, j, k);
                S.hash_of(km[j], hs[j], st[j]);
                best = std::min(best, hs[j]);
            }
            bool have_edge = false;
            for (uint32_t j = 0; j < w; ++j) {
                if (hs[j] != best) continue;
                KPath p = km[j];
                // a k-mer that can go no further reaches the terminus through trailing empty nodes
                uint32_t ln = p.back().node;
                if (S.walk(ln, L.nodes[ln].e, w + k - 1).empty()) {
                    while (p.back().e >= L.nodes[ln].e && L.nodes[ln].out.size() == 1 &&
                           L.node_len(L.nodes[ln].out[0]) == 0) {
                        uint32_t o = L.nodes[ln].out[0];
                        p.push_back(Seg{o, L.nodes[o].s, L.nodes[o].e});
                        ln = o;
                    }
                }
                uint32_t id = find_or_add(p, hs[j], st[j], created);
                if (!created) continue;
                if (!have_edge) link(0, id);
                have_edge = true;
                if (queued.size() < paths.size()) queued.resize(paths.size(), 0);
                queued[id] = 1;
                leaves.push_back(id);
            }
        }
    }

    while (!leaves.empty()) {
        const uint32_t cur = leaves.front();
        leaves.pop_front();
        const uint64_t cur_hash = hashes[cur];
        std::deque<std::vector<KPath>> chains;
        {
            auto first = S.shift(paths[cur]);
            if (first.empty()) end_leaves.push_back(cur);
            for (auto& p : first) chains.push_back({std::move(p)});
        }
        while (!chains.empty()) {
            std::vector<KPath> v = std::move(chains.front());
            chains.pop_front();
            uint64_t h;
            uint8_t st;
            S.hash_of(v.back(), h, st);
            if (h <= cur_hash) {
                uint32_t id = find_or_add(v.back(), h, st, created);
                link(cur, id);
                enqueue(id, v.back());
            } else if (v.size() == w) {
                std::vector<uint64_t> hs(w);
                std::vector<uint8_t> ss(w);
                uint64_t best = UINT64_MAX;
                for (uint32_t j = 0; j < w; ++j) {
                    S.hash_of(v[j], hs[j], ss[j]);
                    best = std::min(best, hs[j]);
                }
                bool have_edge = false;
                for (uint32_t j = 0; j < w; ++j) {
                    if (hs[j] != best) continue;
                    uint32_t id = find_or_add(v[j], hs[j], ss[j], created);
                    if (!have_edge) link(cur, id);
                    have_edge = true;
                    enqueue(id, v.back());
                }
            } else if (v.back().back().e == S.END) {
                end_leaves.push_back(cur);
            } else {
                for (auto& nx : S.shift(v.back())) {
                    chains.push_back(v);
                    chains.back().push_back(std::move(nx));
                }
            }
        }
    }
    uint32_t term = find_or_add(KPath{Seg{(uint32_t)L.nodes.size() - 1, S.END, S.END}}, UINT64_MAX, 1, created);
    if (end_leaves.empty()) link(0, term);
    for (uint32_t e : end_leaves) link(e, term);

    // shortcut edges a->c with a->b->c and b on the stretch a..c, judged on the original edge set
    std::vector<std::pair<uint32_t, uint32_t>> kill;
    for (uint32_t a = 0; a < paths.size(); ++a)
        for (uint32_t c : outs[a]) {
            if (!kpath_less(paths[a], paths[c])) continue;
            KPath u;
            bool have_u = false;
            for (uint32_t b : outs[a]) {
                if (b == c) continue;
                if (std::find(outs[b].begin(), outs[b].end(), c) == outs[b].end()) continue;
                if (!have_u) {
                    u = join(paths[a], paths[c]);
                    have_u = true;
                }
                if (u.empty()) break;
                if (lies_on(paths[b], u)) {
                    kill.push_back({a, c});
                    break;
                }
            }
        }
    for (auto& e : kill) {
        auto& o = outs[e.first];
        o.erase(std::find(o.begin(), o.end(), e.second));
    }

    // rank order = path order (pandora sorted_nodes): the DP order and the hit tie-break order
    const uint32_t N = (uint32_t)paths.size();
    std::vector<uint32_t> order(N), rank(N);
    for (uint32_t i = 0; i < N; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return kpath_less(paths[a], paths[b]); });
    for (uint32_t r = 0; r < N; ++r) rank[order[r]] = r;
    L.kpath.resize(N);
    L.khash.resize(N);
    L.kstrand.resize(N);
    L.kout.assign(N, {});
    for (uint32_t r = 0; r < N; ++r) {
        uint32_t id = order[r];
        L.kpath[r] = paths[id];
        L.khash[r] = hashes[id];
        L.kstrand[r] = strands[id];
        for (uint32_t o : outs[id]) L.kout[r].push_back(rank[o]);
        std::sort(L.kout[r].begin(), L.kout[r].end());
        if (r != 0 && r != N - 1 && !(id == 0 || id == term)) records.push_back(Record{hashes[id], prg_id, r, strands[id]});
    }
    // fewest edges from the null start to the null end
    std::vector<uint32_t> d(N, UINT32_MAX);
    d[N - 1] = 0;
    for (uint32_t r = N - 1; r-- > 0;)
        for (uint32_t o : L.kout[r])
            if (d[o] != UINT32_MAX && d[o] + 1 < d[r]) d[r] = d[o] + 1;
    L.min_path_len = d[0] == UINT32_MAX ? 0 : d[0];
}

HostIndex build_host_index(const std::string& prg_text, uint32_t w, uint32_t k) {
    if (w < 1 || k < 1 || k > 32) throw std::runtime_error("unsupported w/k");
    HostIndex H;
    H.w = w;
    H.k = k;
    parse_prg_text(prg_text, H.loci);
    if (H.loci.empty()) throw std::runtime_error("no PRG records found");
    H.knode_base.push_back(0);
    for (uint32_t i = 0; i < H.loci.size(); ++i) {
        sketch_locus(H.loci[i], i, w, k, H.records);
        H.knode_base.push_back(H.knode_base.back() + (uint32_t)H.loci[i].kpath.size());
    }
    std::sort(H.records.begin(), H.records.end(), [](const Record& a, const Record& b) {
        if (a.hash != b.hash) return a.hash < b.hash;
        if (a.prg != b.prg) return a.prg < b.prg;
        return a.knode < b.knode;
    });
    return H;
}

}  // namespace drprg
