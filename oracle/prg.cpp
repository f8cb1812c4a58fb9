// ORACLE (test infrastructure, see oracle.hpp).  PRG text -> LocalGraph -> (w,k) graph minimizers
// -> KmerGraph + minimizer index; read sketching.  Restates pandora src/prg/path.cpp, inthash.cpp,
// seq.cpp, localPRG.cpp (build_graph, split_by_site, shift, minimizer_sketch), localgraph.cpp
// (walk), kmergraph.cpp.  Reference call sites: /root/reference/src/lib.rs:479-510 (pandora index),
// :580-642 (pandora map).  PRG grammar pinned by /root/reference/tests/cases/expected/dr.prg.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cassert>
#include <deque>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// ------------------------------------------------------------------------------------- Path ---
uint32_t path_length(const Path& p) {
    uint32_t n = 0;
    for (auto& i : p) n += i.length;
    return n;
}

// pandora prg::Path::operator<
bool path_less(const Path& a, const Path& b) {
    size_t n = std::min(a.size(), b.size());
    for (size_t i = 0; i < n; ++i) {
        if (!(a[i] == b[i])) return a[i] < b[i];
    }
    return a.size() < b.size();
}

// pandora prg::Path::subpath(start, len): position `start` along the path, `len` bases.
Path path_subpath(const Path& p, uint32_t start, uint32_t len) {
    Path out;
    uint32_t covered = 0, added = 0;
    for (const auto& iv : p) {
        if (out.empty() && ((covered <= start && covered + iv.length > start) ||
                            (covered == start && iv.length == 0))) {
            uint32_t s = iv.start + start - covered;
            uint32_t e = std::min(iv.end(), s + len - added);
            out.push_back(Interval(s, e));
            added += std::min(len - added, iv.length - (start - covered));
        } else if (!out.empty() && covered >= start && covered <= start + len) {
            uint32_t e = std::min(iv.end(), iv.start + len - added);
            out.push_back(Interval(iv.start, e));
            added += std::min(len - added, iv.length);
        }
        covered += iv.length;
        if (added >= len && !out.empty()) break;
    }
    return out;
}

// pandora prg::Path::is_branching
bool path_is_branching(const Path& x, const Path& y) {
    if (x.empty() || y.empty()) return false;
    if (path_end(x) < path_start(y) || path_end(y) < path_start(x)) return false;
    bool overlap = false;
    size_t j = 0;
    for (size_t i = 0; i < x.size(); ++i) {
        if (overlap) {
            if (x[i].start != y[j].start) return true;
            ++j;
            if (j == y.size()) return false;
        } else {
            for (j = 0; j < y.size(); ++j) {
                if ((x[i].end() > y[j].start && x[i].start < y[j].end()) || x[i] == y[j]) {
                    overlap = true;
                    if (i != 0 && j != 0 && x[i - 1].end() != y[j - 1].end()) return true;
                    ++j;
                    if (j == y.size()) return false;
                    break;
                }
            }
        }
    }
    return false;
}

// pandora get_union(x, y) with x < y
Path path_union(const Path& x, const Path& y) {
    Path p;
    if (x.empty()) return y;
    if (y.empty()) return p;
    if (path_end(x) < path_start(y) || path_is_branching(x, y)) return p;
    size_t xi = 0, yi = 0;
    while (xi < x.size() && yi < y.size() && x[xi].end() < y[yi].start) {
        p.push_back(x[xi]);
        ++xi;
    }
    if (xi < x.size() && yi < y.size() && x[xi].start <= y[yi].end()) {
        p.push_back(Interval(x[xi].start, std::max(y[yi].end(), x[xi].end())));
        while (yi + 1 < y.size()) {
            ++yi;
            p.push_back(y[yi]);
        }
    }
    return p;
}

// pandora prg::Path::is_subpath(big_path)
bool path_is_subpath(const Path& small, const Path& big) {
    if (small.empty() || big.empty()) return false;
    uint32_t ls = path_length(small), lb = path_length(big);
    if (lb < ls || path_start(big) > path_start(small) || path_end(big) < path_end(small) ||
        path_is_branching(small, big))
        return false;
    uint32_t offset = 0;
    for (const auto& iv : big) {
        if (iv.end() >= path_start(small)) {
            if (path_start(small) < iv.start) return false;
            offset += path_start(small) - iv.start;
            if (offset + ls > lb) return false;
            Path sp = path_subpath(big, offset, ls);
            return sp == small;
        }
        offset += iv.length;
    }
    return false;
}

// ---------------------------------------------------------------------------------- hashing ---
// minimap2-style invertible integer hash, pandora src/inthash.cpp
uint64_t hash64(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

int nt4(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

std::pair<uint64_t, uint64_t> kmerhash(const std::string& s, uint32_t k) {
    uint64_t shift1 = 2 * (k - 1), mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1), kmer[2] = {0, 0};
    for (char ch : s) {
        int c = nt4((uint8_t)ch);
        if (c < 4) {
            kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
            kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
        }
    }
    return {hash64(kmer[0], mask), hash64(kmer[1], mask)};
}

// pandora Seq::minimizer_sketch: union over all w-windows of the k-mers attaining the window
// minimum of the canonical hash (ties all kept); empty if too short or any non-ACGT base.
std::vector<Minimizer> sketch_read(const char* seq, size_t len, uint32_t w, uint32_t k) {
    std::vector<Minimizer> out;
    if (len + 1 < (size_t)w + k) return out;
    uint64_t shift1 = 2 * (k - 1), mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1), kmer[2] = {0, 0};
    std::set<Minimizer> sk;
    std::vector<Minimizer> window;  // last <= w k-mers
    window.reserve(w + 1);
    uint64_t smallest = UINT64_MAX;
    uint32_t buff = 0;
    for (size_t i = 0; i < len; ++i) {
        int c = nt4((uint8_t)seq[i]);
        if (c >= 4) return {};  // "bad letter": whole read contributes nothing
        kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
        kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
        ++buff;
        if (buff < k) continue;
        uint64_t hf = hash64(kmer[0], mask), hr = hash64(kmer[1], mask);
        Minimizer m{std::min(hf, hr), (uint32_t)(i + 1 - k), hf <= hr};
        if (window.size() < w) {
            window.push_back(m);
            if (window.size() == w) {  // first full window
                smallest = UINT64_MAX;
                for (auto& x : window) smallest = std::min(smallest, x.hash);
                for (auto& x : window)
                    if (x.hash == smallest) sk.insert(x);
            }
            continue;
        }
        Minimizer gone = window.front();
        window.erase(window.begin());
        window.push_back(m);
        if (gone.hash == smallest) {  // the minimum may have left: re-minimise the window
            smallest = UINT64_MAX;
            for (auto& x : window) smallest = std::min(smallest, x.hash);
            for (auto& x : window)
                if (x.hash == smallest) sk.insert(x);
        } else if (m.hash <= smallest) {
            smallest = m.hash;
            sk.insert(m);
        }
    }
    out.assign(sk.begin(), sk.end());
    return out;
}

// -------------------------------------------------------------------------------- KmerGraph ---
uint32_t KmerGraph::add_node(const Path& p) {
    auto it = by_path.find(p);
    if (it != by_path.end()) return it->second;
    KmerNode n;
    n.id = (uint32_t)nodes.size();
    n.path = p;
    nodes.push_back(n);
    by_path[p] = n.id;
    return n.id;
}

void KmerGraph::add_edge(uint32_t from, uint32_t to) {
    if (from == to) return;
    auto& o = nodes[from].out;
    if (std::find(o.begin(), o.end(), to) == o.end()) {
        o.push_back(to);
        nodes[to].in.push_back(from);
    }
}

void KmerGraph::finalize() {
    sorted.resize(nodes.size());
    for (uint32_t i = 0; i < nodes.size(); ++i) sorted[i] = i;
    std::sort(sorted.begin(), sorted.end(),
              [&](uint32_t a, uint32_t b) { return path_less(nodes[a].path, nodes[b].path); });
    rank.assign(nodes.size(), 0);
    for (uint32_t i = 0; i < sorted.size(); ++i) rank[sorted[i]] = i;
}

// pandora KmerGraph::remove_shortcut_edges: drop a->c when a->b->c exists and b lies on the
// PRG path spanned by a and c (b.path is a subpath of union(a.path, c.path)).  All removals are
// decided against the original edge set (order-free), then applied together.
void KmerGraph::remove_shortcut_edges() {
    std::vector<std::pair<uint32_t, uint32_t>> kill;
    for (auto& n : nodes) {
        for (uint32_t c : n.out) {
            bool shortcut = false;
            for (uint32_t b : n.out) {
                if (b == c) continue;
                if (std::find(nodes[b].out.begin(), nodes[b].out.end(), c) == nodes[b].out.end()) continue;
                if (!path_less(n.path, nodes[c].path)) continue;
                Path u = path_union(n.path, nodes[c].path);
                if (u.empty() || !path_is_subpath(nodes[b].path, u)) continue;
                shortcut = true;
                break;
            }
            if (shortcut) kill.push_back({n.id, c});
        }
    }
    for (auto& e : kill) {
        auto& o = nodes[e.first].out;
        o.erase(std::find(o.begin(), o.end(), e.second));
        auto& in = nodes[e.second].in;
        in.erase(std::find(in.begin(), in.end(), e.first));
    }
}

// fewest edges on any null-start -> null-end path (pandora KmerGraph::min_path_length; L-confidence)
uint32_t KmerGraph::min_path_length() const {
    if (sorted.size() < 2) return 0;
    std::vector<uint32_t> len(nodes.size(), UINT32_MAX);
    len[sorted.back()] = 0;
    for (size_t j = sorted.size() - 1; j-- > 0;) {
        const KmerNode& n = nodes[sorted[j]];
        for (uint32_t o : n.out)
            if (len[o] != UINT32_MAX && len[o] + 1 < len[n.id]) len[n.id] = len[o] + 1;
    }
    return len[sorted[0]] == UINT32_MAX ? 0 : len[sorted[0]];
}

// --------------------------------------------------------------------------------- LocalPRG ---
namespace {
struct Tok {
    uint32_t start, len;
    bool marker;
    uint32_t value;
};
}  // namespace

// pandora LocalPRG::build_graph / split_by_site.  Every DNA token (possibly empty) between
// markers becomes one LocalNode, ids in order of appearance; odd marker n opens and closes
// site n, even marker n+1 separates its alleles.
void LocalPRG::build_graph() {
    std::vector<Tok> toks;
    {
        uint32_t i = 0, n = (uint32_t)seq.size();
        while (true) {
            uint32_t j = i;
            while (j < n && seq[j] != ' ') ++j;
            Tok t{i, j - i, false, 0};
            if (t.len > 0 && isdigit((unsigned char)seq[i])) {
                t.marker = true;
                t.value = (uint32_t)std::stoul(seq.substr(i, j - i));
            }
            toks.push_back(t);
            if (j >= n) break;
            i = j + 1;
        }
    }
    nodes.clear();
    // recursive descent over tokens [lo, hi): DNA (open alleles close DNA)*
    struct Rec {
        LocalPRG* self;
        std::vector<Tok>* toks;
        uint32_t new_node(const Tok& t, const std::vector<uint32_t>& from) {
            LocalNode nd;
            nd.id = (uint32_t)self->nodes.size();
            nd.seq = self->seq.substr(t.start, t.len);
            nd.pos = Interval(t.start, t.start + t.len);
            self->nodes.push_back(nd);
            for (uint32_t f : from) {
                self->nodes[f].out.push_back(nd.id);
                self->nodes[nd.id].in.push_back(f);
            }
            return nd.id;
        }
        std::vector<uint32_t> build(size_t lo, size_t hi, const std::vector<uint32_t>& from) {
            auto& T = *toks;
            if (lo >= hi || T[lo].marker) throw std::runtime_error("PRG parse: expected sequence token in " + self->name);
            uint32_t pre = new_node(T[lo], from);
            if (hi - lo == 1) return {pre};
            if (!T[lo + 1].marker || (T[lo + 1].value % 2) == 0)
                throw std::runtime_error("PRG parse: expected site-open marker in " + self->name);
            uint32_t site = T[lo + 1].value;
            size_t close = 0;
            for (size_t j = lo + 2; j < hi; ++j)
                if (T[j].marker && T[j].value == site) {
                    close = j;
                    break;
                }
            if (!close) throw std::runtime_error("PRG parse: unclosed site in " + self->name);
            std::vector<uint32_t> ends;
            size_t a = lo + 2;
            for (size_t j = lo + 2; j <= close; ++j) {
                if (j == close || (T[j].marker && T[j].value == site + 1)) {
                    auto e = build(a, j, {pre});
                    ends.insert(ends.end(), e.begin(), e.end());
                    a = j + 1;
                }
            }
            return build(close + 1, hi, ends);
        }
    } rec{this, &toks};
    rec.build(0, toks.size(), {});
    start_to_node.clear();
    for (auto& n : nodes) start_to_node[n.pos.start] = n.id;
}

// pandora LocalGraph::walk: all paths of exactly len bases from position pos in node node_id
std::vector<Path> LocalPRG::walk(uint32_t node_id, uint32_t pos, uint32_t len) const {
    std::vector<Path> ret;
    if (len == 0) return ret;
    const LocalNode& nd = nodes[node_id];
    if (pos + len <= nd.pos.end()) {
        ret.push_back({Interval(pos, pos + len)});
        return ret;
    }
    uint32_t len_added = std::min(nd.pos.end() - pos, len);
    if (len_added < len) {
        for (uint32_t o : nd.out) {
            auto sub = walk(o, nodes[o].pos.start, len - len_added);
            for (auto& s : sub) {
                Path p;
                p.push_back(Interval(pos, nd.pos.end()));
                p.insert(p.end(), s.begin(), s.end());
                if (path_length(p) == len) ret.push_back(std::move(p));
            }
        }
    }
    return ret;
}

// pandora LocalPRG::nodes_along_path
std::vector<uint32_t> LocalPRG::nodes_along_path(const Path& p) const {
    std::vector<uint32_t> v;
    for (size_t i = 0; i < p.size(); ++i) {
        const Interval& iv = p[i];
        if (iv.length == 0) {
            auto it = start_to_node.find(iv.start);
            if (it == start_to_node.end()) continue;
            const LocalNode& nd = nodes[it->second];
            if (nd.pos.length == 0)
                v.push_back(nd.id);  // an empty node
            else if (i + 1 == p.size() && nd.id != 0)
                v.push_back(nd.id);  // cursor at the start of a node, last interval only
        } else {
            auto it = start_to_node.upper_bound(iv.start);
            while (it != start_to_node.begin()) {
                --it;
                const LocalNode& nd = nodes[it->second];
                if (nd.pos.length > 0 && nd.pos.start <= iv.start && iv.end() <= nd.pos.end()) {
                    v.push_back(nd.id);
                    break;
                }
                if (nd.pos.length > 0) break;
            }
        }
    }
    return v;
}

std::string LocalPRG::string_along_path(const Path& p) const {
    std::string s;
    for (auto& iv : p) s += seq.substr(iv.start, iv.length);
    return s;
}

std::string LocalPRG::string_along_nodes(const std::vector<uint32_t>& np) const {
    std::string s;
    for (uint32_t n : np) s += nodes[n].seq;
    return s;
}

// pandora LocalPRG::shift: all paths of the same length shifted one base along the graph
std::vector<Path> LocalPRG::shift(const Path& p0) const {
    std::vector<Path> ret;
    uint32_t L = path_length(p0);
    if (L < 1) return ret;
    Path q = path_subpath(p0, 1, L - 1);
    std::deque<Path> short_paths;
    short_paths.push_back(q);
    std::vector<Path> k_paths;
    const uint32_t END = last_end();
    while (!short_paths.empty()) {
        Path p = short_paths.front();
        short_paths.pop_front();
        if (p.empty()) continue;
        auto n = nodes_along_path(p);
        if (n.empty()) continue;
        const LocalNode& last = nodes[n.back()];
        if (path_end(p) < last.pos.end()) {
            p.back().length += 1;
            k_paths.push_back(p);
        } else if (path_end(p) != END) {
            for (uint32_t o : last.out) {
                Path e = p;
                e.push_back(Interval(nodes[o].pos.start, nodes[o].pos.start));
                short_paths.push_back(e);
            }
        }
    }
    // by adding null nodes do we reach the end of the prg?
    for (auto& kp : k_paths) {
        std::deque<Path> sp;
        sp.push_back(kp);
        bool non_terminus = false;
        while (!sp.empty()) {
            Path p = sp.front();
            sp.pop_front();
            auto n = nodes_along_path(p);
            const LocalNode& last = nodes[n.back()];
            if (last.pos.end() == END) {
                ret.push_back(p);
            } else if (last.pos.end() == path_end(p)) {
                for (uint32_t o : last.out) {
                    if (nodes[o].pos.length == 0) {
                        Path e = p;
                        e.push_back(nodes[o].pos);
                        sp.push_back(e);
                    } else {
                        non_terminus = true;
                    }
                }
            } else {
                non_terminus = true;
            }
        }
        if (non_terminus) ret.push_back(kp);
    }
    return ret;
}

// pandora LocalPRG::minimizer_sketch
void LocalPRG::minimizer_sketch(std::unordered_map<uint64_t, std::vector<MiniRecord>>& index, uint32_t w, uint32_t k) {
    kg = KmerGraph();
    const uint32_t END = last_end();
    auto add_record = [&](uint64_t h, uint32_t knode, bool strand) {
        auto& v = index[h];
        for (auto& r : v)
            if (r.prg_id == id && r.knode_id == knode && r.strand == strand) return;
        v.push_back(MiniRecord{id, knode, strand});
    };
    auto hash_of = [&](const Path& p, uint64_t& hmin, bool& strand) {
        auto kh = kmerhash(string_along_path(p), k);
        hmin = std::min(kh.first, kh.second);
        strand = kh.first <= kh.second;
    };
    auto new_knode = [&](const Path& p, uint64_t h, bool strand, bool& created) -> uint32_t {
        size_t before = kg.nodes.size();
        uint32_t idn = kg.add_node(p);
        created = kg.nodes.size() > before;
        if (created) {
            kg.nodes[idn].khash = h;
            kg.nodes[idn].strand = strand;
            std::string s = string_along_path(p);
            kg.nodes[idn].num_AT = (uint32_t)(std::count(s.begin(), s.end(), 'A') + std::count(s.begin(), s.end(), 'T'));
            add_record(h, idn, strand);
        }
        return idn;
    };

    kg.add_node({Interval(0, 0)});  // null start
    std::deque<uint32_t> current_leaves;
    std::vector<uint32_t> end_leaves;
    // pandora re-queues a node every time it is found again; re-processing is idempotent (the
    // LocalGraph and all hashes are static), so each node is expanded once here.
    std::vector<char> queued;
    auto push_leaf = [&](uint32_t kn, const Path& p_last) {
        if (path_end(p_last) == END) {
            end_leaves.push_back(kn);
            return;
        }
        if (queued.size() <= kn) queued.resize(kg.nodes.size() + 16, 0);
        if (!queued[kn]) {
            queued[kn] = 1;
            current_leaves.push_back(kn);
        }
    };

    bool trivial = (nodes.size() == 1 && nodes[0].pos.length < k);
    std::vector<Path> walk_paths;
    if (!trivial) walk_paths = walk(0, 0, w + k - 1);
    // first (w,k) minimizers of every walk from the start
    for (auto& wp : walk_paths) {
        uint64_t smallest = UINT64_MAX;
        std::vector<Path> kp(w);
        std::vector<uint64_t> hs(w);
        std::vector<bool> st(w);
        for (uint32_t j = 0; j < w; ++j) {
            kp[j] = path_subpath(wp, j, k);
            bool s;
            hash_of(kp[j], hs[j], s);
            st[j] = s;
            smallest = std::min(smallest, hs[j]);
        }
        bool mini_found = false;
        for (uint32_t j = 0; j < w; ++j) {
            Path kmer_path = kp[j];
            auto n = nodes_along_path(kmer_path);
            if (!n.empty() && walk(n.back(), nodes[n.back()].pos.end(), w + k - 1).empty()) {
                while (path_end(kmer_path) >= nodes[n.back()].pos.end() && nodes[n.back()].out.size() == 1 &&
                       nodes[nodes[n.back()].out[0]].pos.length == 0) {
                    kmer_path.push_back(nodes[nodes[n.back()].out[0]].pos);
                    n.push_back(nodes[n.back()].out[0]);
                }
            }
            if (hs[j] == smallest) {
                bool created;
                uint32_t kn = new_knode(kmer_path, hs[j], st[j], created);
                if (created) {
                    if (!mini_found) kg.add_edge(0, kn);
                    mini_found = true;
                    if (queued.size() <= kn) queued.resize(kg.nodes.size() + 16, 0);
                    queued[kn] = 1;
                    current_leaves.push_back(kn);
                }
            }
        }
    }

    while (!current_leaves.empty()) {
        uint32_t kn = current_leaves.front();
        current_leaves.pop_front();
        const uint64_t kn_hash = kg.nodes[kn].khash;
        std::deque<std::vector<Path>> shifts;
        {
            auto sp = shift(kg.nodes[kn].path);
            if (sp.empty()) end_leaves.push_back(kn);
            for (auto& s : sp) shifts.push_back({s});
        }
        while (!shifts.empty()) {
            std::vector<Path> v = std::move(shifts.front());
            shifts.pop_front();
            uint64_t h;
            bool s;
            hash_of(v.back(), h, s);
            if (h <= kn_hash) {  // next minimizer
                bool created;
                uint32_t nk = new_knode(v.back(), h, s, created);
                kg.add_edge(kn, nk);
                push_leaf(nk, v.back());
            } else if (v.size() == w) {  // old minimizer left the window: minimise the w new k-mers
                uint64_t smallest = UINT64_MAX;
                std::vector<uint64_t> hs(w);
                std::vector<bool> st(w);
                for (uint32_t j = 0; j < w; ++j) {
                    bool sj;
                    hash_of(v[j], hs[j], sj);
                    st[j] = sj;
                    smallest = std::min(smallest, hs[j]);
                }
                bool mini_found = false;
                for (uint32_t j = 0; j < w; ++j) {
                    if (hs[j] != smallest) continue;
                    bool created;
                    uint32_t nk = new_knode(v[j], hs[j], st[j], created);
                    if (!mini_found) kg.add_edge(kn, nk);
                    mini_found = true;
                    push_leaf(nk, v.back());
                }
            } else if (path_end(v.back()) == END) {
                end_leaves.push_back(kn);
            } else {
                auto sp = shift(v.back());
                for (auto& sft : sp) {
                    shifts.push_back(v);
                    shifts.back().push_back(sft);
                }
            }
        }
    }

    uint32_t term = kg.add_node({Interval(END, END)});  // null end
    if (end_leaves.empty()) kg.add_edge(0, term);
    for (uint32_t e : end_leaves) kg.add_edge(e, term);
    kg.remove_shortcut_edges();
    kg.finalize();
}

std::vector<uint32_t> LocalPRG::top_path() const {
    std::vector<uint32_t> p{0};
    while (!nodes[p.back()].out.empty()) p.push_back(nodes[p.back()].out[0]);
    return p;
}

// pandora LocalPRG::get_valid_vcf_reference: node path from node 0 to the sink spelling s
std::vector<uint32_t> LocalPRG::path_spelling(const std::string& s) const {
    struct St {
        uint32_t node;
        size_t off;
    };
    std::vector<uint32_t> cur;
    std::vector<uint32_t> result;
    // iterative DFS with explicit path
    std::vector<std::pair<uint32_t, size_t>> stack;  // (node, child cursor)
    auto matches = [&](uint32_t n, size_t off) {
        const std::string& q = nodes[n].seq;
        if (off + q.size() > s.size()) return false;
        for (size_t i = 0; i < q.size(); ++i)
            if (toupper(q[i]) != toupper(s[off + i])) return false;
        return true;
    };
    if (!matches(0, 0)) return result;
    std::vector<size_t> offs;
    stack.push_back({0, 0});
    offs.push_back(nodes[0].seq.size());
    while (!stack.empty()) {
        auto& top = stack.back();
        const LocalNode& nd = nodes[top.first];
        size_t off = offs.back();
        if (nd.out.empty()) {
            if (off == s.size()) {
                for (auto& e : stack) result.push_back(e.first);
                return result;
            }
            stack.pop_back();
            offs.pop_back();
            continue;
        }
        if (top.second >= nd.out.size()) {
            stack.pop_back();
            offs.pop_back();
            continue;
        }
        uint32_t c = nd.out[top.second++];
        if (matches(c, off)) {
            stack.push_back({c, 0});
            offs.push_back(off + nodes[c].seq.size());
        }
    }
    return result;
}

// ------------------------------------------------------------------------------------ Index ---
std::unique_ptr<Index> build_index_from_text(const std::string& text, uint32_t w, uint32_t k) {
    if (k < 1 || k > 32 || w < 1) throw std::runtime_error("bad w/k");
    auto idx = std::make_unique<Index>();
    idx->w = w;
    idx->k = k;
    std::istringstream in(text);
    std::string line, name, body;
    auto flush = [&]() {
        if (name.empty()) return;
        LocalPRG p;
        p.id = (uint32_t)idx->prgs.size();
        p.name = name;
        p.seq = body;
        idx->prgs.push_back(std::move(p));
        name.clear();
        body.clear();
    };
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '>') {
            flush();
            name = line.substr(1);
            size_t sp = name.find_first_of(" \t");
            if (sp != std::string::npos) name = name.substr(0, sp);
        } else {
            body += line;
        }
    }
    flush();
    for (auto& p : idx->prgs) {
        p.build_graph();
        p.minimizer_sketch(idx->minhash, w, k);
    }
    idx->knode_base.clear();
    uint32_t base = 0;
    for (auto& p : idx->prgs) {
        idx->knode_base.push_back(base);
        base += (uint32_t)p.kg.nodes.size();
    }
    idx->total_knodes = base;
    return idx;
}

std::unique_ptr<Index> build_index(const std::string& prg_path, uint32_t w, uint32_t k) {
    std::ifstream f(prg_path);
    if (!f) throw std::runtime_error("cannot open PRG file " + prg_path);
    std::stringstream ss;
    ss << f.rdbuf();
    return build_index_from_text(ss.str(), w, k);
}

}  // namespace orc
