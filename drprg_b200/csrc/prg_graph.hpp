// Host side of the device index: PRG text -> local graph -> (w,k) graph minimizers -> flat k-mer
// graphs + minimizer records, laid out for upload (CSR, rank-ordered).  This is stage a1 of
// SURVEY.md §8a (pandora LocalPRG ctor + LocalPRG::minimizer_sketch + Index::add_record — what
// `pandora index`, /root/reference/src/lib.rs:479-510, precomputes); it runs once per PRG in
// milliseconds on one host core and is not part of the per-read hot path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace drprg {

// a stretch of one local node, PRG-string coordinates [s,e); s==e for an empty node
struct Seg {
    uint32_t node, s, e;
    bool operator==(const Seg& o) const { return s == o.s && e == o.e; }
};
using KPath = std::vector<Seg>;

struct LNode {
    uint32_t s, e;  // PRG-string interval of the node's sequence
    std::vector<uint32_t> out;
};

struct Locus {
    std::string name, text;
    std::vector<LNode> nodes;
    // k-mer graph in rank (path-sorted) order; node 0 = null start, last = null end
    std::vector<KPath> kpath;
    std::vector<uint64_t> khash;
    std::vector<uint8_t> kstrand;
    std::vector<std::vector<uint32_t>> kout;  // ascending ranks
    uint32_t min_path_len = 0;

    std::string node_seq(uint32_t n) const { return text.substr(nodes[n].s, nodes[n].e - nodes[n].s); }
    uint32_t node_len(uint32_t n) const { return nodes[n].e - nodes[n].s; }
    uint32_t end_coord() const { return nodes.back().e; }
};

struct Record {
    uint64_t hash;
    uint32_t prg, knode;  // knode = rank within the locus
    uint8_t strand;
};

struct HostIndex {
    uint32_t w = 0, k = 0;
    std::vector<Locus> loci;
    std::vector<uint32_t> knode_base;  // n_loci + 1
    std::vector<Record> records;       // sorted (hash, prg, knode)
    uint32_t total_knodes() const { return knode_base.back(); }
};

uint64_t hash64_host(uint64_t key, uint64_t mask);
uint64_t hash64_inverse_host(uint64_t key, uint64_t mask);
void parse_prg_text(const std::string& text, std::vector<Locus>& loci);
void sketch_locus(Locus& L, uint32_t prg_id, uint32_t w, uint32_t k, std::vector<Record>& records);
HostIndex build_host_index(const std::string& prg_text, uint32_t w, uint32_t k);
std::string read_text_file(const std::string& path);

}  // namespace drprg
