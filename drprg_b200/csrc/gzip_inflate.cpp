// Parallel inflate of a SINGLE-STREAM gzip file (what `drprg predict -i reads.fastq.gz` usually gets:
// /root/reference/src/predict.rs:166-170, docs/src/guide/predict.md:36).  zlib inflates one stream on one core
// (~0.3 GB/s of text, 1 s per million 150 bp reads), 600x the GPU time of the sample; this file cuts the COMPRESSED
// stream into chunks that are decoded side by side:
//   1. block search   every chunk but the first starts at an unknown bit offset: candidate offsets are tried until a
//                     dynamic-Huffman block header parses (complete code-length, literal/length and distance codes, an
//                     end-of-block code) and the block decodes to the end of a plausible text block followed by another
//                     valid header;
//   2. decoding with an unknown window   a chunk does not know the 32 KB of text before it, so it decodes into 16-bit
//                     symbols: 0..255 = a byte, 0x8000 | i = "byte i of the unknown window"; back-references copy
//                     symbols, so the unknowns propagate exactly;
//   3. resolution     chunk 0 has no unknowns; the last 32 KB of chunk t-1 (resolved) turn the symbols of chunk t into
//                     bytes.  Only the 32 KB windows are chained sequentially, the bulk translation is parallel.
// (The scheme is the two-pass decompression of Kerbiriou & Chikhi, "Parallel decompression of gzip-compressed files and
// random access to DNA sequences", 2019 — restated here, no code taken.)
// Concatenated members (`cat a.gz b.gz`, lanes of one sample) and bgzip / BGZF files (one member per <= 64 KB) are handled
// by the same passes: a gzip member header close behind a nominal chunk boundary is the preferred chunk start (nothing
// before it can be referenced: no unknown window), and a chunk decodes through any member ends on its way, recording each
// member's trailer.
// The CRC-32 and length of EVERY member are verified (PCLMUL CRC-32 on the pieces between member ends, combined across
// chunk boundaries); any failure — a corrupt file, or a stream this decoder does not handle — makes the caller fall back
// to zlib's sequential gzread.
#include "gzip_inflate.hpp"

#include <immintrin.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "genotype_host.hpp"  // parallel_for

namespace drprg {
namespace {

constexpr uint32_t WIN = 32768;
constexpr uint16_t UNKNOWN = 0x8000;

// ---- bit reader over a byte range (LSB first, like deflate).  The caller guarantees 16 readable bytes behind p[n)
// (zero padding), so a refill is one unaligned 64-bit load.
struct Bits {
    const uint8_t* p;
    size_t n;        // bytes of real input
    size_t pos = 0;  // next byte to load
    uint64_t buf = 0;
    int cnt = 0;     // valid bits in buf
    bool over = false;
    Bits(const uint8_t* p_, size_t n_, size_t bitpos) : p(p_), n(n_) {
        pos = bitpos >> 3;
        refill();
        const int skip = (int)(bitpos & 7);
        buf >>= skip;
        cnt -= skip;
    }
    inline void refill() {  // afterwards cnt >= 56
        if (pos > n + 8) {  // far past the end: decoding garbage (a wrong block-start guess or a truncated file)
            over = true;
            cnt |= 56;
            return;
        }
        uint64_t w;
        memcpy(&w, p + pos, 8);
        buf |= w << cnt;
        pos += (size_t)((63 - cnt) >> 3);
        cnt |= 56;
    }
    inline uint32_t peek(int k) const { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    inline void drop(int k) {
        buf >>= k;
        cnt -= k;
    }
    inline uint32_t get(int k) {
        if (cnt < k) refill();
        const uint32_t v = peek(k);
        drop(k);
        return v;
    }
    size_t bit_position() const { return pos * 8 - (size_t)cnt; }  // next unread bit
    bool exhausted() const { return bit_position() > n * 8; }
};

// ---- canonical Huffman decoding tables: one level of ROOT bits, longer (rare) codes by a canonical bit-by-bit walk ----
struct Huff {
    static constexpr int ROOT = 11;
    // entry: bits 0..3 = code length (0 = needs the slow path), 4..7 = number of extra bits, 8 = literal, 9 = end of block,
    // 10 = invalid symbol, 16..31 = the literal / the length base / the distance base: one lookup gives everything the
    // decoder needs (no second table for base and extra bits on the match path)
    static constexpr uint32_t F_LIT = 1u << 8, F_EOB = 1u << 9, F_BAD = 1u << 10;
    uint32_t fast[1 << ROOT];
    uint16_t count[16], symbol[320];
    int max_len = 0;
    bool build(const uint8_t* lens, int n) {  // false: over-subscribed or (incomplete with more than one code)
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i) ++count[lens[i]];
        count[0] = 0;
        int left = 1, codes = 0;
        max_len = 0;
        for (int l = 1; l < 16; ++l) {
            left <<= 1;
            left -= count[l];
            if (left < 0) return false;
            if (count[l]) max_len = l;
            codes += count[l];
        }
        if (codes == 0) return false;
        if (left > 0 && !(codes == 1 && count[1] == 1)) return false;  // incomplete: only a lone 1-bit code is allowed
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; ++i)
            if (lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
        return true;
    }
    static inline uint32_t litlen_entry(uint32_t sym);
    static inline uint32_t dist_entry(uint32_t sym);
    void fill_fast(bool is_dist) {  // separate from build(): the block search builds thousands of tables it never decodes with
        memset(fast, 0, sizeof fast);
        int code = 0, idx = 0;
        for (int l = 1; l <= std::min(ROOT, 15); ++l) {
            for (int k = 0; k < count[l]; ++k, ++code, ++idx) {
                // deflate codes are MSB-first in a stream that is read LSB-first: reverse the code
                uint32_t rev = 0;
                for (int b = 0; b < l; ++b) rev |= ((code >> b) & 1u) << (l - 1 - b);
                const uint32_t entry = (is_dist ? dist_entry(symbol[idx]) : litlen_entry(symbol[idx])) | (uint32_t)l;
                for (uint32_t fill = rev; fill < (1u << ROOT); fill += 1u << l) fast[fill] = entry;
            }
            code <<= 1;
        }
    }
    // canonical decode bit by bit
    inline int slow(Bits& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= max_len; ++l) {
            code |= (int)b.get(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t Huff::litlen_entry(uint32_t sym) {
    if (sym < 256) return (sym << 16) | F_LIT;
    if (sym == 256) return F_EOB;
    if (sym > 285) return F_BAD;
    return ((uint32_t)LEN_BASE[sym - 257] << 16) | ((uint32_t)LEN_EXTRA[sym - 257] << 4);
}
inline uint32_t Huff::dist_entry(uint32_t sym) {
    if (sym > 29) return F_BAD;
    return ((uint32_t)DIST_BASE[sym] << 16) | ((uint32_t)DIST_EXTRA[sym] << 4);
}

struct BlockCodes {
    Huff lit, dist;
    bool have_dist = false;
};

// parse a dynamic block header at the reader's position (after BFINAL/BTYPE); false = not a valid header
bool read_dynamic_header(Bits& b, BlockCodes& C, bool fill = true) {
    const uint32_t hlit = b.get(5) + 257, hdist = b.get(5) + 1, hclen = b.get(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t cl[19] = {0};
    for (uint32_t i = 0; i < hclen; ++i) cl[CL_ORDER[i]] = (uint8_t)b.get(3);
    Huff clh;
    if (!clh.build(cl, 19)) return false;
    uint8_t lens[320];
    uint32_t i = 0;
    while (i < hlit + hdist) {
        if (b.cnt < 32) b.refill();
        const int sym = clh.slow(b);
        if (sym < 0 || b.over) return false;
        if (sym < 16) lens[i++] = (uint8_t)sym;
        else {
            uint32_t rep, val = 0;
            if (sym == 16) {
                if (i == 0) return false;
                val = lens[i - 1];
                rep = 3 + b.get(2);
            } else if (sym == 17) rep = 3 + b.get(3);
            else rep = 11 + b.get(7);
            if (i + rep > hlit + hdist) return false;
            while (rep--) lens[i++] = (uint8_t)val;
        }
    }
    if (lens[256] == 0) return false;  // no end-of-block code
    if (!C.lit.build(lens, (int)hlit)) return false;
    // a distance alphabet with no code at all is legal when the block has only literals
    bool any = false;
    for (uint32_t k = 0; k < hdist; ++k) any = any || lens[hlit + k];
    C.have_dist = any;
    if (any && !C.dist.build(lens + hlit, (int)hdist)) return false;
    if (fill) {
        C.lit.fill_fast(false);
        if (any) C.dist.fill_fast(true);
    }
    return true;
}

void fixed_codes(BlockCodes& C) {
    uint8_t l[288];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    C.lit.build(l, 288);
    C.lit.fill_fast(false);
    uint8_t d[30];
    for (int i = 0; i < 30; ++i) d[i] = 5;
    C.dist.build(d, 30);
    C.dist.fill_fast(true);
    C.have_dist = true;
}

// ---- a chunk's output: 16-bit symbols, the first WIN entries are the (unknown) window ----------------------------
// (a plain growable array: std::vector would zero-fill every resize and check capacity on every literal)
struct SymBuf {
    uint16_t* p = nullptr;
    size_t n = 0, cap = 0;  // n includes the window
    SymBuf() = default;
    SymBuf(const SymBuf&) = delete;
    SymBuf& operator=(const SymBuf&) = delete;
    SymBuf(SymBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    SymBuf& operator=(SymBuf&& o) noexcept {
        if (this != &o) {
            free(p);
            p = o.p; n = o.n; cap = o.cap;
            o.p = nullptr; o.n = o.cap = 0;
        }
        return *this;
    }
    ~SymBuf() { free(p); }
    size_t size() const { return n - WIN; }
    void reserve(size_t want) {
        if (want <= cap) return;
        size_t c = std::max(want, cap + cap / 2 + 4096);
        uint16_t* q = (uint16_t*)realloc(p, c * sizeof(uint16_t));
        if (!q) throw std::bad_alloc();
        p = q;
        cap = c;
    }
    inline void room(size_t extra) {
        if (n + extra > cap) reserve(n + extra);
    }
    void init_unknown() {
        reserve(WIN + 65536);
        for (uint32_t i = 0; i < WIN; ++i) p[i] = (uint16_t)(UNKNOWN | i);
        n = WIN;
    }
    void init_empty() {
        reserve(WIN + 65536);
        memset(p, 0, WIN * sizeof(uint16_t));
        n = WIN;
    }
    void release() {
        free(p);
        p = nullptr;
        n = cap = 0;
    }
};

enum class Stop { EndOfMember, Limit, Error };

// One Huffman-coded block, from after its header to its end-of-block code.  The input of this program is FASTQ text at
// low compression levels: ~90 % of the output comes from short matches (deflate_fast accepts any 3-base match), so the
// match path is the hot one: one table lookup for the length (base and extra-bit count packed in the entry), one for the
// distance, and a copy that always moves 16 symbols before it looks at the length.  Bit budget: the buffer holds >= 48
// bits at the top of the loop; two literals take <= 30, a length <= 20 and a distance <= 28 bits.
template <bool TEXT_ONLY>
Stop huffman_block(Bits& b, SymBuf& out, const BlockCodes& C) {
    constexpr uint32_t MASK = (1u << Huff::ROOT) - 1u;
    const uint32_t* lt = C.lit.fast;
    const uint32_t* dt = C.dist.fast;
    auto plausible = [](uint32_t sym) { return sym < 0x80 && (sym >= 0x20 || sym == '\n' || sym == '\r' || sym == '\t'); };
    for (;;) {
        out.room(520);
        if (b.cnt < 48) {
            b.refill();
            if (b.over) return Stop::Error;
        }
        uint32_t e = lt[b.buf & MASK];
        if (e & Huff::F_LIT) {
            b.drop(e & 15);
            if (TEXT_ONLY && !plausible(e >> 16)) return Stop::Error;
            out.p[out.n++] = (uint16_t)(e >> 16);
            e = lt[b.buf & MASK];
            if (e & Huff::F_LIT) {
                b.drop(e & 15);
                if (TEXT_ONLY && !plausible(e >> 16)) return Stop::Error;
                out.p[out.n++] = (uint16_t)(e >> 16);
                continue;
            }
            if (b.cnt < 48) {  // the low bits stay where they are: e is still the entry of the next code
                b.refill();
                if (b.over) return Stop::Error;
            }
        }
        if (!(e & 15)) {  // a code longer than ROOT bits
            const int sym = C.lit.slow(b);
            if (sym < 0) return Stop::Error;
            e = Huff::litlen_entry((uint32_t)sym);
            if (e & Huff::F_LIT) {
                if (TEXT_ONLY && !plausible(e >> 16)) return Stop::Error;
                out.p[out.n++] = (uint16_t)(e >> 16);
                continue;
            }
            if (b.cnt < 48) b.refill();
        } else {
            b.drop(e & 15);
        }
        if (e & (Huff::F_EOB | Huff::F_BAD)) return (e & Huff::F_EOB) ? Stop::EndOfMember : Stop::Error;  // EndOfMember = "block ended" here
        if (!C.have_dist) return Stop::Error;
        const uint32_t lx = (e >> 4) & 15u;
        const uint32_t len = (e >> 16) + b.peek((int)lx);
        b.drop((int)lx);
        uint32_t d = dt[b.buf & MASK];
        if (!(d & 15)) {
            const int ds = C.dist.slow(b);
            if (ds < 0) return Stop::Error;
            d = Huff::dist_entry((uint32_t)ds);
            if (b.cnt < 16) b.refill();
        } else {
            b.drop(d & 15);
        }
        if (d & Huff::F_BAD) return Stop::Error;
        const uint32_t dx = (d >> 4) & 15u;
        const uint32_t dist = (d >> 16) + b.peek((int)dx);
        b.drop((int)dx);
        if (dist > out.n) return Stop::Error;
        uint16_t* dst = out.p + out.n;
        const uint16_t* src = dst - dist;
        if (dist >= 8) {  // 16-byte steps (may run up to 15 symbols past the end: room() keeps the slack)
            memcpy(dst, src, 16);
            if (len > 8)
                for (uint32_t i = 8; i < len; i += 8) memcpy(dst + i, src + i, 16);
        } else if (dist == 1) {  // a run (quality strings): broadcast
            const uint64_t v = 0x0001000100010001ull * src[0];
            for (uint32_t i = 0; i < len; i += 4) memcpy(dst + i, &v, 8);
        } else {
            for (uint32_t i = 0; i < len; ++i) dst[i] = src[i];
        }
        out.n += len;
    }
}

// Decode blocks from the reader's position until the final block of the member ends or a block ends at/after bit
// `limit_bit` (blocks are never cut).  text_only: a literal outside the plausible text range is an error (used while
// validating a guessed block start).  Returns how it stopped; `end_bit` = position after the last decoded block.
Stop decode_blocks(Bits& b, SymBuf& out, size_t limit_bit, bool text_only, size_t max_blocks, size_t& end_bit, size_t* n_blocks = nullptr) {
    BlockCodes C;
    size_t blocks = 0;
    for (;;) {
        if (b.over) return Stop::Error;
        const uint32_t bfinal = b.get(1), btype = b.get(2);
        if (btype == 3) return Stop::Error;
        if (btype == 0) {
            b.drop(b.cnt & 7);  // to the byte boundary
            const uint32_t len = b.get(16), nlen = b.get(16);
            if ((len ^ nlen) != 0xffffu) return Stop::Error;
            out.room(len);
            for (uint32_t i = 0; i < len; ++i) {
                if (b.over) return Stop::Error;
                out.p[out.n++] = (uint16_t)b.get(8);
            }
        } else {
            if (btype == 1) fixed_codes(C);
            else if (!read_dynamic_header(b, C)) return Stop::Error;
            const Stop r = text_only ? huffman_block<true>(b, out, C) : huffman_block<false>(b, out, C);
            if (r == Stop::Error) return r;
        }
        ++blocks;
        if (n_blocks) *n_blocks = blocks;
        end_bit = b.bit_position();
        if (b.exhausted()) return Stop::Error;
        if (bfinal) return Stop::EndOfMember;
        if (end_bit >= limit_bit || blocks >= max_blocks) return Stop::Limit;
    }
}

// first bit offset >= from_bit (and < to_bit) at which a non-final dynamic block starts, decodes to plausible text and
// is followed by another parseable block header; SIZE_MAX if none
size_t find_block_start(const uint8_t* p, size_t n, size_t from_bit, size_t to_bit) {
    for (size_t bit = from_bit; bit < to_bit; ++bit) {
        // cheap reject first: BFINAL = 0, BTYPE = 2 and sane HLIT / HDIST without building anything
        const size_t byte = bit >> 3;
        if (byte + 4 >= n) return SIZE_MAX;
        uint32_t w;
        memcpy(&w, p + byte, 4);
        w >>= bit & 7;
        if ((w & 7u) != 4u) continue;               // bfinal 0, btype 10b (LSB first: 0, then 0,1)
        if (((w >> 3) & 31u) > 29u) continue;       // hlit
        if (((w >> 8) & 31u) > 29u) continue;       // hdist
        Bits b(p, n, bit + 3);
        BlockCodes C;
        if (!read_dynamic_header(b, C, false)) continue;
        // full check: the block (and the next one's header) must decode as text
        Bits b2(p, n, bit);
        SymBuf tmp;
        tmp.init_unknown();
        size_t end = 0;
        const Stop s = decode_blocks(b2, tmp, SIZE_MAX, true, 2, end);
        if (s == Stop::Error) continue;
        if (tmp.size() < 1024 && s != Stop::EndOfMember) continue;  // implausibly small blocks: keep looking
        return bit;
    }
    return SIZE_MAX;
}

// gzip member header at byte offset `at`; returns the offset of the deflate data or SIZE_MAX
size_t skip_gzip_header(const uint8_t* p, size_t n, size_t at) {
    if (at + 10 > n || p[at] != 0x1f || p[at + 1] != 0x8b || p[at + 2] != 8) return SIZE_MAX;
    const uint8_t flg = p[at + 3];
    size_t q = at + 10;
    if (flg & 4) {
        if (q + 2 > n) return SIZE_MAX;
        q += 2 + (p[q] | (p[q + 1] << 8));
    }
    if (flg & 8) {
        while (q < n && p[q]) ++q;
        ++q;
    }
    if (flg & 16) {
        while (q < n && p[q]) ++q;
        ++q;
    }
    if (flg & 2) q += 2;
    return q <= n ? q : SIZE_MAX;
}


// ---- CRC-32 by carry-less multiplication (4 x 128-bit folding, then Barrett reduction; the folding constants are the
// x^n mod P values of the reflected CRC-32 polynomial published in Intel's "Fast CRC Computation for Generic Polynomials
// Using PCLMULQDQ").  zlib's table-driven crc32 runs at 1-3 GB/s per core, which made the checksum a visible part of
// the inflate; the result is checked against zlib once per process and zlib is used if the CPU lacks PCLMUL/SSE4.1.
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul_raw(const uint8_t* buf, size_t len, uint32_t crc) {
    // len >= 64 and a multiple of 16; crc is the raw (already inverted) register
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
    x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
    x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = _mm_load_si128((const __m128i*)k1k2);
    buf += 64;
    len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
        y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
        y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }
    x0 = _mm_load_si128((const __m128i*)k3k4);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i*)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64((const __m128i*)k5k0);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128((const __m128i*)poly);
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}

uint32_t crc32_zlib(const uint8_t* d, size_t n) {
    uLong c = crc32(0L, Z_NULL, 0);
    while (n) {
        const size_t step = std::min<size_t>(n, 1u << 30);
        c = crc32(c, d, (uInt)step);
        d += step;
        n -= step;
    }
    return (uint32_t)c;
}

uint32_t crc32_fast_unchecked(const uint8_t* d, size_t n) {
    if (n < 64) return crc32_zlib(d, n);
    const size_t body = n & ~(size_t)15;
    const uint32_t raw = crc32_clmul_raw(d, body, 0xffffffffu);
    uLong c = (uLong)(raw ^ 0xffffffffu);
    if (n > body) c = crc32(c, d + body, (uInt)(n - body));
    return (uint32_t)c;
}

uint32_t crc32_fast(const uint8_t* d, size_t n) {
    static const bool usable = [] {
        if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
        uint8_t t[4099];
        uint32_t x = 12345;
        for (auto& b : t) {
            x = x * 1664525u + 1013904223u;
            b = (uint8_t)(x >> 24);
        }
        for (size_t len : {64ul, 65ul, 80ul, 127ul, 128ul, 1000ul, 4096ul, 4099ul})
            if (crc32_fast_unchecked(t, len) != crc32_zlib(t, len)) return false;
        return true;
    }();
    return usable ? crc32_fast_unchecked(d, n) : crc32_zlib(d, n);
}

// symbols -> bytes: 32 symbols at a time when none of them refers to the unknown window (the common case after the first
// few hundred KB of a chunk), through the lookup table otherwise
__attribute__((target("avx2"))) void translate_avx2(const uint16_t* s, uint8_t* d, size_t n, const uint8_t* L) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(s + i));
        const __m256i b = _mm256_loadu_si256((const __m256i*)(s + i + 16));
        if ((_mm256_movemask_epi8(_mm256_or_si256(a, b)) & 0xaaaaaaaa) == 0) {
            const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi16(a, b), 0xd8);
            _mm256_storeu_si256((__m256i*)(d + i), p);
        } else {
            for (size_t j = i; j < i + 32; ++j) d[j] = L[s[j]];
        }
    }
    for (; i < n; ++i) d[i] = L[s[i]];
}
void translate_scalar(const uint16_t* s, uint8_t* d, size_t n, const uint8_t* L) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        d[i] = L[s[i]]; d[i + 1] = L[s[i + 1]]; d[i + 2] = L[s[i + 2]]; d[i + 3] = L[s[i + 3]];
        d[i + 4] = L[s[i + 4]]; d[i + 5] = L[s[i + 5]]; d[i + 6] = L[s[i + 6]]; d[i + 7] = L[s[i + 7]];
    }
    for (; i < n; ++i) d[i] = L[s[i]];
}

struct MemberEnd {
    size_t out_pos;        // symbols of the chunk's output that belong to members ending at or before this one
    uint32_t crc, len;     // the member's trailer
};
struct Chunk {
    size_t start_bit = 0;     // first block of the chunk, or (member_start) the first byte of a gzip member header, in bits
    size_t end_bit = 0;       // after its last block / member
    bool unknown_window = true;
    bool member_start = false;         // the chunk begins with a member header: nothing before it can be referenced
    std::vector<MemberEnd> members;    // members that end inside this chunk
    SymBuf out;
    Stop stop = Stop::Error;
    bool ok = false;
    size_t out_off = 0;       // byte offset in the final text
    uint8_t last_window[WIN];  // resolved
    uint32_t last_n = 0;
};

// first gzip member header at a byte offset in [from_byte, to_byte) whose deflate data starts with blocks that decode as
// text (concatenated .gz files, bgzip / BGZF blocks); SIZE_MAX if none.  A member start is the best chunk start there is:
// back-references cannot cross it, so the chunk needs no unknown window.
size_t find_member_start(const uint8_t* p, size_t n, size_t from_byte, size_t to_byte) {
    size_t at = from_byte;
    while (at + 18 < n && at < to_byte) {
        const uint8_t* h = (const uint8_t*)memchr(p + at, 0x1f, std::min(to_byte, n - 18) - at);
        if (!h) return SIZE_MAX;
        at = (size_t)(h - p);
        if (p[at + 1] == 0x8b && p[at + 2] == 8 && (p[at + 3] & 0xe0) == 0) {
            const size_t q = skip_gzip_header(p, n, at);
            if (q != SIZE_MAX && q + 8 <= n) {
                Bits b(p, n, q * 8);
                SymBuf tmp;
                tmp.init_empty();
                size_t end = 0;
                if (decode_blocks(b, tmp, SIZE_MAX, true, 2, end) != Stop::Error) return at;
            }
        }
        ++at;
    }
    return SIZE_MAX;
}

// Decodes a chunk: from its start (a block start inside a member, or a member header) through any number of member
// boundaries up to `limit_bit` — the start of the next chunk (same two kinds) — or, for the last chunk, to the end of the
// file.  Every member that ends inside the chunk is recorded with its trailer.  false: corrupt data, a wrong start guess,
// or the chunk did not end exactly where the next one starts.
bool decode_chunk(const uint8_t* gz, size_t n, Chunk& c, size_t limit_bit, bool last) {
    size_t pos_bit = c.start_bit;
    bool at_header = c.member_start;
    for (;;) {
        if (at_header) {
            const size_t byte = pos_bit >> 3;
            if (!last && pos_bit == limit_bit) {
                c.end_bit = pos_bit;
                return true;
            }
            if (byte == n) {  // end of the file
                c.end_bit = pos_bit;
                return last;
            }
            if (!last && pos_bit > limit_bit) return false;
            const size_t q = skip_gzip_header(gz, n, byte);
            if (q == SIZE_MAX || q + 8 > n) return false;
            pos_bit = q * 8;
        }
        Bits b(gz, n, pos_bit);
        size_t end = pos_bit;
        const Stop r = decode_blocks(b, c.out, last ? SIZE_MAX : limit_bit, false, SIZE_MAX, end);
        if (r == Stop::Error) return false;
        if (r == Stop::Limit) {  // stopped after a block at or beyond the next chunk's first block
            c.end_bit = end;
            return !last && end == limit_bit;
        }
        // end of a member: trailer, then the next header (or the end of the file)
        const size_t trailer = (end + 7) / 8;
        if (trailer + 8 > n) return false;
        MemberEnd m;
        m.out_pos = c.out.size();
        memcpy(&m.crc, gz + trailer, 4);
        memcpy(&m.len, gz + trailer + 4, 4);
        c.members.push_back(m);
        pos_bit = (trailer + 8) * 8;
        at_header = true;
    }
}

}  // namespace

bool parallel_gunzip(const uint8_t* gz, size_t n, uint32_t threads, char** out, size_t* out_n) {
    *out = nullptr;
    *out_n = 0;
    static const bool timing = getenv("DRPRG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    const size_t data0 = skip_gzip_header(gz, n, 0);
    if (data0 == SIZE_MAX || n < data0 + 18) return false;
    static const size_t chunk_min = [] {  // compressed bytes a chunk must have (tests lower it to cut small files)
        const char* e = getenv("DRPRG_PARALLEL_GZIP_CHUNK");
        return e && atol(e) > 0 ? (size_t)atol(e) : (size_t)(1u << 20);
    }();
    // more chunks than threads (they decode at different speeds and are handed out dynamically), and never more than
    // 8 MB of compressed bytes per chunk: a large file becomes many chunks that are decoded GROUP BY GROUP, so the 16-bit
    // symbol buffers (2 bytes per byte of text) only ever exist for one group
    const size_t n_threads = std::max(1u, threads);
    static const size_t group_env = [] {  // tests set a small group so that small files take several
        const char* e = getenv("DRPRG_PARALLEL_GZIP_GROUP");
        return e && atol(e) > 0 ? (size_t)atol(e) : (size_t)0;
    }();
    const size_t group = group_env ? group_env : n_threads * 4;
    const size_t chunk_bytes = std::min<size_t>(std::max<size_t>(chunk_min, (n - data0) / (n_threads * 4)), std::max<size_t>(chunk_min, 8u << 20));
    size_t T = std::max<size_t>(1, std::min<size_t>((n - data0) / chunk_bytes, (size_t)1 << 16));
    // whole groups: a last group of a few chunks would leave most threads idle for a full chunk's decode time
    if (T > group && (n - data0) / ((T + group - 1) / group * group) >= chunk_min) T = (T + group - 1) / group * group;
    if (T < 2) return false;  // small files: zlib is as fast
    // ---- 1. chunk starts
    std::vector<Chunk> ch(T);
    ch[0].start_bit = data0 * 8;
    ch[0].unknown_window = false;
    std::vector<size_t> guess(T);
    for (size_t t = 1; t < T; ++t) guess[t] = (data0 + (n - data0) * t / T) * 8;
    parallel_for_io(T - 1, [&](size_t i) {
        const size_t t = i + 1;
        const size_t hi = t + 1 < T ? guess[t + 1] : (n - 8) * 8;
        // a member header close behind the nominal boundary (bgzip: every <= 64 KB; `cat a.gz b.gz`: wherever the files
        // meet) is taken as it is; otherwise the first dynamic block start, unless a member starts before it
        const size_t near = std::min(hi >> 3, (guess[t] >> 3) + (256u << 10));
        size_t m = find_member_start(gz, n, guess[t] >> 3, near);
        if (m == SIZE_MAX) {
            const size_t blk = find_block_start(gz, n, guess[t], hi);
            m = find_member_start(gz, n, near, blk == SIZE_MAX ? (hi >> 3) : (blk >> 3));
            if (m == SIZE_MAX) {
                ch[t].start_bit = blk;
                return;
            }
        }
        ch[t].start_bit = m * 8;
        ch[t].member_start = true;
        ch[t].unknown_window = false;
    }, n_threads);
    const double t1 = now();
    // chunks whose start was not found are merged into their predecessor
    std::vector<Chunk> live;
    live.reserve(T);
    for (size_t t = 0; t < T; ++t)
        if (t == 0 || ch[t].start_bit != SIZE_MAX) live.push_back(std::move(ch[t]));
    T = live.size();
    if (T < 2) return false;
    // ---- 2 + 3, group by group: decode every chunk up to the start of the next one; chain the 32 KB windows; translate
    char* text = nullptr;
    size_t total = 0;
    struct TextGuard {  // freed unless the function succeeds
        char*& p;
        bool keep = false;
        ~TextGuard() {
            if (!keep) {
                free(p);
                p = nullptr;
            }
        }
    } guard{text};
    std::vector<std::vector<uint32_t>> crcs(T);  // per chunk: CRC-32 of every piece between member ends
    std::vector<size_t> out_size(T, 0);
    std::vector<double> tt(T, 0), tc(T, 0), tf(T, 0);
    double ms_decode = 0, ms_chain = 0, ms_translate = 0;
    for (size_t g0 = 0; g0 < T; g0 += group) {
        const size_t g1 = std::min(T, g0 + group);
        const double ta = now();
        parallel_for_io(g1 - g0, [&](size_t i) {
            const size_t t = g0 + i;
            Chunk& c = live[t];
            if (c.unknown_window) c.out.init_unknown();
            else c.out.init_empty();
            c.out.reserve(WIN + (size_t)((n / T) * 4));
            c.ok = decode_chunk(gz, n, c, t + 1 < T ? live[t + 1].start_bit : SIZE_MAX, t + 1 == T);
        }, n_threads);
        const double tb = now();
        for (size_t t = g0; t < g1; ++t)
            if (!live[t].ok) return false;  // a guessed start was wrong, or corrupt data
        for (size_t t = g0; t < g1; ++t) {
            Chunk& c = live[t];
            c.out_off = total;
            out_size[t] = c.out.size();
            total += out_size[t];
            const size_t sz = out_size[t];
            const uint32_t keep = (uint32_t)std::min<size_t>(WIN, sz);
            // window after this chunk = last WIN bytes of (previous window ++ this chunk's output)
            uint8_t w[WIN];
            uint32_t have = 0;
            if (keep < WIN && t > 0) {
                const uint32_t from_prev = std::min<uint32_t>(WIN - keep, live[t - 1].last_n);
                memcpy(w, live[t - 1].last_window + (live[t - 1].last_n - from_prev), from_prev);
                have = from_prev;
            }
            const uint16_t* s = c.out.p + WIN + (sz - keep);
            for (uint32_t i = 0; i < keep; ++i) {
                uint16_t x = s[i];
                if (x & UNKNOWN) {
                    if (t == 0) return false;
                    const uint32_t off = x & 0x7fffu;  // index into the previous window, which holds its last last_n bytes
                    const Chunk& pc = live[t - 1];
                    if (off < WIN - pc.last_n) return false;  // before the start of the text
                    x = pc.last_window[off - (WIN - pc.last_n)];
                }
                w[have + i] = (uint8_t)x;
            }
            c.last_n = have + keep;
            memcpy(c.last_window, w, c.last_n);
        }
        const double tc0 = now();
        {   // the text grows by this group's output (realloc moves large blocks by remapping, not by copying)
            char* q = (char*)realloc(text, total + 1);
            if (!q) return false;
            text = q;
        }
        parallel_for_io(g1 - g0, [&](size_t i) {
            const size_t t = g0 + i;
            Chunk& c = live[t];
            const size_t sz = out_size[t];
            const uint16_t* s = c.out.p + WIN;
            uint8_t* d = (uint8_t*)text + c.out_off;
            const Chunk* pc = t ? &live[t - 1] : nullptr;
            // one table turns every symbol into its byte: 0..255 map to themselves, 0x8000 | i to byte i of the window
            // before this chunk (a reference before the start of the text maps to 0 and fails the CRC check below)
            std::vector<uint8_t> lut(65536, 0);
            for (uint32_t v = 0; v < 256; ++v) lut[v] = (uint8_t)v;
            if (pc)
                for (uint32_t off = WIN - pc->last_n; off < WIN; ++off) lut[UNKNOWN | off] = pc->last_window[off - (WIN - pc->last_n)];
            static const bool avx2 = __builtin_cpu_supports("avx2");
            const double a0 = timing ? now() : 0;
            if (avx2) translate_avx2(s, d, sz, lut.data());
            else translate_scalar(s, d, sz, lut.data());
            const double a1 = timing ? now() : 0;
            {   // the chunk's output is cut at the member ends inside it; the pieces are combined per member afterwards
                size_t from = 0;
                for (size_t k = 0; k <= c.members.size(); ++k) {
                    const size_t to = k < c.members.size() ? c.members[k].out_pos : sz;
                    crcs[t].push_back(to > from ? crc32_fast(d + from, to - from) : 0u);
                    from = to;
                }
            }
            const double a2 = timing ? now() : 0;
            c.out.release();
            if (timing) {
                tt[t] = a1 - a0;
                tc[t] = a2 - a1;
                tf[t] = now() - a2;
            }
        }, n_threads);
        ms_decode += tb - ta;
        ms_chain += tc0 - tb;
        ms_translate += now() - tc0;
    }
    // the last chunk must have run to the end of the file, through the last member's trailer
    if ((live[T - 1].end_bit >> 3) != n || live[T - 1].members.empty()) return false;
    // every member's CRC-32 and length (mod 2^32), pieces combined across chunk boundaries
    bool ok = true;
    {
        uint32_t run_crc = 0;
        size_t run_len = 0, n_members = 0;
        for (size_t t = 0; t < T && ok; ++t) {
            const Chunk& c = live[t];
            const size_t sz = out_size[t];
            size_t from = 0;
            for (size_t i = 0; i <= c.members.size() && ok; ++i) {
                const size_t to = i < c.members.size() ? c.members[i].out_pos : sz;
                const size_t len = to - from;
                if (len) {
                    run_crc = run_len ? (uint32_t)crc32_combine(run_crc, crcs[t][i], (z_off_t)len) : crcs[t][i];
                    run_len += len;
                }
                if (i < c.members.size()) {
                    ok = run_crc == c.members[i].crc && (uint32_t)run_len == c.members[i].len;
                    run_crc = 0;
                    run_len = 0;
                    ++n_members;
                }
                from = to;
            }
        }
        ok = ok && run_len == 0 && n_members > 0;
    }
    if (!ok) return false;
    text[total] = 0;
    guard.keep = true;
    *out = text;
    *out_n = total;
    if (timing) {
        double st = 0, sc = 0, sf = 0;
        for (size_t t = 0; t < T; ++t) {
            st += tt[t];
            sc += tc[t];
            sf += tf[t];
        }
        fprintf(stderr, "[drprg-cuda] parallel gunzip: %zu chunks in groups of %zu, %.1f MB -> %.1f MB, block search %.1f ms, decode %.1f ms, "
                        "window chain %.1f ms, translate + crc %.1f ms (thread-ms: translate %.0f, crc %.0f, free %.0f)\n", T, group,
                n / 1e6, total / 1e6, t1 - t0, ms_decode, ms_chain, ms_translate, st, sc, sf);
    }
    return true;
}

}  // namespace drprg
