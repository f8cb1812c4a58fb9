// Host-side FASTQ framing for the ingest pipeline (the reads file of
// /root/reference/src/predict.rs:166-170, 288-294).  A 150 bp FASTQ record is ~315 bytes of which only the 150 sequence
// bytes matter to the map step; sending the whole text over PCIe made the drop-in call transfer-bound.  The host threads
// therefore find the record structure (newline bitmasks, 32 bytes per AVX2 step) and copy ONLY the sequence lines into
// the pinned staging buffer; packing to 2 bits, the non-ACGT check and everything else stay on the device.
#include "fastq_frame.hpp"

#include <immintrin.h>

#include <unistd.h>

#include <atomic>
#include <cstring>
#include <stdexcept>

#include "genotype_host.hpp"  // parallel_for_io

namespace drprg {
namespace {

// offsets of the '\n' bytes of p[0..n) into nl (room for cap), stops early when nl is full; returns the count and sets
// `scanned` to the number of bytes looked at
__attribute__((target("avx2"))) size_t newlines_avx2(const char* p, size_t n, uint32_t* nl, size_t cap, size_t& scanned) {
    size_t k = 0, i = 0;
    const __m256i needle = _mm256_set1_epi8('\n');
    for (; i + 32 <= n && k + 32 <= cap; i += 32) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(p + i));
        uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, needle));
        while (m) {
            nl[k++] = (uint32_t)i + (uint32_t)__builtin_ctz(m);
            m &= m - 1;
        }
    }
    if (i + 32 > n)
        for (; i < n && k < cap; ++i)
            if (p[i] == '\n') nl[k++] = (uint32_t)i;
    scanned = i;
    return k;
}
size_t newlines_scalar(const char* p, size_t n, uint32_t* nl, size_t cap, size_t& scanned) {
    size_t k = 0;
    const char* q = p;
    const char* e = p + n;
    while (q < e && k < cap) {
        const char* h = (const char*)memchr(q, '\n', (size_t)(e - q));
        if (!h) {
            q = e;
            break;
        }
        nl[k++] = (uint32_t)(h - p);
        q = h + 1;
    }
    scanned = (size_t)(q - p);
    return k;
}

}  // namespace

size_t fastq_next_record(const char* t, size_t n, size_t p) {
    if (p == 0) return 0;
    if (p >= n) return n;
    // start of the line after the one containing p - 1
    const char* h = (const char*)memchr(t + p - 1, '\n', n - (p - 1));
    if (!h) return n;
    size_t q = (size_t)(h - t) + 1;
    while (q < n) {
        const char* e1 = (const char*)memchr(t + q, '\n', n - q);
        if (!e1) return n;
        if (t[q] == '@') {  // a header line is followed, two lines later, by a '+' line
            const size_t l2 = (size_t)(e1 - t) + 1;
            const char* e2 = l2 < n ? (const char*)memchr(t + l2, '\n', n - l2) : nullptr;
            if (e2 && (size_t)(e2 - t) + 1 < n && e2[1] == '+') return q;
        }
        q = (size_t)(e1 - t) + 1;
    }
    return n;
}

inline bool FastqFramer::line(const char* ls, const char* le) {
    switch (phase_) {
        case 0:
            if (le - ls < 2 || *ls != '@') return false;  // blank line, empty header or broken framing
            break;
        case 1: {
            size_t len = (size_t)(le - ls);
            if (len && le[-1] == '\r') --len;
            if (fill_ + len > cap_ || base_ + fill_ + len > 0xfffffff0ull) return false;
            memcpy(out_ + fill_, ls, len);
            starts_.push_back(base_ + (uint32_t)fill_);
            fill_ += len;
            seq_len_ = len;
            break;
        }
        case 2:
            if (le == ls || *ls != '+') return false;
            break;
        default: {
            size_t len = (size_t)(le - ls);
            if (len && le[-1] == '\r') --len;
            if (len != seq_len_) return false;  // wrapped or damaged record
            lens_.push_back((uint32_t)seq_len_);
            if (S_.n_reads == 0) S_.first_len = (uint32_t)seq_len_;
            ++S_.n_reads;
            S_.max_len = std::max<uint32_t>(S_.max_len, (uint32_t)seq_len_);
            S_.min_len = std::min<uint32_t>(S_.min_len, (uint32_t)seq_len_);
            S_.total_bases += seq_len_;
            S_.seq_bytes = fill_;
            break;
        }
    }
    phase_ = (phase_ + 1) & 3u;
    return true;
}

bool FastqFramer::feed(const char* p, size_t n, size_t& used) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    constexpr size_t NL = 4096;
    uint32_t nl[NL];
    size_t line_start = 0, at = 0;
    while (at < n) {
        const size_t stretch = std::min<size_t>(n - at, 1u << 30);
        size_t scanned = 0;
        const size_t k = avx2 ? newlines_avx2(p + at, stretch, nl, NL, scanned) : newlines_scalar(p + at, stretch, nl, NL, scanned);
        for (size_t i = 0; i < k; ++i) {
            const size_t le = at + nl[i];
            if (!line(p + line_start, p + le)) return false;
            line_start = le + 1;
        }
        at += scanned;
    }
    used = line_start;
    return true;
}

bool FastqFramer::finish(const char* p, size_t n, bool end_of_file) {
    if (n) {  // text after the last newline: the final line of a file without a trailing newline
        if (!end_of_file) return false;
        if (!line(p, p + n)) return false;
    }
    return phase_ == 0;  // the slice must end on a record boundary
}

size_t TextSource::read(char* buf, size_t at, size_t n) const {
    n = at < size ? std::min(n, size - at) : 0;
    if (mem) {
        memcpy(buf, mem + at, n);
        return n;
    }
    size_t done = 0;
    while (done < n) {
        const ssize_t r = pread(fd, buf + done, n - done, (off_t)(at + done));
        if (r <= 0) break;
        done += (size_t)r;
    }
    return done;
}

size_t TextSource::boundary(size_t b, std::vector<char>& scratch) const {
    if (b == 0) return 0;
    if (b >= size) return size;
    if (mem) return fastq_next_record(mem, size, b);
    for (size_t window = 64u << 10;; window *= 4) {
        const size_t want = std::min(window, size - (b - 1));
        if (scratch.size() < want) scratch.resize(want);
        if (read(scratch.data(), b - 1, want) != want) return SIZE_MAX;
        const size_t q = fastq_next_record(scratch.data(), want, 1);
        if (q < want) return b - 1 + q;
        if (b - 1 + want >= size) return size;
        if (window > (512u << 20)) return SIZE_MAX;
    }
}

bool fastq_frame_text(const TextSource& src, uint32_t threads, char* seq_buf, std::vector<FramedSlice>& sl,
                      const std::function<bool(size_t, size_t)>& on_slice) {
    static const size_t slice_bytes = [] {  // tests lower it to cut small files into many slices
        const char* e = getenv("DRPRG_FRAME_SLICE");
        return e && atol(e) > 0 ? (size_t)atol(e) : (size_t)(1u << 20);
    }();
    const size_t S = std::max<size_t>(1, std::min<size_t>(src.size / slice_bytes, 8192));
    sl.clear();
    sl.resize(S);
    std::atomic<bool> failed{false};
    parallel_for_io(S, [&](size_t s) {
        FramedSlice& Z = sl[s];
        if (failed.load(std::memory_order_relaxed)) return;
        thread_local std::vector<char> buf;
        const size_t lo = src.boundary(src.size / S * s + std::min(s, src.size % S), buf);
        const size_t hi = s + 1 == S ? src.size : src.boundary(src.size / S * (s + 1) + std::min(s + 1, src.size % S), buf);
        if (lo == SIZE_MAX || hi == SIZE_MAX) {
            Z.state = -2;
            failed = true;
            return;
        }
        Z.lo = lo;
        Z.state = 1;
        if (lo >= hi) return;
        Z.starts.reserve((hi - lo) / 256 + 16);
        Z.lens.reserve((hi - lo) / 256 + 16);
        FastqFramer F(seq_buf + lo / 2, (hi - lo) / 2, (uint32_t)(lo / 2), Z.starts, Z.lens);
        bool ok = true;
        if (src.mem) {
            size_t used = 0;
            ok = F.feed(src.mem + lo, hi - lo, used) && F.finish(src.mem + lo + used, hi - lo - used, hi == src.size);
        } else {
            constexpr size_t CH = 512u << 10;  // stays in the core's L2: the file bytes reach DRAM once (page cache -> here)
            if (buf.size() < 2 * CH) buf.resize(2 * CH);
            size_t at = lo, carry = 0;
            while (ok && at < hi) {
                const size_t want = std::min(CH, hi - at);
                if (buf.size() < carry + want) buf.resize(std::max(buf.size() * 2, carry + want));
                if (src.read(buf.data() + carry, at, want) != want) {
                    Z.state = -2;
                    failed = true;
                    return;
                }
                at += want;
                size_t used = 0;
                ok = F.feed(buf.data(), carry + want, used);
                carry = carry + want - used;
                if (ok && carry) memmove(buf.data(), buf.data() + used, carry);
            }
            ok = ok && F.finish(buf.data(), carry, hi == src.size);
        }
        if (!ok) {
            Z.state = -1;
            failed = true;
            return;
        }
        Z.st = F.stats();
        if (Z.st.seq_bytes && on_slice && !on_slice(lo / 2, Z.st.seq_bytes)) failed = true;
    }, std::max<size_t>(1, threads));
    for (const FramedSlice& Z : sl)
        if (Z.state == -2) throw std::runtime_error("reads file: read error");
    return !failed;
}

}  // namespace drprg
