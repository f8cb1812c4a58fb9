// Host-side FASTQ framing (fastq_frame.cpp): record structure found on the host, sequence lines only go to the device.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace drprg {

struct FrameStats {
    size_t n_reads = 0, seq_bytes = 0;
    uint64_t total_bases = 0;
    uint32_t max_len = 0, min_len = 0xffffffffu, first_len = 0;
};

// first record start at or after byte p of the text t[0..n): a line starting with '@' whose line after next starts with '+'
// (a quality line may start with '@', but the line two below it is a sequence line).  n if there is none.
size_t fastq_next_record(const char* t, size_t n, size_t p);

// Streaming framer of strict 4-line records.  The text of one slice (which starts at a record start) is fed in pieces;
// every piece must start at a line start.  The sequence lines are copied back to back into seq_out and (start, length)
// of every read appended to `starts` / `lens` (start = offset in seq_out + start_base).  A false return means the text
// is not strict 4-line FASTQ (wrapped sequence, blank line, damaged record) or seq_out is full: the caller falls back to
// the general parser.
class FastqFramer {
  public:
    FastqFramer(char* seq_out, size_t seq_cap, uint32_t start_base, std::vector<uint32_t>& starts, std::vector<uint32_t>& lens)
        : out_(seq_out), cap_(seq_cap), base_(start_base), starts_(starts), lens_(lens) {}
    // consumes the whole lines of p[0..n); `used` = bytes consumed (everything up to and including the last '\n')
    bool feed(const char* p, size_t n, size_t& used);
    // the end of the slice: `p[0..n)` is what feed() left over (a last line without '\n', only legal at the end of the file)
    bool finish(const char* p, size_t n, bool end_of_file);
    const FrameStats& stats() const { return S_; }

  private:
    bool line(const char* ls, const char* le);
    // The sequence bytes go through a small cache-resident staging buffer and leave it as whole 64-byte lines written
    // with non-temporal stores: an ordinary memcpy into the (pinned, never re-read by the CPU) output would first READ
    // every destination line into the cache — a quarter of the framing thread's memory traffic.
    void put(const char* p, size_t len);
    void flush(bool all);
    static constexpr size_t STAGE = 8192;
    alignas(64) char stage_[STAGE + 64];
    size_t staged_ = 0;    // bytes in stage_ (from stage_ + phase_off_)
    size_t flushed_ = 0;   // bytes of the output already written
    size_t phase_off_ = 0; // (out_ + flushed_) & 63 at the last flush: stage_ keeps the destination's line phase
    char* out_;
    size_t cap_, fill_ = 0;
    uint32_t base_;
    std::vector<uint32_t>&starts_, &lens_;
    uint32_t phase_ = 0;  // 0 header, 1 sequence, 2 plus, 3 quality
    size_t seq_len_ = 0;
    bool plus_plain_ = false, plain_prev_ = false;  // the last '+' line was a lone '+'; the last record may serve as a template
    FrameStats S_;
};

// The text to frame: a plain file (read in pieces with pread by the framing threads) or text already in host memory (a
// gzip file inflated ahead of time).
struct TextSource {
    int fd = -1;
    const char* mem = nullptr;
    size_t size = 0;
    size_t origin = 0;  // file offset of this source's byte 0 (a wave of a larger file); `mem` already points at byte 0
    size_t read(char* buf, size_t at, size_t n) const;                 // bytes [at, at + n) -> buf; returns the count read
    size_t boundary(size_t b, std::vector<char>& scratch) const;       // first record start at or after byte b (size = none); SIZE_MAX: read error
};

struct FramedSlice {
    std::vector<uint32_t> starts, lens;  // per read: offset of its sequence in the sequence buffer, length
    FrameStats st;
    size_t lo = 0;                       // the slice's first byte in the text; its sequences start at lo / 2 in the buffer
    int state = 0;                       // 1 framed, -1 not strict FASTQ, -2 read error
};

// Cuts the text into slices at record starts and frames them side by side on the IO pool.  The sequence lines of the
// slice [lo, hi) go to seq_buf + lo / 2 (seq_buf holds size / 2 + 64 bytes): a record's sequence is less than half of its
// bytes, so the slices cannot collide.  on_slice(offset, bytes), if given, is called by the thread that framed a slice
// (e.g. to queue its H2D copy while the other slices are still being framed); returning false aborts.
// false = not strict 4-line FASTQ; throws on read errors.
bool fastq_frame_text(const TextSource& src, uint32_t threads, char* seq_buf, std::vector<FramedSlice>& slices,
                      const std::function<bool(size_t, size_t)>& on_slice = nullptr);

}  // namespace drprg
