// Hand-written sm_100a kernels for the map hot path: read sketching + index lookup (S1+S2), hit
// clustering (S3/S4), k-mer coverage (S5), ML path (S7) and genotyping (S8).  These replace the
// per-read and per-locus loops of `pandora map` that drprg launches at
// /root/reference/src/lib.rs:580-642 (argv :594-609, src/predict.rs:288-294); stage semantics
// follow pandora's Seq::minimizer_sketch, add_read_hits, define_clusters, filter_clusters(2),
// add_hits_to_kmergraphs, KmerGraphWithCoverage::find_max_path and SampleInfo.
// This file: S3 - S5 (hit ordering, clustering + filters, k-mer coverage).
#include <cub/cub.cuh>
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "kernels_common.cuh"

namespace drprg {

// ============================================================================================
// hit ordering: stable LSD radix sort on lo then hi  ==  order by (hi, lo)
// ============================================================================================
size_t sort_hits_temp_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int64_t)n, 0, 64);
    return bytes;
}

// When read, locus, strand, read_start and k-mer node fit 64 bits together (they do for every BASELINE shape: 46 bits
// for 1 M x 150 bp reads on a 30-locus panel) the hits are packed into ONE key, sorted keys-only over exactly the bits
// in use (6 radix passes of 8 B instead of 10 passes of 16 B) and unpacked again.
struct HitPacking {
    int knode_bits, start_bits, prg_bits, read_bits;
    __host__ __device__ int total() const { return knode_bits + start_bits + 1 + prg_bits + read_bits; }
};

__global__ void pack_hits_kernel(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo,
                                 unsigned long long n, HitPacking B, unsigned long long* __restrict__ key) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long h = hi[i], l = lo[i];
    unsigned long long k = h >> 32;                                  // read
    k = (k << B.prg_bits) | ((h >> 16) & 0xffffull);                 // locus
    k = (k << 1) | ((h >> 15) & 1ull);                               // !forward
    k = (k << B.start_bits) | (l >> 32);                             // read_start
    k = (k << B.knode_bits) | (l & 0xffffffffull);                   // k-mer node rank
    key[i] = k;
}

__global__ void unpack_hits_kernel(const unsigned long long* __restrict__ key, unsigned long long n, HitPacking B,
                                   unsigned long long* __restrict__ hi, unsigned long long* __restrict__ lo) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = key[i];
    const unsigned long long knode = k & ((1ull << B.knode_bits) - 1ull);
    k >>= B.knode_bits;
    const unsigned long long start = k & ((1ull << B.start_bits) - 1ull);
    k >>= B.start_bits;
    const unsigned long long rev = k & 1ull;
    k >>= 1;
    const unsigned long long prg = k & ((1ull << B.prg_bits) - 1ull);
    k >>= B.prg_bits;
    hi[i] = (k << 32) | (prg << 16) | (rev << 15);
    lo[i] = (start << 32) | knode;
}

void sort_hits(void* d_temp, size_t temp_bytes, unsigned long long* hi_in, unsigned long long* lo_in,
               unsigned long long* hi_tmp, unsigned long long* lo_tmp, uint64_t n, int read_bits, int start_bits,
               int knode_bits, int prg_bits, cudaStream_t st) {
    if (n == 0) return;
    const HitPacking B{knode_bits, start_bits, prg_bits, read_bits};
    static const bool packed_on = [] {
        const char* e = getenv("DRPRG_PACKED_SORT");
        return !e || atoi(e) != 0;
    }();
    if (packed_on && B.total() <= 64 && knode_bits < 32 && start_bits < 32 && prg_bits <= 16) {
        const unsigned grid = (unsigned)((n + 255) / 256);
        pack_hits_kernel<<<grid, 256, 0, st>>>(hi_in, lo_in, n, B, hi_tmp);
        size_t need = temp_bytes;
        cub::DeviceRadixSort::SortKeys(d_temp, need, hi_tmp, lo_tmp, (int64_t)n, 0, B.total(), st);
        unpack_hits_kernel<<<grid, 256, 0, st>>>(lo_tmp, n, B, hi_in, lo_in);
        g_launches += 3;
        return;
    }
    // pass A: key = lo (start | knode), value = hi.  knode occupies bits [0,knode_bits), start [32,32+start_bits)
    cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, lo_in, lo_tmp, hi_in, hi_tmp, (int64_t)n, 0, 32 + start_bits, st);
    // pass B: key = hi (read | prg | strand), value = lo
    cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, hi_tmp, hi_in, lo_tmp, lo_in, (int64_t)n, 15, 32 + read_bits, st);
    g_launches += 2;
}

// ============================================================================================
// S3 + S4 : clustering.  Hits are sorted (read, prg, fwd-first, read_start, knode), so a read's
// hits are contiguous; the thread sitting on a read's first hit walks that read: splits clusters
// (pandora define_clusters), applies the size threshold, then filter_clusters (adjacent pairs in
// clusterComp order) and filter_clusters2 (by decreasing size, drop clusters whose read span is
// already covered).  Reads carry tens of hits and a handful of clusters, so per-read work is tiny.
// ============================================================================================
__device__ __forceinline__ uint32_t hit_read(unsigned long long hi) { return (uint32_t)(hi >> 32); }
__device__ __forceinline__ uint32_t hit_prg(unsigned long long hi) { return (uint32_t)(hi >> 16) & 0xffffu; }
__device__ __forceinline__ uint32_t hit_fwd(unsigned long long hi) { return (((uint32_t)hi >> 15) & 1u) ^ 1u; }
__device__ __forceinline__ uint32_t hit_start(unsigned long long lo) { return (uint32_t)(lo >> 32); }

__global__ void cluster_filter_kernel(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo,
                                      unsigned long long n, uint32_t max_diff, const uint32_t* __restrict__ thresh,
                                      uint32_t* __restrict__ clist, uint32_t* __restrict__ clist2,
                                      uint32_t* __restrict__ cend, uint8_t* __restrict__ calive,
                                      uint8_t* __restrict__ kept, int32_t* __restrict__ locus_reads) {
    const unsigned long long i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= n) return;
    const uint32_t read = hit_read(hi[i0]);
    if (i0 > 0 && hit_read(hi[i0 - 1]) == read) return;  // not the first hit of its read
    // ---- define_clusters
    uint32_t ncl = 0;
    unsigned long long b = i0, i = i0 + 1;
    while (true) {
        bool split = true, end_of_read = true;
        if (i < n && hit_read(hi[i]) == read) {
            end_of_read = false;
            const unsigned long long hp = hi[i - 1], hc = hi[i];
            const long long d = (long long)hit_start(lo[i]) - (long long)hit_start(lo[i - 1]);
            split = (hit_prg(hp) != hit_prg(hc)) || (hit_fwd(hp) != hit_fwd(hc)) || ((d < 0 ? -d : d) > (long long)max_diff);
        }
        if (split) {
            const uint32_t size = (uint32_t)(i - b);
            if (size > thresh[hit_prg(hi[b])]) {
                clist[i0 + ncl] = (uint32_t)(b - i0);
                cend[b] = (uint32_t)(i - i0);
                calive[b] = 1;
                ++ncl;
            }
            b = i;
        }
        if (end_of_read) break;
        ++i;
    }
    if (ncl == 0) return;
    auto c_first = [&](uint32_t c) { return hit_start(lo[i0 + c]); };
    auto c_last = [&](uint32_t c) { return hit_start(lo[i0 + cend[i0 + c] - 1]); };
    auto c_size = [&](uint32_t c) { return cend[i0 + c] - c; };
    auto c_prg = [&](uint32_t c) { return hit_prg(hi[i0 + c]); };
    auto c_fwd = [&](uint32_t c) { return hit_fwd(hi[i0 + c]); };
    // ---- filter_clusters: order (first start, size desc, prg, fwd asc); adjacent-pair sweep
    if (ncl > 1) {
        auto before = [&](uint32_t x, uint32_t y) {
            if (c_first(x) != c_first(y)) return c_first(x) < c_first(y);
            if (c_size(x) != c_size(y)) return c_size(x) > c_size(y);
            if (c_prg(x) != c_prg(y)) return c_prg(x) < c_prg(y);
            return c_fwd(x) < c_fwd(y);
        };
        for (uint32_t a = 1; a < ncl; ++a) {  // insertion sort of clist[i0 .. i0+ncl)
            const uint32_t v = clist[i0 + a];
            uint32_t j = a;
            while (j > 0 && before(v, clist[i0 + j - 1])) {
                clist[i0 + j] = clist[i0 + j - 1];
                --j;
            }
            clist[i0 + j] = v;
        }
        uint32_t prev = clist[i0];
        for (uint32_t t = 1; t < ncl; ++t) {
            const uint32_t cur = clist[i0 + t];
            const bool cond = (c_prg(cur) == c_prg(prev) && c_fwd(cur) != c_fwd(prev)) || (c_last(cur) <= c_last(prev));
            if (cond) {
                if (c_size(prev) >= c_size(cur)) {
                    calive[i0 + cur] = 0;
                    continue;
                }
                calive[i0 + prev] = 0;
            }
            prev = cur;
        }
        // ---- filter_clusters2
        uint32_t n2 = 0;
        for (uint32_t t = 0; t < ncl; ++t)
            if (calive[i0 + clist[i0 + t]]) clist2[i0 + n2++] = clist[i0 + t];
        auto before2 = [&](uint32_t x, uint32_t y) {
            if (c_size(x) != c_size(y)) return c_size(x) > c_size(y);
            if (c_first(x) != c_first(y)) return c_first(x) < c_first(y);
            if (c_prg(x) != c_prg(y)) return c_prg(x) < c_prg(y);
            return c_fwd(x) < c_fwd(y);
        };
        for (uint32_t a = 1; a < n2; ++a) {
            const uint32_t v = clist2[i0 + a];
            uint32_t j = a;
            while (j > 0 && before2(v, clist2[i0 + j - 1])) {
                clist2[i0 + j] = clist2[i0 + j - 1];
                --j;
            }
            clist2[i0 + j] = v;
        }
        for (uint32_t t = 1; t < n2; ++t) {
            const uint32_t c = clist2[i0 + t];
            const uint32_t z = c_last(c);
            uint32_t cur = c_first(c);
            bool contained = true;
            while (cur < z) {
                uint32_t best = cur;
                for (uint32_t u = 0; u < t; ++u) {
                    const uint32_t pc = clist2[i0 + u];
                    if (!calive[i0 + pc]) continue;  // erased clusters never marked the read
                    if (c_first(pc) <= cur && cur < c_last(pc)) best = max(best, c_last(pc));
                }
                if (best == cur) {
                    contained = false;
                    break;
                }
                cur = best;
            }
            if (contained) calive[i0 + c] = 0;
        }
    }
    // ---- add_clusters_to_pangraph: mark kept hits, count supporting reads per locus
    for (uint32_t t = 0; t < ncl; ++t) {
        const uint32_t c = clist[i0 + t];
        if (!calive[i0 + c]) continue;
        const uint32_t e = cend[i0 + c];
        for (uint32_t j = c; j < e; ++j) kept[i0 + j] = 1;
        atomicAdd(locus_reads + c_prg(c), 1);
    }
}

void launch_cluster_filter(const unsigned long long* hi, const unsigned long long* lo, uint64_t n, uint32_t max_diff,
                           const uint32_t* d_thresh_per_prg, uint32_t* d_clist, uint32_t* d_clist2, uint32_t* d_cend,
                           uint8_t* d_calive, uint8_t* d_kept, int32_t* d_locus_reads, cudaStream_t st) {
    if (n == 0) return;
    cudaMemsetAsync(d_kept, 0, n, st);
    cudaMemsetAsync(d_calive, 0, n, st);
    const int threads = 128;
    cluster_filter_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(hi, lo, n, max_diff, d_thresh_per_prg,
                                                                                      d_clist, d_clist2, d_cend, d_calive,
                                                                                      d_kept, d_locus_reads);
    ++g_launches;
}

// ============================================================================================
// S5 : coverage.  key = 2 * global knode + (reverse ? 1 : 0) for kept hits; sorted keys; the thread
// on the first element of each run finds the run's end by binary search and adds the run length
// to that counter (one writer per counter: no atomics).
// ============================================================================================
__global__ void cov_keys_kernel(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo,
                                const uint8_t* __restrict__ kept, unsigned long long n,
                                const uint32_t* __restrict__ knode_base, uint32_t* __restrict__ keys, uint32_t sentinel) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key = sentinel;  // above every real key: discarded hits sort to the end
    if (kept[i]) {
        const unsigned long long h = hi[i];
        const uint32_t g = knode_base[hit_prg(h)] + (uint32_t)lo[i];
        key = 2u * g + (hit_fwd(h) ^ 1u);
    }
    keys[i] = key;
}

__global__ void cov_runs_kernel(const uint32_t* __restrict__ keys, unsigned long long n, int32_t* __restrict__ cov,
                                unsigned long long* __restrict__ n_kept, uint32_t sentinel) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t key = keys[i];
    if (i > 0 && keys[i - 1] == key) return;
    if (key == sentinel) {  // first discarded hit: everything before it was kept
        *n_kept += i;
        return;
    }
    unsigned long long lo_ = i, hi_ = n;  // first index with keys[idx] > key
    while (lo_ < hi_) {
        const unsigned long long mid = (lo_ + hi_) >> 1;
        if (keys[mid] <= key) lo_ = mid + 1;
        else hi_ = mid;
    }
    cov[key] += (int32_t)(lo_ - i);
    if (lo_ == n) *n_kept += n;  // no discarded hits at all
}

size_t sort_cov_temp_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)n, 0, 32);
    return bytes;
}

void launch_coverage(const unsigned long long* hi, const unsigned long long* lo, const uint8_t* kept, uint64_t n,
                     const uint32_t* d_knode_base, uint32_t* d_keys, uint32_t* d_keys_sorted, void* d_temp,
                     size_t temp_bytes, int key_bits, int32_t* d_cov, unsigned long long* d_n_kept, cudaStream_t st) {
    if (n == 0) return;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    // real keys use key_bits bits; discarded hits carry 1 << key_bits, so only key_bits + 1 bits are sorted
    const int kb = key_bits < 31 ? key_bits : 31;
    const uint32_t sentinel = kb < 31 ? (1u << kb) : 0xffffffffu;
    cov_keys_kernel<<<blocks, threads, 0, st>>>(hi, lo, kept, n, d_knode_base, d_keys, sentinel);
    cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, d_keys, d_keys_sorted, (int64_t)n, 0, kb + 1, st);
    cov_runs_kernel<<<blocks, threads, 0, st>>>(d_keys_sorted, n, d_cov, d_n_kept, sentinel);
    g_launches += 3;
}

}  // namespace drprg
