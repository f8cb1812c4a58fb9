// Hand-written sm_100a kernels for the map hot path of `pandora map` that drprg launches at
// /root/reference/src/lib.rs:580-642 (argv :594-609, src/predict.rs:288-294); stage semantics follow pandora's
// add_read_hits, define_clusters, filter_clusters(2), add_clusters_to_pangraph and
// pangenome::Graph::add_hits_to_kmergraphs.
//
// This file: S3 - S5 (hit grouping, clustering + filters, k-mer coverage) with no library sort and no host round trip.
//
// The lookup kernels append hits in no particular order.  pandora orders them (read, prg, strand, read_start, k-mer
// node) and everything downstream is per read, so nothing here sorts globally:
//   count_reads_kernel    per-read hit counters; the first hit of a read enrols it in the list of active reads
//                         (~1.4 % of whole-genome reads touch the panel)
//   assign_kernel         every active read gets a contiguous slice of the grouped-hit array (order of the slices is
//                         immaterial: all later results are sums) and a class: <= 64 hits -> a warp, more -> a CTA
//   scatter_kernel        hits move into their read's slice as ONE 64-bit key prg | strand | read_start | k-mer node whose
//                         integer order is pandora's MinimizerHit order within a read; the per-read counters count back
//                         down to zero, so they never need a memset
//   cluster_warp_kernel   one warp per read: register bitonic sort of its <= 64 keys by shuffles, cluster boundaries by
//                         ballot (define_clusters), size thresholds, then filter_clusters / filter_clusters2 on the
//                         surviving clusters (usually one)
//   cluster_cta_kernel    one CTA per long read: the same with a shared-memory (or, beyond 8 k hits, in-place global)
//                         bitonic sort and a block scan for the cluster boundaries
//   both emit, for every kept hit, the coverage key 2 * k-mer node + strand, and one key 2N + locus per kept cluster
//   cov_tile_kernel       S5: sorted-key reduction without atomics.  A persistent CTA sorts tiles of 8 k keys in shared
//                         memory, turns them into (key, run length) and adds the run lengths to its PRIVATE partial
//                         accumulator (each key occurs once per tile after the sort: plain adds, no conflicts)
//   cov_merge_kernel      column sum of the partials into the sample's accumulator — or, on a read-sharded run, straight
//                         into the ROOT GPU's accumulator over NVLink with red.global.add (integers: bit-exact for any
//                         number of GPUs), which is the whole "allreduce" of the path
// Overflowing buffers are detected at the end of the batch (the kernels skip their work when a counter exceeds its
// capacity, the accumulators stay untouched) and the batch is redone with larger buffers.
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "kernels_common.cuh"

namespace drprg {

// ---- grouped-hit key: prg 16 | reverse 1 | read_start 25 | k-mer node rank 22 ---------------------------------
__device__ __forceinline__ unsigned long long group_key(unsigned long long hi, unsigned long long lo) {
    const unsigned long long prg = (hi >> 16) & 0xffffull, rev = (hi >> 15) & 1ull;
    return (prg << 48) | (rev << 47) | ((lo >> 32) << GKEY_KNODE_BITS) | (lo & ((1ull << GKEY_KNODE_BITS) - 1ull));
}
__device__ __forceinline__ uint32_t gk_pf(unsigned long long k) { return (uint32_t)(k >> 47); }                // prg << 1 | reverse
__device__ __forceinline__ uint32_t gk_prg(unsigned long long k) { return (uint32_t)(k >> 48); }
__device__ __forceinline__ uint32_t gk_rev(unsigned long long k) { return (uint32_t)(k >> 47) & 1u; }
__device__ __forceinline__ uint32_t gk_start(unsigned long long k) { return (uint32_t)(k >> GKEY_KNODE_BITS) & ((1u << GKEY_START_BITS) - 1u); }
__device__ __forceinline__ uint32_t gk_knode(unsigned long long k) { return (uint32_t)k & ((1u << GKEY_KNODE_BITS) - 1u); }

// counters (unsigned long long): the layout is shared with sketch.cu / capi.cu through kernels.cuh (CTR_*)
__device__ __forceinline__ bool batch_overflowed(const unsigned long long* __restrict__ ctr, PostCaps C) {
    return ctr[CTR_HITS] > C.hit_cap || ctr[CTR_QUEUE_NEED] > C.queue_cap;
}

// ============================================================================================
// grouping by read
// ============================================================================================
__global__ void count_reads_kernel(const unsigned long long* __restrict__ hi, unsigned long long* __restrict__ ctr, PostCaps C,
                                   uint32_t read_id_base, int32_t* __restrict__ read_count, uint32_t* __restrict__ act_read) {
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n = ctr[CTR_HITS];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(hi[i] >> 32) - read_id_base;
        if (atomicAdd(read_count + r, 1) == 0) act_read[atomicAdd(ctr + CTR_ACTIVE, 1ull)] = r;
    }
}

__global__ void assign_kernel(unsigned long long* __restrict__ ctr, PostCaps C, const int32_t* __restrict__ read_count,
                              const uint32_t* __restrict__ act_read, uint32_t* __restrict__ act_base,
                              uint32_t* __restrict__ act_count, uint32_t* __restrict__ read_base,
                              uint32_t* __restrict__ big_list) {
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n = ctr[CTR_ACTIVE];
    const int lane = threadIdx.x & 31;
    for (unsigned long long a0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; a0 < n;
         a0 += (unsigned long long)gridDim.x * blockDim.x) {  // warp-uniform trip count
        const unsigned long long a = a0 + lane;
        uint32_t r = 0, c = 0;
        if (a < n) {
            r = act_read[a];
            c = (uint32_t)read_count[r];
        }
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(ctr + CTR_CURSOR, (unsigned long long)incl);  // one slice allocation per warp
        base = __shfl_sync(FULL, base, 31) + (incl - c);
        if (a < n) {
            act_base[a] = (uint32_t)base;
            act_count[a] = c;
            read_base[r] = (uint32_t)base;
            if (c > CLUSTER_WARP_MAX) big_list[atomicAdd(ctr + CTR_BIG, 1ull)] = (uint32_t)a;
        }
    }
}

__global__ void scatter_kernel(const unsigned long long* __restrict__ hi, const unsigned long long* __restrict__ lo,
                               const unsigned long long* __restrict__ ctr, PostCaps C, uint32_t read_id_base,
                               int32_t* __restrict__ read_count, const uint32_t* __restrict__ read_base,
                               unsigned long long* __restrict__ gkey) {
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n = ctr[CTR_HITS];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h = hi[i];
        const uint32_t r = (uint32_t)(h >> 32) - read_id_base;
        const uint32_t slot = read_base[r] + (uint32_t)(atomicSub(read_count + r, 1) - 1);  // the counter ends at 0 again
        gkey[slot] = group_key(h, lo[i]);
    }
}

// ============================================================================================
// S3 + S4 on one read's clusters that passed the size threshold (pandora filter_clusters, filter_clusters2), run by
// ONE WARP.  Cluster t: first / last = read_start of its first / last hit, size, pf = prg << 1 | reverse.
//   filter_clusters : clusters in the order (first, size desc, prg, strand forward-first); adjacent-pair sweep: two
//                     clusters of the same prg on opposite strands, or one ending no later than its predecessor, lose
//                     the smaller one (ties: the later one);
//   filter_clusters2: survivors by decreasing size; a cluster whose read span is already covered by the spans of the
//                     bigger survivors is dropped.
// The orders are total (two clusters of one read cannot agree on prg, strand and first start), so they are built by
// rank counting across the lanes; the sweep is sequential like pandora's.
// ============================================================================================
struct ClusterArrays {
    uint32_t *first, *last, *size, *pf;  // per cluster
    uint32_t *ord, *ord2;                // work: the two orders
    uint32_t* alive;                     // in: 1 for every cluster; out: survivors
};

__device__ void filter_clusters_warp(uint32_t ncl, const ClusterArrays& A, int lane) {
    // ---- filter_clusters
    for (uint32_t c = lane; c < ncl; c += 32) {
        const uint32_t f = A.first[c], s = A.size[c], p = A.pf[c];
        uint32_t rank = 0;
        for (uint32_t d = 0; d < ncl; ++d) {
            const uint32_t fd = A.first[d], sd = A.size[d], pd = A.pf[d];
            const bool before = fd != f ? fd < f : (sd != s ? sd > s : (pd != p ? pd < p : d < c));
            rank += before ? 1u : 0u;
        }
        A.ord[rank] = c;
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t prev = A.ord[0];
        for (uint32_t t = 1; t < ncl; ++t) {
            const uint32_t cur = A.ord[t];
            const uint32_t pc = A.pf[cur], pp = A.pf[prev];
            const bool cond = ((pc >> 1) == (pp >> 1) && (pc & 1u) != (pp & 1u)) || (A.last[cur] <= A.last[prev]);
            if (cond) {
                if (A.size[prev] >= A.size[cur]) {
                    A.alive[cur] = 0;
                    continue;
                }
                A.alive[prev] = 0;
            }
            prev = cur;
        }
    }
    __syncwarp();
    // ---- filter_clusters2: order (size desc, first, prg, strand) among the survivors
    uint32_t n2 = 0;
    for (uint32_t c0 = 0; c0 < ncl; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool live = c < ncl && A.alive[c];
        if (live) {
            const uint32_t f = A.first[c], s = A.size[c], p = A.pf[c];
            uint32_t rank = 0;
            for (uint32_t d = 0; d < ncl; ++d) {
                if (!A.alive[d]) continue;
                const uint32_t fd = A.first[d], sd = A.size[d], pd = A.pf[d];
                const bool before = sd != s ? sd > s : (fd != f ? fd < f : (pd != p ? pd < p : d < c));
                rank += before ? 1u : 0u;
            }
            A.ord2[rank] = c;
        }
        n2 += __popc(__ballot_sync(FULL, live));
    }
    __syncwarp();
    for (uint32_t t = 1; t < n2; ++t) {
        const uint32_t c = A.ord2[t];
        const uint32_t z = A.last[c];
        uint32_t cur = A.first[c];
        bool contained = true;
        while (cur < z) {
            uint32_t best = cur;
            for (uint32_t u = lane; u < t; u += 32) {
                const uint32_t pc = A.ord2[u];
                if (!A.alive[pc]) continue;  // erased clusters never marked the read
                if (A.first[pc] <= cur && cur < A.last[pc]) best = max(best, A.last[pc]);
            }
            best = __reduce_max_sync(FULL, best);
            if (best == cur) {
                contained = false;
                break;
            }
            cur = best;
        }
        if (contained && lane == 0) A.alive[c] = 0;
        __syncwarp();
    }
}

// ---- coverage-key emission shared by the two cluster kernels: one slot allocation per warp --------------------
__device__ __forceinline__ uint32_t warp_alloc(unsigned long long* counter, uint32_t mine, int lane, uint32_t& total) {
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += t;
    }
    total = __shfl_sync(FULL, incl, 31);
    unsigned long long base = 0;
    if (total && lane == 31) base = atomicAdd(counter, (unsigned long long)total);
    base = __shfl_sync(FULL, base, 31);
    return (uint32_t)base + (incl - mine);
}

// ============================================================================================
// one warp per read with <= 64 hits
// ============================================================================================
constexpr int CW_WARPS = 8;

__device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int m) {
    return ((unsigned long long)__shfl_xor_sync(FULL, (uint32_t)(v >> 32), m) << 32) | __shfl_xor_sync(FULL, (uint32_t)v, m);
}
__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int d) {
    return ((unsigned long long)__shfl_up_sync(FULL, (uint32_t)(v >> 32), d) << 32) | __shfl_up_sync(FULL, (uint32_t)v, d);
}
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    return ((unsigned long long)__shfl_sync(FULL, (uint32_t)(v >> 32), src) << 32) | __shfl_sync(FULL, (uint32_t)v, src);
}

__global__ void __launch_bounds__(CW_WARPS * 32) cluster_warp_kernel(
    unsigned long long* __restrict__ ctr, PostCaps C, const uint32_t* __restrict__ act_base, const uint32_t* __restrict__ act_count,
    unsigned long long* __restrict__ gkey, uint8_t* __restrict__ gkept, uint32_t max_diff, const uint32_t* __restrict__ thresh,
    const uint32_t* __restrict__ knode_base, uint32_t locus_key_base, uint32_t* __restrict__ cov_keys) {
    __shared__ uint32_t s_u32[CW_WARPS][6][64];  // first, last, size, pf, ord, ord2
    __shared__ uint32_t s_alive[CW_WARPS][64];
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n_active = ctr[CTR_ACTIVE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned long long a = (unsigned long long)blockIdx.x * CW_WARPS + wid; a < n_active; a += (unsigned long long)gridDim.x * CW_WARPS) {
        const uint32_t c = act_count[a];
        if (c > CLUSTER_WARP_MAX) continue;  // the CTA kernel's read
        const uint32_t base = act_base[a];
        // element i lives in lane i & 31, register i >> 5; padding sorts to the end
        unsigned long long x0 = (uint32_t)lane < c ? gkey[base + lane] : ~0ull;
        unsigned long long x1 = (uint32_t)lane + 32u < c ? gkey[base + 32 + lane] : ~0ull;
        const bool two = c > 32u;  // warp-uniform
        const int top = two ? 64 : 32;
        for (int k = 2; k <= top; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j == 32) {  // k == 64: the last merge is ascending over all 64 elements
                    if (x0 > x1) {
                        const unsigned long long t = x0;
                        x0 = x1;
                        x1 = t;
                    }
                } else {
                    const bool lower = (lane & j) == 0;
                    {
                        const unsigned long long y = shfl_xor64(x0, j);
                        const bool asc = (lane & k) == 0;
                        x0 = (lower == asc) ? (x0 < y ? x0 : y) : (x0 < y ? y : x0);
                    }
                    if (two) {
                        const unsigned long long y = shfl_xor64(x1, j);
                        const bool asc = ((lane + 32) & k) == 0;
                        x1 = (lower == asc) ? (x1 < y ? x1 : y) : (x1 < y ? y : x1);
                    }
                }
            }
        if ((uint32_t)lane < c) gkey[base + lane] = x0;  // the slice stays sorted for the hooks / later consumers
        if ((uint32_t)lane + 32u < c) gkey[base + 32 + lane] = x1;
        // ---- define_clusters: a hit opens a cluster when prg or strand change or the gap exceeds max_diff
        const unsigned long long p0 = shfl_up64(x0, 1), x0_31 = shfl64(x0, 31);
        unsigned long long p1 = shfl_up64(x1, 1);
        if (lane == 0) p1 = x0_31;
        auto opens = [&](unsigned long long cur, unsigned long long prev, uint32_t i) {
            if (i == 0u || i == c) return true;  // i == c: the first padding element closes the last cluster
            if (i > c) return false;
            return gk_pf(cur) != gk_pf(prev) || gk_start(cur) - gk_start(prev) > max_diff;
        };
        const uint32_t m0 = __ballot_sync(FULL, opens(x0, p0, (uint32_t)lane));
        const uint32_t m1 = __ballot_sync(FULL, opens(x1, p1, (uint32_t)lane + 32u));
        const unsigned long long S = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
        uint32_t cb[2], ce[2];
        bool pass[2], begin_pass[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t i = (uint32_t)lane + 32u * r;
            const unsigned long long le = i == 63u ? ~0ull : ((2ull << i) - 1ull);
            cb[r] = 63u - (uint32_t)__clzll((long long)(S & le));
            const unsigned long long gt = S & ~le;
            ce[r] = gt ? (uint32_t)__ffsll((long long)gt) - 1u : c;
            const unsigned long long xr = r ? x1 : x0;
            pass[r] = i < c && (ce[r] - cb[r]) > thresh[gk_prg(xr)];
            begin_pass[r] = pass[r] && cb[r] == i;
        }
        const uint32_t b0 = __ballot_sync(FULL, begin_pass[0]), b1 = __ballot_sync(FULL, begin_pass[1]);
        const unsigned long long BP = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
        const uint32_t ncl = (uint32_t)__popcll(BP);
        if (ncl == 0u) {
            if ((uint32_t)lane < c) gkept[base + lane] = 0;
            if ((uint32_t)lane + 32u < c) gkept[base + 32 + lane] = 0;
            continue;
        }
        bool alive[2] = {pass[0], pass[1]};
        if (ncl > 1u) {
            uint32_t* first = s_u32[wid][0];
            uint32_t* last = s_u32[wid][1];
            uint32_t* size = s_u32[wid][2];
            uint32_t* pf = s_u32[wid][3];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const unsigned long long xr = r ? x1 : x0;
                const uint32_t i = (uint32_t)lane + 32u * r;
                if (begin_pass[r]) {
                    const uint32_t t = (uint32_t)__popcll(BP & ((1ull << i) - 1ull));
                    first[t] = gk_start(xr);
                    size[t] = ce[r] - cb[r];
                    pf[t] = gk_pf(xr);
                    s_alive[wid][t] = 1u;
                }
                if (pass[r] && i + 1u == ce[r]) last[(uint32_t)__popcll(BP & ((1ull << cb[r]) - 1ull))] = gk_start(xr);
            }
            __syncwarp();
            ClusterArrays A{first, last, size, pf, s_u32[wid][4], s_u32[wid][5], s_alive[wid]};
            filter_clusters_warp(ncl, A, lane);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (pass[r]) alive[r] = s_alive[wid][(uint32_t)__popcll(BP & ((1ull << cb[r]) - 1ull))] != 0u;
        }
        // ---- add_clusters_to_pangraph: kept flags, one coverage key per kept hit, one locus key per kept cluster
        if ((uint32_t)lane < c) gkept[base + lane] = alive[0] ? 1 : 0;
        if ((uint32_t)lane + 32u < c) gkept[base + 32 + lane] = alive[1] ? 1 : 0;
        const uint32_t mine = (alive[0] ? 1u : 0u) + (alive[1] ? 1u : 0u) + ((alive[0] && begin_pass[0]) ? 1u : 0u) +
                              ((alive[1] && begin_pass[1]) ? 1u : 0u);
        uint32_t total;
        uint32_t o = warp_alloc(ctr + CTR_COVKEYS, mine, lane, total);
        const uint32_t nkept = __popc(__ballot_sync(FULL, alive[0])) + __popc(__ballot_sync(FULL, alive[1]));
        if (lane == 0 && nkept) atomicAdd(ctr + CTR_KEPT, (unsigned long long)nkept);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!alive[r]) continue;
            const unsigned long long xr = r ? x1 : x0;
            cov_keys[o++] = 2u * (knode_base[gk_prg(xr)] + gk_knode(xr)) + gk_rev(xr);
            if (begin_pass[r]) cov_keys[o++] = locus_key_base + gk_prg(xr);
        }
    }
}

// ============================================================================================
// one CTA per read with more than 64 hits (long reads)
// ============================================================================================
constexpr int CC_THREADS = 256;
constexpr uint32_t CC_SMEM_KEYS = 8192;  // 64 KB of keys sorted in shared memory; longer slices are sorted in place in global memory

// ascending-only bitonic network (first step of every merge compares i with i ^ (k-1), the rest i ^ j): elements past n
// act as +infinity and never move, so n need not be a power of two
template <class T>
__device__ void block_bitonic(T* a, uint32_t n, int tid, int nthreads) {
    uint32_t P = 2;
    while (P < n) P <<= 1;
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const uint32_t mask = (j == (k >> 1)) ? (k - 1u) : j;
            for (uint32_t i = tid; i < n; i += nthreads) {
                const uint32_t l = i ^ mask;
                if (l > i && l < n) {
                    const T x = a[i], y = a[l];
                    if (x > y) {
                        a[i] = y;
                        a[l] = x;
                    }
                }
            }
            __syncthreads();
        }
}

struct BigScratch {  // per-hit scratch, indexed like the grouped-hit array
    uint32_t *cbeg, *cend;                      // cluster begin of every hit; end, stored at the begin
    uint32_t *first, *last, *size, *pf, *ord, *ord2, *alive;  // per cluster that passed the threshold, at [slice base + t]
    uint32_t* slot;                             // cluster number of a begin, stored at the begin
};

__global__ void __launch_bounds__(CC_THREADS) cluster_cta_kernel(
    unsigned long long* __restrict__ ctr, PostCaps C, const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ act_base,
    const uint32_t* __restrict__ act_count, unsigned long long* __restrict__ gkey, uint8_t* __restrict__ gkept, uint32_t max_diff,
    const uint32_t* __restrict__ thresh, const uint32_t* __restrict__ knode_base, uint32_t locus_key_base,
    uint32_t* __restrict__ cov_keys, BigScratch B) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ uint32_t s_carry, s_ncl, s_warp_max[CC_THREADS / 32];
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n_big = ctr[CTR_BIG];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (unsigned long long bi = blockIdx.x; bi < n_big; bi += gridDim.x) {
        const uint32_t a = big_list[bi];
        const uint32_t c = act_count[a], base = act_base[a];
        unsigned long long* K = gkey + base;
        // ---- sort the slice
        if (c <= CC_SMEM_KEYS) {
            for (uint32_t i = tid; i < c; i += CC_THREADS) s_keys[i] = K[i];
            __syncthreads();
            block_bitonic(s_keys, c, tid, CC_THREADS);
            for (uint32_t i = tid; i < c; i += CC_THREADS) K[i] = s_keys[i];
        } else {
            block_bitonic(K, c, tid, CC_THREADS);
        }
        if (tid == 0) {
            s_carry = 0;
            s_ncl = 0;
        }
        __syncthreads();
        // ---- define_clusters: cbeg[i] = index of the last hit <= i that opens a cluster (running maximum)
        for (uint32_t i0 = 0; i0 < c; i0 += CC_THREADS) {
            const uint32_t i = i0 + tid;
            uint32_t v = 0;
            if (i < c && i > 0) {
                const unsigned long long cur = K[i], prev = K[i - 1];
                if (gk_pf(cur) != gk_pf(prev) || gk_start(cur) - gk_start(prev) > max_diff) v = i;
            }
            uint32_t m = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) m = max(m, __shfl_up_sync(FULL, m, d) * (lane >= d ? 1u : 0u));
            if (lane == 31) s_warp_max[wid] = m;
            __syncthreads();
            uint32_t pre = s_carry;
            for (int w = 0; w < wid; ++w) pre = max(pre, s_warp_max[w]);
            m = max(m, pre);
            if (i < c) B.cbeg[base + i] = m;
            __syncthreads();
            if (tid == CC_THREADS - 1) s_carry = m;
            __syncthreads();
        }
        // a hit that is the last of its cluster knows the end
        for (uint32_t i = tid; i < c; i += CC_THREADS) {
            const uint32_t b = B.cbeg[base + i];
            if (i + 1 == c || B.cbeg[base + i + 1] != b) B.cend[base + b] = i + 1;
        }
        __syncthreads();
        // ---- size threshold: the passing clusters form the list the filters work on
        for (uint32_t i = tid; i < c; i += CC_THREADS) {
            uint32_t sl = 0xffffffffu;
            if (B.cbeg[base + i] == i) {
                const uint32_t e = B.cend[base + i];
                const unsigned long long k0 = K[i];
                if (e - i > thresh[gk_prg(k0)]) {
                    sl = atomicAdd(&s_ncl, 1u);
                    B.first[base + sl] = gk_start(k0);
                    B.last[base + sl] = gk_start(K[e - 1]);
                    B.size[base + sl] = e - i;
                    B.pf[base + sl] = gk_pf(k0);
                    B.alive[base + sl] = 1u;
                }
                B.slot[base + i] = sl;
            }
        }
        __syncthreads();
        const uint32_t ncl = s_ncl;
        if (ncl > 1u && wid == 0) {
            ClusterArrays A{B.first + base, B.last + base, B.size + base, B.pf + base, B.ord + base, B.ord2 + base, B.alive + base};
            __threadfence_block();
            filter_clusters_warp(ncl, A, lane);
        }
        __syncthreads();
        // ---- kept flags + coverage keys
        for (uint32_t i0 = 0; i0 < c; i0 += CC_THREADS) {
            const uint32_t i = i0 + tid;
            bool keep = false, is_begin = false;
            unsigned long long k0 = 0;
            if (i < c) {
                const uint32_t b = B.cbeg[base + i];
                const uint32_t sl = B.slot[base + b];
                keep = sl != 0xffffffffu && B.alive[base + sl] != 0u;
                is_begin = keep && b == i;
                k0 = K[i];
                gkept[base + i] = keep ? 1 : 0;
            }
            const uint32_t mine = (keep ? 1u : 0u) + (is_begin ? 1u : 0u);
            uint32_t total;
            uint32_t o = warp_alloc(ctr + CTR_COVKEYS, mine, lane, total);
            const uint32_t nkept = __popc(__ballot_sync(FULL, keep));
            if (lane == 0 && nkept) atomicAdd(ctr + CTR_KEPT, (unsigned long long)nkept);
            if (keep) {
                cov_keys[o++] = 2u * (knode_base[gk_prg(k0)] + gk_knode(k0)) + gk_rev(k0);
                if (is_begin) cov_keys[o++] = locus_key_base + gk_prg(k0);
            }
        }
        __syncthreads();
    }
}

// ============================================================================================
// S5 : coverage by sorted-key reduction, no atomics on the counters
// ============================================================================================
constexpr int CT_THREADS = 512;
constexpr uint32_t CT_TILE = 8192;

__global__ void __launch_bounds__(CT_THREADS) cov_tile_kernel(const unsigned long long* __restrict__ ctr, PostCaps C,
                                                             const uint32_t* __restrict__ cov_keys, int32_t* __restrict__ partials,
                                                             uint32_t n_accum) {
    __shared__ uint32_t s_k[CT_TILE];
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n = ctr[CTR_COVKEYS];
    const unsigned long long n_tiles = (n + CT_TILE - 1) / CT_TILE;
    int32_t* mine = partials + (size_t)blockIdx.x * n_accum;
    const int tid = threadIdx.x;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long t0 = tile * CT_TILE;
        const uint32_t m = (uint32_t)min((unsigned long long)CT_TILE, n - t0);
        for (uint32_t i = tid; i < m; i += CT_THREADS) s_k[i] = cov_keys[t0 + i];
        __syncthreads();
        block_bitonic(s_k, m, tid, CT_THREADS);
        // the first element of every run of equal keys finds the run's end and adds its length: after the sort a key is
        // touched by exactly one thread of this CTA, and the partial belongs to this CTA alone
        for (uint32_t i = tid; i < m; i += CT_THREADS) {
            const uint32_t key = s_k[i];
            if (i > 0 && s_k[i - 1] == key) continue;
            uint32_t e = i + 1;
            while (e < m && e < i + 8u && s_k[e] == key) ++e;
            if (e < m && s_k[e] == key) {  // a long run (deep coverage): first index with s_k[idx] > key
                uint32_t lo = e, hi = m;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_k[mid] <= key) lo = mid + 1;
                    else hi = mid;
                }
                e = lo;
            }
            mine[key] += (int32_t)(e - i);
        }
        __syncthreads();
    }
}

// Column sum of the partials of the CTAs that had tiles into `dst` (re-zeroing them for the next batch).  dst is the
// sample's accumulator on this GPU, or — remote != 0 — the ROOT GPU's accumulator mapped over NVLink, which receives one
// fire-and-forget red.global.add per non-zero counter: the fused form of the path's only collective.
__global__ void cov_merge_kernel(const unsigned long long* __restrict__ ctr, PostCaps C, int32_t* __restrict__ partials,
                                 uint32_t n_partials, uint32_t n_accum, uint32_t n_keys, int32_t* __restrict__ dst, int remote) {
    if (batch_overflowed(ctr, C)) return;
    const unsigned long long n_tiles = (ctr[CTR_COVKEYS] + CT_TILE - 1) / CT_TILE;
    const uint32_t used = (uint32_t)min((unsigned long long)n_partials, n_tiles);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_keys; k += gridDim.x * blockDim.x) {
        int32_t sum = 0;
        for (uint32_t g = 0; g < used; ++g) {
            int32_t* p = partials + (size_t)g * n_accum + k;
            const int32_t v = *p;
            if (v) {
                sum += v;
                *p = 0;
            }
        }
        if (!sum) continue;
        if (remote) asm volatile("red.global.add.s32 [%0], %1;" ::"l"(dst + k), "r"(sum) : "memory");
        else dst[k] += sum;
    }
}

uint32_t cov_partial_ctas(int sm_count) { return (uint32_t)sm_count; }

void launch_postprocess(const PostBuffers& P, PostCaps C, uint32_t read_id_base, uint64_t n_reads, uint32_t max_diff,
                        const uint32_t* d_thresh_per_prg, const uint32_t* d_knode_base, uint32_t total_knodes, uint32_t n_loci,
                        int32_t* d_accum_dst, int remote_dst, int sm_count, cudaStream_t st, cudaEvent_t ev_grouped,
                        cudaEvent_t ev_clustered) {
    (void)n_reads;
    const unsigned sm = (unsigned)sm_count;
    count_reads_kernel<<<sm * 4, 256, 0, st>>>(P.hi, P.ctr, C, read_id_base, P.read_count, P.act_read);
    assign_kernel<<<sm, 256, 0, st>>>(P.ctr, C, P.read_count, P.act_read, P.act_base, P.act_count, P.read_base, P.big_list);
    scatter_kernel<<<sm * 4, 256, 0, st>>>(P.hi, P.lo, P.ctr, C, read_id_base, P.read_count, P.read_base, P.gkey);
    if (ev_grouped) cudaEventRecord(ev_grouped, st);
    const uint32_t locus_key_base = 2u * total_knodes;
    cluster_warp_kernel<<<sm * 4, CW_WARPS * 32, 0, st>>>(P.ctr, C, P.act_base, P.act_count, P.gkey, P.gkept, max_diff, d_thresh_per_prg,
                                                         d_knode_base, locus_key_base, P.cov_keys);
    ensure_dyn_smem(cluster_cta_kernel, (size_t)CC_SMEM_KEYS * 8);
    BigScratch B{P.scratch[0], P.scratch[1], P.scratch[2], P.scratch[3], P.scratch[4], P.scratch[5], P.scratch[6], P.scratch[7],
                 P.scratch[8], P.scratch[9]};
    cluster_cta_kernel<<<sm * 2, CC_THREADS, (size_t)CC_SMEM_KEYS * 8, st>>>(P.ctr, C, P.big_list, P.act_base, P.act_count, P.gkey,
                                                                            P.gkept, max_diff, d_thresh_per_prg, d_knode_base,
                                                                            locus_key_base, P.cov_keys, B);
    if (ev_clustered) cudaEventRecord(ev_clustered, st);
    const uint32_t n_accum = 2u * total_knodes + n_loci + 4u, n_keys = 2u * total_knodes + n_loci;
    cov_tile_kernel<<<P.n_partials, CT_THREADS, 0, st>>>(P.ctr, C, P.cov_keys, P.partials, n_accum);
    cov_merge_kernel<<<(n_keys + 255) / 256, 256, 0, st>>>(P.ctr, C, P.partials, P.n_partials, n_accum, n_keys, d_accum_dst, remote_dst);
    g_launches += 7;
}

}  // namespace drprg
