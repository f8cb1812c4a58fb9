// Hand-written sm_100a kernels for the map hot path: read sketching + index lookup (S1+S2), hit
// clustering (S3/S4), k-mer coverage (S5), ML path (S7) and genotyping (S8).  These replace the
// per-read and per-locus loops of `pandora map` that drprg launches at
// /root/reference/src/lib.rs:580-642 (argv :594-609, src/predict.rs:288-294); stage semantics
// follow pandora's Seq::minimizer_sketch, add_read_hits, define_clusters, filter_clusters(2),
// add_hits_to_kmergraphs, KmerGraphWithCoverage::find_max_path and SampleInfo.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cub/cub.cuh>

#include "kernels.cuh"

namespace drprg {

static uint64_t g_launches = 0;
uint64_t launch_count() { return g_launches; }

#define FULL 0xffffffffu

// ============================================================================================
// S1 + S2 : sketch + lookup.  One warp per read; lanes own consecutive k-mer positions.
//   * bases are 2-bit packed, first base in the top bits, so the forward k-mer at position p is a
//     funnel shift of two words and the reverse complement is brev + pair swap of its complement;
//   * k-mers are kept LEFT-ALIGNED in 32 bits (value << (32-2k)): every "& mask" of pandora's
//     hash64 becomes the natural 2^32 wrap, the three shift-add steps become single IMADs
//     (x2097151, x265, x21) and hash order is preserved, so the canonical min works in place;
//   * window minima with all ties (pandora keeps every k-mer attaining a window minimum) are a
//     sliding min followed by a sliding max of the minima, both by doubling in shared memory:
//     position i is a minimizer  <=>  h[i] == max over windows s containing i of min(h[s..s+w)).
// ============================================================================================
constexpr int WARPS = 8;
constexpr int EXT_MAX = CHUNK + 2 * (W_MAX - 1);
constexpr int BUF_N = EXT_MAX + W_MAX + 2;
constexpr int SW_N = (EXT_MAX + K_MAX + 15) / 16 + 3;

__device__ __forceinline__ uint32_t hash_left_aligned(uint32_t K, uint32_t S, uint32_t hm) {
    K = K * 2097151u - (1u << S);  // (~key + (key << 21)) & mask
    K ^= (K >> 24) & hm;           // key ^= key >> 24
    K *= 265u;                     // (key + (key << 3) + (key << 8)) & mask
    K ^= (K >> 14) & hm;
    K *= 21u;                      // (key + (key << 2) + (key << 4)) & mask
    K ^= (K >> 28) & hm;
    K += K << 31;                  // (key + (key << 31)) & mask : only bit 31 can change (k = 16)
    return K;
}

// hash of the k-mer held RIGHT-aligned (possibly with garbage above bit 2k): the left alignment (<< S) is folded
// into the first multiply
template <uint32_t S>
__device__ __forceinline__ uint32_t hash_right_aligned(uint32_t F) {
    constexpr uint32_t hm = ~((1u << S) - 1u);
    uint32_t K = F * (2097151u << S) - (1u << S);
    K ^= (K >> 24) & hm;
    K *= 265u;
    K ^= (K >> 14) & hm;
    K *= 21u;
    K ^= (K >> 28) & hm;
    return K;  // S >= 1: the final (key + (key << 31)) & mask step cannot change a kept bit
}

__device__ __forceinline__ uint32_t table_slot(uint32_t h, uint32_t bits) { return (h * 0x9E3779B1u) >> (32 - bits); }

template <bool LOOKUP>
__global__ void __launch_bounds__(WARPS * 32) sketch_kernel(DevReads R, DevTable T, uint32_t w, uint32_t k,
                                                           unsigned long long* __restrict__ out_a,
                                                           unsigned long long* __restrict__ out_b,
                                                           unsigned long long* __restrict__ out_count,
                                                           unsigned long long cap) {
    __shared__ uint32_t s_words[WARPS][SW_N];
    __shared__ uint32_t s_h[WARPS][BUF_N];
    __shared__ uint32_t s_a[WARPS][BUF_N];
    __shared__ uint32_t s_b[WARPS][BUF_N];
    __shared__ uint32_t s_strand[WARPS][(EXT_MAX + 31) / 32 + 1];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* sw = s_words[wid];
    uint32_t* H = s_h[wid];
    uint32_t* A = s_a[wid];
    uint32_t* B = s_b[wid];
    uint32_t* SS = s_strand[wid];
    const uint32_t S = 32 - 2 * k;
    const uint32_t hm = (S == 0) ? 0xffffffffu : ~((1u << S) - 1u);
    const unsigned long long nwarps = (unsigned long long)gridDim.x * WARPS;

    for (unsigned long long r = (unsigned long long)blockIdx.x * WARPS + wid; r < R.n_reads; r += nwarps) {
        const uint32_t len = R.lens[r];
        if (len + 1 < w + k) continue;  // too short, or flagged 0 (non-ACGT): contributes nothing
        const uint32_t nk = len - k + 1;
        const unsigned long long wbase = R.stride_words ? r * R.stride_words : R.word_off[r];
        const uint32_t nwords_read = (len + 15) >> 4;

        for (uint32_t c0 = 0; c0 < nk; c0 += CHUNK) {
            const uint32_t ext_lo = (c0 >= w - 1) ? c0 - (w - 1) : 0;
            const uint32_t ext_hi = min(c0 + CHUNK + (w - 1), nk);
            const uint32_t n_ext = ext_hi - ext_lo;
            const uint32_t w0 = ext_lo >> 4;
            const uint32_t nw = ((ext_hi + k - 2) >> 4) - w0 + 1;
            __syncwarp();
            for (uint32_t i = lane; i < nw + 1; i += 32) {
                uint32_t wi = w0 + i;
                sw[i] = (wi < nwords_read) ? __ldg(R.words + wbase + wi) : 0u;
            }
            __syncwarp();
            // ---- canonical hashes of positions [ext_lo, ext_hi)
            for (uint32_t e0 = 0; e0 < n_ext; e0 += 32) {
                const uint32_t e = e0 + lane;
                const uint32_t b = 2u * (ext_lo + e - (w0 << 4));
                const uint32_t wi = min(b >> 5, nw - 1);
                const uint32_t v = __funnelshift_l(sw[wi + 1], sw[wi], b & 31u);
                const uint32_t F = v & hm;
                uint32_t y = __brev(~v & hm);
                y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
                const uint32_t Rc = y << S;
                const uint32_t hf = hash_left_aligned(F, S, hm), hr = hash_left_aligned(Rc, S, hm);
                const bool valid = e < n_ext;
                if (valid) H[e] = min(hf, hr);
                const uint32_t bal = __ballot_sync(FULL, valid && hf <= hr);
                if (lane == 0) SS[e0 >> 5] = bal;
            }
            __syncwarp();
            // ---- sliding minimum over w consecutive hashes: wm[e] = min(H[e .. e+w-1])
            uint32_t span = 1;
            const uint32_t* src = H;
            uint32_t* dst = A;
            while (span * 2 <= w) {
                const uint32_t cnt = n_ext - (2 * span - 1);
                for (uint32_t e = lane; e < cnt; e += 32) dst[e] = min(src[e], src[e + span]);
                __syncwarp();
                src = dst;
                dst = (dst == A) ? B : A;
                span *= 2;
            }
            // padded array P[t], t in [0, n_ext + w - 1): windows start at ext_lo - (w-1) + t
            const uint32_t n_win = n_ext - w + 1;
            const uint32_t n_pad = n_ext + w - 1;
            for (uint32_t t = lane; t < n_pad; t += 32) {
                uint32_t val = 0;
                if (t >= w - 1 && t - (w - 1) < n_win) {
                    const uint32_t e = t - (w - 1);
                    val = min(src[e], src[e + w - span]);
                }
                dst[t] = val;
            }
            __syncwarp();
            // ---- sliding maximum of the window minima: X[e] = max(P[e .. e+w-1])
            src = dst;
            dst = (dst == A) ? B : A;
            span = 1;
            while (span * 2 <= w) {
                const uint32_t cnt = n_pad - (2 * span - 1);
                for (uint32_t t = lane; t < cnt; t += 32) dst[t] = max(src[t], src[t + span]);
                __syncwarp();
                src = dst;
                dst = (dst == A) ? B : A;
                span *= 2;
            }
            // ---- minimizers of this chunk
            const uint32_t chunk_hi = min(c0 + CHUNK, nk);
            for (uint32_t p0 = c0; p0 < chunk_hi; p0 += 32) {
                const uint32_t p = p0 + lane;
                const uint32_t e = p - ext_lo;
                bool is_min = false;
                uint32_t hv = 0;
                if (p < chunk_hi) {
                    const uint32_t x = max(src[e], src[e + w - span]);
                    hv = H[e];
                    is_min = (hv == x);
                    hv >>= S;
                }
                const uint32_t read_strand = (SS[e >> 5] >> (e & 31)) & 1u;
                if (!LOOKUP) {
                    const uint32_t bal = __ballot_sync(FULL, is_min);
                    if (bal) {
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(out_count, (unsigned long long)__popc(bal));
                        base = __shfl_sync(FULL, base, 0);
                        if (is_min) {
                            const unsigned long long o = base + __popc(bal & ((1u << lane) - 1u));
                            if (o < cap) {
                                out_a[o] = ((unsigned long long)(R.read_id_base + (uint32_t)r) << 32) | p;
                                out_b[o] = ((unsigned long long)hv << 1) | read_strand;
                            }
                        }
                    }
                } else {
                    bool pass = false;
                    if (is_min) {
                        const uint32_t fw = __ldg(T.filter + (hv & ((1u << T.filter_bits) - 1u)));
                        const uint32_t m = (1u << ((hv >> T.filter_bits) & 31u)) | (1u << ((hv >> (T.filter_bits + 5)) & 31u));
                        pass = (fw & m) == m;
                    }
                    if (__any_sync(FULL, pass)) {
                        uint32_t rec_begin = 0, rec_n = 0;
                        if (pass) {
                            uint32_t slot = table_slot(hv, T.slot_bits);
                            const uint32_t smask = (1u << T.slot_bits) - 1u;
                            while (true) {
                                const uint2 ent = __ldg(T.slots + slot);
                                if (ent.y == 0u) break;
                                if (ent.x == hv) {
                                    rec_begin = ent.y & 0xffffffu;
                                    rec_n = ent.y >> 24;
                                    break;
                                }
                                slot = (slot + 1) & smask;
                            }
                        }
                        // warp-aggregated append
                        uint32_t incl = rec_n;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t t = __shfl_up_sync(FULL, incl, d);
                            if (lane >= d) incl += t;
                        }
                        const uint32_t total = __shfl_sync(FULL, incl, 31);
                        if (total) {
                            unsigned long long base = 0;
                            if (lane == 0) base = atomicAdd(out_count, (unsigned long long)total);
                            base = __shfl_sync(FULL, base, 0) + (incl - rec_n);
                            for (uint32_t j = 0; j < rec_n; ++j) {
                                const uint2 rc = __ldg(T.recs + rec_begin + j);
                                const uint32_t fwd = ((rc.y & 1u) == read_strand) ? 1u : 0u;
                                if (base + j < cap) {
                                    out_a[base + j] = ((unsigned long long)(R.read_id_base + (uint32_t)r) << 32) |
                                                      ((unsigned long long)(rc.y >> 1) << 16) | ((unsigned long long)(fwd ^ 1u) << 15);
                                    out_b[base + j] = ((unsigned long long)p << 32) | rc.x;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

// ============================================================================================
// S1 + S2, short reads (Illumina): ONE THREAD PER READ, W and K compile-time.
// The warp-per-read kernel above spends ~90 % of its issue slots on the shared-memory min/max passes
// and partial rounds (ncu: 1236 warp instructions per 150 bp read, hashing only 8 % of them).  Here all
// 32 lanes of a warp walk 32 different reads in lockstep, everything lives in registers and the
// per-position cost is ~50 instructions:
//   * rolling k-mers: forward by one funnel shift taking the next base from the top of the current
//     word, reverse complement by one funnel shift taking the complemented base from a rotating copy;
//   * window minima with ties by the van Herk / Gil-Werman block decomposition with block = W and the
//     loop unrolled by W so every array index is static: prefix/suffix minima give the minimum of each
//     window, prefix/suffix maxima of those give, per position, the largest window minimum among the
//     windows containing it; position i is a minimizer iff h[i] equals that value;
//   * positions past the read end (and windows before its start) carry hash 0 == "-infinity": a window
//     touching them has minimum 0 and so can never certify a real minimizer (a real hash of 0 is the
//     minimum of its valid windows anyway), which removes every boundary branch;
//   * the block's hashes are parked in shared memory ([slot][thread], conflict free) only so that the
//     rare flagged positions can be fetched with a dynamic index when they probe the index.
// ============================================================================================
constexpr int SHORT_THREADS = 512;
constexpr uint32_t SMEM_FILTER_BITS = 14;  // a pre-filter of <= 2^14 words (64 KB) is copied into shared memory
#ifndef DRPRG_DEFAULT_VARIANT
#define DRPRG_DEFAULT_VARIANT 3
#endif

template <int W, int K, bool LOOKUP, int VARIANT, bool SMEM_FILTER, int THREADS>
__global__ void __launch_bounds__(THREADS) sketch_short_kernel(DevReads R, DevTable T,
                                                                     unsigned long long* __restrict__ out_a,
                                                                     unsigned long long* __restrict__ out_b,
                                                                     unsigned long long* __restrict__ out_count,
                                                                     unsigned long long cap) {
    static_assert(K >= 2 && K <= 15, "left-aligned hash with a spare low bit range needs k <= 15");
    constexpr uint32_t S = 32 - 2 * K;
    constexpr uint32_t HM = ~((1u << S) - 1u);
    extern __shared__ uint32_t s_short[];
    uint32_t(*s_h)[W][THREADS] = reinterpret_cast<uint32_t(*)[W][THREADS]>(s_short);
    const uint32_t* s_filter = s_short + 2 * W * THREADS;
    const int tid = threadIdx.x;
    if (LOOKUP && SMEM_FILTER) {  // "hot buckets in shared memory": the whole negative filter, once per persistent CTA
        uint32_t* f = s_short + 2 * W * THREADS;
        const uint32_t nwf = 1u << T.filter_bits;
        for (uint32_t i = tid * 4; i < nwf; i += THREADS * 4)
            *reinterpret_cast<uint4*>(f + i) = __ldg(reinterpret_cast<const uint4*>(T.filter + i));
        __syncthreads();
    }
    // A work item is a whole read, or — for long reads — a SEGMENT of one (R.seg_read != nullptr): seg_len k-mer
    // positions whose minimizer status only depends on the w-1 positions either side, so a thread streams
    // [seg_start-(w-1), seg_end+(w-1)) and reports [seg_start, seg_end).  Padding with hash 0 outside the streamed
    // range is exact at the true read ends and harmless inside the read (it only affects the halo).
    const unsigned long long n_items = R.seg_read ? R.n_segs : R.n_reads;
    const unsigned long long n_tiles = (n_items + THREADS - 1) / THREADS;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long item = tile * THREADS + tid;
    const bool have = item < n_items;
    const unsigned long long r = have ? (R.seg_read ? (unsigned long long)__ldg(R.seg_read + item) : item) : 0ull;
    uint32_t len = have ? __ldg(R.lens + r) : 0u;
    if (len + 1 < (uint32_t)(W + K)) len = 0;  // too short or dropped: no k-mer position is valid
    const uint32_t nk_read = len ? len - K + 1 : 0;
    const uint32_t seg_s = (have && R.seg_read) ? __ldg(R.seg_start + item) : 0u;
    const uint32_t seg_e = R.seg_read ? min(seg_s + R.seg_len, nk_read) : nk_read;   // report [seg_s, seg_e)
    const uint32_t str_lo = seg_s >= (uint32_t)(W - 1) ? seg_s - (W - 1) : 0u;        // stream [str_lo, str_lo + nk)
    const uint32_t nk = nk_read ? min(nk_read, seg_e + (W - 1)) - str_lo : 0u;
    const uint32_t nk_max = __reduce_max_sync(FULL, nk);
    if (nk_max == 0) continue;
    const uint32_t nk_min = __reduce_min_sync(FULL, nk);
    const uint32_t* wp = R.words + (have ? (R.stride_words ? r * R.stride_words : __ldg(R.word_off + r)) : 0ull) + (str_lo >> 4);
    const uint32_t nwords = ((len + 15) >> 4) - (nk_read ? (str_lo >> 4) : 0u);

    // Bases are served from a 64-bit shift register (hi:lo) holding `avail` bases, top aligned; it is
    // topped up with the next 16-base word once per block of W positions (W <= 16 bases are consumed per
    // block), so the refill test is per block, not per base, and warp-uniform: all lanes are in lockstep.
    static_assert(W <= 16, "one refill per block must cover the block");
    uint32_t hi = 0, lo = 0, avail = 0, widx = 0, F = 0, Rc = 0;
    uint32_t wnext = nwords ? __ldg(wp) : 0u;  // always one word ahead: the load has a whole block to land
    auto refill = [&]() {
        if (avail <= 16u) {
            const uint32_t word = wnext;
            ++widx;
            wnext = (widx < nwords) ? __ldg(wp + widx) : 0u;
            const uint32_t t = 2u * avail;  // 0..32 bits already occupied at the top of hi; lo is empty
            hi |= __funnelshift_rc(word, 0u, t);
            lo = __funnelshift_rc(0u, word, t);
            avail += 16u;
        }
    };
    auto next_base = [&]() {
        const uint32_t c = hi >> 30;
        hi = __funnelshift_l(lo, hi, 2);
        lo <<= 2;
        F = F * 4u + c;                                     // garbage above bit 2K wraps away in the first hash multiply
        Rc = (__funnelshift_r(Rc, c, 2) & HM) ^ 0xC0000000u;  // complemented base enters at the top; bases older than K fall off
    };
    {
        // skip to the first streamed base inside its word, then prime the first K-1 bases
        const uint32_t prime = (str_lo & 15u) + (uint32_t)(K - 1);
#pragma unroll 1
        for (uint32_t i = 0; i < prime; ++i) {
            refill();
            next_base();
            --avail;
        }
    }

    uint32_t hp[W], Sp[W], SXo[W + 1];
#pragma unroll
    for (int j = 0; j < W; ++j) hp[j] = Sp[j] = SXo[j] = 0u;
    SXo[W] = 0u;
    uint32_t strand_prev = 0;
    const uint32_t n_blocks = (nk_max + W - 1) / W + 1;  // one extra all-padding block resolves the last real one
#pragma unroll 1
    for (uint32_t b = 0; b < n_blocks; ++b) {
        uint32_t h[W];
        uint32_t not_strand = 0;  // bit (W-1-j) = !(hf <= hr) of position j
        const uint32_t p0 = b * W;
        uint32_t* sh = &s_h[b & 1][0][tid];
        refill();
        avail -= (uint32_t)W;
        // blocks that lie inside every lane's read (all but the last one or two) may skip the per-position padding select
        auto strand_bit = [&](uint32_t hf, uint32_t hr, int j) {
            if (VARIANT & 1)
                asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(not_strand) : "r"(hf), "r"(hr), "r"(1u << (W - 1 - j)));
            else
                not_strand |= (hf > hr ? 1u : 0u) << (W - 1 - j);
        };
        if ((VARIANT & 2) && p0 + W <= nk_min) {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                next_base();
                const uint32_t hf = hash_right_aligned<S>(F), hr = hash_left_aligned(Rc, S, HM);
                const uint32_t hv = min(hf, hr);
                strand_bit(hf, hr, j);
                h[j] = hv;
                sh[j * THREADS] = hv;
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j) {
                next_base();
                const uint32_t hf = hash_right_aligned<S>(F), hr = hash_left_aligned(Rc, S, HM);
                uint32_t hv = min(hf, hr);
                strand_bit(hf, hr, j);
                hv = (p0 + j < nk) ? hv : 0u;
                h[j] = hv;
                sh[j * THREADS] = hv;
            }
        }
        const uint32_t strand_cur = ~not_strand;
        // windows starting in the previous block: offset t covers prev[t..W-1] + cur[0..t-1]
        uint32_t wm[W];
        wm[0] = Sp[0];
        {
            uint32_t pmin = h[0];
#pragma unroll
            for (int t = 1; t < W; ++t) {
                wm[t] = min(Sp[t], pmin);
                pmin = min(pmin, h[t]);
            }
        }
        // previous block's positions: best window minimum among the windows containing them
        uint32_t flags = 0;
        {
            uint32_t pmax = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                pmax = max(pmax, wm[j]);
                const uint32_t best = max(pmax, SXo[j + 1]);
                if (VARIANT & 1)
                    asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(flags) : "r"(hp[j]), "r"(best), "r"(1u << j));
                else
                    flags |= (hp[j] == best ? 1u : 0u) << j;
            }
        }
        if (b > 0) {
            const uint32_t prev0 = str_lo + p0 - W;  // read coordinate of the previous block's first position
            // positions to report: prev0 + j in [seg_s, seg_e)
            const uint32_t first = seg_s > prev0 ? min(seg_s - prev0, (uint32_t)W) : 0u;
            const uint32_t last = seg_e > prev0 ? min(seg_e - prev0, (uint32_t)W) : 0u;
            uint32_t fm = flags & ((1u << last) - 1u) & ~((1u << first) - 1u);
            while (fm) {
                const int j = __ffs(fm) - 1;
                fm &= fm - 1;
                const uint32_t hv = s_h[(b - 1) & 1][j][tid] >> S;
                const uint32_t pos = prev0 + j;
                const uint32_t read_strand = (strand_prev >> (W - 1 - j)) & 1u;
                if (!LOOKUP) {
                    const unsigned long long o = atomicAdd(out_count, 1ull);
                    if (o < cap) {
                        out_a[o] = ((unsigned long long)(R.read_id_base + (uint32_t)r) << 32) | pos;
                        out_b[o] = ((unsigned long long)hv << 1) | read_strand;
                    }
                } else {
                    const uint32_t fidx = hv & ((1u << T.filter_bits) - 1u);
                    const uint32_t fw = SMEM_FILTER ? s_filter[fidx] : __ldg(T.filter + fidx);
                    // both filter bits set?  (funnel shifts take the shift amount modulo 32)
                    const uint32_t hb = hv >> T.filter_bits;
                    if (!(__funnelshift_r(fw, 0u, hb) & __funnelshift_r(fw, 0u, hb >> 5) & 1u)) continue;
                    uint32_t slot = table_slot(hv, T.slot_bits);
                    const uint32_t smask = (1u << T.slot_bits) - 1u;
                    uint32_t rec_begin = 0, rec_n = 0;
                    while (true) {
                        const uint2 ent = __ldg(T.slots + slot);
                        if (ent.y == 0u) break;
                        if (ent.x == hv) {
                            rec_begin = ent.y & 0xffffffu;
                            rec_n = ent.y >> 24;
                            break;
                        }
                        slot = (slot + 1) & smask;
                    }
                    if (rec_n) {
                        const unsigned long long base = atomicAdd(out_count, (unsigned long long)rec_n);
                        for (uint32_t q = 0; q < rec_n; ++q) {
                            const uint2 rc = __ldg(T.recs + rec_begin + q);
                            const uint32_t fwd = ((rc.y & 1u) == read_strand) ? 1u : 0u;
                            if (base + q < cap) {
                                out_a[base + q] = ((unsigned long long)(R.read_id_base + (uint32_t)r) << 32) |
                                                  ((unsigned long long)(rc.y >> 1) << 16) | ((unsigned long long)(fwd ^ 1u) << 15);
                                out_b[base + q] = ((unsigned long long)pos << 32) | rc.x;
                            }
                        }
                    }
                }
            }
        }
        // roll the block state
        {
            uint32_t smax = 0;
#pragma unroll
            for (int j = W - 1; j >= 0; --j) {
                smax = max(smax, wm[j]);
                SXo[j] = smax;
            }
            uint32_t smin = 0xffffffffu;
#pragma unroll
            for (int j = W - 1; j >= 0; --j) {
                smin = min(smin, h[j]);
                Sp[j] = smin;
                hp[j] = h[j];
            }
        }
        strand_prev = strand_cur;
    }
    }  // persistent tile loop
}

template <int W, int K, bool LOOKUP, int V, bool SF>
static void launch_short_one(const DevReads& R, const DevTable& T, unsigned long long* a, unsigned long long* b,
                             unsigned long long* cnt, uint64_t cap, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)2 * W * SHORT_THREADS * 4 + (SF ? (size_t)4 << T.filter_bits : 0);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(sketch_short_kernel<W, K, LOOKUP, V, SF, SHORT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * W * SHORT_THREADS * 4 + (SF ? (4u << SMEM_FILTER_BITS) : 0)));
        configured = true;
    }
    // persistent CTAs: two per SM (register file: 2 x 512 threads x 62 registers), each loops over read tiles
    const unsigned long long n_items = R.seg_read ? R.n_segs : R.n_reads;
    const unsigned long long n_tiles = (n_items + SHORT_THREADS - 1) / SHORT_THREADS;
    const unsigned grid = (unsigned)std::min<unsigned long long>(n_tiles, 2ull * (unsigned)sm_count);
    sketch_short_kernel<W, K, LOOKUP, V, SF, SHORT_THREADS><<<grid, SHORT_THREADS, smem, st>>>(R, T, a, b, cnt, cap);
    ++g_launches;
}

template <bool LOOKUP>
static bool launch_short(const DevReads& R, const DevTable& T, uint32_t w, uint32_t k, unsigned long long* a,
                         unsigned long long* b, unsigned long long* cnt, uint64_t cap, int sm_count, cudaStream_t st) {
    static const int variant = [] {
        const char* e = getenv("DRPRG_SKETCH_VARIANT");  // tuning switch: bit0 = predicated-OR flag accumulation, bit1 = padding-free fast path
        return e ? atoi(e) & 3 : DRPRG_DEFAULT_VARIANT;
    }();
    const bool sf = LOOKUP && T.filter_bits <= SMEM_FILTER_BITS;
#define DRPRG_SHORT_V(WW, KK, VV)                                                                 \
    if (variant == VV) {                                                                          \
        if (sf) launch_short_one<WW, KK, LOOKUP, VV, true>(R, T, a, b, cnt, cap, sm_count, st);   \
        else launch_short_one<WW, KK, LOOKUP, VV, false>(R, T, a, b, cnt, cap, sm_count, st);     \
        return true;                                                                              \
    }
#define DRPRG_SHORT(WW, KK)      \
    if (w == WW && k == KK) {    \
        DRPRG_SHORT_V(WW, KK, 0) \
        DRPRG_SHORT_V(WW, KK, 1) \
        DRPRG_SHORT_V(WW, KK, 2) \
        DRPRG_SHORT_V(WW, KK, 3) \
    }
    DRPRG_SHORT(11, 15)  // drprg defaults (src/builder.rs:40-41)
    DRPRG_SHORT(14, 15)  // pandora's default w, used by the reference's build tests (src/builder.rs:1181)
#undef DRPRG_SHORT
#undef DRPRG_SHORT_V
    return false;
}

// ============================================================================================
// K-mer screen.  ~99 % of whole-genome reads share no k-mer with the panel, yet the sketch above spends
// ~70 instructions per k-mer position on them (two hashes + window minima).  pandora's hash64 is a
// bijection on 2k-bit values, so "this minimizer is in the index" implies "this FORWARD k-mer of the read is
// one of the indexed k-mers or their reverse complements" — a set-membership test on the raw 2-bit k-mer
// that needs no hash and no window logic.
//   screen_kernel   streams every read once (a warp takes 32 reads at a time from a global ticket counter,
//                   words in registers) and tests each k-mer against a blocked 2-bit Bloom filter held in
//                   shared memory: one multiply, one LDS, two shifts, ~10 instructions per position.  The
//                   flagged positions (~1 % false positives + the real ones) are appended to a queue.
//   resolve_kernel  one thread per queued (read, position): canonical hash of that k-mer, index probe (exact,
//                   drops the false positives), then the minimizer test restricted to that position — it is a
//                   (w,k)-minimizer with pandora's "all ties kept" rule iff the run of neighbours whose hash
//                   is >= its own covers a whole window — and the hit records.
// The hits are the same set the full sketch + lookup kernels emit; only ~2 % of the positions ever get hashed.
// ============================================================================================
constexpr int SCREEN_THREADS = 1024;
constexpr uint32_t SCREEN_MUL = 0x9E3779B1u;

void screen_filter_insert(uint32_t* filter, uint32_t n_words, uint32_t kmer, uint32_t k) {
    const uint32_t p = kmer * (SCREEN_MUL << (32u - 2u * k));  // bits above 2k wrap away
    const uint32_t idx = (uint32_t)(((unsigned long long)p * n_words) >> 32);
    filter[idx] |= (1u << (kmer & 31u)) | (1u << ((p >> 11) & 31u));
}

// V: pipe-balance variants (the ALU pipe — SHF/LOP3/LEA — binds first, the multiplier pipe has room):
//   bit 0: shared-memory address by IMAD with a run-time 4 instead of LEA;  bit 1: the second bit index by
//   multiply-high with a run-time 2^21 instead of a shift.
struct ScreenConsts { uint32_t four, two21, prefetch; };
template <int K, int CW, int V>
__global__ void __launch_bounds__(SCREEN_THREADS, 1) screen_kernel(DevReads R, DevTable T, uint32_t wk, ScreenConsts SC,
                                                                   unsigned long long* __restrict__ queue,
                                                                   uint32_t* __restrict__ queue_kmer,
                                                                   unsigned long long* __restrict__ queue_count,
                                                                   unsigned long long queue_cap,
                                                                   unsigned long long* __restrict__ ticket) {
    static_assert(K >= 8 && K <= 15, "the filter bit choice needs >= 16 k-mer bits; 2k < 32");
    constexpr int NF = (CW + 1) / 2;
    extern __shared__ uint32_t s_kf[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t i = tid * 4; i < T.kfilter_words; i += SCREEN_THREADS * 4)
        *reinterpret_cast<uint4*>(s_kf + i) = __ldg(reinterpret_cast<const uint4*>(T.kfilter + i));
    __syncthreads();
    const uint32_t n_fw = T.kfilter_words;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_kf);
    const unsigned long long n_items = R.seg_read ? R.n_segs : R.n_reads;
    const unsigned long long n_tiles = (n_items + 31) / 32;
    const bool wide = R.stride_words && !(R.stride_words & 1u) && !R.seg_read &&
                      (reinterpret_cast<unsigned long long>(R.words) & 7ull) == 0ull;  // every read starts 8-byte aligned
    // tickets are drawn two tiles ahead: the next tile's id is known when a tile starts, so its words can be
    // prefetched into L2 while this tile is screened (a cold read otherwise costs the full HBM latency per tile)
    unsigned long long tile = 0, next_tile = 0;
    if (lane == 0) {
        tile = atomicAdd(ticket, 1ull);
        next_tile = atomicAdd(ticket, 1ull);
    }
    tile = __shfl_sync(FULL, tile, 0);
    next_tile = __shfl_sync(FULL, next_tile, 0);
    while (tile < n_tiles) {
        unsigned long long after_next = 0;
        if (lane == 0) after_next = atomicAdd(ticket, 1ull);  // lands while this tile is screened
        if (SC.prefetch && !R.seg_read && R.stride_words && next_tile * 32 + lane < n_items) {
            const uint32_t* np = R.words + (next_tile * 32 + lane) * R.stride_words;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(np));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(np + R.stride_words - 1));
            if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(R.lens + next_tile * 32));
        }
        const unsigned long long item = tile * 32 + lane;
        const bool have = item < n_items;
        const unsigned long long r = have ? (R.seg_read ? (unsigned long long)__ldg(R.seg_read + item) : item) : 0ull;
        uint32_t len = have ? __ldg(R.lens + r) : 0u;
        if (len + 1 < wk) len = 0;  // too short or dropped: the sketch skips it too
        const uint32_t nk_read = len ? len - K + 1 : 0;
        const uint32_t seg_s = (have && R.seg_read) ? __ldg(R.seg_start + item) : 0u;
        const uint32_t seg_e = R.seg_read ? min(seg_s + R.seg_len, nk_read) : nk_read;  // screen positions [seg_s, seg_e)
        const uint32_t q0 = seg_s & ~15u;                                             // streamed from a word boundary
        const uint32_t span = seg_e > seg_s ? seg_e - q0 : 0u;
        const uint32_t span_max = __reduce_max_sync(FULL, span);
        const uint32_t* wp = R.words + (have ? (R.stride_words ? r * R.stride_words : __ldg(R.word_off + r)) : 0ull) + (q0 >> 4);
        const uint32_t nwords = span ? ((len + 15) >> 4) - (q0 >> 4) : 0u;  // words that may be read from wp
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < span_max; c0 += CW * 16) {
            const uint32_t w0 = c0 >> 4;
            uint32_t cw[CW + 2];
            if (wide) {
#pragma unroll
                for (int i = 0; i <= CW; i += 2) {
                    uint2 t = make_uint2(0u, 0u);
                    if (w0 + i < nwords) t = __ldg(reinterpret_cast<const uint2*>(wp + w0 + i));  // nwords bounds the pair: the stride is even
                    cw[i] = t.x;
                    cw[i + 1] = (w0 + i + 1 < nwords) ? t.y : 0u;
                }
            } else {
#pragma unroll
                for (int i = 0; i <= CW; ++i) cw[i] = (w0 + i < nwords) ? __ldg(wp + w0 + i) : 0u;
            }
            uint32_t f[NF];
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = 0u;
#pragma unroll
            for (int i = 0; i < CW; ++i) {
                if (c0 + i * 16 < span_max) {  // warp-uniform: skip words past every lane's last position
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // 32 bits ENDING at the last base of the k-mer that starts at base j of word i; the older
                        // bases above bit 2K are removed by the multiply
                        const int n = 64 - 2 * (j + K);
                        const uint32_t v = (n >= 32) ? (cw[i] >> ((n - 32) & 31)) : __funnelshift_r(cw[i + 1], cw[i], n & 31);
                        const uint32_t p = v * (SCREEN_MUL << (32 - 2 * K));
                        uint32_t word;
                        if (V & 1) {
                            uint32_t addr;
                            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(__umulhi(p, n_fw)), "r"(SC.four), "r"(s_base));
                            asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
                        } else {
                            word = s_kf[__umulhi(p, n_fw)];
                        }
                        const uint32_t s2 = (V & 2) ? __umulhi(p, SC.two21) : (p >> 11);
                        const uint32_t t = __funnelshift_r(word, 0u, v) & __funnelshift_r(word, 0u, s2);
                        asm("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %1, 1;\n\tsetp.ne.u32 q, t, 0;\n\t@q or.b32 %0, %0, %2;\n\t}"
                            : "+r"(f[i >> 1])
                            : "r"(t), "r"(1u << ((i & 1) * 16 + j)));
                    }
                }
            }
            // keep the flags of positions inside [seg_s, seg_e), count them, reserve queue space per warp
            uint32_t cnt = 0;
            {
                const uint32_t lo_cut = (c0 == 0) ? (seg_s - q0) : 0u;                          // < 16
                const uint32_t hi_cut = span > c0 ? min(span - c0, (uint32_t)(CW * 16)) : 0u;  // valid positions in this chunk
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    const uint32_t base = i * 32;
                    const uint32_t hi_n = hi_cut > base ? min(hi_cut - base, 32u) : 0u;
                    uint32_t m = hi_n >= 32u ? 0xffffffffu : ((1u << hi_n) - 1u);
                    if (i == 0) m &= ~((1u << lo_cut) - 1u);
                    f[i] &= m;
                    cnt += __popc(f[i]);
                }
            }
            const uint32_t total = __reduce_add_sync(FULL, cnt);
            if (total) {
                unsigned long long o = 0;
                if (lane == 0) o = atomicAdd(queue_count, (unsigned long long)total);  // in flight during the prefix scan
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += t;
                }
                o = __shfl_sync(FULL, o, 0);
                // a warp whose entries do not all fit writes none of them: the host sees the overflow in the counter and
                // redoes the batch with a larger queue, so one bounds test per warp is enough
                if (o + total <= queue_cap) {
                    unsigned long long* qp = queue + o + (incl - cnt);
                    uint32_t* kp = queue_kmer + o + (incl - cnt);
                    const uint32_t tag = (uint32_t)r, pos0 = q0 + c0;
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        uint32_t m = f[i];
                        while (m) {
                            const uint32_t bit = __ffs(m) - 1;
                            m &= m - 1u;
                            *reinterpret_cast<uint2*>(qp++) = make_uint2(pos0 + i * 32 + bit, tag);  // read << 32 | position
                            // the k-mer travels with the entry, so the resolve kernel's index probe needs no access to the read
                            const bool up = bit >= 16u;
                            const uint32_t a = up ? cw[2 * i + 1] : cw[2 * i], b = up ? cw[2 * i + 2] : cw[2 * i + 1];
                            *kp++ = __funnelshift_l(b, a, 2u * (bit & 15u)) >> (32 - 2 * K);
                        }
                    }
                }
            }
        }
        tile = next_tile;
        next_tile = __shfl_sync(FULL, after_next, 0);
    }
}

// Resolve the flagged (read, position) pairs.  Phase 1, one thread per queue entry: canonical hash of the k-mer, index
// probe — exact, so the Bloom filter's false positives (~80 % of the queue) end here; the survivors are compacted into
// shared memory.  Phase 2, one thread per survivor in dense warps: the minimizer test restricted to that position.
//   W > 0: compile-time window (2W-2+K bases around the position fit three aligned words): the words are fetched once,
//          shifted so that every neighbour sits at a static offset, and all 2W-1 canonical hashes are computed branch-free;
//   W == 0: any window, neighbours fetched one by one with early exit.
constexpr int RESOLVE_THREADS = 256;
template <int W, int K>
__global__ void __launch_bounds__(RESOLVE_THREADS) resolve_kernel(DevReads R, DevTable T, uint32_t w_rt, uint32_t k_rt,
                                                                  const unsigned long long* __restrict__ queue,
                                                                  const uint32_t* __restrict__ queue_kmer,
                                                                  const unsigned long long* __restrict__ queue_count,
                                                                  unsigned long long queue_cap,
                                                                  unsigned long long* __restrict__ queue_need,
                                                                  unsigned long long* __restrict__ out_a,
                                                                  unsigned long long* __restrict__ out_b,
                                                                  unsigned long long* __restrict__ out_count,
                                                                  unsigned long long cap) {
    static_assert(W == 0 || 2 * W - 2 + K <= 48, "window + k-mer must fit three aligned words");
    __shared__ unsigned long long s_q[RESOLVE_THREADS];
    __shared__ uint32_t s_rec[RESOLVE_THREADS];
    __shared__ uint32_t s_n;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(queue_need, *queue_count);  // sticky over the chunks of a batch: the host regrows and redoes
    if (*queue_count > queue_cap) return;  // overflow: the queue has unwritten slots and the batch is redone anyway
    const unsigned long long n = *queue_count;
    const uint32_t w = W ? (uint32_t)W : w_rt, k = W ? (uint32_t)K : k_rt;
    const uint32_t S = 32 - 2 * k;
    const uint32_t hm = (S == 0) ? 0xffffffffu : ~((1u << S) - 1u);
    const int tid = threadIdx.x, lane = tid & 31;
    auto canon_of = [&](uint32_t v, uint32_t& strand) {  // v: 32 bits starting at the k-mer's first base
        const uint32_t F = v & hm;
        uint32_t y = __brev(~v & hm);
        y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
        const uint32_t hf = hash_left_aligned(F, S, hm), hr = hash_left_aligned(y << S, S, hm);
        strand = hf <= hr ? 1u : 0u;
        return min(hf, hr);
    };
    for (unsigned long long base = (unsigned long long)blockIdx.x * RESOLVE_THREADS; base < n;
         base += (unsigned long long)gridDim.x * RESOLVE_THREADS) {  // CTA-uniform trip count
        if (tid == 0) s_n = 0;
        __syncthreads();
        // ---- phase 1
        {
            const unsigned long long e = base + tid;
            uint32_t rec = 0;
            unsigned long long q = 0;
            if (e < n) {
                q = queue[e];
                uint32_t strand;
                const uint32_t hv = canon_of(queue_kmer[e] << S, strand) >> S;  // the queued k-mer, left aligned
                uint32_t slot = table_slot(hv, T.slot_bits);
                const uint32_t smask = (1u << T.slot_bits) - 1u;
                while (true) {
                    const uint2 ent = __ldg(T.slots + slot);
                    if (ent.y == 0u) break;
                    if (ent.x == hv) {
                        rec = ent.y;  // rec_begin | rec_count << 24, count >= 1
                        break;
                    }
                    slot = (slot + 1) & smask;
                }
            }
            const uint32_t bal = __ballot_sync(FULL, rec != 0u);
            if (bal) {
                uint32_t o = 0;
                if (lane == 0) o = atomicAdd(&s_n, (uint32_t)__popc(bal));
                o = __shfl_sync(FULL, o, 0) + __popc(bal & ((1u << lane) - 1u));
                if (rec) {
                    s_q[o] = q;
                    s_rec[o] = rec;
                }
            }
        }
        __syncthreads();
        // ---- phase 2
        const uint32_t nf = s_n;
        if ((uint32_t)(tid & ~31) < nf) {  // warp-uniform
            uint32_t emit_n = 0, rec_begin = 0, read_strand = 0, r = 0, pos = 0;
            if ((uint32_t)tid < nf) {
                const unsigned long long q = s_q[tid];
                const uint32_t rec = s_rec[tid];
                r = (uint32_t)(q >> 32);
                pos = (uint32_t)q;
                rec_begin = rec & 0xffffffu;
                const uint32_t len = __ldg(R.lens + r);
                const uint32_t nk = len - k + 1;
                const uint32_t* wp = R.words + (R.stride_words ? (unsigned long long)r * R.stride_words : __ldg(R.word_off + r));
                const uint32_t nwords = (len + 15) >> 4;
                uint32_t run = 1, dummy;
                if (W) {
                    // minimizer test: the neighbours with hash >= h on both sides must cover a window of w positions
                    const int start = (int)pos - (W - 1);  // may be negative near the read start: those words read as 0
                    const int sw = start >> 4;
                    const uint32_t sh = 2u * ((uint32_t)start & 15u);
                    uint32_t x[4], y[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = (sw + i >= 0 && sw + i < (int)nwords) ? __ldg(wp + sw + i) : 0u;
#pragma unroll
                    for (int i = 0; i < 3; ++i) y[i] = __funnelshift_l(x[i + 1], x[i], sh);  // base `start + j` now sits at base j
                    y[3] = 0u;
                    constexpr int c0 = W - 1;
                    const uint32_t h = canon_of(__funnelshift_l(y[(c0 >> 4) + 1], y[c0 >> 4], 2 * (c0 & 15)), read_strand);
                    bool ok = true;
#pragma unroll
                    for (int d = 1; d < (W ? W : 1); ++d) {
                        const int c = W - 1 - d;
                        const uint32_t hn = canon_of(__funnelshift_l(y[(c >> 4) + 1], y[c >> 4], 2 * (c & 15)), dummy);
                        ok = ok && (uint32_t)d <= pos && hn >= h;
                        run += ok ? 1u : 0u;
                    }
                    ok = true;
#pragma unroll
                    for (int d = 1; d < (W ? W : 1); ++d) {
                        const int c = W - 1 + d;
                        const uint32_t hn = canon_of(__funnelshift_l(y[(c >> 4) + 1], y[c >> 4], 2 * (c & 15)), dummy);
                        ok = ok && pos + d < nk && hn >= h;
                        run += ok ? 1u : 0u;
                    }
                } else {
                    auto canon_at = [&](uint32_t x, uint32_t& strand) {  // fetch the two words of position x
                        const uint32_t wi = x >> 4;
                        const uint32_t a = __ldg(wp + wi), b = (wi + 1 < nwords) ? __ldg(wp + wi + 1) : 0u;
                        return canon_of(__funnelshift_l(b, a, 2u * (x & 15u)), strand);
                    };
                    const uint32_t h = canon_at(pos, read_strand);
                    for (uint32_t d = 1; d < w && d <= pos && run < w; ++d) {
                        if (canon_at(pos - d, dummy) < h) break;
                        ++run;
                    }
                    for (uint32_t d = 1; d < w && pos + d < nk && run < w; ++d) {
                        if (canon_at(pos + d, dummy) < h) break;
                        ++run;
                    }
                }
                if (run >= w) emit_n = rec >> 24;
            }
            // one atomic per warp: a single hit counter takes ~1 atomic per clock
            uint32_t incl = emit_n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            if (total) {
                unsigned long long ob = 0;
                if (lane == 0) ob = atomicAdd(out_count, (unsigned long long)total);
                ob = __shfl_sync(FULL, ob, 0) + (incl - emit_n);
                for (uint32_t j = 0; j < emit_n; ++j) {
                    const uint2 rc = __ldg(T.recs + rec_begin + j);
                    const uint32_t fwd = ((rc.y & 1u) == read_strand) ? 1u : 0u;
                    if (ob + j < cap) {
                        out_a[ob + j] = ((unsigned long long)(R.read_id_base + r) << 32) | ((unsigned long long)(rc.y >> 1) << 16) |
                                        ((unsigned long long)(fwd ^ 1u) << 15);
                        out_b[ob + j] = ((unsigned long long)pos << 32) | rc.x;
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <int K, int CW, int V>
static void launch_screen_v(const DevReads& R, const DevTable& T, uint32_t wk, unsigned long long* queue, uint32_t* queue_kmer,
                            unsigned long long* counters, uint64_t queue_cap, int sm_count, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(screen_kernel<K, CW, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SCREEN_MAX_FILTER_WORDS * 4));
        configured = true;
    }
    static const uint32_t prefetch = [] {
        const char* e = getenv("DRPRG_SCREEN_PREFETCH");  // L2 prefetch of the next tile: measured neutral (cold == warm L2 time), off
        return e ? (uint32_t)atoi(e) : 0u;
    }();
    const unsigned long long n_items = R.seg_read ? R.n_segs : R.n_reads;
    const unsigned long long n_ctas = (n_items + SCREEN_THREADS - 1) / SCREEN_THREADS;
    const unsigned grid = (unsigned)std::min<unsigned long long>(n_ctas, (unsigned long long)sm_count);  // one persistent CTA per SM
    screen_kernel<K, CW, V><<<grid, SCREEN_THREADS, (size_t)T.kfilter_words * 4, st>>>(R, T, wk, ScreenConsts{4u, 1u << 21, prefetch}, queue, queue_kmer, counters, queue_cap, counters + 1);
    ++g_launches;
}

#ifndef DRPRG_SCREEN_DEFAULT_VARIANT
#define DRPRG_SCREEN_DEFAULT_VARIANT 1
#endif
template <int K, int CW>
static void launch_screen_one(const DevReads& R, const DevTable& T, uint32_t wk, unsigned long long* queue, uint32_t* queue_kmer,
                              unsigned long long* counters, uint64_t queue_cap, int sm_count, cudaStream_t st) {
    static const int variant = [] {
        const char* e = getenv("DRPRG_SCREEN_VARIANT");
        return e ? atoi(e) & 3 : DRPRG_SCREEN_DEFAULT_VARIANT;
    }();
    switch (variant) {
        case 1: return launch_screen_v<K, CW, 1>(R, T, wk, queue, queue_kmer, counters, queue_cap, sm_count, st);
        case 2: return launch_screen_v<K, CW, 2>(R, T, wk, queue, queue_kmer, counters, queue_cap, sm_count, st);
        case 3: return launch_screen_v<K, CW, 3>(R, T, wk, queue, queue_kmer, counters, queue_cap, sm_count, st);
        default: return launch_screen_v<K, CW, 0>(R, T, wk, queue, queue_kmer, counters, queue_cap, sm_count, st);
    }
}

// screen + resolve; counters = {queue length, ticket, largest queue length wanted (not reset here)}
template <int K>
static void launch_screened(const DevReads& R, const DevTable& T, uint32_t w, unsigned long long* a, unsigned long long* b,
                            unsigned long long* cnt, uint64_t cap, int sm_count, uint32_t max_len,
                            unsigned long long* queue, uint32_t* queue_kmer, uint64_t queue_cap, unsigned long long* counters,
                            cudaStream_t st) {
    cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), st);
    if (!R.seg_read && max_len >= (uint32_t)K && max_len - K + 1 <= 10 * 16) launch_screen_one<K, 10>(R, T, w + K, queue, queue_kmer, counters, queue_cap, sm_count, st);
    else launch_screen_one<K, 8>(R, T, w + K, queue, queue_kmer, counters, queue_cap, sm_count, st);
    // ~2 queue entries per read; the grid-stride loop reads the real length on the device
    const unsigned long long n_items = R.seg_read ? R.n_segs : R.n_reads;
    const unsigned grid = (unsigned)std::min<unsigned long long>((n_items * 2 + 255) / 256 + 1, 8ull * (unsigned)sm_count);
    if (w == 11) resolve_kernel<11, K><<<grid, RESOLVE_THREADS, 0, st>>>(R, T, w, K, queue, queue_kmer, counters, queue_cap, counters + 2, a, b, cnt, cap);
    else if (w == 14) resolve_kernel<14, K><<<grid, RESOLVE_THREADS, 0, st>>>(R, T, w, K, queue, queue_kmer, counters, queue_cap, counters + 2, a, b, cnt, cap);
    else resolve_kernel<0, K><<<grid, RESOLVE_THREADS, 0, st>>>(R, T, w, K, queue, queue_kmer, counters, queue_cap, counters + 2, a, b, cnt, cap);
    ++g_launches;
}

static int grid_for(int sm_count, uint64_t n_reads) {
    // persistent-style grid: a multiple of the SM count, capped by the work available
    long long want = (long long)((n_reads + WARPS - 1) / WARPS);
    long long g = (long long)sm_count * 8;
    if (want < g) g = want;
    return (int)(g < 1 ? 1 : g);
}

void launch_sketch_lookup(const DevReads& R, const DevTable& T, uint32_t w, uint32_t k, unsigned long long* d_hi,
                          unsigned long long* d_lo, unsigned long long* d_hit_count, uint64_t hit_cap, int sm_count,
                          uint32_t max_len, cudaStream_t st, unsigned long long* d_queue, uint64_t queue_cap,
                          unsigned long long* d_screen_counters, uint32_t* d_queue_kmer) {
    if (R.n_reads == 0) return;
    static const bool screen_on = [] {
        const char* e = getenv("DRPRG_SCREEN");  // DRPRG_SCREEN=0 sketches every read (A/B measurements, parity tests)
        return !e || atoi(e) != 0;
    }();
    if (screen_on && d_queue && d_queue_kmer && T.kfilter && (max_len <= SHORT_READ_MAX || R.seg_read) && k == 15)
        return launch_screened<15>(R, T, w, d_hi, d_lo, d_hit_count, hit_cap, sm_count, max_len, d_queue, d_queue_kmer, queue_cap, d_screen_counters, st);
    if ((max_len <= SHORT_READ_MAX || R.seg_read) && launch_short<true>(R, T, w, k, d_hi, d_lo, d_hit_count, hit_cap, sm_count, st)) return;
    sketch_kernel<true><<<grid_for(sm_count, R.n_reads), WARPS * 32, 0, st>>>(R, T, w, k, d_hi, d_lo, d_hit_count, hit_cap);
    ++g_launches;
}

void launch_sketch_only(const DevReads& R, uint32_t w, uint32_t k, unsigned long long* d_key, unsigned long long* d_val,
                        unsigned long long* d_count, uint64_t cap, int sm_count, uint32_t max_len, cudaStream_t st) {
    if (R.n_reads == 0) return;
    DevTable T{};
    if ((max_len <= SHORT_READ_MAX || R.seg_read) && launch_short<false>(R, T, w, k, d_key, d_val, d_count, cap, sm_count, st)) return;
    sketch_kernel<false><<<grid_for(sm_count, R.n_reads), WARPS * 32, 0, st>>>(R, T, w, k, d_key, d_val, d_count, cap);
    ++g_launches;
}

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

// ============================================================================================
// S7 : node log-probabilities and the max-likelihood path
// ============================================================================================
__device__ __forceinline__ uint32_t cov_sat(int32_t c) { return c > 65535 ? 65535u : (uint32_t)c; }  // uint16 upstream

__device__ double node_log_prob(const ModelParams& P, uint32_t f, uint32_t r, bool terminal) {
    if (P.bin) {
        if (terminal) return 0.0;
        const uint32_t s = f + r;
        const double n = (double)(s > P.exp_depth ? s : P.exp_depth);
        const double lnck2 = lgamma(n + 1.0) - lgamma((double)f + 1.0) - lgamma((double)r + 1.0) - lgamma(n - (double)f - (double)r + 1.0);
        if (s > P.exp_depth) return lnck2 + (double)s * log(P.bin_p / 2);
        return lnck2 + (double)s * log(P.bin_p / 2) + (double)(P.exp_depth - s) * log(1 - P.bin_p);
    }
    const double c = (double)f + (double)r;
    const double v = lgamma(c + P.nb_r) - lgamma(P.nb_r) - lgamma(c + 1.0) + P.nb_r * log(P.nb_p) + c * log(1.0 - P.nb_p);
    const double FLOOR = -(double)FLT_MAX / 1000.0;
    return v > FLOOR ? v : FLOOR;
}

__global__ void node_prob_kernel(const int32_t* __restrict__ cov, uint32_t total, const uint8_t* __restrict__ is_terminal,
                                 ModelParams P, double* __restrict__ prob) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    prob[g] = node_log_prob(P, cov_sat(cov[2 * g]), cov_sat(cov[2 * g + 1]), is_terminal[g] != 0);
}

void launch_node_prob(const int32_t* d_cov, uint32_t total_knodes, const uint8_t* d_is_terminal, ModelParams P,
                      double* d_prob, cudaStream_t st) {
    if (!total_knodes) return;
    node_prob_kernel<<<(total_knodes + 255) / 256, 256, 0, st>>>(d_cov, total_knodes, d_is_terminal, P, d_prob);
    ++g_launches;
}

// histogram of floor(log-prob + 200) over the inner k-mer nodes of the loci present in the sample: the
// data-parallel half of pandora's estimate_parameters (the valley search itself is a 200-bin host scan)
__global__ void prob_hist_kernel(const double* __restrict__ prob, uint32_t total, const uint8_t* __restrict__ is_terminal,
                                 const uint32_t* __restrict__ knode_locus, const int32_t* __restrict__ locus_reads,
                                 uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[200];
    for (int i = threadIdx.x; i < 200; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < total && !is_terminal[g] && locus_reads[knode_locus[g]] > 0) {
        const double p = prob[g];
        if (p >= -200.0 && p < 0.0) {
            const int j = (int)floor(p + 200.0);
            if (j >= 0 && j < 200) atomicAdd(&sh[j], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 200; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// histogram of the per-node total coverage (0..999) over the inner k-mer nodes of the loci present in the sample:
// the data-parallel half of pandora's estimate_parameters moments / peak search (the 1000-bin scans stay on the host)
__global__ void cov_hist_kernel(const int32_t* __restrict__ cov, uint32_t total, const uint8_t* __restrict__ is_terminal,
                                const uint32_t* __restrict__ knode_locus, const int32_t* __restrict__ locus_reads,
                                uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[1000];
    for (int i = threadIdx.x; i < 1000; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x)
        if (!is_terminal[g] && locus_reads[knode_locus[g]] > 0) {
            const uint32_t c = cov_sat(max(cov[2 * g], 0)) + cov_sat(max(cov[2 * g + 1], 0));
            if (c < 1000u) atomicAdd(&sh[c], 1u);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 1000; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

void launch_cov_hist(const int32_t* d_cov, uint32_t total, const uint8_t* d_is_terminal, const uint32_t* d_knode_locus,
                     const int32_t* d_locus_reads, uint32_t* d_hist1000, cudaStream_t st) {
    cudaMemsetAsync(d_hist1000, 0, 1000 * sizeof(uint32_t), st);
    if (!total) return;
    const unsigned grid = std::min<unsigned>((total + 1023) / 1024, 64u);
    cov_hist_kernel<<<grid, 1024, 0, st>>>(d_cov, total, d_is_terminal, d_knode_locus, d_locus_reads, d_hist1000);
    ++g_launches;
}

void launch_prob_hist(const double* d_prob, uint32_t total, const uint8_t* d_is_terminal, const uint32_t* d_knode_locus,
                      const int32_t* d_locus_reads, uint32_t* d_hist, cudaStream_t st) {
    cudaMemsetAsync(d_hist, 0, 200 * sizeof(uint32_t), st);
    if (!total) return;
    prob_hist_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_prob, total, d_is_terminal, d_knode_locus, d_locus_reads, d_hist);
    ++g_launches;
}

// pandora's find_prob_thresh on the 200-bin histogram (same scan as prob_threshold() on the host): the valley between
// the error peak and the signal peak.  One thread; lets the ML-path kernel start without a host round trip.
__global__ void prob_thresh_kernel(const uint32_t* __restrict__ hist, int any_present, int fallback, double* __restrict__ out_f64,
                                   int* __restrict__ out_i32) {
    __shared__ uint32_t ph[200];  // three scans by one thread: from shared memory they cost ~1 us, from global ~20 us
    for (int i = threadIdx.x; i < 200; i += blockDim.x) ph[i] = hist[i];
    __syncthreads();
    if (threadIdx.x) return;
    int t = fallback;
    if (any_present) {
        int p1 = 0, p2 = -1;
        for (int i = 1; i < 200; ++i)
            if (ph[i] > ph[p1]) p1 = i;  // first maximum
        for (int i = 0; i < 200; ++i) {
            const int d = i > p1 ? i - p1 : p1 - i;
            if (d <= 10 || ph[i] == 0) continue;
            if (p2 < 0 || ph[i] > ph[p2]) p2 = i;
        }
        if (p2 < 0) {
            t = p1 - 200 - 10 > -200 ? p1 - 200 - 10 : -200;
        } else {
            const int a = p1 < p2 ? p1 : p2, b = p1 < p2 ? p2 : p1;
            int m = a;
            for (int i = a + 1; i <= b; ++i)
                if (ph[i] < ph[m]) m = i;  // first minimum
            t = m - 200;
        }
    }
    *out_i32 = t;
    *out_f64 = (double)t;
}

void launch_prob_thresh(const uint32_t* d_hist200, bool any_present, int fallback, double* d_thresh_f64, int* d_thresh_i32,
                        cudaStream_t st) {
    prob_thresh_kernel<<<1, 256, 0, st>>>(d_hist200, any_present ? 1 : 0, fallback, d_thresh_f64, d_thresh_i32);
    ++g_launches;
}

// One warp per locus.  The recurrence is a chain (node j needs its successors), and the choice
// among successors is order dependent (1e-6 tolerance, longer path wins ties), so lane 0 walks the
// nodes in reverse rank order; the windowed mean needs the node `window` steps down the chosen
// path, found in O(log window) with binary-lifting pointers instead of pandora's linear walk.
template <bool IN_SMEM, int LVT>
__device__ __forceinline__ void mlpath_chain(uint32_t n, double* __restrict__ M, double* __restrict__ mean,
                                             const double* __restrict__ prb, uint32_t* __restrict__ len,
                                             uint32_t* __restrict__ up, uint32_t up_stride,
                                             const uint32_t* __restrict__ eo, const uint32_t* __restrict__ ed,
                                             const ModelParams& P) {
    const double tol = 0.000001;
    const uint32_t term = n - 1;
    const uint32_t steps = P.window - 1;  // the node `window` steps down the chosen path = window-1 steps after v
    M[term] = 0.0;
    len[term] = 0;
    if (IN_SMEM) mean[term] = 0.0;  // never read: the terminus competes with the threshold instead
#pragma unroll
    for (int v = 0; v < LVT; ++v) up[v * up_stride + term] = term;
    for (uint32_t j = term; j-- > 0;) {
        double max_mean = -(double)FLT_MAX;
        uint32_t max_len = 0;
        double Mj = 0.0;
        uint32_t lenj = 0, prevj = term;
        const double pj = prb[j];
        const uint32_t e1 = eo[j + 1];
        for (uint32_t e = eo[j]; e < e1; ++e) {
            const uint32_t v = ed[e];
            const bool is_term = (v == term);
            const uint32_t lv = len[v];
            double mean_v = 0.0;
            if (!is_term) mean_v = IN_SMEM ? mean[v] : M[v] / (double)lv;
            const bool take = is_term ? (P.thresh > max_mean + tol)
                                      : ((mean_v > max_mean + tol) || (max_mean - mean_v <= tol && lv > max_len));
            if (!take) continue;
            Mj = pj + M[v];
            lenj = 1 + lv;
            prevj = v;
            if (lenj > P.window) {
                uint32_t pn = v;
#pragma unroll
                for (int b = 0; b < LVT; ++b)
                    if ((steps >> b) & 1u) pn = up[b * up_stride + pn];
                Mj -= prb[pn];
                lenj -= 1;
            }
            max_mean = is_term ? P.thresh : mean_v;
            if (!is_term) max_len = lv;
        }
        M[j] = Mj;
        len[j] = lenj;
        if (IN_SMEM) mean[j] = Mj / (double)lenj;  // 0/0 = NaN for a dead end: never chosen, like pandora
        up[j] = prevj;
        uint32_t a = prevj;
#pragma unroll
        for (int v = 1; v < LVT; ++v) {
            a = up[(v - 1) * up_stride + a];
            up[v * up_stride + j] = a;
        }
    }
}

template <int LVT>
__global__ void mlpath_kernel(uint32_t n_loci, const uint32_t* __restrict__ knode_base, const uint32_t* __restrict__ edge_off,
                              const uint32_t* __restrict__ edges, const double* __restrict__ prob,
                              const int32_t* __restrict__ locus_reads, ModelParams P, double* __restrict__ gM,
                              uint32_t* __restrict__ glen, uint32_t* __restrict__ gup, uint32_t total,
                              uint32_t* __restrict__ path, uint32_t* __restrict__ path_len, uint32_t smem_nodes,
                              uint32_t smem_edges) {
    extern __shared__ double s_dyn[];
    const uint32_t l = blockIdx.x;
    if (l >= n_loci) return;
    const uint32_t base = knode_base[l], n = knode_base[l + 1] - base;
    if (locus_reads[l] <= 0 || n < 2) {
        if (threadIdx.x == 0) path_len[l] = 0xffffffffu;
        return;
    }
    // The chain's whole working set lives in shared memory when the locus fits: running sum, cached mean
    // (sum / length: one fp64 division per node instead of one per edge visit), node score, length, the
    // binary-lifting pointers and the locus's CSR edges.  Every step of the serial dependency is then an
    // LDS with 32-bit addressing instead of an L2 round trip.
    const uint32_t e_base = edge_off[base], n_edges = edge_off[base + n] - e_base;
    const bool in_smem = n <= smem_nodes && n_edges <= smem_edges;
    const uint32_t term = n - 1;
    uint32_t cnt = 0;
    if (in_smem) {
        double* M = s_dyn;
        double* mean = s_dyn + smem_nodes;
        double* pr = s_dyn + 2 * (size_t)smem_nodes;
        uint32_t* len = (uint32_t*)(s_dyn + 3 * (size_t)smem_nodes);
        uint32_t* up = len + smem_nodes;
        uint32_t* s_eoff = up + (size_t)LVT * smem_nodes;
        uint32_t* s_edges = s_eoff + smem_nodes + 2;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) pr[i] = prob[base + i];
        for (uint32_t i = threadIdx.x; i <= n; i += blockDim.x) s_eoff[i] = edge_off[base + i] - e_base;
        for (uint32_t i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = edges[e_base + i];
        __syncwarp();
        if (threadIdx.x != 0) return;
        mlpath_chain<true, LVT>(n, M, mean, pr, len, up, smem_nodes, s_eoff, s_edges, P);
        uint32_t p = up[0];
        while (p < term && cnt < n) {
            path[base + cnt++] = p;
            p = up[p];
        }
    } else {
        if (threadIdx.x != 0) return;
        uint32_t* up = gup + base;
        mlpath_chain<false, LVT>(n, gM + base, nullptr, prob + base, glen + base, up, total, edge_off + base, edges, P);
        uint32_t p = up[0];
        while (p < term && cnt < n) {
            path[base + cnt++] = p;
            p = up[p];
        }
    }
    path_len[l] = cnt;
}

// ---- fast variant: 64-byte node records addressed by their shared-memory address ---------------------------
// The generic kernel above spends ~160 instructions and ~1000 cycles per node, two thirds of it address
// arithmetic and dependent-issue waits (ncu: stall_wait 2.4, stall_short_scoreboard 2.6 per issued instruction).
// Here every pointer the chain follows (successor, lifting pointers, edge targets) is stored as the 32-bit
// shared-memory ADDRESS of the target's 64-byte record, so one hop is a single LDS with an immediate offset.
constexpr uint32_t REC = 64, R_M = 0, R_MEAN = 8, R_PR = 16, R_LEN = 24, R_EOFF = 28, R_UP = 32, R_T = 60;
// record: sum f64 | mean f64 | score f64 | len u32 | edge-list address u32 | lifting pointers up[0..6] | T
// T = the node window-1 steps down the chosen path (what a predecessor subtracts when its window is full)

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }

// After the successor v of node a is known, two pointer chases remain, both through tables of OLDER nodes only and
// therefore independent of each other and of the fp chain: the lifting pointers of a (level l of a = level l-1 of
// the node 2^(l-1) steps down) and T(a) = succ^(window-2)(v).  Their loads are issued interleaved so the two chains
// overlap instead of adding up (volatile asm keeps this order).
__device__ __forceinline__ void mlpath_link(uint32_t a, uint32_t v, uint32_t steps2) {
    sts32(a + R_UP, v);
    uint32_t x = v, t = v;
#pragma unroll
    for (int l = 0; l < 7; ++l) {
        const uint32_t xn = (l < 6) ? lds32(x + R_UP + 4 * l) : 0u;          // level l+1 of a
        const uint32_t tn = ((steps2 >> l) & 1u) ? lds32(t + R_UP + 4 * l) : t;  // walk window-2 steps from v
        if (l < 6) {
            sts32(a + R_UP + 4 * (l + 1), xn);
            x = xn;
        }
        t = tn;
    }
    sts32(a + R_T, t);
}

__global__ void mlpath_rec_kernel(uint32_t n_loci, const uint32_t* __restrict__ knode_base, const uint32_t* __restrict__ edge_off,
                                  const uint32_t* __restrict__ edges, const double* __restrict__ prob,
                                  const int32_t* __restrict__ locus_reads, const uint8_t* __restrict__ needs_mean,
                                  ModelParams P, uint32_t* __restrict__ path, uint32_t* __restrict__ path_len,
                                  uint32_t max_nodes) {
    extern __shared__ double s_dyn[];
    const uint32_t l = blockIdx.x;
    if (l >= n_loci) return;
    const uint32_t base = knode_base[l], n = knode_base[l + 1] - base;
    if (locus_reads[l] <= 0 || n < 2) {
        if (threadIdx.x == 0) path_len[l] = 0xffffffffu;
        return;
    }
    const uint32_t e_base = edge_off[base], n_edges = edge_off[base + n] - e_base;
    const uint32_t recs = (uint32_t)__cvta_generic_to_shared(s_dyn);  // n + 1 records (the extra one carries the edge end)
    const uint32_t edg = recs + (max_nodes + 1) * REC;                // successor record addresses
    for (uint32_t i = threadIdx.x; i <= n; i += blockDim.x) {
        const uint32_t a = recs + i * REC;
        if (i < n) sts64(a + R_PR, prob[base + i]);
        // edge list address (4-byte aligned) | bit0: some predecessor has a choice, so this node's mean is compared
        sts32(a + R_EOFF, (edg + (edge_off[base + i] - e_base) * 4u) | (i < n ? (uint32_t)needs_mean[base + i] : 0u));
    }
    for (uint32_t i = threadIdx.x; i < n_edges; i += blockDim.x) sts32(edg + i * 4u, recs + edges[e_base + i] * REC);
    __syncwarp();
    if (threadIdx.x != 0) return;
    const double tol = 0.000001;
    const uint32_t term = recs + (n - 1) * REC;
    const uint32_t steps2 = P.window - 2;  // T(a) = succ^(window-1)(a) = succ^(window-2)(chosen successor)
    sts64(term + R_M, 0.0);
    sts32(term + R_LEN, 0u);
#pragma unroll
    for (int v = 0; v < 8; ++v) sts32(term + R_UP + 4 * v, term);  // includes R_T
    for (uint32_t a = term - REC; a + REC > recs; a -= REC) {  // nodes n-2 .. 0
        double Mj = 0.0;
        uint32_t lenj = 0, prevj = term;
        const double pj = lds64(a + R_PR);
        const uint32_t e1 = lds32(a + REC + R_EOFF) & ~3u;
        const uint32_t e0w = lds32(a + R_EOFF);
        const uint32_t e0 = e0w & ~3u;
        if (e1 - e0 == 4u) {
            // single successor (most nodes): no choice to make.  pandora's comparison against the initial
            // -FLT_MAX accepts any successor with a real mean, i.e. any successor that is not a dead end.
            const uint32_t v = lds32(e0);
            const uint32_t lv = lds32(v + R_LEN);
            const uint32_t tv = lds32(v + R_T);
            const double Mv = lds64(v + R_M);
            if (v == term || lv > 0u) {
                mlpath_link(a, v, steps2);  // independent of the sums below: overlaps them
                prevj = v;
                lenj = 1 + lv;
                Mj = pj + Mv;
                if (lenj > P.window) {
                    Mj -= lds64(tv + R_PR);
                    lenj -= 1;
                }
            }
        } else {
            double max_mean = -(double)FLT_MAX;
            uint32_t max_len = 0;
            for (uint32_t e = e0; e < e1; e += 4u) {
                const uint32_t v = lds32(e);
                const bool is_term = (v == term);
                const uint32_t lv = lds32(v + R_LEN);
                const uint32_t tv = lds32(v + R_T);
                const double mean_v = lds64(v + R_MEAN);
                const double Mv = lds64(v + R_M);
                const bool take = is_term ? (P.thresh > max_mean + tol)
                                          : ((mean_v > max_mean + tol) || (max_mean - mean_v <= tol && lv > max_len));
                if (!take) continue;
                Mj = pj + Mv;
                lenj = 1 + lv;
                prevj = v;
                if (lenj > P.window) {
                    Mj -= lds64(tv + R_PR);
                    lenj -= 1;
                }
                max_mean = is_term ? P.thresh : mean_v;
                if (!is_term) max_len = lv;
            }
            if (lenj) mlpath_link(a, prevj, steps2);
        }
        if (!lenj) {  // dead end (or no acceptable successor): points at the terminus like pandora's default
#pragma unroll
            for (int v = 0; v < 8; ++v) sts32(a + R_UP + 4 * v, term);
        }
        sts64(a + R_M, Mj);
        sts32(a + R_LEN, lenj);
        if (e0w & 1u) sts64(a + R_MEAN, Mj / (double)lenj);  // 0/0 = NaN for a dead end: never chosen, like pandora
    }
    uint32_t cnt = 0, p = lds32(recs + R_UP);
    while (p != term && cnt < n) {
        path[base + cnt++] = (p - recs) / REC;
        p = lds32(p + R_UP);
    }
    path_len[l] = cnt;
}

// ---- run-parallel variant ---------------------------------------------------------------------------------
// Most k-mer nodes have a single successor, so which node follows them is known from the graph alone; only the
// running sums are sequential.  The host cuts every locus into UNITS in processing order: a node with a choice (or
// none), or a RUN of up to 32 single-successor nodes j_1 <- j_2 <- ... hanging off an already finished node b.  For
// a run every pointer a node needs (lifting pointers, window tail T, the node whose score leaves the window) is
// succ^m(j_i) = j_(i-m) inside the run or a walk of m-i steps from b through finished tables: all 32 lanes resolve
// their node's pointers at once (independent loads, no stores in between), then lane 0 adds up the sums in pandora's
// order from staged operands (two dependent DADDs per node instead of ~10 dependent shared-memory hops).
struct MlUnitsDev {
    const uint32_t* locus_unit_off;  // n_loci + 1
    const uint32_t* unit_start;      // n_units + 1 -> unit_nodes
    const uint32_t* unit_nodes;      // ranks within the locus, chain order (highest rank first)
};

__device__ __forceinline__ uint32_t ldsx32(uint32_t a, uint32_t token) {  // reorderable load, tied to the unit by `token`
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a), "r"(token));
    return v;
}
__device__ __forceinline__ double ldsx64(uint32_t a, uint32_t token) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a), "r"(token));
    return v;
}
__device__ __forceinline__ uint32_t ml_walk(uint32_t x, uint32_t steps, uint32_t token) {
#pragma unroll
    for (int l = 0; l < 7; ++l)
        if ((steps >> l) & 1u) x = ldsx32(x + R_UP + 4 * l, token);
    return x;
}

__global__ void mlpath_unit_kernel(uint32_t n_loci, const uint32_t* __restrict__ knode_base, const uint32_t* __restrict__ edge_off,
                                   const uint32_t* __restrict__ edges, const double* __restrict__ prob,
                                   const int32_t* __restrict__ locus_reads, const uint8_t* __restrict__ needs_mean,
                                   MlUnitsDev U, ModelParams P, uint32_t* __restrict__ path, uint32_t* __restrict__ path_len,
                                   uint32_t max_nodes, uint32_t max_edges) {
    extern __shared__ double s_dyn[];
    const uint32_t l = blockIdx.x;
    if (l >= n_loci) return;
    const uint32_t lane = threadIdx.x;
    const uint32_t base = knode_base[l], n = knode_base[l + 1] - base;
    if (locus_reads[l] <= 0 || n < 2) {
        if (lane == 0) path_len[l] = 0xffffffffu;
        return;
    }
    const uint32_t e_base = edge_off[base], n_edges = edge_off[base + n] - e_base;
    const uint32_t recs = (uint32_t)__cvta_generic_to_shared(s_dyn);
    const uint32_t edg = recs + (max_nodes + 1) * REC;
    const uint32_t scr = (edg + (max_edges + 1) * 4u + 7u) & ~7u;  // staging: 32 x {p f64, q f64, node u32}
    const uint32_t s_p = scr, s_q = scr + 256, s_n = scr + 512;
    for (uint32_t i = lane; i <= n; i += 32) {
        const uint32_t a = recs + i * REC;
        if (i < n) sts64(a + R_PR, prob[base + i]);
        sts32(a + R_EOFF, (edg + (edge_off[base + i] - e_base) * 4u) | (i < n ? (uint32_t)needs_mean[base + i] : 0u));
    }
    for (uint32_t i = lane; i < n_edges; i += 32) sts32(edg + i * 4u, recs + edges[e_base + i] * REC);
    const double tol = 0.000001;
    const uint32_t term = recs + (n - 1) * REC;
    const uint32_t W = P.window;
    if (lane == 0) {
        sts64(term + R_M, 0.0);
        sts32(term + R_LEN, 0u);
#pragma unroll
        for (int v = 0; v < 8; ++v) sts32(term + R_UP + 4 * v, term);  // includes R_T
    }
    __syncwarp();
    const uint32_t u0 = U.locus_unit_off[l], u1 = U.locus_unit_off[l + 1];
    for (uint32_t u = u0; u < u1; ++u) {
        const uint32_t s0 = U.unit_start[u], k = U.unit_start[u + 1] - s0;
        const uint32_t a1 = recs + U.unit_nodes[s0] * REC;
        const uint32_t e0w = lds32(a1 + R_EOFF);
        const uint32_t e0 = e0w & ~3u, e1 = lds32(a1 + REC + R_EOFF) & ~3u;
        if (e1 - e0 != 4u) {
            // ---- a node with a choice (or a dead end): pandora's sequential comparison, one lane
            if (lane == 0) {
                const uint32_t a = a1;
                double Mj = 0.0, max_mean = -(double)FLT_MAX;
                uint32_t lenj = 0, prevj = term, max_len = 0;
                const double pj = lds64(a + R_PR);
                for (uint32_t e = e0; e < e1; e += 4u) {
                    const uint32_t v = lds32(e);
                    const bool is_term = (v == term);
                    const uint32_t lv = lds32(v + R_LEN);
                    const uint32_t tv = lds32(v + R_T);
                    const double mean_v = lds64(v + R_MEAN);
                    const double Mv = lds64(v + R_M);
                    const bool take = is_term ? (P.thresh > max_mean + tol)
                                              : ((mean_v > max_mean + tol) || (max_mean - mean_v <= tol && lv > max_len));
                    if (!take) continue;
                    Mj = pj + Mv;
                    lenj = 1 + lv;
                    prevj = v;
                    if (lenj > W) {
                        Mj -= lds64(tv + R_PR);
                        lenj -= 1;
                    }
                    max_mean = is_term ? P.thresh : mean_v;
                    if (!is_term) max_len = lv;
                }
                if (lenj) {
                    mlpath_link(a, prevj, W - 2);
                } else {
#pragma unroll
                    for (int v = 0; v < 8; ++v) sts32(a + R_UP + 4 * v, term);
                }
                sts64(a + R_M, Mj);
                sts32(a + R_LEN, lenj);
                if (e0w & 1u) sts64(a + R_MEAN, Mj / (double)lenj);
            }
            __syncwarp();
            continue;
        }
        // ---- a run of k single-successor nodes hanging off b
        const uint32_t b = lds32(e0);
        const uint32_t lb = lds32(b + R_LEN);
        const double Mb = lds64(b + R_M);
        const bool active = lane < k;
        const uint32_t me = active ? recs + U.unit_nodes[s0 + lane] * REC : term;
        if (b != term && lb == 0u) {  // hanging off a dead end: the whole run is dead (pandora never takes a NaN mean)
            if (active) {
#pragma unroll
                for (int v = 0; v < 8; ++v) sts32(me + R_UP + 4 * v, term);
                sts64(me + R_M, 0.0);
                sts32(me + R_LEN, 0u);
                if (lds32(me + R_EOFF) & 1u) sts64(me + R_MEAN, 0.0 / 0.0);
            }
            __syncwarp();
            continue;
        }
        if (active) {
            const uint32_t i = lane + 1;  // j_i
            auto succ = [&](uint32_t m) -> uint32_t {  // succ^m(j_i), m >= 1
                return (m < i) ? recs + U.unit_nodes[s0 + lane - m] * REC : ml_walk(b, m - i, u);
            };
            uint32_t up[7];
#pragma unroll
            for (int lv = 0; lv < 7; ++lv) up[lv] = succ(1u << lv);
            const uint32_t T = succ(W - 1);
            const uint32_t qn = succ(W);  // = T(successor of j_i): its score leaves the window when j_i joins a full one
            const double p = ldsx64(me + R_PR, u), q = ldsx64(qn + R_PR, u);
#pragma unroll
            for (int lv = 0; lv < 7; ++lv) sts32(me + R_UP + 4 * lv, up[lv]);
            sts32(me + R_T, T);
            sts64(s_p + 8 * lane, p);
            sts64(s_q + 8 * lane, q);
            sts32(s_n + 4 * lane, me);
        }
        __syncwarp();
        if (lane == 0) {
            double M = Mb;
            uint32_t len = lb;
            for (uint32_t i = 0; i < k; ++i) {
                const uint32_t node = lds32(s_n + 4 * i);
                M = lds64(s_p + 8 * i) + M;
                len += 1;
                if (len > W) {
                    M -= lds64(s_q + 8 * i);
                    len = W;
                }
                sts64(node + R_M, M);
                sts32(node + R_LEN, len);
            }
        }
        __syncwarp();
        if (active && (lds32(me + R_EOFF) & 1u)) sts64(me + R_MEAN, lds64(me + R_M) / (double)lds32(me + R_LEN));
        __syncwarp();
    }
    if (lane == 0) {
        uint32_t cnt = 0, p = lds32(recs + R_UP);
        while (p != term && cnt < n) {
            path[base + cnt++] = (p - recs) / REC;
            p = lds32(p + R_UP);
        }
        path_len[l] = cnt;
    }
}

// ---- level-parallel variant --------------------------------------------------------------------------------
// A node only needs its successors' finished records, so all nodes at the same distance-to-sink ("level": 1 + the
// largest level among the successors) are independent: the alleles of a bubble advance side by side.  The host
// sorts every locus by level, single-successor nodes first within a level.  A locus gets a CTA of four warps: warps
// 0-1 take the level's single-successor nodes, warps 2-3 the nodes with a choice, one node per lane, so the two code
// paths run on different schedulers instead of serialising inside one warp (a lone warp issues one instruction every
// ~6 cycles here: the kernel is bound by its own instruction latency, not by shared memory).  Each node runs exactly
// the per-node step of mlpath_rec_kernel (successors still visited in rank order, so pandora's order-dependent
// tie-breaking is unchanged); the CTA synchronises between levels.  The serial chain shrinks from the number of nodes
// to the number of levels (benchmark panel: 38 251 nodes -> 15 128 levels, widest level 22 nodes).
// sum / length on the level kernel's critical path.  The length is an integer <= 128, so the correctly rounded quotient
// comes from a table of correctly rounded reciprocals and two Markstein corrections (q += fma(-q, n, a) * y): five
// dependent FP64 operations instead of the ~15 of the generic division sequence.  The first correction makes q
// faithful, the second one makes it RN(a / n) because y = RN(1 / n) (Markstein's theorem; also checked against the
// hardware division on 7.7e8 random and adversarial operands, tools/divtest.c).  n == 0 gives NaN like 0.0 / 0.
__constant__ double c_rcp_small[129];
static void upload_rcp_table() {
    static bool done = false;  // per process; every device of the process gets it at its first launch
    static int done_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (done && done_dev == dev) return;
    double h[129];
    h[0] = 0.0;
    for (int n = 1; n <= 128; ++n) h[n] = 1.0 / (double)n;
    cudaMemcpyToSymbol(c_rcp_small, h, sizeof h);
    done = true;
    done_dev = dev;
}
__device__ __forceinline__ double div_small(double a, uint32_t n) {
    if (n == 0u) return __longlong_as_double(0x7ff8000000000000ll);  // what 0.0 / 0 gives: never compares true
    const double y = c_rcp_small[n], b = (double)n;
    double q = a * y;
    q = fma(fma(-q, b, a), y, q);
    q = fma(fma(-q, b, a), y, q);
    return q;
}

// The level kernel's own record layout (64 B), arranged for vector shared-memory accesses: a lone warp per scheduler
// pays ~6 cycles per instruction, so fewer, wider accesses matter more than anything else here.
//   0 sum f64 | 8 mean f64 | 16 score f64 | 24 len u32 | 28 T u32 | 32 edge-list address u32 | 36 up[0..6] u32
// sum+mean come with one LDS.128, len+T with one LDS.64; up[1..6] sit at 40/48/56 and are stored as three STS.64.
constexpr uint32_t L_M = 0, L_MEAN = 8, L_PR = 16, L_LEN = 24, L_T = 28, L_EOFF = 32, L_UP = 36;
static_assert(L_MEAN == L_M + 8 && L_T == L_LEN + 4 && (L_UP + 4) % 8 == 0 && L_UP + 28 == REC, "vector accesses rely on this layout");
__device__ __forceinline__ void lds_sum_mean(uint32_t a, double& m, double& mean) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m), "=d"(mean) : "r"(a + L_M));
}
__device__ __forceinline__ void lds_len_t(uint32_t a, uint32_t& len, uint32_t& t) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(len), "=r"(t) : "r"(a + L_LEN));
}
__device__ __forceinline__ void sts_pair32(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y));
}
// lifting pointers of a (successor v) and T(a) = succ^(window-2)(v): the two pointer chases are issued interleaved like
// mlpath_link; the stores are batched: up[0] | (up[1],up[2]) | (up[3],up[4]) | (up[5],up[6]); T is returned
__device__ __forceinline__ uint32_t level_link(uint32_t a, uint32_t v, uint32_t steps2) {
    uint32_t up[7];
    up[0] = v;
    uint32_t x = v, t = v;
#pragma unroll
    for (int l = 0; l < 7; ++l) {
        const uint32_t xn = (l < 6) ? lds32(x + L_UP + 4 * l) : 0u;              // level l+1 of a
        const uint32_t tn = ((steps2 >> l) & 1u) ? lds32(t + L_UP + 4 * l) : t;  // walk window-2 steps from v
        if (l < 6) {
            up[l + 1] = xn;
            x = xn;
        }
        t = tn;
    }
    sts32(a + L_UP, up[0]);
    sts_pair32(a + L_UP + 4, up[1], up[2]);
    sts_pair32(a + L_UP + 12, up[3], up[4]);
    sts_pair32(a + L_UP + 20, up[5], up[6]);
    return t;
}

constexpr int ML_LEVEL_THREADS = 128;
__global__ void __launch_bounds__(ML_LEVEL_THREADS) mlpath_level_kernel(
    uint32_t n_loci, const uint32_t* __restrict__ knode_base, const uint32_t* __restrict__ edge_off,
    const uint32_t* __restrict__ edges, const double* __restrict__ prob, const int32_t* __restrict__ locus_reads,
    const uint8_t* __restrict__ needs_mean, MlUnitsDev L, const uint32_t* __restrict__ level_singles, ModelParams P,
    uint32_t* __restrict__ path, uint32_t* __restrict__ path_len, uint32_t max_nodes, uint32_t max_edges,
    volatile uint32_t* done,     // done != nullptr: path / path_len are host-mapped and done[l] tells the host that locus l is there
    const double* d_thresh) {    // != nullptr: the probability threshold is read from device memory (computed by prob_thresh_kernel)
    extern __shared__ double s_dyn[];
    const uint32_t l = blockIdx.x;
    if (l >= n_loci) return;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t base = knode_base[l], n = knode_base[l + 1] - base;
    if (locus_reads[l] <= 0 || n < 2) {
        if (tid == 0) {
            path_len[l] = 0xffffffffu;
            if (done) {
                __threadfence_system();
                done[l] = 1u;
            }
        }
        return;
    }
    const uint32_t e_base = edge_off[base], n_edges = edge_off[base + n] - e_base;
    const uint32_t recs = (uint32_t)__cvta_generic_to_shared(s_dyn);
    const uint32_t edg = recs + (max_nodes + 1) * REC;
    const uint32_t lvn = edg + (max_edges + 1) * 4u;  // record addresses in level order
    const uint32_t lvs = (lvn + max_nodes * 4u + 7u) & ~7u;  // per level: first index into lvn, number of single-successor nodes (8-byte aligned pairs)
    for (uint32_t i = tid; i <= n; i += ML_LEVEL_THREADS) {
        const uint32_t a = recs + i * REC;
        if (i < n) sts64(a + L_PR, prob[base + i]);
        sts32(a + L_EOFF, (edg + (edge_off[base + i] - e_base) * 4u) | (i < n ? (uint32_t)needs_mean[base + i] : 0u));
    }
    for (uint32_t i = tid; i < n_edges; i += ML_LEVEL_THREADS) sts32(edg + i * 4u, recs + edges[e_base + i] * REC);
    const uint32_t u0 = L.locus_unit_off[l], n_levels = L.locus_unit_off[l + 1] - u0;
    const uint32_t s00 = L.unit_start[u0];
    for (uint32_t i = tid; i <= n_levels; i += ML_LEVEL_THREADS) {
        sts32(lvs + 8u * i, L.unit_start[u0 + i] - s00);
        sts32(lvs + 8u * i + 4u, i < n_levels ? level_singles[u0 + i] : 0u);
    }
    for (uint32_t i = tid; i + 1 < n; i += ML_LEVEL_THREADS) sts32(lvn + 4u * i, recs + L.unit_nodes[s00 + i] * REC);
    const double tol = 0.000001;
    const double thresh = d_thresh ? *d_thresh : P.thresh;
    const uint32_t term = recs + (n - 1) * REC;
    const uint32_t steps2 = P.window - 2;
    if (tid == 0) {
        sts64(term + L_M, 0.0);
        sts_pair32(term + L_LEN, 0u, term);
#pragma unroll
        for (int v = 0; v < 7; ++v) sts32(term + L_UP + 4 * v, term);
    }
    __syncthreads();
    const bool single_warp = warp < 2;
    uint32_t sub = (warp & 1u) + 2u * lane;  // this lane's slot among the level's nodes of its kind (64 per pass)
    asm volatile("" : "+r"(sub));            // keep it in a register: recomputing it every level costs issue slots
    // level bounds: one LDS.64 per level ({first index, number of single-successor nodes} of the NEXT level; the
    // current pair is carried in registers)
    uint32_t s0, ns;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(s0), "=r"(ns) : "r"(lvs));
    for (uint32_t u = 0; u < n_levels; ++u) {
        uint32_t s1, ns1;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(s1), "=r"(ns1) : "r"(lvs + 8u * (u + 1)));
        const uint32_t lo = single_warp ? s0 : s0 + ns, hi = single_warp ? s0 + ns : s1;
        s0 = s1;
        ns = ns1;
        for (uint32_t idx = lo + sub; idx < hi; idx += 64) {
            const uint32_t a = lds32(lvn + 4u * idx);
            double Mj = 0.0;
            uint32_t lenj = 0, prevj = term;
            const double pj = lds64(a + L_PR);
            const uint32_t e0w = lds32(a + L_EOFF);
            const uint32_t e0 = e0w & ~3u;
            uint32_t Tj = term;
            if (single_warp) {
                // single successor: pandora's comparison against the initial -FLT_MAX accepts any successor that is
                // not a dead end
                const uint32_t v = lds32(e0);
                uint32_t lv, tv;
                lds_len_t(v, lv, tv);
                const double Mv = lds64(v + L_M);
                if (v == term || lv > 0u) {
                    Tj = level_link(a, v, steps2);  // independent of the sums below: overlaps them
                    prevj = v;
                    lenj = 1 + lv;
                    Mj = pj + Mv;
                    if (lenj > P.window) {
                        Mj -= lds64(tv + L_PR);
                        lenj -= 1;
                    }
                }
            } else {
                const uint32_t e1 = lds32(a + REC + L_EOFF) & ~3u;
                double max_mean = -(double)FLT_MAX;
                uint32_t max_len = 0;
                for (uint32_t e = e0; e < e1; e += 4u) {
                    const uint32_t v = lds32(e);
                    const bool is_term = (v == term);
                    uint32_t lv, tv;
                    lds_len_t(v, lv, tv);
                    double Mv, mean_v;
                    lds_sum_mean(v, Mv, mean_v);
                    const bool take = is_term ? (thresh > max_mean + tol)
                                              : ((mean_v > max_mean + tol) || (max_mean - mean_v <= tol && lv > max_len));
                    if (!take) continue;
                    Mj = pj + Mv;
                    lenj = 1 + lv;
                    prevj = v;
                    if (lenj > P.window) {
                        Mj -= lds64(tv + L_PR);
                        lenj -= 1;
                    }
                    max_mean = is_term ? thresh : mean_v;
                    if (!is_term) max_len = lv;
                }
                if (lenj) Tj = level_link(a, prevj, steps2);
            }
            if (!lenj) {  // dead end (or no acceptable successor): points at the terminus like pandora's default
#pragma unroll
                for (int v = 0; v < 7; ++v) sts32(a + L_UP + 4 * v, term);
            }
            sts_pair32(a + L_LEN, lenj, Tj);
            if (e0w & 1u) {  // 0/0 = NaN for a dead end: never chosen, like pandora
                const double mean_j = div_small(Mj, lenj);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a + L_M), "d"(Mj), "d"(mean_j));
            } else {
                sts64(a + L_M, Mj);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t cnt = 0, p = lds32(recs + L_UP);
        while (p != term && cnt < n) {
            path[base + cnt++] = (p - recs) / REC;
            p = lds32(p + L_UP);
        }
        path_len[l] = cnt;
        if (done) {  // the loci finish at different times (170 .. 860 levels): the host verifies each one as it lands
            __threadfence_system();
            done[l] = 1u;
        }
    }
}

bool launch_mlpath(uint32_t n_loci, const uint32_t* d_knode_base, const uint32_t* d_edge_off, const uint32_t* d_edges,
                   const double* d_prob, const int32_t* d_locus_reads, ModelParams P, double* d_M, uint32_t* d_len,
                   uint32_t* d_up, uint32_t total_knodes, uint32_t* d_path, uint32_t* d_path_len,
                   uint32_t max_locus_knodes, uint32_t max_locus_edges, const uint8_t* d_needs_mean,
                   const uint32_t* d_locus_unit_off, const uint32_t* d_unit_start, const uint32_t* d_unit_nodes,
                   float mean_run_len, cudaStream_t st, const uint32_t* d_locus_level_off, const uint32_t* d_level_start,
                   const uint32_t* d_level_nodes, const uint32_t* d_level_singles, uint32_t* h_path, uint32_t* h_path_len,
                   uint32_t* h_done, const double* d_thresh) {
    if (!n_loci) return false;
    {   // default: level-parallel kernel (any of the older switches selects the older kernels)
        static const bool levels_on = [] {
            const char* e = getenv("DRPRG_MLPATH_LEVELS");
            if (e) return atoi(e) != 0;
            return !getenv("DRPRG_MLPATH_UNITS") && !getenv("DRPRG_MLPATH_GENERIC");
        }();
        const size_t lvl_smem = ((size_t)max_locus_knodes + 1) * REC + (size_t)(max_locus_edges + 1) * 4 + (size_t)max_locus_knodes * 12 + 32;
        if (levels_on && d_level_nodes && d_level_singles && d_needs_mean && P.window >= 2 && P.window <= 128 && lvl_smem <= 220u * 1024u) {
            static size_t configured = 0;
            if (lvl_smem > configured) {
                cudaFuncSetAttribute(mlpath_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lvl_smem);
                configured = lvl_smem;
            }
            MlUnitsDev L{d_locus_level_off, d_level_start, d_level_nodes};
            upload_rcp_table();
            const bool streamed = h_path && h_path_len && h_done;  // results straight into host-mapped memory, locus by locus
            mlpath_level_kernel<<<n_loci, ML_LEVEL_THREADS, lvl_smem, st>>>(
                n_loci, d_knode_base, d_edge_off, d_edges, d_prob, d_locus_reads, d_needs_mean, L, d_level_singles, P,
                streamed ? h_path : d_path, streamed ? h_path_len : d_path_len, max_locus_knodes, max_locus_edges,
                streamed ? h_done : nullptr, d_thresh);
            ++g_launches;
            return streamed;
        }
    }
    if (d_thresh) {  // the older kernels take the threshold by value: one small read-back
        double t = P.thresh;
        cudaMemcpyAsync(&t, d_thresh, sizeof t, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        P.thresh = t;
    }
    // The run-parallel kernel pays ~1.5k cycles of per-unit overhead (warp syncs, ~50 loads per lane): it wins when
    // runs of single-successor nodes are long (sparse panels) and loses on bubble-dense graphs (the benchmark panel:
    // 32 % of the nodes have a choice, mean run 2.6 nodes: 1.23 ms vs 0.78 ms for the chain kernel), so among the older
    // kernels it is chosen by run length; the level-parallel kernel above is the default.
    static const char* force_units = getenv("DRPRG_MLPATH_UNITS");
    const bool want_units = force_units ? atoi(force_units) != 0 : mean_run_len >= 8.0f;
    if (want_units) {
        const size_t unit_smem = ((size_t)max_locus_knodes + 1) * REC + (size_t)(max_locus_edges + 1) * 4 + 8 + 32 * 20;
        static const bool no_units = getenv("DRPRG_MLPATH_GENERIC") != nullptr;
        if (P.window >= 2 && P.window <= 127 && unit_smem <= 220u * 1024u && !no_units && d_needs_mean && d_unit_nodes) {
            static size_t configured = 0;
            if (unit_smem > configured) {
                cudaFuncSetAttribute(mlpath_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)unit_smem);
                configured = unit_smem;
            }
            MlUnitsDev U{d_locus_unit_off, d_unit_start, d_unit_nodes};
            mlpath_unit_kernel<<<n_loci, 32, unit_smem, st>>>(n_loci, d_knode_base, d_edge_off, d_edges, d_prob, d_locus_reads,
                                                             d_needs_mean, U, P, d_path, d_path_len, max_locus_knodes,
                                                             max_locus_edges);
            ++g_launches;
            return false;
        }
    }
    // shared memory: per k-mer node sum, mean, score (f64), length, LV lifting pointers, edge offset (u32);
    // per edge one u32.  Up to the 227 KB a CTA may own; larger loci fall back to global memory.
    int LV = 1;
    while ((1u << LV) <= P.window && LV < LV_MAX) ++LV;
    LV = LV <= 7 ? 7 : LV_MAX;  // the two instantiated level counts
    const size_t per_node = 24 + 4 + 4 * (size_t)LV + 4;
    const size_t budget = 220u * 1024u;
    uint32_t smem_nodes = (max_locus_knodes + 3) & ~1u, smem_edges = max_locus_edges + 2;
    if ((size_t)smem_nodes * per_node + 16 + 4 * (size_t)smem_edges > budget) {  // biggest locus does not fit: size for the rest
        smem_nodes = (uint32_t)((budget / 2) / per_node) & ~1u;
        smem_edges = (uint32_t)((budget / 2) / 4);
    }
    const size_t smem = (size_t)smem_nodes * per_node + 16 + 4 * (size_t)smem_edges;
    {
        const size_t rec_smem = ((size_t)max_locus_knodes + 1) * REC + (size_t)(max_locus_edges + 1) * 4;
        static const bool force_generic = getenv("DRPRG_MLPATH_GENERIC") != nullptr;
        if (P.window >= 2 && P.window <= 128 && rec_smem <= budget && !force_generic && d_needs_mean) {
            static size_t configured = 0;
            if (rec_smem > configured) {
                cudaFuncSetAttribute(mlpath_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rec_smem);
                configured = rec_smem;
            }
            mlpath_rec_kernel<<<n_loci, 32, rec_smem, st>>>(n_loci, d_knode_base, d_edge_off, d_edges, d_prob, d_locus_reads,
                                                           d_needs_mean, P, d_path, d_path_len, max_locus_knodes);
            ++g_launches;
            return false;
        }
    }
    auto go = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<n_loci, 32, smem, st>>>(n_loci, d_knode_base, d_edge_off, d_edges, d_prob, d_locus_reads, P, d_M, d_len, d_up,
                                         total_knodes, d_path, d_path_len, smem_nodes, smem_edges);
    };
    switch (LV) {  // levels needed for the window (pandora's default 100 -> 7)
        case 7: go(mlpath_kernel<7>); break;
        case 1: case 2: case 3: case 4: case 5: case 6: go(mlpath_kernel<7>); break;
        default: go(mlpath_kernel<LV_MAX>); break;
    }
    ++g_launches;
    return false;
}

// ============================================================================================
// S8 : per-allele statistics and genotype likelihoods
// ============================================================================================
__global__ void allele_stats_kernel(const int32_t* __restrict__ cov, DevGenotype G, uint32_t min_kmer_covg) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= G.n_alleles) return;
    const uint32_t b = G.allele_off[a], e = G.allele_off[a + 1], n = e - b;
    uint32_t sf = 0, sr = 0, gaps = 0;
    for (uint32_t i = b; i < e; ++i) {
        const uint32_t g = G.allele_kn[i];
        const uint32_t f = cov_sat(cov[2 * g]), r = cov_sat(cov[2 * g + 1]);
        sf += f;
        sr += r;
        if (f + r < min_kmer_covg) ++gaps;
    }
    // integer median by rank selection (n is a handful of k-mers): no scratch memory needed
    uint32_t med[2] = {0, 0};
    if (n) {
        const uint32_t r_hi = n / 2, r_lo = (n % 2) ? n / 2 : n / 2 - 1;
        for (int s = 0; s < 2; ++s) {
            uint32_t v_lo = 0, v_hi = 0;
            for (uint32_t i = b; i < e; ++i) {
                const uint32_t vi = cov_sat(cov[2 * G.allele_kn[i] + s]);
                uint32_t less = 0, leq = 0;
                for (uint32_t j = b; j < e; ++j) {
                    const uint32_t vj = cov_sat(cov[2 * G.allele_kn[j] + s]);
                    less += vj < vi;
                    leq += vj <= vi;
                }
                if (less <= r_lo && r_lo < leq) v_lo = vi;
                if (less <= r_hi && r_hi < leq) v_hi = vi;
            }
            med[s] = (n % 2) ? v_hi : (v_lo + v_hi) / 2;
        }
    }
    G.sum_fwd[a] = sf;
    G.sum_rev[a] = sr;
    G.mean_fwd[a] = n ? sf / n : 0;
    G.mean_rev[a] = n ? sr / n : 0;
    G.med_fwd[a] = med[0];
    G.med_rev[a] = med[1];
    G.gaps[a] = n ? (double)gaps / (double)n : 0.0;
}

__global__ void genotype_kernel(DevGenotype G, ModelParams P) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.n_records) return;
    const uint32_t b = G.rec_off[r], e = G.rec_off[r + 1];
    const double E = (double)P.exp_depth;
    double total = 0.0;
    for (uint32_t a = b; a < e; ++a) total += (double)G.mean_fwd[a] + (double)G.mean_rev[a];
    const double lnE = log(E), lnerr = log(P.gt_err), ln1m = log(1.0 - exp(-E));
    uint32_t best = b;
    for (uint32_t a = b; a < e; ++a) {
        const double c = (double)G.mean_fwd[a] + (double)G.mean_rev[a];
        const double g = G.gaps[a];
        const double L = -E + c * lnE - lgamma(c + 1.0) + (total - c) * lnerr - E * g + ln1m * (1.0 - g);
        G.lik[a] = L;
        if (L > G.lik[best]) best = a;
    }
    double second = -INFINITY;
    for (uint32_t a = b; a < e; ++a)
        if (a != best && G.lik[a] > second) second = G.lik[a];
    const double conf = (e - b > 1) ? fabs(G.lik[best] - second) : 0.0;
    G.gt_conf[r] = conf;
    G.gt[r] = (conf >= P.gt_conf) ? (int32_t)(best - b) : -1;
}

void launch_genotype(const int32_t* d_cov, const DevGenotype& G, ModelParams P, cudaStream_t st) {
    if (!G.n_records) return;
    allele_stats_kernel<<<(G.n_alleles + 127) / 128, 128, 0, st>>>(d_cov, G, P.min_kmer_covg);
    genotype_kernel<<<(G.n_records + 127) / 128, 128, 0, st>>>(G, P);
    g_launches += 2;
}

}  // namespace drprg
