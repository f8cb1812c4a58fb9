// Hand-written sm_100a kernels for the map hot path: read sketching + index lookup (S1+S2), hit
// clustering (S3/S4), k-mer coverage (S5), ML path (S7) and genotyping (S8).  These replace the
// per-read and per-locus loops of `pandora map` that drprg launches at
// /root/reference/src/lib.rs:580-642 (argv :594-609, src/predict.rs:288-294); stage semantics
// follow pandora's Seq::minimizer_sketch, add_read_hits, define_clusters, filter_clusters(2),
// add_hits_to_kmergraphs, KmerGraphWithCoverage::find_max_path and SampleInfo.
// This file: S8 (per-allele statistics, likelihoods, GT, GT_CONF).
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "kernels_common.cuh"

namespace drprg {

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
    const int32_t gt_pandora = (conf >= P.gt_conf) ? (int32_t)(best - b) : -1;
    G.gt[r] = gt_pandora;
    if (!G.covg_gt) return;
    // ---- what drprg computes next from this record, fused here.  Everything below is f32 / i32
    // like the Rust code; "None" is NaN.  drprg first nulls the call of a record without depth and GT_CONF 0
    // (src/predict.rs:440-444), and its filters see that call.
    const float NONE = __int_as_float(0x7fc00000);
    const uint32_t na = e - b;
    int32_t tot_f = 0, tot_r = 0;
    for (uint32_t a = b; a < e; ++a) {
        tot_f += (int32_t)G.mean_fwd[a];
        tot_r += (int32_t)G.mean_rev[a];
    }
    const int32_t depth = tot_f + tot_r;
    const int32_t gt = (depth == 0 && (float)conf == 0.0f) ? -1 : gt_pandora;
    auto dp = [&](uint32_t i) { return (int32_t)G.mean_fwd[b + i] + (int32_t)G.mean_rev[b + i]; };
    // Filterer::_covg_for_gt
    G.covg_gt[r] = gt < 0 ? depth : dp((uint32_t)gt);
    // VcfExt::fraction_read_support
    float frs = NONE;
    if (na < 2) frs = 1.0f;
    else if (gt >= 0) {
        const float called = (float)dp((uint32_t)gt);
        int32_t other = 0;
        if (gt > 0) other = dp(0);
        else
            for (uint32_t i = 0; i < na; ++i)
                if ((int32_t)i != gt && dp(i) > other) other = dp(i);
        frs = called / (called + (float)other);  // 0 / 0 = NaN = None
    }
    G.frs[r] = frs;
    // Filterer::has_strand_bias: the ratio it compares with --min-strand-bias
    float sb = NONE;
    if (gt < 0) {
        const float tf = (float)tot_f, tr = (float)tot_r, tt = tf + tr;
        if (tt != 0.0f) sb = fminf(tf, tr) / tt;
    } else {
        const float f = (float)G.mean_fwd[b + gt], rv = (float)G.mean_rev[b + gt], sum = f + rv;
        if (sum != 0.0f) sb = fminf(f, rv) / sum;
    }
    G.sb_ratio[r] = sb;
    // VcfExt::depth_proportions (the PDP tag)
    const float total_depth = (float)depth;
    for (uint32_t i = 0; i < na; ++i) G.pdp[b + i] = depth ? (float)dp(i) / total_depth : NONE;
    // MinorAllele::check_for_minor_alternate: the non-called allele with the largest depth proportion >= maf whose GAPS
    // pass (stable ascending sort walked backwards: among equal proportions the higher allele index comes first)
    int32_t minor = -1;
    if (na >= 2 && depth != 0 && gt >= 0 && (float)G.gaps[b + gt] <= 0.39f) {
        const float called_gaps = (float)G.gaps[b + gt];
        int32_t pick = -1;
        float pick_d = -1.0f;
        for (uint32_t i = 0; i < na; ++i) {
            if ((int32_t)i == gt) continue;
            const float d = (float)dp(i) / total_depth, g = (float)G.gaps[b + i];
            if (d >= P.minor_af && g <= 0.5f && g - called_gaps <= 0.2f && d >= pick_d) {
                pick = (int32_t)i;
                pick_d = d;
            }
        }
        if (pick >= 0) {
            const int32_t c = dp((uint32_t)pick);
            const float f = (float)G.mean_fwd[b + pick], rv = (float)G.mean_rev[b + pick], sum = f + rv;
            const bool bias = sum == 0.0f ? true : fminf(f, rv) / sum < 0.01f;
            if (c >= 3 && !bias) minor = pick;
        }
    }
    G.minor_gt[r] = minor;
}

// ============================================================================================
// VCF text on the device: the per-record lines of pandora_genotyped.vcf
// ============================================================================================
// The CHROM..FORMAT columns of a record never change for a site table, so they sit in device memory as a byte array;
// only the sample column (GT : six integer vectors : GAPS : LIKELIHOOD : GT_CONF) is formatted per sample.  "%g" with
// six significant digits is the host formatter's fast path (genotype_host.cpp::format_g6, pinned against printf by
// tests/test_host_cpu.py) restated with explicit round-to-nearest operations: scale to a six-digit integer, refuse
// anything within 1e-6 of a rounding tie and everything that needs exponent notation.  A refused value raises a flag
// and the host formats that sample's text itself, so the text is byte-identical either way.
__device__ __forceinline__ char* dev_put_uint(char* o, uint32_t v) {
    char tmp[10];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10u);
        v /= 10u;
    } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}

__device__ char* dev_format_g6(double v, char* o, uint32_t& refused) {
    if (v == 0.0) {
        if (signbit(v)) refused = 1u;  // "-0": left to the host
        *o++ = '0';
        return o;
    }
    const double a = fabs(v);
    if (!(a >= 1e-4 && a < 999999.0)) {  // exponent notation, inf, nan
        refused = 1u;
        return o;
    }
    const double P10[11] = {1e-4, 1e-3, 1e-2, 1e-1, 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6};
    const double SC[10] = {1e9, 1e8, 1e7, 1e6, 1e5, 1e4, 1e3, 1e2, 1e1, 1e0};  // 10^(5 - e10)
    int e10 = -4;
    while (a >= P10[e10 + 5]) ++e10;  // a in [10^e10, 10^(e10+1))
    const double x = __dmul_rn(a, SC[e10 + 4]);
    const double fl = floor(x);
    const double frac = __dsub_rn(x, fl);
    if (fabs(__dsub_rn(frac, 0.5)) < 1e-6) {
        refused = 1u;
        return o;
    }
    uint32_t n = (uint32_t)fl + (frac > 0.5 ? 1u : 0u);
    if (n >= 1000000u) {
        n = 100000u;
        ++e10;
        if (e10 > 5) {
            refused = 1u;
            return o;
        }
    }
    char d[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        d[i] = (char)('0' + n % 10u);
        n /= 10u;
    }
    int last = 5;
    while (last > 0 && d[last] == '0') --last;  // significant digits d[0..last]
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
    return o;
}

// one thread per record: the sample column into the record's slot (slot_off: static upper bounds), its length
__global__ void vcf_sample_column_kernel(DevGenotype G, DevVcfText V) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= G.n_records) return;
    const uint32_t b = G.rec_off[r], e = G.rec_off[r + 1];
    char* const o0 = V.slots + V.slot_off[r];
    char* o = o0;
    uint32_t refused = 0;
    const int32_t gt = G.gt[r];
    if (gt < 0) *o++ = '.';
    else o = dev_put_uint(o, (uint32_t)gt);
    const uint32_t* cols[6] = {G.mean_fwd, G.mean_rev, G.med_fwd, G.med_rev, G.sum_fwd, G.sum_rev};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        *o++ = ':';
        for (uint32_t a = b; a < e; ++a) {
            if (a > b) *o++ = ',';
            o = dev_put_uint(o, cols[c][a]);
        }
    }
    *o++ = ':';
    for (uint32_t a = b; a < e; ++a) {
        if (a > b) *o++ = ',';
        o = dev_format_g6(G.gaps[a], o, refused);
    }
    *o++ = ':';
    for (uint32_t a = b; a < e; ++a) {
        if (a > b) *o++ = ',';
        o = dev_format_g6(G.lik[a], o, refused);
    }
    *o++ = ':';
    o = dev_format_g6(G.gt_conf[r], o, refused);
    *o++ = '\n';
    V.line_len[r] = (V.prefix_off[r + 1] - V.prefix_off[r]) + (uint32_t)(o - o0);
    if (refused) atomicOr(V.flags, 1u);
}

// one CTA: exclusive scan of the line lengths -> where every line starts in the text; total length
__global__ void vcf_line_offsets_kernel(uint32_t n, const uint32_t* __restrict__ line_len, uint32_t* __restrict__ out_off,
                                        uint32_t* __restrict__ total /* host-mapped */) {
    __shared__ uint32_t s_part[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const uint32_t i = i0 + tid;
        const uint32_t v = i < n ? line_len[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += t;
        }
        if (lane == 31) s_part[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t p = lane < (blockDim.x >> 5) ? s_part[lane] : 0u, pi = p;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, pi, d);
                if (lane >= (uint32_t)d) pi += t;
            }
            s_part[lane] = pi - p;  // exclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n) out_off[i] = carry + s_part[warp] + incl - v;
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = carry + s_part[warp] + incl;
        __syncthreads();
    }
    if (tid == 0) {
        out_off[n] = s_carry;
        *total = s_carry;
    }
}

// The text goes to host-mapped pinned memory, i.e. over PCIe: every thread assembles 16 consecutive bytes of the OUTPUT
// (finding the record its first byte belongs to by binary search over the line offsets) and stores them with one 128-bit
// write, so a warp emits 512 contiguous bytes — byte stores per record would cross the bus as 32-byte packets.
__global__ void vcf_gather_kernel(uint32_t n, DevVcfText V, const uint32_t* __restrict__ out_off) {
    const uint32_t total = out_off[n];
    const uint32_t o0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    if (o0 >= total) return;
    uint32_t lo = 0, hi = n;  // the record r with out_off[r] <= o0 < out_off[r + 1] (empty lines cannot occur)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (out_off[mid] <= o0) lo = mid;
        else hi = mid;
    }
    uint32_t r = lo;
    uint32_t line0 = out_off[r], line1 = out_off[r + 1];
    uint32_t p0 = V.prefix_off[r], pl = V.prefix_off[r + 1] - p0;
    const char* slot = V.slots + V.slot_off[r];
    union {
        uint4 v;
        char c[16];
    } w;
    w.v = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t o = o0 + k;
        if (o >= total) break;
        while (o >= line1) {  // next record
            ++r;
            line0 = line1;
            line1 = out_off[r + 1];
            p0 = V.prefix_off[r];
            pl = V.prefix_off[r + 1] - p0;
            slot = V.slots + V.slot_off[r];
        }
        const uint32_t i = o - line0;
        w.c[k] = i < pl ? V.prefix[p0 + i] : slot[i - pl];
    }
    *reinterpret_cast<uint4*>(V.text + o0) = w.v;  // text is 16-byte aligned; the buffer has room past `total`
}

// test hook: the device "%g" on arbitrary values (48 bytes per value, its length, and whether the device refused it)
__global__ void format_g6_batch_kernel(const double* __restrict__ v, uint32_t n, char* __restrict__ out, uint8_t* __restrict__ len,
                                       uint8_t* __restrict__ refused) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r = 0;
    char* o0 = out + 48ull * i;
    char* o = dev_format_g6(v[i], o0, r);
    len[i] = (uint8_t)(o - o0);
    refused[i] = (uint8_t)r;
}
void launch_format_g6_batch(const double* d_v, uint32_t n, char* d_out, uint8_t* d_len, uint8_t* d_refused, cudaStream_t st) {
    if (!n) return;
    format_g6_batch_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_v, n, d_out, d_len, d_refused);
    ++g_launches;
}

void launch_vcf_text(const DevGenotype& G, const DevVcfText& V, cudaStream_t st) {
    if (!G.n_records) return;
    vcf_sample_column_kernel<<<(G.n_records + 127) / 128, 128, 0, st>>>(G, V);
    vcf_line_offsets_kernel<<<1, 1024, 0, st>>>(G.n_records, V.line_len, V.out_off, V.total);
    // one thread per 16 bytes of text; the grid is sized for the upper bound of the text (threads past the end return)
    const uint32_t max_chunks = (V.text_bound + 15u) / 16u;
    vcf_gather_kernel<<<(max_chunks + 255) / 256, 256, 0, st>>>(G.n_records, V, V.out_off);
    g_launches += 3;
}

void launch_genotype(const int32_t* d_cov, const DevGenotype& G, ModelParams P, cudaStream_t st) {
    if (!G.n_records) return;
    allele_stats_kernel<<<(G.n_alleles + 127) / 128, 128, 0, st>>>(d_cov, G, P.min_kmer_covg);
    genotype_kernel<<<(G.n_records + 127) / 128, 128, 0, st>>>(G, P);
    g_launches += 2;
}

void launch_genotype_rows(const DevGenotype& G, ModelParams P, cudaStream_t st) {
    if (!G.n_records) return;
    genotype_kernel<<<(G.n_records + 127) / 128, 128, 0, st>>>(G, P);
    ++g_launches;
}

}  // namespace drprg
