// exact_kernel.cuh — K4: exhaustive genotype calling over known haplotypes (mchap call-exact).
//
// Reference restated (paths relative to the reference repository):
//   calling/exact.py:17-61 _call_posterior_mode, 64-105 _genotype_support_log_joint,
//   108-153 _posterior_allele_frequencies, 156-249 posterior_mode, 252-292 genotype_likelihoods,
//   295-329 genotype_posteriors, 332-369 posterior_allele_frequencies;
//   calling/prior.py:116-179; jitutils.py:113-146 (increment), 253-318 (rank / unrank).
//
// Design: one CTA per (locus, sample) item.
//   1. the read x haplotype table t[r][h] = (prod_j reads[r,j,hap[h,j]], gaps skipped) / ploidy is
//      built once in shared memory (a gather-product over positions: not a GEMM, stays SIMT);
//   2. the G = C(H+P-1, P) genotypes are split into contiguous VCF-order chunks, one per thread:
//      the thread unranks its first genotype and then walks with increment_genotype; the
//      log-likelihood of a genotype sums log(sum_k t[r][g_k]) * count_r over reads IN READ ORDER,
//      exactly the reference's operation order, so values differ only by log() ULPs;
//   3. mode (first maximum, strict >), normaliser (log-sum-exp) and allele statistics are block
//      reductions; log joints are parked in a per-CTA global scratch row for the second pass.
#pragma once
#include "common.cuh"

namespace mchb {

struct ExactArgs {
    const mchb_call_item *items;
    int32_t n_items;
    const double *reads;
    const int64_t *counts;      // may be null
    const int8_t *haplotypes;
    const double *freqs;        // may be null
    int32_t mode;               // 0: posterior_mode (fused), 1: genotype_likelihoods (float32 array)
    int64_t *out_alleles;       // [n_items, pstride]
    int32_t pstride;
    double *out_stats;          // [n_items, 4]: mode_llk, mode_prob, support_prob, total_ljoint
    double *out_freqs;          // per item at hap_out_off: posterior mean allele frequencies
    double *out_occur;          // per item at hap_out_off: posterior occurrence
    float *out_gl;              // per item at gl_off (mode 1)
    double *scratch;            // [gridDim.x, scratch_stride] log joints
    int64_t scratch_stride;
    int32_t *work_counter;
    mchb_item_result *results;
    int32_t umax, hmax, pmax;   // shared memory geometry
};

// up to 16 allele indices (< 256) packed in two words
struct Geno {
    uint64_t lo, hi;
    __device__ __forceinline__ int get(int k) const { return (int)(((k < 8 ? lo : hi) >> (8 * (k & 7))) & 255u); }
    __device__ __forceinline__ void set(int k, int v) {
        uint64_t m = ~(255ull << (8 * (k & 7)));
        uint64_t x = (uint64_t)v << (8 * (k & 7));
        if (k < 8) lo = (lo & m) | x;
        else hi = (hi & m) | x;
    }
};

// jitutils.py:113-146 on a sorted packed genotype
__device__ __forceinline__ void geno_increment(Geno &g, int P) {
    if (P == 1) {
        g.set(0, g.get(0) + 1);
        return;
    }
    int prev = g.get(0);
    for (int i = 1; i < P; i++) {
        int al = g.get(i);
        if (al == prev) continue;
        g.set(i - 1, g.get(i - 1) + 1);
        for (int m = 0; m < i - 1; m++) g.set(m, 0);
        return;
    }
    g.set(P - 1, g.get(P - 1) + 1);
    for (int m = 0; m < P - 1; m++) g.set(m, 0);
}

// jitutils.py:279-318
__device__ inline Geno geno_unrank(long long index, int P) {
    Geno g;
    g.lo = 0;
    g.hi = 0;
    long long remainder = index;
    for (int it = 0; it < P; it++) {
        int p = P - it;
        long long a = -1, nw = 0, prev = 0;
        while (nw <= remainder) {
            a += 1;
            prev = nw;
            nw = comb_with_replacement(a, p);
        }
        a -= 1;
        remainder -= prev;
        g.set(p - 1, (int)a);
    }
    return g;
}

// jitutils.py:253-276
__device__ inline long long geno_rank(const Geno &g, int P) {
    long long index = 0;
    for (int i = 0; i < P; i++) index += comb_with_replacement(g.get(i), i + 1);
    return index;
}

// calling/prior.py:116-179 for a packed genotype.  lgA[a] = lgamma(alpha_a) and
// lgDA[a*(P+1)+d] = lgamma(d + alpha_a) are per-item tables in shared memory (only when
// inbreeding > 0); freqs may be null.
__device__ __forceinline__ double exact_log_prior(const Geno &g, int P, int H, double inbreeding, const double *freqs,
                                                  const double *lgA, const double *lgDA, double lg_left,
                                                  double log_H) {
    double acc = 0.0;  // ln_denom (null prior) or prod (Dirichlet-multinomial)
    const bool null_prior = inbreeding == 0.0;
    for (int i = 0; i < P; i++) {
        const int ai = g.get(i);
        int cntv = 0;
        bool first = true;
        for (int k = 0; k < P; k++) {
            bool eq = g.get(k) == ai;
            cntv += eq;
            first = first && !(eq && k < i);
        }
        const int dose = first ? cntv : 0;
        if (null_prior) acc += LGAMMA_INT[dose + 1];
        else if (dose > 0) acc += lgDA[ai * (P + 1) + dose] - (LGAMMA_INT[dose + 1] + lgA[ai]);
    }
    if (null_prior) {
        const double ln_perms = LGAMMA_INT[P + 1] - acc;
        if (!freqs) return ln_perms - (double)P * log_H;
        double prod = 1.0;
        for (int i = 0; i < P; i++) prod *= freqs[g.get(i)];
        return ln_perms + log(prod);
    }
    return lg_left + acc;
}

struct ModeRec {
    double ljoint, llk;
    long long idx;
};

__device__ __forceinline__ ModeRec mode_better(const ModeRec &a, const ModeRec &b) {
    // strict > in index order == largest value, smallest index among equals; NaN never wins
    if (b.ljoint > a.ljoint || (b.ljoint == a.ljoint && b.idx < a.idx)) return b;
    return a;
}

__global__ void __launch_bounds__(128) exact_kernel(const __grid_constant__ ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5;
    // shared layout
    double *tab = reinterpret_cast<double *>(smem_raw);             // [umax][hmax]
    double *cnt = tab + (size_t)a.umax * a.hmax;                     // [umax]
    double *lgA = cnt + a.umax;                                      // [hmax]
    double *lgDA = lgA + a.hmax;                                     // [hmax][pmax+1]
    double *sfreq = lgDA + (size_t)a.hmax * (a.pmax + 1);            // [hmax]
    double *soccur = sfreq + a.hmax;                                 // [hmax]
    double *red_d = soccur + a.hmax;                                 // [8] reductions
    ModeRec *red_m = reinterpret_cast<ModeRec *>(red_d + 8);         // [4]
    __shared__ int s_item;
    double *scratch = a.scratch ? a.scratch + (size_t)blockIdx.x * a.scratch_stride : nullptr;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;
        const mchb_call_item it = a.items[item];
        const int U = it.n_reads, N = it.n_pos, A = it.max_allele, P = it.ploidy, H = it.n_haps;
        const double *R = a.reads + it.reads_off;
        const int8_t *haps = a.haplotypes + it.haps_off;
        const double *freqs = (a.freqs && it.freqs_off >= 0) ? a.freqs + it.freqs_off : nullptr;
        const bool has_prior = !isnan(it.inbreeding);
        const double inbreeding = it.inbreeding;
        const double dP = (double)P;
        const long long G = comb_exact((long long)H + P - 1, P);

        // ---- 1. table t[r][h] (likelihood.py:48-60 per haplotype) and counts
        for (int i = tid; i < U * H; i += nthr) {
            const int r = i / H, h = i - r * H;
            double prod = 1.0;
            for (int j = 0; j < N; j++) {
                double v = __ldg(R + ((size_t)r * N + j) * A + haps[h * N + j]);
                if (!isnan(v)) prod *= v;
            }
            tab[r * H + h] = prod / dP;
        }
        for (int r = tid; r < U; r += nthr) cnt[r] = a.counts ? (double)__ldg(a.counts + it.counts_off + r) : 1.0;
        // ---- prior tables
        double lg_left = 0.0;
        const double log_H = log((double)H);
        if (has_prior && inbreeding != 0.0) {
            const double scale = (1.0 - inbreeding) / inbreeding;
            const double alpha_const = (1.0 / (double)H) * scale;
            for (int i = tid; i < H * (P + 1); i += nthr) {
                const int al = i / (P + 1), d = i - al * (P + 1);
                const double alpha = freqs ? freqs[al] * scale : alpha_const;
                if (d == 0) lgA[al] = lgamma(alpha);
                else lgDA[al * (P + 1) + d] = lgamma((double)d + alpha);
            }
            double sum_alphas = 0.0;
            if (freqs) for (int al = 0; al < H; al++) sum_alphas += freqs[al] * scale;
            else sum_alphas = alpha_const * (double)H;
            lg_left = (LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma(dP + sum_alphas);
        }
        for (int i = tid; i < H; i += nthr) {
            sfreq[i] = 0.0;
            soccur[i] = 0.0;
        }
        __syncthreads();

        // ---- 2. enumerate this thread's chunk of genotypes in VCF order
        const long long chunk = (G + nthr - 1) / nthr;
        const long long g0 = (long long)tid * chunk;
        const long long g1 = g0 + chunk < G ? g0 + chunk : G;
        ModeRec best;
        best.ljoint = -INFINITY;
        best.llk = -INFINITY;
        best.idx = 0x7fffffffffffffffLL;
        double total = -INFINITY;
        if (g0 < G) {
            Geno g = geno_unrank(g0, P);
            for (long long gi = g0; gi < g1; gi++) {
                double llk = 0.0;
                for (int r = 0; r < U; r++) {
                    const double *row = tab + r * H;
                    double rp = 0.0;
                    for (int k = 0; k < P; k++) rp += row[g.get(k)];
                    llk += log(rp) * cnt[r];
                }
                if (a.mode == 1) {
                    a.out_gl[it.gl_off + gi] = (float)llk;
                } else {
                    double lpr = 0.0;
                    if (has_prior) lpr = exact_log_prior(g, P, H, inbreeding, freqs, lgA, lgDA, lg_left, log_H);
                    const double ljoint = llk + lpr;
                    if (ljoint > best.ljoint) {
                        best.ljoint = ljoint;
                        best.llk = llk;
                        best.idx = gi;
                    }
                    total = add_log_prob(total, ljoint);
                    scratch[gi] = ljoint;
                }
                geno_increment(g, P);
            }
        }
        if (a.mode == 1) {
            if (tid == 0) {
                mchb_item_result r;
                r.status = MCHB_ITEM_OK;
                r.n_het = 0;
                r.rng_words = 0;
                r.llk_evals = G;
                a.results[item] = r;
            }
            continue;
        }
        // ---- 3. block reductions: first maximum and log-sum-exp
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            ModeRec o;
            o.ljoint = __shfl_xor_sync(MCHB_FULL, best.ljoint, m);
            o.llk = __shfl_xor_sync(MCHB_FULL, best.llk, m);
            o.idx = __shfl_xor_sync(MCHB_FULL, best.idx, m);
            best = mode_better(best, o);
            total = add_log_prob(total, __shfl_xor_sync(MCHB_FULL, total, m));
        }
        if (lane == 0) {
            red_m[warp] = best;
            red_d[warp] = total;
        }
        __syncthreads();
        const int nwarp = nthr >> 5;
        best = red_m[0];
        total = red_d[0];
        for (int w = 1; w < nwarp; w++) {
            best = mode_better(best, red_m[w]);
            total = add_log_prob(total, red_d[w]);
        }
        if (!(best.ljoint > -INFINITY)) {  // nothing ever exceeded -inf: the reference keeps index 0
            best.idx = 0;
            best.llk = -INFINITY;
            best.ljoint = -INFINITY;
        }
        // ---- 4. mode alleles, support probability (thread 0), then allele statistics (all)
        if (tid == 0) {
            const Geno mode = geno_unrank(best.idx, P);
            for (int k = 0; k < P; k++) a.out_alleles[(size_t)item * a.pstride + k] = mode.get(k);
            for (int k = P; k < a.pstride; k++) a.out_alleles[(size_t)item * a.pstride + k] = -2;
            // exact.py:64-105: all dosage variants of the mode's haplotype set, in
            // combinations_with_replacement order
            int support[MCHB_MAX_PLOIDY];
            int ns = 0;
            for (int k = 0; k < P; k++)
                if (k == 0 || mode.get(k) != mode.get(k - 1)) support[ns++] = mode.get(k);
            const int rem = P - ns;
            int idx[MCHB_MAX_PLOIDY];
            for (int k = 0; k < rem; k++) idx[k] = 0;
            double support_ljoint = -INFINITY;
            for (;;) {
                // merge support + chosen extras into a sorted genotype (counting sort over support)
                Geno t;
                t.lo = 0;
                t.hi = 0;
                int pos = 0;
                for (int s = 0; s < ns; s++) {
                    int c = 1;
                    for (int k = 0; k < rem; k++) c += (idx[k] == s);
                    for (int k = 0; k < c; k++) t.set(pos++, support[s]);
                }
                support_ljoint = add_log_prob(support_ljoint, scratch[geno_rank(t, P)]);
                int i = rem - 1;
                while (i >= 0 && idx[i] == ns - 1) i--;
                if (i < 0) break;
                const int v = idx[i] + 1;
                for (int k = i; k < rem; k++) idx[k] = v;
            }
            double *st = a.out_stats + (size_t)item * 4;
            st[0] = best.llk;
            st[1] = exp(best.ljoint - total);
            st[2] = exp(support_ljoint - total);
            st[3] = total;
        }
        __syncthreads();  // scratch complete and visible
        if (g0 < G) {
            Geno g = geno_unrank(g0, P);
            for (long long gi = g0; gi < g1; gi++) {
                const double prob = exp(scratch[gi] - total);
                for (int k = 0; k < P; k++) {
                    const int al = g.get(k);
                    atomicAdd(&sfreq[al], prob);
                    if (k == 0 || al != g.get(k - 1)) atomicAdd(&soccur[al], prob);
                }
                geno_increment(g, P);
            }
        }
        __syncthreads();
        for (int i = tid; i < H; i += nthr) {
            a.out_freqs[it.hap_out_off + i] = sfreq[i] / dP;
            a.out_occur[it.hap_out_off + i] = soccur[i];
        }
        if (tid == 0) {
            mchb_item_result r;
            r.status = MCHB_ITEM_OK;
            r.n_het = 0;
            r.rng_words = 0;
            r.llk_evals = G;
            a.results[item] = r;
        }
    }
}

// calling/exact.py:295-329 genotype_posteriors + 332-369 posterior_allele_frequencies for one
// item per CTA from a stored llk array (float32 as the reference's CLI branch stores it, or
// float64).  posteriors[i] = exp(x_i - logsumexp(x)), x_i = STORED(llk_i + lprior_i) where the
// store rounds to float32 when the input was float32 (numba keeps the array dtype, exact.py:311).
struct PosteriorArgs {
    const mchb_call_item *items;
    int32_t n_items;
    const double *freqs;
    const float *llk32;     // one of llk32 / llk64 is set; per item at gl_off
    const double *llk64;
    double *out_gp;         // per item at gl_off
    double *out_freqs;      // per item at hap_out_off (may be null)
    double *out_counts;
    double *out_occur;
    int32_t hmax, pmax;
};

__global__ void __launch_bounds__(128) posterior_kernel(const __grid_constant__ PosteriorArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    double *lgA = reinterpret_cast<double *>(smem_raw);
    double *lgDA = lgA + a.hmax;
    double *sfreq = lgDA + (size_t)a.hmax * (a.pmax + 1);
    double *soccur = sfreq + a.hmax;
    double *red_d = soccur + a.hmax;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        __syncthreads();
        const mchb_call_item it = a.items[item];
        const int P = it.ploidy, H = it.n_haps;
        const double *freqs = (a.freqs && it.freqs_off >= 0) ? a.freqs + it.freqs_off : nullptr;
        const bool has_prior = !isnan(it.inbreeding);
        const double inbreeding = it.inbreeding;
        const double dP = (double)P;
        const long long G = comb_exact((long long)H + P - 1, P);
        double lg_left = 0.0;
        const double log_H = log((double)H);
        if (has_prior && inbreeding != 0.0) {
            const double scale = (1.0 - inbreeding) / inbreeding;
            const double alpha_const = (1.0 / (double)H) * scale;
            for (int i = tid; i < H * (P + 1); i += nthr) {
                const int al = i / (P + 1), d = i - al * (P + 1);
                const double alpha = freqs ? freqs[al] * scale : alpha_const;
                if (d == 0) lgA[al] = lgamma(alpha);
                else lgDA[al * (P + 1) + d] = lgamma((double)d + alpha);
            }
            double sum_alphas = 0.0;
            if (freqs) for (int al = 0; al < H; al++) sum_alphas += freqs[al] * scale;
            else sum_alphas = alpha_const * (double)H;
            lg_left = (LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma(dP + sum_alphas);
        }
        for (int i = tid; i < H; i += nthr) {
            sfreq[i] = 0.0;
            soccur[i] = 0.0;
        }
        __syncthreads();
        const long long chunk = (G + nthr - 1) / nthr;
        const long long g0 = (long long)tid * chunk;
        const long long g1 = g0 + chunk < G ? g0 + chunk : G;
        double *gp = a.out_gp + it.gl_off;
        double total = -INFINITY;
        if (g0 < G) {
            Geno g = geno_unrank(g0, P);
            for (long long gi = g0; gi < g1; gi++) {
                double lpr = 0.0;
                if (has_prior) lpr = exact_log_prior(g, P, H, inbreeding, freqs, lgA, lgDA, lg_left, log_H);
                double x;
                if (a.llk32) x = (double)(float)((double)a.llk32[it.gl_off + gi] + lpr);
                else x = a.llk64[it.gl_off + gi] + lpr;
                gp[gi] = x;
                total = add_log_prob(total, x);
                geno_increment(g, P);
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) total = add_log_prob(total, __shfl_xor_sync(MCHB_FULL, total, m));
        if (lane == 0) red_d[warp] = total;
        __syncthreads();
        total = red_d[0];
        for (int w = 1; w < (nthr >> 5); w++) total = add_log_prob(total, red_d[w]);
        if (g0 < G) {
            Geno g = geno_unrank(g0, P);
            for (long long gi = g0; gi < g1; gi++) {
                const double prob = exp(gp[gi] - total);
                gp[gi] = prob;
                if (a.out_freqs) {
                    for (int k = 0; k < P; k++) {
                        const int al = g.get(k);
                        atomicAdd(&sfreq[al], prob);
                        if (k == 0 || al != g.get(k - 1)) atomicAdd(&soccur[al], prob);
                    }
                }
                geno_increment(g, P);
            }
        }
        __syncthreads();
        if (a.out_freqs) {
            for (int i = tid; i < H; i += nthr) {
                a.out_freqs[it.hap_out_off + i] = sfreq[i] / dP;
                a.out_counts[it.hap_out_off + i] = sfreq[i];
                a.out_occur[it.hap_out_off + i] = soccur[i];
            }
        }
    }
}

}  // namespace mchb
