// exact_kernel.cuh — K4: exhaustive genotype calling over known haplotypes (mchap call-exact).
//
// Reference restated (paths relative to the reference repository):
//   calling/exact.py:17-61 _call_posterior_mode, 64-105 _genotype_support_log_joint,
//   108-153 _posterior_allele_frequencies, 156-249 posterior_mode, 252-292 genotype_likelihoods,
//   295-329 genotype_posteriors, 332-369 posterior_allele_frequencies;
//   calling/prior.py:116-179; jitutils.py:113-146 (increment), 253-318 (rank / unrank).
//
// Design: one CTA per (locus, sample) item, one thread per contiguous VCF-order chunk of the
// G = C(H+P-1, P) genotypes.
//   1. the read x haplotype table t[h][r] = (prod_j reads[r,j,hap[h,j]], gaps skipped) / ploidy is
//      built once in shared memory, haplotype-major with an odd row stride: a thread keeps one row
//      pointer per genotype slot in registers and the inner loop is P shared loads with immediate
//      offsets and P-1 adds per read (a gather-product / gather-sum: not a GEMM, stays SIMT);
//   2. the per-read probability rp = sum_k t[g_k][r] keeps the reference's left fold, and the
//      log-likelihood sum_r c_r log(rp_r) is evaluated as the log of a product: the mantissas of the
//      rp_r are multiplied and their exponents added as integers, with ONE log per genotype instead
//      of one per read.  That is the same real number; its rounding error (<= n_reads * 2^-53
//      relative on the product, nothing on the exponents) is below the accumulated rounding of the
//      reference's own sum.  Reads with a count above 16, zero / denormal / non-finite rp take the
//      reference's form (log(rp) * count);
//   3. a genotype moves to its VCF-order successor in registers; a thread's first genotype is
//      unranked from a binomial table C(n+p-1, p) that the CTA keeps in shared memory;
//   4. every reduction is in a fixed order, so results are reproducible bit for bit from run to
//      run: per-thread partials over the contiguous chunk, combined in thread order.  The
//      normaliser is M + log(sum_g exp(ljoint_g - M)) with M the largest log joint, and the allele
//      statistics accumulate exp(ljoint_g - M) in per-thread shared-memory rows;
//   5. log joints are parked in a per-CTA global scratch row between the passes when the rows of
//      the whole grid fit the caller's budget, otherwise they are evaluated again in the second pass
//      (the reference's own low-memory scheme, exact.py:108-153).
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
    double *scratch;            // [gridDim.x, scratch_stride] log joints, row layout [step][thread]; null: evaluate twice
    int64_t scratch_stride;
    int32_t *work_counter;
    mchb_item_result *results;
    int32_t umax, hmax, pmax;   // shared memory geometry
    int32_t part_threads;       // threads that own a partial row of allele statistics (power of two <= 128)
};

// fdlibm's split of ln 2: E * LN2_HI is exact for |E| < 2^20
#define MCHB_LN2_HI 6.93147180369123816490e-01
#define MCHB_LN2_LO 1.90821492927058770002e-10
// read counts up to this take the product form (count factors of the mantissa); larger ones log(rp) * count
#define MCHB_EXACT_POW_MAX 16

// Shared-memory carve-up of the exhaustive kernels (doubles unless noted), host mirror: exact_smem().
struct ExactSmem {
    double *tab;      // [hmax][us]     read x haplotype table, haplotype-major
    double *lgA;      // [hmax]         lgamma(alpha_a)
    double *lgDA;     // [hmax][pmax+1] Dirichlet-multinomial term of allele a at dosage d
    double *fr;       // [hmax]         prior allele frequencies (staged)
    double *part;     // [part_threads][hs][2] partial allele statistics
    double *red;      // [32]: [0..7] per-warp sums, [8..] per-warp mode records / maxima
    long long *cwr;   // [(hmax+1)][pmax+1] C(n+p-1, p)
    int *cnt;         // [umax] read counts in column order
    int *col;         // [umax] table column of read r
    int us, hs, ps;   // row strides (doubles): table, -, partial rows (odd: rows of a warp start in different banks)
};

__device__ __forceinline__ ExactSmem exact_carve(unsigned char *raw, int umax, int hmax, int pmax, int part_threads) {
    ExactSmem s;
    s.us = umax | 1;
    s.hs = hmax | 1;
    s.ps = (2 * hmax) | 1;
    s.tab = reinterpret_cast<double *>(raw);
    s.lgA = s.tab + (size_t)hmax * s.us;
    s.lgDA = s.lgA + hmax;
    s.fr = s.lgDA + (size_t)hmax * (pmax + 1);
    s.part = s.fr + hmax;
    s.red = s.part + (size_t)part_threads * s.ps;
    s.cwr = reinterpret_cast<long long *>(s.red + 32);
    s.cnt = reinterpret_cast<int *>(s.cwr + (size_t)(hmax + 1) * (pmax + 1));
    s.col = s.cnt + umax;
    return s;
}

// C(n+p-1, p) for n = 0..H, p = 0..P by Pascal's rule (jitutils.py:213-250 conventions:
// cwr(n, 0) = 1 for n >= 1, cwr(0, p) = 0 including p = 0).  Values beyond int64 saturate.
__device__ __forceinline__ void exact_fill_cwr(long long *cwr, int H, int P, int pstride, int tid, int nthr) {
    // column p needs column p-1: P is small, so one thread per row sweeps p with a barrier per column
    for (int p = 0; p <= P; p++) {
        if (p == 0) {
            for (int n = tid; n <= H; n += nthr) cwr[n * pstride] = n == 0 ? 0 : 1;
        } else if (tid == 0) {
            long long acc = 0;
            cwr[p] = 0;
            for (int n = 1; n <= H; n++) {
                // cwr(n, p) = cwr(n-1, p) + cwr(n, p-1)
                const long long add = cwr[n * pstride + p - 1];
                acc = (acc > 0x7fffffffffffffffLL - add) ? 0x7fffffffffffffffLL : acc + add;
                cwr[n * pstride + p] = acc;
            }
        }
        __syncthreads();
    }
}

// Genotype of PM register slots (slots >= P unused), the row pointer of every slot and the
// successor / unrank operations on them.
template <int PM>
struct GenoR {
    int g[PM];

    // jitutils.py:279-318 with the shared binomial table
    __device__ __forceinline__ void unrank(long long index, int P, const long long *cwr, int pstride, int H) {
        long long rem = index;
#pragma unroll
        for (int k = PM - 1; k >= 0; k--) {
            if (k < P) {
                // largest a with cwr(a, k+1) <= rem
                int a = 0;
                while (a < H && cwr[(a + 1) * pstride + k + 1] <= rem) a++;
                rem -= cwr[a * pstride + k + 1];
                g[k] = a;
            } else {
                g[k] = 0;
            }
        }
    }
    // jitutils.py:113-146
    __device__ __forceinline__ void increment(int P) {
        int i = P;  // all equal: bump the last slot
#pragma unroll
        for (int k = PM - 1; k >= 1; k--)
            if (k < P && g[k] != g[k - 1]) i = k;
#pragma unroll
        for (int k = 0; k < PM; k++) {
            if (k == i - 1) g[k] += 1;
            else if (k < i - 1) g[k] = 0;
        }
    }
    // jitutils.py:253-276
    __device__ __forceinline__ long long rank(int P, const long long *cwr, int pstride) const {
        long long index = 0;
#pragma unroll
        for (int k = 0; k < PM; k++)
            if (k < P) index += cwr[g[k] * pstride + k + 1];
        return index;
    }
};

struct ExactItem {
    int U, P, H;
    int nA, nB;        // read columns [0, nA): count 1; [nA, nB): count 2..MCHB_EXACT_POW_MAX; [nB, U): the rest
    int n_factors;     // mantissa factors of the product form = sum of the counts of the columns below nB
    bool checked;      // the table holds entries that are not positive normal numbers
    bool has_prior, null_prior, has_freqs;
    double lg_left, p_log_h;
};

// per-read probability of a genotype: the reference's left fold over the slots (likelihood.py:60-66)
template <int PM>
__device__ __forceinline__ double exact_rp(const double *const (&row)[PM], int P, int r) {
    double rp = row[0][r];
#pragma unroll
    for (int k = 1; k < PM; k++)
        if (k < P) rp += row[k][r];
    return rp;
}

// mantissa in [1, 2) of a positive normal double; its biased exponent is hi >> 20
__device__ __forceinline__ double mantissa_of(double x, int hi) {
    return __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}

// sum_r count_r * log(sum_k t[g_k][r]) for one genotype (header, point 2).  The reads of the table
// are grouped by count: columns [0, nA) have count 1, [nA, nB) counts 2..MCHB_EXACT_POW_MAX,
// [nB, U) anything else; s.cnt holds the counts in column order.  CHECKED: table entries may be
// zero / denormal, so every rp is tested and an odd one sends the genotype to the reference's form.
template <int PM, bool CHECKED>
__device__ __forceinline__ double exact_llk(const GenoR<PM> &g, const ExactSmem &s, const ExactItem &it) {
    const double *row[PM];
#pragma unroll
    for (int k = 0; k < PM; k++) row[k] = s.tab + (size_t)g.g[k] * s.us;
    double mant = 1.0, slow = 0.0;
    int expo = 0;
    unsigned worst = 0;  // max over reads of (hi - 0x00100000) as unsigned: >= 0x7fe00000 iff some rp is odd
    // ---- count 1: one factor per read, renormalised every 512 reads
    for (int base = 0; base < it.nA; base += 512) {
        const int stop = min(it.nA, base + 512);
#pragma unroll 4
        for (int r = base; r < stop; r++) {
            const double rp = exact_rp<PM>(row, it.P, r);
            const int hi = __double2hiint(rp);
            if (CHECKED) worst = max(worst, (unsigned)(hi - 0x00100000));
            expo += hi >> 20;
            mant *= mantissa_of(rp, hi);
        }
        const int mh = __double2hiint(mant);
        expo += (mh >> 20) - 1023;
        mant = mantissa_of(mant, mh);
    }
    // ---- small counts: count factors per read (uniform inner loop), renormalised every 32 reads
    for (int base = it.nA; base < it.nB; base += 32) {
        const int stop = min(it.nB, base + 32);
#pragma unroll 1
        for (int r = base; r < stop; r++) {
            const double rp = exact_rp<PM>(row, it.P, r);
            const int hi = __double2hiint(rp);
            if (CHECKED) worst = max(worst, (unsigned)(hi - 0x00100000));
            const int c = s.cnt[r];
            const double m = mantissa_of(rp, hi);
            expo += (hi >> 20) * c;
#pragma unroll 1
            for (int q = 0; q < c; q++) mant *= m;
        }
        const int mh = __double2hiint(mant);
        expo += (mh >> 20) - 1023;
        mant = mantissa_of(mant, mh);
    }
    // ---- everything else in the reference's form
#pragma unroll 1
    for (int r = it.nB; r < it.U; r++) {
        const double rp = exact_rp<PM>(row, it.P, r);
        if (CHECKED) worst = max(worst, (unsigned)(__double2hiint(rp) - 0x00100000));
        slow += log(rp) * (double)s.cnt[r];
    }
    if (CHECKED && worst >= 0x7fe00000u) {
        // some rp is zero, denormal, negative, inf or NaN: read by read (log(0) = -inf, NaN propagates)
        double llk = 0.0;
#pragma unroll 1
        for (int r = 0; r < it.U; r++) llk += log(exact_rp<PM>(row, it.P, r)) * (double)s.cnt[r];
        return llk;
    }
    const double e = (double)(expo - 1023 * it.n_factors);
    return (e * MCHB_LN2_HI + (e * MCHB_LN2_LO + log(mant))) + slow;
}

// log joint of a genotype; the table of almost every item is free of zeros (CHECKED = false)
template <int PM>
__device__ __forceinline__ double exact_llk_any(const GenoR<PM> &g, const ExactSmem &s, const ExactItem &it) {
    return it.checked ? exact_llk<PM, true>(g, s, it) : exact_llk<PM, false>(g, s, it);
}

// calling/prior.py:116-179 for a sorted genotype: the dosage of an allele is the length of its run
// and the reference adds the terms at the first copy of every allele, i.e. run by run.
template <int PM>
__device__ __forceinline__ double exact_log_prior(const GenoR<PM> &g, const ExactSmem &s, const ExactItem &it) {
    double acc = 0.0;
    int run = 1;
#pragma unroll
    for (int k = 0; k < PM; k++) {
        if (k < it.P) {
            const bool last = (k == it.P - 1) || (g.g[k + 1 < PM ? k + 1 : k] != g.g[k]);
            if (last) {
                if (it.null_prior) acc += LGAMMA_INT[run + 1];
                else acc += s.lgDA[g.g[k] * (it.P + 1) + run];  // num - denom of prior.py:170-176, tabulated per item
                run = 1;
            } else {
                run += 1;
            }
        }
    }
    if (!it.null_prior) return it.lg_left + acc;
    const double ln_perms = LGAMMA_INT[it.P + 1] - acc;
    if (!it.has_freqs) return ln_perms - it.p_log_h;
    double prod = 1.0;
#pragma unroll
    for (int k = 0; k < PM; k++)
        if (k < it.P) prod *= s.fr[g.g[k]];
    return ln_perms + log(prod);
}

// prior tables of an item (cooperative) -> lg_left; call between barriers
__device__ __forceinline__ double exact_prior_tables(const ExactSmem &s, int H, int P, double inbreeding,
                                                     const double *freqs, int tid, int nthr) {
    for (int i = tid; i < H; i += nthr) s.fr[i] = freqs ? freqs[i] : 0.0;
    if (isnan(inbreeding) || inbreeding == 0.0) return 0.0;
    const double scale = (1.0 - inbreeding) / inbreeding;
    const double alpha_const = (1.0 / (double)H) * scale;
    for (int i = tid; i < H * (P + 1); i += nthr) {
        const int al = i / (P + 1), d = i - al * (P + 1);
        const double alpha = freqs ? freqs[al] * scale : alpha_const;
        // term of an allele with dosage d: lgamma(d + alpha) - (lgamma(d + 1) + lgamma(alpha))
        if (d > 0) s.lgDA[al * (P + 1) + d] = lgamma((double)d + alpha) - (LGAMMA_INT[d + 1] + lgamma(alpha));
    }
    double sum_alphas = 0.0;
    if (freqs) for (int al = 0; al < H; al++) sum_alphas += freqs[al] * scale;
    else sum_alphas = alpha_const * (double)H;
    return (LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma((double)P + sum_alphas);
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

// Fixed-order sum of one value per thread: xor butterfly inside the warp, warps in order.
__device__ __forceinline__ double block_sum_ordered(double v, double *red, int tid, int nthr) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(MCHB_FULL, v, m);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = red[0];
    for (int w = 1; w < (nthr >> 5); w++) t += red[w];
    return t;
}

// Allele statistics of one genotype with weight p into the thread's partial row:
// row[2a] += p per copy, row[2a+1] += p once per distinct allele (exact.py:139-148)
template <int PM>
__device__ __forceinline__ void exact_tally(const GenoR<PM> &g, int P, double p, double *row) {
#pragma unroll
    for (int k = 0; k < PM; k++) {
        if (k < P) {
            double *e = row + 2 * g.g[k];
            e[0] += p;
            if (k == 0 || g.g[k] != g.g[k > 0 ? k - 1 : 0]) e[1] += p;
        }
    }
}

// PM: register slots of a genotype; FIXED: every item of the launch has ploidy == PM (the slot
// loops carry no predicates); RECOMP: no parked log joints, the second pass evaluates them again.
template <int PM, bool FIXED, bool RECOMP>
__global__ void __launch_bounds__(128, PM <= 8 ? 8 : 4) exact_kernel(const __grid_constant__ ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const ExactSmem s = exact_carve(smem_raw, a.umax, a.hmax, a.pmax, a.part_threads);
    const int cstride = a.pmax + 1;
    ModeRec *red_m = reinterpret_cast<ModeRec *>(s.red + 8);  // [4]
    __shared__ int s_item;
    __shared__ int s_tabHP;  // (H << 8) | P the binomial table was filled for
    __shared__ int s_groups[3];
    double *scratch = RECOMP ? nullptr : a.scratch + (size_t)blockIdx.x * a.scratch_stride;
    if (tid == 0) s_tabHP = -1;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;
        const mchb_call_item itd = a.items[item];
        const int U = itd.n_reads, N = itd.n_pos, A = itd.max_allele, H = itd.n_haps;
        const int P = FIXED ? PM : itd.ploidy;
        const double *R = a.reads + itd.reads_off;
        const int8_t *haps = a.haplotypes + itd.haps_off;
        const double *freqs = (a.freqs && itd.freqs_off >= 0) ? a.freqs + itd.freqs_off : nullptr;
        const double dP = (double)P;
        ExactItem it;
        it.U = U;
        it.P = P;
        it.H = H;
        it.has_prior = !isnan(itd.inbreeding);
        it.null_prior = !(it.has_prior && itd.inbreeding != 0.0);
        it.has_freqs = freqs != nullptr;
        it.p_log_h = dP * log((double)H);

        // ---- 1. read columns grouped by count (thread 0; the order of the reads does not matter to a
        // product), table t[h][column] (likelihood.py:48-60 per haplotype), binomials, prior tables
        if (tid == 0) {
            int nA = 0, nB = 0, nf = 0;
            for (int pass = 0; pass < 3; pass++) {
                int col = pass == 0 ? 0 : (pass == 1 ? nA : nB);
                for (int r = 0; r < U; r++) {
                    const long long c = a.counts ? __ldg(a.counts + itd.counts_off + r) : 1;
                    const int cls = c == 1 ? 0 : ((c >= 2 && c <= MCHB_EXACT_POW_MAX) ? 1 : 2);
                    if (cls != pass) continue;
                    s.col[r] = col;
                    s.cnt[col] = (int)c;
                    if (pass < 2) nf += (int)c;
                    col++;
                }
                if (pass == 0) nA = col;
                if (pass == 1) nB = col;
            }
            s_groups[0] = nA;
            s_groups[1] = nB;
            s_groups[2] = nf;
        }
        __syncthreads();
        it.nA = s_groups[0];
        it.nB = s_groups[1];
        it.n_factors = s_groups[2];
        int odd_entry = 0;
        for (int i = tid; i < U * H; i += nthr) {
            const int h = i / U, r = i - h * U;
            double prod = 1.0;
            for (int j = 0; j < N; j++) {
                double v = __ldg(R + ((size_t)r * N + j) * A + haps[h * N + j]);
                if (!isnan(v)) prod *= v;
            }
            const double t = prod / dP;
            s.tab[h * s.us + s.col[r]] = t;
            odd_entry |= (unsigned)(__double2hiint(t) - 0x00100000) >= 0x7fe00000u;
        }
        it.checked = __syncthreads_or(odd_entry) != 0;
        if (s_tabHP != ((H << 8) | P)) {  // uniform: read before the barrier inside the fill
            exact_fill_cwr(s.cwr, H, P, cstride, tid, nthr);
            if (tid == 0) s_tabHP = (H << 8) | P;
        }
        it.lg_left = exact_prior_tables(s, H, P, itd.inbreeding, freqs, tid, nthr);
        __syncthreads();
        const long long G = s.cwr[H * cstride + P];

        // ---- 2. first pass over this thread's chunk of genotypes in VCF order
        const long long chunk = (G + nthr - 1) / nthr;
        const long long g0 = (long long)tid * chunk;
        const long long g1 = g0 + chunk < G ? g0 + chunk : G;
        ModeRec best;
        best.ljoint = -INFINITY;
        best.llk = -INFINITY;
        best.idx = 0x7fffffffffffffffLL;
        if (g0 < G) {
            GenoR<PM> g;
            g.unrank(g0, P, s.cwr, cstride, H);
            for (long long gi = g0; gi < g1; gi++) {
                const double llk = exact_llk_any<PM>(g, s, it);
                if (a.mode == 1) {
                    a.out_gl[itd.gl_off + gi] = (float)llk;
                } else {
                    const double ljoint = llk + (it.has_prior ? exact_log_prior<PM>(g, s, it) : 0.0);
                    if (ljoint > best.ljoint) {
                        best.ljoint = ljoint;
                        best.llk = llk;
                        best.idx = gi;
                    }
                    if (!RECOMP) scratch[(gi - g0) * nthr + tid] = ljoint;  // [step][thread]: coalesced
                }
                g.increment(P);
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
        // ---- 3. first maximum over the block
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            ModeRec o;
            o.ljoint = __shfl_xor_sync(MCHB_FULL, best.ljoint, m);
            o.llk = __shfl_xor_sync(MCHB_FULL, best.llk, m);
            o.idx = __shfl_xor_sync(MCHB_FULL, best.idx, m);
            best = mode_better(best, o);
        }
        if ((tid & 31) == 0) red_m[tid >> 5] = best;
        __syncthreads();  // also: every log joint of the first pass is in scratch
        best = red_m[0];
        for (int w = 1; w < (nthr >> 5); w++) best = mode_better(best, red_m[w]);
        if (!(best.ljoint > -INFINITY)) {  // nothing ever exceeded -inf: the reference keeps index 0
            best.idx = 0;
            best.llk = -INFINITY;
            best.ljoint = -INFINITY;
        }
        const double top = best.ljoint;
        // ---- 4. second pass: sum of exp(ljoint - top) and the allele statistics in per-thread rows
        const int pt = a.part_threads;
        double ssum = 0.0;
        if (tid < pt) {
            double *row = s.part + (size_t)tid * s.ps;
            for (int i = 0; i < 2 * H; i++) row[i] = 0.0;
            const long long chunk2 = (G + pt - 1) / pt;
            const long long b0 = (long long)tid * chunk2;
            const long long b1 = b0 + chunk2 < G ? b0 + chunk2 : G;
            if (b0 < G) {
                GenoR<PM> g;
                g.unrank(b0, P, s.cwr, cstride, H);
                long long owner = b0 / chunk, step = b0 - owner * chunk;  // where the first pass parked genotype b0
                for (long long gi = b0; gi < b1; gi++) {
                    double lj;
                    if (!RECOMP) lj = scratch[step * nthr + owner];
                    else lj = exact_llk_any<PM>(g, s, it) + (it.has_prior ? exact_log_prior<PM>(g, s, it) : 0.0);
                    const double p = exp(lj - top);
                    ssum += p;
                    exact_tally<PM>(g, P, p, row);
                    g.increment(P);
                    if (++step == chunk) {
                        step = 0;
                        owner++;
                    }
                }
            }
        }
        const double stot = block_sum_ordered(ssum, s.red, tid, nthr);
        const double total = top + log(stot);
        // ---- 5. outputs: allele statistics (partial rows added in thread order), mode, support
        for (int i = tid; i < 2 * H; i += nthr) {
            double acc = 0.0;
            for (int t = 0; t < pt; t++) acc += s.part[(size_t)t * s.ps + i];
            const double v = acc / stot;
            if (i & 1) a.out_occur[itd.hap_out_off + (i >> 1)] = v;
            else a.out_freqs[itd.hap_out_off + (i >> 1)] = v / dP;
        }
        if (tid == 0) {
            GenoR<PM> mode;
            mode.unrank(best.idx, P, s.cwr, cstride, H);
#pragma unroll
            for (int k = 0; k < PM; k++)
                if (k < P) a.out_alleles[(size_t)item * a.pstride + k] = mode.g[k];
            for (int k = P; k < a.pstride; k++) a.out_alleles[(size_t)item * a.pstride + k] = -2;
            // exact.py:64-105: all dosage variants of the mode's haplotype set, in
            // combinations_with_replacement order
            int support[MCHB_MAX_PLOIDY];
            int ns = 0;
#pragma unroll
            for (int k = 0; k < PM; k++)
                if (k < P && (k == 0 || mode.g[k] != mode.g[k > 0 ? k - 1 : 0])) support[ns++] = mode.g[k];
            const int rem = P - ns;
            int idx[MCHB_MAX_PLOIDY];
            for (int k = 0; k < rem; k++) idx[k] = 0;
            double support_ljoint = -INFINITY;
            for (;;) {
                // merge support + chosen extras into a sorted genotype (counting sort over support)
                int merged[MCHB_MAX_PLOIDY];
                int pos = 0;
                for (int q = 0; q < ns; q++) {
                    int c = 1;
                    for (int k = 0; k < rem; k++) c += (idx[k] == q);
                    for (int k = 0; k < c; k++) merged[pos++] = support[q];
                }
                GenoR<PM> t;
#pragma unroll
                for (int k = 0; k < PM; k++) t.g[k] = k < P ? merged[k] : 0;
                double lj;
                if (!RECOMP) {
                    const long long rk = t.rank(P, s.cwr, cstride), owner = rk / chunk;
                    lj = scratch[(rk - owner * chunk) * nthr + owner];
                }
                else lj = exact_llk_any<PM>(t, s, it) + (it.has_prior ? exact_log_prior<PM>(t, s, it) : 0.0);
                support_ljoint = add_log_prob(support_ljoint, lj);
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
// Same enumeration, prior and fixed-order reductions as exact_kernel.
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
    int32_t part_threads;
};

template <int PM>
__global__ void __launch_bounds__(128) posterior_kernel(const __grid_constant__ PosteriorArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const ExactSmem s = exact_carve(smem_raw, 0, a.hmax, a.pmax, a.part_threads);
    const int cstride = a.pmax + 1;
    __shared__ int s_tabHP;
    if (tid == 0) s_tabHP = -1;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        __syncthreads();
        const mchb_call_item itd = a.items[item];
        const int P = itd.ploidy, H = itd.n_haps;
        const double *freqs = (a.freqs && itd.freqs_off >= 0) ? a.freqs + itd.freqs_off : nullptr;
        const double dP = (double)P;
        ExactItem it;
        it.U = it.nA = it.nB = it.n_factors = 0;
        it.P = P;
        it.H = H;
        it.checked = false;
        it.has_prior = !isnan(itd.inbreeding);
        it.null_prior = !(it.has_prior && itd.inbreeding != 0.0);
        it.has_freqs = freqs != nullptr;
        it.p_log_h = dP * log((double)H);
        if (s_tabHP != ((H << 8) | P)) {
            exact_fill_cwr(s.cwr, H, P, cstride, tid, nthr);
            if (tid == 0) s_tabHP = (H << 8) | P;
        }
        it.lg_left = exact_prior_tables(s, H, P, itd.inbreeding, freqs, tid, nthr);
        __syncthreads();
        const long long G = s.cwr[H * cstride + P];
        const long long chunk = (G + nthr - 1) / nthr;
        const long long g0 = (long long)tid * chunk;
        const long long g1 = g0 + chunk < G ? g0 + chunk : G;
        double *gp = a.out_gp + itd.gl_off;
        // ---- log joints (stored like the reference stores them) and their maximum
        double top = -INFINITY;
        bool any_nan = false;
        if (g0 < G) {
            GenoR<PM> g;
            g.unrank(g0, P, s.cwr, cstride, H);
            for (long long gi = g0; gi < g1; gi++) {
                const double lpr = it.has_prior ? exact_log_prior<PM>(g, s, it) : 0.0;
                double x;
                if (a.llk32) x = (double)(float)((double)a.llk32[itd.gl_off + gi] + lpr);
                else x = a.llk64[itd.gl_off + gi] + lpr;
                gp[gi] = x;
                any_nan = any_nan || isnan(x);
                if (x > top) top = x;
                g.increment(P);
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) top = fmax(top, __shfl_xor_sync(MCHB_FULL, top, m));
        __syncthreads();
        if ((tid & 31) == 0) s.red[8 + (tid >> 5)] = top;
        const int has_nan = __syncthreads_or(any_nan);
        top = s.red[8];
        for (int w = 1; w < (nthr >> 5); w++) top = fmax(top, s.red[8 + w]);
        if (has_nan) top = NAN;  // the reference's fold propagates NaN to every posterior
        // ---- normaliser: sum of exp(x - top) in fixed order
        double ssum = 0.0;
        for (long long gi = g0; gi < g1; gi++) {
            const double p = exp(gp[gi] - top);
            gp[gi] = p;
            ssum += p;
        }
        const double stot = block_sum_ordered(ssum, s.red, tid, nthr);
        // exp(x - (top + log(stot))) evaluated as exp(x - top) / stot
        const int pt = a.part_threads;
        if (a.out_freqs && tid < pt) {
            double *row = s.part + (size_t)tid * s.ps;
            for (int i = 0; i < 2 * H; i++) row[i] = 0.0;
        }
        for (long long gi = g0; gi < g1; gi++) gp[gi] = gp[gi] / stot;
        __syncthreads();
        if (a.out_freqs) {
            if (tid < pt) {
                double *row = s.part + (size_t)tid * s.ps;
                const long long chunk2 = (G + pt - 1) / pt;
                const long long b0 = (long long)tid * chunk2;
                const long long b1 = b0 + chunk2 < G ? b0 + chunk2 : G;
                if (b0 < G) {
                    GenoR<PM> g;
                    g.unrank(b0, P, s.cwr, cstride, H);
                    for (long long gi = b0; gi < b1; gi++) {
                        exact_tally<PM>(g, P, gp[gi], row);
                        g.increment(P);
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < 2 * H; i += nthr) {
                double acc = 0.0;
                for (int t = 0; t < pt; t++) acc += s.part[(size_t)t * s.ps + i];
                if (i & 1) {
                    a.out_occur[itd.hap_out_off + (i >> 1)] = acc;
                } else {
                    a.out_freqs[itd.hap_out_off + (i >> 1)] = acc / dP;
                    a.out_counts[itd.hap_out_off + (i >> 1)] = acc;
                }
            }
        }
    }
}

}  // namespace mchb
