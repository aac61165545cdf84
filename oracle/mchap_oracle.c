/*
 * mchap_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the MCHap (v0.11.1) haplotype /
 * genotype inference hot path.  It exists to CHECK the CUDA path; it is never
 * shipped, never imported by the product package (mchap_b200/) and never used
 * as a fallback.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs may load it.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the real
 * reference (numba path) in the build container and stores input/output
 * fixtures under tests/golden/; tests/test_oracle_golden.py checks this file
 * against them (bit-exact for integers and, on the same libm, for the fp64
 * log-likelihoods; trajectories step-for-step).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).  Arithmetic order follows the reference
 * (numba compiles without fast-math: no re-association, no FMA contraction;
 * build this file with -ffp-contract=off).
 *
 * Third-party arithmetic restated here (not vendored in the reference):
 *   numba 0.65.0 numba/cpython/randomimpl.py  (MT19937 + np.random lowering):
 *     get_next_int32 109-132, get_next_double 134-147, get_next_int 149-196,
 *     _randrange_impl 454-520, do_shuffle_impl 1929-1956, permutation 1968-1980,
 *     choice 2021-2070;  numba/_random.c numba_rnd_init / numba_rnd_shuffle.
 *   numba np.searchsorted: numba/np/arraymath.py 3841-3860 (binary search,
 *     NaN-aware <=), np.cumsum / ndarray.sum = sequential loops.
 *   libm log/exp/log1p/lgamma (numba lowers to the C library's).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_NAN_LLK 1        /* assemble/mcmc.py:330-331 */
#define ORC_ERR_BREAKS 2         /* assemble/structural.py:49-50 */
#define ORC_ERR_UNSORTED 3       /* jitutils.py:142-143 */
#define ORC_ERR_CHOICE_RANGE 4   /* random_choice returned len(p) where the reference would index out of range */
#define ORC_ERR_INITIAL_SHAPE 5  /* assemble/mcmc.py:207 */
#define ORC_ERR_RNG_EXHAUSTED 6  /* replay stream too short */
#define ORC_ERR_STEP_TYPE 7

/* ------------------------------------------------------------------------- */
/* RNG: MT19937 exactly as numba drives it                                   */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint32_t mt[624];
    int idx;
    int64_t words;            /* 32-bit words consumed since seeding */
    const uint32_t *replay;   /* optional pre-drawn (tempered) word stream */
    int64_t n_replay;
    int exhausted;
} orc_rng;

/* numba/_random.c numba_rnd_init */
void orc_rng_seed(orc_rng *s, uint32_t seed)
{
    for (int pos = 0; pos < 624; pos++) {
        s->mt[pos] = seed;
        seed = 1812433253U * (seed ^ (seed >> 30)) + (uint32_t)pos + 1U;
    }
    s->idx = 624;
    s->words = 0;
    s->replay = NULL;
    s->n_replay = 0;
    s->exhausted = 0;
}

void orc_rng_replay(orc_rng *s, const uint32_t *words, int64_t n)
{
    s->idx = 624;
    s->words = 0;
    s->replay = words;
    s->n_replay = n;
    s->exhausted = 0;
}

orc_rng *orc_rng_new(uint32_t seed)
{
    orc_rng *s = (orc_rng *)malloc(sizeof(orc_rng));
    orc_rng_seed(s, seed);
    return s;
}

orc_rng *orc_rng_new_replay(const uint32_t *words, int64_t n)
{
    orc_rng *s = (orc_rng *)malloc(sizeof(orc_rng));
    orc_rng_seed(s, 0);
    orc_rng_replay(s, words, n);
    return s;
}

void orc_rng_free(orc_rng *s) { free(s); }
int64_t orc_rng_words(const orc_rng *s) { return s->words; }

/* numba/_random.c numba_rnd_shuffle: regenerate the 624-word block */
static void rng_refill(orc_rng *s)
{
    uint32_t *mt = s->mt;
    int i;
    uint32_t y;
    for (i = 0; i < 624 - 397; i++) {
        y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU);
        mt[i] = mt[i + 397] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    for (; i < 623; i++) {
        y = (mt[i] & 0x80000000U) | (mt[i + 1] & 0x7fffffffU);
        mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
}

/* randomimpl.py:109-132 get_next_int32 */
static uint32_t rng_u32(orc_rng *s)
{
    uint32_t y;
    if (s->replay) {
        if (s->words >= s->n_replay) {
            s->exhausted = 1;
            s->words++;
            return 0;
        }
        return s->replay[s->words++];
    }
    if (s->idx >= 624) {
        rng_refill(s);
        s->idx = 0;
    }
    y = s->mt[s->idx++];
    s->words++;
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= (y >> 18);
    return y;
}

/* randomimpl.py:134-147 get_next_double (np.random.random / rand) */
static double rng_double(orc_rng *s)
{
    uint32_t a = rng_u32(s) >> 5;
    uint32_t b = rng_u32(s) >> 6;
    return ((double)b + (double)a * 67108864.0) / 9007199254740992.0;
}

/* randomimpl.py:149-196 get_next_int with is_numpy=True */
static uint64_t rng_bits(orc_rng *s, int nbits)
{
    if (nbits <= 32) {
        uint32_t y = rng_u32(s);
        uint32_t mask = 0xffffffffU >> (32 - nbits);
        return (uint64_t)(y & mask);
    } else {
        int hb = nbits - 32;
        uint32_t high = rng_u32(s) & (0xffffffffU >> (32 - hb));
        uint32_t low = rng_u32(s);
        return (uint64_t)low + ((uint64_t)high << 32);
    }
}

/* randomimpl.py:454-520 _randrange_impl, state == "np", int64 arguments */
static int64_t rng_randint(orc_rng *s, int64_t n)
{
    if (n == 1)
        return 0;
    int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
    for (;;) {
        int64_t r = (int64_t)rng_bits(s, nbits);
        if (r < n)
            return r;
        if (s->exhausted)
            return 0;
    }
}

/* exported probes used by the golden tests */
uint32_t orc_rng_next_u32(orc_rng *s) { return rng_u32(s); }
double orc_rng_next_double(orc_rng *s) { return rng_double(s); }
int64_t orc_rng_next_randint(orc_rng *s, int64_t n) { return rng_randint(s, n); }

/* fill a buffer with the tempered 32-bit output stream of seed */
void orc_mt19937_words(uint32_t seed, uint32_t *out, int64_t n)
{
    orc_rng s;
    orc_rng_seed(&s, seed);
    for (int64_t i = 0; i < n; i++)
        out[i] = rng_u32(&s);
}

/* randomimpl.py:1929-1956 do_shuffle_impl (1-D int64) */
void orc_rng_shuffle_i64(orc_rng *s, int64_t *x, int64_t n)
{
    for (int64_t i = n - 1; i > 0; i--) {
        int64_t j = rng_randint(s, i + 1);
        int64_t t = x[i];
        x[i] = x[j];
        x[j] = t;
    }
}

/* ------------------------------------------------------------------------- */
/* jitutils: log-space helpers, categorical choice                           */
/* ------------------------------------------------------------------------- */

/* jitutils.py:6-26 add_log_prob */
double orc_add_log_prob(double x, double y)
{
    if (x == -INFINITY && y == -INFINITY)
        return -INFINITY;
    if (x > y)
        return x + log1p(exp(y - x));
    else
        return y + log1p(exp(x - y));
}

/* jitutils.py:29-48 sum_log_probs */
double orc_sum_log_probs(const double *a, int64_t n)
{
    double acc = a[0];
    for (int64_t i = 1; i < n; i++)
        acc = orc_add_log_prob(acc, a[i]);
    return acc;
}

/* jitutils.py:51-74 normalise_log_probs */
void orc_normalise_log_probs(const double *llks, int64_t n, double *out)
{
    double denom = orc_sum_log_probs(llks, n);
    for (int64_t i = 0; i < n; i++)
        out[i] = exp(llks[i] - denom);
}

/* numba np.searchsorted(a, v, side="right"): arraymath.py:3841-3860 with the
 * NaN-aware <= of 3801-3809 (v is never NaN here) */
static int64_t searchsorted_right(const double *a, int64_t n, double v)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* jitutils.py:77-92 random_choice: searchsorted(cumsum(p), random(), "right") */
static int64_t random_choice(orc_rng *s, const double *p, int64_t n)
{
    double stackbuf[64];
    double *cs = n <= 64 ? stackbuf : (double *)malloc(sizeof(double) * (size_t)n);
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) {
        acc += p[i];
        cs[i] = acc;
    }
    double u = rng_double(s);
    int64_t r = searchsorted_right(cs, n, u);
    if (cs != stackbuf)
        free(cs);
    return r;
}

int64_t orc_random_choice(orc_rng *s, const double *p, int64_t n) { return random_choice(s, p, n); }

/* ------------------------------------------------------------------------- */
/* jitutils: combinatorics, VCF-order genotype rank / unrank                 */
/* ------------------------------------------------------------------------- */

static int64_t gcd_i64(int64_t x, int64_t y) /* jitutils.py:186-192 */
{
    while (y != 0) {
        int64_t t = x % y;
        x = y;
        y = t;
    }
    return x;
}

/* jitutils.py:195-210 _comb (exact, gcd-reduced) ; the 100x12 table of
 * 213-228 holds the same values so no table is needed for correctness */
int64_t orc_comb(int64_t n, int64_t k)
{
    if (n < 0 || k < 0)
        return -1; /* reference raises ValueError */
    if (k > n)
        return 0;
    int64_t r = 1;
    for (int64_t d = 1; d <= k; d++) {
        int64_t g = gcd_i64(r, d);
        r /= g;
        r *= n;
        r /= d / g;
        n -= 1;
    }
    return r;
}

/* jitutils.py:231-250 comb_with_replacement; note (0,0) -> 0 (232-233) */
int64_t orc_comb_with_replacement(int64_t n, int64_t k)
{
    if (n < 0)
        return -1;
    if (n == 0 && k == 0)
        return 0;
    return orc_comb(n + k - 1, k);
}

/* jitutils.py:113-146 increment_genotype; returns ORC_ERR_UNSORTED on the
 * reference's ValueError */
int orc_increment_genotype(int64_t *g, int ploidy)
{
    if (ploidy == 1) {
        g[0] += 1;
        return ORC_OK;
    }
    int64_t previous = g[0];
    for (int i = 1; i < ploidy; i++) {
        int64_t allele = g[i];
        if (allele == previous) {
            continue;
        } else if (allele > previous) {
            int k = i - 1;
            g[k] += 1;
            for (int m = 0; m < k; m++)
                g[m] = 0;
            return ORC_OK;
        } else {
            return ORC_ERR_UNSORTED;
        }
    }
    g[ploidy - 1] += 1;
    for (int m = 0; m < ploidy - 1; m++)
        g[m] = 0;
    return ORC_OK;
}

/* jitutils.py:253-276 genotype_alleles_as_index; -1 on negative allele */
int64_t orc_genotype_alleles_as_index(const int64_t *alleles, int ploidy)
{
    int64_t index = 0;
    for (int i = 0; i < ploidy; i++) {
        int64_t a = alleles[i];
        if (a >= 0)
            index += orc_comb_with_replacement(a, i + 1);
        else
            return -1;
    }
    return index;
}

/* jitutils.py:279-318 index_as_genotype_alleles. Returns 1 when the reference
 * returns None (index < 0, 300-303), else 0 */
int orc_index_as_genotype_alleles(int64_t index, int ploidy, int64_t *out)
{
    for (int i = 0; i < ploidy; i++)
        out[i] = -2;
    if (index < 0) {
        for (int i = 0; i < ploidy; i++)
            out[i] = -1;
        return 1;
    }
    int64_t remainder = index;
    for (int it = 0; it < ploidy; it++) {
        int p = ploidy - it;
        int64_t n = -1, nw = 0, prev = 0;
        while (nw <= remainder) {
            n += 1;
            prev = nw;
            nw = orc_comb_with_replacement(n, p);
        }
        n -= 1;
        remainder -= prev;
        out[p - 1] = n;
    }
    return 0;
}

/* jitutils.py:149-171 ln_equivalent_permutations */
static double ln_equivalent_permutations_i64(const int64_t *dosage, int n)
{
    int64_t ploidy = 0;
    for (int i = 0; i < n; i++)
        ploidy += dosage[i];
    double ln_num = lgamma((double)(ploidy + 1));
    double ln_denom = 0.0;
    for (int i = 0; i < n; i++)
        ln_denom += lgamma((double)(dosage[i] + 1));
    return ln_num - ln_denom;
}

double orc_ln_equivalent_permutations(const int64_t *dosage, int n)
{
    return ln_equivalent_permutations_i64(dosage, n);
}

/* ------------------------------------------------------------------------- */
/* jitutils: haplotype bookkeeping                                            */
/* ------------------------------------------------------------------------- */

/* jitutils.py:321-346 array_equal over a half-open interval */
static int row_equal(const int8_t *x, const int8_t *y, int start, int stop)
{
    for (int i = start; i < stop; i++)
        if (x[i] != y[i])
            return 0;
    return 1;
}

/* jitutils.py:349-374 count_haplotype_copies */
int orc_count_haplotype_copies(const int8_t *genotype, int P, int N, int h)
{
    int count = 1;
    for (int i = 0; i < P; i++) {
        if (i == h)
            continue;
        if (row_equal(genotype + (size_t)i * N, genotype + (size_t)h * N, 0, N))
            count += 1;
    }
    return count;
}

/* jitutils.py:377-422 get_haplotype_dosage (rows of width N, compare [start,stop)) */
static void haplotype_dosage(int8_t *dosage, const int8_t *rows, int P, int N, int start, int stop)
{
    for (int h = 0; h < P; h++)
        dosage[h] = 1;
    for (int h = 0; h < P; h++) {
        if (dosage[h] == 0)
            continue;
        for (int p = h + 1; p < P; p++) {
            if (dosage[p] == 0)
                continue;
            if (row_equal(rows + (size_t)h * N, rows + (size_t)p * N, start, stop)) {
                dosage[h] += 1;
                dosage[p] = 0;
            }
        }
    }
}

void orc_get_haplotype_dosage(int8_t *dosage, const int8_t *genotype, int P, int N)
{
    haplotype_dosage(dosage, genotype, P, N, 0, N);
}

/* jitutils.py:501-544 structural_change; interval [start,stop) */
void orc_structural_change(int8_t *genotype, int P, int N, const int8_t *hap_idx, int start, int stop)
{
    int8_t cache[256];
    for (int j = start; j < stop; j++) {
        for (int h = 0; h < P; h++)
            cache[h] = genotype[(size_t)h * N + j];
        for (int h = 0; h < P; h++)
            genotype[(size_t)h * N + j] = cache[hap_idx[h]];
    }
}

/* ------------------------------------------------------------------------- */
/* assemble/likelihood.py                                                     */
/* ------------------------------------------------------------------------- */

/* assemble/likelihood.py:18-70 log_likelihood.
 * reads f64[U,N,A] C-contiguous, genotype int8[P,N], counts i64[U] or NULL */
double orc_log_likelihood(const double *reads, int U, int N, int A,
                          const int8_t *genotype, int P, const int64_t *counts)
{
    double llk = 0.0;
    for (int r = 0; r < U; r++) {
        double read_prob = 0.0;
        for (int h = 0; h < P; h++) {
            double prod = 1.0;
            for (int j = 0; j < N; j++) {
                int i = genotype[(size_t)h * N + j];
                double val = reads[((size_t)r * N + j) * A + i];
                if (!isnan(val))
                    prod *= val;
            }
            read_prob += prod / (double)P;
        }
        double lrp = log(read_prob);
        if (counts)
            lrp *= (double)counts[r];
        llk += lrp;
    }
    return llk;
}

/* assemble/likelihood.py:74-148 log_likelihood_structural_change */
double orc_log_likelihood_structural_change(const double *reads, int U, int N, int A,
                                            const int8_t *genotype, int P,
                                            const int8_t *hap_idx, int start, int stop,
                                            const int64_t *counts)
{
    double llk = 0.0;
    for (int r = 0; r < U; r++) {
        double read_prob = 0.0;
        for (int h = 0; h < P; h++) {
            double prod = 1.0;
            for (int j = 0; j < N; j++) {
                int h_ = (j >= start && j < stop) ? hap_idx[h] : h;
                int i = genotype[(size_t)h_ * N + j];
                double val = reads[((size_t)r * N + j) * A + i];
                if (!isnan(val))
                    prod *= val;
            }
            read_prob += prod / (double)P;
        }
        double lrp = log(read_prob);
        if (counts)
            lrp *= (double)counts[r];
        llk += lrp;
    }
    return llk;
}

/* ------------------------------------------------------------------------- */
/* assemble/prior.py                                                          */
/* ------------------------------------------------------------------------- */

/* assemble/prior.py:15-36 + 39-78 + 81-112 log_genotype_prior (dosage int8[P]) */
double orc_assemble_log_genotype_prior(const int8_t *dosage, int P,
                                       double log_unique_haplotypes, double inbreeding)
{
    int64_t ploidy = 0;
    for (int i = 0; i < P; i++)
        ploidy += dosage[i];
    if (inbreeding == 0.0) {
        /* null prior: ln(perms) - ploidy * log_unique_haplotypes */
        double ln_num = lgamma((double)(ploidy + 1));
        double ln_denom = 0.0;
        for (int i = 0; i < P; i++)
            ln_denom += lgamma((double)(dosage[i] + 1));
        double ln_perms = ln_num - ln_denom;
        double ln_total = (double)ploidy * log_unique_haplotypes;
        return ln_perms - ln_total;
    }
    double log_dispersion = log((1.0 - inbreeding) / inbreeding) - log_unique_haplotypes;
    double dispersion = exp(log_dispersion);
    double sum_dispersion = exp(log_dispersion + log_unique_haplotypes);
    double num = lgamma((double)(ploidy + 1)) + lgamma(sum_dispersion);
    double denom = lgamma((double)ploidy + sum_dispersion);
    double left = num - denom;
    double prod = 0.0;
    for (int i = 0; i < P; i++) {
        int dose = dosage[i];
        if (dose > 0) {
            double n2 = lgamma((double)dose + dispersion);
            double d2 = lgamma((double)(dose + 1)) + lgamma(dispersion);
            prod += n2 - d2;
        }
    }
    return left + prod;
}

/* ------------------------------------------------------------------------- */
/* calling/utils.py + calling/prior.py                                        */
/* ------------------------------------------------------------------------- */

/* calling/utils.py:7-35 allelic_dosage */
static void allelic_dosage(const int64_t *g, int P, int64_t *dosage)
{
    for (int i = 0; i < P; i++)
        dosage[i] = 0;
    for (int i = 0; i < P; i++) {
        int64_t a = g[i];
        int j = 0;
        while (g[j] != a)
            j++;
        dosage[j] += 1;
    }
}

void orc_allelic_dosage(const int64_t *g, int P, int64_t *dosage) { allelic_dosage(g, P, dosage); }

/* calling/utils.py:38-57 count_allele */
static int count_allele(const int64_t *g, int P, int64_t allele)
{
    int c = 0;
    for (int i = 0; i < P; i++)
        if (g[i] == allele)
            c++;
    return c;
}

/* calling/prior.py:116-179 log_genotype_prior. freqs may be NULL */
double orc_calling_log_genotype_prior(const int64_t *g, int P, int64_t unique_haplotypes,
                                      double inbreeding, const double *freqs)
{
    int64_t dosage[64];
    allelic_dosage(g, P, dosage);
    if (inbreeding == 0.0) {
        double ln_perms = ln_equivalent_permutations_i64(dosage, P);
        if (!freqs)
            return ln_perms - (double)P * log((double)unique_haplotypes);
        double prod = 1.0;
        for (int i = 0; i < P; i++)
            prod *= freqs[g[i]];
        return ln_perms + log(prod);
    }
    double alpha_const = 0.0, sum_alphas = 0.0;
    double scale = (1.0 - inbreeding) / inbreeding;
    if (!freqs) {
        alpha_const = (1.0 / (double)unique_haplotypes) * scale;
        sum_alphas = alpha_const * (double)unique_haplotypes;
    } else {
        /* alphas.sum(): sequential over the array */
        for (int64_t a = 0; a < unique_haplotypes; a++)
            sum_alphas += freqs[a] * scale;
    }
    double num = lgamma((double)(P + 1)) + lgamma(sum_alphas);
    double denom = lgamma((double)P + sum_alphas);
    double left = num - denom;
    double prod = 0.0;
    for (int i = 0; i < P; i++) {
        int64_t dose = dosage[i];
        if (dose > 0) {
            double alpha_i = freqs ? freqs[g[i]] * scale : alpha_const;
            double n2 = lgamma((double)dose + alpha_i);
            double d2 = lgamma((double)(dose + 1)) + lgamma(alpha_i);
            prod += n2 - d2;
        }
    }
    return left + prod;
}

/* calling/prior.py:30-52 log_genotype_allele_flat_prior */
double orc_log_genotype_allele_flat_prior(const int64_t *g, int P, int variable_allele)
{
    int n = 0;
    int64_t a = g[variable_allele];
    for (int i = 0; i < P; i++)
        n += (g[i] == a);
    return log((double)n);
}

/* calling/prior.py:55-113 log_genotype_allele_prior */
double orc_log_genotype_allele_prior(const int64_t *g, int P, int variable_allele,
                                     int64_t unique_haplotypes, double inbreeding,
                                     const double *freqs)
{
    if (inbreeding == 0.0) {
        if (!freqs)
            return log(1.0 / (double)unique_haplotypes);
        return log(freqs[g[variable_allele]]);
    }
    int constant_sum = P - 1;
    int constant_ibs = count_allele(g, P, g[variable_allele]) - 1;
    double scale = (1.0 - inbreeding) / inbreeding;
    double sum_alpha, variable_alpha;
    if (!freqs) {
        double alpha = (1.0 / (double)unique_haplotypes) * scale;
        sum_alpha = (double)constant_sum + alpha * (double)unique_haplotypes;
        variable_alpha = alpha + (double)constant_ibs;
    } else {
        double s = 0.0;
        for (int64_t a = 0; a < unique_haplotypes; a++)
            s += freqs[a] * scale;
        sum_alpha = (double)constant_sum + s;
        variable_alpha = freqs[g[variable_allele]] * scale + (double)constant_ibs;
    }
    double left = lgamma(sum_alpha) - lgamma(1.0 + sum_alpha);
    double right = lgamma(1.0 + variable_alpha) - lgamma(variable_alpha);
    return left + right;
}

/* ------------------------------------------------------------------------- */
/* assemble/mutation.py                                                       */
/* ------------------------------------------------------------------------- */

/* DenovoMCMC.llk_cache_threshold (assemble/mcmc.py:39, default 100; -1 disables) */
static int64_t orc_llk_cache_threshold = 100;
void orc_set_llk_cache_threshold(int64_t t) { orc_llk_cache_threshold = t; }

typedef struct {
    const double *reads; /* f64[U,N,A] */
    int U, N, A;
    const int64_t *counts; /* or NULL */
    int P;
    int has_inbreeding; /* 0 => inbreeding is None (flat prior) */
    double inbreeding;
    double log_unique_haplotypes;
    int64_t llk_evals; /* statistics: number of log_likelihood evaluations requested */
    /* memo of log-likelihoods keyed by the ordered genotype: stands in for the reference's
     * array-map cache (assemble/likelihood.py:152-305, assemble/arraymap.py; enabled iff
     * ploidy * n_base * n_reads > llk_cache_threshold, assemble/mcmc.py:305-312).  Like the
     * reference's cache it never changes a returned value (the llk is a deterministic function of
     * the ordered genotype) and it is emptied when full. */
    int use_cache;
    int cache_bits;
    uint64_t *cache_hash;
    int8_t *cache_key;
    double *cache_val;
    int64_t cache_n;
    int64_t cache_hits;
} asm_ctx;

static uint64_t geno_hash(const int8_t *g, size_t n)
{
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < n; i++) {
        h ^= (uint64_t)(uint8_t)g[i];
        h *= 1099511628211ULL;
    }
    return h | 1ULL; /* 0 marks an empty slot */
}

static void asm_cache_init(asm_ctx *c, int enable)
{
    c->use_cache = enable;
    c->cache_bits = 13;
    c->cache_n = 0;
    c->cache_hits = 0;
    c->cache_hash = NULL;
    c->cache_key = NULL;
    c->cache_val = NULL;
    if (enable) {
        size_t cap = (size_t)1 << c->cache_bits;
        c->cache_hash = (uint64_t *)calloc(cap, sizeof(uint64_t));
        c->cache_key = (int8_t *)malloc(cap * (size_t)c->P * c->N);
        c->cache_val = (double *)malloc(cap * sizeof(double));
    }
}

static void asm_cache_free(asm_ctx *c)
{
    free(c->cache_hash);
    free(c->cache_key);
    free(c->cache_val);
    c->cache_hash = NULL;
    c->cache_key = NULL;
    c->cache_val = NULL;
}

/* log_likelihood_cached (assemble/likelihood.py:195-235) on an explicit genotype */
static double asm_llk_cached(asm_ctx *c, const int8_t *genotype)
{
    c->llk_evals++;
    if (!c->use_cache)
        return orc_log_likelihood(c->reads, c->U, c->N, c->A, genotype, c->P, c->counts);
    const size_t gsz = (size_t)c->P * c->N;
    const size_t cap = (size_t)1 << c->cache_bits;
    const uint64_t h = geno_hash(genotype, gsz);
    size_t s = (size_t)(h >> 17) & (cap - 1);
    while (c->cache_hash[s]) {
        if (c->cache_hash[s] == h && memcmp(c->cache_key + s * gsz, genotype, gsz) == 0) {
            c->cache_hits++;
            return c->cache_val[s];
        }
        s = (s + 1) & (cap - 1);
    }
    double v = orc_log_likelihood(c->reads, c->U, c->N, c->A, genotype, c->P, c->counts);
    if ((size_t)c->cache_n * 2 >= cap) { /* empty the cache when (half) full */
        memset(c->cache_hash, 0, cap * sizeof(uint64_t));
        c->cache_n = 0;
        s = (size_t)(h >> 17) & (cap - 1);
    }
    c->cache_hash[s] = h;
    memcpy(c->cache_key + s * gsz, genotype, gsz);
    c->cache_val[s] = v;
    c->cache_n++;
    return v;
}

static double np_minimum0(double x) /* np.minimum(0.0, x): NaN propagates */
{
    if (isnan(x))
        return x;
    return x < 0.0 ? x : 0.0;
}

static double asm_prior_of_rows(const asm_ctx *c, const int8_t *rows, int width)
{
    int8_t dosage[256];
    haplotype_dosage(dosage, rows, c->P, width, 0, width);
    return orc_assemble_log_genotype_prior(dosage, c->P, c->log_unique_haplotypes, c->inbreeding);
}

/* assemble/mutation.py:15-161 base_step */
static double base_step(asm_ctx *c, orc_rng *rng, int8_t *genotype, double llk, int h, int j,
                        int n_alleles, double temp, int *err)
{
    int P = c->P, N = c->N;
    double llks[64], log_accept[64], probs[64];
    double lhapcount = log((double)orc_count_haplotype_copies(genotype, P, N, h));
    double lprior = 0.0;
    if (c->has_inbreeding)
        lprior = asm_prior_of_rows(c, genotype, N);
    int current = genotype[(size_t)h * N + j];
    int n_options = 0;
    for (int i = 0; i < n_alleles; i++) {
        if (i == current) {
            llks[i] = llk;
            log_accept[i] = -INFINITY;
        } else {
            n_options += 1;
            genotype[(size_t)h * N + j] = (int8_t)i;
            double llk_i = asm_llk_cached(c, genotype);
            llks[i] = llk_i;
            double llk_ratio = llk_i - llk;
            double lprior_ratio = 0.0;
            if (c->has_inbreeding) {
                double lprior_i = asm_prior_of_rows(c, genotype, N);
                lprior_ratio = lprior_i - lprior;
            }
            double lhapcount_i = log((double)orc_count_haplotype_copies(genotype, P, N, h));
            double lproposal_ratio = lhapcount_i - lhapcount;
            double mh_ratio = (llk_ratio + lprior_ratio) * temp + lproposal_ratio;
            log_accept[i] = np_minimum0(mh_ratio);
        }
    }
    double ln_opts = log((double)n_options);
    double sum = 0.0;
    for (int i = 0; i < n_alleles; i++) {
        log_accept[i] -= ln_opts;
        probs[i] = exp(log_accept[i]);
    }
    for (int i = 0; i < n_alleles; i++)
        sum += probs[i];
    probs[current] = 1 - sum;
    int64_t choice = random_choice(rng, probs, n_alleles);
    if (choice >= n_alleles) {
        /* the reference would write allele == n_alleles and read llks out of range */
        genotype[(size_t)h * N + j] = (int8_t)current;
        *err = ORC_ERR_CHOICE_RANGE;
        return llk;
    }
    genotype[(size_t)h * N + j] = (int8_t)choice;
    return llks[choice];
}

/* assemble/mutation.py:165-246 compound_step. n_alleles int8[N] */
static double mutation_compound_step(asm_ctx *c, orc_rng *rng, int8_t *genotype, double llk,
                                     const int8_t *n_alleles, double temp, int *err)
{
    int P = c->P, N = c->N;
    int n = P * N;
    int8_t *sub = (int8_t *)malloc((size_t)n * 2);
    for (int h = 0; h < P; h++)
        for (int j = 0; j < N; j++) {
            sub[2 * (h * N + j)] = (int8_t)h;
            sub[2 * (h * N + j) + 1] = (int8_t)j;
        }
    /* np.random.shuffle on the rows: randomimpl.py:1947-1954 */
    for (int i = n - 1; i > 0; i--) {
        int64_t k = rng_randint(rng, (int64_t)i + 1);
        int8_t t0 = sub[2 * i], t1 = sub[2 * i + 1];
        sub[2 * i] = sub[2 * k];
        sub[2 * i + 1] = sub[2 * k + 1];
        sub[2 * k] = t0;
        sub[2 * k + 1] = t1;
    }
    for (int i = 0; i < n && !*err; i++) {
        int h = sub[2 * i], j = sub[2 * i + 1];
        llk = base_step(c, rng, genotype, llk, h, j, n_alleles[j], temp, err);
    }
    free(sub);
    return llk;
}

/* ------------------------------------------------------------------------- */
/* assemble/structural.py                                                     */
/* ------------------------------------------------------------------------- */

/* assemble/structural.py:23-71 random_breaks. out i64[(breaks+1)*2] */
static int random_breaks(orc_rng *rng, int64_t breaks, int64_t n, int64_t *out)
{
    if (breaks >= n || n < 0)
        return ORC_ERR_BREAKS;
    uint8_t *ind = (uint8_t *)malloc((size_t)n + 1);
    int64_t *opts = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n + 1));
    for (int64_t i = 0; i <= n; i++)
        ind[i] = 1;
    ind[0] = 0;
    ind[n] = 0;
    for (int64_t b = 0; b < breaks; b++) {
        int64_t m = 0;
        for (int64_t i = 0; i <= n; i++)
            if (ind[i])
                opts[m++] = i;
        if (m == 0)
            break;
        int64_t point = opts[rng_randint(rng, m)]; /* np.random.choice(options) */
        ind[point] = 0;
    }
    int64_t m = 0;
    for (int64_t i = 0; i <= n; i++)
        if (!ind[i])
            opts[m++] = i;
    for (int64_t i = 0; i < breaks + 1; i++) {
        out[2 * i] = opts[i];
        out[2 * i + 1] = opts[i + 1];
    }
    free(ind);
    free(opts);
    return ORC_OK;
}

int orc_random_breaks(orc_rng *rng, int64_t breaks, int64_t n, int64_t *out)
{
    return random_breaks(rng, breaks, n, out);
}

/* assemble/structural.py:311-360 _label_haplotypes over the listed columns */
static void label_haplotypes(int8_t *labels /* stride 2 */, const int8_t *genotype, int P, int N,
                             const int *cols, int ncols)
{
    for (int j = 0; j < P; j++)
        labels[2 * j] = 0;
    for (int ci = 0; ci < ncols; ci++) {
        int i = cols[ci];
        for (int j = 1; j < P; j++) {
            if (genotype[(size_t)j * N + i] == genotype[(size_t)labels[2 * j] * N + i])
                continue;
            int prev = labels[2 * j];
            labels[2 * j] = (int8_t)j;
            for (int k = j + 1; k < P; k++)
                if (labels[2 * k] == prev &&
                    genotype[(size_t)j * N + i] == genotype[(size_t)k * N + i])
                    labels[2 * k] = (int8_t)j;
        }
    }
}

/* assemble/structural.py:394-430 haplotype_segment_labels.
 * has_interval == 0 reproduces interval=None. labels int8[P,2] */
void orc_haplotype_segment_labels(const int8_t *genotype, int P, int N, int has_interval,
                                  int start, int stop, int8_t *labels)
{
    int *cols = (int *)calloc((size_t)N + 1, sizeof(int));
    int n = 0;
    if (!has_interval) {
        start = 0;
        stop = N;
    }
    for (int i = start; i < stop; i++)
        cols[n++] = i;
    label_haplotypes(labels, genotype, P, N, cols, n);
    n = 0;
    if (has_interval)
        for (int i = 0; i < N; i++)
            if (!(i >= start && i < stop))
                cols[n++] = i;
    label_haplotypes(labels + 1, genotype, P, N, cols, n);
    free(cols);
}

/* assemble/structural.py:75-121 recombination_step_n_options */
int orc_recombination_step_n_options(const int8_t *labels, int P)
{
    int8_t dosage[256];
    haplotype_dosage(dosage, labels, P, 2, 0, 2);
    int n = 0;
    for (int h0 = 0; h0 < P; h0++) {
        if (dosage[h0] == 0)
            continue;
        for (int h1 = h0 + 1; h1 < P; h1++) {
            if (dosage[h1] == 0)
                continue;
            if (labels[2 * h0] == labels[2 * h1] || labels[2 * h0 + 1] == labels[2 * h1 + 1])
                continue;
            n++;
        }
    }
    return n;
}

/* assemble/structural.py:124-178 recombination_step_options. out int8[n,P,2] */
int orc_recombination_step_options(const int8_t *labels, int P, int8_t *out)
{
    int8_t dosage[256];
    haplotype_dosage(dosage, labels, P, 2, 0, 2);
    int opt = 0;
    for (int h0 = 0; h0 < P; h0++) {
        if (dosage[h0] == 0)
            continue;
        for (int h1 = h0 + 1; h1 < P; h1++) {
            if (dosage[h1] == 0)
                continue;
            if (labels[2 * h0] == labels[2 * h1] || labels[2 * h0 + 1] == labels[2 * h1 + 1])
                continue;
            int8_t *o = out + (size_t)opt * P * 2;
            memcpy(o, labels, (size_t)P * 2);
            o[2 * h0] = labels[2 * h1];
            o[2 * h1] = labels[2 * h0];
            opt++;
        }
    }
    return opt;
}

/* assemble/structural.py:182-236 dosage_step_n_options */
int orc_dosage_step_n_options(const int8_t *labels, int P)
{
    int8_t hd[256], sd[256];
    haplotype_dosage(hd, labels, P, 2, 0, 2);
    haplotype_dosage(sd, labels, P, 2, 0, 1);
    int n = 0;
    for (int h0 = 0; h0 < P; h0++) {
        if (hd[h0] == 0 || sd[h0] == 1)
            continue;
        for (int h1 = 0; h1 < P; h1++) {
            if (sd[h1] == 0)
                continue;
            if (labels[2 * h0] == labels[2 * h1])
                continue;
            n++;
        }
    }
    return n;
}

/* assemble/structural.py:239-307 dosage_step_options. out int8[n,P,2] */
int orc_dosage_step_options(const int8_t *labels, int P, int8_t *out)
{
    int8_t hd[256], sd[256];
    haplotype_dosage(hd, labels, P, 2, 0, 2);
    haplotype_dosage(sd, labels, P, 2, 0, 1);
    int opt = 0;
    for (int h0 = 0; h0 < P; h0++) {
        if (hd[h0] == 0 || sd[h0] == 1)
            continue;
        for (int h1 = 0; h1 < P; h1++) {
            if (sd[h1] == 0)
                continue;
            if (labels[2 * h0] == labels[2 * h1])
                continue;
            int8_t *o = out + (size_t)opt * P * 2;
            memcpy(o, labels, (size_t)P * 2);
            o[2 * h0] = labels[2 * h1];
            opt++;
        }
    }
    return opt;
}

/* assemble/structural.py:434-587 interval_step */
static double interval_step(asm_ctx *c, orc_rng *rng, int8_t *genotype, double llk, int start,
                            int stop, int step_type, double temp, int *err)
{
    int P = c->P, N = c->N;
    int8_t labels[512];
    orc_haplotype_segment_labels(genotype, P, N, 1, start, stop, labels);
    int max_opts = P * P;
    int8_t *options = (int8_t *)malloc((size_t)max_opts * P * 2 + 2);
    int n_options;
    if (step_type == 0)
        n_options = orc_recombination_step_options(labels, P, options);
    else if (step_type == 1)
        n_options = orc_dosage_step_options(labels, P, options);
    else {
        free(options);
        *err = ORC_ERR_STEP_TYPE;
        return llk;
    }
    if (n_options == 0) {
        free(options);
        return llk;
    }
    double log_proposal_prob = log(1.0 / (double)n_options);
    double lprior = 0.0;
    if (c->has_inbreeding)
        lprior = asm_prior_of_rows(c, genotype, N);
    double *llks = (double *)malloc(sizeof(double) * ((size_t)n_options + 1) * 3);
    double *log_accept = llks + n_options + 1;
    double *probs = log_accept + n_options + 1;
    llks[n_options] = -INFINITY;
    log_accept[n_options] = -INFINITY;
    int8_t hap_idx[256];
    for (int i = 0; i < n_options; i++) {
        const int8_t *opt = options + (size_t)i * P * 2;
        for (int h = 0; h < P; h++)
            hap_idx[h] = opt[2 * h];
        double llk_i;
        if (c->use_cache) { /* likelihood.py:239-305: key = the changed genotype */
            int8_t *gnew = (int8_t *)malloc((size_t)P * N);
            memcpy(gnew, genotype, (size_t)P * N);
            orc_structural_change(gnew, P, N, hap_idx, start, stop);
            llk_i = asm_llk_cached(c, gnew);
            free(gnew);
        } else {
            llk_i = orc_log_likelihood_structural_change(c->reads, c->U, N, c->A, genotype, P, hap_idx,
                                                         start, stop, c->counts);
            c->llk_evals++;
        }
        llks[i] = llk_i;
        double llk_ratio = llk_i - llk;
        double lprior_ratio = 0.0;
        if (c->has_inbreeding) {
            double lprior_i = asm_prior_of_rows(c, opt, 2); /* structural.py:546 */
            lprior_ratio = lprior_i - lprior;
        }
        int n_return = step_type == 0 ? orc_recombination_step_n_options(opt, P)
                                      : orc_dosage_step_n_options(opt, P);
        double log_return_prob = log(1.0 / (double)n_return);
        double lproposal_ratio = log_return_prob - log_proposal_prob;
        double mh_ratio = (llk_ratio + lprior_ratio) * temp + lproposal_ratio;
        log_accept[i] = np_minimum0(mh_ratio);
    }
    double ln_opts = log((double)n_options);
    double sum = 0.0;
    for (int i = 0; i <= n_options; i++) {
        log_accept[i] -= ln_opts;
        probs[i] = exp(log_accept[i]);
    }
    for (int i = 0; i <= n_options; i++)
        sum += probs[i];
    probs[n_options] = 1 - sum;
    int64_t choice = random_choice(rng, probs, n_options + 1);
    if (choice < n_options) {
        const int8_t *opt = options + (size_t)choice * P * 2;
        for (int h = 0; h < P; h++)
            hap_idx[h] = opt[2 * h];
        orc_structural_change(genotype, P, N, hap_idx, start, stop);
        llk = llks[choice];
    }
    free(llks);
    free(options);
    return llk;
}

/* assemble/structural.py:591-673 compound_step. intervals i64[n,2] */
static double structural_compound_step(asm_ctx *c, orc_rng *rng, int8_t *genotype, double llk,
                                       const int64_t *intervals, int n_intervals, int step_type,
                                       double temp, int *err)
{
    int64_t *perm = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_intervals > 0 ? n_intervals : 1));
    for (int i = 0; i < n_intervals; i++)
        perm[i] = i;
    orc_rng_shuffle_i64(rng, perm, n_intervals); /* np.random.permutation(np.arange(n)) */
    for (int i = 0; i < n_intervals && !*err; i++) {
        const int64_t *iv = intervals + 2 * perm[i];
        llk = interval_step(c, rng, genotype, llk, (int)iv[0], (int)iv[1], step_type, temp, err);
    }
    free(perm);
    return llk;
}

/* ------------------------------------------------------------------------- */
/* assemble/tempering.py                                                      */
/* ------------------------------------------------------------------------- */

/* assemble/tempering.py:11-58 chain_swap_acceptance */
double orc_chain_swap_acceptance(double llk_i, double lprior_i, double temp_i, double llk_j,
                                 double lprior_j, double temp_j)
{
    double post_i = llk_i + lprior_i;
    double post_j = llk_j + lprior_j;
    double frac_1 = (post_j - post_i) * temp_i;
    double frac_2 = (post_i - post_j) * temp_j;
    double a = exp(frac_1 + frac_2);
    if (a > 1.0)
        a = 1.0;
    return a;
}

/* assemble/tempering.py:62-151 chain_swap_step. Updates llk_i/llk_j in place */
static void chain_swap_step(asm_ctx *c, orc_rng *rng, int8_t *gi, double *llk_i, double temp_i,
                            int8_t *gj, double *llk_j, double temp_j)
{
    int P = c->P, N = c->N;
    double prior_i = 0.0, prior_j = 0.0;
    if (c->has_inbreeding) {
        prior_i = asm_prior_of_rows(c, gi, N);
        prior_j = asm_prior_of_rows(c, gj, N);
    }
    double acc = orc_chain_swap_acceptance(*llk_i, prior_i, temp_i, *llk_j, prior_j, temp_j);
    double val = rng_double(rng);
    if (acc >= val) {
        size_t n = (size_t)P * N;
        int8_t *tmp = (int8_t *)malloc(n);
        memcpy(tmp, gi, n);
        memcpy(gi, gj, n);
        memcpy(gj, tmp, n);
        free(tmp);
        double t = *llk_i;
        *llk_i = *llk_j;
        *llk_j = t;
    }
}

/* ------------------------------------------------------------------------- */
/* exported single-step wrappers (used by the unit / golden tests)            */
/* ------------------------------------------------------------------------- */

static void ctx_init(asm_ctx *c, const double *reads, int U, int N, int A, const int64_t *counts,
                     int P, double inbreeding /* NaN => None */, double log_unique_haplotypes)
{
    c->reads = reads;
    c->U = U;
    c->N = N;
    c->A = A;
    c->counts = counts;
    c->P = P;
    c->has_inbreeding = !isnan(inbreeding);
    c->inbreeding = inbreeding;
    c->log_unique_haplotypes = log_unique_haplotypes;
    c->llk_evals = 0;
    asm_cache_init(c, 0);
}

double orc_mutation_base_step(orc_rng *rng, int8_t *genotype, int P, int N, const double *reads,
                              int U, int A, const int64_t *counts, double llk, int h, int j,
                              int n_alleles, double log_unique_haplotypes, double inbreeding,
                              double temp, int *err)
{
    asm_ctx c;
    ctx_init(&c, reads, U, N, A, counts, P, inbreeding, log_unique_haplotypes);
    *err = 0;
    return base_step(&c, rng, genotype, llk, h, j, n_alleles, temp, err);
}

double orc_mutation_compound_step(orc_rng *rng, int8_t *genotype, int P, int N,
                                  const double *reads, int U, int A, const int64_t *counts,
                                  double llk, const int8_t *n_alleles,
                                  double log_unique_haplotypes, double inbreeding, double temp,
                                  int *err)
{
    asm_ctx c;
    ctx_init(&c, reads, U, N, A, counts, P, inbreeding, log_unique_haplotypes);
    *err = 0;
    return mutation_compound_step(&c, rng, genotype, llk, n_alleles, temp, err);
}

double orc_structural_interval_step(orc_rng *rng, int8_t *genotype, int P, int N,
                                    const double *reads, int U, int A, const int64_t *counts,
                                    double llk, int start, int stop, int step_type,
                                    double log_unique_haplotypes, double inbreeding, double temp,
                                    int *err)
{
    asm_ctx c;
    ctx_init(&c, reads, U, N, A, counts, P, inbreeding, log_unique_haplotypes);
    *err = 0;
    return interval_step(&c, rng, genotype, llk, start, stop, step_type, temp, err);
}

double orc_structural_compound_step(orc_rng *rng, int8_t *genotype, int P, int N,
                                    const double *reads, int U, int A, const int64_t *counts,
                                    double llk, const int64_t *intervals, int n_intervals,
                                    int step_type, double log_unique_haplotypes,
                                    double inbreeding, double temp, int *err)
{
    asm_ctx c;
    ctx_init(&c, reads, U, N, A, counts, P, inbreeding, log_unique_haplotypes);
    *err = 0;
    return structural_compound_step(&c, rng, genotype, llk, intervals, n_intervals, step_type,
                                    temp, err);
}

/* ------------------------------------------------------------------------- */
/* assemble/mcmc.py                                                           */
/* ------------------------------------------------------------------------- */

/* assemble/mcmc.py:294 `np.log(n_alleles).sum()` with n_alleles an int8 array: numba resolves
 * the ufunc loop to float32 (int8 -> 'f'), and ndarray.sum() accumulates in the array dtype,
 * so the reference's log_unique_haplotypes is a float32 quantity (verified against numba;
 * fixture luh_*). */
double orc_log_unique_haplotypes(const int8_t *n_alleles, int N)
{
    float s = 0.0f;
    for (int j = 0; j < N; j++)
        s += logf((float)n_alleles[j]);
    return (double)s;
}

/* assemble/mcmc.py:269-426 _denovo_assembler (return_heated_trace=False).
 * genotype int8[P,N] initial state (not modified); n_alleles int8[N];
 * break_dist f64[n_break_dist]; temperatures ascending f64[T];
 * out_genotypes int8[steps,P,N]; out_llks f64[steps]. Returns error code. */
int orc_denovo_assembler(orc_rng *rng, const int8_t *genotype, int P, int N, const double *reads,
                         int U, int A, const int64_t *counts, const int8_t *n_alleles,
                         double inbreeding, int steps, const double *break_dist,
                         int n_break_dist, double p_recomb, double p_partial, double p_dosage,
                         const double *temperatures, int T, int8_t *out_genotypes,
                         double *out_llks, int64_t *out_llk_evals)
{
    asm_ctx c;
    double log_unique_haplotypes = orc_log_unique_haplotypes(n_alleles, N); /* mcmc.py:294 */
    ctx_init(&c, reads, U, N, A, counts, P, inbreeding, log_unique_haplotypes);
    /* assemble/mcmc.py:305-312 */
    if (orc_llk_cache_threshold >= 0 && (int64_t)P * N * U > orc_llk_cache_threshold)
        asm_cache_init(&c, 1);
    size_t gsz = (size_t)P * N;
    int8_t *genotypes = (int8_t *)malloc(gsz * (size_t)T + 1);
    double *llks = (double *)malloc(sizeof(double) * (size_t)T);
    int64_t *intervals = (int64_t *)malloc(sizeof(int64_t) * 2 * ((size_t)N + 2));
    double llk0 = orc_log_likelihood(reads, U, N, A, genotype, P, counts);
    for (int t = 0; t < T; t++) {
        memcpy(genotypes + gsz * t, genotype, gsz);
        llks[t] = llk0;
    }
    int err = 0;
    for (int i = 0; i < steps && !err; i++) {
        int8_t *g = NULL;
        double llk = 0.0;
        for (int t = 0; t < T && !err; t++) {
            llk = llks[t];
            g = genotypes + gsz * t;
            double temp = temperatures[t];
            if (isnan(llk)) {
                err = ORC_ERR_NAN_LLK;
                break;
            }
            llk = mutation_compound_step(&c, rng, g, llk, n_alleles, temp, &err);
            if (err)
                break;
            if (rng_double(rng) <= p_recomb) {
                int64_t n_breaks = random_choice(rng, break_dist, n_break_dist);
                err = random_breaks(rng, n_breaks, N, intervals);
                if (err)
                    break;
                llk = structural_compound_step(&c, rng, g, llk, intervals, (int)n_breaks + 1, 0,
                                               temp, &err);
                if (err)
                    break;
            }
            if (rng_double(rng) <= p_partial) {
                int64_t n_breaks = random_choice(rng, break_dist, n_break_dist);
                err = random_breaks(rng, n_breaks, N, intervals);
                if (err)
                    break;
                llk = structural_compound_step(&c, rng, g, llk, intervals, (int)n_breaks + 1, 1,
                                               temp, &err);
                if (err)
                    break;
            }
            if (rng_double(rng) <= p_dosage) {
                int64_t full[2] = {0, N};
                llk = structural_compound_step(&c, rng, g, llk, full, 1, 1, temp, &err);
                if (err)
                    break;
            }
            if (t > 0) {
                double llk_prev = llks[t - 1];
                chain_swap_step(&c, rng, g, &llk, temp, genotypes + gsz * (t - 1), &llk_prev,
                                temperatures[t - 1]);
                llks[t - 1] = llk_prev;
            }
            llks[t] = llk;
        }
        if (err)
            break;
        memcpy(out_genotypes + gsz * i, g, gsz);
        out_llks[i] = llk;
    }
    if (!err && rng->exhausted)
        err = ORC_ERR_RNG_EXHAUSTED;
    if (out_llk_evals)
        *out_llk_evals = c.llk_evals;
    asm_cache_free(&c);
    free(genotypes);
    free(llks);
    free(intervals);
    return err;
}

/* assemble/snpcalling.py:14-70 snp_posterior for position `pos` of reads[U,N,A].
 * probs f64[comb_with_replacement(n_alleles, ploidy)] */
static void snp_posterior(const double *reads, int U, int N, int A, int pos, int n_alleles,
                          int ploidy, int has_inbreeding, double inbreeding,
                          const int64_t *counts, double *probs, int64_t u_gens)
{
    int64_t g[64];
    int8_t g8[64];
    double *col = (double *)malloc(sizeof(double) * (size_t)(U > 0 ? U : 1) * A);
    int Ueff = U;
    if (U == 0) { /* snpcalling.py:41-45 */
        Ueff = 1;
        for (int a = 0; a < A; a++)
            col[a] = NAN;
    } else {
        for (int r = 0; r < U; r++)
            for (int a = 0; a < A; a++)
                col[(size_t)r * A + a] = reads[((size_t)r * N + pos) * A + a];
    }
    double *lp = (double *)malloc(sizeof(double) * (size_t)(u_gens > 0 ? u_gens : 1));
    for (int i = 0; i < ploidy; i++)
        g[i] = 0;
    for (int64_t i = 0; i < u_gens; i++) {
        double lprior = 0.0;
        if (has_inbreeding)
            lprior = orc_calling_log_genotype_prior(g, ploidy, n_alleles, inbreeding, NULL);
        for (int k = 0; k < ploidy; k++)
            g8[k] = (int8_t)g[k];
        double llk = orc_log_likelihood(col, Ueff, 1, A, g8, ploidy, U == 0 ? NULL : counts);
        lp[i] = lprior + llk;
        orc_increment_genotype(g, ploidy);
    }
    orc_normalise_log_probs(lp, u_gens, probs);
    free(lp);
    free(col);
}

/* assemble/mcmc.py:495-541 _homozygosity_probabilities -> f64[N,A] */
void orc_homozygosity_probabilities(const double *reads, int U, int N, int A,
                                    const int8_t *n_alleles, int ploidy, double inbreeding,
                                    const int64_t *counts, double *out)
{
    int has_inb = !isnan(inbreeding);
    for (int i = 0; i < N * A; i++)
        out[i] = 0.0;
    int64_t g[64];
    for (int i = 0; i < N; i++) {
        int n = n_alleles[i];
        int64_t u_gens = orc_comb_with_replacement(n, ploidy);
        double *probs = (double *)malloc(sizeof(double) * (size_t)(u_gens > 0 ? u_gens : 1));
        snp_posterior(reads, U, N, A, i, n, ploidy, has_inb, inbreeding, counts, probs, u_gens);
        for (int a = 0; a < n; a++) {
            for (int k = 0; k < ploidy; k++)
                g[k] = a;
            int64_t idx = orc_genotype_alleles_as_index(g, ploidy);
            out[(size_t)i * A + a] = probs[idx];
        }
        free(probs);
    }
}

/* assemble/mcmc.py:455-491 _read_mean_dist (numpy semantics restated:
 * nanmean over axis 0 = in-order sum of the non-NaN entries / their count;
 * row sums over the short last axis are sequential). reads f64[U,N,A] -> f64[N,A] */
void orc_read_mean_dist(const double *reads, int U, int N, int A, double *dist)
{
    for (int j = 0; j < N; j++) {
        int n_nonzero_alleles = 0;
        int gap[64];
        for (int a = 0; a < A; a++) {
            int all_nan = 1;
            for (int r = 0; r < U; r++)
                if (!isnan(reads[((size_t)r * N + j) * A + a])) {
                    all_nan = 0;
                    break;
                }
            gap[a] = all_nan;
            double tot = 0.0;
            int cnt = 0;
            int all_zero = 1;
            for (int r = 0; r < U; r++) {
                double v = all_nan ? 1.0 : reads[((size_t)r * N + j) * A + a];
                if (!isnan(v)) {
                    tot += v;
                    cnt++;
                }
                if (!(v == 0.0))
                    all_zero = 0;
            }
            dist[(size_t)j * A + a] = tot / (double)cnt;
            if (!all_zero)
                n_nonzero_alleles++;
        }
        for (int a = 0; a < A; a++)
            if (gap[a])
                dist[(size_t)j * A + a] = 1.0 / (double)n_nonzero_alleles;
        double s = 0.0;
        for (int a = 0; a < A; a++)
            s += dist[(size_t)j * A + a];
        for (int a = 0; a < A; a++)
            dist[(size_t)j * A + a] /= s;
    }
}

/* jitutils.py:465-498 sample_snv_alleles for one haplotype: dist f64[N,A] -> int8[N] */
static void sample_snv_alleles(orc_rng *rng, const double *dist, int N, int A, int8_t *out)
{
    double d[64];
    for (int j = 0; j < N; j++) {
        double s = 0.0;
        for (int a = 0; a < A; a++)
            s += dist[(size_t)j * A + a];
        for (int a = 0; a < A; a++)
            d[a] = dist[(size_t)j * A + a] / s;
        out[j] = (int8_t)random_choice(rng, d, A);
    }
}

/* assemble/mcmc.py:103-265 DenovoMCMC.fit + _mcmc.
 * reads f64[U,N,A]; n_alleles int8[N]; temperatures must be sorted ascending;
 * break_table f64[(N+1) * break_stride]: row n = break distribution to use when
 * n_het == n, with break_len[n] entries (the host computes the rows with scipy
 * exactly as _point_beta_probabilities 429-452, or the n_intervals hack 214-217);
 * initial int8[C,P,initial_nhet] or NULL;
 * out_genotypes int8[C,S,P,N]; out_llks f64[C,S]; out_nhet: number of
 * non-fixed positions; returns error code. */
int orc_denovo_fit(uint32_t seed, const uint32_t *replay_words, int64_t n_replay,
                   const double *reads_in, int U, int N, int A, const int64_t *counts,
                   const int8_t *n_alleles, int P, double inbreeding, int steps, int chains,
                   double fix_homozygous, const double *break_table, const int32_t *break_len,
                   int break_stride, double p_recomb, double p_partial, double p_dosage,
                   const double *temperatures, int T, const int8_t *initial, int initial_nhet,
                   int8_t *out_genotypes, double *out_llks, int32_t *out_nhet,
                   int64_t *out_words, int64_t *out_llk_evals)
{
    orc_rng rng;
    orc_rng_seed(&rng, seed);
    if (replay_words)
        orc_rng_replay(&rng, replay_words, n_replay);
    const double *reads = reads_in;
    double *mock = NULL;
    if (U == 0) { /* mcmc.py:132-137 */
        mock = (double *)malloc(sizeof(double) * (size_t)N * A + 8);
        for (int i = 0; i < N * A; i++)
            mock[i] = NAN;
        reads = mock;
        U = 1;
        counts = NULL; /* log(1) * count == 0 whatever the (out of range) count is */
    }
    int err = 0;
    int64_t evals = 0;
    double *hom = (double *)malloc(sizeof(double) * ((size_t)N * A + 1));
    int *het_idx = (int *)malloc(sizeof(int) * ((size_t)N + 1));
    int8_t *fixed_allele = (int8_t *)malloc((size_t)N + 1);
    size_t step_sz = (size_t)P * N;
    for (int chain = 0; chain < chains && !err; chain++) {
        int8_t *og = out_genotypes + (size_t)chain * steps * step_sz;
        double *ol = out_llks + (size_t)chain * steps;
        /* mcmc.py:165-186 */
        orc_homozygosity_probabilities(reads, U, N, A, n_alleles, P, inbreeding, counts, hom);
        int n_het = 0;
        for (int j = 0; j < N; j++) {
            int any = 0;
            fixed_allele[j] = 0;
            for (int a = 0; a < A; a++)
                if (hom[(size_t)j * A + a] >= fix_homozygous) {
                    any = 1;
                    fixed_allele[j] = (int8_t)a; /* np.where order: last wins */
                }
            if (!any)
                het_idx[n_het++] = j;
        }
        if (out_nhet)
            *out_nhet = n_het;
        if (n_het == 0) { /* mcmc.py:188-199 */
            for (int s = 0; s < steps; s++) {
                for (int h = 0; h < P; h++)
                    for (int j = 0; j < N; j++)
                        og[(size_t)s * step_sz + (size_t)h * N + j] = fixed_allele[j];
                ol[s] = NAN;
            }
            continue;
        }
        /* reads_het = reads[:, heterozygous] */
        double *reads_het = (double *)malloc(sizeof(double) * (size_t)U * n_het * A);
        for (int r = 0; r < U; r++)
            for (int k = 0; k < n_het; k++)
                for (int a = 0; a < A; a++)
                    reads_het[((size_t)r * n_het + k) * A + a] =
                        reads[((size_t)r * N + het_idx[k]) * A + a];
        int8_t *na_het = (int8_t *)malloc((size_t)n_het);
        for (int k = 0; k < n_het; k++)
            na_het[k] = n_alleles[het_idx[k]];
        int8_t *genotype = (int8_t *)malloc((size_t)P * n_het);
        if (!initial) { /* mcmc.py:202-204 */
            double *dist = (double *)malloc(sizeof(double) * (size_t)n_het * A);
            orc_read_mean_dist(reads_het, U, n_het, A, dist);
            for (int h = 0; h < P; h++)
                sample_snv_alleles(&rng, dist, n_het, A, genotype + (size_t)h * n_het);
            free(dist);
        } else {
            if (initial_nhet != n_het) {
                err = ORC_ERR_INITIAL_SHAPE;
            } else {
                memcpy(genotype, initial + (size_t)chain * P * n_het, (size_t)P * n_het);
            }
        }
        if (!err) {
            int8_t *tg = (int8_t *)malloc((size_t)steps * P * n_het + 1);
            int64_t ev = 0;
            err = orc_denovo_assembler(&rng, genotype, P, n_het, reads_het, U, A, counts, na_het,
                                       inbreeding, steps, break_table + (size_t)n_het * break_stride,
                                       break_len[n_het], p_recomb, p_partial, p_dosage,
                                       temperatures, T, tg, ol, &ev);
            evals += ev;
            /* mcmc.py:255-265: re-insert the fixed alleles */
            if (!err)
                for (int s = 0; s < steps; s++)
                    for (int h = 0; h < P; h++) {
                        int8_t *row = og + (size_t)s * step_sz + (size_t)h * N;
                        for (int j = 0; j < N; j++)
                            row[j] = fixed_allele[j];
                        for (int k = 0; k < n_het; k++)
                            row[het_idx[k]] = tg[((size_t)s * P + h) * n_het + k];
                    }
            free(tg);
        }
        free(genotype);
        free(na_het);
        free(reads_het);
    }
    if (out_words)
        *out_words = rng.words;
    if (out_llk_evals)
        *out_llk_evals = evals;
    free(hom);
    free(het_idx);
    free(fixed_allele);
    free(mock);
    return err;
}

/* ------------------------------------------------------------------------- */
/* calling/likelihood.py + calling/mcmc.py                                    */
/* ------------------------------------------------------------------------- */

typedef struct {
    const double *reads;
    int U, N, A;
    const int64_t *counts;
    const int8_t *haplotypes; /* int8[H,N] */
    int H;
    int has_prior;
    double inbreeding;
    const double *freqs; /* or NULL */
    /* dict cache keyed on the rank of the sorted alleles (calling/likelihood.py:67-77) */
    int use_cache;
    int64_t cache_cap;
    int64_t *cache_keys;
    double *cache_vals;
    int64_t cache_n;
    int64_t llk_evals;
} call_ctx;

/* calling/likelihood.py:8-33 log_likelihood_alleles */
static double llk_alleles(call_ctx *c, const int64_t *g, int P)
{
    int8_t *geno = (int8_t *)malloc((size_t)P * c->N + 1);
    for (int k = 0; k < P; k++)
        memcpy(geno + (size_t)k * c->N, c->haplotypes + (size_t)g[k] * c->N, (size_t)c->N);
    double v = orc_log_likelihood(c->reads, c->U, c->N, c->A, geno, P, c->counts);
    c->llk_evals++;
    free(geno);
    return v;
}

static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

static void cache_grow(call_ctx *c)
{
    int64_t ncap = c->cache_cap ? c->cache_cap * 2 : 1024;
    int64_t *nk = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncap);
    double *nv = (double *)malloc(sizeof(double) * (size_t)ncap);
    for (int64_t i = 0; i < ncap; i++)
        nk[i] = -1;
    for (int64_t i = 0; i < c->cache_cap; i++) {
        if (c->cache_keys[i] < 0)
            continue;
        uint64_t hsh = (uint64_t)c->cache_keys[i] * 0x9E3779B97F4A7C15ULL;
        int64_t s = (int64_t)(hsh >> 20) & (ncap - 1);
        while (nk[s] >= 0)
            s = (s + 1) & (ncap - 1);
        nk[s] = c->cache_keys[i];
        nv[s] = c->cache_vals[i];
    }
    free(c->cache_keys);
    free(c->cache_vals);
    c->cache_keys = nk;
    c->cache_vals = nv;
    c->cache_cap = ncap;
}

/* calling/likelihood.py:36-78 log_likelihood_alleles_cached */
static double llk_alleles_cached(call_ctx *c, const int64_t *g, int P)
{
    if (!c->use_cache)
        return llk_alleles(c, g, P);
    int64_t sorted[64];
    memcpy(sorted, g, sizeof(int64_t) * (size_t)P);
    qsort(sorted, (size_t)P, sizeof(int64_t), cmp_i64);
    int64_t key = orc_genotype_alleles_as_index(sorted, P);
    if (c->cache_n * 2 >= c->cache_cap)
        cache_grow(c);
    uint64_t hsh = (uint64_t)key * 0x9E3779B97F4A7C15ULL;
    int64_t s = (int64_t)(hsh >> 20) & (c->cache_cap - 1);
    while (c->cache_keys[s] >= 0) {
        if (c->cache_keys[s] == key)
            return c->cache_vals[s];
        s = (s + 1) & (c->cache_cap - 1);
    }
    double v = llk_alleles(c, g, P);
    c->cache_keys[s] = key;
    c->cache_vals[s] = v;
    c->cache_n++;
    return v;
}

static void cache_reset(call_ctx *c)
{
    free(c->cache_keys);
    free(c->cache_vals);
    c->cache_keys = NULL;
    c->cache_vals = NULL;
    c->cache_cap = 0;
    c->cache_n = 0;
}

static double call_prior(const call_ctx *c, const int64_t *g, int P)
{
    return orc_calling_log_genotype_prior(g, P, c->H, c->inbreeding, c->freqs);
}

/* calling/mcmc.py:143-229 gibbs_options */
static void gibbs_options(call_ctx *c, int64_t *g, int P, int k, double *llks, double *lpriors,
                          double *probs)
{
    int64_t current = g[k];
    int H = c->H;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)H);
    for (int a = 0; a < H; a++) {
        g[k] = a;
        if (!c->has_prior)
            lpriors[a] = orc_log_genotype_allele_flat_prior(g, P, k);
        else
            lpriors[a] = orc_log_genotype_allele_prior(g, P, k, H, c->inbreeding, c->freqs);
        llks[a] = llk_alleles_cached(c, g, P);
    }
    for (int a = 0; a < H; a++)
        tmp[a] = llks[a] + lpriors[a];
    orc_normalise_log_probs(tmp, H, probs);
    g[k] = current;
    free(tmp);
}

/* calling/mcmc.py:15-140 mh_options */
static void mh_options(call_ctx *c, int64_t *g, int P, int k, double *llks, double *lpriors,
                       double *probs)
{
    int H = c->H;
    int64_t current = g[k];
    int copies = count_allele(g, P, g[k]);
    double lprior = c->has_prior ? call_prior(c, g, P) : 0.0;
    double llk = llk_alleles_cached(c, g, P);
    double *lprop = (double *)malloc(sizeof(double) * (size_t)H);
    for (int a = 0; a < H; a++) {
        if (g[k] == a) {
            lprop[a] = 0.0;
            lpriors[a] = lprior;
            llks[a] = llk;
        } else {
            g[k] = a;
            lpriors[a] = c->has_prior ? call_prior(c, g, P) : 0.0;
            llks[a] = llk_alleles_cached(c, g, P);
            int copies_i = count_allele(g, P, g[k]);
            lprop[a] = log((double)copies_i / (double)copies);
        }
    }
    /* note: after the first proposal g[k] != current, so `g[k] == a` (mcmc.py:93)
     * compares against the previously proposed allele, exactly as the reference */
    for (int a = 0; a < H; a++) {
        double mh = (llks[a] - llk) + (lpriors[a] - lprior) + lprop[a];
        probs[a] = exp(np_minimum0(mh));
    }
    probs[current] = 0;
    for (int a = 0; a < H; a++)
        probs[a] /= (double)(H - 1);
    double s = 0.0;
    for (int a = 0; a < H; a++)
        s += probs[a];
    probs[current] = 1 - s;
    g[k] = current;
    free(lprop);
}

/* calling/mcmc.py:232-327 compound_step */
static double call_compound_step(call_ctx *c, orc_rng *rng, int64_t *g, int P, int step_type,
                                 int *err)
{
    int H = c->H;
    double *llks = (double *)malloc(sizeof(double) * (size_t)H * 3);
    double *lpriors = llks + H;
    double *probs = lpriors + H;
    int64_t order[64];
    for (int i = 0; i < P; i++)
        order[i] = i;
    orc_rng_shuffle_i64(rng, order, P);
    int64_t choice = 0;
    for (int j = 0; j < P; j++) {
        int k = (int)order[j];
        if (step_type == 0)
            gibbs_options(c, g, P, k, llks, lpriors, probs);
        else
            mh_options(c, g, P, k, llks, lpriors, probs);
        choice = random_choice(rng, probs, H);
        if (choice >= H) {
            *err = ORC_ERR_CHOICE_RANGE;
            choice = g[k];
        }
        g[k] = choice;
    }
    qsort(g, (size_t)P, sizeof(int64_t), cmp_i64);
    double out = llks[choice];
    free(llks);
    return out;
}

/* calling/mcmc.py:393-453 greedy_caller */
static void greedy_caller(call_ctx *c, int P, int64_t *out)
{
    int64_t g[64];
    for (int i = 0; i < P; i++) {
        int k = i + 1;
        double best = -INFINITY;
        int64_t best_a = -1;
        for (int a = 0; a < c->H; a++) {
            g[i] = a;
            double llk = llk_alleles(c, g, k);
            double lprior = c->has_prior ? call_prior(c, g, k) : 0.0;
            double lprob = llk + lprior;
            if (lprob > best) {
                best = lprob;
                best_a = a;
            }
        }
        g[i] = best_a;
    }
    qsort(g, (size_t)P, sizeof(int64_t), cmp_i64);
    for (int i = 0; i < P; i++)
        out[i] = g[i];
}

static void call_ctx_init(call_ctx *c, const double *reads, int U, int N, int A,
                          const int64_t *counts, const int8_t *haplotypes, int H,
                          double inbreeding /* NaN => prior None */, const double *freqs)
{
    memset(c, 0, sizeof(*c));
    c->reads = reads;
    c->U = U;
    c->N = N;
    c->A = A;
    c->counts = counts;
    c->haplotypes = haplotypes;
    c->H = H;
    c->has_prior = !isnan(inbreeding);
    c->inbreeding = inbreeding;
    c->freqs = freqs;
}

void orc_greedy_caller(const double *reads, int U, int N, int A, const int64_t *counts,
                       const int8_t *haplotypes, int H, int P, double inbreeding,
                       const double *freqs, int64_t *out)
{
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, inbreeding, freqs);
    greedy_caller(&c, P, out);
}

/* probabilities of one Gibbs / MH sub-step (for the transition-matrix tests) */
void orc_calling_step_options(const double *reads, int U, int N, int A, const int64_t *counts,
                              const int8_t *haplotypes, int H, int64_t *g, int P, int k,
                              double inbreeding, const double *freqs, int step_type,
                              double *llks, double *lpriors, double *probs)
{
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, inbreeding, freqs);
    if (step_type == 0)
        gibbs_options(&c, g, P, k, llks, lpriors, probs);
    else
        mh_options(&c, g, P, k, llks, lpriors, probs);
}

/* calling/classes.py:49-124 CallingMCMC.fit + calling/mcmc.py:330-390 mcmc_sampler.
 * initial i64[P] or NULL; out_alleles i64[C,S,P]; out_llks f64[C,S] */
int orc_calling_fit(uint32_t seed, const uint32_t *replay_words, int64_t n_replay,
                    const double *reads, int U, int N, int A, const int64_t *counts,
                    const int8_t *haplotypes, int H, int P, double inbreeding,
                    const double *freqs, int steps, int chains, int step_type,
                    const int64_t *initial, int64_t *out_alleles, double *out_llks,
                    int64_t *out_words, int64_t *out_llk_evals)
{
    orc_rng rng;
    orc_rng_seed(&rng, seed);
    if (replay_words)
        orc_rng_replay(&rng, replay_words, n_replay);
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, inbreeding, freqs);
    int64_t init[64], g[64];
    int err = 0;
    if (initial)
        memcpy(init, initial, sizeof(int64_t) * (size_t)P);
    else
        greedy_caller(&c, P, init);
    for (int ch = 0; ch < chains; ch++) {
        memcpy(g, init, sizeof(int64_t) * (size_t)P);
        c.use_cache = 1; /* classes.py:117 cache=True: fresh dict per chain */
        cache_reset(&c);
        for (int s = 0; s < steps; s++) {
            double llk = call_compound_step(&c, &rng, g, P, step_type, &err);
            out_llks[(size_t)ch * steps + s] = llk;
            memcpy(out_alleles + ((size_t)ch * steps + s) * P, g, sizeof(int64_t) * (size_t)P);
        }
    }
    cache_reset(&c);
    if (!err && rng.exhausted)
        err = ORC_ERR_RNG_EXHAUSTED;
    if (out_words)
        *out_words = rng.words;
    if (out_llk_evals)
        *out_llk_evals = c.llk_evals;
    return err;
}

/* ------------------------------------------------------------------------- */
/* calling/exact.py                                                           */
/* ------------------------------------------------------------------------- */

/* calling/exact.py:252-263 _genotype_likelihoods (float32 output, 254) */
void orc_genotype_likelihoods(const double *reads, int U, int N, int A, const int64_t *counts,
                              const int8_t *haplotypes, int H, int P, int64_t n_genotypes,
                              float *out)
{
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, NAN, NULL);
    int64_t g[64];
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        out[i] = (float)llk_alleles(&c, g, P);
        orc_increment_genotype(g, P);
    }
}

/* same enumeration in full precision (not in the reference API; used to check
 * the device's fp64 table before the float32 rounding) */
void orc_genotype_likelihoods_f64(const double *reads, int U, int N, int A,
                                  const int64_t *counts, const int8_t *haplotypes, int H, int P,
                                  int64_t n_genotypes, double *out)
{
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, NAN, NULL);
    int64_t g[64];
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        out[i] = llk_alleles(&c, g, P);
        orc_increment_genotype(g, P);
    }
}

/* calling/exact.py:295-329 genotype_posteriors for a float32 llk array */
void orc_genotype_posteriors_f32(const float *llks, int64_t n_genotypes, int P, int64_t n_alleles,
                                 double inbreeding, const double *freqs, double *out)
{
    int has_prior = !isnan(inbreeding);
    float *post = (float *)malloc(sizeof(float) * (size_t)(n_genotypes > 0 ? n_genotypes : 1));
    int64_t g[64];
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        float llk = llks[i];
        double lpr = 0.0;
        if (has_prior)
            lpr = orc_calling_log_genotype_prior(g, P, n_alleles, inbreeding, freqs);
        post[i] = (float)((double)llk + lpr);
        orc_increment_genotype(g, P);
    }
    /* numba unifies the accumulator of sum_log_probs to float64 (add_log_prob returns the
     * float64 constant -inf on one path), so only the STORED values are float32-rounded;
     * the log-sum-exp and the final exp run in float64 (verified against the reference). */
    double acc = (double)post[0];
    for (int64_t i = 1; i < n_genotypes; i++)
        acc = orc_add_log_prob(acc, (double)post[i]);
    for (int64_t i = 0; i < n_genotypes; i++)
        out[i] = exp((double)post[i] - acc);
    free(post);
}

/* calling/exact.py:295-329 genotype_posteriors for a float64 llk array */
void orc_genotype_posteriors_f64(const double *llks, int64_t n_genotypes, int P,
                                 int64_t n_alleles, double inbreeding, const double *freqs,
                                 double *out)
{
    int has_prior = !isnan(inbreeding);
    double *post = (double *)malloc(sizeof(double) * (size_t)(n_genotypes > 0 ? n_genotypes : 1));
    int64_t g[64];
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        double lpr = 0.0;
        if (has_prior)
            lpr = orc_calling_log_genotype_prior(g, P, n_alleles, inbreeding, freqs);
        post[i] = llks[i] + lpr;
        orc_increment_genotype(g, P);
    }
    orc_normalise_log_probs(post, n_genotypes, out);
    free(post);
}

/* calling/exact.py:332-369 posterior_allele_frequencies */
void orc_posterior_allele_frequencies(const double *posteriors, int64_t n_genotypes, int P,
                                      int64_t n_alleles, double *freqs, double *counts_out,
                                      double *occur)
{
    int64_t g[64];
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t a = 0; a < n_alleles; a++) {
        counts_out[a] = 0.0;
        occur[a] = 0.0;
    }
    for (int64_t i = 0; i < n_genotypes; i++) {
        double p = posteriors[i];
        for (int j = 0; j < P; j++) {
            int64_t a = g[j];
            counts_out[a] += p;
            if (j == 0)
                occur[a] += p;
            else if (a != g[j - 1])
                occur[a] += p;
        }
        orc_increment_genotype(g, P);
    }
    for (int64_t a = 0; a < n_alleles; a++)
        freqs[a] = counts_out[a] / (double)P;
}

/* itertools.combinations_with_replacement(support, r) in lexicographic order,
 * used by calling/exact.py:64-105 and 372-407 */
static int cwr_next(int *idx, int r, int n)
{
    int i = r - 1;
    while (i >= 0 && idx[i] == n - 1)
        i--;
    if (i < 0)
        return 0;
    int v = idx[i] + 1;
    for (int k = i; k < r; k++)
        idx[k] = v;
    return 1;
}

/* calling/exact.py:156-249 posterior_mode (all optional outputs computed).
 * out_mode i64[P]; out_freqs / out_occur f64[H]. */
void orc_posterior_mode(const double *reads, int U, int N, int A, const int64_t *counts,
                        const int8_t *haplotypes, int H, int P, int64_t n_genotypes,
                        double inbreeding, const double *freqs_prior, int64_t *out_mode,
                        double *out_mode_llk, double *out_mode_prob, double *out_support_prob,
                        double *out_freqs, double *out_occur)
{
    call_ctx c;
    call_ctx_init(&c, reads, U, N, A, counts, haplotypes, H, inbreeding, freqs_prior);
    int64_t g[64];
    /* exact.py:17-61 _call_posterior_mode */
    int64_t mode_idx = 0;
    double mode_llk = -INFINITY, mode_ljoint = -INFINITY, total_ljoint = -INFINITY;
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        double llk = llk_alleles(&c, g, P);
        double lpr = c.has_prior ? call_prior(&c, g, P) : 0.0;
        double ljoint = llk + lpr;
        if (ljoint > mode_ljoint) {
            mode_idx = i;
            mode_llk = llk;
            mode_ljoint = ljoint;
        }
        total_ljoint = orc_add_log_prob(total_ljoint, ljoint);
        orc_increment_genotype(g, P);
    }
    orc_index_as_genotype_alleles(mode_idx, P, out_mode);
    *out_mode_llk = mode_llk;
    *out_mode_prob = exp(mode_ljoint - total_ljoint);
    /* exact.py:64-105 _genotype_support_log_joint */
    {
        int64_t support[64], tmp[64];
        int ns = 0;
        for (int i = 0; i < P; i++) { /* np.unique of a sorted genotype */
            if (i == 0 || out_mode[i] != out_mode[i - 1])
                support[ns++] = out_mode[i];
        }
        int rem = P - ns;
        int idx[64];
        for (int i = 0; i < rem; i++)
            idx[i] = 0;
        double support_ljoint = -INFINITY;
        do {
            for (int i = 0; i < ns; i++)
                tmp[i] = support[i];
            for (int i = 0; i < rem; i++)
                tmp[ns + i] = support[idx[i]];
            qsort(tmp, (size_t)P, sizeof(int64_t), cmp_i64);
            double llk = llk_alleles(&c, tmp, P);
            double lpr = c.has_prior ? call_prior(&c, tmp, P) : 0.0;
            support_ljoint = orc_add_log_prob(support_ljoint, llk + lpr);
        } while (rem > 0 && cwr_next(idx, rem, ns));
        *out_support_prob = exp(support_ljoint - total_ljoint);
    }
    /* exact.py:108-153 _posterior_allele_frequencies */
    for (int a = 0; a < H; a++) {
        out_freqs[a] = 0.0;
        out_occur[a] = 0.0;
    }
    for (int i = 0; i < P; i++)
        g[i] = 0;
    for (int64_t i = 0; i < n_genotypes; i++) {
        double llk = llk_alleles(&c, g, P);
        double lpr = c.has_prior ? call_prior(&c, g, P) : 0.0;
        double prob = exp((llk + lpr) - total_ljoint);
        for (int k = 0; k < P; k++) {
            int64_t a = g[k];
            out_freqs[a] += prob;
            if (k == 0)
                out_occur[a] += prob;
            else if (a != g[k - 1])
                out_occur[a] += prob;
        }
        orc_increment_genotype(g, P);
    }
    for (int a = 0; a < H; a++)
        out_freqs[a] = out_freqs[a] / (double)P;
}
