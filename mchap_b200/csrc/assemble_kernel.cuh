// assemble_kernel.cuh — K2/K3: the whole of DenovoMCMC.fit for one (locus, sample) item per warp.
//
// Reference path restated B200-first (paths relative to the reference repository):
//   assemble/mcmc.py:103-265 (fit/_mcmc), 269-426 (_denovo_assembler), 455-541 (initial state,
//   homozygous fixing), snpcalling.py:14-70, mutation.py:15-246, structural.py:23-673,
//   tempering.py:11-151, prior.py:15-112, likelihood.py:18-148, jitutils.py (random_choice,
//   dosage, structural_change).
//
// Design:
//   * one warp = one item, chains and temperatures processed in the reference's serial order so
//     that the MT19937 word stream is consumed exactly as numba consumes it (bit-exact replay);
//   * the item's reads tensor is staged once into shared memory, transposed to
//     [position][allele][read] so that lane r reads column r conflict-free, gaps (NaN) -> 1.0
//     (multiplying by 1.0 is exactly "skip", likelihood.py:55-58);
//   * haplotypes are packed bit keys (B bits per allele) held uniformly: equality tests,
//     dosage and segment labels are integer compares;
//   * per temperature the per-(haplotype, read) products q[h][r] = prod_j reads[r,j,g[h,j]] / P
//     are cached in shared memory; a proposal recomputes only the changed haplotype(s).  The
//     evaluated value is a deterministic function of the ordered genotype (same instruction
//     sequence whether cached or recomputed), as in the reference;
//   * log-sum over reads = per-lane log() + xor-butterfly (uniform result in all lanes).
#pragma once
#include "common.cuh"

namespace mchb {

struct AsmArgs {
    const mchb_assemble_item *items;
    const int32_t *order;       // item ids handled by this launch
    int32_t n_order;
    const double *reads;
    const int64_t *counts;      // may be null
    const int8_t *n_alleles;
    const int8_t *initial;      // may be null
    int8_t *out_genotypes;
    double *out_llks;
    mchb_item_result *results;
    const uint32_t *words;      // [n_streams][stream_len]
    const int32_t *item_stream; // stream index per item
    int64_t stream_len;
    int32_t steps, chains;
    double fix_homozygous, p_recomb, p_partial, p_dosage;
    const double *break_table;
    const int32_t *break_len;
    int32_t break_rows, break_stride;
    const double *temperatures;
    int32_t *work_counter;
    // shared memory geometry (per warp)
    int32_t nmax, amax, pmax, tmax, maxopt;
    int32_t smem_per_warp;      // bytes
    // byte offsets of the per-warp arrays (host computed, asm_layout()); Rt is at offset 0
    int32_t o_cnt, o_q, o_dist, o_oll, o_opr, o_lgdisp, o_homlp, o_llk_t, o_key, o_sc, o_perm, o_het, o_fixa,
        o_nall, o_opt0, o_opt1, o_ivb, o_ivp;
};

// uniform per-item scalars parked in shared memory (sc[]) to keep them out of registers
enum { SC_LUH = 0, SC_LG_SUMDISP, SC_LG_P_SUMDISP, SC_LG_DISP, SC_INBREEDING, SC_COUNT };

// nibble helpers for packed small-integer vectors (labels, SNP genotypes), P <= 16
__device__ __forceinline__ int nib(uint64_t v, int i) { return (int)((v >> (4 * i)) & 15u); }
__device__ __forceinline__ uint64_t nib_set(uint64_t v, int i, int x) {
    return (v & ~(15ull << (4 * i))) | ((uint64_t)x << (4 * i));
}

// increment a nibble-packed sorted genotype to its VCF-order successor (jitutils.py:113-146)
__device__ __forceinline__ uint64_t increment_packed(uint64_t g, int P) {
    if (P == 1) return g + 1;
    int prev = nib(g, 0);
    for (int i = 1; i < P; i++) {
        int al = nib(g, i);
        if (al == prev) continue;
        g = nib_set(g, i - 1, nib(g, i - 1) + 1);  // al > prev for sorted input
        for (int m = 0; m < i - 1; m++) g = nib_set(g, m, 0);
        return g;
    }
    g = nib_set(g, P - 1, nib(g, P - 1) + 1);
    for (int m = 0; m < P - 1; m++) g = nib_set(g, m, 0);
    return g;
}

// calling/prior.py:116-179 with frequencies=None on a nibble-packed genotype (used by the
// per-SNP posterior, snpcalling.py:50-57)
__device__ __noinline__ double snp_log_genotype_prior(uint64_t g, int P, int n_alleles, double inbreeding) {
    double acc = 0.0;
    double alpha = 0.0, lg_alpha = 0.0;
    const bool null_prior = inbreeding == 0.0;
    if (!null_prior) {
        alpha = (1.0 / (double)n_alleles) * ((1.0 - inbreeding) / inbreeding);
        lg_alpha = lgamma(alpha);
    }
    for (int i = 0; i < P; i++) {
        int cntv = 0;
        bool first = true;
        for (int k = 0; k < P; k++) {
            bool eq = nib(g, k) == nib(g, i);
            cntv += eq;
            first = first && !(eq && k < i);
        }
        int dose = first ? cntv : 0;
        if (null_prior) acc += LGAMMA_INT[dose + 1];
        else if (dose > 0) acc += lgamma((double)dose + alpha) - (LGAMMA_INT[dose + 1] + lg_alpha);
    }
    if (null_prior) return (LGAMMA_INT[P + 1] - acc) - (double)P * log((double)n_alleles);
    double sum_alphas = alpha * (double)n_alleles;
    return ((LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma((double)P + sum_alphas)) + acc;
}

template <int CH>
struct AsmCtx {
    static constexpr int UPAD = CH * 32;
    const AsmArgs &a;
    unsigned char *sm;  // this warp's shared memory region
    int lane;
    int N, A, P, B;
    uint32_t amask;
    bool pow2, has_inb;
    double invP;
    uint32_t slots;  // nibble t -> state slot (parallel tempering swaps exchange slots)
    WordStream ws;
    int err;
    long long evals;

    __device__ __forceinline__ AsmCtx(const AsmArgs &args, unsigned char *s, int l) : a(args), sm(s), lane(l) {}

    __device__ __forceinline__ double *Rt() const { return reinterpret_cast<double *>(sm); }
    __device__ __forceinline__ double *cnt() const { return reinterpret_cast<double *>(sm + a.o_cnt); }
    __device__ __forceinline__ double *q() const { return reinterpret_cast<double *>(sm + a.o_q); }
    __device__ __forceinline__ double *dist() const { return reinterpret_cast<double *>(sm + a.o_dist); }
    __device__ __forceinline__ double *oll() const { return reinterpret_cast<double *>(sm + a.o_oll); }
    __device__ __forceinline__ double *opr() const { return reinterpret_cast<double *>(sm + a.o_opr); }
    __device__ __forceinline__ double *lgdisp() const { return reinterpret_cast<double *>(sm + a.o_lgdisp); }
    __device__ __forceinline__ double *homlp() const { return reinterpret_cast<double *>(sm + a.o_homlp); }
    __device__ __forceinline__ double *llk_t() const { return reinterpret_cast<double *>(sm + a.o_llk_t); }
    __device__ __forceinline__ double *sc() const { return reinterpret_cast<double *>(sm + a.o_sc); }
    __device__ __forceinline__ uint64_t *key() const { return reinterpret_cast<uint64_t *>(sm + a.o_key); }
    __device__ __forceinline__ uint16_t *perm() const { return reinterpret_cast<uint16_t *>(sm + a.o_perm); }
    __device__ __forceinline__ uint8_t *het() const { return sm + a.o_het; }
    __device__ __forceinline__ uint8_t *fixa() const { return sm + a.o_fixa; }
    __device__ __forceinline__ uint8_t *nall() const { return sm + a.o_nall; }
    __device__ __forceinline__ uint8_t *opt0() const { return sm + a.o_opt0; }
    __device__ __forceinline__ uint8_t *opt1() const { return sm + a.o_opt1; }
    __device__ __forceinline__ uint8_t *ivb() const { return sm + a.o_ivb; }
    __device__ __forceinline__ uint8_t *ivp() const { return sm + a.o_ivp; }

    __device__ __forceinline__ int slot(int t) const { return (int)((slots >> (4 * t)) & 15u); }
    __device__ __forceinline__ uint64_t *keys(int s) const { return key() + s * P; }

    // jitutils.random_choice (77-92): p (uniform, shared memory) is overwritten by its cumsum
    __device__ __forceinline__ int random_choice_inplace(double *p, int n) {
        double acc = 0.0;
        for (int i = 0; i < n; i++) {
            acc += p[i];
            p[i] = acc;
        }
        double u = ws.next_double(lane);
        return searchsorted_right(p, n, u);
    }

    // products of one haplotype key over all positions, divided by the ploidy
    // (likelihood.py:48-60): out[ch] belongs to read ch*32+lane
    __device__ __forceinline__ void hap_products(uint64_t k, double (&out)[CH]) const {
#pragma unroll
        for (int ch = 0; ch < CH; ch++) out[ch] = 1.0;
        const double *base = Rt() + lane;
        const int stride = A * UPAD;
        for (int j = 0; j < N; j++) {
            int al = (int)((uint32_t)k & amask);
            k >>= B;
            const double *p = base + al * UPAD;
#pragma unroll
            for (int ch = 0; ch < CH; ch++) out[ch] *= p[ch * 32];
            base += stride;
        }
        if (pow2) {
#pragma unroll
            for (int ch = 0; ch < CH; ch++) out[ch] = out[ch] * invP;
        } else {
            const double dP = (double)P;
#pragma unroll
            for (int ch = 0; ch < CH; ch++) out[ch] = out[ch] / dP;
        }
    }

    // log-likelihood of state slot s with haplotype hA (and hB) replaced by the given product
    // vectors (likelihood.py:45-68 order: sum over h in order, log, * count, sum over reads)
    __device__ __forceinline__ double eval_llk(int s, int hA, const double (&qa)[CH], int hB,
                                               const double (&qb)[CH]) {
        double acc = 0.0;
        const double *qq = q() + (size_t)(s * P) * UPAD + lane;
        const double *cn = cnt() + lane;
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
            double rp = 0.0;
            for (int h = 0; h < P; h++) {
                double v = qq[h * UPAD + ch * 32];
                v = (h == hA) ? qa[ch] : v;
                v = (h == hB) ? qb[ch] : v;
                rp += v;
            }
            acc += log(rp) * cn[ch * 32];
        }
        evals++;
        return warp_sum(acc);
    }

    __device__ __forceinline__ void commit(int s, int h, uint64_t k, const double (&qa)[CH]) {
        keys(s)[h] = k;
        double *row = q() + (size_t)(s * P + h) * UPAD + lane;
#pragma unroll
        for (int ch = 0; ch < CH; ch++) row[ch * 32] = qa[ch];
    }

    // copies of key kh among the haplotypes ks with haplotype hs replaced by kh
    // (jitutils.count_haplotype_copies 349-374)
    __device__ __forceinline__ int count_copies(const uint64_t *ks, int hs, uint64_t kh) const {
        int c = 0;
        for (int i = 0; i < P; i++) {
            uint64_t k = (i == hs) ? kh : ks[i];
            c += (k == kh);
        }
        return c;
    }

    // assemble/prior.py:15-112 from P comparable row identifiers sel(i); the first-occurrence
    // dosage (jitutils.get_haplotype_dosage 377-422) is evaluated on the fly, terms are
    // accumulated in row order like the reference
    template <typename F>
    __device__ __forceinline__ double prior_generic(F sel) const {
        const double *scv = sc();
        const bool null_prior = scv[SC_INBREEDING] == 0.0;
        const double *lgd = lgdisp();
        const double lg_disp = scv[SC_LG_DISP];
        double acc = 0.0;
        for (int i = 0; i < P; i++) {
            const uint64_t ki = sel(i);
            int cntv = 0;
            bool first = true;
            for (int k = 0; k < P; k++) {
                bool eq = sel(k) == ki;
                cntv += eq;
                first = first && !(eq && k < i);
            }
            const int dose = first ? cntv : 0;
            if (null_prior) acc += LGAMMA_INT[dose + 1];
            else if (dose > 0) acc += lgd[dose] - (LGAMMA_INT[dose + 1] + lg_disp);
        }
        if (null_prior) return (LGAMMA_INT[P + 1] - acc) - (double)P * scv[SC_LUH];
        return ((LGAMMA_INT[P + 1] + scv[SC_LG_SUMDISP]) - scv[SC_LG_P_SUMDISP]) + acc;
    }

    __device__ double prior_of_keys(const uint64_t *ks, int hA, uint64_t kA, int hB, uint64_t kB) const {
        return prior_generic([=](int i) {
            uint64_t v = ks[i];
            v = (i == hA) ? kA : v;
            v = (i == hB) ? kB : v;
            return v;
        });
    }

    // prior of a label matrix (structural.py:546: dosage of the (inside, outside) label rows)
    __device__ double prior_of_labels(uint64_t lin, uint64_t lout) const {
        return prior_generic([=](int i) { return (uint64_t)((nib(lin, i) << 4) | nib(lout, i)); });
    }

    // ------------------------------------------------------------------ mutation.py:15-161
    __device__ __forceinline__ void base_step(int s, int h, int j, int n_all, double temp, double &llk) {
        uint64_t *ks = keys(s);
        const uint64_t kh = ks[h];
        const int shift = B * j;
        const uint64_t clr = ~((uint64_t)amask << shift);
        const int cur = (int)((uint32_t)(kh >> shift) & amask);
        const double lhap = LOG_INT[count_copies(ks, -1, kh)];
        double lprior = 0.0;
        if (has_inb) lprior = prior_of_keys(ks, -1, 0, -1, 0);
        double *ol = oll(), *op = opr();
        int n_options = 0;
        double qn[CH];
        for (int i = 0; i < n_all; i++) {
            if (i == cur) {
                ol[i] = llk;
                op[i] = -INFINITY;
            } else {
                n_options++;
                const uint64_t kn = (kh & clr) | ((uint64_t)i << shift);
                hap_products(kn, qn);
                const double llk_i = eval_llk(s, h, qn, -1, qn);
                ol[i] = llk_i;
                const double llk_ratio = llk_i - llk;
                double lprior_ratio = 0.0;
                if (has_inb) lprior_ratio = prior_of_keys(ks, h, kn, -1, 0) - lprior;
                const double lprop = LOG_INT[count_copies(ks, h, kn)] - lhap;
                const double mh = (llk_ratio + lprior_ratio) * temp + lprop;
                op[i] = np_minimum0(mh);
            }
        }
        const double ln_opts = LOG_INT[n_options];
        for (int i = 0; i < n_all; i++) op[i] = exp(op[i] - ln_opts);
        double sum = 0.0;
        for (int i = 0; i < n_all; i++) sum += op[i];
        op[cur] = 1 - sum;
        const int choice = random_choice_inplace(op, n_all);
        if (choice >= n_all) {
            err = MCHB_ITEM_CHOICE_RANGE;
            return;
        }
        if (choice != cur) {
            const uint64_t kn = (kh & clr) | ((uint64_t)choice << shift);
            hap_products(kn, qn);
            commit(s, h, kn, qn);
        }
        llk = ol[choice];
    }

    // ------------------------------------------------------------------ mutation.py:165-246
    __device__ __forceinline__ void mutation_compound_step(int s, double temp, double &llk) {
        const int n = P * N;
        uint16_t *pm = perm();
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            int h = i / N, j = i - h * N;
            pm[i] = (uint16_t)((h << 8) | j);
        }
        __syncwarp();
        for (int i = n - 1; i > 0; i--) {  // np.random.shuffle of the rows
            int k = ws.randint(i + 1, lane);
            uint16_t x = pm[i], y = pm[k];
            __syncwarp();
            pm[i] = y;
            pm[k] = x;
        }
        __syncwarp();
        const uint8_t *na = nall();
        for (int i = 0; i < n && !err; i++) {
            int hj = pm[i];
            int j = hj & 255;
            base_step(s, hj >> 8, j, na[j], temp, llk);
        }
    }

    // ------------------------------------------------------------------ structural.py:311-430
    __device__ __forceinline__ void segment_labels(const uint64_t *ks, uint64_t mask_in, uint64_t &lin,
                                                   uint64_t &lout) const {
        lin = 0;
        lout = 0;
        for (int h = 1; h < P; h++) {
            const uint64_t kk = ks[h];
            int fi = h, fo = h;
            for (int k = h - 1; k >= 0; k--) {
                const uint64_t d = ks[k] ^ kk;
                if ((d & mask_in) == 0) fi = k;
                if ((d & ~mask_in) == 0) fo = k;
            }
            lin |= (uint64_t)fi << (4 * h);
            lout |= (uint64_t)fo << (4 * h);
        }
    }

    // bit h set when row h of the (lin, lout) label matrix is not a duplicate of an earlier row
    __device__ __forceinline__ uint32_t first_full(uint64_t lin, uint64_t lout) const {
        uint32_t m = 0;
        for (int h = 0; h < P; h++) {
            bool dup = false;
            for (int k = 0; k < h; k++) dup = dup || (nib(lin, k) == nib(lin, h) && nib(lout, k) == nib(lout, h));
            m |= (dup ? 0u : 1u) << h;
        }
        return m;
    }

    // bit h set when h is the first haplotype carrying its inside segment / the only one carrying it
    __device__ __forceinline__ void segment_stats(uint64_t lin, uint32_t &seg_first, uint32_t &seg_single) const {
        seg_first = 0;
        seg_single = 0;
        for (int h = 0; h < P; h++) {
            bool dup = false;
            int c = 0;
            for (int k = 0; k < P; k++) {
                bool eq = nib(lin, k) == nib(lin, h);
                c += eq;
                dup = dup || (eq && k < h);
            }
            seg_first |= (dup ? 0u : 1u) << h;
            seg_single |= (c == 1 ? 1u : 0u) << h;
        }
    }

    // structural.py:75-121 (count) / 124-178 (enumerate, WRITE)
    template <bool WRITE>
    __device__ __forceinline__ int recomb_options(uint64_t lin, uint64_t lout) const {
        const uint32_t ff = first_full(lin, lout);
        uint8_t *o0 = opt0(), *o1 = opt1();
        int n = 0;
        for (int h0 = 0; h0 < P; h0++) {
            if (!((ff >> h0) & 1)) continue;
            for (int h1 = h0 + 1; h1 < P; h1++) {
                if (!((ff >> h1) & 1)) continue;
                if (nib(lin, h0) == nib(lin, h1) || nib(lout, h0) == nib(lout, h1)) continue;
                if (WRITE) {
                    o0[n] = (uint8_t)h0;
                    o1[n] = (uint8_t)h1;
                }
                n++;
            }
        }
        return n;
    }

    // structural.py:182-236 (count) / 239-307 (enumerate, WRITE)
    template <bool WRITE>
    __device__ __forceinline__ int dosage_options(uint64_t lin, uint64_t lout) const {
        const uint32_t ff = first_full(lin, lout);
        uint32_t sf, ss;
        segment_stats(lin, sf, ss);
        uint8_t *o0 = opt0(), *o1 = opt1();
        int n = 0;
        for (int h0 = 0; h0 < P; h0++) {
            if (!((ff >> h0) & 1)) continue;   // full duplicate of an earlier haplotype
            if ((ss >> h0) & 1) continue;      // would delete the only copy of its segment
            for (int h1 = 0; h1 < P; h1++) {
                if (!((sf >> h1) & 1)) continue;  // donor segment already visited
                if (nib(lin, h0) == nib(lin, h1)) continue;
                if (WRITE) {
                    o0[n] = (uint8_t)h0;
                    o1[n] = (uint8_t)h1;
                }
                n++;
            }
        }
        return n;
    }

    // ------------------------------------------------------------------ structural.py:434-587
    __device__ void interval_step(int s, int start, int stop, int step_type, double temp, double &llk) {
        uint64_t *ks = keys(s);
        const int width = B * (stop - start);
        uint64_t mask_in = 0;
        if (width >= 64) mask_in = ~0ull;
        else if (width > 0) mask_in = ((1ull << width) - 1ull) << (B * start);
        uint64_t lin, lout;
        segment_labels(ks, mask_in, lin, lout);
        __syncwarp();
        const int n_options = (step_type == 0) ? recomb_options<true>(lin, lout) : dosage_options<true>(lin, lout);
        __syncwarp();
        if (n_options == 0) return;  // no draw (structural.py:504-506)
        const double log_proposal = LOG_INV_INT[n_options];
        double lprior = 0.0;
        if (has_inb) lprior = prior_of_keys(ks, -1, 0, -1, 0);
        double *ol = oll(), *op = opr();
        const uint8_t *o0 = opt0(), *o1 = opt1();
        double qa[CH], qb[CH];
        for (int i = 0; i < n_options; i++) {
            const int h0 = o0[i], h1 = o1[i];
            const uint64_t k0 = ks[h0], k1 = ks[h1];
            const uint64_t k0n = (k1 & mask_in) | (k0 & ~mask_in);
            uint64_t lin_o = nib_set(lin, h0, nib(lin, h1));
            double llk_i;
            hap_products(k0n, qa);
            if (step_type == 0) {
                const uint64_t k1n = (k0 & mask_in) | (k1 & ~mask_in);
                lin_o = nib_set(lin_o, h1, nib(lin, h0));
                hap_products(k1n, qb);
                llk_i = eval_llk(s, h0, qa, h1, qb);
            } else {
                llk_i = eval_llk(s, h0, qa, -1, qa);
            }
            ol[i] = llk_i;
            const double llk_ratio = llk_i - llk;
            double lprior_ratio = 0.0;
            if (has_inb) lprior_ratio = prior_of_labels(lin_o, lout) - lprior;
            const int n_return =
                (step_type == 0) ? recomb_options<false>(lin_o, lout) : dosage_options<false>(lin_o, lout);
            const double lprop = LOG_INV_INT[n_return] - log_proposal;
            const double mh = (llk_ratio + lprior_ratio) * temp + lprop;
            op[i] = np_minimum0(mh);
        }
        ol[n_options] = -INFINITY;
        op[n_options] = -INFINITY;
        const double ln_opts = LOG_INT[n_options];
        for (int i = 0; i <= n_options; i++) op[i] = exp(op[i] - ln_opts);
        double sum = 0.0;
        for (int i = 0; i <= n_options; i++) sum += op[i];
        op[n_options] = 1 - sum;
        const int choice = random_choice_inplace(op, n_options + 1);
        if (choice < n_options) {
            const int h0 = o0[choice], h1 = o1[choice];
            const uint64_t k0 = ks[h0], k1 = ks[h1];
            const uint64_t k0n = (k1 & mask_in) | (k0 & ~mask_in);
            hap_products(k0n, qa);
            if (step_type == 0) {
                const uint64_t k1n = (k0 & mask_in) | (k1 & ~mask_in);
                hap_products(k1n, qb);
                commit(s, h1, k1n, qb);
            }
            commit(s, h0, k0n, qa);
            llk = ol[choice];
        }
    }

    // structural.py:23-71 random_breaks + 591-673 compound_step
    __device__ void structural_step(int s, int n_breaks, int step_type, double temp, double &llk) {
        if (n_breaks >= N) {
            err = MCHB_ITEM_BREAKS;
            return;
        }
        uint64_t avail = 0;  // candidate cut points 1..N-1
        if (N >= 2) avail = ((N - 1 >= 64) ? ~0ull : ((1ull << (N - 1)) - 1ull)) << 1;
        uint64_t cuts = 0;
        for (int b = 0; b < n_breaks; b++) {
            int m = __popcll(avail);
            if (m == 0) break;
            int k = ws.randint(m, lane);  // np.random.choice(options)
            uint64_t t = avail;
            for (int i = 0; i < k; i++) t &= t - 1;
            int point = __ffsll((long long)t) - 1;
            avail &= ~(1ull << point);
            cuts |= 1ull << point;
        }
        uint8_t *vb = ivb(), *vp = ivp();
        __syncwarp();
        int nb = 0;
        vb[nb++] = 0;
        for (uint64_t t = cuts; t; t &= t - 1) vb[nb++] = (uint8_t)(__ffsll((long long)t) - 1);
        vb[nb] = (uint8_t)N;
        const int n_int = n_breaks + 1;
        for (int i = 0; i < n_int; i++) vp[i] = (uint8_t)i;
        __syncwarp();
        for (int i = n_int - 1; i > 0; i--) {  // np.random.permutation(np.arange(n))
            int k = ws.randint(i + 1, lane);
            uint8_t x = vp[i], y = vp[k];
            __syncwarp();
            vp[i] = y;
            vp[k] = x;
        }
        __syncwarp();
        for (int i = 0; i < n_int && !err; i++) {
            int p = vp[i];
            interval_step(s, vb[p], vb[p + 1], step_type, temp, llk);
        }
    }

    // tempering.py:62-151 between temperature t (cooler, i) and t-1 (warmer, j)
    __device__ void chain_swap_step(int t, double temp_i, double temp_j, double &llk_i) {
        const int si = slot(t), sj = slot(t - 1);
        double *lt = llk_t();
        double llk_j = lt[t - 1];
        double prior_i = 0.0, prior_j = 0.0;
        if (has_inb) {
            prior_i = prior_of_keys(keys(si), -1, 0, -1, 0);
            prior_j = prior_of_keys(keys(sj), -1, 0, -1, 0);
        }
        double post_i = llk_i + prior_i, post_j = llk_j + prior_j;
        double frac_1 = (post_j - post_i) * temp_i;
        double frac_2 = (post_i - post_j) * temp_j;
        double acc = exp(frac_1 + frac_2);
        if (acc > 1.0) acc = 1.0;
        double val = ws.next_double(lane);
        if (acc >= val) {
            slots = (slots & ~((15u << (4 * t)) | (15u << (4 * (t - 1))))) | ((uint32_t)sj << (4 * t)) |
                    ((uint32_t)si << (4 * (t - 1)));
            __syncwarp();
            lt[t - 1] = llk_i;
            llk_i = llk_j;
        }
    }
};

template <int CH>
__global__ void __launch_bounds__(128) assemble_kernel(const __grid_constant__ AsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int UPAD = CH * 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    AsmCtx<CH> c(a, smem_raw + (size_t)warp * a.smem_per_warp, lane);

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(MCHB_FULL, w, 0);
        if (w >= a.n_order) break;
        const int item_id = a.order[w];
        const mchb_assemble_item *itp = a.items + item_id;
        const int Nf = itp->n_pos, A = itp->max_allele, P = itp->ploidy, T = itp->n_temps;
        const int Uin = itp->n_reads;
        const int U = Uin > 0 ? Uin : 1;  // mcmc.py:132-137: one all-gap read stands in for none
        const double inbreeding = itp->inbreeding;
        const bool has_initial = a.initial != nullptr && itp->initial_off >= 0;
        c.A = A;
        c.P = P;
        c.err = 0;
        c.evals = 0;
        c.invP = 1.0 / (double)P;
        c.pow2 = (P & (P - 1)) == 0;
        c.has_inb = !isnan(inbreeding);
        c.ws.init(a.words + (size_t)a.item_stream[item_id] * a.stream_len, a.stream_len, lane);
        const int8_t *nal_full = a.n_alleles + itp->nalleles_off;
        int8_t *og = a.out_genotypes + itp->genotypes_off;
        double *ol = a.out_llks + itp->llks_off;
        const int step_sz = P * Nf;
        double *Rt = c.Rt();
        double *cnt = c.cnt();

        // ---- stage reads transposed: Rt[(j*A + al)*UPAD + r], raw values (NaN kept for now)
        __syncwarp();
        for (int i = lane; i < Nf * A * UPAD; i += 32) Rt[i] = 1.0;
        for (int i = lane; i < UPAD; i += 32) cnt[i] = 0.0;
        __syncwarp();
        if (Uin > 0) {
            const double *src = a.reads + itp->reads_off;
            const int row = Nf * A;
            const int tot = Uin * row;
            for (int i = lane; i < tot; i += 32) {
                int r = i / row;
                int ja = i - r * row;
                Rt[ja * UPAD + r] = __ldg(src + i);
            }
            for (int r = lane; r < Uin; r += 32)
                cnt[r] = a.counts ? (double)__ldg(a.counts + itp->counts_off + r) : 1.0;
        } else {
            for (int i = lane; i < Nf * A; i += 32) Rt[i * UPAD] = NAN;
            if (lane == 0) cnt[0] = 1.0;
        }
        if (lane == 0) c.sc()[SC_INBREEDING] = inbreeding;
        __syncwarp();

        // ---- homozygous fixing: mcmc.py:495-541 + snpcalling.py:14-70
        int n_het = 0;
        {
            double *homlp = c.homlp();
            uint8_t *het = c.het(), *fixa = c.fixa(), *nall = c.nall();
            for (int j = 0; j < Nf; j++) {
                const int nA = nal_full[j];
                const long long u_gens = comb_with_replacement(nA, P);
                uint64_t g = 0;
                double denom = 0.0;
                for (long long i = 0; i < u_gens; i++) {
                    double lprior = 0.0;
                    if (c.has_inb) lprior = snp_log_genotype_prior(g, P, nA, inbreeding);
                    double acc = 0.0;
#pragma unroll
                    for (int ch = 0; ch < CH; ch++) {
                        double rp = 0.0;
                        for (int h = 0; h < P; h++) {
                            double v = Rt[(j * A + nib(g, h)) * UPAD + ch * 32 + lane];
                            double prod = isnan(v) ? 1.0 : v;
                            rp += c.pow2 ? prod * c.invP : prod / (double)P;
                        }
                        acc += log(rp) * cnt[ch * 32 + lane];
                    }
                    double lp = lprior + warp_sum(acc);
                    denom = (i == 0) ? lp : add_log_prob(denom, lp);
                    if (nib(g, 0) == nib(g, P - 1)) homlp[nib(g, 0)] = lp;
                    g = increment_packed(g, P);
                }
                __syncwarp();
                int fixed = 0, fa = 0;
                for (int al = 0; al < nA; al++) {
                    double prob = exp(homlp[al] - denom);
                    if (prob >= a.fix_homozygous) {
                        fixed = 1;
                        fa = al;
                    }
                }
                __syncwarp();
                fixa[j] = (uint8_t)fa;
                if (!fixed) {
                    het[n_het] = (uint8_t)j;
                    nall[n_het] = (uint8_t)nA;
                    n_het++;
                }
            }
        }
        __syncwarp();
        const int N = n_het;
        c.N = N;

        int status = 0;
        if (N == 0) {
            // mcmc.py:188-199: every position fixed -> tiled haplotype, NaN llks, no RNG use
            const uint8_t *fixa = c.fixa();
            for (int ch = 0; ch < a.chains; ch++) {
                for (int i = lane; i < a.steps * step_sz; i += 32)
                    og[(size_t)ch * a.steps * step_sz + i] = (int8_t)fixa[i % Nf];
                for (int i = lane; i < a.steps; i += 32) ol[(size_t)ch * a.steps + i] = NAN;
            }
        } else {
            // bits per allele and shape limits
            const uint8_t *nall = c.nall();
            int amax_het = 0;
            for (int k = 0; k < N; k++) amax_het = max(amax_het, (int)nall[k]);
            if (has_initial) amax_het = max(amax_het, A);  // user states may use any allele < max_allele
            int B = 1;
            while ((1 << B) < amax_het) B++;
            c.B = B;
            c.amask = (1u << B) - 1u;
            if (N * B > 64 || P > MCHB_MAX_PLOIDY || T > MCHB_MAX_TEMPS || T < 1) status = MCHB_ITEM_UNSUPPORTED;
            if (!status && has_initial && itp->initial_nhet != N) status = MCHB_ITEM_INITIAL_SHAPE;
        }
        if (N > 0 && !status) {
            const uint8_t *het = c.het(), *nall = c.nall();
            double *dist = c.dist(), *opr = c.opr();
            // ---- compact the variable positions (het[k] >= k, ascending: in-place is safe)
            for (int k = 0; k < N; k++) {
                int j = het[k];
                if (j != k)
                    for (int i = lane; i < A * UPAD; i += 32) Rt[k * A * UPAD + i] = Rt[j * A * UPAD + i];
                __syncwarp();
            }
            // ---- initial-state distribution: mcmc.py:455-491 then jitutils.py:483-487
            if (!has_initial) {
                for (int k = 0; k < N; k++) {
                    int n_nonzero = 0;
                    uint32_t gapmask = 0;
                    for (int al = 0; al < A; al++) {
                        const double *col = Rt + (k * A + al) * UPAD;
                        bool all_nan = true;
                        for (int r = 0; r < U; r++) all_nan = all_nan && isnan(col[r]);
                        double tot = 0.0;
                        int cn = 0;
                        bool all_zero = true;
                        for (int r = 0; r < U; r++) {
                            double v = all_nan ? 1.0 : col[r];
                            if (!isnan(v)) {
                                tot += v;
                                cn++;
                            }
                            if (!(v == 0.0)) all_zero = false;
                        }
                        __syncwarp();
                        dist[k * A + al] = tot / (double)cn;
                        if (all_nan) gapmask |= 1u << al;
                        if (!all_zero) n_nonzero++;
                    }
                    __syncwarp();
                    double s1 = 0.0;
                    for (int al = 0; al < A; al++) {
                        double v = dist[k * A + al];
                        if ((gapmask >> al) & 1) v = 1.0 / (double)n_nonzero;
                        opr[al] = v;
                        s1 += v;
                    }
                    double s2 = 0.0;
                    for (int al = 0; al < A; al++) {
                        double v = opr[al] / s1;
                        dist[k * A + al] = v;
                        s2 += v;
                    }
                    for (int al = 0; al < A; al++) dist[k * A + al] = dist[k * A + al] / s2;
                    __syncwarp();
                }
            }
            // ---- gaps become 1.0 from here on (likelihood.py:55-58)
            for (int i = lane; i < N * A * UPAD; i += 32) {
                double v = Rt[i];
                if (isnan(v)) Rt[i] = 1.0;
            }
            __syncwarp();
            // ---- per-item prior constants
            {
                float s = 0.0f;  // mcmc.py:294: float32 arithmetic in numba (int8 array)
                for (int k = 0; k < N; k++) s += LOGF_INT[nall[k]];
                const double luh = (double)s;
                double *scv = c.sc();
                double *lgd = c.lgdisp();
                __syncwarp();
                scv[SC_LUH] = luh;
                if (c.has_inb && inbreeding != 0.0) {
                    double log_disp = log((1.0 - inbreeding) / inbreeding) - luh;
                    double disp = exp(log_disp);
                    double sum_disp = exp(log_disp + luh);
                    scv[SC_LG_SUMDISP] = lgamma(sum_disp);
                    scv[SC_LG_P_SUMDISP] = lgamma((double)P + sum_disp);
                    scv[SC_LG_DISP] = lgamma(disp);
                    for (int d = 1; d <= P; d++) lgd[d] = lgamma((double)d + disp);
                }
                __syncwarp();
            }
            const int brow_i = min(N, a.break_rows - 1);
            const double *brow = a.break_table + (size_t)brow_i * a.break_stride;
            const int blen = a.break_len[brow_i];
            const double *temps = a.temperatures + itp->temps_off;
            double *lt = c.llk_t();

            for (int chain = 0; chain < a.chains && !c.err; chain++) {
                // ---- initial genotype written into state slot 0
                uint64_t *k0 = c.keys(0);
                __syncwarp();
                if (has_initial) {
                    const int8_t *src = a.initial + itp->initial_off + (size_t)chain * P * N;
                    for (int h = 0; h < P; h++) {
                        uint64_t k = 0;
                        for (int j = 0; j < N; j++) k |= (uint64_t)((uint32_t)(uint8_t)src[h * N + j] & c.amask) << (c.B * j);
                        k0[h] = k;
                    }
                } else {
                    for (int h = 0; h < P; h++) {
                        uint64_t k = 0;
                        for (int j = 0; j < N; j++) {
                            for (int al = 0; al < A; al++) opr[al] = dist[j * A + al];
                            int choice = c.random_choice_inplace(opr, A);
                            if (choice >= A) c.err = MCHB_ITEM_CHOICE_RANGE;
                            k |= (uint64_t)((uint32_t)choice & c.amask) << (c.B * j);
                        }
                        k0[h] = k;
                    }
                }
                if (c.err) break;
                __syncwarp();
                // ---- all temperatures start from the same state (mcmc.py:296-303)
                c.slots = 0x76543210u;
                {
                    double qn[CH];
                    for (int h = 0; h < P; h++) {
                        const uint64_t k = k0[h];
                        c.hap_products(k, qn);
                        for (int t = 0; t < T; t++) c.commit(t, h, k, qn);
                    }
                    __syncwarp();
                    double llk0 = c.eval_llk(0, -1, qn, -1, qn);
                    c.evals--;  // the initial evaluation is not a proposal
                    for (int t = 0; t < T; t++) lt[t] = llk0;
                    __syncwarp();
                }
                int8_t *ogc = og + (size_t)chain * a.steps * step_sz;
                double *olc = ol + (size_t)chain * a.steps;
                for (int step = 0; step < a.steps && !c.err; step++) {
                    double llk = 0.0;
                    for (int t = 0; t < T && !c.err; t++) {
                        llk = lt[t];
                        const int s = c.slot(t);
                        const double temp = temps[t];
                        if (isnan(llk)) {
                            c.err = MCHB_ITEM_NAN_LLK;
                            break;
                        }
                        c.mutation_compound_step(s, temp, llk);
                        if (c.err) break;
                        if (c.ws.next_double(lane) <= a.p_recomb) {
                            for (int i = 0; i < blen; i++) opr[i] = brow[i];
                            int n_breaks = c.random_choice_inplace(opr, blen);
                            c.structural_step(s, n_breaks, 0, temp, llk);
                            if (c.err) break;
                        }
                        if (c.ws.next_double(lane) <= a.p_partial) {
                            for (int i = 0; i < blen; i++) opr[i] = brow[i];
                            int n_breaks = c.random_choice_inplace(opr, blen);
                            c.structural_step(s, n_breaks, 1, temp, llk);
                            if (c.err) break;
                        }
                        if (c.ws.next_double(lane) <= a.p_dosage) {
                            c.interval_step(s, 0, N, 1, temp, llk);  // permutation of one interval draws nothing
                            if (c.err) break;
                        }
                        if (t > 0) c.chain_swap_step(t, temp, temps[t - 1], llk);
                        __syncwarp();
                        lt[t] = llk;
                    }
                    if (c.err) break;
                    // ---- record the cold chain (state of the last temperature, mcmc.py:418-425)
                    {
                        const uint64_t *ks = c.keys(c.slot(T - 1));
                        const uint8_t *fixa = c.fixa();
                        int8_t *dst = ogc + (size_t)step * step_sz;
                        __syncwarp();
                        for (int i = lane; i < step_sz; i += 32) {
                            int h = i / Nf, j = i - h * Nf;
                            dst[i] = (int8_t)fixa[j];
                        }
                        __syncwarp();
                        for (int i = lane; i < P * N; i += 32) {
                            int h = i / N, k = i - h * N;
                            dst[h * Nf + het[k]] = (int8_t)((uint32_t)(ks[h] >> (c.B * k)) & c.amask);
                        }
                        if (lane == 0) olc[step] = llk;
                        __syncwarp();
                    }
                }
            }
            status = c.err;
            if (!status && c.ws.exhausted) status = MCHB_ITEM_RNG_EXHAUSTED;
        }
        if (lane == 0) {
            mchb_item_result r;
            r.status = status;
            r.n_het = N;
            r.rng_words = c.ws.cur;
            r.llk_evals = c.evals;
            a.results[item_id] = r;
        }
        __syncwarp();
    }
}

}  // namespace mchb
