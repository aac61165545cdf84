// call_mcmc_kernel.cuh — K5: CallingMCMC.fit (mchap call) for one (locus, sample) item per warp.
//
// Reference restated (paths relative to the reference repository):
//   calling/classes.py:49-124 CallingMCMC.fit, calling/mcmc.py:15-140 mh_options, 143-229
//   gibbs_options, 232-327 compound_step, 330-390 mcmc_sampler, 393-453 greedy_caller,
//   calling/prior.py:10-179, calling/likelihood.py:8-78, calling/utils.py:7-57.
//
// Design: the read x haplotype table t[r][h] = prod_j reads[r,j,hap[h,j]] (gaps skipped) is built
// once per item in shared memory.  A Gibbs / MH sub-step evaluates all H candidate alleles of one
// genotype slot at once: lane a owns candidate a (rounds of 32 when H > 32) and walks the reads
// in order, so each candidate's log-likelihood has the reference's operation order (sum over
// slots in slot order, log, * count, sum over reads in read order).  The categorical draw uses a
// warp log-sum-exp, an inclusive scan and a ballot.  Chains run one after another on the item's
// MT19937 word stream exactly as numba consumes it.
#pragma once
#include "common.cuh"

namespace mchb {

struct CallMcmcArgs {
    const mchb_call_item *items;
    const int32_t *order;
    int32_t n_order;
    const double *reads;
    const int64_t *counts;
    const int8_t *haplotypes;
    const double *freqs;
    const int32_t *initial;     // [n_items, pstride] or null; a row starting with < 0 = greedy
    int32_t pstride;
    int32_t *out_alleles;       // per item at gl_off: int32[chains, steps, P]
    double *out_llks;           // per item at hap_out_off: f64[chains, steps]
    mchb_item_result *results;
    const uint32_t *words;
    const int32_t *item_stream;
    int64_t stream_len;
    int32_t steps, chains, step_type;
    int32_t *work_counter;
    int32_t umax, hmax, pmax;
    int32_t smem_per_warp;
    int32_t memo_off;           // byte offset of the per-warp memo region, -1: no memo
};

// host-initialised: LOG_RATIO[i * 17 + j] = log((double)i / (double)j), 1 <= i, j <= 16
__constant__ double LOG_RATIO[17 * 17];

__device__ __forceinline__ double warp_lse(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v = add_log_prob(v, __shfl_xor_sync(MCHB_FULL, v, m));
    return v;
}

// calling/prior.py:116-179 on a small uniform genotype array (greedy initialisation only)
__device__ __noinline__ double call_log_genotype_prior_slow(const int *g, int P, int H, double inbreeding,
                                                            const double *freqs) {
    return calling_log_genotype_prior(g, P, H, inbreeding, freqs);
}

__global__ void __launch_bounds__(128) call_mcmc_kernel(const __grid_constant__ CallMcmcArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *sm = smem_raw + (size_t)warp * a.smem_per_warp;
    double *tab = reinterpret_cast<double *>(sm);                  // [umax][hmax]
    double *cnt = tab + (size_t)a.umax * a.hmax;                    // [umax]
    double *xs = cnt + a.umax;                                      // [hmax] llk + lprior / probabilities
    double *ls = xs + a.hmax;                                       // [hmax] llks
    double *lgA = ls + a.hmax;                                      // [hmax][pmax+2] lgamma(alpha_a + c)
    int *gs = reinterpret_cast<int *>(lgA + (size_t)a.hmax * (a.pmax + 2));  // [pmax] genotype
    int *ord = gs + a.pmax;                                         // [pmax] slot order
    int *ini = ord + a.pmax;                                        // [pmax] initial genotype
    uint32_t *ring = reinterpret_cast<uint32_t *>(ini + a.pmax);    // [128] RNG word ring
    // Memo: the categorical distribution of slot k is a function of (k, genotype) only, and a
    // chain that sits in a mode asks for the same one step after step.  Per slot: the genotype it
    // was computed for, its cumulative sums (what random_choice searches) and the candidates' llks.
    // Two entries (ways) per slot, least recently used replaced: a chain that leaves its mode for a
    // step or two finds the mode's distributions again when it returns (measured: 233 M -> 298 M MCMC
    // steps/s at configs[4]; three and four ways lose more to shared memory than they gain).  The Gibbs
    // conditional of slot k does not depend on the allele slot k holds, so that allele is not part of
    // its key (the Metropolis-Hastings distribution is relative to the current allele: there it is).
    const bool memo = a.memo_off >= 0;
    double *mcs = reinterpret_cast<double *>(sm + (memo ? a.memo_off : 0));   // [pmax][2][hmax] cumulative sums
    double *mls = mcs + (size_t)2 * a.pmax * a.hmax;                          // [pmax][2][hmax] llks
    int *mkey = reinterpret_cast<int *>(mls + (size_t)2 * a.pmax * a.hmax);   // [pmax][2][pmax] genotype of the entry
    int *mvalid = mkey + 2 * a.pmax * a.pmax;                                 // [pmax][2]
    int *mlru = mvalid + 2 * a.pmax;                                          // [pmax] way to replace next

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(MCHB_FULL, w, 0);
        if (w >= a.n_order) break;
        const int item_id = a.order[w];
        const mchb_call_item it = a.items[item_id];
        const int U = it.n_reads, N = it.n_pos, A = it.max_allele, P = it.ploidy, H = it.n_haps;
        const double *R = a.reads + it.reads_off;
        const int8_t *haps = a.haplotypes + it.haps_off;
        const double *freqs = (a.freqs && it.freqs_off >= 0) ? a.freqs + it.freqs_off : nullptr;
        const bool has_prior = !isnan(it.inbreeding);
        const double inbreeding = it.inbreeding;
        const double dP = (double)P;
        WordStream ws;
        ws.init(a.words + (size_t)a.item_stream[item_id] * a.stream_len, a.stream_len, ring, lane);
        long long evals = 0;
        int err = 0;
        if (memo) {
            for (int k = lane; k < 2 * a.pmax; k += 32) mvalid[k] = 0;
            for (int k = lane; k < a.pmax; k += 32) mlru[k] = 0;
        }

        // ---- raw table t[r][h] (likelihood.py:48-58 per haplotype), counts
        __syncwarp();
        for (int i = lane; i < U * H; i += 32) {
            const int r = i / H, h = i - r * H;
            double prod = 1.0;
            for (int j = 0; j < N; j++) {
                double v = __ldg(R + ((size_t)r * N + j) * A + haps[h * N + j]);
                if (!isnan(v)) prod *= v;
            }
            tab[r * H + h] = prod;
        }
        for (int r = lane; r < U; r += 32) cnt[r] = a.counts ? (double)__ldg(a.counts + it.counts_off + r) : 1.0;
        __syncwarp();

        // ---- initial genotype: given or greedy (calling/mcmc.py:393-453)
        const int32_t *init_row = a.initial ? a.initial + (size_t)item_id * a.pstride : nullptr;
        if (init_row && init_row[0] >= 0) {
            for (int k = lane; k < P; k += 32) ini[k] = init_row[k];
            __syncwarp();
        } else {
            for (int i = 0; i < P; i++) {
                const int k = i + 1;
                const double dk = (double)k;
                double best = -INFINITY;
                int best_a = -1;
                for (int a0 = 0; a0 < H; a0 += 32) {
                    const int al = a0 + lane;
                    double lprob = -INFINITY;
                    if (al < H) {
                        double llk = 0.0;
                        for (int r = 0; r < U; r++) {
                            const double *row = tab + r * H;
                            double rp = 0.0;
                            for (int s = 0; s < i; s++) rp += row[ini[s]] / dk;
                            rp += row[al] / dk;
                            llk += log(rp) * cnt[r];
                        }
                        double lprior = 0.0;
                        if (has_prior) {
                            int g[MCHB_MAX_PLOIDY];
                            for (int s = 0; s < i; s++) g[s] = ini[s];
                            g[i] = al;
                            lprior = call_log_genotype_prior_slow(g, k, H, inbreeding, freqs);
                        }
                        lprob = llk + lprior;
                    }
                    // first maximum in allele order (strict >)
                    double v = lprob;
                    int idx = al < H ? al : 0x7fffffff;
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) {
                        double ov = __shfl_xor_sync(MCHB_FULL, v, m);
                        int oi = __shfl_xor_sync(MCHB_FULL, idx, m);
                        if (ov > v || (ov == v && oi < idx)) {
                            v = ov;
                            idx = oi;
                        }
                    }
                    if (v > best) {
                        best = v;
                        best_a = idx;
                    }
                }
                __syncwarp();
                ini[i] = best_a;
                __syncwarp();
            }
            // genotype.sort()
            for (int i = 1; i < P; i++) {
                int v = ini[i];
                int j = i - 1;
                __syncwarp();
                while (j >= 0) {  // uniform code on shared memory: read, barrier, then write
                    const int x = ini[j];
                    if (!(x > v)) break;
                    __syncwarp();
                    ini[j + 1] = x;
                    j--;
                }
                __syncwarp();
                ini[j + 1] = v;
                __syncwarp();
            }
        }
        // ---- the sampler works on t / P (likelihood.py:60)
        for (int i = lane; i < U * H; i += 32) tab[i] = tab[i] / dP;
        // ---- prior tables: lgA[a][c] = lgamma(alpha_a + c), c = 0..P
        double left_gibbs = 0.0, lg_left_full = 0.0, scale = 0.0;
        if (has_prior && inbreeding != 0.0) {
            scale = (1.0 - inbreeding) / inbreeding;
            const double alpha_const = (1.0 / (double)H) * scale;
            for (int i = lane; i < H * (P + 2); i += 32) {
                const int al = i / (P + 2), c = i - al * (P + 2);
                const double alpha = freqs ? freqs[al] * scale : alpha_const;
                lgA[i] = lgamma(alpha + (double)c);
            }
            double sum_alphas = 0.0;
            if (freqs) for (int al = 0; al < H; al++) sum_alphas += freqs[al] * scale;
            else sum_alphas = alpha_const * (double)H;
            // calling/prior.py:93-113: sum_alpha = (P - 1) + sum(alphas)
            const double sa = (double)(P - 1) + sum_alphas;
            left_gibbs = lgamma(sa) - lgamma(1.0 + sa);
            lg_left_full = (LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma(dP + sum_alphas);
        }
        __syncwarp();
        const double log_H = log((double)H);
        const double log_invH = log(1.0 / (double)H);

        int32_t *og = a.out_alleles + it.gl_off;
        double *ol = a.out_llks + it.hap_out_off;
        for (int chain = 0; chain < a.chains && !err; chain++) {
            __syncwarp();
            for (int k = lane; k < P; k += 32) gs[k] = ini[k];
            __syncwarp();
            for (int step = 0; step < a.steps && !err; step++) {
                // order = arange(P); np.random.shuffle(order)
                for (int k = lane; k < P; k += 32) ord[k] = k;
                __syncwarp();
                for (int i = P - 1; i > 0; i--) {
                    int k = ws.randint(i + 1);
                    int x = ord[i], y = ord[k];
                    __syncwarp();
                    ord[i] = y;
                    ord[k] = x;
                }
                __syncwarp();
                double llk_last = 0.0;
                for (int jj = 0; jj < P && !err; jj++) {
                    const int k = ord[jj];
                    if (memo) {
                        bool same0 = mvalid[2 * k] != 0, same1 = mvalid[2 * k + 1] != 0;
                        for (int s = lane; s < P; s += 32) {
                            const bool own = a.step_type == 0 && s == k;
                            same0 = same0 && (own || mkey[(2 * k) * P + s] == gs[s]);
                            same1 = same1 && (own || mkey[(2 * k + 1) * P + s] == gs[s]);
                        }
                        const bool hit0 = __all_sync(MCHB_FULL, same0), hit1 = __all_sync(MCHB_FULL, same1);
                        if (hit0 || hit1) {
                            // same draw from the remembered cumulative sums (jitutils.py:77-92)
                            const int way = hit0 ? 0 : 1;
                            const double *wcs = mcs + (size_t)(2 * k + way) * H, *wls = mls + (size_t)(2 * k + way) * H;
                            evals += H;
                            const double u = ws.next_double();
                            int choice = 0;
                            for (int a0 = 0; a0 < H; a0 += 32) {
                                const int al = a0 + lane;
                                choice += __popc(__ballot_sync(MCHB_FULL, al < H && wcs[al] <= u));
                            }
                            if (choice >= H) {
                                err = MCHB_ITEM_CHOICE_RANGE;
                                break;
                            }
                            llk_last = wls[choice];
                            __syncwarp();
                            gs[k] = choice;
                            if (lane == 0) mlru[k] = 1 - way;
                            __syncwarp();
                            continue;
                        }
                    }
                    const int current = gs[k];
                    int copies_cur = 0;  // mcmc.py:66-68
                    for (int s = 0; s < P; s++) copies_cur += (gs[s] == current);
                    for (int a0 = 0; a0 < H; a0 += 32) {
                        const int al = a0 + lane;
                        const int ae = al < H ? al : 0;
                        // log-likelihood of the genotype with slot k = al (likelihood.py:45-68 order)
                        double llk = 0.0;
                        for (int r = 0; r < U; r++) {
                            const double *row = tab + r * H;
                            double rp = 0.0;
                            for (int s = 0; s < P; s++) rp += row[s == k ? ae : gs[s]];
                            llk += log(rp) * cnt[r];
                        }
                        // copies of al among the genotype with slot k = al
                        int copies = 1;
                        for (int s = 0; s < P; s++) copies += (s != k && gs[s] == ae);
                        double lprior;
                        if (a.step_type == 0) {
                            if (!has_prior) {
                                lprior = LOG_INT[copies];                       // prior.py:30-52
                            } else if (inbreeding == 0.0) {
                                lprior = freqs ? log(freqs[ae]) : log_invH;     // prior.py:84-88
                            } else {
                                // prior.py:90-113 with constant_ibs = copies - 1
                                lprior = left_gibbs + (lgA[ae * (P + 2) + copies] - lgA[ae * (P + 2) + copies - 1]);
                            }
                        } else {
                            // full genotype prior (prior.py:116-179) of the proposed genotype
                            lprior = 0.0;
                            if (has_prior) {
                                double acc = 0.0;
                                const bool null_prior = inbreeding == 0.0;
                                for (int i = 0; i < P; i++) {
                                    const int ai = (i == k) ? ae : gs[i];
                                    int c = 0;
                                    bool first = true;
                                    for (int s = 0; s < P; s++) {
                                        const int as = (s == k) ? ae : gs[s];
                                        const bool eq = as == ai;
                                        c += eq;
                                        first = first && !(eq && s < i);
                                    }
                                    const int dose = first ? c : 0;
                                    if (null_prior) acc += LGAMMA_INT[dose + 1];
                                    else if (dose > 0)
                                        acc += lgA[ai * (P + 2) + dose] - (LGAMMA_INT[dose + 1] + lgA[ai * (P + 2)]);
                                }
                                if (null_prior) {
                                    const double ln_perms = LGAMMA_INT[P + 1] - acc;
                                    if (!freqs) lprior = ln_perms - dP * log_H;
                                    else {
                                        double prod = 1.0;
                                        for (int i = 0; i < P; i++) prod *= freqs[(i == k) ? ae : gs[i]];
                                        lprior = ln_perms + log(prod);
                                    }
                                } else {
                                    lprior = lg_left_full + acc;
                                }
                            }
                        }
                        if (al < H) {
                            ls[al] = llk;
                            xs[al] = (a.step_type == 0) ? llk + lprior : lprior;
                        }
                    }
                    evals += H;
                    __syncwarp();
                    // MH: llk and prior of the current genotype are the entries of the current allele
                    const double llk_cur = ls[current];
                    const double lprior_cur = xs[current];
                    // ---- probabilities
                    if (a.step_type == 0) {
                        // normalise_log_probs (jitutils.py:51-74): log-sum-exp then exp
                        double part = -INFINITY;
                        for (int a0 = 0; a0 < H; a0 += 32) {
                            const int al = a0 + lane;
                            if (al < H) part = add_log_prob(part, xs[al]);
                        }
                        const double denom = warp_lse(part);
                        for (int a0 = 0; a0 < H; a0 += 32) {
                            const int al = a0 + lane;
                            if (al < H) xs[al] = exp(xs[al] - denom);
                        }
                    } else {
                        // mcmc.py:123-132
                        double part = 0.0;
                        for (int a0 = 0; a0 < H; a0 += 32) {
                            const int al = a0 + lane;
                            if (al < H) {
                                int copies = 1;
                                for (int s = 0; s < P; s++) copies += (s != k && gs[s] == al);
                                const double lprop = (al == current) ? 0.0 : LOG_RATIO[copies * 17 + copies_cur];
                                const double mh = (ls[al] - llk_cur) + (xs[al] - lprior_cur) + lprop;
                                double p = exp(np_minimum0(mh));
                                if (al == current) p = 0;
                                p = p / (double)(H - 1);
                                xs[al] = p;
                                part += p;
                            }
                        }
                        const double sum = warp_sum(part);
                        __syncwarp();
                        if (lane == 0) xs[current] = 1 - sum;
                    }
                    __syncwarp();
                    // ---- random_choice (jitutils.py:77-92): cumsum, searchsorted right
                    const double u = ws.next_double();
                    double carry = 0.0;
                    int choice = 0;
                    for (int a0 = 0; a0 < H; a0 += 32) {
                        const int al = a0 + lane;
                        double v = al < H ? xs[al] : 0.0;
#pragma unroll
                        for (int m = 1; m < 32; m <<= 1) {
                            double o = __shfl_up_sync(MCHB_FULL, v, m);
                            if (lane >= m) v += o;
                        }
                        v += carry;
                        if (memo && al < H) {
                            const int way = mlru[k];
                            mcs[(size_t)(2 * k + way) * H + al] = v;
                            mls[(size_t)(2 * k + way) * H + al] = ls[al];
                        }
                        choice += __popc(__ballot_sync(MCHB_FULL, al < H && v <= u));
                        carry = __shfl_sync(MCHB_FULL, v, 31);
                    }
                    if (choice >= H) {
                        err = MCHB_ITEM_CHOICE_RANGE;
                        break;
                    }
                    llk_last = ls[choice];
                    __syncwarp();
                    if (memo) {
                        const int way = mlru[k];
                        for (int s = lane; s < P; s += 32) mkey[(2 * k + way) * P + s] = gs[s];
                        __syncwarp();
                        if (lane == 0) {
                            mvalid[2 * k + way] = 1;
                            mlru[k] = 1 - way;
                        }
                    }
                    __syncwarp();
                    gs[k] = choice;
                    __syncwarp();
                }
                if (err) break;
                // genotype_alleles.sort()
                for (int i = 1; i < P; i++) {
                    int v = gs[i];
                    int j = i - 1;
                    __syncwarp();
                    while (j >= 0) {  // uniform code on shared memory: read, barrier, then write
                        const int x = gs[j];
                        if (!(x > v)) break;
                        __syncwarp();
                        gs[j + 1] = x;
                        j--;
                    }
                    __syncwarp();
                    gs[j + 1] = v;
                    __syncwarp();
                }
                int32_t *dst = og + ((size_t)chain * a.steps + step) * P;
                for (int s = lane; s < P; s += 32) dst[s] = gs[s];
                if (lane == 0) ol[(size_t)chain * a.steps + step] = llk_last;
            }
        }
        int status = err;
        if (!status && ws.exhausted()) status = MCHB_ITEM_RNG_EXHAUSTED;
        if (lane == 0) {
            mchb_item_result r;
            r.status = status;
            r.n_het = 0;
            r.rng_words = ws.cur;
            r.llk_evals = evals;
            a.results[item_id] = r;
        }
        __syncwarp();
    }
}

}  // namespace mchb
