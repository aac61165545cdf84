/*
 * mchap_b200.h — C ABI of the B200-native MCHap inference hot path.
 *
 * One shared library (libmchap_b200.so), plain pointers and sizes, no torch / numpy types.
 * Every entry point returns an int status (MCHB_OK == 0); nothing throws across the ABI.
 * The reference (PlantandFoodResearch/MCHap v0.11.1) has no FFI layer: its boundary is the
 * Python call surface, so each entry point cites the reference function(s) it replaces
 * (paths relative to the reference repository).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Memory spaces: `mem` tells where the BULK arrays of a call live (reads, counts, haplotypes,
 * traces ...).  MCHB_MEM_HOST: the library stages them through the handle's stream
 * (H2D before, D2H after).  MCHB_MEM_DEVICE: they are device pointers on the handle's device
 * and no bulk copy is made.  Item descriptors, parameter structs and per-item result records
 * are always HOST memory.
 *
 * A handle owns one device, one stream and its scratch; calls on one handle are serialised,
 * different handles (e.g. one per host thread) run concurrently.
 */
#ifndef MCHAP_B200_H
#define MCHAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCHB_VERSION 1

/* ---- status codes (library level) ---- */
#define MCHB_OK 0
#define MCHB_ERR_CUDA 100         /* a CUDA runtime call failed: see mchb_last_error */
#define MCHB_ERR_ARGUMENT 101     /* inconsistent sizes / null pointers */
#define MCHB_ERR_NO_DEVICE 102    /* no usable sm_100 device: the library never falls back to the CPU */

/* ---- per-item status codes (mchb_item_result.status) ---- */
#define MCHB_ITEM_OK 0
#define MCHB_ITEM_NAN_LLK 1        /* reference: ValueError("Encountered log likelihood of nan"), assemble/mcmc.py:330-331 */
#define MCHB_ITEM_BREAKS 2         /* reference: ValueError("breaks must be smaller then n"), assemble/structural.py:49-50 */
#define MCHB_ITEM_CHOICE_RANGE 4   /* random_choice returned len(p) where the reference indexes out of range (jitutils.py:92) */
#define MCHB_ITEM_INITIAL_SHAPE 5  /* reference: AssertionError, assemble/mcmc.py:207 */
#define MCHB_ITEM_RNG_EXHAUSTED 6  /* pre-drawn word stream too short (the host shim retries with a longer one) */
#define MCHB_ITEM_UNSUPPORTED 8    /* shape outside the compiled limits (see mchb_limits); never silently approximated */
#define MCHB_ITEM_TALLY_OVERFLOW 9 /* trace tally: more distinct genotypes than max_unique (call again with a larger table) */

#define MCHB_MEM_HOST 0
#define MCHB_MEM_DEVICE 1
#define MCHB_MEM_LAST_TRACE 2  /* mchb_trace_tally_batch only: the trace that the last mchb_assemble_tally_batch left in the handle's scratch */

typedef struct mchb_handle mchb_handle;

typedef struct {
    int32_t max_ploidy;          /* 16 */
    int32_t max_key_bits;        /* 64: n_het_positions * bits_per_allele must fit */
    int32_t max_unique_reads;    /* 1024 per item for the warp-resident assemble kernels (subject to shared memory) */
    int32_t max_temperatures;    /* 8 */
    int32_t max_haplotypes;      /* 256 known haplotypes per locus for call / call-exact */
} mchb_limits;

typedef struct {
    int32_t status;       /* MCHB_ITEM_* */
    int32_t n_het;        /* assemble: positions left variable after homozygous fixing */
    int64_t rng_words;    /* 32-bit MT19937 words consumed by the item (== numba's cursor advance) */
    int64_t llk_evals;    /* log_likelihood evaluations performed (algorithmic work counter) */
} mchb_item_result;

/* ---- lifecycle ---- */
int mchb_create(int device, mchb_handle **out);
void mchb_destroy(mchb_handle *h);
const char *mchb_last_error(const mchb_handle *h);
void mchb_get_limits(mchb_limits *out);
/* device time (ms, CUDA events on the handle's stream) of the kernels of the last call and
 * how many kernels this library launched in it */
float mchb_last_kernel_ms(const mchb_handle *h);
int32_t mchb_last_kernel_launches(const mchb_handle *h);
/* host-buffer assemble calls: into how many chunks the last call cut the trace copy (1 = one
 * copy after the kernels; > 1 = chunk copies overlapped with the kernels) */
int32_t mchb_last_host_chunks(const mchb_handle *h);
/* the cudaStream_t the handle launches on (for external event timing) */
void *mchb_stream(const mchb_handle *h);
int mchb_sm_count(const mchb_handle *h);
/* warps the GPU held at once in the most populated assemble launch of the last call (one warp = one item):
 * callers that size batches in whole waves avoid the tail of a partially filled last wave */
int32_t mchb_last_resident_warps(const mchb_handle *h);

/* Page-locked host memory for the bulk arrays of MCHB_MEM_HOST calls (no reference counterpart: the
 * reference never leaves the host).  Copies from / to such buffers run at the full PCIe rate and
 * overlap with kernels; pageable buffers work too, through the driver's staging. */
int mchb_host_alloc(mchb_handle *h, int64_t bytes, void **out);
int mchb_host_free(mchb_handle *h, void *p);

/* Profiling aid (no reference counterpart): per-temperature cycle / event counters of the assemble
 * kernel, [8 temperatures][16 counters], filled only by a library built with -DMCHB_PROFILE
 * (profiles/phase_profile.py); the product build returns MCHB_ERR_ARGUMENT. */
int mchb_debug_counters(mchb_handle *h, uint64_t *out, int32_t n, int32_t reset);

/* Measurement aid (no reference counterpart): sustained FP64 FMA throughput of the device in
 * TFLOP/s from a register-resident DFMA loop; the roofline denominator for the FP64-SIMT-bound
 * MCMC kernels. */
int mchb_measure_fp64_peak(mchb_handle *h, double *out_tflops);

/* ---- RNG: numba's np.random MT19937 stream -------------------------------------------
 * Replaces: numba's thread-local generator as the reference drives it (jitutils.py:180-183
 * seed_numba; numba/cpython/randomimpl.py get_next_int32).  Fills `out` (n words, space `mem`)
 * with the tempered 32-bit output stream of init_genrand(seed). */
int mchb_mt19937_words(mchb_handle *h, int mem, uint32_t seed, uint32_t *out, int64_t n);

/* ---- multiset rank / unrank (bit-exact integers) -------------------------------------
 * Replaces: jitutils.py:253-276 genotype_alleles_as_index, 279-318 index_as_genotype_alleles
 * (VCF order).  alleles int64[n, ploidy] (sorted ascending per row), index int64[n].
 * unrank writes -1 alleles for index < 0 (the reference returns None). */
int mchb_genotype_rank(mchb_handle *h, int mem, const int64_t *alleles, int64_t n, int32_t ploidy,
                       int64_t *out_index);
int mchb_genotype_unrank(mchb_handle *h, int mem, const int64_t *index, int64_t n, int32_t ploidy,
                         int64_t *out_alleles);

/* ---- read log-likelihood, batched -----------------------------------------------------
 * Replaces: assemble/likelihood.py:18-70 log_likelihood for n (reads, genotype) pairs.
 * Item i: reads f64[U,N,A] at reads + reads_off[i]; counts i64[U] at counts + counts_off[i]
 * (counts == NULL: unweighted); genotype int8[P,N] at genotypes + geno_off[i]. */
typedef struct {
    int64_t reads_off, counts_off, geno_off;
    int32_t n_reads, n_pos, max_allele, ploidy;
} mchb_llk_item;

int mchb_log_likelihood_batch(mchb_handle *h, int mem, const mchb_llk_item *items, int64_t n_items,
                              const double *reads, int64_t reads_len, const int64_t *counts,
                              int64_t counts_len, const int8_t *genotypes, int64_t genotypes_len,
                              double *out_llk);

/* ---- de novo assembly MCMC, batched -----------------------------------------------------
 * Replaces: assemble/mcmc.py:103-265 DenovoMCMC.fit/_mcmc (homozygous fixing 495-541 via
 * snpcalling.py:14-70, initial state 455-491 + jitutils.py:465-498), 269-426 _denovo_assembler,
 * mutation.py:15-246, structural.py:23-673, tempering.py:11-151, prior.py:15-112 — one call
 * per batch of (locus, sample) items instead of one fit() per item.
 *
 * Item i: reads f64[U,N,A], counts i64[U] (or none), n_alleles int8[N], optional initial
 * int8[chains, P, initial_nhet]; outputs genotypes int8[chains, steps, P, N] at
 * out_genotypes + genotypes_off (fixed homozygous alleles re-inserted, haplotypes in sampler
 * order — the reference sorts them later in GenotypeMultiTrace.__post_init__) and llks
 * f64[chains, steps] at out_llks + llks_off.  The RNG stream of an item is the MT19937
 * stream of `seed` from its start (the reference re-seeds at the top of every fit), consumed
 * with numba's word->value rules in the reference's call order: chains one after another,
 * temperatures ascending inside a step.
 *
 * With MCHB_MEM_HOST outputs of 256 MB or more the batch is cut into chunks of consecutive
 * items and the traces of a finished chunk are copied to the host while later chunks are still
 * being sampled (see mchb_last_host_chunks); page-locked output buffers make those copies
 * asynchronous, pageable ones work too. */
typedef struct {
    int64_t reads_off;       /* doubles */
    int64_t counts_off;      /* int64 elements; ignored if counts == NULL */
    int64_t nalleles_off;    /* int8 elements */
    int64_t initial_off;     /* int8 elements; -1: sample the initial state like the reference */
    int64_t genotypes_off;   /* int8 elements into out_genotypes */
    int64_t llks_off;        /* doubles into out_llks */
    int32_t n_reads, n_pos, max_allele, ploidy;
    int32_t temps_off, n_temps; /* temperatures[temps_off .. +n_temps): ascending, last == 1.0 */
    int32_t initial_nhet;
    uint32_t seed;
    double inbreeding;       /* NaN: flat prior (reference inbreeding=None) */
} mchb_assemble_item;

typedef struct {
    int32_t steps, chains;
    double fix_homozygous;
    double p_recombination, p_partial_dosage, p_dosage;
    /* row n (0..break_rows-1) = distribution of the number of break points when n positions
     * stay variable (assemble/mcmc.py:211-217; rows are computed on the host with scipy exactly
     * like _point_beta_probabilities 429-452); f64[break_rows * break_stride], lengths int32 */
    const double *break_table;
    const int32_t *break_len;
    int32_t break_rows, break_stride;
    const double *temperatures; /* pool indexed by item.temps_off */
    int32_t temperatures_len;
    /* != 0: every recorded step has its haplotypes sorted lexicographically (first position most
     * significant), the form GenotypeMultiTrace.__post_init__ gives a trace (assemble/classes.py:265-278,
     * encoding/integer/sequence.py:78-110); 0: the sampler's own row order (the order the reference's
     * _denovo_assembler returns, assemble/mcmc.py:418-425) */
    int32_t sort_haplotypes;
    /* replay harness: when replay_words != NULL every item reads this pre-drawn (tempered)
     * 32-bit stream from word 0 instead of MT19937(seed) (always HOST memory) */
    const uint32_t *replay_words;
    int64_t replay_len;
    int64_t rng_words_hint;  /* 0: estimate; else initial per-stream length in words */
} mchb_assemble_params;

int mchb_assemble_batch(mchb_handle *h, int mem, const mchb_assemble_params *params,
                        const mchb_assemble_item *items, int64_t n_items, const double *reads,
                        int64_t reads_len, const int64_t *counts, int64_t counts_len,
                        const int8_t *n_alleles, int64_t n_alleles_len, const int8_t *initial,
                        int64_t initial_len, int8_t *out_genotypes, int64_t out_genotypes_len,
                        double *out_llks, int64_t out_llks_len, mchb_item_result *results);


/* ---- exhaustive genotype calling over known haplotypes (mchap call-exact) ----------------
 * Item i: reads f64[U,N,A], counts i64[U] (or none), haplotypes int8[H,N] in VCF allele order,
 * optional prior = (inbreeding, frequencies f64[H] or flat).  G = C(H+P-1, P) genotypes are
 * enumerated in VCF order (jitutils.py:113-146). */
typedef struct {
    int64_t reads_off;       /* doubles */
    int64_t counts_off;      /* int64 elements; ignored if counts == NULL */
    int64_t haps_off;        /* int8 elements: haplotypes[H, N] */
    int64_t freqs_off;       /* doubles: prior allele frequencies [H]; -1: flat frequencies (None) */
    int64_t hap_out_off;     /* element offset of the item's [H] row in per-haplotype outputs */
    int64_t gl_off;          /* element offset of the item's [G] row in per-genotype arrays */
    int32_t n_reads, n_pos, max_allele, ploidy, n_haps;
    int32_t reserved;
    double inbreeding;       /* NaN: prior None */
} mchb_call_item;

/* Replaces: calling/exact.py:156-249 posterior_mode with every optional result
 * (17-61 _call_posterior_mode, 64-105 _genotype_support_log_joint, 108-153
 * _posterior_allele_frequencies), all in float64 like the reference's low-memory branch.
 * out_alleles i64[n_items, pstride] (mode alleles, sorted; unused slots -2);
 * out_stats f64[n_items, 4] = mode llk, mode probability, support probability, log normaliser;
 * out_freqs / out_occur f64 rows at hap_out_off (posterior mean allele frequencies, occurrence). */
int mchb_call_exact_mode_batch(mchb_handle *h, int mem, const mchb_call_item *items, int64_t n_items,
                               const double *reads, int64_t reads_len, const int64_t *counts,
                               int64_t counts_len, const int8_t *haplotypes, int64_t haplotypes_len,
                               const double *freqs, int64_t freqs_len, int64_t *out_alleles,
                               int32_t pstride, double *out_stats, double *out_freqs,
                               double *out_occur, int64_t hap_out_len, mchb_item_result *results);

/* Replaces: calling/exact.py:252-292 genotype_likelihoods — float32[G] per item at gl_off, like
 * the reference (exact.py:254). */
int mchb_genotype_likelihoods_batch(mchb_handle *h, int mem, const mchb_call_item *items,
                                    int64_t n_items, const double *reads, int64_t reads_len,
                                    const int64_t *counts, int64_t counts_len,
                                    const int8_t *haplotypes, int64_t haplotypes_len, float *out_gl,
                                    int64_t gl_len, mchb_item_result *results);

/* Replaces: calling/exact.py:295-329 genotype_posteriors (+ 332-369
 * posterior_allele_frequencies when out_freqs != NULL).  llks: float32 (llk_is_f32 != 0; the
 * sum llk + log-prior is rounded to float32 before the float64 normalisation exactly as numba
 * does for a float32 input array) or float64.  out_gp f64[G] rows at gl_off. */
int mchb_genotype_posteriors_batch(mchb_handle *h, int mem, const mchb_call_item *items,
                                   int64_t n_items, const double *freqs, int64_t freqs_len,
                                   const void *llks, int llk_is_f32, int64_t gl_len, double *out_gp,
                                   double *out_freqs, double *out_counts, double *out_occur,
                                   int64_t hap_out_len);

/* ---- genotype calling MCMC over known haplotypes (mchap call) -----------------------------
 * Replaces: calling/classes.py:49-124 CallingMCMC.fit (greedy_caller 393-453 for the initial
 * genotype, mcmc_sampler 330-390, compound_step 232-327, gibbs_options 143-229 / mh_options
 * 15-140).  Items are mchb_call_item records with two fields re-used: `reserved` = RNG seed
 * (uint32), gl_off = element offset of the item's int32[chains, steps, P] trace in out_alleles,
 * hap_out_off = element offset of its f64[chains, steps] llks in out_llks.
 * initial: int32[n_items, pstride] or NULL; a row whose first entry is negative asks for the
 * reference's greedy initialisation. */
typedef struct {
    int32_t steps, chains;
    int32_t step_type;           /* 0 = Gibbs, 1 = Metropolis-Hastings */
    int32_t reserved;
    const uint32_t *replay_words; /* replay harness (HOST memory) or NULL */
    int64_t replay_len;
    int64_t rng_words_hint;
} mchb_call_mcmc_params;

int mchb_call_mcmc_batch(mchb_handle *h, int mem, const mchb_call_mcmc_params *params,
                         const mchb_call_item *items, int64_t n_items, const double *reads,
                         int64_t reads_len, const int64_t *counts, int64_t counts_len,
                         const int8_t *haplotypes, int64_t haplotypes_len, const double *freqs,
                         int64_t freqs_len, const int32_t *initial, int32_t pstride,
                         int32_t *out_alleles, int64_t out_alleles_len, double *out_llks,
                         int64_t out_llks_len, mchb_item_result *results);

/* ---- trace post-processing (tallies of a de novo assembly trace) ---------------------------
 * Replaces, for a batch of traces: assemble/classes.py:265-278 GenotypeMultiTrace.__post_init__
 * (per-step lexicographic haplotype sort, encoding/integer/sequence.py:78-110), 280-305 burn,
 * 307-325 posterior and 327-339 split (unique genotypes in first-occurrence order and their
 * counts, mset.py:242-284, 361-392) — the inputs of mode_genotype_support (classes.py:87-128)
 * and replicate_incongruence (341-376).
 *
 * Item i: trace int8[chains, steps, ploidy, n_pos] at genotypes + genotypes_off (haplotypes in
 * any order).  Steps burn..steps-1 of every chain are tallied.  Outputs, per item:
 *   out_states int8[max_unique, ploidy, n_pos] at states_off: the distinct genotypes, haplotypes
 *     sorted, in order of first occurrence in the chain-major flattened trace;
 *   out_counts int32[max_unique, chains] at tallies_off: occurrences of state u in chain c;
 *   out_first  int32[max_unique, chains] at tallies_off: step (after burn) of the first
 *     occurrence of state u in chain c, -1 if it never occurs there (the per-chain
 *     first-occurrence order);
 *   results[i].n_het = number of distinct genotypes; status MCHB_ITEM_TALLY_OVERFLOW if there
 *     are more than max_unique (the arrays then hold the first max_unique states, counts are
 *     incomplete). */
typedef struct {
    int64_t genotypes_off;   /* int8 elements into genotypes */
    int64_t states_off;      /* int8 elements into out_states */
    int64_t tallies_off;     /* int32 elements into out_counts and out_first */
    int32_t n_pos, ploidy, chains, steps;
    int32_t burn, max_unique;
} mchb_tally_item;

/* mem_in: where `genotypes` lives (MCHB_MEM_LAST_TRACE: `genotypes` is ignored and the offsets
 * index the trace kept by the last mchb_assemble_tally_batch call on this handle, e.g. to tally
 * the few items that overflowed max_unique again with a larger table, without sampling again);
 * mem_out: where the three output arrays live. */
int mchb_trace_tally_batch(mchb_handle *h, int mem_in, int mem_out, const mchb_tally_item *items,
                           int64_t n_items, const int8_t *genotypes, int64_t genotypes_len,
                           int8_t *out_states, int64_t out_states_len, int32_t *out_counts,
                           int32_t *out_first, int64_t tallies_len, mchb_item_result *results);

/* mchb_assemble_batch followed by mchb_trace_tally_batch with the trace kept on the device: the
 * inputs and the tallies are HOST memory, the traces never leave HBM (they live in the handle's
 * scratch; items[i].genotypes_off / llks_off index that scratch like in mchb_assemble_batch and
 * tally_items[i].genotypes_off must repeat items[i].genotypes_off).  This is the call behind
 * `DenovoMCMC(...).fit(...).burn(n).posterior()` of mchap/application/assemble.py:123-170 for
 * a batch.  results: the assemble results; tally_results: the tally results. */
int mchb_assemble_tally_batch(mchb_handle *h, const mchb_assemble_params *params,
                              const mchb_assemble_item *items, const mchb_tally_item *tally_items,
                              int64_t n_items, const double *reads, int64_t reads_len,
                              const int64_t *counts, int64_t counts_len, const int8_t *n_alleles,
                              int64_t n_alleles_len, const int8_t *initial, int64_t initial_len,
                              int64_t genotypes_len, int64_t llks_len, int8_t *out_states,
                              int64_t out_states_len, int32_t *out_counts, int32_t *out_first,
                              int64_t tallies_len, mchb_item_result *results,
                              mchb_item_result *tally_results);

/* The same for calling traces (mchap call): replaces calling/classes.py:178-242
 * GenotypeAllelesMultiTrace.burn / posterior / split and feeds replicate_incongruence (221-256) and
 * posterior_frequencies (258-297).  Trace int32[chains, steps, ploidy] at alleles + genotypes_off;
 * tally items have n_pos == 1 (a "row" is one allele index; the sampler already keeps the alleles
 * of a step sorted, calling/mcmc.py:325-326); out_states is int32[max_unique, ploidy]. */
int mchb_call_trace_tally_batch(mchb_handle *h, int mem_in, int mem_out, const mchb_tally_item *items,
                                int64_t n_items, const int32_t *alleles, int64_t alleles_len,
                                int32_t *out_states, int64_t out_states_len, int32_t *out_counts,
                                int32_t *out_first, int64_t tallies_len, mchb_item_result *results);

/* mchb_call_mcmc_batch followed by mchb_call_trace_tally_batch with the trace kept on the device
 * (HOST inputs and tallies; tally_items[i].genotypes_off must repeat items[i].gl_off): the call
 * behind `CallingMCMC(...).fit(...).burn(n)` of mchap/application/call.py:134-182 for a batch. */
int mchb_call_mcmc_tally_batch(mchb_handle *h, const mchb_call_mcmc_params *params,
                               const mchb_call_item *items, const mchb_tally_item *tally_items,
                               int64_t n_items, const double *reads, int64_t reads_len,
                               const int64_t *counts, int64_t counts_len, const int8_t *haplotypes,
                               int64_t haplotypes_len, const double *freqs, int64_t freqs_len,
                               const int32_t *initial, int32_t pstride, int64_t alleles_len,
                               int64_t llks_len, int32_t *out_states, int64_t out_states_len,
                               int32_t *out_counts, int32_t *out_first, int64_t tallies_len,
                               mchb_item_result *results, mchb_item_result *tally_results);

/* ---- read encoding + de-duplication ------------------------------------------------------
 * Replaces, for a batch of (locus, sample) items, what mchap/application/baseclass.py:194-209
 * does between the BAM reader and the samplers: encoding/integer/transcode.py:16-77
 * as_probabilistic (through io/bam.py:251-289 encode_read_distributions) followed by
 * mset.unique_counts (mset.py:242-284, 361-392).
 *
 * Item i: calls int8[n_reads, n_pos] (allele index per read and position, < 0 = gap) at
 * calls + calls_off; probs f64[n_reads, n_pos] = probability that the call is correct (the host
 * computes (1 - error_rate) * prob_of_qual(quals) like io/bam.py:281-286) at probs + probs_off;
 * n_alleles int8[n_pos] at n_alleles + nalleles_off.  Outputs: the distinct encoded reads
 * f64[n_unique, n_pos, max_allele] in order of first occurrence at out_reads + reads_off
 * (capacity n_reads rows) and their counts int64[n_unique] at out_counts + counts_off (capacity
 * n_reads); results[i].n_het = n_unique.  Element (r, j, a): probs[r, j] if a is the call,
 * (1 - probs[r, j]) / error_factor otherwise, NaN (numpy's np.nan bit pattern) at gaps, 0 where
 * a >= n_alleles[j] — bit-identical to the reference, so that the byte-wise de-duplication and
 * the order of the unique reads (which fixes the summation order of every log-likelihood) are
 * the reference's. */
typedef struct {
    int64_t calls_off;       /* int8 elements */
    int64_t probs_off;       /* doubles */
    int64_t nalleles_off;    /* int8 elements */
    int64_t reads_off;       /* doubles into out_reads */
    int64_t counts_off;      /* int64 elements into out_counts */
    int32_t n_reads, n_pos, max_allele, reserved;
} mchb_encode_item;

int mchb_encode_reads_batch(mchb_handle *h, int mem, const mchb_encode_item *items, int64_t n_items,
                            const int8_t *calls, int64_t calls_len, const double *probs,
                            int64_t probs_len, const int8_t *n_alleles, int64_t n_alleles_len,
                            double error_factor, double *out_reads, int64_t out_reads_len,
                            int64_t *out_counts, int64_t out_counts_len, mchb_item_result *results);

/* The whole per-sample device path of mchap/application/assemble.py:95-170 for a batch, one call:
 * mchb_encode_reads_batch -> mchb_assemble_batch -> mchb_trace_tally_batch with every intermediate
 * (encoded unique reads, counts, traces) kept in the handle's scratch on the device.  HOST memory
 * in (calls, probs, n_alleles) and out (tallies).  assemble_items[i] carries ploidy, temperatures,
 * seed, inbreeding, genotypes_off / llks_off (into the scratch trace) and initial_off = -1; its
 * reads_off / counts_off / nalleles_off / n_reads / n_pos / max_allele are filled in by the library
 * from encode_items[i] and the number of distinct reads.  tally_items as in
 * mchb_assemble_tally_batch.  encode_results[i].n_het = distinct reads of item i. */
int mchb_encode_assemble_tally_batch(mchb_handle *h, const mchb_assemble_params *params,
                                     const mchb_encode_item *encode_items,
                                     const mchb_assemble_item *assemble_items,
                                     const mchb_tally_item *tally_items, int64_t n_items,
                                     const int8_t *calls, int64_t calls_len, const double *probs,
                                     int64_t probs_len, const int8_t *n_alleles, int64_t n_alleles_len,
                                     double error_factor, int64_t genotypes_len, int64_t llks_len,
                                     int8_t *out_states, int64_t out_states_len, int32_t *out_counts,
                                     int32_t *out_first, int64_t tallies_len,
                                     mchb_item_result *encode_results, mchb_item_result *results,
                                     mchb_item_result *tally_results);

/* ---- minimum error correction ------------------------------------------------------------
 * Replaces: encoding/integer/stats.py:18-39 minimum_error_correction as the CLIs use it
 * (application/assemble.py:158-163, call.py:170-175, call_exact.py:186-191): for every read the
 * number of called positions (call >= 0) at which it differs from the closest haplotype of the
 * called genotype.  Item i: calls int8[n_reads, n_pos] at calls + calls_off, genotype
 * int8[ploidy, n_pos] at genotypes + geno_off.  out_mec[i] = sum over reads; out_called[i]
 * (optional) = number of calls >= 0; out_per_read (optional) int32 rows at per_read_off. */
typedef struct {
    int64_t calls_off;       /* int8 elements */
    int64_t geno_off;        /* int8 elements */
    int64_t per_read_off;    /* int32 elements into out_per_read */
    int32_t n_reads, n_pos, ploidy, reserved;
} mchb_mec_item;

int mchb_mec_batch(mchb_handle *h, int mem, const mchb_mec_item *items, int64_t n_items,
                   const int8_t *calls, int64_t calls_len, const int8_t *genotypes, int64_t genotypes_len,
                   int64_t *out_mec, int64_t *out_called, int32_t *out_per_read, int64_t per_read_len);

#ifdef __cplusplus
}
#endif
#endif /* MCHAP_B200_H */
