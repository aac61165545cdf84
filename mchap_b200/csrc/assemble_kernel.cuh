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
//   * log-sum over reads = per-lane log() + xor-butterfly (uniform result in all lanes);
//   * the steady-state loop is kept small enough for the instruction cache: one copy of each
//     step function (structural sub-steps share one call site), rarely executed set-up code in
//     separate non-inlined functions, the prior code compiled only into the PRIOR variant.
#pragma once
#include "common.cuh"

// resident CTAs per SM the register allocation is tuned for (4 warps per CTA)
#ifndef MCHB_ASM_MINBLOCKS
#define MCHB_ASM_MINBLOCKS 4
#endif
// Launch bounds per read-chunk count.  One and two chunks: 4 CTAs of 4 warps, 128 registers (16 warps
// per SM; shared memory allows as many at the headline shape).  Three and four chunks: CTAs of one or
// two warps, 204 registers, so that the 10 warps per SM that shared memory allows at configs[3] fit
// the register file.  Eight and more chunks: one or two warps per SM, no register cap.
#define MCHB_ASM_MAXTHREADS(CH) (((CH) == 3 || (CH) == 4) ? 64 : 128)
#define MCHB_ASM_MINCTAS(CH) ((CH) <= 2 ? MCHB_ASM_MINBLOCKS : ((CH) <= 4 ? 5 : 1))

namespace mchb {

// Profiling build (-DMCHB_PROFILE, profiles/phase_profile.py): per-temperature cycle and event
// counters of the assemble kernel, read back with mchb_debug_counters.  Absent from the product build.
#ifdef MCHB_PROFILE
__device__ unsigned long long g_asm_prof[8 * 16];
#define MCHB_PROF_ADD(t, k, v)                                                                   \
    do {                                                                                         \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_asm_prof[((t) & 7) * 16 + (k)], (unsigned long long)(v)); \
    } while (0)
#define MCHB_PROF_CLOCK() clock64()
#else
#define MCHB_PROF_ADD(t, k, v) do { (void)(v); } while (0)
#define MCHB_PROF_CLOCK() 0ll
#endif
enum { PK_CYC_MUT = 0, PK_CYC_STR, PK_CYC_SLOT, PK_MUT_ACCEPT, PK_WINDOWS, PK_T2A_EVALS, PK_T2B_WINDOWS, PK_BASE_STEPS,
       PK_STR_EXACT, PK_STR_SCREEN_STAY, PK_STR_MEMO_STAY, PK_INTERVALS, PK_CYC_SWAPSTEP, PK_STR_ACCEPT, PK_T2B_DONE };

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
    // host-buffer path: items finished per chunk of chunk_items consecutive item ids; the copy
    // stream waits on these counters (stream memory operation) before moving a chunk's traces
    uint32_t *chunk_done;       // may be null
    uint32_t *chunk_flags;      // host-mapped: set to 1 by the warp that finishes a chunk's last item
    int32_t chunk_items, n_items;
    // shared memory geometry (per warp)
    int32_t nmax, amax, pmax, tmax, maxopt;
    int32_t smem_per_warp;      // bytes
    // byte offsets of the per-warp arrays (host computed, asm_layout()); Rt is at offset 0
    int32_t o_cnt, o_q, o_dist, o_oll, o_opr, o_lgdisp, o_homlp, o_llk_t, o_key, o_sc, o_perm, o_het, o_fixa,
        o_nall, o_opt0, o_opt1, o_ivb, o_ivp, o_ring, o_q32, o_rat, o_c32, o_rpc, o_bcs, o_epoch, o_mcache, o_scache,
        o_wmap, o_inv, o_hot, o_hmap, o_rank;
    int32_t scache_n, scache_tri; // entries of the structural memo per slot; > 0: directly indexed, intervals per type
    int32_t sort_recorded;       // record every step with its haplotypes sorted (assemble/classes.py:265-278)
    // tres = state slots resident in shared memory: tmax normally; 1 when the per-temperature
    // tables of a large shape would otherwise leave only one warp per SM — then the slot of the
    // temperature being stepped is swapped in from / out to slot_backing (global memory, L2)
    int32_t tres;
    int32_t slot_bytes;          // bytes of one slot in the backing store
    unsigned char *slot_backing; // [grid warps][tmax][slot_bytes], only if tres < tmax
    // The transposed reads tensor Rt is the largest per-item table and the exact paths (row
    // recomputation) are its only readers: the one-chunk kernels keep it in shared memory, the
    // kernels of larger items in a per-warp region of global memory (L2 resident), which roughly
    // halves their shared memory per warp and lets twice as many warps share an SM.
    double *rt_backing;          // [grid warps][rt_stride], only if CH >= 2
    int64_t rt_stride;
    double *q_backing;           // [grid warps][q_stride] = [tmax][pmax][UPAD] rows + 2 spare rows, only if CH >= 3
    int64_t q_stride;
};

// Kernels of three and more read chunks also keep the double-precision product rows q[slot][h][r]
// in global memory, every state slot in its own place (the resident-slot swap then moves only the
// float32 shadows, keys and memo tables): configs[3] goes from 40 to 30 KB of shared memory per warp.
#define MCHB_ASM_Q_GLOBAL(CH) ((CH) >= 3)
// ... and hold one direction of the float32 allele-ratio table, rat[j][r] = fl(R[j][1] / R[j][0]); the
// other direction is its correctly rounded reciprocal, taken on the fly (one rounding more per factor,
// accounted for in the error bounds of assemble_item_setup).  Ratios below the float32 normal range are
// stored as 0 so that no denormal is ever inverted.
#define MCHB_ASM_RAT_HALF(CH) ((CH) >= 3)

template <int CH>
__device__ __forceinline__ double *asm_rt(const AsmArgs &a, unsigned char *sm) {
    if (CH == 1) return reinterpret_cast<double *>(sm);
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    return a.rt_backing + (size_t)warp_global * a.rt_stride;
}

// uniform per-item scalars parked in shared memory (sc[]) to keep them out of registers
enum { SC_LUH = 0, SC_LG_SUMDISP, SC_LG_P_SUMDISP, SC_LG_DISP, SC_INBREEDING, SC_ERR_MUT, SC_ERR_STR, SC_COUNT };

// Float32 screening: a proposal is settled without an exact evaluation only when its screened
// Metropolis-Hastings ratio is below log(t) by more than  temp * E + MCHB_SCREEN_SLACK,  where E
// (SC_ERR_MUT / SC_ERR_STR, per item) bounds |screened llk - exact llk| — derivation in DESIGN.md
// section 4 'Error bound of the float32 screening' — and the slack covers the float32 log of the
// uniform (<= 3.2e-5), the float conversion of t and the double roundings of the exact path.
#ifndef MCHB_SCREEN_SLACK
#define MCHB_SCREEN_SLACK 1.0e-3
#endif
// Test aid (profiles/bound_sensitivity.sh): scales both error bounds; a build with a bound that is
// too small must be caught by the parity harness (it is: see profiles/r02_bound_sensitivity.txt).
#ifndef MCHB_SCREEN_ERR_SCALE
#define MCHB_SCREEN_ERR_SCALE 1.0
#endif
// reads whose screened probability falls below this fraction of the current one make the sub-step
// needy (the error bound is proportional to 1 / this)
#define MCHB_SCREEN_MIN_RATIO 1.0e-4f

// more needy sub-steps than this in one window: evaluate the window with the lane-parallel exact
// loop instead of one exact evaluation per needy sub-step
#define MCHB_EXACT_SERIAL_MAX 8

// Per-slot memo of screened structural steps, keyed by (type, start, stop).  When the table can hold
// every key of the class's longest item — 2 types x nmax (nmax + 1) / 2 intervals: 72 entries at 8
// SNVs — it is indexed directly (no collisions: a chain sitting in a mode never screens an interval
// twice); otherwise it is a direct-mapped hash table of MCHB_SCACHE_HASH_N(CH) entries.  The budgets
// keep 16 warps per SM at the headline shape (one and two read chunks) and are small where shared
// memory decides how many warps an SM holds (three and more chunks).
#define MCHB_SCACHE_HASH_LOG2(CH) ((CH) == 1 ? 7 : ((CH) == 2 ? 6 : 5))
#define MCHB_SCACHE_HASH_N(CH) (1 << MCHB_SCACHE_HASH_LOG2(CH))
#define MCHB_SCACHE_DIRECT_MAX(CH) ((CH) == 1 ? 128 : ((CH) == 2 ? 96 : 32))
struct ScEntry {
    uint32_t epoch;     // state epoch of the slot when the entry was filled
    uint32_t key;       // (type, start, stop)
    float smax;         // max over options of the screened mh (+inf: screening not conclusive)
    float temp;         // inverse temperature the entry was computed for
    int32_t n_options;
    int32_t pad;
};

// nibble helpers for packed small-integer vectors (labels, SNP genotypes), P <= 16
__device__ __forceinline__ int nib(uint64_t v, int i) { return (int)((v >> (4 * i)) & 15u); }
__device__ __forceinline__ uint64_t nib_set(uint64_t v, int i, int x) {
    return (v & ~(15ull << (4 * i))) | ((uint64_t)x << (4 * i));
}

// increment a nibble-packed sorted genotype to its VCF-order successor (jitutils.py:113-146)
__device__ __noinline__ uint64_t increment_packed(uint64_t g, int P) {
    if (P == 1) return g + 1;
    int prev = nib(g, 0);
#pragma unroll 1
    for (int i = 1; i < P; i++) {
        int al = nib(g, i);
        if (al == prev) continue;
        g = nib_set(g, i - 1, nib(g, i - 1) + 1);  // al > prev for sorted input
        g &= ~((1ull << (4 * (i - 1))) - 1ull);
        return g;
    }
    g = nib_set(g, P - 1, nib(g, P - 1) + 1);
    g &= ~((1ull << (4 * (P - 1))) - 1ull);
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
#pragma unroll 1
    for (int i = 0; i < P; i++) {
        int cntv = 0;
        bool first = true;
#pragma unroll 1
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

// ---------------------------------------------------------------------------------------
// structural.py: options of a label matrix (lin = inside-interval labels, lout = outside labels,
// nibble packed).  type 0: recombination (75-178), type 1: dosage swap (182-307).  Returns the
// number of options; when o0 != nullptr the (h0, h1) pairs are written in the reference's order.
// All tests are nibble-parallel bit operations on masks with one bit per haplotype (bit 4h):
//   ff: row h is not a duplicate of an earlier row          (haplotype_dosage[h] != 0)
//   sf: h is the first haplotype carrying its inside segment (segment_dosage[h] != 0)
//   ss: h is the only haplotype carrying its inside segment  (segment_dosage[h] == 1)
// ---------------------------------------------------------------------------------------
template <typename W>
__device__ __forceinline__ W nib_eq(W v, int x, W ones) {
    W t = v ^ (ones * (W)x);  // zero nibble where equal
    t |= t >> 1;
    t |= t >> 2;              // bit 0 of every nibble = OR of the nibble's bits
    return ~t & ones;
}

template <typename W>
__device__ __forceinline__ int structural_options_w(W lin, W lout, int P, int type, uint8_t *o0, uint8_t *o1) {
    const W all = (4 * P >= (int)(8 * sizeof(W))) ? ~(W)0 : ((((W)1) << (4 * P)) - 1);
    const W ones = (W)0x1111111111111111ull & all;
    W ff = 0, sf = 0, ss = 0;
#pragma unroll 1
    for (int h = 0; h < P; h++) {
        const int sh = 4 * h;
        const W eqI = nib_eq<W>(lin, (int)((lin >> sh) & 15), ones);
        const W eqO = nib_eq<W>(lout, (int)((lout >> sh) & 15), ones);
        const W bit = ((W)1) << sh;
        const W below = bit - 1;
        if (!(eqI & eqO & below)) ff |= bit;
        if (!(eqI & below)) sf |= bit;
        if (!(eqI & (eqI - 1))) ss |= bit;
    }
    int n = 0;
#pragma unroll 1
    for (int h0 = 0; h0 < P; h0++) {
        const int sh = 4 * h0;
        const W bit = ((W)1) << sh;
        if (!(ff & bit)) continue;               // duplicate copy of a haplotype
        if (type == 1 && (ss & bit)) continue;   // would delete the only copy of the segment
        const W eqI = nib_eq<W>(lin, (int)((lin >> sh) & 15), ones);
        W cand;
        if (type == 0) {
            const W eqO = nib_eq<W>(lout, (int)((lout >> sh) & 15), ones);
            cand = ff & ~eqI & ~eqO & ~((bit << 1) - 1);  // h1 > h0, both segments differ
        } else {
            cand = sf & ~eqI;                             // first carrier of a different segment
        }
        if (!o0) {
            n += (sizeof(W) == 8) ? __popcll((unsigned long long)cand) : __popc((unsigned)cand);
        } else {
            while (cand) {
                const int b = (sizeof(W) == 8) ? __ffsll((long long)cand) - 1 : __ffs((int)cand) - 1;
                o0[n] = (uint8_t)h0;
                o1[n] = (uint8_t)(b >> 2);
                n++;
                cand &= cand - 1;
            }
        }
    }
    return n;
}

__device__ __noinline__ int structural_options(uint64_t lin, uint64_t lout, int P, int type, uint8_t *o0,
                                               uint8_t *o1) {
    if (P <= 8) return structural_options_w<uint32_t>((uint32_t)lin, (uint32_t)lout, P, type, o0, o1);
    return structural_options_w<uint64_t>(lin, lout, P, type, o0, o1);
}

// structural.py:311-430 haplotype_segment_labels on packed keys: first-occurrence labels of the
// inside-interval (.x) and outside-interval (.y) segments
__device__ __noinline__ ulonglong2 segment_labels(const uint64_t *ks, int P, uint64_t mask_in) {
    uint64_t li = 0, lo = 0;
#pragma unroll 1
    for (int h = 1; h < P; h++) {
        const uint64_t kk = ks[h];
        int fi = h, fo = h;
#pragma unroll 1
        for (int k = h - 1; k >= 0; k--) {
            const uint64_t d = ks[k] ^ kk;
            if ((d & mask_in) == 0) fi = k;
            if ((d & ~mask_in) == 0) fo = k;
        }
        li |= (uint64_t)fi << (4 * h);
        lo |= (uint64_t)fo << (4 * h);
    }
    return make_ulonglong2(li, lo);
}

// assemble/prior.py:15-112 from P comparable row identifiers (keys or label pairs) held in a
// small uniform array; the first-occurrence dosage (jitutils.get_haplotype_dosage 377-422) is
// evaluated on the fly and the terms are accumulated in row order like the reference.
__device__ __noinline__ double assemble_prior_rows(const uint64_t *rows, int P, const double *scv,
                                                   const double *lgd) {
    const bool null_prior = scv[SC_INBREEDING] == 0.0;
    const double lg_disp = scv[SC_LG_DISP];
    double acc = 0.0;
#pragma unroll 1
    for (int i = 0; i < P; i++) {
        const uint64_t ki = rows[i];
        int cntv = 0;
        bool first = true;
#pragma unroll 1
        for (int k = 0; k < P; k++) {
            bool eq = rows[k] == ki;
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

// ---------------------------------------------------------------------------------------
// The two numeric building blocks of the lanes-over-reads paths, kept out of line so that the
// steady-state loop holds exactly one copy of each (the kernel is instruction-cache bound).
// ---------------------------------------------------------------------------------------
struct RowGeom {
    int N, A, P, B;
    uint32_t amask;
    int pow2;
    double invP;
};

// likelihood.py:48-60 for one haplotype key: row[r] = (prod_j Rt[j][allele_j][r]) / ploidy
template <int CH>
__device__ __noinline__ void compute_row(const double *Rt_lane, double *row_lane, float *row32_lane, uint64_t k,
                                         RowGeom g) {
    constexpr int UPAD = CH * 32;
    double out[CH];
#pragma unroll
    for (int ch = 0; ch < CH; ch++) out[ch] = 1.0;
    const double *base = Rt_lane;
    const int pstride = g.A * UPAD;
    // (the kernels of three and more read chunks run one warp per scheduler: their loops are unrolled so that
    // the shared-memory loads overlap; the one-chunk kernel is instruction-cache bound and is not)
#pragma unroll(CH >= 3 ? 4 : 1)
    for (int j = 0; j < g.N; j++) {
        const double *p = base + ((uint32_t)k & g.amask) * UPAD;
        k >>= g.B;
#pragma unroll
        for (int ch = 0; ch < CH; ch++) out[ch] *= p[ch * 32];
        base += pstride;
    }
#pragma unroll
    for (int ch = 0; ch < CH; ch++) {
        const double v = g.pow2 ? out[ch] * g.invP : out[ch] / (double)g.P;
        row_lane[ch * 32] = v;
        row32_lane[ch * 32] = (float)v;  // shadow copy for the float32 screening pass
    }
}

// likelihood.py:45-68 from the cached rows of one state: sum over haplotypes in order, log,
// * count, then the sum over reads as an xor butterfly (uniform result)
template <int CH>
__device__ __noinline__ double eval_rows(const double *q_lane, const double *cnt_lane, int P) {
    constexpr int UPAD = CH * 32;
    double acc = 0.0;
#pragma unroll
    for (int ch = 0; ch < CH; ch++) {
        double rp = 0.0;
#pragma unroll(CH >= 3 ? 4 : 1)
        for (int h = 0; h < P; h++) rp += q_lane[h * UPAD + ch * 32];
        acc += log(rp) * cnt_lane[ch * 32];
    }
    return warp_sum(acc);
}

// Swap one state slot (keys, product rows and their float32 shadows, per-read sums, epochs, memo
// tables) between the resident slot 0 in shared memory and slot `slot` of this warp's backing
// store.  Out of line: only shapes whose tables do not fit use it.
template <int CH>
__device__ __noinline__ void slot_copy(const AsmArgs &a, unsigned char *sm, int lane, int slot, bool store) {
    constexpr int UPAD = CH * 32;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned char *back = a.slot_backing + ((size_t)warp_global * a.tmax + slot) * a.slot_bytes;
    const int32_t offs[7] = {a.o_q, a.o_mcache, a.o_key, a.o_scache, a.o_q32, a.o_rpc, a.o_epoch};
    const int32_t lens[7] = {MCHB_ASM_Q_GLOBAL(CH) ? 0 : a.pmax * UPAD * 8, 2 * a.pmax * a.nmax * 8, a.pmax * 8, a.scache_n * (int)sizeof(ScEntry),
                             a.pmax * UPAD * 4, UPAD * 4, 8};
    __syncwarp();
    size_t boff = 0;
#pragma unroll 1
    for (int r = 0; r < 7; r++) {
        uint32_t *sp = reinterpret_cast<uint32_t *>(sm + offs[r]);
        uint32_t *bp = reinterpret_cast<uint32_t *>(back + boff);
        const int words = lens[r] >> 2;
        if (store) {
#pragma unroll 4
            for (int i = lane; i < words; i += 32) bp[i] = sp[i];
        } else {
#pragma unroll 4
            for (int i = lane; i < words; i += 32) sp[i] = bp[i];
        }
        boff += (size_t)((lens[r] + 7) & ~7);
    }
    __syncwarp();
}

template <int CH, bool PRIOR>
struct AsmCtx {
    static constexpr int UPAD = CH * 32;
    const AsmArgs &a;
    unsigned char *sm;  // this warp's shared memory region
    int lane;
    int N, A, P, B, U;
    uint32_t amask;
    bool pow2;
    double invP;
    uint32_t slots;  // nibble t -> state slot (parallel tempering swaps exchange slots)
    WordStream ws;
    int err;
    long long evals;
    int prof_t = 0;  // index of the temperature being stepped
    int n_acc;       // proposals accepted by the mutation compound step under way
    int qslot = 0;   // state slot behind the resident position 0 (resident-slot swap with q in global memory)

    __device__ __forceinline__ AsmCtx(const AsmArgs &args, unsigned char *s, int l) : a(args), sm(s), lane(l) {}

    __device__ __forceinline__ double *Rt() const { return asm_rt<CH>(a, sm); }
    __device__ __forceinline__ double *cnt() const { return reinterpret_cast<double *>(sm + a.o_cnt); }
    // product rows of the resident state slot(s): shared memory, or (three and more chunks) this
    // warp's region of q_backing, where qslot is the slot the resident position 0 stands for
    __device__ __forceinline__ double *qwarp() const {
        const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        return a.q_backing + (size_t)warp_global * a.q_stride;
    }
    __device__ __forceinline__ double *q() const {
        if (!MCHB_ASM_Q_GLOBAL(CH)) return reinterpret_cast<double *>(sm + a.o_q);
        return qwarp() + (size_t)qslot * a.pmax * UPAD;
    }
    __device__ __forceinline__ double *dist() const { return reinterpret_cast<double *>(sm + a.o_dist); }
    __device__ __forceinline__ double *oll() const { return reinterpret_cast<double *>(sm + a.o_oll); }
    __device__ __forceinline__ double *opr() const { return reinterpret_cast<double *>(sm + a.o_opr); }
    __device__ __forceinline__ double *lgdisp() const { return reinterpret_cast<double *>(sm + a.o_lgdisp); }
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
    // scratch rows for the prior (reuses the interval permutation area + padding: P u64 values)
    __device__ __forceinline__ uint64_t *prow() const { return reinterpret_cast<uint64_t *>(sm + a.o_homlp); }

    __device__ __forceinline__ int slot(int t) const { return (int)((slots >> (4 * t)) & 15u); }
    __device__ __forceinline__ uint64_t *keys(int s) const { return key() + s * P; }
    // keys of state slot s wherever it lives: shared memory, or the backing store when only one
    // slot is resident (every slot is stored right after its temperature was stepped)
    __device__ __forceinline__ const uint64_t *slot_keys(int s) const {
        if (CH < 2 || a.tres >= a.tmax) return keys(s);  // (the CH = 1 kernels never swap: see launch_assemble)
        const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const size_t key_off = (MCHB_ASM_Q_GLOBAL(CH) ? 0 : (size_t)a.pmax * UPAD * 8) + (size_t)2 * a.pmax * a.nmax * 8;  // after q and mcache
        return reinterpret_cast<const uint64_t *>(a.slot_backing + ((size_t)warp_global * a.tmax + s) * a.slot_bytes + key_off);
    }

    // jitutils.random_choice (77-92): p (uniform, shared memory) is overwritten by its cumsum
    __device__ __forceinline__ int random_choice_inplace(double *p, int n) {
        double acc = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            acc += p[i];
            __syncwarp();  // every lane has read p[i] before any lane overwrites it
            p[i] = acc;
        }
        double u = ws.next_double();
        return searchsorted_right(p, n, u);
    }

    __device__ __forceinline__ RowGeom geom() const {
        RowGeom g;
        g.N = N;
        g.A = A;
        g.P = P;
        g.B = B;
        g.amask = amask;
        g.pow2 = pow2 ? 1 : 0;
        g.invP = invP;
        return g;
    }
    __device__ __forceinline__ double *qrow_lane(int s, int h) const { return q() + (size_t)(s * P + h) * UPAD + lane; }
    // spare rows (two per warp) that hold the old rows of a proposal while it is installed
    __device__ __forceinline__ double *spare_lane(int k) const {
        if (MCHB_ASM_Q_GLOBAL(CH)) return qwarp() + (size_t)(a.tmax * a.pmax + k) * UPAD + lane;
        return q() + (size_t)(a.tres * a.pmax + k) * UPAD + lane;
    }

    __device__ __forceinline__ float *q32() const { return reinterpret_cast<float *>(sm + a.o_q32); }
    __device__ __forceinline__ float *rat() const { return reinterpret_cast<float *>(sm + a.o_rat); }
    __device__ __forceinline__ float *c32() const { return reinterpret_cast<float *>(sm + a.o_c32); }
    __device__ __forceinline__ float *qrow32_lane(int s, int h) const { return q32() + (size_t)(s * P + h) * UPAD + lane; }
    __device__ __forceinline__ float *spare32_lane(int k) const { return q32() + (size_t)(a.tres * a.pmax + k) * UPAD + lane; }

    // ---- memo of screened quantities.  Chains sit in one state for long stretches; the screened
    // (float32) acceptance bounds are functions of the state only, so they are kept per slot and
    // dropped whenever the slot's state changes (epoch bump in commit / accept).
    __device__ __forceinline__ uint32_t *epoch() const { return reinterpret_cast<uint32_t *>(sm + a.o_epoch); }  // [T] state, [T] mutation-memo
    __device__ __forceinline__ double *mcache(int s) const {
        return reinterpret_cast<double *>(sm + a.o_mcache) + (size_t)s * 2 * a.pmax * a.nmax;
    }
    __device__ __forceinline__ ScEntry *scache(int s) const {
        return reinterpret_cast<ScEntry *>(sm + a.o_scache) + (size_t)s * a.scache_n;
    }
    __device__ __forceinline__ void bump_epoch(int s) {
        uint32_t *e = epoch();
        __syncwarp();
        if (lane == 0) e[s] = e[s] + 1;
        __syncwarp();
    }

    // float32 per-read probability of the current state of slot s (sum of its shadow rows);
    // refreshed whenever a row of the slot changes for good
    __device__ __forceinline__ float *rpc_lane(int s) const { return reinterpret_cast<float *>(sm + a.o_rpc) + (size_t)s * UPAD + lane; }
    __device__ __forceinline__ void refresh_rpc(int s) {
        const float *qs = q32() + (size_t)(s * P) * UPAD + lane;
        float *dst = rpc_lane(s);
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
            float rp = 0.f;
#pragma unroll(CH >= 3 ? 4 : 1)
            for (int hh = 0; hh < P; hh++) rp += qs[hh * UPAD + ch * 32];
            dst[ch * 32] = rp;
        }
        __syncwarp();
    }

    // log-likelihood of state slot s from its cached product rows
    __device__ __forceinline__ double eval_llk(int s) {
        evals++;
        return eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
    }

    // install the product row of key k as haplotype h of slot s, keeping the old row in spare[k_spare]
    __device__ __forceinline__ void install_row(int s, int h, uint64_t k, int k_spare) {
        double *row = qrow_lane(s, h), *sp = spare_lane(k_spare);
        float *row32 = qrow32_lane(s, h), *sp32 = spare32_lane(k_spare);
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
            sp[ch * 32] = row[ch * 32];
            sp32[ch * 32] = row32[ch * 32];
        }
        compute_row<CH>(Rt() + lane, row, row32, k, geom());
    }
    __device__ __forceinline__ void restore_row(int s, int h, int k_spare) {
        double *row = qrow_lane(s, h), *sp = spare_lane(k_spare);
        float *row32 = qrow32_lane(s, h), *sp32 = spare32_lane(k_spare);
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
            row[ch * 32] = sp[ch * 32];
            row32[ch * 32] = sp32[ch * 32];
        }
    }
    // make key k haplotype h of slot s (row recomputed)
    __device__ __forceinline__ void commit(int s, int h, uint64_t k) {
        keys(s)[h] = k;
        compute_row<CH>(Rt() + lane, qrow_lane(s, h), qrow32_lane(s, h), k, geom());
        refresh_rpc(s);
        bump_epoch(s);
    }

    // prior of the haplotype keys of a slot with up to two haplotypes replaced
    __device__ __forceinline__ double prior_of_keys(const uint64_t *ks, int hA, uint64_t kA, int hB, uint64_t kB) const {
        uint64_t *rows = prow();
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < P; i++) {
            uint64_t v = ks[i];
            v = (i == hA) ? kA : v;
            v = (i == hB) ? kB : v;
            rows[i] = v;
        }
        __syncwarp();
        return assemble_prior_rows(rows, P, sc(), lgdisp());
    }

    // lane-local variant (different lanes evaluate different proposals): haplotype hA replaced by kA
    __device__ __forceinline__ double prior_of_keys_lane(const uint64_t *ks, int hA, uint64_t kA) const {
        const double *scv = sc();
        const double *lgd = lgdisp();
        const bool null_prior = scv[SC_INBREEDING] == 0.0;
        const double lg_disp = scv[SC_LG_DISP];
        double acc = 0.0;
#pragma unroll 1
        for (int i = 0; i < P; i++) {
            const uint64_t ki = (i == hA) ? kA : ks[i];
            int cntv = 0;
            bool first = true;
#pragma unroll 1
            for (int k = 0; k < P; k++) {
                const uint64_t kk = (k == hA) ? kA : ks[k];
                const bool eq = kk == ki;
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

    // prior of a label matrix (structural.py:546: dosage of the (inside, outside) label rows)
    __device__ __forceinline__ double prior_of_labels(uint64_t lin, uint64_t lout) const {
        uint64_t *rows = prow();
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < P; i++) rows[i] = (uint64_t)((nib(lin, i) << 4) | nib(lout, i));
        __syncwarp();
        return assemble_prior_rows(rows, P, sc(), lgdisp());
    }

    // ------------------------------------------------------------------ mutation.py:15-161
    // Serial form of one sub-step (any number of alleles).  u: the step's only uniform draw
    // (random_choice), supplied by the caller.  Bi-allelic sub-steps normally take the
    // lane-parallel path of mutation_compound_step instead.
    __device__ __forceinline__ void base_step(int s, int h, int j, int n_all, double temp, double &llk, const double u) {
        uint64_t *ks = keys(s);
        MCHB_PROF_ADD(prof_t, PK_BASE_STEPS, 1);
        const uint64_t kh = ks[h];
        const int shift = B * j;
        const uint64_t clr = ~((uint64_t)amask << shift);
        const int cur = (int)((uint32_t)(kh >> shift) & amask);
        // lanes 0..P-1 compare one haplotype each (jitutils.count_haplotype_copies 349-374)
        const bool mine = lane < P;
        const uint64_t myk = mine ? ks[lane] : 0ull;
        const double lhap = LOG_INT[__popc(__ballot_sync(MCHB_FULL, mine && myk == kh))];
        double lprior = 0.0;
        if (PRIOR) lprior = prior_of_keys(ks, -1, 0, -1, 0);
        double *ol = oll(), *op = opr();
        int n_options = 0;
#pragma unroll 1
        for (int i = 0; i < n_all; i++) {
            if (i == cur) {
                ol[i] = llk;
                op[i] = -INFINITY;
            } else {
                n_options++;
                const uint64_t kn = (kh & clr) | ((uint64_t)i << shift);
                install_row(s, h, kn, 0);
                const double llk_i = eval_llk(s);
                restore_row(s, h, 0);
                ol[i] = llk_i;
                const double llk_ratio = llk_i - llk;
                double lprior_ratio = 0.0;
                if (PRIOR) lprior_ratio = prior_of_keys(ks, h, kn, -1, 0) - lprior;
                const int copies_n = 1 + __popc(__ballot_sync(MCHB_FULL, mine && lane != h && myk == kn));
                const double lprop = LOG_INT[copies_n] - lhap;
                const double mh = (llk_ratio + lprior_ratio) * temp + lprop;
                op[i] = np_minimum0(mh);
            }
        }
        const double ln_opts = LOG_INT[n_options];
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i < n_all; i++) {
            const double p = dexp(op[i] - ln_opts);
            __syncwarp();  // uniform code updating shared memory in place: read by all lanes, then written
            op[i] = p;
            sum += p;
        }
        __syncwarp();
        op[cur] = 1 - sum;
        __syncwarp();
        double cacc = 0.0;
#pragma unroll 1
        for (int i = 0; i < n_all; i++) {
            cacc += op[i];
            __syncwarp();
            op[i] = cacc;
        }
        __syncwarp();
        const int choice = searchsorted_right(op, n_all, u);
        if (choice >= n_all) {
            err = MCHB_ITEM_CHOICE_RANGE;
            return;
        }
        if (choice != cur) {
            MCHB_PROF_ADD(prof_t, PK_MUT_ACCEPT, 1);
            n_acc++;
            commit(s, h, (kh & clr) | ((uint64_t)choice << shift));
        }
        llk = ol[choice];
    }

    // ------------------------------------------------------------------ mutation.py:165-246
    // The P*N sub-steps are applied in shuffled order.  Each consumes exactly one uniform, so
    // the draw of the i-th pending sub-step sits at a known offset of the word stream; and as
    // long as no proposal is accepted they all start from the same state.  Lanes therefore
    // evaluate up to 32 consecutive bi-allelic sub-steps AT ONCE (lane = sub-step, each lane
    // walks the reads in read order like likelihood.py:45-68); the first accepting lane commits
    // and the sub-steps behind it are re-evaluated from the new state.  A sub-step at a position
    // with more than two alleles is a barrier handled by the serial base_step.
    // ------------------------------------------------------------------ hot mode of mutation.py:165-246
    // The shuffled sub-steps of a replica that accepts most proposals, one after the other (lanes =
    // reads).  Each bi-allelic sub-step is first screened in float32: the proposal's log-likelihood
    // a32 = sum_r c_r logf(rpc_r + q32[h][r] (R_new / R_old - 1)) is within SC_ERR_MUT of the exact one.
    // The exact step accepts iff exp(min(0, mh)) reaches t = u (current allele 1) or t = 1 - u
    // (current allele 0); with d = temp * (error of a32 + error of the current state's value) + slack,
    //   mh32 < log t - d  -> certainly rejected,   mh32 > log t + d  -> certainly accepted,
    // anything else is decided exactly (install the row, eval_rows, the reference's arithmetic).
    // After a certain acceptance the exact log-likelihood of the new state is not needed until a later
    // decision is ambiguous or the compound step ends: until then the screened value a32 of the accepted
    // proposal stands in for it (its error is the same SC_ERR_MUT, doubling d) and the exact value is
    // evaluated from the rows when it is asked for — a deterministic function of the state, the
    // value the reference carries.  Decisions, draws and evaluation counts are the reference's.
    __device__ __forceinline__ void hot_compound_step(int s, int t, double temp, double &llk, int n) {
        const uint16_t *pm = perm();
        const uint8_t *na = nall();
        uint64_t *ks = keys(s);
        const double e_mut = sc()[SC_ERR_MUT];
        bool known = true;   // llk is the exact log-likelihood of the current state
        double l32 = 0.0;    // otherwise: the screened value of the proposal that was accepted last
#pragma unroll 1
        for (int i = 0; i < n && !err; i++) {
            const int hj = pm[i];
            const int h = hj >> 8, j = hj & 255;
            const uint64_t kh = ks[h];
            const int shift = B * j;
            const int cur = (int)((uint32_t)(kh >> shift) & amask);
            const double u = ws.next_double();
            if (na[j] != 2 || cur >= 2) {
                if (!known) {
                    llk = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
                    known = true;
                }
                base_step(s, h, j, na[j], temp, llk, u);
                continue;
            }
            evals++;
            const uint64_t kn = (kh & ~((uint64_t)amask << shift)) | ((uint64_t)(cur ^ 1) << shift);
            // lanes 0..P-1 compare one haplotype each (jitutils.count_haplotype_copies 349-374)
            const bool hl = lane < P;
            const uint64_t myk = hl ? ks[lane] : 0ull;
            const int copies_o = __popc(__ballot_sync(MCHB_FULL, hl && myk == kh));
            const int copies_n = 1 + __popc(__ballot_sync(MCHB_FULL, hl && lane != h && myk == kn));
            const double lprop = LOG_INT[copies_n] - LOG_INT[copies_o];
            double lprior_ratio = 0.0;
            if (PRIOR) lprior_ratio = prior_of_keys(ks, h, kn, -1, 0) - prior_of_keys(ks, -1, 0, -1, 0);
            // ---- float32 screen of this proposal
            float acc = 0.f;
            bool ok = true;
            {
                const float *qh = q32() + (size_t)(s * P + h) * UPAD + lane;
                const float *rc = reinterpret_cast<const float *>(sm + a.o_rpc) + (size_t)s * UPAD + lane;
                const float *rt = rat() + (size_t)(MCHB_ASM_RAT_HALF(CH) ? j : j * 2 + (cur & 1)) * UPAD + lane;
                const float *cw = c32() + lane;
#pragma unroll
                for (int ch = 0; ch < CH; ch++) {
                    const float rc_r = rc[ch * 32];
                    float rt_r = rt[ch * 32];
                    if (MCHB_ASM_RAT_HALF(CH) && (cur & 1)) rt_r = __frcp_rn(rt_r);
                    const float rp = fmaf(qh[ch * 32], rt_r - 1.0f, rc_r);
                    ok = ok && (rp > MCHB_SCREEN_MIN_RATIO * rc_r) && (rp > 1e-30f) && (rp < 1e30f);
                    acc = fmaf(__logf(rp), cw[ch * 32], acc);
                }
            }
            const double a32 = warp_sum((double)acc);
            const bool sane = __all_sync(MCHB_FULL, ok);
            const double mh32 = ((a32 - (known ? llk : l32)) + lprior_ratio) * temp + lprop;
            const double t_acc = (cur == 1) ? u : 1.0 - u;
            const double thr = (double)__logf((float)t_acc);
            const double dlt = temp * (known ? e_mut : 2.0 * e_mut) + MCHB_SCREEN_SLACK;
            // (u away from 0 and 1: log t stays inside __logf's range and the cs1 <= u corner is excluded)
            const bool decidable = sane && u > 8.8817841970012523e-16 && u < 0.99999999999999911182;
            if (decidable && mh32 < thr - dlt) continue;  // certainly rejected
            if (decidable && mh32 > thr + dlt) {          // certainly accepted
                MCHB_PROF_ADD(prof_t, PK_MUT_ACCEPT, 1);
                n_acc++;
                commit(s, h, kn);
                l32 = a32;
                known = false;
                continue;
            }
            // ---- decided exactly
            if (!known) {
                llk = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
                known = true;
            }
            MCHB_PROF_ADD(prof_t, PK_T2A_EVALS, 1);
            install_row(s, h, kn, 0);
            const double llk_x = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
            const double mh = ((llk_x - llk) + lprior_ratio) * temp + lprop;
            const double la = np_minimum0(mh);
            int choice;
            if (la < -40.0 && u >= 1.1102230246251565e-16) {
                choice = cur;
            } else {
                const double p_o = dexp(la - 0.0);
                const double p_c = 1 - p_o;  // 1 - (0 + p_o)
                const double cs0 = cur == 0 ? p_c : p_o;
                const double cs1 = cs0 + (cur == 0 ? p_o : p_c);
                choice = (cs1 <= u) ? 2 : ((cs0 <= u) ? 1 : 0);
            }
            if (choice == cur) {
                restore_row(s, h, 0);
            } else if (choice >= 2) {
                restore_row(s, h, 0);
                err = MCHB_ITEM_CHOICE_RANGE;
            } else {
                __syncwarp();
                MCHB_PROF_ADD(prof_t, PK_MUT_ACCEPT, 1);
                n_acc++;
                ks[h] = kn;  // the row is already installed
                llk = llk_x;
                refresh_rpc(s);
                bump_epoch(s);
            }
        }
        if (!known && !err) llk = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
        __syncwarp();
        if (lane == 0) hot()[4 * t] = n_acc;
        __syncwarp();
    }

    // Per-temperature history that steers the choice between equivalent code paths (never the
    // results): [4 t] accepted proposals of the previous mutation compound step at temperature t.
    __device__ __forceinline__ int32_t *hot() const { return reinterpret_cast<int32_t *>(sm + a.o_hot); }

    __device__ __forceinline__ void mutation_compound_step(int s, int t, double temp, double &llk) {
        const int n = P * N;
        // A replica that accepted many proposals in its last compound step (a heated one, or a
        // chain still far from a mode) gains nothing from screening whole windows of sub-steps from
        // one state: every acceptance invalidates the window behind it.  Such steps decide their
        // sub-steps one after the other, exactly, with no screening pass (tier 2a below).
        const bool hot_mode = hot()[4 * t] * 4 >= n;
        n_acc = 0;
        uint16_t *pm = perm();
        __syncwarp();
        if (n <= 32) {
            // np.random.shuffle of the rows: for i = n-1 .. 1 swap rows i and randint(i + 1).
            // (1) The draws.  Lane l looks at word cursor + l.  Which draw a word serves depends
            // on how many words before it were rejected, so the rejection flags are iterated to
            // their fixed point: a consistent assignment is unique (induction over the words), it
            // is the serial rejection loop's, and word l is right after l + 1 rounds at the latest
            // (in practice 2-4 rounds per 32 words).  Accepted draws are scattered to pm[i].
            int i0 = n - 1;
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
            while (i0 > 0) {
                const uint32_t w = ws.word_at(lane);
                unsigned b = 0, bp;
                int ik;
                uint32_t v;
                bool act;
                do {
                    bp = b;
                    ik = i0 - lane + __popc(bp & lt);  // the draw this word would serve: randint(ik + 1)
                    act = ik >= 1;
                    v = w & (0xffffffffu >> __clz(act ? ik : 1));
                    b = __ballot_sync(MCHB_FULL, act && (int)v > ik);
                } while (b != bp);
                const bool acc = act && !((b >> lane) & 1u);
                if (acc) pm[ik] = (uint16_t)v;
                const unsigned last = __ballot_sync(MCHB_FULL, acc && ik == 1);
                const int used = last ? __ffs(last) : 32;
                i0 = last ? 0 : i0 - (32 - __popc(b));
                ws.advance(used);
            }
            __syncwarp();
            // (2) The swaps.  Row i is final after swap i; what it receives is the content of row
            // k_i at that time, traced back through the earlier swaps (i' > i) that moved something
            // into that row: c = k_i; for i' = i + 1 .. n - 1: if (k_i' == c) c = i'.  With
            // inv[x] = {i' : k_i' == x} the first hop is the lowest member of inv[k_i] above i and
            // every later hop is succ(c) = the lowest member of inv[c] above c.  Lane = final row.
            const int kreg = (lane >= 1 && lane < n) ? (int)pm[lane] : 0;
            uint32_t *inv = reinterpret_cast<uint32_t *>(sm + a.o_inv);  // 32 words
            inv[lane] = 0u;
            __syncwarp();
            if (lane >= 1 && lane < n) atomicOr(inv + kreg, 1u << lane);
            __syncwarp();
            const uint32_t above = 0xfffffffeu << lane;  // lanes above mine
            const uint32_t mine_in = inv[lane] & above;
            const int succ = mine_in ? __ffs(mine_in) - 1 : -1;
            int c = kreg;
            {
                const uint32_t first = __shfl_sync(MCHB_FULL, inv[lane], kreg) & above;
                int nxt = first ? __ffs(first) - 1 : -1;
#pragma unroll 1
                while (__any_sync(MCHB_FULL, nxt >= 0)) {
                    if (nxt >= 0) c = nxt;
                    const int s2 = __shfl_sync(MCHB_FULL, succ, c);
                    nxt = nxt >= 0 ? s2 : -1;
                }
            }
            __syncwarp();
            if (lane < n) {
                // c / N for c < 32: (c + 0.5) / N is at least 1 / (2 N) away from an integer
                const int h = __float2int_rz(((float)c + 0.5f) * __frcp_rn((float)N));
                pm[lane] = (uint16_t)((h << 8) | (c - h * N));
            }
        } else {
            for (int i = lane; i < n; i += 32) {
                int h = i / N, j = i - h * N;
                pm[i] = (uint16_t)((h << 8) | j);
            }
            __syncwarp();
#pragma unroll 1
            for (int i = n - 1; i > 0; i--) {  // np.random.shuffle of the rows
                int k = ws.randint(i + 1);
                uint16_t x = pm[i], y = pm[k];
                __syncwarp();
                pm[i] = y;
                pm[k] = x;
            }
        }
        __syncwarp();
        const uint8_t *na = nall();
        uint64_t *ks = keys(s);
        if (hot_mode) {
            hot_compound_step(s, t, temp, llk, n);
            return;
        }
        const double *Rt0 = Rt();
        const double *q0 = q() + (size_t)(s * P) * UPAD;
        const double *cn = cnt();
        int done = 0;
        const uint32_t epoch_at_start = epoch()[s];
#pragma unroll 1
        while (done < n && !err) {
            // screened quantities of this slot's sub-steps are still valid if the state is the
            // one they were computed for
            const bool memo_ok = epoch()[a.tres + s] == epoch()[s];
            const int i = done + lane;
            const bool active = i < n;
            const int hj = pm[active ? i : done];
            const int h = hj >> 8, j = hj & 255;
            const uint64_t kh = ks[h];
            const int shift = B * j;
            const int cur = (int)((uint32_t)(kh >> shift) & amask);
            const bool fast = na[j] == 2 && cur < 2;
            const unsigned slow_mask = __ballot_sync(MCHB_FULL, active && !fast);
            int limit = min(32, n - done);
            if (slow_mask) limit = min(limit, __ffs(slow_mask) - 1);
            if (limit == 0) {
                // the next sub-step is not bi-allelic: serial path (uniform: all lanes agree)
                const int hj0 = pm[done];
                const double u0 = ws.next_double();
                base_step(s, hj0 >> 8, hj0 & 255, na[hj0 & 255], temp, llk, u0);
                done += 1;
                continue;
            }
            const bool mine = lane < limit;
            MCHB_PROF_ADD(prof_t, PK_WINDOWS, 1);
            const uint64_t kn = (kh & ~((uint64_t)amask << shift)) | ((uint64_t)(cur ^ 1) << shift);
            const double u = ws.double_at(2 * (mine ? lane : 0));
            double lprior_ratio, lprop, d32;
            double *mc = mcache(s) + 2 * (h * N + j);
            if (memo_ok) {
                // the state has not changed since these were computed
                d32 = mc[0];
                lprop = mc[1];
                lprior_ratio = 0.0;  // folded into d32
            } else {
                // ---- the parts of my Metropolis-Hastings ratio that do not need the likelihood
                int copies_o = 0, copies_n = 1;
#pragma unroll 1
                for (int k = 0; k < P; k++) {
                    const uint64_t kk = ks[k];
                    copies_o += (kk == kh);
                    copies_n += (k != h && kk == kn);
                }
                lprior_ratio = 0.0;
                if (PRIOR) lprior_ratio = prior_of_keys_lane(ks, h, kn) - prior_of_keys_lane(ks, -1, 0);
                lprop = LOG_INT[copies_n] - LOG_INT[copies_o];
                // ---- tier 1: float32 screening.  The proposal's row is the cached row of haplotype
                // h times R[j][new] / R[j][old] (exact in real arithmetic); with float32 roundings and
                // __logf the log-likelihood stays within SC_ERR_MUT of the exact one.  The exact step
                // accepts only if exp(min(0, mh)) reaches t = u (current allele 1) or t = 1 - u
                // (current allele 0) — see the cumulative sums in base_step — so a sub-step whose
                // screened mh is below log(t) - (temp * SC_ERR_MUT + slack) is certainly rejected and needs no exact
                // evaluation; everything else ("needy") is decided exactly below.
                float a32f = 0.f;  // float32 accumulation: its rounding is part of SC_ERR_MUT
                bool sane = true;
                {
                    // rp_new = rp_cur + q[h] * (R_new / R_old - 1): one fused multiply-add per read.
                    // The subtraction hidden in it can lose relative accuracy when the proposal
                    // removes almost all of a read's probability, so reads with rp_new < 1e-4 rp_cur
                    // make the sub-step needy; above that the relative error stays below 1e-3.
                    const float *qh = q32() + (size_t)(s * P + h) * UPAD;
                    const float *rc = reinterpret_cast<const float *>(sm + a.o_rpc) + (size_t)s * UPAD;
                    const float *rt = rat() + (size_t)(MCHB_ASM_RAT_HALF(CH) ? j : j * 2 + (cur & 1)) * UPAD;
                    const float *cw = c32();
#pragma unroll 2
                    for (int r = 0; r < U; r++) {
                        const float rc_r = rc[r];
                        float rt_r = rt[r];
                        if (MCHB_ASM_RAT_HALF(CH) && (cur & 1)) rt_r = __frcp_rn(rt_r);
                        const float rp = fmaf(qh[r], rt_r - 1.0f, rc_r);
                        sane = sane && (rp > MCHB_SCREEN_MIN_RATIO * rc_r) && (rp > 1e-30f) && (rp < 1e30f);
                        a32f = fmaf(__logf(rp), cw[r], a32f);
                    }
                }
                const double a32 = (double)a32f;
                d32 = sane ? (a32 - llk) + lprior_ratio : INFINITY;  // +inf: never screened out
                if (mine) {
                    mc[0] = d32;
                    mc[1] = lprop;
                }
            }
            const double mh32 = d32 * temp + lprop;
            const double t_acc = (cur == 1) ? u : 1.0 - u;
            const bool hopeless = (mh32 < (double)__logf((float)t_acc) - (temp * sc()[SC_ERR_MUT] + MCHB_SCREEN_SLACK)) &&
                                  (u < 0.99999999999999911182);  // 1 - 2^-50: keep clear of the cs1 <= u corner
            const unsigned needy = __ballot_sync(MCHB_FULL, mine && !hopeless);
            if (PRIOR && memo_ok && needy != 0)
                lprior_ratio = prior_of_keys_lane(ks, h, kn) - prior_of_keys_lane(ks, -1, 0);
            // (the kernels of larger items always take the serial tier: their Rt lives in global memory
            // and the lane-per-sub-step loop of tier 2b would read it with 32 different rows per load)
            if (CH >= 2 || __popc(needy) <= MCHB_EXACT_SERIAL_MAX) {
                // ---- tier 2a: exact decisions for the needy sub-steps, in order (uniform code)
                int completed = limit;
                unsigned m = needy;
#pragma unroll 1
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    const int hl = __shfl_sync(MCHB_FULL, h, l);
                    const int curl = __shfl_sync(MCHB_FULL, cur, l);
                    const uint64_t knl = __shfl_sync(MCHB_FULL, kn, l);
                    const double ul = __shfl_sync(MCHB_FULL, u, l);
                    const double lrest = __shfl_sync(MCHB_FULL, lprior_ratio, l);
                    const double lpropl = __shfl_sync(MCHB_FULL, lprop, l);
                    install_row(s, hl, knl, 0);
                    MCHB_PROF_ADD(prof_t, PK_T2A_EVALS, 1);
                    const double llk_x = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
                    const double mh = ((llk_x - llk) + lrest) * temp + lpropl;
                    const double la = np_minimum0(mh);
                    int choice;
                    if (la < -40.0 && ul >= 1.1102230246251565e-16) {
                        choice = curl;
                    } else {
                        const double p_o = dexp(la - 0.0);
                        const double p_c = 1 - p_o;  // 1 - (0 + p_o)
                        const double cs0 = curl == 0 ? p_c : p_o;
                        const double cs1 = cs0 + (curl == 0 ? p_o : p_c);
                        choice = (cs1 <= ul) ? 2 : ((cs0 <= ul) ? 1 : 0);
                    }
                    if (choice == curl) {
                        restore_row(s, hl, 0);
                        continue;
                    }
                    completed = l + 1;
                    if (choice >= 2) {
                        restore_row(s, hl, 0);
                        err = MCHB_ITEM_CHOICE_RANGE;
                    } else {
                        __syncwarp();
                        MCHB_PROF_ADD(prof_t, PK_MUT_ACCEPT, 1);
                        n_acc++;
                        ks[hl] = knl;  // the row is already installed
                        llk = llk_x;
                        refresh_rpc(s);
                        bump_epoch(s);
                    }
                    break;
                }
                evals += completed;
                ws.advance(2 * completed);
                done += completed;
                continue;
            }
            // ---- tier 2b: many needy sub-steps: exact log-likelihood of every proposal of the
            // window, lane-parallel (reads in order, haplotypes in order)
            double llk_o = 0.0;
            MCHB_PROF_ADD(prof_t, PK_T2B_WINDOWS, 1);
            {
                const int astride = UPAD, pstride = A * UPAD;
#pragma unroll 1
                for (int r = 0; r < U; r++) {
                    double prod = 1.0;
                    const double *col = Rt0 + r;
                    uint64_t kk = kn;
#pragma unroll 2
                    for (int jj = 0; jj < N; jj++) {
                        prod *= col[((uint32_t)kk & amask) * astride];
                        kk >>= B;
                        col += pstride;
                    }
                    prod = pow2 ? prod * invP : prod / (double)P;
                    double rp = 0.0;
#pragma unroll 2
                    for (int hh = 0; hh < P; hh++) {
                        const double v = q0[hh * UPAD + r];
                        rp += (hh == h) ? prod : v;
                    }
                    llk_o += log(rp) * cn[r];
                }
            }
            const double mh = ((llk_o - llk) + lprior_ratio) * temp + lprop;
            const double la = np_minimum0(mh);
            int choice;
            if (la < -40.0 && u >= 1.1102230246251565e-16) {
                choice = cur;  // exp(la) < 2^-54: see base_step
            } else {
                const double p_o = dexp(la - 0.0);
                const double p_c = 1 - p_o;
                const double cs0 = cur == 0 ? p_c : p_o;
                const double cs1 = cs0 + (cur == 0 ? p_o : p_c);
                choice = (cs1 <= u) ? 2 : ((cs0 <= u) ? 1 : 0);
            }
            const unsigned event = __ballot_sync(MCHB_FULL, mine && choice != cur);
            if (event == 0) {
                // every evaluated proposal was rejected
                MCHB_PROF_ADD(prof_t, PK_T2B_DONE, limit);
                evals += limit;
                ws.advance(2 * limit);
                done += limit;
                continue;
            }
            const int first = __ffs(event) - 1;
            MCHB_PROF_ADD(prof_t, PK_T2B_DONE, first + 1);
            MCHB_PROF_ADD(prof_t, PK_MUT_ACCEPT, 1);
            n_acc++;
            evals += first + 1;
            ws.advance(2 * (first + 1));
            done += first + 1;
            const int ch1 = __shfl_sync(MCHB_FULL, choice, first);
            if (ch1 >= 2) {
                err = MCHB_ITEM_CHOICE_RANGE;
                break;
            }
            // commit the first accepted proposal (uniform code: products across lanes = reads)
            const int h1 = __shfl_sync(MCHB_FULL, h, first);
            const uint64_t kn1 = __shfl_sync(MCHB_FULL, kn, first);
            llk = __shfl_sync(MCHB_FULL, llk_o, first);
            __syncwarp();
            commit(s, h1, kn1);
            __syncwarp();
        }
        // every bi-allelic sub-step was screened from one and the same state: keep the memo
        __syncwarp();
        if (lane == 0) {
            if (!err && epoch()[s] == epoch_at_start) epoch()[a.tres + s] = epoch_at_start;
            hot()[4 * t] = n_acc;
        }
        __syncwarp();
    }

    // ------------------------------------------------------------------ structural.py:434-587
    __device__ __forceinline__ void interval_step(int s, int start, int stop, int step_type, double temp,
                                                  double &llk) {
        uint64_t *ks = keys(s);
        MCHB_PROF_ADD(prof_t, PK_INTERVALS, 1);
        // ---- memo of this (type, interval) for the slot's current state
        const uint32_t ep = epoch()[s];
        const uint32_t skey = ((uint32_t)step_type << 16) | ((uint32_t)start << 8) | (uint32_t)stop;
        ScEntry *ent = scache(s) + (a.scache_tri > 0 ? step_type * a.scache_tri + ((stop * (stop - 1)) >> 1) + start
                                                      : (int)((skey * 2654435761u) >> (32 - MCHB_SCACHE_HASH_LOG2(CH))));
        double u = 0.0;
        bool have_u = false;
        if (ent->epoch == ep && ent->key == skey && ent->temp == (float)temp) {
            const int n_opt = ent->n_options;
            if (n_opt == 0) return;  // no option, no draw
            u = ws.next_double();
            have_u = true;
            if (u > 0.0 && u < 0.99999999999999911182 &&
                (double)ent->smax < (double)__logf((float)u) - (temp * sc()[SC_ERR_STR] + MCHB_SCREEN_SLACK)) {
                evals += n_opt;  // certainly "stay" (see the screening below)
                MCHB_PROF_ADD(prof_t, PK_STR_MEMO_STAY, 1);
                return;
            }
        }
        const int width = B * (stop - start);
        uint64_t mask_in = 0;
        if (width >= 64) mask_in = ~0ull;
        else if (width > 0) mask_in = ((1ull << width) - 1ull) << (B * start);
        const ulonglong2 labels = segment_labels(ks, P, mask_in);
        const uint64_t lin = labels.x, lout = labels.y;
        uint8_t *o0 = opt0(), *o1 = opt1();
        __syncwarp();
        const int n_options = structural_options(lin, lout, P, step_type, o0, o1);
        __syncwarp();
        if (n_options == 0) {  // no draw (structural.py:504-506)
            if (lane == 0) {
                ent->epoch = ep;
                ent->key = skey;
                ent->smax = INFINITY;
                ent->temp = (float)temp;
                ent->n_options = 0;
            }
            __syncwarp();
            return;
        }
        if (!have_u) u = ws.next_double();  // the step's only draw (random_choice at the end)
        const double log_proposal = LOG_INV_INT[n_options];
        double lprior = 0.0;
        if (PRIOR) lprior = prior_of_keys(ks, -1, 0, -1, 0);
        // lane i owns option i: its label matrix and its number of reverse moves are evaluated by
        // all lanes at once (the option functions are pure integer code)
        uint64_t my_lin = lin;
        if (lane < n_options) {
            const int h0 = o0[lane], h1 = o1[lane];
            my_lin = nib_set(lin, h0, nib(lin, h1));
            if (step_type == 0) my_lin = nib_set(my_lin, h1, nib(lin, h0));
        }
        // ---- float32 screening (all variable positions bi-allelic, <= 32 options): the step
        // stays put unless sum_i exp(min(0, mh_i)) / n exceeds u; with every screened mh_i below
        // log(u) - (temp * SC_ERR_STR + slack) that sum is below u, so "stay" is certain and no exact
        // log-likelihood is needed (same error budget as in mutation_compound_step).  The largest
        // screened mh is a function of the state only and is kept in the memo.
        {
            double smax = INFINITY;
            // a replica whose last 16 screening passes all failed to prove "stay" (a heated one)
            // skips the pass and probes again every 16th interval
            int32_t *fails = hot() + 4 * prof_t + 1;
            const int nf = *fails;
            const bool try_screen = nf < 16 || (nf & 15) == 0;
            __syncwarp();
            if (lane == 0) *fails = nf + 1;
            double my_d32 = 0.0;   // lane i: screened (llk_i - llk) + prior ratio of option i; NaN: not usable
            bool screened = false;
            if (B == 1 && n_options <= 32 && try_screen) {
                const float *qs = q32() + (size_t)(s * P) * UPAD + lane;
                const float *rt = rat() + lane;
                const float *cw = c32() + lane;
                smax = -INFINITY;
                screened = true;
#pragma unroll 1
                for (int k = 0; k < n_options; k++) {
                    const int h0 = o0[k], h1 = o1[k];
                    const uint64_t k0 = ks[h0], k1 = ks[h1];
                    float ra[CH], rb[CH];
#pragma unroll
                    for (int ch = 0; ch < CH; ch++) {
                        ra[ch] = qs[h0 * UPAD + ch * 32];
                        rb[ch] = qs[h1 * UPAD + ch * 32];
                    }
                    uint64_t d = (k0 ^ k1) & mask_in;  // positions where the swapped / copied segment differs
#pragma unroll 1
                    while (d) {
                        const int jp = __ffsll((long long)d) - 1;
                        d &= d - 1;
                        const int c0 = (int)((k0 >> jp) & 1ull);
#pragma unroll
                        for (int ch = 0; ch < CH; ch++) {
                            if (MCHB_ASM_RAT_HALF(CH)) {
                                const float v01 = rt[jp * UPAD + ch * 32], v10 = __frcp_rn(v01);
                                ra[ch] *= c0 ? v10 : v01;  // h0 takes h1's allele
                                rb[ch] *= c0 ? v01 : v10;  // h1 takes h0's allele
                            } else {
                                ra[ch] *= rt[(jp * 2 + c0) * UPAD + ch * 32];        // h0 takes h1's allele
                                rb[ch] *= rt[(jp * 2 + (c0 ^ 1)) * UPAD + ch * 32];  // h1 takes h0's allele
                            }
                        }
                    }
                    float acc = 0.f;
                    bool ok = true;
#pragma unroll
                    for (int ch = 0; ch < CH; ch++) {
                        float rp = 0.f;
#pragma unroll(CH >= 3 ? 4 : 1)
                        for (int hh = 0; hh < P; hh++) {
                            float v = qs[hh * UPAD + ch * 32];
                            v = (hh == h0) ? ra[ch] : v;
                            v = (step_type == 0 && hh == h1) ? rb[ch] : v;
                            rp += v;
                        }
                        ok = ok && (rp > 1e-30f) && (rp < 1e30f);
                        acc += __logf(rp) * cw[ch * 32];
                    }
                    const double a32 = warp_sum((double)acc);
                    const bool sane = __all_sync(MCHB_FULL, ok);
                    double lprior_ratio = 0.0;
                    if (PRIOR) lprior_ratio = prior_of_labels(__shfl_sync(MCHB_FULL, my_lin, k), lout) - lprior;
                    // the proposal ratio is log(1 / n_reverse) - log(1 / n_options) <= log(n_options): the
                    // bound spares the screening pass the reverse-move count of every option
                    const double lprop = -log_proposal;
                    const double d32 = (a32 - llk) + lprior_ratio;
                    const double mh32 = d32 * temp + lprop;
                    smax = sane ? fmax(smax, mh32) : INFINITY;  // fmax ignores NaN: treat NaN as inconclusive
                    if (isnan(mh32)) smax = INFINITY;
                    if (lane == k) my_d32 = sane ? d32 : NAN;
                }
            }
            __syncwarp();
            if (lane == 0) {
                ent->epoch = ep;
                ent->key = skey;
                ent->smax = __double2float_ru(smax);
                ent->temp = (float)temp;
                ent->n_options = n_options;
            }
            __syncwarp();
            if (u > 0.0 && u < 0.99999999999999911182 &&
                smax < (double)__logf((float)u) - (temp * sc()[SC_ERR_STR] + MCHB_SCREEN_SLACK)) {
                evals += n_options;
                MCHB_PROF_ADD(prof_t, PK_STR_SCREEN_STAY, 1);
                if (lane == 0) *fails = 0;
                __syncwarp();
                return;
            }
            // ---- the categorical draw by interval arithmetic.  The step picks the first k whose
            // cumulative sum c_k = sum_{i<=k} exp(min(0, mh_i)) / n exceeds u ("stay" has the rest of
            // the mass).  Every mh_i is known to within d = temp * SC_ERR_STR + slack of its screened
            // value (here with the option's true proposal ratio), so c_k lies between the cumulative
            // sums of exp(min(0, mh32_i -+ d)) / n: when the upper sum up to k - 1 is below u and the
            // lower sum up to k is above u the choice is k whatever the exact values are — only the
            // chosen option's log-likelihood (the new state's) is then evaluated exactly; when the
            // upper sum of all options is below u the step stays.  Anything else is decided by the
            // exact path below.  (Heated replicas: almost every structural step ends here.)
            if (screened && u > 0.0 && u < 0.99999999999999911182) {
                const int n_ret = structural_options(my_lin, lout, P, step_type, nullptr, nullptr);
                const bool mine = lane < n_options;
                const double dlt = temp * sc()[SC_ERR_STR] + MCHB_SCREEN_SLACK;
                const double mh32 = my_d32 * temp + (LOG_INV_INT[mine ? n_ret : 1] - log_proposal);
                const bool usable = mine && !isnan(mh32);
                const double ln_opts = LOG_INT[n_options];
                double lo = 0.0, hi = 0.0;
                if (mine) {
                    hi = dexp(-ln_opts);  // unusable option: anything up to 1 / n
                    if (usable) {
                        lo = dexp(fmin(0.0, mh32 - dlt) - ln_opts);
                        hi = dexp(fmin(0.0, mh32 + dlt) - ln_opts);
                    }
                }
                double slo = lo, shi = hi;  // inclusive prefix sums over the options
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double a_ = __shfl_up_sync(MCHB_FULL, slo, o), b_ = __shfl_up_sync(MCHB_FULL, shi, o);
                    if (lane >= o) {
                        slo += a_;
                        shi += b_;
                    }
                }
                const double eps = 1e-12;  // roundings of the reference's own cumulative sum
                const bool is_k = mine && (shi - hi) + eps < u && slo - eps > u;
                const unsigned pick = __ballot_sync(MCHB_FULL, is_k);
                const double shi_all = __shfl_sync(MCHB_FULL, shi, n_options - 1);
                if (pick != 0 || shi_all + eps < u) {
                    evals += n_options;
                    if (lane == 0) *fails = 0;
                    __syncwarp();
                    if (pick != 0) {
                        const int choice = __ffs(pick) - 1;
                        MCHB_PROF_ADD(prof_t, PK_STR_ACCEPT, 1);
                        const int h0 = o0[choice], h1 = o1[choice];
                        const uint64_t k0 = ks[h0], k1 = ks[h1];
                        if (step_type == 0) commit(s, h1, (k0 & mask_in) | (k1 & ~mask_in));
                        commit(s, h0, (k1 & mask_in) | (k0 & ~mask_in));
                        llk = eval_rows<CH>(q() + (size_t)(s * P) * UPAD + lane, cnt() + lane, P);
                    } else {
                        MCHB_PROF_ADD(prof_t, PK_STR_SCREEN_STAY, 1);
                    }
                    return;
                }
            }
        }
        double my_la = -INFINITY;
#pragma unroll 1
        for (int base = 0; base < n_options; base += 32) {
            // (more than 32 options only for ploidy > 6: processed in rounds of 32)
            uint64_t lin_r = my_lin;
            if (base > 0) {
                lin_r = lin;
                if (base + lane < n_options) {
                    const int h0 = o0[base + lane], h1 = o1[base + lane];
                    lin_r = nib_set(lin, h0, nib(lin, h1));
                    if (step_type == 0) lin_r = nib_set(lin_r, h1, nib(lin, h0));
                }
            }
            const int n_ret = structural_options(lin_r, lout, P, step_type, nullptr, nullptr);
            const int rounds = min(32, n_options - base);
#pragma unroll 1
            for (int k = 0; k < rounds; k++) {
                const int i = base + k;
                const int h0 = o0[i], h1 = o1[i];
                const uint64_t k0 = ks[h0], k1 = ks[h1];
                const uint64_t k0n = (k1 & mask_in) | (k0 & ~mask_in);
                const uint64_t lin_o = __shfl_sync(MCHB_FULL, lin_r, k);
                const int n_return = __shfl_sync(MCHB_FULL, n_ret, k);
                install_row(s, h0, k0n, 0);
                if (step_type == 0) install_row(s, h1, (k0 & mask_in) | (k1 & ~mask_in), 1);
                const double llk_i = eval_llk(s);
                MCHB_PROF_ADD(prof_t, PK_STR_EXACT, 1);
                restore_row(s, h0, 0);
                if (step_type == 0) restore_row(s, h1, 1);
                double lprior_ratio = 0.0;
                if (PRIOR) lprior_ratio = prior_of_labels(lin_o, lout) - lprior;
                const double lprop = LOG_INV_INT[n_return] - log_proposal;
                const double mh = ((llk_i - llk) + lprior_ratio) * temp + lprop;
                const double la = np_minimum0(mh);
                if (i < 32) {
                    if (lane == i) {
                        my_la = la;
                    }
                }
                oll()[i] = llk_i;
                opr()[i] = la;
            }
        }
        double *ol = oll(), *op = opr();
        // all proposals hopeless (exp(la) < 2^-54 each) and u > 0: the cumulative sums stay below
        // u until the final "stay" slot, so the outcome is "stay" without evaluating exp()
        bool hopeless = u >= 1.1102230246251565e-16;
        if (n_options <= 32) {
            hopeless = hopeless && __all_sync(MCHB_FULL, lane >= n_options || my_la < -40.0);
        } else {
#pragma unroll 1
            for (int i = 0; i < n_options; i++) hopeless = hopeless && (op[i] < -40.0);
        }
        if (hopeless) return;
        ol[n_options] = -INFINITY;
        op[n_options] = -INFINITY;
        const double ln_opts = LOG_INT[n_options];
        double sum = 0.0;
#pragma unroll 1
        for (int i = 0; i <= n_options; i++) {
            const double p = dexp(op[i] - ln_opts);
            __syncwarp();  // uniform code updating shared memory in place: read by all lanes, then written
            op[i] = p;
            sum += p;
        }
        __syncwarp();
        op[n_options] = 1 - sum;
        __syncwarp();
        double acc = 0.0;
#pragma unroll 1
        for (int i = 0; i <= n_options; i++) {
            acc += op[i];
            __syncwarp();
            op[i] = acc;
        }
        __syncwarp();
        const int choice = searchsorted_right(op, n_options + 1, u);
        if (choice < n_options) {
            MCHB_PROF_ADD(prof_t, PK_STR_ACCEPT, 1);
            const int h0 = o0[choice], h1 = o1[choice];
            const uint64_t k0 = ks[h0], k1 = ks[h1];
            const uint64_t k0n = (k1 & mask_in) | (k0 & ~mask_in);
            if (step_type == 0) commit(s, h1, (k0 & mask_in) | (k1 & ~mask_in));
            commit(s, h0, k0n);
            llk = ol[choice];
        }
    }

    // the three structural sub-steps of mcmc.py:347-394 share this single call site of
    // interval_step: sub 0 = recombination over random intervals, sub 1 = dosage swap over
    // random intervals (structural.py:23-71 random_breaks + 591-673 compound_step),
    // sub 2 = dosage swap over the full length
    __device__ __forceinline__ void structural_substeps(int s, double temp, double &llk, const double *brow, int blen) {
        uint8_t *vb = ivb(), *vp = ivp();
#pragma unroll 1
        for (int sub = 0; sub < 3 && !err; sub++) {
            const double p_sub = sub == 0 ? a.p_recomb : (sub == 1 ? a.p_partial : a.p_dosage);
            if (!(ws.next_double() <= p_sub)) continue;
            int n_int = 1;
            __syncwarp();
            if (sub < 2) {
                // random_choice(break_dist): cumulative sums prepared once per item (bcs)
                const double ub = ws.next_double();
                const int n_breaks = searchsorted_right(reinterpret_cast<const double *>(sm + a.o_bcs), blen, ub);
                if (n_breaks >= N) {
                    err = MCHB_ITEM_BREAKS;
                    return;
                }
                uint64_t avail = 0;  // candidate cut points 1..N-1
                if (N >= 2) avail = ((N - 1 >= 64) ? ~0ull : ((1ull << (N - 1)) - 1ull)) << 1;
                uint64_t cuts = 0;
#pragma unroll 1
                for (int b = 0; b < n_breaks; b++) {
                    int m = __popcll(avail);
                    if (m == 0) break;
                    int k = ws.randint(m);  // np.random.choice(options)
                    uint64_t t = avail;
#pragma unroll 1
                    for (int i = 0; i < k; i++) t &= t - 1;
                    int point = __ffsll((long long)t) - 1;
                    avail &= ~(1ull << point);
                    cuts |= 1ull << point;
                }
                int nb = 0;
                vb[nb++] = 0;
#pragma unroll 1
                for (uint64_t t = cuts; t; t &= t - 1) vb[nb++] = (uint8_t)(__ffsll((long long)t) - 1);
                vb[nb] = (uint8_t)N;
                n_int = n_breaks + 1;
#pragma unroll 1
                for (int i = 0; i < n_int; i++) vp[i] = (uint8_t)i;
                __syncwarp();
#pragma unroll 1
                for (int i = n_int - 1; i > 0; i--) {  // np.random.permutation(np.arange(n))
                    int k = ws.randint(i + 1);
                    uint8_t x = vp[i], y = vp[k];
                    __syncwarp();
                    vp[i] = y;
                    vp[k] = x;
                }
            } else {
                vb[0] = 0;
                vb[1] = (uint8_t)N;
                vp[0] = 0;
            }
            __syncwarp();
            const int step_type = sub == 0 ? 0 : 1;
#pragma unroll 1
            for (int i = 0; i < n_int && !err; i++) {
                int p = vp[i];
                interval_step(s, vb[p], vb[p + 1], step_type, temp, llk);
            }
        }
    }

    // tempering.py:62-151 between temperature t (cooler, i) and t-1 (warmer, j)
    __device__ __forceinline__ void chain_swap_step(int t, double temp_i, double temp_j, double &llk_i) {
        const int si = slot(t), sj = slot(t - 1);
        double *lt = llk_t();
        double llk_j = lt[t - 1];
        double prior_i = 0.0, prior_j = 0.0;
        if (PRIOR) {
            prior_i = prior_of_keys(slot_keys(si), -1, 0, -1, 0);
            prior_j = prior_of_keys(slot_keys(sj), -1, 0, -1, 0);
        }
        double post_i = llk_i + prior_i, post_j = llk_j + prior_j;
        double frac_1 = (post_j - post_i) * temp_i;
        double frac_2 = (post_i - post_j) * temp_j;
        double acc = dexp(frac_1 + frac_2);
        if (acc > 1.0) acc = 1.0;
        double val = ws.next_double();
        if (acc >= val) {
            slots = (slots & ~((15u << (4 * t)) | (15u << (4 * (t - 1))))) | ((uint32_t)sj << (4 * t)) |
                    ((uint32_t)si << (4 * (t - 1)));
            __syncwarp();
            lt[t - 1] = llk_i;
            llk_i = llk_j;
        }
    }
};

// ---------------------------------------------------------------------------------------
// Per-item set-up (runs once per item, kept out of line): stage the reads, homozygous fixing
// (mcmc.py:495-541 + snpcalling.py:14-70), compaction of the variable positions, initial-state
// distribution (mcmc.py:455-491, jitutils.py:483-487), gap -> 1.0, prior constants.
// Returns n_het (>= 0) in the low 16 bits, bits per allele in bits 16..23.
// ---------------------------------------------------------------------------------------
template <int CH>
__device__ __noinline__ int assemble_item_setup(const AsmArgs &a, unsigned char *sm, int lane,
                                                const mchb_assemble_item *itp, bool has_initial) {
    constexpr int UPAD = CH * 32;
    const int Nf = itp->n_pos, A = itp->max_allele, P = itp->ploidy;
    const int Uin = itp->n_reads;
    const int U = Uin > 0 ? Uin : 1;  // mcmc.py:132-137: one all-gap read stands in for none
    const double inbreeding = itp->inbreeding;
    const bool has_inb = !isnan(inbreeding);
    const bool pow2 = (P & (P - 1)) == 0;
    const double invP = 1.0 / (double)P;
    double *Rt = asm_rt<CH>(a, sm);
    double *cnt = reinterpret_cast<double *>(sm + a.o_cnt);
    double *homlp = reinterpret_cast<double *>(sm + a.o_homlp);
    double *dist = reinterpret_cast<double *>(sm + a.o_dist);
    double *opr = reinterpret_cast<double *>(sm + a.o_opr);
    double *scv = reinterpret_cast<double *>(sm + a.o_sc);
    double *lgd = reinterpret_cast<double *>(sm + a.o_lgdisp);
    uint8_t *het = sm + a.o_het, *fixa = sm + a.o_fixa, *nall = sm + a.o_nall;
    const int8_t *nal_full = a.n_alleles + itp->nalleles_off;

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
        for (int r = lane; r < Uin; r += 32) cnt[r] = a.counts ? (double)__ldg(a.counts + itp->counts_off + r) : 1.0;
    } else {
        for (int i = lane; i < Nf * A; i += 32) Rt[i * UPAD] = NAN;
        if (lane == 0) cnt[0] = 1.0;
    }
    if (lane == 0) scv[SC_INBREEDING] = inbreeding;
    __syncwarp();

    // ---- homozygous fixing
    int n_het = 0;
#pragma unroll 1
    for (int j = 0; j < Nf; j++) {
        const int nA = nal_full[j];
        const long long u_gens = comb_with_replacement(nA, P);
        uint64_t g = 0;
        double denom = 0.0;
#pragma unroll 1
        for (long long i = 0; i < u_gens; i++) {
            double lprior = 0.0;
            if (has_inb) lprior = snp_log_genotype_prior(g, P, nA, inbreeding);
            double acc = 0.0;
#pragma unroll
            for (int ch = 0; ch < CH; ch++) {
                double rp = 0.0;
#pragma unroll 1
                for (int h = 0; h < P; h++) {
                    double v = Rt[(j * A + nib(g, h)) * UPAD + ch * 32 + lane];
                    double prod = isnan(v) ? 1.0 : v;
                    rp += pow2 ? prod * invP : prod / (double)P;
                }
                acc += dlog(rp) * cnt[ch * 32 + lane];
            }
            double lp = lprior + warp_sum(acc);
            denom = (i == 0) ? lp : add_log_prob(denom, lp);
            if (nib(g, 0) == nib(g, P - 1)) homlp[nib(g, 0)] = lp;
            g = increment_packed(g, P);
        }
        __syncwarp();
        int fixed = 0, fa = 0;
#pragma unroll 1
        for (int al = 0; al < nA; al++) {
            double prob = dexp(homlp[al] - denom);
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
    __syncwarp();
    const int N = n_het;
    if (N == 0) return 0;
    int amax_het = 0;
#pragma unroll 1
    for (int k = 0; k < N; k++) amax_het = max(amax_het, (int)nall[k]);
    if (has_initial) amax_het = max(amax_het, A);  // user states may use any allele < max_allele
    int B = 1;
    while ((1 << B) < amax_het) B++;

    // ---- compact the variable positions (het[k] >= k, ascending: in-place is safe)
#pragma unroll 1
    for (int k = 0; k < N; k++) {
        int j = het[k];
        if (j != k)
            for (int i = lane; i < A * UPAD; i += 32) Rt[k * A * UPAD + i] = Rt[j * A * UPAD + i];
        __syncwarp();
    }
    // ---- initial-state distribution
    if (!has_initial) {
#pragma unroll 1
        for (int k = 0; k < N; k++) {
            int n_nonzero = 0;
            uint32_t gapmask = 0;
#pragma unroll 1
            for (int al = 0; al < A; al++) {
                const double *col = Rt + (k * A + al) * UPAD;
                bool all_nan = true;
#pragma unroll 1
                for (int r = 0; r < U; r++) all_nan = all_nan && isnan(col[r]);
                double tot = 0.0;
                int cn = 0;
                bool all_zero = true;
#pragma unroll 1
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
#pragma unroll 1
            for (int al = 0; al < A; al++) {
                double v = dist[k * A + al];
                if ((gapmask >> al) & 1) v = 1.0 / (double)n_nonzero;
                opr[al] = v;
                s1 += v;
            }
            __syncwarp();  // dist is rewritten below: every lane is done reading it
            double s2 = 0.0;
#pragma unroll 1
            for (int al = 0; al < A; al++) {
                double v = opr[al] / s1;
                dist[k * A + al] = v;
                s2 += v;
            }
#pragma unroll 1
            for (int al = 0; al < A; al++) {
                const double v = dist[k * A + al] / s2;
                __syncwarp();
                dist[k * A + al] = v;
            }
            __syncwarp();
        }
    }
    // ---- gaps become 1.0 from here on (likelihood.py:55-58)
    for (int i = lane; i < N * A * UPAD; i += 32) {
        double v = Rt[i];
        if (isnan(v)) Rt[i] = 1.0;
    }
    __syncwarp();
    // ---- float32 screening tables: allele ratios of bi-allelic flips, counts, error bounds
    {
        float *rat = reinterpret_cast<float *>(sm + a.o_rat);
        float *c32 = reinterpret_cast<float *>(sm + a.o_c32);
        for (int i = lane; i < N * UPAD; i += 32) {
            const int k = i / UPAD, r = i - k * UPAD;
            const double r0 = Rt[(k * A + 0) * UPAD + r];
            const double r1 = A > 1 ? Rt[(k * A + 1) * UPAD + r] : r0;
            if (MCHB_ASM_RAT_HALF(CH)) {
                float f01 = (float)(r1 / r0);
                if (fabsf(f01) < 1.17549435e-38f) f01 = 0.f;  // (NaN stays NaN: the comparison is false)
                rat[k * UPAD + r] = f01;
            } else {
                rat[(k * 2 + 0) * UPAD + r] = (float)(r1 / r0);  // current allele 0 -> 1
                rat[(k * 2 + 1) * UPAD + r] = (float)(r0 / r1);  // current allele 1 -> 0
            }
        }
        double csum = 0.0;
        for (int r = lane; r < UPAD; r += 32) {
            c32[r] = (float)cnt[r];
            csum += cnt[r];
        }
        csum = warp_sum(csum);
        if (lane == 0) {
            // Bounds of |screened - exact| log-likelihood, u = 2^-24:
            //  mutation: a read's screened probability fma(q32[h], rat - 1, rpc) has relative error
            //    <= 1.001 u ((P + 2) rpc / rp + 5) with rpc / rp <= 1.5 / MCHB_SCREEN_MIN_RATIO under the
            //    `sane` test, i.e. <= 9.0e-4 (P + 2); log(1 + e) <= 1.02 e; __logf errs by <= 3.2e-5 on
            //    (1e-30, 1e30); the float32 fma accumulation over U reads by <= 4.3e-6 U per unit count;
            //  structural: products and sums of positive float32 values, relative error
            //    <= 1.001 u (3 N + P + 1) (a factor: u for the stored ratio, u for its reciprocal where
            //    only one direction is stored, u for the product); accumulation over CH chunks per lane.
            const double umax_reads = (double)U;
            //  both: float32 underflow of a term (<= 2^-126 against rp > 1e-30): 1.2e-8 per operation.
            const double under = 1.2e-8 * (double)(N + P + 2);
            scv[SC_ERR_MUT] = MCHB_SCREEN_ERR_SCALE * csum * (1.02 * ((double)(P + 2) * 9.0e-4) + 3.2e-5 + 4.3e-6 * umax_reads + under);
            scv[SC_ERR_STR] = MCHB_SCREEN_ERR_SCALE * csum * (1.02 * 6.0e-8 * (double)(3 * N + P + 1) + 3.2e-5 + 4.3e-6 * (double)(CH + 1) + under);
        }
        __syncwarp();
    }
    // ---- per-item prior constants
    {
        float s = 0.0f;  // mcmc.py:294: float32 arithmetic in numba (int8 array)
#pragma unroll 1
        for (int k = 0; k < N; k++) s += LOGF_INT[nall[k]];
        const double luh = (double)s;
        __syncwarp();
        scv[SC_LUH] = luh;
        if (has_inb && inbreeding != 0.0) {
            double log_disp = log((1.0 - inbreeding) / inbreeding) - luh;
            double disp = exp(log_disp);
            double sum_disp = exp(log_disp + luh);
            scv[SC_LG_SUMDISP] = lgamma(sum_disp);
            scv[SC_LG_P_SUMDISP] = lgamma((double)P + sum_disp);
            scv[SC_LG_DISP] = lgamma(disp);
#pragma unroll 1
            for (int d = 1; d <= P; d++) lgd[d] = lgamma((double)d + disp);
        }
        __syncwarp();
    }
    return N | (B << 16);
}

template <int CH, bool PRIOR>
__global__ void __launch_bounds__(MCHB_ASM_MAXTHREADS(CH), MCHB_ASM_MINCTAS(CH)) assemble_kernel(const __grid_constant__ AsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // The launches of one call (rare shape classes first, the most populated class last) are
    // chained with programmatic dependent launch: the next kernel may start as soon as every CTA
    // of this one is resident, so the few long-running rare items start first and the big class
    // fills the rest of the machine, all in one stream.  (There is no data dependency.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    AsmCtx<CH, PRIOR> c(a, smem_raw + (size_t)warp * a.smem_per_warp, lane);
    {
        // state epochs start at 1 and only grow; memo entries start at epoch 0 (= never valid)
        uint32_t *e = c.epoch();
        for (int i = lane; i < 2 * a.tres; i += 32) e[i] = i < a.tres ? 1u : 0u;
        ScEntry *sc0 = c.scache(0);
        for (int i = lane; i < a.tres * a.scache_n; i += 32) sc0[i].epoch = 0u;
        __syncwarp();
        if (CH >= 2 && a.tres < a.tmax) {
            // every slot of the backing store starts from the same empty memo state
            for (int t = 0; t < a.tmax; t++) slot_copy<CH>(a, c.sm, lane, t, true);
        }
    }

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(MCHB_FULL, w, 0);
        if (w >= a.n_order) break;
        const int item_id = a.order[w];
        const mchb_assemble_item *itp = a.items + item_id;
        const int Nf = itp->n_pos, A = itp->max_allele, P = itp->ploidy, T = itp->n_temps;
        const bool has_initial = a.initial != nullptr && itp->initial_off >= 0;
        c.A = A;
        c.P = P;
        c.U = itp->n_reads > 0 ? itp->n_reads : 1;
        c.err = 0;
        c.evals = 0;
        c.invP = 1.0 / (double)P;
        c.pow2 = (P & (P - 1)) == 0;
        c.ws.init(a.words + (size_t)a.item_stream[item_id] * a.stream_len, a.stream_len,
                  reinterpret_cast<uint32_t *>(c.sm + a.o_ring), lane);
        int8_t *og = a.out_genotypes + itp->genotypes_off;
        double *ol = a.out_llks + itp->llks_off;
        const int step_sz = P * Nf;

        const int setup = assemble_item_setup<CH>(a, c.sm, lane, itp, has_initial);
        const int N = setup & 0xffff;
        c.N = N;
        c.B = setup >> 16;
        c.amask = (1u << c.B) - 1u;

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
            if (N * c.B > 64 || P > MCHB_MAX_PLOIDY || T > MCHB_MAX_TEMPS || T < 1) status = MCHB_ITEM_UNSUPPORTED;
            if (!status && has_initial && itp->initial_nhet != N) status = MCHB_ITEM_INITIAL_SHAPE;
        }
        if (N > 0 && !status) {
            const uint8_t *het = c.het();
            // how each byte of a recorded genotype is produced: 0x8000 | haplotype << 8 | key shift
            // for a variable position, the fixed allele otherwise
            uint16_t *wmap = reinterpret_cast<uint16_t *>(c.sm + a.o_wmap);
            {
                const uint8_t *fixa = c.fixa();
                __syncwarp();
                uint16_t *hmap = reinterpret_cast<uint16_t *>(c.sm + a.o_hmap);
                for (int i = lane; i < step_sz; i += 32) {
                    wmap[i] = (uint16_t)fixa[i % Nf];
                    hmap[i] = (uint16_t)(((i / Nf) << 8) | (i % Nf));  // (haplotype, position) of byte i
                }
                __syncwarp();
                for (int i = lane; i < P * N; i += 32) {
                    const int h = i / N, k = i - h * N;
                    wmap[h * Nf + het[k]] = (uint16_t)(0x8000 | (h << 8) | (c.B * k));
                }
                __syncwarp();
            }
            double *dist = c.dist(), *opr = c.opr();
            const int brow_i = min(N, a.break_rows - 1);
            const double *brow = a.break_table + (size_t)brow_i * a.break_stride;
            const int blen = a.break_len[brow_i];
            const double *temps = a.temperatures + itp->temps_off;
            double *lt = c.llk_t();
            {
                // np.cumsum(break_dist), sequential like numba (jitutils.py:92)
                double *bcs = reinterpret_cast<double *>(c.sm + a.o_bcs);
                double acc = 0.0;
                __syncwarp();
#pragma unroll 1
                for (int i = 0; i < blen; i++) {
                    acc += brow[i];
                    bcs[i] = acc;
                }
                __syncwarp();
            }

#pragma unroll 1
            for (int chain = 0; chain < a.chains && !c.err; chain++) {
                // ---- initial genotype written into state slot 0
                uint64_t *k0 = c.keys(0);
                __syncwarp();
                if (has_initial) {
                    const int8_t *src = a.initial + itp->initial_off + (size_t)chain * P * N;
#pragma unroll 1
                    for (int h = 0; h < P; h++) {
                        uint64_t k = 0;
#pragma unroll 1
                        for (int j = 0; j < N; j++) k |= (uint64_t)((uint32_t)(uint8_t)src[h * N + j] & c.amask) << (c.B * j);
                        k0[h] = k;
                    }
                } else {
#pragma unroll 1
                    for (int h = 0; h < P; h++) {
                        uint64_t k = 0;
#pragma unroll 1
                        for (int j = 0; j < N; j++) {
#pragma unroll 1
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
                for (int i = lane; i < 4 * T; i += 32) c.hot()[i] = 0;
                // one resident slot, the others in the backing store (never in the CH = 1 kernels,
                // whose hot loop stays free of the swap code)
                const bool swap = CH >= 2 && a.tres < a.tmax && T > 1;
                c.qslot = 0;
                if (swap) {
                    uint64_t *kk = reinterpret_cast<uint64_t *>(c.oll());  // the initial keys survive the slot loads
                    for (int h = lane; h < P; h += 32) kk[h] = k0[h];
                    __syncwarp();
#pragma unroll 1
                    for (int t = 0; t < T; t++) {
                        slot_copy<CH>(a, c.sm, lane, t, false);
                        c.qslot = t;
#pragma unroll 1
                        for (int h = 0; h < P; h++) c.commit(0, h, kk[h]);
                        slot_copy<CH>(a, c.sm, lane, t, true);
                    }
                }
                {
                    if (!swap) {
#pragma unroll 1
                        for (int h = 0; h < P; h++) {
                            const uint64_t k = k0[h];
#pragma unroll 1
                            for (int t = 0; t < T; t++) c.commit(t, h, k);
                        }
                    }
                    __syncwarp();
                    double llk0 = c.eval_llk(0);
                    c.evals--;  // the initial evaluation is not a proposal
#pragma unroll 1
                    for (int t = 0; t < T; t++) lt[t] = llk0;
                    __syncwarp();
                }
                int8_t *ogc = og + (size_t)chain * a.steps * step_sz;
                double *olc = ol + (size_t)chain * a.steps;
#pragma unroll 1
                for (int step = 0; step < a.steps && !c.err; step++) {
                    double llk = 0.0;
#pragma unroll 1
                    for (int t = 0; t < T && !c.err; t++) {
                        llk = lt[t];
                        int s = c.slot(t);
                        const double temp = temps[t];
                        if (isnan(llk)) {
                            c.err = MCHB_ITEM_NAN_LLK;
                            break;
                        }
                        c.prof_t = t;
                        const long long pc0 = MCHB_PROF_CLOCK();
                        if (swap) {
                            slot_copy<CH>(a, c.sm, lane, s, false);
                            c.qslot = s;
                            s = 0;
                        }
                        const long long pc1 = MCHB_PROF_CLOCK();
                        c.mutation_compound_step(s, t, temp, llk);
                        const long long pc2 = MCHB_PROF_CLOCK();
                        if (c.err) break;
                        c.structural_substeps(s, temp, llk, brow, blen);
                        const long long pc3 = MCHB_PROF_CLOCK();
                        if (swap) slot_copy<CH>(a, c.sm, lane, c.slot(t), true);
                        const long long pc4 = MCHB_PROF_CLOCK();
                        MCHB_PROF_ADD(t, PK_CYC_MUT, pc2 - pc1);
                        MCHB_PROF_ADD(t, PK_CYC_STR, pc3 - pc2);
                        MCHB_PROF_ADD(t, PK_CYC_SLOT, (pc1 - pc0) + (pc4 - pc3));
                        if (c.err) break;
                        if (t > 0) c.chain_swap_step(t, temp, temps[t - 1], llk);
                        __syncwarp();
                        lt[t] = llk;
                    }
                    if (c.err) break;
                    // ---- record the cold chain (state of the last temperature, mcmc.py:418-425)
                    {
                        const uint64_t *ks = c.slot_keys(c.slot(T - 1));
                        int8_t *dst = ogc + (size_t)step * step_sz;
                        __syncwarp();
                        if (a.sort_recorded) {
                            // GenotypeMultiTrace keeps every step with its haplotypes sorted
                            // lexicographically, first position most significant (assemble/classes.py:
                            // 265-278 -> encoding/integer/sequence.py:78-110).  Fixed positions are equal
                            // in all haplotypes, so the packed keys decide: lane h ranks its key
                            // (ties keep their order), and the rows are written at their ranks.
                            uint8_t *rk = c.sm + a.o_rank;
                            const uint16_t *hmap = reinterpret_cast<const uint16_t *>(c.sm + a.o_hmap);
                            if (lane < P) {
                                const uint64_t km = ks[lane];
                                int r = 0;
#pragma unroll 1
                                for (int k = 0; k < P; k++) {
                                    const uint64_t kk = ks[k];
                                    const uint64_t d = kk ^ km;
                                    bool less = k < lane;  // equal keys
                                    if (d != 0) {
                                        const int sh = ((__ffsll((long long)d) - 1) / c.B) * c.B;  // first differing position
                                        less = ((uint32_t)(kk >> sh) & c.amask) < ((uint32_t)(km >> sh) & c.amask);
                                    }
                                    r += less ? 1 : 0;
                                }
                                rk[lane] = (uint8_t)r;
                            }
                            __syncwarp();
#pragma unroll 1
                            for (int i = lane; i < step_sz; i += 32) {
                                const uint32_t m = wmap[i], hp = hmap[i];
                                const uint32_t v = (uint32_t)(ks[(m >> 8) & 0x7f] >> (m & 63)) & c.amask;
                                dst[(int)rk[hp >> 8] * Nf + (int)(hp & 255u)] = (int8_t)((m & 0x8000u) ? v : m);
                            }
                        } else {
#pragma unroll 1
                            for (int i = lane; i < step_sz; i += 32) {
                                const uint32_t m = wmap[i];
                                const uint32_t v = (uint32_t)(ks[(m >> 8) & 0x7f] >> (m & 63)) & c.amask;
                                dst[i] = (int8_t)((m & 0x8000u) ? v : m);
                            }
                        }
                        if (lane == 0) olc[step] = llk;
                        __syncwarp();
                    }
                }
            }
            status = c.err;
            if (!status && c.ws.exhausted()) status = MCHB_ITEM_RNG_EXHAUSTED;
        }
        if (lane == 0) {
            mchb_item_result r;
            r.status = status;
            r.n_het = N;
            r.rng_words = c.ws.cur;
            r.llk_evals = c.evals;
            a.results[item_id] = r;
        }
        if (a.chunk_done) {
            __threadfence_system();  // every lane: its trace stores are visible before the count moves
            __syncwarp();
            if (lane == 0) {
                const int k = item_id / a.chunk_items;
                const uint32_t target = (uint32_t)min(a.chunk_items, a.n_items - k * a.chunk_items);
                if (atomicAdd(a.chunk_done + k, 1u) + 1u == target) {
                    __threadfence_system();
                    *reinterpret_cast<volatile uint32_t *>(a.chunk_flags + k) = 1u;
                    __threadfence_system();
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace mchb
