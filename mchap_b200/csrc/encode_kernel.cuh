// Read encoding + de-duplication on the device: what the reference
// does on the host between the BAM reader and the samplers (mchap/application/baseclass.py:194-209)
//   * as_probabilistic (mchap/encoding/integer/transcode.py:16-77, called through
//     io/bam.py:251-289 encode_read_distributions): integer allele calls + P(call correct) ->
//     f64[n_reads, n_pos, max_allele] probability rows (gaps NaN, alleles beyond a position's
//     allele count 0),
//   * mset.unique_counts (mchap/mset.py:242-284, 361-392): the distinct encoded reads in order of
//     first occurrence (byte equality) and how often each occurs
// — as one pass per (locus, sample) item: one warp = one item, reads one after another, lanes over
// the N * A elements of a read.  The output is exactly the (reads, read_counts) pair the samplers
// take, already in the packed layout of mchb_assemble_batch / mchb_call_*.
#pragma once

#include <cstdint>

#include "../../include/mchap_b200.h"

namespace mchb {

struct EncodeArgs {
    const mchb_encode_item *items;
    int32_t n_items;
    const int8_t *calls;
    const double *probs;
    const int8_t *n_alleles;
    double error_factor;
    double *out_reads;
    int64_t *out_counts;
    mchb_item_result *results;
    int32_t *work_counter;
    int32_t smem_per_warp;  // bytes
    int32_t rmax;           // most reads of an item in the batch
    int32_t emax;           // most elements (n_pos * max_allele) of a read in the batch
};

__global__ void __launch_bounds__(128) encode_reads_kernel(const __grid_constant__ EncodeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *sm = smem_raw + (size_t)warp * a.smem_per_warp;
    double *row = reinterpret_cast<double *>(sm);                       // [emax] the read being encoded
    uint32_t *hashes = reinterpret_cast<uint32_t *>(row + a.emax);      // [rmax] hashes of the distinct reads
    const unsigned long long NAN_BITS = 0x7ff8000000000000ull;          // numpy's np.nan

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= a.n_items) break;
        const mchb_encode_item it = a.items[w];
        const int R = it.n_reads, N = it.n_pos, A = it.max_allele;
        const int E = N * A;
        const int8_t *calls = a.calls + it.calls_off;
        const double *probs = a.probs + it.probs_off;
        const int8_t *nall = a.n_alleles + it.nalleles_off;
        double *out = a.out_reads + it.reads_off;
        int64_t *cnt = a.out_counts + it.counts_off;
        int n_unique = 0;
#pragma unroll 1
        for (int r = 0; r < R; r++) {
            // ---- as_probabilistic for read r (transcode.py:61-75, in its order of assignments)
            __syncwarp();
            uint32_t hsh = 0;
            for (int e = lane; e < E; e += 32) {
                const int j = e / A, al = e - j * A;
                const int c = calls[(size_t)r * N + j];
                const double p = probs[(size_t)r * N + j];
                double v = (c == al) ? p : (1.0 - p) / a.error_factor;
                if (c < 0) v = __longlong_as_double((long long)NAN_BITS);
                if ((int)nall[j] <= al) v = 0.0;
                row[e] = v;
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                hsh += ((uint32_t)bits ^ (uint32_t)(bits >> 32) ^ 0x9e3779b9u) * (2654435761u * (uint32_t)(2 * e + 1));
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) hsh += __shfl_xor_sync(0xffffffffu, hsh, m);
            hsh ^= hsh >> 15;
            __syncwarp();
            // ---- mset.unique / count: first-occurrence order, byte equality
            int found = -1;
#pragma unroll 1
            for (int base = 0; base < n_unique && found < 0; base += 32) {
                const int i = base + lane;
                unsigned cand = __ballot_sync(0xffffffffu, i < n_unique && hashes[i] == hsh);
                while (cand && found < 0) {
                    const int idx = base + __ffs(cand) - 1;
                    cand &= cand - 1;
                    const double *st = out + (size_t)idx * E;
                    bool eq = true;
                    for (int e = lane; e < E; e += 32)
                        eq = eq && (__double_as_longlong(st[e]) == __double_as_longlong(row[e]));
                    if (__all_sync(0xffffffffu, eq)) found = idx;
                }
            }
            if (found < 0) {
                found = n_unique++;
                double *st = out + (size_t)found * E;
                for (int e = lane; e < E; e += 32) st[e] = row[e];
                if (lane == 0) {
                    hashes[found] = hsh;
                    cnt[found] = 1;
                }
                __syncwarp();
            } else if (lane == 0) {
                cnt[found] += 1;
            }
        }
        if (lane == 0) {
            mchb_item_result res;
            res.status = MCHB_ITEM_OK;
            res.n_het = n_unique;
            res.rng_words = 0;
            res.llk_evals = 0;
            a.results[w] = res;
        }
        __syncwarp();
    }
}

}  // namespace mchb
