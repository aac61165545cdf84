// Trace post-processing on the device: what the reference does on
// the host with per-step Python loops —
//   * GenotypeMultiTrace.__post_init__: lexicographic sort of the haplotypes of every recorded
//     genotype (mchap/assemble/classes.py:265-278, encoding/integer/sequence.py:78-110),
//   * .burn(n) (classes.py:280-305),
//   * .posterior() / .split(): unique genotypes in order of first occurrence with their counts,
//     merged over the chains and per chain (classes.py:307-339, mset.py:242-284, 361-392)
// — as one pass over the trace where it lies in HBM.  One warp = one item.  The host only sees
// the tallies (a few hundred bytes per item instead of the 120 KB trace of the headline shape);
// probabilities, np.flip(np.argsort(.)) and the incongruence test stay on the host on those
// tallies, so ties come out exactly as in the reference.
#pragma once

#include <cstdint>

#include "../../include/mchap_b200.h"

namespace mchb {

struct TallyArgs {
    const mchb_tally_item *items;
    int32_t n_items;
    const int8_t *genotypes;  // elements of ES bytes (ES = template parameter of the kernel)
    int8_t *out_states;
    int32_t *out_counts;
    int32_t *out_first;
    mchb_item_result *results;
    int32_t *work_counter;
    int32_t smem_per_warp;   // bytes
    int32_t pn_max;          // largest ploidy * n_pos of the batch
    int32_t tile_bytes;      // staging tile size (>= pn_max)
    int32_t unique_max;      // largest max_unique of the batch
};

// lexicographic order of two rows of n signed elements in shared memory
template <typename T>
__device__ __forceinline__ int row_compare(const T *x, const T *y, int n) {
#pragma unroll 1
    for (int j = 0; j < n; j++) {
        const T xv = x[j], yv = y[j];
        if (xv != yv) return xv < yv ? -1 : 1;
    }
    return 0;
}

// T = int8_t: assembly traces int8[chains, steps, ploidy, n_pos]; T = int32_t: calling traces
// int32[chains, steps, ploidy] (n_pos = 1: a row is one allele index)
template <typename T>
__global__ void __launch_bounds__(128) tally_kernel(const __grid_constant__ TallyArgs a) {
    constexpr int ES = (int)sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char *sm = smem_raw + (size_t)warp * a.smem_per_warp;
    uint32_t *hashes = reinterpret_cast<uint32_t *>(sm);                 // [unique_max]
    int8_t *tile = reinterpret_cast<int8_t *>(hashes + a.unique_max);    // [tile_bytes] staged steps
    int8_t *prev = tile + a.tile_bytes;                                  // [pn_max] last distinct raw step
    int8_t *sorted = prev + a.pn_max;                                    // [pn_max] its sorted form

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(a.work_counter, 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= a.n_items) break;
        const mchb_tally_item it = a.items[w];
        const int P = it.ploidy, N = it.n_pos, C = it.chains, S = it.steps;
        const int PN = P * N * ES;   // bytes of one recorded genotype
        const int NB = N * ES;       // bytes of one row
        const int burn = min(max(it.burn, 0), S);
        const int U = it.max_unique;
        const int8_t *g = a.genotypes + it.genotypes_off * ES;
        int8_t *states = a.out_states + it.states_off * ES;
        int32_t *counts = a.out_counts + it.tallies_off;
        int32_t *first = a.out_first + it.tallies_off;
        for (int i = lane; i < U * C; i += 32) {
            counts[i] = 0;
            first[i] = -1;
        }
        __syncwarp();
        int n_unique = 0;
        int status = MCHB_ITEM_OK;
        const int tile_steps = PN > 0 ? max(1, a.tile_bytes / PN) : 1;
#pragma unroll 1
        for (int c = 0; c < C && !status; c++) {
            int cur = -1;   // index of the running state (-1: none yet in this chain)
            int run = 0;    // steps of the running state not yet added to counts
#pragma unroll 1
            for (int s0 = burn; s0 < S && !status; s0 += tile_steps) {
                const int nst = min(tile_steps, S - s0);
                const int nbytes = nst * PN;
                const int8_t *src = g + ((size_t)c * S + s0) * PN;
                __syncwarp();
#pragma unroll 8
                for (int b = lane; b < nbytes; b += 32) tile[b] = src[b];
                __syncwarp();
#pragma unroll 1
                for (int t = 0; t < nst; t++) {
                    const int8_t *raw = tile + t * PN;
                    bool same = cur >= 0;
                    for (int b = lane; b < PN; b += 32) same = same && (raw[b] == prev[b]);
                    if (__all_sync(0xffffffffu, same)) {
                        run++;
                        continue;
                    }
                    // ---- a different raw step: canonical form = haplotypes in lexicographic order
                    __syncwarp();
                    for (int b = lane; b < PN; b += 32) prev[b] = raw[b];
                    __syncwarp();
                    if (lane < P) {
                        int rank = 0;
#pragma unroll 1
                        for (int k = 0; k < P; k++) {
                            const int d = row_compare<T>(reinterpret_cast<const T *>(prev + k * NB),
                                                         reinterpret_cast<const T *>(prev + lane * NB), N);
                            rank += (d < 0) || (d == 0 && k < lane);
                        }
                        for (int j = 0; j < NB; j++) sorted[rank * NB + j] = prev[lane * NB + j];
                    }
                    __syncwarp();
                    uint32_t hsh = 0;
                    for (int b = lane; b < PN; b += 32)
                        hsh += ((uint32_t)(uint8_t)sorted[b] + 1u) * (2654435761u * (uint32_t)(2 * b + 1));
#pragma unroll
                    for (int m = 16; m > 0; m >>= 1) hsh += __shfl_xor_sync(0xffffffffu, hsh, m);
                    hsh ^= hsh >> 15;
                    // ---- look it up among the states seen so far (first-occurrence order)
                    int found = -1;
#pragma unroll 1
                    for (int base = 0; base < n_unique && found < 0; base += 32) {
                        const int i = base + lane;
                        unsigned cand = __ballot_sync(0xffffffffu, i < n_unique && hashes[i] == hsh);
                        while (cand && found < 0) {
                            const int idx = base + __ffs(cand) - 1;
                            cand &= cand - 1;
                            const int8_t *st = states + (size_t)idx * PN;
                            bool eq = true;
                            for (int b = lane; b < PN; b += 32) eq = eq && (st[b] == sorted[b]);
                            if (__all_sync(0xffffffffu, eq)) found = idx;
                        }
                    }
                    if (found < 0) {
                        if (n_unique >= U) {
                            status = MCHB_ITEM_TALLY_OVERFLOW;
                            break;
                        }
                        found = n_unique++;
                        int8_t *st = states + (size_t)found * PN;
                        for (int b = lane; b < PN; b += 32) st[b] = sorted[b];
                        if (lane == 0) hashes[found] = hsh;
                        __syncwarp();
                    }
                    if (found != cur) {
                        if (lane == 0) {
                            if (cur >= 0) counts[cur * C + c] += run;
                            if (first[found * C + c] < 0) first[found * C + c] = s0 + t - burn;
                        }
                        run = 0;
                        cur = found;
                    }
                    run++;
                }
            }
            if (lane == 0 && cur >= 0 && !status) counts[cur * C + c] += run;
            __syncwarp();
        }
        if (lane == 0) {
            mchb_item_result r;
            r.status = status;
            r.n_het = n_unique;
            r.rng_words = 0;
            r.llk_evals = 0;
            a.results[w] = r;
        }
        __threadfence();
        __syncwarp();
    }
}

}  // namespace mchb
