// aux_kernels.cuh — MT19937 stream generator, batched read log-likelihood (K1), rank/unrank.
#pragma once
#include "common.cuh"

namespace mchb {

// ---------------------------------------------------------------------------------------
// MT19937 as numba seeds and steps it (numba/_random.c numba_rnd_init / numba_rnd_shuffle,
// numba/cpython/randomimpl.py:109-132 tempering).  One CTA per stream; the 624-word state lives
// in shared memory and is regenerated in four dependency phases per block.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v, uint32_t m) {
    uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__global__ void __launch_bounds__(256) mt19937_fill_kernel(const uint32_t *seeds, uint32_t *out, int64_t len) {
    __shared__ uint32_t mt[624];
    const int tid = threadIdx.x;
    uint32_t *dst = out + (size_t)blockIdx.x * len;
    if (tid == 0) {
        uint32_t seed = seeds[blockIdx.x];
        for (int pos = 0; pos < 624; pos++) {
            mt[pos] = seed;
            seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)pos + 1u;
        }
    }
    __syncthreads();
    for (int64_t base = 0; base < len; base += 624) {
        uint32_t v = 0;
        // phase A: i in [0, 227) reads old mt[i], mt[i+1], mt[i+397]
        if (tid < 227) v = mt_twist(mt[tid], mt[tid + 1], mt[tid + 397]);
        __syncthreads();
        if (tid < 227) mt[tid] = v;
        __syncthreads();
        // phase B: i in [227, 454) reads old mt[i], mt[i+1], new mt[i-227]
        if (tid < 227) v = mt_twist(mt[tid + 227], mt[tid + 228], mt[tid]);
        __syncthreads();
        if (tid < 227) mt[tid + 227] = v;
        __syncthreads();
        // phase C: i in [454, 623) reads old mt[i], mt[i+1], new mt[i-227]
        if (tid < 169) v = mt_twist(mt[tid + 454], mt[tid + 455], mt[tid + 227]);
        __syncthreads();
        if (tid < 169) mt[tid + 454] = v;
        __syncthreads();
        // phase D: i = 623 reads old mt[623], new mt[0], new mt[396]
        if (tid == 0) mt[623] = mt_twist(mt[623], mt[0], mt[396]);
        __syncthreads();
        for (int i = tid; i < 624; i += 256) {
            if (base + i < len) {
                uint32_t y = mt[i];
                y ^= (y >> 11);
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= (y >> 18);
                dst[base + i] = y;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K1: assemble/likelihood.py:18-70 for a batch of (reads, genotype) pairs, one warp per pair.
// Lane r evaluates read r (products over positions in order, gaps skipped, sum over haplotypes
// in order, log, * count); the sum over reads is accumulated IN READ ORDER by shuffling each
// lane's term to the accumulator, so the value differs from the reference only through the ULP
// differences between CUDA's and glibc's log().
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) llk_batch_kernel(const mchb_llk_item *items, int64_t n_items,
                                                        const double *reads, const int64_t *counts,
                                                        const int8_t *genotypes, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_items) return;
    const mchb_llk_item it = items[w];
    const int U = it.n_reads, N = it.n_pos, A = it.max_allele, P = it.ploidy;
    const double *R = reads + it.reads_off;
    const int8_t *G = genotypes + it.geno_off;
    const double dP = (double)P;
    double llk = 0.0;
    for (int r0 = 0; r0 < U; r0 += 32) {
        const int r = r0 + lane;
        double term = 0.0;
        if (r < U) {
            double rp = 0.0;
            for (int h = 0; h < P; h++) {
                double prod = 1.0;
                for (int j = 0; j < N; j++) {
                    int a = G[h * N + j];
                    double v = __ldg(R + ((size_t)r * N + j) * A + a);
                    if (!isnan(v)) prod *= v;
                }
                rp += prod / dP;
            }
            term = log(rp);
            if (counts) term *= (double)__ldg(counts + it.counts_off + r);
        }
        const int n = min(32, U - r0);
        for (int k = 0; k < n; k++) llk += __shfl_sync(MCHB_FULL, term, k);
    }
    if (lane == 0) out[w] = llk;
}

// ---------------------------------------------------------------------------------------
// jitutils.py:253-276 / 279-318: VCF-order multiset rank and unrank, one thread per genotype
// ---------------------------------------------------------------------------------------
__global__ void rank_kernel(const int64_t *alleles, int64_t n, int ploidy, int64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t *g = alleles + i * ploidy;
    int64_t index = 0;
    bool bad = false;
    for (int k = 0; k < ploidy; k++) {
        int64_t al = g[k];
        if (al >= 0) index += comb_with_replacement(al, k + 1);
        else bad = true;
    }
    out[i] = bad ? -1 : index;
}

__global__ void unrank_kernel(const int64_t *index, int64_t n, int ploidy, int64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t *g = out + i * ploidy;
    int64_t remainder = index[i];
    if (remainder < 0) {
        for (int k = 0; k < ploidy; k++) g[k] = -1;
        return;
    }
    for (int it = 0; it < ploidy; it++) {
        int p = ploidy - it;
        int64_t a = -1, nw = 0, prev = 0;
        while (nw <= remainder) {
            a += 1;
            prev = nw;
            nw = comb_with_replacement(a, p);
        }
        a -= 1;
        remainder -= prev;
        g[p - 1] = a;
    }
}

// ---------------------------------------------------------------------------------------
// encoding/integer/stats.py:18-39 minimum_error_correction for a batch of (read calls, genotype)
// pairs, one warp per pair: lane = read; per read the number of called positions (call >= 0) that
// differ from a haplotype, minimised over the haplotypes.  Byte/integer work: reads the int8 calls
// once (coalesced along the read row), the genotype through the read-only path.
// Outputs per item: sum over reads (what the CLIs report as MEC), number of calls >= 0 (the MECP
// denominator, application/assemble.py:160-163) and optionally the per-read values.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mec_batch_kernel(const mchb_mec_item *items, int64_t n_items,
                                                        const int8_t *calls, const int8_t *genotypes,
                                                        int64_t *out_mec, int64_t *out_called, int32_t *out_per_read) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_items) return;
    const mchb_mec_item it = items[w];
    const int R = it.n_reads, N = it.n_pos, P = it.ploidy;
    const int8_t *C = calls + it.calls_off;
    const int8_t *G = genotypes + it.geno_off;
    long long total = 0, called = 0;
    for (int r = lane; r < R; r += 32) {
        const int8_t *row = C + (size_t)r * N;
        int best = 0x7fffffff, ncall = 0;
        for (int j = 0; j < N; j++) ncall += row[j] >= 0;
        for (int h = 0; h < P; h++) {
            int d = 0;
            for (int j = 0; j < N; j++) {
                const int c = row[j];
                d += (c >= 0) && (c != (int)__ldg(G + h * N + j));
            }
            best = min(best, d);
        }
        if (P == 0) best = 0;
        total += best;
        called += ncall;
        if (out_per_read) out_per_read[it.per_read_off + r] = best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        total += __shfl_xor_sync(MCHB_FULL, total, o);
        called += __shfl_xor_sync(MCHB_FULL, called, o);
    }
    if (lane == 0) {
        out_mec[w] = total;
        if (out_called) out_called[w] = called;
    }
}

// ---------------------------------------------------------------------------------------
// FP64 SIMT pipe probe: 8 independent DFMA chains per thread.  Used by bench.py to measure the
// roofline denominator of the (FP64-bound) MCMC kernels live on the device under test.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double x, double y) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, x, y);
        a1 = fma(a1, x, y);
        a2 = fma(a2, x, y);
        a3 = fma(a3, x, y);
        a4 = fma(a4, x, y);
        a5 = fma(a5, x, y);
        a6 = fma(a6, x, y);
        a7 = fma(a7, x, y);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace mchb
