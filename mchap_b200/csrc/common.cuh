// common.cuh — device-side building blocks shared by the sm_100a kernels.
//
// Conventions
//  * one warp owns one (locus, sample) item; "uniform" values are held identically by all
//    32 lanes (every lane executes the scalar control flow, so no broadcasts are needed);
//  * per-read data is distributed over lanes: read r lives in lane r%32, chunk r/32;
//  * compiled with -fmad=false: the reference (numba, no fast-math) never contracts a*b+c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mchap_b200.h"

#define MCHB_MAX_PLOIDY 16
#define MCHB_MAX_TEMPS 8
#define MCHB_MAX_READS 1024  // distinct reads per assemble item (read chunks of 32: 1, 2, 3, 4, 8, 16, 32 per lane)
#define MCHB_FULL 0xffffffffu

namespace mchb {

// Host-initialised tables (glibc values, so integer-argument logs / lgammas are bit-identical
// to what numba calls on the CPU).  k < MCHB_TABLE_N.
//   LOG_INT[k] = log(k) (LOG_INT[0] = -inf), LOG_INV_INT[k] = log(1.0 / k),
//   LGAMMA_INT[k] = lgamma(k), LOGF_INT[k] = logf((float)k)
#define MCHB_TABLE_N 272
__constant__ double LOG_INT[MCHB_TABLE_N];
__constant__ double LOG_INV_INT[MCHB_TABLE_N];
__constant__ double LGAMMA_INT[MCHB_TABLE_N];
__constant__ float LOGF_INT[MCHB_TABLE_N];

// ---------------------------------------------------------------------------------------
// Word source: numba's MT19937 output stream, pre-generated in global memory (tempered words).
// A 128-word ring in shared memory (4 blocks of 32) always holds the block under the cursor and the
// next two, so that the sequential consumer (next_u32: one broadcast LDS) and lane-parallel
// consumers (word_at(off), off < 64: lane i decodes the draw of the i-th pending sub-step) read
// without touching global memory.  The cursor is warp-uniform.
// Reference semantics: numba/cpython/randomimpl.py get_next_int32 109-132.
// ---------------------------------------------------------------------------------------
// One copy of the ring refill for all its call sites (instruction-cache footprint): lane l loads
// word blk * 32 + l of the stream (0 past its end) into the slot of block blk.
__device__ __noinline__ void word_ring_refill(const uint32_t *base, uint32_t *ring, int len, int blk, int lane) {
    const int i = blk * 32 + lane;
    __syncwarp();
    ring[(blk & 3) * 32 + lane] = (i < len) ? __ldg(base + i) : 0u;
    __syncwarp();
}

struct WordStream {
    const uint32_t *base;
    uint32_t *ring;   // shared memory, 128 words
    int len;          // words available (< 2^31; reads beyond it return 0 and flag exhaustion at the end)
    int cur;          // words consumed so far (uniform)
    int lane;

    __device__ __forceinline__ uint32_t load_block(int blk) const {
        const int i = blk * 32 + lane;
        return (i < len) ? __ldg(base + i) : 0u;
    }
    __device__ __forceinline__ void init(const uint32_t *b, int64_t n, uint32_t *ring_smem, int lane_id) {
        base = b;
        len = (int)n;
        ring = ring_smem;
        lane = lane_id;
        cur = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; k++) ring[k * 32 + lane] = load_block(k);
        __syncwarp();
    }
    __device__ __forceinline__ bool exhausted() const { return cur > len; }
    // the cursor just entered block cur/32: fetch block cur/32 + 2 into the slot of block cur/32 - 2
    __device__ __forceinline__ void refill() { word_ring_refill(base, ring, len, (cur >> 5) + 2, lane); }
    __device__ __forceinline__ uint32_t next_u32() {
        const uint32_t w = ring[cur & 127];
        cur++;
        if ((cur & 31) == 0) refill();
        return w;
    }
    // word at cursor + off without consuming it (0 <= off < 64)
    __device__ __forceinline__ uint32_t word_at(int off) const { return ring[(cur + off) & 127]; }
    // consume m words at once (m <= 64)
    __device__ __forceinline__ void advance(int m) {
        const int target = cur + m;
        while ((cur >> 5) < (target >> 5)) {
            cur = ((cur >> 5) + 1) << 5;
            refill();
        }
        cur = target;
    }
    __device__ __forceinline__ static double to_double(uint32_t w0, uint32_t w1) {
        const uint32_t a = w0 >> 5, b = w1 >> 6;  // randomimpl.py:134-147 get_next_double
        return ((double)b + (double)a * 67108864.0) / 9007199254740992.0;
    }
    __device__ __forceinline__ double next_double() {
        const uint32_t w0 = next_u32();
        const uint32_t w1 = next_u32();
        return to_double(w0, w1);
    }
    // the double that the (off/2)-th next call of next_double() would return (off even, < 63)
    __device__ __forceinline__ double double_at(int off) const { return to_double(word_at(off), word_at(off + 1)); }
    // randomimpl.py:454-520 _randrange_impl (state "np", n <= 2^31 here); n == 1 draws nothing.
    // Past the end of the stream the words are 0, which ends the rejection loop.
    __device__ __forceinline__ int randint(int n) {
        if (n == 1) return 0;
        const uint32_t mask = 0xffffffffu >> __clz(n - 1);
        for (;;) {
            uint32_t r = next_u32() & mask;
            if ((int)r < n) return (int)r;
        }
    }
};

// Out-of-line log / exp for code that runs once per item (set-up), to keep it small.  In the
// steady-state loop the libdevice bodies stay inlined: sharing them through calls was measured
// 5 % slower (round-1 profile notes in profiles/README.md).
__device__ __noinline__ double dlog(double x) { return log(x); }
__device__ __noinline__ double dexp(double x) { return exp(x); }

// ---------------------------------------------------------------------------------------
// log-space helpers (jitutils.py:6-74)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double add_log_prob(double x, double y) {
    if (x == -INFINITY && y == -INFINITY) return -INFINITY;
    if (x > y) return x + log1p(exp(y - x));
    return y + log1p(exp(x - y));
}

// np.minimum(0.0, x): NaN propagates
__device__ __forceinline__ double np_minimum0(double x) {
    if (isnan(x)) return x;
    return x < 0.0 ? x : 0.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(MCHB_FULL, v, m);
    return v;  // bit-identical in every lane (fp add commutes)
}

// searchsorted(cumsum, u, side="right") on a uniform smem array (numba arraymath.py:3841-3860)
__device__ __forceinline__ int searchsorted_right(const double *cs, int n, double u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + ((hi - lo) >> 1);
        if (cs[mid] <= u) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------
// exact integer binomials (jitutils.py:186-250)
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int64_t gcd_i64(int64_t x, int64_t y) {
    while (y != 0) {
        int64_t t = x % y;
        x = y;
        y = t;
    }
    return x;
}
__host__ __device__ inline int64_t comb_exact(int64_t n, int64_t k) {
    if (k > n) return 0;
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
__host__ __device__ inline int64_t comb_with_replacement(int64_t n, int64_t k) {
    if (n == 0 && k == 0) return 0;  // jitutils.py:232-233
    return comb_exact(n + k - 1, k);
}

// calling/prior.py:116-179 log_genotype_prior with frequencies == None or given.
// g: sorted or unsorted allele indices (uniform array in local/shared memory), P <= 16.
__device__ inline double calling_log_genotype_prior(const int *g, int P, int H, double inbreeding,
                                                    const double *freqs) {
    int dosage[MCHB_MAX_PLOIDY];
    for (int i = 0; i < P; i++) dosage[i] = 0;
    for (int i = 0; i < P; i++) {  // calling/utils.py:7-35 allelic_dosage
        int j = 0;
        while (g[j] != g[i]) j++;
        dosage[j] += 1;
    }
    if (inbreeding == 0.0) {
        double ln_num = LGAMMA_INT[P + 1];
        double ln_denom = 0.0;
        for (int i = 0; i < P; i++) ln_denom += LGAMMA_INT[dosage[i] + 1];
        double ln_perms = ln_num - ln_denom;
        if (!freqs) return ln_perms - (double)P * log((double)H);
        double prod = 1.0;
        for (int i = 0; i < P; i++) prod *= freqs[g[i]];
        return ln_perms + log(prod);
    }
    double scale = (1.0 - inbreeding) / inbreeding;
    double alpha_const = 0.0, sum_alphas = 0.0;
    if (!freqs) {
        alpha_const = (1.0 / (double)H) * scale;
        sum_alphas = alpha_const * (double)H;
    } else {
        for (int a = 0; a < H; a++) sum_alphas += freqs[a] * scale;
    }
    double left = (LGAMMA_INT[P + 1] + lgamma(sum_alphas)) - lgamma((double)P + sum_alphas);
    double prod = 0.0;
    for (int i = 0; i < P; i++) {
        int dose = dosage[i];
        if (dose > 0) {
            double alpha_i = freqs ? freqs[g[i]] * scale : alpha_const;
            double num = lgamma((double)dose + alpha_i);
            double den = LGAMMA_INT[dose + 1] + lgamma(alpha_i);
            prod += num - den;
        }
    }
    return left + prod;
}

}  // namespace mchb
