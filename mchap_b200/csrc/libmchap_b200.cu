// libmchap_b200.cu — the C ABI (include/mchap_b200.h) over the sm_100a kernels.
// Unity build: the kernel headers are included here so that the __constant__ tables exist once.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//             -Xcompiler -fPIC -shared -cudart static
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include <algorithm>

#include "common.cuh"
#include "aux_kernels.cuh"
#include "assemble_kernel.cuh"
#include "exact_kernel.cuh"
#include "call_mcmc_kernel.cuh"
#include "tally_kernel.cuh"
#include "encode_kernel.cuh"

using namespace mchb;

#define MCHB_MAX_CHUNKS 64
// cuStreamWaitValue32(stream, device address, value, flags): flags 0 = wait until *addr >= value
typedef int (*wait_value32_fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct mchb_handle {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t cs = nullptr;       // chunked host path: device-to-host copies behind wait-value operations
    wait_value32_fn wait_value32 = nullptr;
    uint32_t *chunk_flags = nullptr;  // pinned, device-mapped host memory: one "chunk finished" flag per chunk
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr;
    std::string err;
    float kernel_ms = 0.f;
    int32_t launches = 0;
    int32_t host_chunks = 1;
    int32_t resident_warps = 0;   // warps the GPU holds at once in the most populated assemble launch of the last call
    int32_t resident_class_items = 0;
    int64_t last_trace_len = 0;  // int8 elements of the trace left in scratch by mchb_assemble_tally_batch
    int64_t last_call_trace_len = 0;  // int32 elements left by mchb_call_mcmc_tally_batch
    std::vector<DevBuf> bufs;  // scratch slots, grown on demand
    std::vector<cudaEvent_t> evs;  // events of the piecewise host-to-device pipelines, created on demand
};

namespace {

enum Slot {
    S_ITEMS = 0, S_ORDER, S_READS, S_COUNTS, S_NALLELES, S_INITIAL, S_OUT_G, S_OUT_L, S_RESULTS,
    S_WORDS, S_SEEDS, S_STREAM, S_BREAKS, S_BREAKLEN, S_TEMPS, S_COUNTER, S_GENO, S_AUX0, S_AUX1,
    S_HAPS, S_FREQS, S_SCRATCH, S_INIT32, S_OUT_A32, S_OUT_A, S_OUT_S, S_OUT_F, S_OUT_O, S_OUT_C, S_OUT_GL, S_OUT_GP, S_LLKS, S_CHUNKS,
    S_TITEMS, S_TGENO, S_TSTATES, S_TCOUNTS, S_TFIRST, S_TRESULTS,
    S_EITEMS, S_ECALLS, S_EPROBS, S_ENALL, S_EREADS, S_ECOUNTS, S_ERESULTS,
    S_BACKING0, S_BACKING1, S_BACKING2, S_BACKING3, S_BACKING4, S_BACKING5, S_BACKING6, S_BACKING7, S_BACKING8, S_BACKING9, S_BACKING10, S_BACKING11, S_BACKING12, S_BACKING13,
    S_RTBACK0, S_RTBACK1, S_RTBACK2, S_RTBACK3, S_RTBACK4, S_RTBACK5, S_RTBACK6, S_RTBACK7, S_RTBACK8, S_RTBACK9, S_RTBACK10, S_RTBACK11, S_RTBACK12, S_RTBACK13,
    S_QBACK0, S_QBACK1, S_QBACK2, S_QBACK3, S_QBACK4, S_QBACK5, S_QBACK6, S_QBACK7, S_QBACK8, S_QBACK9, S_QBACK10, S_QBACK11, S_QBACK12, S_QBACK13,
    S_NSLOTS
};

#define CK(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
            return MCHB_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

int ensure(mchb_handle *h, int slot, size_t bytes, void **out) {
    DevBuf &b = h->bufs[slot];
    if (bytes > b.cap) {
        if (b.p) CK(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
        size_t cap = bytes + bytes / 4 + 256;
        CK(cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    *out = b.p;
    return MCHB_OK;
}

// bring a bulk input array to the device (or pass a device pointer through)
template <typename T>
int stage_in(mchb_handle *h, int mem, int slot, const T *src, int64_t n, const T **dev) {
    if (!src || n <= 0) {
        *dev = nullptr;
        return MCHB_OK;
    }
    if (mem == MCHB_MEM_DEVICE) {
        *dev = src;
        return MCHB_OK;
    }
    void *p;
    int rc = ensure(h, slot, sizeof(T) * (size_t)n, &p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(p, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    *dev = static_cast<const T *>(p);
    return MCHB_OK;
}

template <typename T>
int stage_out(mchb_handle *h, int mem, int slot, T *dst, int64_t n, T **dev) {
    if (mem == MCHB_MEM_DEVICE) {
        *dev = dst;
        return MCHB_OK;
    }
    void *p;
    int rc = ensure(h, slot, sizeof(T) * (size_t)std::max<int64_t>(n, 1), &p);
    if (rc) return rc;
    *dev = static_cast<T *>(p);
    return MCHB_OK;
}

bool g_tables_ready[64] = {false};

int init_tables(mchb_handle *h) {
    if (h->device < 64 && g_tables_ready[h->device]) return MCHB_OK;
    static double log_int[MCHB_TABLE_N], log_inv[MCHB_TABLE_N], lgam[MCHB_TABLE_N];
    static float logf_int[MCHB_TABLE_N];
    for (int k = 0; k < MCHB_TABLE_N; k++) {
        log_int[k] = std::log((double)k);  // log(0) = -inf
        log_inv[k] = k > 0 ? std::log(1.0 / (double)k) : INFINITY;
        lgam[k] = k > 0 ? std::lgamma((double)k) : INFINITY;
        logf_int[k] = ::logf((float)k);
    }
    CK(cudaMemcpyToSymbol(LOG_INT, log_int, sizeof(log_int)));
    CK(cudaMemcpyToSymbol(LOG_INV_INT, log_inv, sizeof(log_inv)));
    CK(cudaMemcpyToSymbol(LGAMMA_INT, lgam, sizeof(lgam)));
    CK(cudaMemcpyToSymbol(LOGF_INT, logf_int, sizeof(logf_int)));
    static double log_ratio[17 * 17];
    for (int i = 0; i < 17; i++)
        for (int j = 0; j < 17; j++) log_ratio[i * 17 + j] = (i > 0 && j > 0) ? std::log((double)i / (double)j) : 0.0;
    CK(cudaMemcpyToSymbol(LOG_RATIO, log_ratio, sizeof(log_ratio)));
    if (h->device < 64) g_tables_ready[h->device] = true;
    return MCHB_OK;
}

void begin_call(mchb_handle *h) {
    h->err.clear();
    h->kernel_ms = 0.f;
    h->launches = 0;
    h->host_chunks = 1;
    h->resident_warps = 0;
    h->resident_class_items = 0;
}

}  // namespace

extern "C" {

int mchb_create(int device, mchb_handle **out) {
    if (!out) return MCHB_ERR_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return MCHB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MCHB_ERR_NO_DEVICE;
    if (prop.major != 10) return MCHB_ERR_NO_DEVICE;  // sm_100a code only: no fallback of any kind
    mchb_handle *h = new mchb_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->bufs.resize(S_NSLOTS);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) {
        delete h;
        return MCHB_ERR_CUDA;
    }
    int rc = init_tables(h);
    if (rc) {
        delete h;
        return rc;
    }
    *out = h;
    return MCHB_OK;
}

void mchb_destroy(mchb_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto &b : h->bufs)
        if (b.p) cudaFree(b.p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->cs) cudaStreamDestroy(h->cs);
    if (h->chunk_flags) cudaFreeHost(h->chunk_flags);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *mchb_last_error(const mchb_handle *h) { return h ? h->err.c_str() : "null handle"; }

void mchb_get_limits(mchb_limits *out) {
    if (!out) return;
    out->max_ploidy = MCHB_MAX_PLOIDY;
    out->max_key_bits = 64;
    out->max_unique_reads = MCHB_MAX_READS;
    out->max_temperatures = MCHB_MAX_TEMPS;
    out->max_haplotypes = 256;
}

float mchb_last_kernel_ms(const mchb_handle *h) { return h ? h->kernel_ms : 0.f; }
int32_t mchb_last_kernel_launches(const mchb_handle *h) { return h ? h->launches : 0; }
int32_t mchb_last_host_chunks(const mchb_handle *h) { return h ? h->host_chunks : 0; }
void *mchb_stream(const mchb_handle *h) { return h ? (void *)h->stream : nullptr; }
int mchb_sm_count(const mchb_handle *h) { return h ? h->sm_count : 0; }
int32_t mchb_last_resident_warps(const mchb_handle *h) { return h ? h->resident_warps : 0; }

int mchb_host_alloc(mchb_handle *h, int64_t bytes, void **out) {
    if (!h || !out || bytes < 0) return MCHB_ERR_ARGUMENT;
    CK(cudaSetDevice(h->device));
    *out = nullptr;
    CK(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable));
    return MCHB_OK;
}

int mchb_host_free(mchb_handle *h, void *p) {
    if (!h) return MCHB_ERR_ARGUMENT;
    if (p) CK(cudaFreeHost(p));
    return MCHB_OK;
}

// ------------------------------------------------------------------------------------- RNG
// One pre-generated stream per DISTINCT seed of the call (the CLIs seed every fit identically: one
// stream of 1-3 MB serves the whole batch).  Per-item seeds multiply that: the total is bounded here,
// loudly, instead of failing in cudaMalloc — callers with many distinct seeds split their batch
// (the Python layer does: DenovoMCMC / CallingMCMC `seeds=`).
#ifndef MCHB_STREAM_BUDGET
#define MCHB_STREAM_BUDGET (16ull << 30)
#endif
static int fill_streams(mchb_handle *h, const std::vector<uint32_t> &seeds, int64_t len, uint32_t **words) {
    if (len <= 0 || len > 0x7fffffffll) {
        h->err = "MT19937 stream of " + std::to_string(len) + " words per seed is out of range";
        return MCHB_ERR_ARGUMENT;
    }
    if ((unsigned long long)seeds.size() * (unsigned long long)len * 4ull > MCHB_STREAM_BUDGET) {
        h->err = std::to_string(seeds.size()) + " distinct seeds x " + std::to_string(len) +
                 " words exceed the stream budget: split the batch (fewer distinct seeds per call)";
        return MCHB_ERR_ARGUMENT;
    }
    void *dseeds, *dwords;
    int rc = ensure(h, S_SEEDS, sizeof(uint32_t) * seeds.size(), &dseeds);
    if (rc) return rc;
    rc = ensure(h, S_WORDS, sizeof(uint32_t) * seeds.size() * (size_t)len, &dwords);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dseeds, seeds.data(), sizeof(uint32_t) * seeds.size(), cudaMemcpyHostToDevice, h->stream));
    mt19937_fill_kernel<<<(unsigned)seeds.size(), 256, 0, h->stream>>>((const uint32_t *)dseeds, (uint32_t *)dwords, len);
    CK(cudaGetLastError());
    h->launches++;
    *words = (uint32_t *)dwords;
    return MCHB_OK;
}

int mchb_mt19937_words(mchb_handle *h, int mem, uint32_t seed, uint32_t *out, int64_t n) {
    if (!h || !out || n < 0) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n == 0) return MCHB_OK;
    std::vector<uint32_t> seeds(1, seed);
    uint32_t *words;
    int rc = fill_streams(h, seeds, n, &words);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, words, sizeof(uint32_t) * (size_t)n,
                       mem == MCHB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCHB_OK;
}

// ------------------------------------------------------------------- profiling-build counters
int mchb_debug_counters(mchb_handle *h, uint64_t *out, int32_t n, int32_t reset) {
    if (!h || !out || n < 0) return MCHB_ERR_ARGUMENT;
    CK(cudaSetDevice(h->device));
    for (int i = 0; i < n; i++) out[i] = 0;
#ifdef MCHB_PROFILE
    unsigned long long tmp[8 * 16];
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyFromSymbol(tmp, g_asm_prof, sizeof(tmp)));
    for (int i = 0; i < n && i < 8 * 16; i++) out[i] = tmp[i];
    if (reset) {
        memset(tmp, 0, sizeof(tmp));
        CK(cudaMemcpyToSymbol(g_asm_prof, tmp, sizeof(tmp)));
    }
    return MCHB_OK;
#else
    (void)reset;
    h->err = "library built without -DMCHB_PROFILE";
    return MCHB_ERR_ARGUMENT;
#endif
}

// --------------------------------------------------------------------------- FP64 pipe probe
int mchb_measure_fp64_peak(mchb_handle *h, double *out_tflops) {
    if (!h || !out_tflops) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    const int threads = 256, blocks = h->sm_count * 8, iters = 1 << 15;
    void *p;
    int rc = ensure(h, S_AUX0, sizeof(double) * (size_t)threads * blocks, &p);
    if (rc) return rc;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(h->ev0, h->stream));
        fp64_peak_kernel<<<blocks, threads, 0, h->stream>>>((double *)p, iters, 1.0000001, 1e-9);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev1, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->launches++;
        if (rep > 0 && ms < best) best = ms;
    }
    h->kernel_ms = best;
    *out_tflops = (double)threads * blocks * (double)iters * 8.0 * 2.0 / ((double)best * 1e-3) / 1e12;
    return MCHB_OK;
}

// ----------------------------------------------------------------------------- rank / unrank
int mchb_genotype_rank(mchb_handle *h, int mem, const int64_t *alleles, int64_t n, int32_t ploidy, int64_t *out_index) {
    if (!h || !alleles || !out_index || n < 0 || ploidy < 1) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n == 0) return MCHB_OK;
    const int64_t *din;
    int64_t *dout;
    int rc = stage_in(h, mem, S_AUX0, alleles, n * ploidy, &din);
    if (rc) return rc;
    rc = stage_out(h, mem, S_AUX1, out_index, n, &dout);
    if (rc) return rc;
    rank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(din, n, ploidy, dout);
    CK(cudaGetLastError());
    h->launches++;
    if (mem == MCHB_MEM_HOST)
        CK(cudaMemcpyAsync(out_index, dout, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCHB_OK;
}

int mchb_genotype_unrank(mchb_handle *h, int mem, const int64_t *index, int64_t n, int32_t ploidy, int64_t *out_alleles) {
    if (!h || !index || !out_alleles || n < 0 || ploidy < 1) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n == 0) return MCHB_OK;
    const int64_t *din;
    int64_t *dout;
    int rc = stage_in(h, mem, S_AUX0, index, n, &din);
    if (rc) return rc;
    rc = stage_out(h, mem, S_AUX1, out_alleles, n * ploidy, &dout);
    if (rc) return rc;
    unrank_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(din, n, ploidy, dout);
    CK(cudaGetLastError());
    h->launches++;
    if (mem == MCHB_MEM_HOST)
        CK(cudaMemcpyAsync(out_alleles, dout, sizeof(int64_t) * (size_t)n * ploidy, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MCHB_OK;
}

// ------------------------------------------------------------------------------ K1 llk batch
int mchb_log_likelihood_batch(mchb_handle *h, int mem, const mchb_llk_item *items, int64_t n_items,
                              const double *reads, int64_t reads_len, const int64_t *counts,
                              int64_t counts_len, const int8_t *genotypes, int64_t genotypes_len,
                              double *out_llk) {
    if (!h || !items || !out_llk || n_items < 0) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_llk_item &it = items[i];
        int64_t rsz = (int64_t)it.n_reads * it.n_pos * it.max_allele;
        if (it.n_reads < 0 || it.n_pos < 0 || it.max_allele < 0 || it.ploidy < 0 || it.reads_off < 0 ||
            it.reads_off + rsz > reads_len || it.geno_off < 0 ||
            it.geno_off + (int64_t)it.ploidy * it.n_pos > genotypes_len ||
            (counts && (it.counts_off < 0 || it.counts_off + it.n_reads > counts_len))) {
            h->err = "llk item " + std::to_string(i) + " exceeds the given array lengths";
            return MCHB_ERR_ARGUMENT;
        }
    }
    void *ditems;
    int rc = ensure(h, S_ITEMS, sizeof(mchb_llk_item) * (size_t)n_items, &ditems);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_llk_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    const double *dreads;
    const int64_t *dcounts;
    const int8_t *dgeno;
    double *dout;
    if ((rc = stage_in(h, mem, S_READS, reads, reads_len, &dreads))) return rc;
    if ((rc = stage_in(h, mem, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, mem, S_GENO, genotypes, genotypes_len, &dgeno))) return rc;
    if ((rc = stage_out(h, mem, S_OUT_L, out_llk, n_items, &dout))) return rc;
    CK(cudaEventRecord(h->ev0, h->stream));
    llk_batch_kernel<<<(unsigned)((n_items + 3) / 4), 128, 0, h->stream>>>((const mchb_llk_item *)ditems, n_items, dreads,
                                                                            dcounts, dgeno, dout);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    if (mem == MCHB_MEM_HOST)
        CK(cudaMemcpyAsync(out_llk, dout, sizeof(double) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->kernel_ms, h->ev0, h->ev1));
    return MCHB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------- K2 assemble
namespace {

struct AsmGeom {
    int nmax = 1, amax = 1, pmax = 1, tmax = 1, maxopt = 1;
};

// per-warp shared memory layout: fills the byte offsets in args, returns the region size
size_t asm_layout(const AsmGeom &g, int ch, int tres, AsmArgs &args) {
    const size_t upad = (size_t)ch * 32;
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
        off = (off + align - 1) & ~(align - 1);
        size_t o = off;
        off += bytes;
        return (int32_t)o;
    };
    // Rt at 0 in the one-chunk kernels; the kernels of larger items keep it in global memory (L2)
    take(ch >= 2 ? 0 : (size_t)g.nmax * g.amax * upad * 8, 16);
    args.o_cnt = take(upad * 8, 8);
    args.o_q = take(MCHB_ASM_Q_GLOBAL(ch) ? 0 : ((size_t)tres * g.pmax + 2) * upad * 8, 8);  // + 2 spare rows
    args.o_dist = take((size_t)g.nmax * g.amax * 8, 8);
    args.o_oll = take((size_t)(g.maxopt + 1) * 8, 8);
    args.o_opr = take((size_t)(g.maxopt + 1) * 8, 8);
    args.o_lgdisp = take((size_t)(g.pmax + 2) * 8, 8);
    args.o_homlp = take((size_t)std::max(g.amax, g.pmax) * 8, 8);  // also the prior's row scratch
    args.o_llk_t = take((size_t)g.tmax * 8, 8);
    args.o_key = take((size_t)tres * g.pmax * 8, 8);
    args.o_sc = take((size_t)SC_COUNT * 8, 8);
    args.o_perm = take((size_t)g.pmax * g.nmax * 2, 2);
    args.o_het = take((size_t)g.nmax, 1);
    args.o_fixa = take((size_t)g.nmax, 1);
    args.o_nall = take((size_t)g.nmax, 1);
    args.o_opt0 = take((size_t)g.maxopt + 1, 1);
    args.o_opt1 = take((size_t)g.maxopt + 1, 1);
    args.o_ivb = take((size_t)g.nmax + 2, 1);
    args.o_ivp = take((size_t)g.nmax + 2, 1);
    args.o_ring = take(128 * 4, 4);
    args.o_q32 = take(((size_t)tres * g.pmax + 2) * upad * 4, 4);
    args.o_rat = take((size_t)g.nmax * (MCHB_ASM_RAT_HALF(ch) ? 1 : 2) * upad * 4, 4);
    args.o_c32 = take(upad * 4, 4);
    args.o_rpc = take((size_t)tres * upad * 4, 4);
    args.o_bcs = take((size_t)(g.maxopt + 1) * 8, 8);
    args.o_epoch = take((size_t)tres * 2 * 4, 8);
    args.o_mcache = take((size_t)tres * 2 * g.pmax * g.nmax * 8, 8);
    {
        const int tri = g.nmax * (g.nmax + 1) / 2;
        const int direct_max = MCHB_SCACHE_DIRECT_MAX(ch);
        args.scache_tri = 2 * tri <= direct_max ? tri : 0;
        args.scache_n = args.scache_tri > 0 ? 2 * tri : MCHB_SCACHE_HASH_N(ch);
    }
    args.o_scache = take((size_t)tres * args.scache_n * sizeof(ScEntry), 8);
    args.o_wmap = take((size_t)g.pmax * g.nmax * 2, 2);
    args.o_inv = take(32 * 4, 4);
    args.o_hot = take((size_t)g.tmax * 4 * 4, 4);
    args.o_hmap = take((size_t)g.pmax * g.nmax * 2, 2);
    args.o_rank = take(16, 1);
    return (off + 15) & ~(size_t)15;
}

template <int CH, bool PRIOR>
int launch_assemble(mchb_handle *h, cudaStream_t stream, AsmArgs &args, const AsmGeom &g, int n_items_class,
                    bool overlap_previous, int backing_slot) {
    // all temperatures' state slots resident — unless that leaves fewer than four warps per CTA:
    // then one resident slot, the others swapped through a backing store in global memory
    int tres = g.tmax;
    size_t per_warp = asm_layout(g, CH, tres, args);
    if (CH >= 2 && g.tmax > 1) {
        // one resident slot when all of them do not fit four warps, or (three and more chunks: a
        // handful of warps per SM at best) when swapping lets an SM hold 1.5 times as many warps
        AsmArgs tmp = args;
        const size_t per_warp_swap = asm_layout(g, CH, 1, tmp);
        const size_t room = (size_t)h->smem_optin;
        const size_t w_full = std::min<size_t>(16, room / per_warp), w_swap = std::min<size_t>(16, room / per_warp_swap);
        if (per_warp * 4 > room || (CH >= 3 && 2 * w_swap >= 3 * w_full)) {
            tres = 1;
            per_warp = asm_layout(g, CH, tres, args);
        }
    }
    // warps per CTA: the value that keeps most warps resident per SM (shared memory and registers
    // both count: the occupancy query knows the kernel's register allocation); ties go to 4, then
    // to the larger CTA
    constexpr int max_warps = MCHB_ASM_MAXTHREADS(CH) / 32;
    int warps_per_cta = 0, best = 0, ctas_per_sm = 0;
    for (int w = max_warps; w >= 1; w--) {
        if (per_warp * w > (size_t)h->smem_optin) continue;
        CK(cudaFuncSetAttribute(assemble_kernel<CH, PRIOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * w)));
        int blocks = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, assemble_kernel<CH, PRIOR>, w * 32, per_warp * w));
        const int resident = blocks * w;
        if (resident > best || (resident == best && w == 4)) {
            best = resident;
            warps_per_cta = w;
            ctas_per_sm = blocks;
        }
    }
    if (warps_per_cta == 0 || ctas_per_sm < 1) {
        h->err = "assemble item needs more shared memory than one CTA can have";
        return MCHB_ERR_ARGUMENT;
    }
    const size_t smem = per_warp * warps_per_cta;
    CK(cudaFuncSetAttribute(assemble_kernel<CH, PRIOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_items_class > h->resident_class_items) {
        h->resident_class_items = n_items_class;
        h->resident_warps = (int32_t)(h->sm_count * ctas_per_sm * warps_per_cta);
    }
    long long want = ((long long)n_items_class + warps_per_cta - 1) / warps_per_cta;
    long long grid = std::min<long long>(want, (long long)h->sm_count * ctas_per_sm);
    if (grid < 1) grid = 1;
    args.nmax = g.nmax;
    args.amax = g.amax;
    args.pmax = g.pmax;
    args.tmax = g.tmax;
    args.maxopt = g.maxopt;
    args.smem_per_warp = (int)per_warp;
    args.tres = tres;
    args.slot_bytes = 0;
    args.slot_backing = nullptr;
    if (tres < g.tmax) {
        const size_t upad = (size_t)CH * 32;
        auto r8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
        const size_t slot_bytes = (r8(MCHB_ASM_Q_GLOBAL(CH) ? 0 : g.pmax * upad * 8) + r8((size_t)2 * g.pmax * g.nmax * 8) + r8((size_t)g.pmax * 8) +
                                   r8((size_t)args.scache_n * sizeof(ScEntry)) + r8(g.pmax * upad * 4) + r8(upad * 4) + r8(8) + 15) &
                                  ~(size_t)15;
        void *back;
        int rc = ensure(h, backing_slot, slot_bytes * (size_t)g.tmax * (size_t)grid * warps_per_cta, &back);
        if (rc) return rc;
        args.slot_bytes = (int32_t)slot_bytes;
        args.slot_backing = (unsigned char *)back;
    }
    args.rt_backing = nullptr;
    args.rt_stride = 0;
    if (CH >= 2) {
        const size_t rt_elems = (size_t)g.nmax * g.amax * CH * 32;
        void *back;
        int rc = ensure(h, backing_slot - S_BACKING0 + S_RTBACK0, rt_elems * 8 * (size_t)grid * warps_per_cta, &back);
        if (rc) return rc;
        args.rt_backing = (double *)back;
        args.rt_stride = (int64_t)rt_elems;
    }
    args.q_backing = nullptr;
    args.q_stride = 0;
    if (MCHB_ASM_Q_GLOBAL(CH)) {
        const size_t q_elems = ((size_t)g.tmax * g.pmax + 2) * CH * 32;
        void *back;
        int rc = ensure(h, backing_slot - S_BACKING0 + S_QBACK0, q_elems * 8 * (size_t)grid * warps_per_cta, &back);
        if (rc) return rc;
        args.q_backing = (double *)back;
        args.q_stride = (int64_t)q_elems;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(warps_per_cta * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap_previous ? 1 : 0;  // start once the previous kernel's CTAs are all resident
    CK(cudaLaunchKernelEx(&cfg, assemble_kernel<CH, PRIOR>, args));
    CK(cudaGetLastError());
    h->launches++;
    return MCHB_OK;
}

}  // namespace

extern "C" {

int mchb_assemble_batch(mchb_handle *h, int mem, const mchb_assemble_params *params,
                        const mchb_assemble_item *items, int64_t n_items, const double *reads,
                        int64_t reads_len, const int64_t *counts, int64_t counts_len,
                        const int8_t *n_alleles, int64_t n_alleles_len, const int8_t *initial,
                        int64_t initial_len, int8_t *out_genotypes, int64_t out_genotypes_len,
                        double *out_llks, int64_t out_llks_len, mchb_item_result *results) {
    if (!h || !params || !items || !results || n_items < 0 || !out_genotypes || !out_llks) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    if (n_items > 0x7fffffff) {
        h->err = "too many items in one call";
        return MCHB_ERR_ARGUMENT;
    }
    const mchb_assemble_params &pp = *params;
    if (pp.steps < 0 || pp.chains < 0 || !pp.break_table || !pp.break_len || pp.break_rows < 1 || !pp.temperatures) {
        h->err = "bad assemble parameters";
        return MCHB_ERR_ARGUMENT;
    }
    constexpr int NCLS = 14;
    static const int CLS_CH[7] = {1, 2, 3, 4, 8, 16, 32};
    // ---- validate, classify by unique-read chunk count and prior use, collect seeds
    std::vector<int32_t> order[NCLS];  // class = 2 * (index of CH in {1, 2, 3, 4, 8, 16, 32}) + has_prior
    AsmGeom geom[NCLS];
    std::map<uint32_t, int32_t> seed_index;
    std::vector<uint32_t> seeds;
    std::vector<int32_t> item_stream((size_t)n_items, 0);
    int64_t words_needed = 0;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_assemble_item &it = items[i];
        const int64_t rsz = (int64_t)it.n_reads * it.n_pos * it.max_allele;
        const int64_t gsz = (int64_t)pp.chains * pp.steps * it.ploidy * it.n_pos;
        bool bad = it.n_reads < 0 || it.n_pos < 0 || it.max_allele < 1 || it.ploidy < 1 || it.n_temps < 1 ||
                   it.reads_off < 0 || it.reads_off + rsz > reads_len || it.nalleles_off < 0 ||
                   it.nalleles_off + it.n_pos > n_alleles_len || it.genotypes_off < 0 ||
                   it.genotypes_off + gsz > out_genotypes_len || it.llks_off < 0 ||
                   it.llks_off + (int64_t)pp.chains * pp.steps > out_llks_len || it.temps_off < 0 ||
                   it.temps_off + it.n_temps > pp.temperatures_len ||
                   (counts && it.n_reads > 0 && (it.counts_off < 0 || it.counts_off + it.n_reads > counts_len)) ||
                   (initial && it.initial_off >= 0 &&
                    it.initial_off + (int64_t)pp.chains * it.ploidy * it.initial_nhet > initial_len);
        if (bad) {
            h->err = "assemble item " + std::to_string(i) + " exceeds the given array lengths";
            return MCHB_ERR_ARGUMENT;
        }
        results[i].status = MCHB_ITEM_OK;
        results[i].n_het = 0;
        results[i].rng_words = 0;
        results[i].llk_evals = 0;
        if (it.n_reads > MCHB_MAX_READS || it.ploidy > MCHB_MAX_PLOIDY || it.n_temps > MCHB_MAX_TEMPS || it.n_pos > 255 ||
            it.max_allele > 255) {
            results[i].status = MCHB_ITEM_UNSUPPORTED;
            continue;
        }
        if (it.n_pos == 0) {  // nothing to sample: empty traces, NaN llks are written by the host shim
            continue;
        }
        const int chi = it.n_reads <= 32 ? 0 : it.n_reads <= 64 ? 1 : it.n_reads <= 96 ? 2 : it.n_reads <= 128 ? 3 :
                        it.n_reads <= 256 ? 4 : it.n_reads <= 512 ? 5 : 6;
        {
            // an item whose own tables exceed one CTA's shared memory is reported as unsupported
            // (per item: it must not take the rest of the batch with it)
            AsmGeom gi;
            gi.nmax = it.n_pos;
            gi.amax = it.max_allele;
            gi.pmax = it.ploidy;
            gi.tmax = it.n_temps;
            gi.maxopt = std::max({gi.pmax * (gi.pmax - 1), gi.amax, (int)pp.break_stride, 2});
            AsmArgs tmp;
            size_t pw = asm_layout(gi, CLS_CH[chi], gi.tmax, tmp);
            if (pw > (size_t)h->smem_optin && gi.tmax > 1 && CLS_CH[chi] >= 2) pw = asm_layout(gi, CLS_CH[chi], 1, tmp);
            if (pw > (size_t)h->smem_optin) {
                results[i].status = MCHB_ITEM_UNSUPPORTED;
                continue;
            }
        }
        int cls = 2 * chi + (std::isnan(it.inbreeding) ? 0 : 1);
        order[cls].push_back((int32_t)i);
        AsmGeom &g = geom[cls];
        g.nmax = std::max(g.nmax, it.n_pos);
        g.amax = std::max(g.amax, it.max_allele);
        g.pmax = std::max(g.pmax, it.ploidy);
        g.tmax = std::max(g.tmax, it.n_temps);
        if (!pp.replay_words) {
            auto f = seed_index.find(it.seed);
            if (f == seed_index.end()) {
                f = seed_index.emplace(it.seed, (int32_t)seeds.size()).first;
                seeds.push_back(it.seed);
            }
            item_stream[(size_t)i] = f->second;
        }
        const int64_t pn = (int64_t)it.ploidy * it.n_pos;
        const int64_t per_step = (4 * pn + 10 * (int64_t)it.n_pos + 24) * it.n_temps;
        words_needed = std::max(words_needed, (int64_t)pp.chains * (2 * pn + 8 + (int64_t)pp.steps * per_step) + 64);
    }
    for (int c = 0; c < NCLS; c++) {
        AsmGeom &g = geom[c];
        g.maxopt = std::max({g.pmax * (g.pmax - 1), g.amax, (int)pp.break_stride, 2});
        // a class is launched with the largest dimensions of its items: if that combination does not
        // fit one CTA (each item alone does), the class is run item by item is not worth the code for
        // a case this rare — its items are reported as unsupported instead of failing the call
        if (order[c].empty()) continue;
        AsmArgs tmp;
        size_t pw = asm_layout(g, CLS_CH[c / 2], g.tmax, tmp);
        if (pw > (size_t)h->smem_optin && g.tmax > 1 && CLS_CH[c / 2] >= 2) pw = asm_layout(g, CLS_CH[c / 2], 1, tmp);
        if (pw > (size_t)h->smem_optin) {
            for (int32_t id : order[c]) results[id].status = MCHB_ITEM_UNSUPPORTED;
            order[c].clear();
        }
    }
    // ---- device copies of the small tables
    void *ditems, *dstream, *dbreaks, *dbreaklen, *dtemps, *dcounter, *dresults;
    int rc;
    if ((rc = ensure(h, S_ITEMS, sizeof(mchb_assemble_item) * (size_t)n_items, &ditems))) return rc;
    if ((rc = ensure(h, S_STREAM, sizeof(int32_t) * (size_t)n_items, &dstream))) return rc;
    if ((rc = ensure(h, S_BREAKS, sizeof(double) * (size_t)pp.break_rows * pp.break_stride, &dbreaks))) return rc;
    if ((rc = ensure(h, S_BREAKLEN, sizeof(int32_t) * (size_t)pp.break_rows, &dbreaklen))) return rc;
    if ((rc = ensure(h, S_TEMPS, sizeof(double) * (size_t)pp.temperatures_len, &dtemps))) return rc;
    if ((rc = ensure(h, S_COUNTER, sizeof(int32_t) * 16, &dcounter))) return rc;
    if ((rc = ensure(h, S_RESULTS, sizeof(mchb_item_result) * (size_t)n_items, &dresults))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_assemble_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dstream, item_stream.data(), sizeof(int32_t) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dbreaks, pp.break_table, sizeof(double) * (size_t)pp.break_rows * pp.break_stride,
                       cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dbreaklen, pp.break_len, sizeof(int32_t) * (size_t)pp.break_rows, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dtemps, pp.temperatures, sizeof(double) * (size_t)pp.temperatures_len, cudaMemcpyHostToDevice,
                       h->stream));
    CK(cudaMemcpyAsync(dresults, results, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    // ---- bulk arrays
    const double *dreads;
    const int64_t *dcounts;
    const int8_t *dnall, *dinit;
    int8_t *dog;
    double *dol;
    if ((rc = stage_in(h, mem, S_READS, reads, reads_len, &dreads))) return rc;
    if ((rc = stage_in(h, mem, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, mem, S_NALLELES, n_alleles, n_alleles_len, &dnall))) return rc;
    if ((rc = stage_in(h, mem, S_INITIAL, initial, initial_len, &dinit))) return rc;
    if (mem == MCHB_MEM_HOST) h->last_trace_len = 0;  // the scratch trace of an earlier tally call is overwritten
    if ((rc = stage_out(h, mem, S_OUT_G, out_genotypes, out_genotypes_len, &dog))) return rc;
    if ((rc = stage_out(h, mem, S_OUT_L, out_llks, out_llks_len, &dol))) return rc;

    // ---- word streams + launches (retry with longer streams if an item runs out of words)
    int64_t stream_len = pp.rng_words_hint > 0 ? pp.rng_words_hint : words_needed;
    if (pp.replay_words) stream_len = pp.replay_len;
    stream_len = (stream_len + 31) & ~(int64_t)31;
    if (stream_len > 0x7fffff00LL) {
        h->err = "random word stream longer than 2^31 words: split the batch into fewer steps per call";
        return MCHB_ERR_ARGUMENT;
    }
    std::vector<int32_t> todo[NCLS];
    for (int c = 0; c < NCLS; c++) todo[c] = order[c];
    // ---- host buffers: the batch is cut into chunks of consecutive item ids.  Every launch counts
    // the items it finishes per chunk (device counters); the warp that finishes the last item of a
    // chunk raises the chunk's flag in device-mapped host memory, and the copy stream waits on that
    // flag with a stream memory operation (cuStreamWaitValue32) before it moves the chunk's traces
    // to the host.  The device-to-host transfer of the (large) trace so hides behind the sampling
    // of the later chunks, with one launch per class and no tail between chunks.  (Waiting on the
    // device counters themselves was measured not to release before the kernels end.)
    int n_chunks = 1;
    int64_t chunk_items = n_items;
    if (mem == MCHB_MEM_HOST && !pp.replay_words) {
        const int64_t out_bytes = out_genotypes_len + 8 * out_llks_len;
        n_chunks = (int)std::min<int64_t>({(int64_t)MCHB_MAX_CHUNKS, out_bytes >> 28, n_items / 1024});
        if (const char *e = getenv("MCHB_HOST_CHUNKS"))  // test hook: force the chunk count
            n_chunks = (int)std::min<int64_t>({(int64_t)MCHB_MAX_CHUNKS, (int64_t)atoi(e), n_items});
        if (n_chunks < 1) n_chunks = 1;
        if (n_chunks > 1 && !h->wait_value32) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess || !fn) {
                (void)cudaGetLastError();
                n_chunks = 1;  // no stream memory operations: one copy after the kernels
            } else {
                h->wait_value32 = (wait_value32_fn)fn;
            }
        }
        if (n_chunks > 1 && !h->cs) {
            CK(cudaStreamCreateWithFlags(&h->cs, cudaStreamNonBlocking));
            CK(cudaHostAlloc((void **)&h->chunk_flags, sizeof(uint32_t) * MCHB_MAX_CHUNKS, cudaHostAllocMapped));
        }
        chunk_items = (n_items + n_chunks - 1) / n_chunks;
        n_chunks = (int)((n_items + chunk_items - 1) / chunk_items);
    }
    h->host_chunks = n_chunks;
    void *dchunk = nullptr;
    std::vector<uint32_t> chunk_skip((size_t)n_chunks, 0u);
    if (n_chunks > 1) {
        if ((rc = ensure(h, S_CHUNKS, sizeof(uint32_t) * (size_t)n_chunks, &dchunk))) return rc;
        // items that no launch handles count as finished from the start
        std::vector<uint8_t> launched((size_t)n_items, 0);
        for (int c = 0; c < NCLS; c++)
            for (int32_t id : order[c]) launched[(size_t)id] = 1;
        for (int64_t i = 0; i < n_items; i++)
            if (!launched[(size_t)i]) chunk_skip[(size_t)(i / chunk_items)]++;
        // no kernel of an earlier call is running any more: the flags can be written directly
        for (int k = 0; k < n_chunks; k++) {
            const int64_t size_k = std::min<int64_t>(n_items, chunk_items * (k + 1)) - chunk_items * k;
            h->chunk_flags[k] = chunk_skip[(size_t)k] >= (uint64_t)size_k ? 1u : 0u;
        }
    }
    int attempts_used = 0;
    for (int attempt = 0; attempt < 6; attempt++) {
        int64_t total = 0;
        for (int c = 0; c < NCLS; c++) total += (int64_t)todo[c].size();
        if (total == 0) break;
        attempts_used++;
        const bool piped = attempt == 0 && n_chunks > 1;
        uint32_t *dwords = nullptr;
        CK(cudaEventRecord(h->ev0, h->stream));  // the word-stream fill belongs to the timed device work
        if (pp.replay_words) {
            void *p;
            if ((rc = ensure(h, S_WORDS, sizeof(uint32_t) * (size_t)stream_len, &p))) return rc;
            CK(cudaMemsetAsync(p, 0, sizeof(uint32_t) * (size_t)stream_len, h->stream));
            CK(cudaMemcpyAsync(p, pp.replay_words, sizeof(uint32_t) * (size_t)pp.replay_len, cudaMemcpyHostToDevice,
                               h->stream));
            dwords = (uint32_t *)p;
        } else {
            if ((rc = fill_streams(h, seeds, stream_len, &dwords))) return rc;
        }
        void *dorder;
        if ((rc = ensure(h, S_ORDER, sizeof(int32_t) * (size_t)total, &dorder))) return rc;
        CK(cudaMemsetAsync(dcounter, 0, sizeof(int32_t) * 16, h->stream));
        if (piped)
            CK(cudaMemcpyAsync(dchunk, chunk_skip.data(), sizeof(uint32_t) * (size_t)n_chunks, cudaMemcpyHostToDevice,
                               h->stream));
        // rare classes first, the most populated class last, chained by programmatic dependent
        // launch (see assemble_kernel): the few long-running rare items overlap the main launch
        int main_cls = 0;
        for (int c = 1; c < NCLS; c++)
            if (todo[c].size() > todo[main_cls].size()) main_cls = c;
        int64_t offs[NCLS];
        {
            int64_t off = 0;
            for (int c = 0; c < NCLS; c++) {
                offs[c] = off;
                if (todo[c].empty()) continue;
                CK(cudaMemcpyAsync((int32_t *)dorder + off, todo[c].data(), sizeof(int32_t) * todo[c].size(),
                                   cudaMemcpyHostToDevice, h->stream));
                off += (int64_t)todo[c].size();
            }
        }
        if (piped) {
            CK(cudaEventRecord(h->ev_fork, h->stream));
            CK(cudaStreamWaitEvent(h->cs, h->ev_fork, 0));
        }
        bool chained = false;
        for (int pass = 0; pass < 2; pass++) {
            for (int c = NCLS - 1; c >= 0; c--) {
                if (todo[c].empty()) continue;
                if ((pass == 0) == (c == main_cls)) continue;  // pass 0: rare classes, pass 1: main class
                cudaStream_t st = h->stream;
                AsmArgs args;
                memset(&args, 0, sizeof(args));
                args.items = (const mchb_assemble_item *)ditems;
                args.order = (int32_t *)dorder + offs[c];
                args.n_order = (int32_t)todo[c].size();
                args.reads = dreads;
                args.counts = dcounts;
                args.n_alleles = dnall;
                args.initial = dinit;
                args.out_genotypes = dog;
                args.out_llks = dol;
                args.results = (mchb_item_result *)dresults;
                args.words = dwords;
                args.item_stream = (const int32_t *)dstream;
                args.stream_len = pp.replay_words ? pp.replay_len : stream_len;
                args.steps = pp.steps;
                args.chains = pp.chains;
                args.sort_recorded = pp.sort_haplotypes;
                args.fix_homozygous = pp.fix_homozygous;
                args.p_recomb = pp.p_recombination;
                args.p_partial = pp.p_partial_dosage;
                args.p_dosage = pp.p_dosage;
                args.break_table = (const double *)dbreaks;
                args.break_len = (const int32_t *)dbreaklen;
                args.break_rows = pp.break_rows;
                args.break_stride = pp.break_stride;
                args.temperatures = (const double *)dtemps;
                args.work_counter = (int32_t *)dcounter + c;
                args.chunk_done = piped ? (uint32_t *)dchunk : nullptr;
                args.chunk_items = (int32_t)chunk_items;
                args.n_items = (int32_t)n_items;
                if (piped) {
                    void *dflags = nullptr;
                    CK(cudaHostGetDevicePointer(&dflags, h->chunk_flags, 0));
                    args.chunk_flags = (uint32_t *)dflags;
                }
                const int n_c = (int)todo[c].size();
                switch (c) {
                    case 0: rc = launch_assemble<1, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 1: rc = launch_assemble<1, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 2: rc = launch_assemble<2, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 3: rc = launch_assemble<2, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 4: rc = launch_assemble<3, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 5: rc = launch_assemble<3, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 6: rc = launch_assemble<4, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 7: rc = launch_assemble<4, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 8: rc = launch_assemble<8, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 9: rc = launch_assemble<8, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 10: rc = launch_assemble<16, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 11: rc = launch_assemble<16, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    case 12: rc = launch_assemble<32, false>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                    default: rc = launch_assemble<32, true>(h, st, args, geom[c], n_c, chained, S_BACKING0 + c); break;
                }
                if (rc) return rc;
                chained = true;
            }
        }
        CK(cudaEventRecord(h->ev1, h->stream));
        // the chunk copies are queued after the launches, so that a pageable destination (whose
        // copies block the host) still overlaps with the kernels already in the streams
        if (piped) {
            const bool trace_chunks = getenv("MCHB_TRACE_CHUNKS") != nullptr;  // debugging aid: when did each copy end
            std::vector<cudaEvent_t> tev;
            for (int k = 0; k < n_chunks; k++) {
                const int64_t item_lo = chunk_items * k, item_hi = std::min<int64_t>(n_items, chunk_items * (k + 1));
                int64_t g_lo = INT64_MAX, g_hi = 0, l_lo = INT64_MAX, l_hi = 0;
                for (int64_t i = item_lo; i < item_hi; i++) {
                    const mchb_assemble_item &it = items[i];
                    const int64_t gsz = (int64_t)pp.chains * pp.steps * it.ploidy * it.n_pos;
                    g_lo = std::min(g_lo, it.genotypes_off);
                    g_hi = std::max(g_hi, it.genotypes_off + gsz);
                    l_lo = std::min(l_lo, it.llks_off);
                    l_hi = std::max(l_hi, it.llks_off + (int64_t)pp.chains * pp.steps);
                }
                void *dflags = nullptr;
                CK(cudaHostGetDevicePointer(&dflags, h->chunk_flags, 0));
                if (trace_chunks && k == 0) {
                    cudaEvent_t e;
                    CK(cudaEventCreate(&e));
                    CK(cudaEventRecord(e, h->cs));
                    tev.push_back(e);
                }
                if (h->wait_value32(h->cs, (unsigned long long)((uint32_t *)dflags + k), 1u, 0u)) {
                    h->err = "cuStreamWaitValue32 failed";
                    return MCHB_ERR_CUDA;
                }
                if (trace_chunks) {
                    cudaEvent_t e;
                    CK(cudaEventCreate(&e));
                    CK(cudaEventRecord(e, h->cs));
                    tev.push_back(e);
                }
                if (g_hi > g_lo)
                    CK(cudaMemcpyAsync(out_genotypes + g_lo, dog + g_lo, (size_t)(g_hi - g_lo), cudaMemcpyDeviceToHost, h->cs));
                if (l_hi > l_lo)
                    CK(cudaMemcpyAsync(out_llks + l_lo, dol + l_lo, sizeof(double) * (size_t)(l_hi - l_lo),
                                       cudaMemcpyDeviceToHost, h->cs));
                if (trace_chunks) {
                    cudaEvent_t e;
                    CK(cudaEventCreate(&e));
                    CK(cudaEventRecord(e, h->cs));
                    tev.push_back(e);
                }
            }
            CK(cudaStreamSynchronize(h->cs));
            if (trace_chunks) {
                CK(cudaStreamSynchronize(h->stream));
                float t1 = 0.f;
                CK(cudaEventElapsedTime(&t1, h->ev0, h->ev1));
                fprintf(stderr, "[mchb] kernels done at %.1f ms; chunk copies done at:", t1);
                for (cudaEvent_t e : tev) {
                    float t = 0.f;
                    CK(cudaEventElapsedTime(&t, h->ev0, e));
                    fprintf(stderr, " %.1f", t);
                    cudaEventDestroy(e);
                }
                fprintf(stderr, "\n");
            }
        }
        // (queued behind the chunk copies: a copy that waits for the kernels at the head of the
        // device-to-host queue would hold the chunk copies back)
        CK(cudaMemcpyAsync(results, dresults, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->kernel_ms += ms;
        if (pp.replay_words) break;
        // collect the items that ran out of words
        bool again = false;
        for (int c = 0; c < NCLS; c++) {
            std::vector<int32_t> next;
            for (int32_t id : todo[c])
                if (results[id].status == MCHB_ITEM_RNG_EXHAUSTED) next.push_back(id);
            todo[c].swap(next);
            again = again || !todo[c].empty();
        }
        if (!again) break;
        stream_len *= 2;
    }
    if (mem == MCHB_MEM_HOST && (n_chunks == 1 || attempts_used > 1)) {
        CK(cudaMemcpyAsync(out_genotypes, dog, (size_t)out_genotypes_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(out_llks, dol, sizeof(double) * (size_t)out_llks_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return MCHB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------ K4 call-exact
namespace {

struct CallGeom {
    int umax = 1, hmax = 1, pmax = 1, pmin = 1 << 30;
    long long gmax = 1;
};

// validates the descriptors; G per item must fit the per-genotype arrays when use_gl
int check_call_items(mchb_handle *h, const mchb_call_item *items, int64_t n_items, int64_t reads_len,
                     bool has_counts, int64_t counts_len, int64_t haps_len, bool has_freqs, int64_t freqs_len,
                     int64_t hap_out_len, int64_t gl_len, bool use_reads, CallGeom &g) {
    int memo_h = -1, memo_p = -1;
    long long memo_g = 0;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_call_item &it = items[i];
        bool bad = it.n_reads < 0 || it.n_pos < 0 || it.max_allele < 0 || it.ploidy < 1 || it.n_haps < 1 ||
                   it.ploidy > MCHB_MAX_PLOIDY || it.n_haps > 256;
        if (!bad && use_reads) {
            const int64_t rsz = (int64_t)it.n_reads * it.n_pos * it.max_allele;
            bad = it.reads_off < 0 || it.reads_off + rsz > reads_len || it.haps_off < 0 ||
                  it.haps_off + (int64_t)it.n_haps * it.n_pos > haps_len ||
                  (has_counts && (it.counts_off < 0 || it.counts_off + it.n_reads > counts_len));
        }
        if (!bad && has_freqs && it.freqs_off >= 0) bad = it.freqs_off + it.n_haps > freqs_len;
        if (!bad && hap_out_len >= 0) bad = it.hap_out_off < 0 || it.hap_out_off + it.n_haps > hap_out_len;
        long long G = 0;
        if (!bad) {
            if (it.n_haps == memo_h && it.ploidy == memo_p) {
                G = memo_g;  // batches repeat a few (H, P) pairs
            } else {
                // C(H+P-1, P) with overflow guard (long double estimate first)
                long double est = 1.0L;
                for (int k = 1; k <= it.ploidy; k++) est = est * (long double)(it.n_haps + it.ploidy - k) / (long double)k;
                if (est > 4.0e18L) bad = true;
                else {
                    G = comb_exact((long long)it.n_haps + it.ploidy - 1, it.ploidy);
                    memo_h = it.n_haps;
                    memo_p = it.ploidy;
                    memo_g = G;
                }
            }
        }
        if (!bad && gl_len >= 0) bad = it.gl_off < 0 || it.gl_off + G > gl_len;
        if (bad) {
            h->err = "call item " + std::to_string(i) + " is inconsistent with the given array lengths or limits";
            return MCHB_ERR_ARGUMENT;
        }
        g.umax = std::max(g.umax, it.n_reads);
        g.hmax = std::max(g.hmax, it.n_haps);
        g.pmax = std::max(g.pmax, it.ploidy);
        g.pmin = std::min(g.pmin, it.ploidy);
        g.gmax = std::max(g.gmax, G);
    }
    return MCHB_OK;
}

// threads of a CTA that own a partial row of allele statistics: as many as fit 32 KB
int exact_part_threads(const CallGeom &g) {
    int pt = 128;
    while (pt > 1 && (size_t)pt * ((2 * g.hmax) | 1) * 8 > 32768) pt >>= 1;
    return pt;
}

// host mirror of exact_carve()
size_t exact_smem(const CallGeom &g, int umax, int part_threads) {
    const size_t us = (size_t)(umax | 1), hs = (size_t)(g.hmax | 1);
    const size_t ps = (size_t)((2 * g.hmax) | 1);
    (void)hs;
    size_t d = (size_t)g.hmax * us + g.hmax + (size_t)g.hmax * (g.pmax + 1) + g.hmax + (size_t)part_threads * ps + 32 +
               (size_t)(g.hmax + 1) * (g.pmax + 1);
    return d * 8 + (size_t)umax * 8 + 64;
}

// Instantiations of exact_kernel: ploidies 2, 4, 6, 8 have kernels specialised for batches in which
// every item has that ploidy; anything else (mixed or odd ploidies, ploidy > 8, or no room for parked
// log joints) runs the generic 16-slot kernel.
#define MCHB_EXACT_DISPATCH(fixed_p, recomp, CALL)                  \
    do {                                                            \
        if (recomp) { CALL(16, false, true); }                      \
        else if ((fixed_p) == 2) { CALL(2, true, false); }          \
        else if ((fixed_p) == 4) { CALL(4, true, false); }          \
        else if ((fixed_p) == 6) { CALL(6, true, false); }          \
        else if ((fixed_p) == 8) { CALL(8, true, false); }          \
        else { CALL(16, false, false); }                            \
    } while (0)

// posterior_kernel: smallest compiled slot count >= pmax
#define MCHB_POSTERIOR_DISPATCH(pmax, CALL) \
    do {                                    \
        if ((pmax) <= 4) { CALL(4); }       \
        else if ((pmax) <= 8) { CALL(8); }  \
        else { CALL(16); }                  \
    } while (0)

// bytes of parked log joints the whole grid may hold (mchb_call_exact_mode_batch); beyond it the
// grid shrinks, and an item too large for a single row is evaluated twice instead
#ifndef MCHB_EXACT_SCRATCH_BUDGET
#define MCHB_EXACT_SCRATCH_BUDGET (2ull << 30)
#endif

int run_exact(mchb_handle *h, int mem, int mode, const mchb_call_item *items, int64_t n_items, const double *reads,
              int64_t reads_len, const int64_t *counts, int64_t counts_len, const int8_t *haplotypes,
              int64_t haplotypes_len, const double *freqs, int64_t freqs_len, int64_t *out_alleles, int32_t pstride,
              double *out_stats, double *out_freqs, double *out_occur, int64_t hap_out_len, float *out_gl,
              int64_t gl_len, mchb_item_result *results) {
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    if (n_items > 0x7fffffff) {
        h->err = "too many items in one call";
        return MCHB_ERR_ARGUMENT;
    }
    CallGeom g;
    int rc = check_call_items(h, items, n_items, reads_len, counts != nullptr, counts_len, haplotypes_len,
                              freqs != nullptr, freqs_len, mode == 0 ? hap_out_len : -1, mode == 1 ? gl_len : -1, true,
                              g);
    if (rc) return rc;
    if (mode == 0 && pstride < g.pmax) {
        h->err = "pstride smaller than the largest ploidy";
        return MCHB_ERR_ARGUMENT;
    }
    const int part_threads = exact_part_threads(g);
    const size_t smem = exact_smem(g, g.umax, part_threads);
    if (smem > (size_t)h->smem_optin) {
        h->err = "call-exact item needs more shared memory than one CTA can have";
        return MCHB_ERR_ARGUMENT;
    }
    void *ditems, *dcounter, *dresults;
    if ((rc = ensure(h, S_ITEMS, sizeof(mchb_call_item) * (size_t)n_items, &ditems))) return rc;
    if ((rc = ensure(h, S_COUNTER, sizeof(int32_t) * 8, &dcounter))) return rc;
    if ((rc = ensure(h, S_RESULTS, sizeof(mchb_item_result) * (size_t)n_items, &dresults))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_call_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dcounter, 0, sizeof(int32_t) * 8, h->stream));
    const double *dreads, *dfreqs;
    const int64_t *dcounts;
    const int8_t *dhaps;
    // Host reads of 32 MB or more whose items lie in ascending order are brought over piece by piece
    // on the copy stream, one piece of consecutive items per launch: the kernel of piece k runs while
    // piece k + 1 is on the bus (the reads are 95 % of the input bytes).
    int n_pieces = 1;
    if (mem == MCHB_MEM_HOST && reads && reads_len * (int64_t)sizeof(double) >= (32ll << 20) && n_items >= 4096) {
        n_pieces = 8;
        for (int64_t i = 1; i < n_items && n_pieces > 1; i++)
            if (items[i].reads_off < items[i - 1].reads_off) n_pieces = 1;
    }
    if (n_pieces > 1) {
        void *p;
        if ((rc = ensure(h, S_READS, sizeof(double) * (size_t)reads_len, &p))) return rc;
        dreads = (const double *)p;
        while ((int)h->evs.size() < 3 * n_pieces) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            h->evs.push_back(e);
        }
    } else if ((rc = stage_in(h, mem, S_READS, reads, reads_len, &dreads))) {
        return rc;
    }
    if ((rc = stage_in(h, mem, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, mem, S_HAPS, haplotypes, haplotypes_len, &dhaps))) return rc;
    if ((rc = stage_in(h, mem, S_FREQS, freqs, freqs_len, &dfreqs))) return rc;
    int ctas_per_sm = 1;
    int threads = 128;
    if (const char *env = getenv("MCHB_EXACT_THREADS")) threads = std::max(32, std::min(128, atoi(env) & ~31));  // tuning aid
    // one ploidy for the whole batch -> specialised kernel; parked log joints need a row per CTA
    const int fixed_p = (g.pmin == g.pmax) ? g.pmax : 0;
    const unsigned long long row_bytes = sizeof(double) * ((unsigned long long)g.gmax + 128);  // [ceil(G / 128)][128]
    unsigned long long budget = MCHB_EXACT_SCRATCH_BUDGET;
    if (const char *env = getenv("MCHB_EXACT_SCRATCH_BYTES")) budget = strtoull(env, nullptr, 10);  // tests: force the low-memory path
    const bool recomp = mode == 0 && row_bytes > budget;
#define EXACT_PREP(PM, FIXED, RECOMP)                                                                                   \
    CK(cudaFuncSetAttribute(exact_kernel<PM, FIXED, RECOMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, exact_kernel<PM, FIXED, RECOMP>, threads, smem))
    MCHB_EXACT_DISPATCH(fixed_p, recomp, EXACT_PREP);
#undef EXACT_PREP
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    long long grid = std::min<long long>(n_items, (long long)h->sm_count * ctas_per_sm);
    ExactArgs a;
    memset(&a, 0, sizeof(a));
    a.items = (const mchb_call_item *)ditems;
    a.n_items = (int32_t)n_items;
    a.reads = dreads;
    a.counts = dcounts;
    a.haplotypes = dhaps;
    a.freqs = dfreqs;
    a.mode = mode;
    a.work_counter = (int32_t *)dcounter;
    a.results = (mchb_item_result *)dresults;
    a.umax = g.umax;
    a.hmax = g.hmax;
    a.pmax = g.pmax;
    a.pstride = pstride;
    a.part_threads = std::min(part_threads, threads);
    int64_t *dalleles = nullptr;
    double *dstats = nullptr, *dfo = nullptr, *doc = nullptr;
    float *dgl = nullptr;
    if (mode == 0) {
        // parked log joints: one row of gmax doubles per CTA, inside the budget
        if (!recomp) {
            grid = std::max<long long>(1, std::min<long long>(grid, (long long)(budget / row_bytes)));
            void *scratch;
            if ((rc = ensure(h, S_SCRATCH, (size_t)row_bytes * (size_t)grid, &scratch))) return rc;
            a.scratch = (double *)scratch;
            a.scratch_stride = (int64_t)(row_bytes / sizeof(double));
        }  // else: a.scratch stays null and the second pass evaluates the log joints again
        if ((rc = stage_out(h, mem, S_OUT_A, out_alleles, n_items * pstride, &dalleles))) return rc;
        if ((rc = stage_out(h, mem, S_OUT_S, out_stats, n_items * 4, &dstats))) return rc;
        if ((rc = stage_out(h, mem, S_OUT_F, out_freqs, hap_out_len, &dfo))) return rc;
        if ((rc = stage_out(h, mem, S_OUT_O, out_occur, hap_out_len, &doc))) return rc;
        a.out_alleles = dalleles;
        a.out_stats = dstats;
        a.out_freqs = dfo;
        a.out_occur = doc;
    } else {
        if ((rc = stage_out(h, mem, S_OUT_GL, out_gl, gl_len, &dgl))) return rc;
        a.out_gl = dgl;
    }
#define EXACT_LAUNCH(PM, FIXED, RECOMP) exact_kernel<PM, FIXED, RECOMP><<<(unsigned)pgrid, threads, smem, h->stream>>>(a)
    if (n_pieces == 1) {
        const long long pgrid = grid;
        CK(cudaEventRecord(h->ev0, h->stream));
        MCHB_EXACT_DISPATCH(fixed_p, recomp, EXACT_LAUNCH);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev1, h->stream));
        h->launches++;
    } else {
        CK(cudaEventRecord(h->ev_fork, h->stream));  // the copy stream starts once the device buffers are ours
        CK(cudaStreamWaitEvent(h->cs, h->ev_fork, 0));
        const ExactArgs base = a;
        for (int k = 0; k < n_pieces; k++) {
            const int64_t i0 = n_items * k / n_pieces, i1 = n_items * (k + 1) / n_pieces;
            const int64_t lo = items[i0].reads_off;
            int64_t hi = lo;
            for (int64_t i = i0; i < i1; i++)
                hi = std::max(hi, items[i].reads_off + (int64_t)items[i].n_reads * items[i].n_pos * items[i].max_allele);
            CK(cudaMemcpyAsync((double *)dreads + lo, reads + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice, h->cs));
            CK(cudaEventRecord(h->evs[3 * k], h->cs));
            CK(cudaStreamWaitEvent(h->stream, h->evs[3 * k], 0));
            a = base;
            a.items = base.items + i0;
            a.n_items = (int32_t)(i1 - i0);
            a.results = base.results + i0;
            a.work_counter = base.work_counter + k;
            if (base.out_alleles) a.out_alleles = base.out_alleles + i0 * pstride;
            if (base.out_stats) a.out_stats = base.out_stats + i0 * 4;
            const long long pgrid = std::max<long long>(1, std::min<long long>(grid, i1 - i0));
            CK(cudaEventRecord(h->evs[3 * k + 1], h->stream));
            MCHB_EXACT_DISPATCH(fixed_p, recomp, EXACT_LAUNCH);
            CK(cudaGetLastError());
            CK(cudaEventRecord(h->evs[3 * k + 2], h->stream));
            h->launches++;
        }
    }
#undef EXACT_LAUNCH
    if (results)
        CK(cudaMemcpyAsync(results, dresults, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
    if (mem == MCHB_MEM_HOST) {
        if (mode == 0) {
            CK(cudaMemcpyAsync(out_alleles, dalleles, sizeof(int64_t) * (size_t)n_items * pstride, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(out_stats, dstats, sizeof(double) * (size_t)n_items * 4, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(out_freqs, dfo, sizeof(double) * (size_t)hap_out_len, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(out_occur, doc, sizeof(double) * (size_t)hap_out_len, cudaMemcpyDeviceToHost, h->stream));
        } else {
            CK(cudaMemcpyAsync(out_gl, dgl, sizeof(float) * (size_t)gl_len, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    if (n_pieces == 1) {
        CK(cudaEventElapsedTime(&h->kernel_ms, h->ev0, h->ev1));
    } else {
        h->kernel_ms = 0.f;  // the kernels' own time (they wait for their piece of the reads in between)
        for (int k = 0; k < n_pieces; k++) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->evs[3 * k + 1], h->evs[3 * k + 2]));
            h->kernel_ms += ms;
        }
    }
    return MCHB_OK;
}

}  // namespace

extern "C" {

int mchb_call_exact_mode_batch(mchb_handle *h, int mem, const mchb_call_item *items, int64_t n_items,
                               const double *reads, int64_t reads_len, const int64_t *counts, int64_t counts_len,
                               const int8_t *haplotypes, int64_t haplotypes_len, const double *freqs,
                               int64_t freqs_len, int64_t *out_alleles, int32_t pstride, double *out_stats,
                               double *out_freqs, double *out_occur, int64_t hap_out_len, mchb_item_result *results) {
    if (!h || !items || n_items < 0 || !out_alleles || !out_stats || !out_freqs || !out_occur) return MCHB_ERR_ARGUMENT;
    return run_exact(h, mem, 0, items, n_items, reads, reads_len, counts, counts_len, haplotypes, haplotypes_len, freqs,
                     freqs_len, out_alleles, pstride, out_stats, out_freqs, out_occur, hap_out_len, nullptr, 0, results);
}

int mchb_genotype_likelihoods_batch(mchb_handle *h, int mem, const mchb_call_item *items, int64_t n_items,
                                    const double *reads, int64_t reads_len, const int64_t *counts, int64_t counts_len,
                                    const int8_t *haplotypes, int64_t haplotypes_len, float *out_gl, int64_t gl_len,
                                    mchb_item_result *results) {
    if (!h || !items || n_items < 0 || !out_gl) return MCHB_ERR_ARGUMENT;
    return run_exact(h, mem, 1, items, n_items, reads, reads_len, counts, counts_len, haplotypes, haplotypes_len, nullptr, 0,
                     nullptr, 0, nullptr, nullptr, nullptr, 0, out_gl, gl_len, results);
}

int mchb_genotype_posteriors_batch(mchb_handle *h, int mem, const mchb_call_item *items, int64_t n_items,
                                   const double *freqs, int64_t freqs_len, const void *llks, int llk_is_f32,
                                   int64_t gl_len, double *out_gp, double *out_freqs, double *out_counts,
                                   double *out_occur, int64_t hap_out_len) {
    if (!h || !items || n_items < 0 || !llks || !out_gp) return MCHB_ERR_ARGUMENT;
    if (out_freqs && (!out_counts || !out_occur)) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    CallGeom g;
    int rc = check_call_items(h, items, n_items, 0, false, 0, 0, freqs != nullptr, freqs_len, out_freqs ? hap_out_len : -1,
                              gl_len, false, g);
    if (rc) return rc;
    const int part_threads = exact_part_threads(g);
    const size_t smem = exact_smem(g, 0, part_threads);
    void *ditems;
    if ((rc = ensure(h, S_ITEMS, sizeof(mchb_call_item) * (size_t)n_items, &ditems))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_call_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    const double *dfreqs;
    if ((rc = stage_in(h, mem, S_FREQS, freqs, freqs_len, &dfreqs))) return rc;
    PosteriorArgs a;
    memset(&a, 0, sizeof(a));
    a.items = (const mchb_call_item *)ditems;
    a.n_items = (int32_t)n_items;
    a.freqs = dfreqs;
    a.hmax = g.hmax;
    a.pmax = g.pmax;
    a.part_threads = part_threads;
    if (llk_is_f32) {
        const float *d;
        if ((rc = stage_in(h, mem, S_LLKS, (const float *)llks, gl_len, &d))) return rc;
        a.llk32 = d;
    } else {
        const double *d;
        if ((rc = stage_in(h, mem, S_LLKS, (const double *)llks, gl_len, &d))) return rc;
        a.llk64 = d;
    }
    double *dgp, *dfo = nullptr, *dco = nullptr, *doc = nullptr;
    if ((rc = stage_out(h, mem, S_OUT_GP, out_gp, gl_len, &dgp))) return rc;
    a.out_gp = dgp;
    if (out_freqs) {
        if ((rc = stage_out(h, mem, S_OUT_F, out_freqs, hap_out_len, &dfo))) return rc;
        if ((rc = stage_out(h, mem, S_OUT_C, out_counts, hap_out_len, &dco))) return rc;
        if ((rc = stage_out(h, mem, S_OUT_O, out_occur, hap_out_len, &doc))) return rc;
        a.out_freqs = dfo;
        a.out_counts = dco;
        a.out_occur = doc;
    }
    long long grid = std::min<long long>(n_items, (long long)h->sm_count * 8);
#define POST_PREP(PM) CK(cudaFuncSetAttribute(posterior_kernel<PM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
    MCHB_POSTERIOR_DISPATCH(g.pmax, POST_PREP);
#undef POST_PREP
    CK(cudaEventRecord(h->ev0, h->stream));
#define POST_LAUNCH(PM) posterior_kernel<PM><<<(unsigned)grid, 128, smem, h->stream>>>(a)
    MCHB_POSTERIOR_DISPATCH(g.pmax, POST_LAUNCH);
#undef POST_LAUNCH
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    if (mem == MCHB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_gp, dgp, sizeof(double) * (size_t)gl_len, cudaMemcpyDeviceToHost, h->stream));
        if (out_freqs) {
            CK(cudaMemcpyAsync(out_freqs, dfo, sizeof(double) * (size_t)hap_out_len, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(out_counts, dco, sizeof(double) * (size_t)hap_out_len, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(out_occur, doc, sizeof(double) * (size_t)hap_out_len, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->kernel_ms, h->ev0, h->ev1));
    return MCHB_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------ K5 call MCMC
extern "C" int mchb_call_mcmc_batch(mchb_handle *h, int mem, const mchb_call_mcmc_params *params,
                                    const mchb_call_item *items, int64_t n_items, const double *reads,
                                    int64_t reads_len, const int64_t *counts, int64_t counts_len,
                                    const int8_t *haplotypes, int64_t haplotypes_len, const double *freqs,
                                    int64_t freqs_len, const int32_t *initial, int32_t pstride, int32_t *out_alleles,
                                    int64_t out_alleles_len, double *out_llks, int64_t out_llks_len,
                                    mchb_item_result *results) {
    if (!h || !params || !items || !results || n_items < 0 || !out_alleles || !out_llks) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    if (n_items > 0x7fffffff) {
        h->err = "too many items in one call";
        return MCHB_ERR_ARGUMENT;
    }
    const mchb_call_mcmc_params &pp = *params;
    if (pp.steps < 0 || pp.chains < 0 || (pp.step_type != 0 && pp.step_type != 1)) {
        h->err = "bad call-mcmc parameters";
        return MCHB_ERR_ARGUMENT;
    }
    CallGeom g;
    int rc = check_call_items(h, items, n_items, reads_len, counts != nullptr, counts_len, haplotypes_len,
                              freqs != nullptr, freqs_len, -1, -1, true, g);
    if (rc) return rc;
    std::map<uint32_t, int32_t> seed_index;
    std::vector<uint32_t> seeds;
    std::vector<int32_t> item_stream((size_t)n_items, 0), order;
    int64_t words_needed = 0;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_call_item &it = items[i];
        const int64_t tsz = (int64_t)pp.chains * pp.steps * it.ploidy;
        if (it.gl_off < 0 || it.gl_off + tsz > out_alleles_len || it.hap_out_off < 0 ||
            it.hap_out_off + (int64_t)pp.chains * pp.steps > out_llks_len || (initial && pstride < it.ploidy)) {
            h->err = "call-mcmc item " + std::to_string(i) + " exceeds the output array lengths";
            return MCHB_ERR_ARGUMENT;
        }
        results[i].status = MCHB_ITEM_OK;
        results[i].n_het = 0;
        results[i].rng_words = 0;
        results[i].llk_evals = 0;
        order.push_back((int32_t)i);
        if (!pp.replay_words) {
            const uint32_t seed = (uint32_t)it.reserved;
            auto f = seed_index.find(seed);
            if (f == seed_index.end()) {
                f = seed_index.emplace(seed, (int32_t)seeds.size()).first;
                seeds.push_back(seed);
            }
            item_stream[(size_t)i] = f->second;
        }
        words_needed = std::max(words_needed, (int64_t)pp.chains * pp.steps * (4 * (int64_t)it.ploidy + 8) + 64);
    }
    size_t per_warp =
        (((size_t)g.umax * g.hmax + g.umax + 2 * (size_t)g.hmax + (size_t)g.hmax * (g.pmax + 2)) * 8 + 3 * (size_t)g.pmax * 4 + 512 + 15) &
        ~(size_t)15;
    // memo of the conditional distributions (cumulative sums + llks per genotype slot): kept in
    // shared memory when it is small next to the read x haplotype table
    const size_t memo_bytes = (2 * ((size_t)g.pmax * g.hmax * 16 + (size_t)g.pmax * g.pmax * 4 + (size_t)g.pmax * 4) + (size_t)g.pmax * 4 + 15) & ~(size_t)15;  // two ways per slot
    const bool use_memo = memo_bytes <= 16384;
    const size_t memo_off = per_warp;
    if (use_memo) per_warp += memo_bytes;
    int warps_per_cta = 4;
    while (warps_per_cta > 1 && per_warp * warps_per_cta > (size_t)h->smem_optin) warps_per_cta >>= 1;
    if (per_warp * warps_per_cta > (size_t)h->smem_optin) {
        h->err = "call-mcmc item needs more shared memory than one CTA can have";
        return MCHB_ERR_ARGUMENT;
    }
    const size_t smem = per_warp * warps_per_cta;
    void *ditems, *dstream, *dcounter, *dresults, *dorder;
    if ((rc = ensure(h, S_ITEMS, sizeof(mchb_call_item) * (size_t)n_items, &ditems))) return rc;
    if ((rc = ensure(h, S_STREAM, sizeof(int32_t) * (size_t)n_items, &dstream))) return rc;
    if ((rc = ensure(h, S_COUNTER, sizeof(int32_t) * 8, &dcounter))) return rc;
    if ((rc = ensure(h, S_RESULTS, sizeof(mchb_item_result) * (size_t)n_items, &dresults))) return rc;
    if ((rc = ensure(h, S_ORDER, sizeof(int32_t) * (size_t)n_items, &dorder))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_call_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dstream, item_stream.data(), sizeof(int32_t) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(dresults, results, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    const double *dreads, *dfreqs;
    const int64_t *dcounts;
    const int8_t *dhaps;
    const int32_t *dinit;
    int32_t *doa;
    double *dol;
    if ((rc = stage_in(h, mem, S_READS, reads, reads_len, &dreads))) return rc;
    if ((rc = stage_in(h, mem, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, mem, S_HAPS, haplotypes, haplotypes_len, &dhaps))) return rc;
    if ((rc = stage_in(h, mem, S_FREQS, freqs, freqs_len, &dfreqs))) return rc;
    if ((rc = stage_in(h, mem, S_INIT32, initial, initial ? n_items * pstride : 0, &dinit))) return rc;
    if (mem == MCHB_MEM_HOST) h->last_call_trace_len = 0;  // the scratch trace of an earlier tally call is overwritten
    if ((rc = stage_out(h, mem, S_OUT_A32, out_alleles, out_alleles_len, &doa))) return rc;
    if ((rc = stage_out(h, mem, S_OUT_L, out_llks, out_llks_len, &dol))) return rc;
    CK(cudaFuncSetAttribute(call_mcmc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, call_mcmc_kernel, warps_per_cta * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    int64_t stream_len = pp.rng_words_hint > 0 ? pp.rng_words_hint : words_needed;
    if (pp.replay_words) stream_len = pp.replay_len;
    stream_len = (stream_len + 31) & ~(int64_t)31;
    std::vector<int32_t> todo = order;
    for (int attempt = 0; attempt < 6 && !todo.empty(); attempt++) {
        uint32_t *dwords = nullptr;
        CK(cudaEventRecord(h->ev0, h->stream));  // the word-stream fill belongs to the timed device work
        if (pp.replay_words) {
            void *p;
            if ((rc = ensure(h, S_WORDS, sizeof(uint32_t) * (size_t)stream_len, &p))) return rc;
            CK(cudaMemsetAsync(p, 0, sizeof(uint32_t) * (size_t)stream_len, h->stream));
            CK(cudaMemcpyAsync(p, pp.replay_words, sizeof(uint32_t) * (size_t)pp.replay_len, cudaMemcpyHostToDevice, h->stream));
            dwords = (uint32_t *)p;
        } else {
            if ((rc = fill_streams(h, seeds, stream_len, &dwords))) return rc;
        }
        CK(cudaMemcpyAsync(dorder, todo.data(), sizeof(int32_t) * todo.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemsetAsync(dcounter, 0, sizeof(int32_t) * 8, h->stream));
        CallMcmcArgs a;
        memset(&a, 0, sizeof(a));
        a.items = (const mchb_call_item *)ditems;
        a.order = (const int32_t *)dorder;
        a.n_order = (int32_t)todo.size();
        a.reads = dreads;
        a.counts = dcounts;
        a.haplotypes = dhaps;
        a.freqs = dfreqs;
        a.initial = dinit;
        a.pstride = pstride;
        a.out_alleles = doa;
        a.out_llks = dol;
        a.results = (mchb_item_result *)dresults;
        a.words = dwords;
        a.item_stream = (const int32_t *)dstream;
        a.stream_len = pp.replay_words ? pp.replay_len : stream_len;
        a.steps = pp.steps;
        a.chains = pp.chains;
        a.step_type = pp.step_type;
        a.work_counter = (int32_t *)dcounter;
        a.umax = g.umax;
        a.hmax = g.hmax;
        a.pmax = g.pmax;
        a.smem_per_warp = (int32_t)per_warp;
        a.memo_off = use_memo ? (int32_t)memo_off : -1;
        long long want = ((long long)todo.size() + warps_per_cta - 1) / warps_per_cta;
        long long grid = std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * ctas_per_sm));
        call_mcmc_kernel<<<(unsigned)grid, warps_per_cta * 32, smem, h->stream>>>(a);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->ev1, h->stream));
        h->launches++;
        CK(cudaMemcpyAsync(results, dresults, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->kernel_ms += ms;
        if (pp.replay_words) break;
        std::vector<int32_t> next;
        for (int32_t id : todo)
            if (results[id].status == MCHB_ITEM_RNG_EXHAUSTED) next.push_back(id);
        todo.swap(next);
        stream_len *= 2;
    }
    if (mem == MCHB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_alleles, doa, sizeof(int32_t) * (size_t)out_alleles_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(out_llks, dol, sizeof(double) * (size_t)out_llks_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return MCHB_OK;
}

// ------------------------------------------------------------------- N1 trace post-processing
namespace {

// elem_size 1: int8 assembly traces; 4: int32 calling traces.  Lengths and offsets are in elements.
int tally_run(mchb_handle *h, int elem_size, int mem_in, int mem_out, const mchb_tally_item *items, int64_t n_items,
              const int8_t *genotypes, int64_t genotypes_len, int8_t *out_states, int64_t out_states_len,
              int32_t *out_counts, int32_t *out_first, int64_t tallies_len, mchb_item_result *results) {
    if (n_items == 0) return MCHB_OK;
    if (n_items > 0x7fffffff) {
        h->err = "too many items in one call";
        return MCHB_ERR_ARGUMENT;
    }
    int pn_max = 1, unique_max = 1;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_tally_item &it = items[i];
        const int64_t pn = (int64_t)it.ploidy * it.n_pos;
        bool bad = it.n_pos < 0 || it.ploidy < 1 || it.chains < 0 || it.steps < 0 || it.max_unique < 1 ||
                   it.ploidy > 32 || pn > 16384 || it.max_unique > 8192 || it.genotypes_off < 0 ||
                   it.genotypes_off + (int64_t)it.chains * it.steps * pn > genotypes_len || it.states_off < 0 ||
                   it.states_off + (int64_t)it.max_unique * pn > out_states_len || it.tallies_off < 0 ||
                   it.tallies_off + (int64_t)it.max_unique * it.chains > tallies_len;
        if (bad) {
            h->err = "tally item " + std::to_string(i) + " exceeds the given array lengths or the limits "
                     "(ploidy <= 32, ploidy * n_pos <= 16384, max_unique <= 8192)";
            return MCHB_ERR_ARGUMENT;
        }
        pn_max = std::max<int>(pn_max, (int)pn * elem_size);  // bytes
        unique_max = std::max(unique_max, it.max_unique);
    }
    int rc;
    void *ditems, *dresults, *dcounter;
    if ((rc = ensure(h, S_TITEMS, sizeof(mchb_tally_item) * (size_t)n_items, &ditems))) return rc;
    if ((rc = ensure(h, S_TRESULTS, sizeof(mchb_item_result) * (size_t)n_items, &dresults))) return rc;
    if ((rc = ensure(h, S_COUNTER, sizeof(int32_t) * 8, &dcounter))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_tally_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dcounter, 0, sizeof(int32_t) * 8, h->stream));
    const int8_t *dgeno;
    int8_t *dstates;
    int32_t *dcounts, *dfirst;
    if ((rc = stage_in(h, mem_in, S_TGENO, genotypes, genotypes_len * elem_size, &dgeno))) return rc;
    if ((rc = stage_out(h, mem_out, S_TSTATES, out_states, out_states_len * elem_size, &dstates))) return rc;
    if ((rc = stage_out(h, mem_out, S_TCOUNTS, out_counts, tallies_len, &dcounts))) return rc;
    if ((rc = stage_out(h, mem_out, S_TFIRST, out_first, tallies_len, &dfirst))) return rc;
    // the states array is only written up to n_unique per item: clear the rest for the host
    if (mem_out == MCHB_MEM_HOST)
        CK(cudaMemsetAsync(dstates, 0, (size_t)std::max<int64_t>(out_states_len * elem_size, 1), h->stream));
    const int tile_bytes = std::max(pn_max, 1024);
    size_t per_warp = (size_t)unique_max * 4 + (size_t)tile_bytes + 2 * (size_t)pn_max;
    per_warp = (per_warp + 15) & ~(size_t)15;
    int warps_per_cta = 4;
    while (warps_per_cta > 1 && per_warp * warps_per_cta > (size_t)h->smem_optin) warps_per_cta >>= 1;
    if (per_warp * warps_per_cta > (size_t)h->smem_optin) {
        h->err = "tally item needs more shared memory than one CTA can have";
        return MCHB_ERR_ARGUMENT;
    }
    const size_t smem = per_warp * warps_per_cta;
    int ctas_per_sm = 1;
    if (elem_size == 1) {
        CK(cudaFuncSetAttribute(tally_kernel<int8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, tally_kernel<int8_t>, warps_per_cta * 32, smem));
    } else {
        CK(cudaFuncSetAttribute(tally_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, tally_kernel<int32_t>, warps_per_cta * 32, smem));
    }
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    TallyArgs a;
    memset(&a, 0, sizeof(a));
    a.items = (const mchb_tally_item *)ditems;
    a.n_items = (int32_t)n_items;
    a.genotypes = dgeno;
    a.out_states = dstates;
    a.out_counts = dcounts;
    a.out_first = dfirst;
    a.results = (mchb_item_result *)dresults;
    a.work_counter = (int32_t *)dcounter;
    a.smem_per_warp = (int32_t)per_warp;
    a.pn_max = pn_max;
    a.tile_bytes = tile_bytes;
    a.unique_max = unique_max;
    long long want = (n_items + warps_per_cta - 1) / warps_per_cta;
    long long grid = std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * ctas_per_sm));
    CK(cudaEventRecord(h->ev0, h->stream));
    if (elem_size == 1) tally_kernel<int8_t><<<(unsigned)grid, warps_per_cta * 32, smem, h->stream>>>(a);
    else tally_kernel<int32_t><<<(unsigned)grid, warps_per_cta * 32, smem, h->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    CK(cudaMemcpyAsync(results, dresults, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
    if (mem_out == MCHB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_states, dstates, (size_t)out_states_len * elem_size, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(out_counts, dcounts, sizeof(int32_t) * (size_t)tallies_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(out_first, dfirst, sizeof(int32_t) * (size_t)tallies_len, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms += ms;
    return MCHB_OK;
}

}  // namespace

extern "C" int mchb_trace_tally_batch(mchb_handle *h, int mem_in, int mem_out, const mchb_tally_item *items,
                                      int64_t n_items, const int8_t *genotypes, int64_t genotypes_len,
                                      int8_t *out_states, int64_t out_states_len, int32_t *out_counts,
                                      int32_t *out_first, int64_t tallies_len, mchb_item_result *results) {
    if (!h || !items || !results || n_items < 0 || !out_states || !out_counts || !out_first ||
        (mem_in != MCHB_MEM_LAST_TRACE && !genotypes && genotypes_len > 0))
        return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (mem_in == MCHB_MEM_LAST_TRACE) {
        if (h->last_trace_len <= 0 || !h->bufs[S_OUT_G].p) {
            h->err = "no trace of an earlier mchb_assemble_tally_batch call is held by this handle";
            return MCHB_ERR_ARGUMENT;
        }
        return tally_run(h, 1, MCHB_MEM_DEVICE, mem_out, items, n_items, (const int8_t *)h->bufs[S_OUT_G].p,
                         h->last_trace_len, out_states, out_states_len, out_counts, out_first, tallies_len, results);
    }
    return tally_run(h, 1, mem_in, mem_out, items, n_items, genotypes, genotypes_len, out_states, out_states_len,
                     out_counts, out_first, tallies_len, results);
}

extern "C" int mchb_assemble_tally_batch(mchb_handle *h, const mchb_assemble_params *params,
                                         const mchb_assemble_item *items, const mchb_tally_item *tally_items,
                                         int64_t n_items, const double *reads, int64_t reads_len,
                                         const int64_t *counts, int64_t counts_len, const int8_t *n_alleles,
                                         int64_t n_alleles_len, const int8_t *initial, int64_t initial_len,
                                         int64_t genotypes_len, int64_t llks_len, int8_t *out_states,
                                         int64_t out_states_len, int32_t *out_counts, int32_t *out_first,
                                         int64_t tallies_len, mchb_item_result *results,
                                         mchb_item_result *tally_results) {
    if (!h || !params || !items || !tally_items || !results || !tally_results || n_items < 0 || !out_states ||
        !out_counts || !out_first || genotypes_len < 0 || llks_len < 0)
        return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    for (int64_t i = 0; i < n_items; i++) {
        if (tally_items[i].genotypes_off != items[i].genotypes_off || tally_items[i].n_pos != items[i].n_pos ||
            tally_items[i].ploidy != items[i].ploidy || tally_items[i].chains != params->chains ||
            tally_items[i].steps != params->steps) {
            h->err = "tally item " + std::to_string(i) + " does not describe the trace of assemble item " +
                     std::to_string(i);
            return MCHB_ERR_ARGUMENT;
        }
    }
    // inputs to the device, traces in the handle's scratch
    int rc;
    const double *dreads;
    const int64_t *dcounts;
    const int8_t *dnall, *dinit;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_READS, reads, reads_len, &dreads))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_NALLELES, n_alleles, n_alleles_len, &dnall))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_INITIAL, initial, initial_len, &dinit))) return rc;
    void *dog, *dol;
    if ((rc = ensure(h, S_OUT_G, (size_t)std::max<int64_t>(genotypes_len, 1), &dog))) return rc;
    if ((rc = ensure(h, S_OUT_L, sizeof(double) * (size_t)std::max<int64_t>(llks_len, 1), &dol))) return rc;
    rc = mchb_assemble_batch(h, MCHB_MEM_DEVICE, params, items, n_items, dreads, reads_len, dcounts, counts_len, dnall,
                             n_alleles_len, dinit, initial_len, (int8_t *)dog, genotypes_len, (double *)dol, llks_len,
                             results);
    if (rc) return rc;
    h->last_trace_len = genotypes_len;
    // items that failed in the sampler have no trace worth tallying: zero steps, empty tallies
    // (kernel_ms and launches keep accumulating: assemble + tally)
    std::vector<mchb_tally_item> titems(tally_items, tally_items + n_items);
    for (int64_t i = 0; i < n_items; i++)
        if (results[i].status != MCHB_ITEM_OK) titems[(size_t)i].steps = 0;
    return tally_run(h, 1, MCHB_MEM_DEVICE, MCHB_MEM_HOST, titems.data(), n_items, (const int8_t *)dog, genotypes_len,
                     out_states, out_states_len, out_counts, out_first, tallies_len, tally_results);
}

extern "C" int mchb_call_trace_tally_batch(mchb_handle *h, int mem_in, int mem_out, const mchb_tally_item *items,
                                           int64_t n_items, const int32_t *alleles, int64_t alleles_len,
                                           int32_t *out_states, int64_t out_states_len, int32_t *out_counts,
                                           int32_t *out_first, int64_t tallies_len, mchb_item_result *results) {
    if (!h || !items || !results || n_items < 0 || !out_states || !out_counts || !out_first ||
        (mem_in != MCHB_MEM_LAST_TRACE && !alleles && alleles_len > 0))
        return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    for (int64_t i = 0; i < n_items; i++)
        if (items[i].n_pos != 1) {
            h->err = "calling-trace tally items must have n_pos == 1 (one allele index per genotype slot)";
            return MCHB_ERR_ARGUMENT;
        }
    if (mem_in == MCHB_MEM_LAST_TRACE) {
        if (h->last_call_trace_len <= 0 || !h->bufs[S_OUT_A32].p) {
            h->err = "no trace of an earlier mchb_call_mcmc_tally_batch call is held by this handle";
            return MCHB_ERR_ARGUMENT;
        }
        return tally_run(h, 4, MCHB_MEM_DEVICE, mem_out, items, n_items, (const int8_t *)h->bufs[S_OUT_A32].p,
                         h->last_call_trace_len, (int8_t *)out_states, out_states_len, out_counts, out_first,
                         tallies_len, results);
    }
    return tally_run(h, 4, mem_in, mem_out, items, n_items, (const int8_t *)alleles, alleles_len, (int8_t *)out_states,
                     out_states_len, out_counts, out_first, tallies_len, results);
}

extern "C" int mchb_call_mcmc_tally_batch(mchb_handle *h, const mchb_call_mcmc_params *params,
                                          const mchb_call_item *items, const mchb_tally_item *tally_items,
                                          int64_t n_items, const double *reads, int64_t reads_len,
                                          const int64_t *counts, int64_t counts_len, const int8_t *haplotypes,
                                          int64_t haplotypes_len, const double *freqs, int64_t freqs_len,
                                          const int32_t *initial, int32_t pstride, int64_t alleles_len,
                                          int64_t llks_len, int32_t *out_states, int64_t out_states_len,
                                          int32_t *out_counts, int32_t *out_first, int64_t tallies_len,
                                          mchb_item_result *results, mchb_item_result *tally_results) {
    if (!h || !params || !items || !tally_items || !results || !tally_results || n_items < 0 || !out_states ||
        !out_counts || !out_first || alleles_len < 0 || llks_len < 0)
        return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    for (int64_t i = 0; i < n_items; i++) {
        if (tally_items[i].genotypes_off != items[i].gl_off || tally_items[i].n_pos != 1 ||
            tally_items[i].ploidy != items[i].ploidy || tally_items[i].chains != params->chains ||
            tally_items[i].steps != params->steps) {
            h->err = "tally item " + std::to_string(i) + " does not describe the trace of call item " + std::to_string(i);
            return MCHB_ERR_ARGUMENT;
        }
    }
    int rc;
    const double *dreads, *dfreqs;
    const int64_t *dcounts;
    const int8_t *dhaps;
    const int32_t *dinit;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_READS, reads, reads_len, &dreads))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_COUNTS, counts, counts_len, &dcounts))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_HAPS, haplotypes, haplotypes_len, &dhaps))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_FREQS, freqs, freqs_len, &dfreqs))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_INIT32, initial, initial ? n_items * pstride : 0, &dinit))) return rc;
    void *doa, *dol;
    if ((rc = ensure(h, S_OUT_A32, sizeof(int32_t) * (size_t)std::max<int64_t>(alleles_len, 1), &doa))) return rc;
    if ((rc = ensure(h, S_OUT_L, sizeof(double) * (size_t)std::max<int64_t>(llks_len, 1), &dol))) return rc;
    rc = mchb_call_mcmc_batch(h, MCHB_MEM_DEVICE, params, items, n_items, dreads, reads_len, dcounts, counts_len, dhaps,
                              haplotypes_len, dfreqs, freqs_len, dinit, pstride, (int32_t *)doa, alleles_len,
                              (double *)dol, llks_len, results);
    if (rc) return rc;
    h->last_call_trace_len = alleles_len;
    std::vector<mchb_tally_item> titems(tally_items, tally_items + n_items);
    for (int64_t i = 0; i < n_items; i++)
        if (results[i].status != MCHB_ITEM_OK) titems[(size_t)i].steps = 0;
    return tally_run(h, 4, MCHB_MEM_DEVICE, MCHB_MEM_HOST, titems.data(), n_items, (const int8_t *)doa, alleles_len,
                     (int8_t *)out_states, out_states_len, out_counts, out_first, tallies_len, tally_results);
}

// ------------------------------------------------------------ N2 read encoding + de-duplication
extern "C" int mchb_encode_reads_batch(mchb_handle *h, int mem, const mchb_encode_item *items, int64_t n_items,
                                       const int8_t *calls, int64_t calls_len, const double *probs, int64_t probs_len,
                                       const int8_t *n_alleles, int64_t n_alleles_len, double error_factor,
                                       double *out_reads, int64_t out_reads_len, int64_t *out_counts,
                                       int64_t out_counts_len, mchb_item_result *results) {
    if (!h || !items || !results || n_items < 0 || !out_reads || !out_counts) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    if (n_items > 0x7fffffff) {
        h->err = "too many items in one call";
        return MCHB_ERR_ARGUMENT;
    }
    int rmax = 1, emax = 1;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_encode_item &it = items[i];
        const int64_t rn = (int64_t)it.n_reads * it.n_pos, e = (int64_t)it.n_pos * it.max_allele;
        bool bad = it.n_reads < 0 || it.n_pos < 0 || it.max_allele < 0 || it.n_reads > 65536 || e > 4096 ||
                   it.calls_off < 0 || it.calls_off + rn > calls_len || it.probs_off < 0 ||
                   it.probs_off + rn > probs_len || it.nalleles_off < 0 || it.nalleles_off + it.n_pos > n_alleles_len ||
                   it.reads_off < 0 || it.reads_off + (int64_t)it.n_reads * e > out_reads_len || it.counts_off < 0 ||
                   it.counts_off + it.n_reads > out_counts_len;
        if (bad) {
            h->err = "encode item " + std::to_string(i) + " exceeds the given array lengths or the limits "
                     "(n_reads <= 65536, n_pos * max_allele <= 4096)";
            return MCHB_ERR_ARGUMENT;
        }
        rmax = std::max(rmax, it.n_reads);
        emax = std::max<int>(emax, (int)e);
    }
    int rc;
    void *ditems, *dresults, *dcounter;
    if ((rc = ensure(h, S_EITEMS, sizeof(mchb_encode_item) * (size_t)n_items, &ditems))) return rc;
    if ((rc = ensure(h, S_ERESULTS, sizeof(mchb_item_result) * (size_t)n_items, &dresults))) return rc;
    if ((rc = ensure(h, S_COUNTER, sizeof(int32_t) * 16, &dcounter))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_encode_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(dcounter, 0, sizeof(int32_t) * 16, h->stream));
    const int8_t *dcalls, *dnall;
    const double *dprobs;
    double *dreads;
    int64_t *dcounts;
    if ((rc = stage_in(h, mem, S_ECALLS, calls, calls_len, &dcalls))) return rc;
    if ((rc = stage_in(h, mem, S_EPROBS, probs, probs_len, &dprobs))) return rc;
    if ((rc = stage_in(h, mem, S_ENALL, n_alleles, n_alleles_len, &dnall))) return rc;
    if ((rc = stage_out(h, mem, S_EREADS, out_reads, out_reads_len, &dreads))) return rc;
    if ((rc = stage_out(h, mem, S_ECOUNTS, out_counts, out_counts_len, &dcounts))) return rc;
    if (mem == MCHB_MEM_HOST) {  // rows and counts beyond n_unique are not written by the kernel
        CK(cudaMemsetAsync(dreads, 0, sizeof(double) * (size_t)std::max<int64_t>(out_reads_len, 1), h->stream));
        CK(cudaMemsetAsync(dcounts, 0, sizeof(int64_t) * (size_t)std::max<int64_t>(out_counts_len, 1), h->stream));
    }
    size_t per_warp = ((size_t)emax * 8 + (size_t)rmax * 4 + 15) & ~(size_t)15;
    int warps_per_cta = 4;
    while (warps_per_cta > 1 && per_warp * warps_per_cta > (size_t)h->smem_optin) warps_per_cta >>= 1;
    if (per_warp * warps_per_cta > (size_t)h->smem_optin) {
        h->err = "encode item needs more shared memory than one CTA can have";
        return MCHB_ERR_ARGUMENT;
    }
    const size_t smem = per_warp * warps_per_cta;
    CK(cudaFuncSetAttribute(encode_reads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, encode_reads_kernel, warps_per_cta * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    EncodeArgs a;
    memset(&a, 0, sizeof(a));
    a.items = (const mchb_encode_item *)ditems;
    a.n_items = (int32_t)n_items;
    a.calls = dcalls;
    a.probs = dprobs;
    a.n_alleles = dnall;
    a.error_factor = error_factor;
    a.out_reads = dreads;
    a.out_counts = dcounts;
    a.results = (mchb_item_result *)dresults;
    a.work_counter = (int32_t *)dcounter;
    a.smem_per_warp = (int32_t)per_warp;
    a.rmax = rmax;
    a.emax = emax;
    long long want = (n_items + warps_per_cta - 1) / warps_per_cta;
    long long grid = std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * ctas_per_sm));
    CK(cudaEventRecord(h->ev0, h->stream));
    encode_reads_kernel<<<(unsigned)grid, warps_per_cta * 32, smem, h->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    CK(cudaMemcpyAsync(results, dresults, sizeof(mchb_item_result) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
    if (mem == MCHB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_reads, dreads, sizeof(double) * (size_t)out_reads_len, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(out_counts, dcounts, sizeof(int64_t) * (size_t)out_counts_len, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->kernel_ms += ms;
    return MCHB_OK;
}

extern "C" int mchb_encode_assemble_tally_batch(mchb_handle *h, const mchb_assemble_params *params,
                                                const mchb_encode_item *encode_items,
                                                const mchb_assemble_item *assemble_items,
                                                const mchb_tally_item *tally_items, int64_t n_items,
                                                const int8_t *calls, int64_t calls_len, const double *probs,
                                                int64_t probs_len, const int8_t *n_alleles, int64_t n_alleles_len,
                                                double error_factor, int64_t genotypes_len, int64_t llks_len,
                                                int8_t *out_states, int64_t out_states_len, int32_t *out_counts,
                                                int32_t *out_first, int64_t tallies_len,
                                                mchb_item_result *encode_results, mchb_item_result *results,
                                                mchb_item_result *tally_results) {
    if (!h || !params || !encode_items || !assemble_items || !tally_items || !encode_results || !results ||
        !tally_results || n_items < 0 || !out_states || !out_counts || !out_first || genotypes_len < 0 || llks_len < 0)
        return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    // ---- inputs to the device; encoded reads and counts in scratch (capacity: one row per read)
    int64_t reads_len = 0, counts_len = 0;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_encode_item &e = encode_items[i];
        if (e.n_reads < 0 || e.n_pos < 0 || e.max_allele < 0 || e.reads_off < 0 || e.counts_off < 0) {
            h->err = "bad encode item " + std::to_string(i);
            return MCHB_ERR_ARGUMENT;
        }
        reads_len = std::max(reads_len, e.reads_off + (int64_t)e.n_reads * e.n_pos * e.max_allele);
        counts_len = std::max(counts_len, e.counts_off + (int64_t)e.n_reads);
    }
    int rc;
    const int8_t *dcalls, *dnall;
    const double *dprobs;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_ECALLS, calls, calls_len, &dcalls))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_EPROBS, probs, probs_len, &dprobs))) return rc;
    if ((rc = stage_in(h, MCHB_MEM_HOST, S_ENALL, n_alleles, n_alleles_len, &dnall))) return rc;
    void *dreads, *dcounts, *dog, *dol;
    if ((rc = ensure(h, S_EREADS, sizeof(double) * (size_t)std::max<int64_t>(reads_len, 1), &dreads))) return rc;
    if ((rc = ensure(h, S_ECOUNTS, sizeof(int64_t) * (size_t)std::max<int64_t>(counts_len, 1), &dcounts))) return rc;
    if ((rc = ensure(h, S_OUT_G, (size_t)std::max<int64_t>(genotypes_len, 1), &dog))) return rc;
    if ((rc = ensure(h, S_OUT_L, sizeof(double) * (size_t)std::max<int64_t>(llks_len, 1), &dol))) return rc;
    rc = mchb_encode_reads_batch(h, MCHB_MEM_DEVICE, encode_items, n_items, dcalls, calls_len, dprobs, probs_len, dnall,
                                 n_alleles_len, error_factor, (double *)dreads, reads_len, (int64_t *)dcounts,
                                 counts_len, encode_results);
    if (rc) return rc;
    float total_ms = h->kernel_ms;
    int32_t total_launches = h->launches;
    // ---- the samplers read the distinct reads where the encoder left them
    std::vector<mchb_assemble_item> aitems(assemble_items, assemble_items + n_items);
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_encode_item &e = encode_items[i];
        mchb_assemble_item &it = aitems[(size_t)i];
        it.reads_off = e.reads_off;
        it.counts_off = e.counts_off;
        it.nalleles_off = e.nalleles_off;
        it.initial_off = -1;
        it.n_reads = encode_results[i].n_het;
        it.n_pos = e.n_pos;
        it.max_allele = std::max(e.max_allele, 1);
        if (tally_items[i].genotypes_off != it.genotypes_off || tally_items[i].n_pos != it.n_pos ||
            tally_items[i].ploidy != it.ploidy || tally_items[i].chains != params->chains ||
            tally_items[i].steps != params->steps) {
            h->err = "tally item " + std::to_string(i) + " does not describe the trace of assemble item " + std::to_string(i);
            return MCHB_ERR_ARGUMENT;
        }
    }
    rc = mchb_assemble_batch(h, MCHB_MEM_DEVICE, params, aitems.data(), n_items, (const double *)dreads, reads_len,
                             (const int64_t *)dcounts, counts_len, dnall, n_alleles_len, nullptr, 0, (int8_t *)dog,
                             genotypes_len, (double *)dol, llks_len, results);
    if (rc) return rc;
    h->last_trace_len = genotypes_len;
    total_ms += h->kernel_ms;
    total_launches += h->launches;
    std::vector<mchb_tally_item> titems(tally_items, tally_items + n_items);
    for (int64_t i = 0; i < n_items; i++)
        if (results[i].status != MCHB_ITEM_OK) titems[(size_t)i].steps = 0;
    h->kernel_ms = 0.f;
    h->launches = 0;
    rc = tally_run(h, 1, MCHB_MEM_DEVICE, MCHB_MEM_HOST, titems.data(), n_items, (const int8_t *)dog, genotypes_len,
                   out_states, out_states_len, out_counts, out_first, tallies_len, tally_results);
    h->kernel_ms += total_ms;
    h->launches += total_launches;
    return rc;
}

// --------------------------------------------------------------- minimum error correction
extern "C" int mchb_mec_batch(mchb_handle *h, int mem, const mchb_mec_item *items, int64_t n_items,
                              const int8_t *calls, int64_t calls_len, const int8_t *genotypes,
                              int64_t genotypes_len, int64_t *out_mec, int64_t *out_called,
                              int32_t *out_per_read, int64_t per_read_len) {
    if (!h || !items || !out_mec || n_items < 0) return MCHB_ERR_ARGUMENT;
    begin_call(h);
    CK(cudaSetDevice(h->device));
    if (n_items == 0) return MCHB_OK;
    for (int64_t i = 0; i < n_items; i++) {
        const mchb_mec_item &it = items[i];
        const int64_t rn = (int64_t)it.n_reads * it.n_pos;
        bool bad = it.n_reads < 0 || it.n_pos < 0 || it.ploidy < 0 || it.calls_off < 0 || it.calls_off + rn > calls_len ||
                   it.geno_off < 0 || it.geno_off + (int64_t)it.ploidy * it.n_pos > genotypes_len ||
                   (out_per_read && (it.per_read_off < 0 || it.per_read_off + it.n_reads > per_read_len));
        if (bad) {
            h->err = "mec item " + std::to_string(i) + " exceeds the given array lengths";
            return MCHB_ERR_ARGUMENT;
        }
    }
    int rc;
    void *ditems;
    if ((rc = ensure(h, S_ITEMS, sizeof(mchb_mec_item) * (size_t)n_items, &ditems))) return rc;
    CK(cudaMemcpyAsync(ditems, items, sizeof(mchb_mec_item) * (size_t)n_items, cudaMemcpyHostToDevice, h->stream));
    const int8_t *dcalls, *dgeno;
    int64_t *dmec, *dcalled = nullptr;
    int32_t *dper = nullptr;
    if ((rc = stage_in(h, mem, S_ECALLS, calls, calls_len, &dcalls))) return rc;
    if ((rc = stage_in(h, mem, S_GENO, genotypes, genotypes_len, &dgeno))) return rc;
    if ((rc = stage_out(h, mem, S_AUX0, out_mec, n_items, &dmec))) return rc;
    if (out_called && (rc = stage_out(h, mem, S_AUX1, out_called, n_items, &dcalled))) return rc;
    if (out_per_read && (rc = stage_out(h, mem, S_OUT_A32, out_per_read, per_read_len, &dper))) return rc;
    CK(cudaEventRecord(h->ev0, h->stream));
    mec_batch_kernel<<<(unsigned)((n_items + 3) / 4), 128, 0, h->stream>>>((const mchb_mec_item *)ditems, n_items, dcalls,
                                                                            dgeno, dmec, dcalled, dper);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    h->launches++;
    if (mem == MCHB_MEM_HOST) {
        CK(cudaMemcpyAsync(out_mec, dmec, sizeof(int64_t) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
        if (out_called)
            CK(cudaMemcpyAsync(out_called, dcalled, sizeof(int64_t) * (size_t)n_items, cudaMemcpyDeviceToHost, h->stream));
        if (out_per_read && per_read_len > 0)
            CK(cudaMemcpyAsync(out_per_read, dper, sizeof(int32_t) * (size_t)per_read_len, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->kernel_ms, h->ev0, h->ev1));
    return MCHB_OK;
}
