"""Batched host API over the C ABI: numpy in, numpy out (or raw device pointers in device mode).

The single-item, reference-shaped entry points (``DenovoMCMC.fit`` ...) live in
``mchap_b200.assemble`` / ``mchap_b200.calling`` and call into this module.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib as L

ASSEMBLE_ITEM_DTYPE = L._np_dtype(L.AssembleItem)
LLK_ITEM_DTYPE = L._np_dtype(L.LlkItem)
ITEM_RESULT_DTYPE = L._np_dtype(L.ItemResult)
TALLY_ITEM_DTYPE = L._np_dtype(L.TallyItem)
CALL_ITEM_DTYPE = L._np_dtype(L.CallItem)
MEC_ITEM_DTYPE = L._np_dtype(L.MecItem)

_ITEM_ERRORS = {
    L.ITEM_NAN_LLK: (ValueError, "Encountered log likelihood of nan"),
    L.ITEM_BREAKS: (ValueError, "breaks must be smaller then n"),
    L.ITEM_CHOICE_RANGE: (IndexError, "random_choice selected an option beyond the end of its probability vector"),
    L.ITEM_INITIAL_SHAPE: (AssertionError, "initial genotype does not have shape (ploidy, n_het_base)"),
    L.ITEM_RNG_EXHAUSTED: (RuntimeError, "pre-drawn random word stream exhausted"),
    L.ITEM_UNSUPPORTED: (NotImplementedError, "item shape is outside the limits of the CUDA kernels (see mchb_get_limits)"),
}


class MchapB200Error(RuntimeError):
    pass


def item_status_error(status, index=None):
    """The exception a per-item device status stands for (reference's type and message), or None."""
    if status == L.ITEM_OK:
        return None
    exc, msg = _ITEM_ERRORS.get(int(status), (MchapB200Error, "device status %d" % status))
    if index is not None:
        msg = "%s (item %d)" % (msg, index)
    return exc(msg)


def raise_item_status(status, index=None):
    """Re-raise a per-item device status with the reference's exception type and message."""
    exc = item_status_error(status, index)
    if exc is not None:
        raise exc


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


class Device:
    """One handle = one GPU + one stream (see include/mchap_b200.h)."""

    def __init__(self, device=0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.mchb_create(int(device), C.byref(h))
        if rc == L.MCHB_ERR_NO_DEVICE:
            raise MchapB200Error(
                "no sm_100 (B200) CUDA device %d: mchap_b200 has no CPU fallback" % device)
        if rc != L.MCHB_OK:
            raise MchapB200Error("mchb_create failed with status %d" % rc)
        self._h = h
        self.device = int(device)
        self._pin_lock = threading.Lock()

    def close(self):
        if getattr(self, "_h", None):
            for size, blocks in self.__dict__.get("_pin_free", {}).items():
                for addr in blocks:
                    self._lib.mchb_host_free(self._h, C.c_void_p(addr))
            self.__dict__["_pin_free"], self.__dict__["_pin_held"] = {}, 0
            self._lib.mchb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != L.MCHB_OK:
            msg = self._lib.mchb_last_error(self._h)
            raise MchapB200Error("status %d: %s" % (rc, msg.decode() if msg else ""))

    @property
    def last_kernel_ms(self):
        return float(self._lib.mchb_last_kernel_ms(self._h))

    @property
    def last_kernel_launches(self):
        return int(self._lib.mchb_last_kernel_launches(self._h))

    @property
    def last_host_chunks(self):
        return int(self._lib.mchb_last_host_chunks(self._h))

    @property
    def last_resident_warps(self):
        """Items (warps) the GPU held at once in the largest assemble launch of the last call."""
        return int(self._lib.mchb_last_resident_warps(self._h))

    @property
    def sm_count(self):
        return int(self._lib.mchb_sm_count(self._h))

    @property
    def stream(self):
        return self._lib.mchb_stream(self._h)

    @staticmethod
    def limits():
        lim = L.Limits()
        L.load().mchb_get_limits(C.byref(lim))
        return {f[0]: getattr(lim, f[0]) for f in lim._fields_}

    def measure_fp64_peak(self):
        """Sustained FP64 FMA TFLOP/s of this device (register-resident DFMA loop)."""
        out = C.c_double(0)
        self._check(self._lib.mchb_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ pinned host buffers
    PIN_POOL_CAP = 24 << 30   # bytes of released page-locked blocks kept for reuse

    @staticmethod
    def _pin_class(nbytes):
        """Size class of a page-locked block: next multiple of 1/8 of the power of two below it
        (at most 12.5 % larger than asked), 64 KB at least."""
        nbytes = max(int(nbytes), 1 << 16)
        step = 1 << max(nbytes.bit_length() - 4, 12)
        return -(-nbytes // step) * step

    def _pin_release(self, addr, size):
        with self._pin_lock:
            pool = self.__dict__.setdefault("_pin_free", {})
            held = self.__dict__.get("_pin_held", 0)
            if self._h and held + size <= self.PIN_POOL_CAP:
                pool.setdefault(size, []).append(addr)
                self.__dict__["_pin_held"] = held + size
                return
        if self._h:
            self._lib.mchb_host_free(self._h, C.c_void_p(addr))

    def pinned_empty(self, shape, dtype=np.float64):
        """Uninitialised numpy array in page-locked host memory.  Use it for the bulk arrays of
        host-buffer calls: transfers run at the full PCIe rate and overlap with kernels.  Page-locking
        is slow (about 2.5 GB/s), so blocks are recycled: when the array and all its views are garbage
        collected the block goes back to a pool of this Device and the next request of that size class
        gets it at no cost."""
        import weakref

        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        size = self._pin_class(n * dtype.itemsize)
        addr = None
        with self._pin_lock:
            free = self.__dict__.setdefault("_pin_free", {}).get(size)
            if free:
                addr = free.pop()
                self.__dict__["_pin_held"] -= size
        if addr is None:
            ptr = C.c_void_p()
            self._check(self._lib.mchb_host_alloc(self._h, size, C.byref(ptr)))
            addr = ptr.value
        buf = (C.c_char * size).from_address(addr)
        arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
        weakref.finalize(buf, self._pin_release, addr, size).atexit = False
        return arr

    def trace_buffer(self, n, dtype):
        """Output array of a batch call: page-locked (pooled) when it is large enough to matter."""
        nbytes = int(n) * np.dtype(dtype).itemsize
        if nbytes >= (8 << 20):
            return self.pinned_empty(int(n), dtype)
        return np.empty(max(int(n), 1), dtype=dtype)

    def pinned_concatenate(self, arrays, dtype):
        """np.concatenate(arrays, axis=None) written straight into a page-locked buffer."""
        n = int(sum(a.size for a in arrays))
        out = self.pinned_empty(n, dtype)
        if arrays and n:
            np.concatenate(arrays, axis=None, out=out)
        return out

    # ------------------------------------------------------------------ RNG / ranking
    def mt19937_words(self, seed, n):
        """numba's MT19937 output stream after np.random.seed(seed) (jitutils.py:180-183)."""
        out = np.empty(int(n), dtype=np.uint32)
        self._check(self._lib.mchb_mt19937_words(self._h, L.MEM_HOST, int(seed), _ptr(out), int(n)))
        return out

    def genotype_alleles_as_index(self, alleles):
        """jitutils.py:253-276 for an int array [n, ploidy] of sorted alleles."""
        a = np.ascontiguousarray(alleles, dtype=np.int64)
        assert a.ndim == 2
        out = np.empty(a.shape[0], dtype=np.int64)
        self._check(self._lib.mchb_genotype_rank(self._h, L.MEM_HOST, _ptr(a), a.shape[0], a.shape[1], _ptr(out)))
        return out

    def index_as_genotype_alleles(self, index, ploidy):
        """jitutils.py:279-318 for an int array [n] (negative index -> -1 alleles)."""
        i = np.ascontiguousarray(index, dtype=np.int64)
        out = np.empty((i.shape[0], int(ploidy)), dtype=np.int64)
        self._check(self._lib.mchb_genotype_unrank(self._h, L.MEM_HOST, _ptr(i), i.shape[0], int(ploidy), _ptr(out)))
        return out

    # ------------------------------------------------------------------ MEC
    def minimum_error_correction_batch(self, calls_list, genotypes_list, per_read=False):
        """encoding/integer/stats.py:18-39 for many (read calls int[R, N], genotype int[P, N]) pairs:
        returns (mec sums int64[n], called-base counts int64[n]) and, with per_read=True, the list of
        per-read arrays the reference returns."""
        n = len(calls_list)
        cs = [np.ascontiguousarray(c, dtype=np.int8) for c in calls_list]
        gs = [np.ascontiguousarray(g, dtype=np.int8) for g in genotypes_list]
        items = np.zeros(n, dtype=MEC_ITEM_DTYPE)
        R_ = np.array([c.shape[0] for c in cs], dtype=np.int64)
        N_ = np.array([c.shape[1] for c in cs], dtype=np.int64)
        P_ = np.array([g.shape[0] for g in gs], dtype=np.int64)
        assert all(g.ndim == 2 and g.shape[1] == c.shape[1] for g, c in zip(gs, cs))
        excl = lambda x: np.concatenate([[0], np.cumsum(x)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        items["calls_off"], items["geno_off"], items["per_read_off"] = excl(R_ * N_), excl(P_ * N_), excl(R_)
        items["n_reads"], items["n_pos"], items["ploidy"] = R_, N_, P_
        calls = np.concatenate(cs, axis=None) if n else np.zeros(0, dtype=np.int8)
        genos = np.concatenate(gs, axis=None) if n else np.zeros(0, dtype=np.int8)
        mec = np.zeros(n, dtype=np.int64)
        called = np.zeros(n, dtype=np.int64)
        rows = np.zeros(max(int(R_.sum()), 1), dtype=np.int32) if per_read else None
        self._check(self._lib.mchb_mec_batch(
            self._h, L.MEM_HOST, _ptr(items), n, _ptr(calls), calls.size, _ptr(genos), genos.size, _ptr(mec),
            _ptr(called), _ptr(rows), int(R_.sum()) if per_read else 0))
        if per_read:
            offs = items["per_read_off"]
            return mec, called, [rows[int(o): int(o) + int(r)].astype(np.int64) for o, r in zip(offs, R_)]
        return mec, called

    # ------------------------------------------------------------------ K1
    def log_likelihood_batch(self, reads_list, genotypes_list, counts_list=None):
        """assemble/likelihood.py:18-70 for many (reads, genotype[, counts]) triples."""
        n = len(reads_list)
        items = np.zeros(n, dtype=LLK_ITEM_DTYPE)
        ro = go = co = 0
        rs, gs, cs = [], [], []
        use_counts = counts_list is not None and any(c is not None for c in counts_list)
        for i in range(n):
            r = np.ascontiguousarray(reads_list[i], dtype=np.float64)
            g = np.ascontiguousarray(genotypes_list[i], dtype=np.int8)
            assert r.ndim == 3 and g.ndim == 2 and g.shape[1] == r.shape[1]
            items[i] = (ro, co, go, r.shape[0], r.shape[1], r.shape[2], g.shape[0])
            rs.append(r.ravel())
            gs.append(g.ravel())
            ro += r.size
            go += g.size
            if use_counts:
                c = counts_list[i]
                c = np.ones(r.shape[0], dtype=np.int64) if c is None else np.ascontiguousarray(c, dtype=np.int64)
                cs.append(c)
                co += c.size
        reads = np.concatenate(rs) if rs else np.zeros(0)
        genos = np.concatenate(gs) if gs else np.zeros(0, dtype=np.int8)
        counts = np.concatenate(cs) if use_counts else None
        out = np.empty(n, dtype=np.float64)
        self._check(self._lib.mchb_log_likelihood_batch(
            self._h, L.MEM_HOST, _ptr(items), n, _ptr(reads), reads.size, _ptr(counts),
            0 if counts is None else counts.size, _ptr(genos), genos.size, _ptr(out)))
        return out

    # ------------------------------------------------------------------ K4
    def call_exact_mode(self, batch):
        """calling/exact.py:156-249 for a CallBatch -> dict of arrays."""
        n = batch.n
        alleles = np.zeros((n, batch.pmax), dtype=np.int64)
        stats = np.zeros((n, 4), dtype=np.float64)
        freqs = np.zeros(max(batch.hap_total, 1), dtype=np.float64)
        occur = np.zeros(max(batch.hap_total, 1), dtype=np.float64)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        self._check(self._lib.mchb_call_exact_mode_batch(
            self._h, L.MEM_HOST, _ptr(batch.items), n, _ptr(batch.reads), batch.reads.size, _ptr(batch.counts),
            0 if batch.counts is None else batch.counts.size, _ptr(batch.haps), batch.haps.size, _ptr(batch.freqs),
            0 if batch.freqs is None else batch.freqs.size, _ptr(alleles), batch.pmax, _ptr(stats), _ptr(freqs),
            _ptr(occur), batch.hap_total, _ptr(results)))
        return dict(alleles=alleles, stats=stats, freqs=freqs, occur=occur, results=results)

    def genotype_likelihoods(self, batch):
        """calling/exact.py:266-292 for a CallBatch -> float32 array (items back to back)."""
        gl = np.zeros(max(batch.gl_total, 1), dtype=np.float32)
        results = np.zeros(batch.n, dtype=ITEM_RESULT_DTYPE)
        self._check(self._lib.mchb_genotype_likelihoods_batch(
            self._h, L.MEM_HOST, _ptr(batch.items), batch.n, _ptr(batch.reads), batch.reads.size, _ptr(batch.counts),
            0 if batch.counts is None else batch.counts.size, _ptr(batch.haps), batch.haps.size, _ptr(gl),
            batch.gl_total, _ptr(results)))
        return gl[: batch.gl_total]

    def genotype_posteriors(self, items, llks, freqs, gl_total, hap_total, with_frequencies=False):
        """calling/exact.py:295-329 (+ 332-369) from stored llk arrays (float32 or float64)."""
        llks = np.ascontiguousarray(llks)
        is32 = llks.dtype == np.float32
        if not is32:
            llks = np.ascontiguousarray(llks, dtype=np.float64)
        gp = np.zeros(max(gl_total, 1), dtype=np.float64)
        of = oc = oo = None
        if with_frequencies:
            of = np.zeros(max(hap_total, 1))
            oc = np.zeros(max(hap_total, 1))
            oo = np.zeros(max(hap_total, 1))
        self._check(self._lib.mchb_genotype_posteriors_batch(
            self._h, L.MEM_HOST, _ptr(items), len(items), _ptr(freqs), 0 if freqs is None else freqs.size, _ptr(llks),
            1 if is32 else 0, gl_total, _ptr(gp), _ptr(of), _ptr(oc), _ptr(oo), hap_total))
        return gp[:gl_total], of, oc, oo

    # ------------------------------------------------------------------ K5
    def call_mcmc(self, batch, steps, chains, step_type, initial=None, pstride=0, replay_words=None):
        """calling/classes.py:49-124 for a CallBatch whose items carry seeds (reserved) and output
        offsets (gl_off / hap_out_off) -> dict(alleles int32, llks f64, results)."""
        n = batch.n
        per = chains * steps
        a_len = int(per * batch.items["ploidy"].astype(np.int64).sum())
        l_len = per * n
        alleles = self.trace_buffer(a_len, np.int32)
        llks = self.trace_buffer(l_len, np.float64)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        p = L.CallMcmcParams()
        p.steps, p.chains, p.step_type = int(steps), int(chains), int(step_type)
        rw = None if replay_words is None else np.ascontiguousarray(replay_words, dtype=np.uint32)
        p.replay_words = None if rw is None else rw.ctypes.data
        p.replay_len = 0 if rw is None else rw.size
        p.rng_words_hint = 0
        ini = None if initial is None else np.ascontiguousarray(initial, dtype=np.int32)
        self._check(self._lib.mchb_call_mcmc_batch(
            self._h, L.MEM_HOST, C.byref(p), _ptr(batch.items), n, _ptr(batch.reads), batch.reads.size,
            _ptr(batch.counts), 0 if batch.counts is None else batch.counts.size, _ptr(batch.haps), batch.haps.size,
            _ptr(batch.freqs), 0 if batch.freqs is None else batch.freqs.size, _ptr(ini), int(pstride),
            _ptr(alleles), a_len, _ptr(llks), l_len, _ptr(results)))
        return dict(alleles=alleles, llks=llks, results=results)

    def call_trace_tally_call(self, items, alleles, alleles_len, out_states, out_counts, out_first,
                              mem_in=L.MEM_HOST, mem_out=L.MEM_HOST):
        """Thin wrapper of mchb_call_trace_tally_batch (int32 traces, items with n_pos == 1)."""
        n = len(items)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        self._check(self._lib.mchb_call_trace_tally_batch(
            self._h, mem_in, mem_out, _ptr(items), n, _ptr(alleles), int(alleles_len), _ptr(out_states),
            int(out_states.size), _ptr(out_counts), _ptr(out_first), int(out_counts.size), _ptr(results)))
        return results

    def call_mcmc_tally(self, batch, tally_items, steps, chains, step_type, out_states, out_counts, out_first,
                        initial=None, pstride=0):
        """mchb_call_mcmc_tally_batch: the sampler of call_mcmc with the traces kept on the device and
        tallied there -> (results, tally_results)."""
        n = batch.n
        per = chains * steps
        a_len = int(per * batch.items["ploidy"].astype(np.int64).sum())
        l_len = per * n
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        tally_results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        p = L.CallMcmcParams()
        p.steps, p.chains, p.step_type = int(steps), int(chains), int(step_type)
        p.replay_words, p.replay_len, p.rng_words_hint = None, 0, 0
        ini = None if initial is None else np.ascontiguousarray(initial, dtype=np.int32)
        self._check(self._lib.mchb_call_mcmc_tally_batch(
            self._h, C.byref(p), _ptr(batch.items), _ptr(tally_items), n, _ptr(batch.reads), batch.reads.size,
            _ptr(batch.counts), 0 if batch.counts is None else batch.counts.size, _ptr(batch.haps), batch.haps.size,
            _ptr(batch.freqs), 0 if batch.freqs is None else batch.freqs.size, _ptr(ini), int(pstride),
            a_len, l_len, _ptr(out_states), int(out_states.size), _ptr(out_counts), _ptr(out_first),
            int(out_counts.size), _ptr(results), _ptr(tally_results)))
        return results, tally_results

    # ------------------------------------------------------------------ K2
    def assemble_call(self, items, params, reads, counts, n_alleles, initial, out_genotypes, out_llks,
                      lens, mem=L.MEM_HOST, keepalive=()):
        """Thin wrapper of mchb_assemble_batch. ``items`` is a structured array of
        ASSEMBLE_ITEM_DTYPE; bulk arrays are numpy arrays (host) or integer device pointers;
        ``lens`` = (reads_len, counts_len, n_alleles_len, initial_len, genotypes_len, llks_len)."""
        n = len(items)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        rc = self._lib.mchb_assemble_batch(
            self._h, mem, C.byref(params), _ptr(items), n, _ptr(reads), int(lens[0]), _ptr(counts), int(lens[1]),
            _ptr(n_alleles), int(lens[2]), _ptr(initial), int(lens[3]), _ptr(out_genotypes), int(lens[4]),
            _ptr(out_llks), int(lens[5]), _ptr(results))
        self._check(rc)
        return results


    # ------------------------------------------------------------------ N1 trace tallies
    def trace_tally_call(self, items, genotypes, genotypes_len, out_states, out_counts, out_first,
                         mem_in=L.MEM_HOST, mem_out=L.MEM_HOST):
        """Thin wrapper of mchb_trace_tally_batch (items: structured array of TALLY_ITEM_DTYPE)."""
        n = len(items)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        rc = self._lib.mchb_trace_tally_batch(
            self._h, mem_in, mem_out, _ptr(items), n, _ptr(genotypes), int(genotypes_len), _ptr(out_states),
            int(out_states.size), _ptr(out_counts), _ptr(out_first), int(out_counts.size), _ptr(results))
        self._check(rc)
        return results

    def assemble_tally_call(self, items, tally_items, params, reads, counts, n_alleles, initial, lens,
                            out_states, out_counts, out_first):
        """Thin wrapper of mchb_assemble_tally_batch: host inputs, traces stay on the device, host
        tallies.  ``lens`` = (reads_len, counts_len, n_alleles_len, initial_len, genotypes_len, llks_len)."""
        n = len(items)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        tally_results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        rc = self._lib.mchb_assemble_tally_batch(
            self._h, C.byref(params), _ptr(items), _ptr(tally_items), n, _ptr(reads), int(lens[0]), _ptr(counts),
            int(lens[1]), _ptr(n_alleles), int(lens[2]), _ptr(initial), int(lens[3]), int(lens[4]), int(lens[5]),
            _ptr(out_states), int(out_states.size), _ptr(out_counts), _ptr(out_first), int(out_counts.size),
            _ptr(results), _ptr(tally_results))
        self._check(rc)
        return results, tally_results


def count_genotypes(n_haplotypes, ploidy):
    """Number of multisets of size ploidy over n_haplotypes (combinatorics.py:35-54, exact integer)."""
    from math import comb

    return comb(int(n_haplotypes) + int(ploidy) - 1, int(ploidy))


class CallBatch:
    """Packed inputs of a batch of call / call-exact items (host arrays + descriptors)."""

    def __init__(self, reads_list, haplotypes_list, ploidy, counts_list=None, priors=None, device=None):
        """device: when given, the flat reads / counts / haplotypes arrays are built in page-locked
        memory of that Device (one copy less on the way to the GPU)."""
        n = len(reads_list)
        self.n = n
        ploidies = np.broadcast_to(np.asarray(ploidy, dtype=np.int64), (n,))
        rs, hs, cs, fs = [], [], [], []
        use_counts = counts_list is not None and any(c is not None for c in counts_list)
        Us, Ns, As, Hs = [0] * n, [0] * n, [0] * n, [0] * n   # per-item scalars, assigned column-wise below
        inbs, foffs = [np.nan] * n, [-1] * n
        fo = 0
        f64, i8, nd = np.float64, np.int8, np.ndarray
        for i in range(n):
            r, hp = reads_list[i], haplotypes_list[i]
            if not (type(r) is nd and r.dtype == f64 and r.flags.c_contiguous):
                r = np.ascontiguousarray(r, dtype=f64)
            if not (type(hp) is nd and hp.dtype == i8 and hp.flags.c_contiguous):
                hp = np.ascontiguousarray(hp, dtype=i8)
            assert r.ndim == 3 and hp.ndim == 2 and hp.shape[1] == r.shape[1]
            Us[i], Ns[i], As[i] = r.shape
            H = Hs[i] = hp.shape[0]
            prior = None if priors is None else priors[i]
            if prior is not None:
                inb, fr = prior
                inbs[i] = float(inb)
                if fr is not None:
                    fr = np.ascontiguousarray(fr, dtype=np.float64)
                    assert len(fr) == H
                    foffs[i] = fo
                    fs.append(fr)
                    fo += H
            rs.append(r)      # (joined with axis=None below: no per-item flattening)
            hs.append(hp)
            if use_counts:
                c = counts_list[i]
                cs.append(np.ones(Us[i], dtype=np.int64) if c is None else np.ascontiguousarray(c, dtype=np.int64))
        U_, N_, A_, H_ = (np.asarray(x, dtype=np.int64) for x in (Us, Ns, As, Hs))
        excl = lambda x: np.concatenate([[0], np.cumsum(x)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        memo = {}
        for k in set(zip(Hs, ploidies.tolist())):
            memo[k] = count_genotypes(*k)
        self.n_genotypes = np.array([memo[k] for k in zip(Hs, ploidies.tolist())], dtype=np.int64)
        items = np.zeros(n, dtype=CALL_ITEM_DTYPE)
        items["reads_off"] = excl(U_ * N_ * A_)
        items["counts_off"] = excl(U_) if use_counts else 0
        items["haps_off"] = excl(H_ * N_)
        items["hap_out_off"], items["gl_off"] = excl(H_), excl(self.n_genotypes)
        items["n_reads"], items["n_pos"], items["max_allele"], items["n_haps"] = U_, N_, A_, H_
        items["ploidy"] = ploidies
        items["freqs_off"], items["inbreeding"] = foffs, inbs
        self.items = items
        if device is not None and n:
            self.reads = device.pinned_concatenate(rs, np.float64)
            self.haps = device.pinned_concatenate(hs, np.int8)
            self.counts = device.pinned_concatenate(cs, np.int64) if use_counts else None
        else:
            self.reads = np.concatenate(rs, axis=None) if rs else np.zeros(0)
            self.haps = np.concatenate(hs, axis=None) if hs else np.zeros(0, dtype=np.int8)
            self.counts = np.concatenate(cs, axis=None) if use_counts else None
        self.freqs = np.concatenate(fs) if fs else None
        self.hap_total = int(H_.sum())
        self.gl_total = int(self.n_genotypes.sum())
        self.pmax = int(ploidies.max()) if n else 1


def make_assemble_params(steps, chains, fix_homozygous, p_recombination, p_partial_dosage, p_dosage,
                         break_table, break_len, temperatures, replay_words=None, rng_words_hint=0,
                         sort_haplotypes=False):
    """Build the parameter struct; returns (params, keepalive tuple of the arrays it points to)."""
    bt = np.ascontiguousarray(break_table, dtype=np.float64)
    bl = np.ascontiguousarray(break_len, dtype=np.int32)
    tp = np.ascontiguousarray(temperatures, dtype=np.float64)
    rw = None if replay_words is None else np.ascontiguousarray(replay_words, dtype=np.uint32)
    p = L.AssembleParams()
    p.steps = int(steps)
    p.chains = int(chains)
    p.fix_homozygous = float(fix_homozygous)
    p.p_recombination = float(p_recombination)
    p.p_partial_dosage = float(p_partial_dosage)
    p.p_dosage = float(p_dosage)
    p.break_table = bt.ctypes.data
    p.break_len = bl.ctypes.data
    p.break_rows = bt.shape[0]
    p.break_stride = bt.shape[1]
    p.temperatures = tp.ctypes.data
    p.temperatures_len = tp.size
    p.sort_haplotypes = 1 if sort_haplotypes else 0
    p.replay_words = None if rw is None else rw.ctypes.data
    p.replay_len = 0 if rw is None else rw.size
    p.rng_words_hint = int(rng_words_hint)
    return p, (bt, bl, tp, rw)


def uniform_assemble_items(offsets, n_pos, max_allele, ploidy, chains, steps, n_temps=1, seed=42,
                           inbreeding=None):
    """Vectorised descriptors for a batch whose items share n_pos / max_allele / ploidy and whose
    reads are packed back to back (``offsets`` in reads, int64[n+1]); outputs are dense per item."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    items = np.zeros(n, dtype=ASSEMBLE_ITEM_DTYPE)
    idx = np.arange(n, dtype=np.int64)
    items["reads_off"] = offsets[:-1] * (n_pos * max_allele)
    items["counts_off"] = offsets[:-1]
    items["nalleles_off"] = idx * n_pos
    items["initial_off"] = -1
    items["genotypes_off"] = idx * (chains * steps * ploidy * n_pos)
    items["llks_off"] = idx * (chains * steps)
    items["n_reads"] = np.diff(offsets)
    items["n_pos"] = n_pos
    items["max_allele"] = max_allele
    items["ploidy"] = ploidy
    items["temps_off"] = 0
    items["n_temps"] = n_temps
    items["seed"] = np.uint32(seed)
    items["inbreeding"] = np.nan if inbreeding is None else float(inbreeding)
    return items


_default_devices = {}


def default_device(device=0):
    """Process-wide handle per GPU ordinal (created on first use)."""
    d = _default_devices.get(device)
    if d is None:
        d = Device(device)
        _default_devices[device] = d
    return d
