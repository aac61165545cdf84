"""``DenovoMCMC`` with the reference's constructor and ``fit`` signature
(reference: mchap/assemble/mcmc.py:24-161), executed by the CUDA kernel ``assemble_kernel``.

``fit`` handles one (locus, sample) item like the reference; ``fit_batch`` sends many items in
one device call (the reason this package exists).  There is no CPU path.
"""
from dataclasses import dataclass

import numpy as np

from .. import _lib as L
import ctypes as C

from ..api import (ASSEMBLE_ITEM_DTYPE, ITEM_RESULT_DTYPE, TALLY_ITEM_DTYPE, _ptr, default_device,
                   make_assemble_params, raise_item_status)
from .._lib import ITEM_TALLY_OVERFLOW as TALLY_OVERFLOW
from .classes import GenotypeMultiTrace, TraceTally

__all__ = ["DenovoMCMC", "point_beta_probabilities", "break_table"]


def point_beta_probabilities(n_base, a=1.0, b=3.0):
    """P(number of break points = k), k = 0..n_base-1: CDF differences of Beta(a, b) on the grid
    k / n_base.  Host-side scipy call, as in the reference (mcmc.py:429-452)."""
    from scipy import stats

    grid = np.arange(1, n_base + 1) / n_base
    cdf = stats.beta(a, b).cdf(grid)
    cdf[1:] = cdf[1:] - cdf[:-1]
    return cdf


def break_table(n_pos, alpha=1.0, beta=3.0, n_intervals=None):
    """Row n = break-point distribution used when n positions stay variable (mcmc.py:211-217)."""
    stride = max(int(n_pos), int(n_intervals or 0), 1)
    table = np.zeros((n_pos + 1, stride), dtype=np.float64)
    lens = np.zeros(n_pos + 1, dtype=np.int32)
    for n in range(1, n_pos + 1):
        if n_intervals is None:
            row = point_beta_probabilities(n, alpha, beta)
        else:
            row = np.zeros(n_intervals, dtype=np.float64)
            row[-1] = 1
        table[n, : len(row)] = row
        lens[n] = len(row)
    return table, lens


def sub2(lst, idx):
    return None if lst is None else [lst[i] for i in idx]


@dataclass
class DenovoMCMC(object):
    ploidy: int
    n_alleles: list
    inbreeding: float = None
    steps: int = 1000
    chains: int = 2
    alpha: float = 1.0
    beta: float = 3.0
    n_intervals: int = None
    fix_homozygous: float = 0.999
    recombination_step_probability: float = 0.5
    partial_dosage_step_probability: float = 0.5
    dosage_step_probability: float = 1.0
    temperatures: tuple = (1.0,)
    random_seed: int = None
    llk_cache_threshold: int = 100  # accepted for signature parity; the GPU path caches per-haplotype products instead
    device: object = None

    @classmethod
    def parameterize(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def _temperatures(self):
        temps = np.sort(np.asarray(self.temperatures, dtype=np.float64))
        assert temps[0] >= 0.0
        assert temps[-1] == 1.0
        return temps

    def _seed(self):
        if self.random_seed is not None:
            return int(self.random_seed) & 0xFFFFFFFF
        return int(np.random.randint(0, 2 ** 32, dtype=np.uint64))

    def fit(self, reads, read_counts=None, initial=None):
        """Same contract as the reference's fit: returns a GenotypeMultiTrace with genotypes
        int8[chains, steps, ploidy, n_positions] (haplotypes sorted per step) and llks."""
        res = self.fit_batch([reads], [read_counts], None if initial is None else [initial])
        return res[0]

    def _pack(self, reads_list, counts_list, initial_list, n_alleles_list, seeds):
        """Flat input arrays + item descriptors of a batch (outputs laid out item after item)."""
        n = len(reads_list)
        temps = self._temperatures()
        use_counts = counts_list is not None and any(c is not None for c in counts_list)
        use_initial = initial_list is not None and any(i is not None for i in initial_list)
        seed0 = self._seed()
        default_na = None if n_alleles_list is not None else np.ascontiguousarray(self.n_alleles, dtype=np.int8)
        rs, cs, ns, ins = [], [], [], []
        Us, Ns, As = [0] * n, [0] * n, [0] * n          # per-item scalars, assigned column-wise below
        init_off, init_nhet = [-1] * n, [0] * n
        io = 0
        for i in range(n):
            r = np.ascontiguousarray(reads_list[i], dtype=np.float64)
            assert r.ndim == 3
            U, N, A = r.shape
            na = default_na if default_na is not None else np.ascontiguousarray(n_alleles_list[i], dtype=np.int8)
            assert len(na) == N
            Us[i], Ns[i], As[i] = U, N, A
            if use_initial and initial_list[i] is not None:
                ini = np.ascontiguousarray(initial_list[i], dtype=np.int8)
                assert ini.ndim == 3 and ini.shape[0] == self.chains and ini.shape[1] == self.ploidy
                init_off[i], init_nhet[i] = io, ini.shape[2]
                ins.append(ini.ravel())
                io += ini.size
            rs.append(r.reshape(-1))
            ns.append(na)
            if use_counts:
                c = counts_list[i]
                c = np.ones(U, dtype=np.int64) if c is None else np.ascontiguousarray(c, dtype=np.int64)
                assert len(c) == U or U == 0
                cs.append(c[:U])
        U_, N_, A_ = (np.asarray(x, dtype=np.int64) for x in (Us, Ns, As))
        excl = lambda x: np.concatenate([[0], np.cumsum(x)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        g_sizes = self.chains * self.steps * self.ploidy * N_
        l_sizes = np.full(n, self.chains * self.steps, dtype=np.int64)
        items = np.zeros(n, dtype=ASSEMBLE_ITEM_DTYPE)
        items["reads_off"] = excl(U_ * N_ * A_)
        items["counts_off"] = excl(U_) if use_counts else 0
        items["nalleles_off"] = excl(N_)
        items["genotypes_off"], items["llks_off"] = excl(g_sizes), excl(l_sizes)
        items["n_reads"], items["n_pos"], items["max_allele"], items["ploidy"] = U_, N_, np.maximum(A_, 1), self.ploidy
        items["temps_off"], items["n_temps"] = 0, len(temps)
        items["seed"] = seed0 if seeds is None else (np.asarray(seeds, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32)
        items["inbreeding"] = np.nan if self.inbreeding is None else float(self.inbreeding)
        items["initial_off"], items["initial_nhet"] = init_off, init_nhet
        go, lo = int(g_sizes.sum()), int(l_sizes.sum())
        shapes = list(zip(Ns, items["genotypes_off"].tolist(), items["llks_off"].tolist()))
        nmax = max(Ns + [1])
        reads = np.concatenate(rs) if rs else np.zeros(0)
        nall = np.concatenate(ns) if ns else np.zeros(0, dtype=np.int8)
        counts = np.concatenate(cs) if use_counts and cs else None
        initial = np.concatenate(ins) if ins else None
        return dict(items=items, shapes=shapes, reads=reads, counts=counts, n_alleles=nall, initial=initial,
                    nmax=nmax, genotypes_len=go, llks_len=lo,
                    lens=(reads.size, 0 if counts is None else counts.size, nall.size,
                          0 if initial is None else initial.size))

    def _params(self, nmax, replay_words=None):
        table, lens = break_table(nmax, self.alpha, self.beta, self.n_intervals)
        return make_assemble_params(
            self.steps, self.chains, self.fix_homozygous, self.recombination_step_probability,
            self.partial_dosage_step_probability, self.dosage_step_probability, table, lens, self._temperatures(),
            replay_words=replay_words)

    def fit_batch(self, reads_list, counts_list=None, initial_list=None, n_alleles_list=None,
                  seeds=None, return_results=False, raw=False, replay_words=None):
        """Run ``fit`` for many items in one device call.

        n_alleles_list: per item allele counts (default: self.n_alleles for every item);
        seeds: per item seeds (default: self.random_seed for every item, like the CLIs);
        raw=True returns unsorted (genotypes, llks) arrays instead of GenotypeMultiTrace."""
        dev = self.device or default_device()
        n = len(reads_list)
        pk = self._pack(reads_list, counts_list, initial_list, n_alleles_list, seeds)
        items, shapes, go, lo = pk["items"], pk["shapes"], pk["genotypes_len"], pk["llks_len"]
        out_g = np.zeros(max(go, 1), dtype=np.int8)
        out_l = np.full(max(lo, 1), np.nan, dtype=np.float64)
        params, keep = self._params(pk["nmax"], replay_words)
        results = dev.assemble_call(
            items, params, pk["reads"], pk["counts"], pk["n_alleles"], pk["initial"], out_g, out_l,
            pk["lens"] + (go, lo))
        out = []
        for i, (N, g0, l0) in enumerate(shapes):
            raise_item_status(int(results["status"][i]), i if n > 1 else None)
            g = out_g[g0: g0 + self.chains * self.steps * self.ploidy * N].reshape(
                self.chains, self.steps, self.ploidy, N)
            l = out_l[l0: l0 + self.chains * self.steps].reshape(self.chains, self.steps)
            if N == 0:
                l[:] = np.nan  # no variable position: nothing was sampled (mcmc.py:188-199)
            out.append((g, l) if raw else GenotypeMultiTrace(g, l))
        if return_results:
            return out, results
        return out

    def _tally_helpers(self, items, burn, n, out):
        """Closures shared by the fit_posterior_* methods: tally descriptors / output arrays for a set of
        items, and the unpacking of the device tallies into TraceTally objects (returns the overflowed)."""
        def tally_items(idx, table):
            t = np.zeros(len(idx), dtype=TALLY_ITEM_DTYPE)
            pn = items["ploidy"][idx].astype(np.int64) * items["n_pos"][idx].astype(np.int64)
            t["genotypes_off"] = items["genotypes_off"][idx]
            t["n_pos"], t["ploidy"] = items["n_pos"][idx], items["ploidy"][idx]
            t["chains"], t["steps"], t["burn"], t["max_unique"] = self.chains, self.steps, burn, table
            t["states_off"] = np.concatenate([[0], np.cumsum(pn * table)[:-1]])
            t["tallies_off"] = np.arange(len(idx), dtype=np.int64) * table * self.chains
            return (t, np.zeros(max(int((pn * table).sum()), 1), dtype=np.int8),
                    np.zeros(max(len(idx) * table * self.chains, 1), dtype=np.int32),
                    np.zeros(max(len(idx) * table * self.chains, 1), dtype=np.int32))

        def collect(idx, t, tres, states, counts, first):
            over = []
            for k, i in enumerate(idx):
                if int(tres["status"][k]) == TALLY_OVERFLOW:
                    over.append(i)
                    continue
                raise_item_status(int(tres["status"][k]), i if n > 1 else None)
                u = int(tres["n_het"][k])
                P, N = int(items["ploidy"][i]), int(items["n_pos"][i])
                so, to = int(t["states_off"][k]), int(t["tallies_off"][k])
                out[i] = TraceTally(
                    states[so: so + u * P * N].reshape(u, P, N).copy(),
                    counts[to: to + u * self.chains].reshape(u, self.chains).astype(np.int64),
                    first[to: to + u * self.chains].reshape(u, self.chains).astype(np.int64))
            return over

        return tally_items, collect

    def fit_posterior_batch(self, reads_list, counts_list=None, burn=0, initial_list=None,
                            n_alleles_list=None, seeds=None, max_unique=128):
        """``fit(...).burn(burn)`` for many items with the traces kept on the device: returns one
        TraceTally per item (``.posterior()``, ``.split()``, ``.replicate_incongruence()`` behave
        like the burnt GenotypeMultiTrace of the reference, mchap/application/assemble.py:123-170).

        Only the tallies (distinct genotypes, counts and first occurrences per chain) cross the
        bus.  Items with more than ``max_unique`` distinct genotypes are tallied a second time from
        the trace still held on the device, with a table as large as the trace."""
        dev = self.device or default_device()
        n = len(reads_list)
        burn = int(burn)
        pk = self._pack(reads_list, counts_list, initial_list, n_alleles_list, seeds)
        items = pk["items"]
        params, keep = self._params(pk["nmax"])
        kept = max(self.steps - max(burn, 0), 0) * self.chains
        out = [None] * n

        tally_items, collect = self._tally_helpers(items, burn, n, out)

        everything = np.arange(n)
        table = max(1, min(int(max_unique), max(kept, 1)))
        t, states, counts, first = tally_items(everything, table)
        results, tres = dev.assemble_tally_call(
            items, t, params, pk["reads"], pk["counts"], pk["n_alleles"], pk["initial"],
            pk["lens"] + (pk["genotypes_len"], pk["llks_len"]), states, counts, first)
        for i in range(n):
            raise_item_status(int(results["status"][i]), i if n > 1 else None)
        over = collect(everything, t, tres, states, counts, first)
        if over and kept <= 8192:
            # second look at the same device-resident trace with a table that cannot overflow
            idx = np.array(over)
            t, states, counts, first = tally_items(idx, kept)
            tres = dev.trace_tally_call(t, None, 0, states, counts, first, mem_in=L.MEM_LAST_TRACE)
            over = collect(idx, t, tres, states, counts, first)
        if over:
            # more distinct genotypes than the device table can hold: bring those traces to the host
            traces = self.fit_batch(sub2(reads_list, over), sub2(counts_list, over), sub2(initial_list, over),
                                    sub2(n_alleles_list, over), sub2(seeds, over))
            for i, tr in zip(over, traces):
                out[i] = TraceTally.from_trace(tr.burn(burn))
        return out

    def fit_posterior_from_calls_batch(self, calls_list, probs_list, burn=0, n_alleles_list=None, seeds=None,
                                       max_unique=128, error_factor=3):
        """The whole per-sample device path in one call: integer allele calls + P(call correct)
        (mchap/application/baseclass.py:194-209 encodes and de-duplicates them on the host) ->
        de novo assembly -> tallies of the burnt trace.  Equivalent to
        ``encode_unique_reads_batch`` followed by ``fit_posterior_batch``; the encoded reads and the
        traces never leave the device.  Returns (tallies, n_unique_reads)."""
        from ..encoding import ENCODE_ITEM_DTYPE, encode_unique_reads_batch

        dev = self.device or default_device()
        n = len(calls_list)
        burn = int(burn)
        temps = self._temperatures()
        cs, ps, ns = [], [], []
        Rs, Ns, As = [0] * n, [0] * n, [0] * n
        default_na = None if n_alleles_list is not None else np.ascontiguousarray(self.n_alleles, dtype=np.int8)
        for i in range(n):
            c = np.ascontiguousarray(calls_list[i], dtype=np.int8)
            assert c.ndim == 2
            R, N = c.shape
            p = np.ascontiguousarray(np.broadcast_to(np.asarray(probs_list[i], dtype=np.float64), (R, N)))
            na = default_na if default_na is not None else np.ascontiguousarray(n_alleles_list[i], dtype=np.int8)
            assert len(na) == N
            Rs[i], Ns[i], As[i] = R, N, (int(na.max()) if N > 0 else 0)
            cs.append(c.reshape(-1))
            ps.append(p.reshape(-1))
            ns.append(na)
        R_, N_, A_ = (np.asarray(x, dtype=np.int64) for x in (Rs, Ns, As))
        excl = lambda x: np.concatenate([[0], np.cumsum(x)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        enc = np.zeros(n, dtype=ENCODE_ITEM_DTYPE)
        enc["calls_off"] = enc["probs_off"] = excl(R_ * N_)
        enc["nalleles_off"] = excl(N_)
        enc["reads_off"], enc["counts_off"] = excl(R_ * N_ * A_), excl(R_)
        enc["n_reads"], enc["n_pos"], enc["max_allele"] = R_, N_, A_
        items = np.zeros(n, dtype=ASSEMBLE_ITEM_DTYPE)
        g_sizes = self.chains * self.steps * self.ploidy * N_
        items["genotypes_off"] = excl(g_sizes)
        items["llks_off"] = np.arange(n, dtype=np.int64) * (self.chains * self.steps)
        items["n_pos"], items["max_allele"], items["ploidy"] = N_, np.maximum(A_, 1), self.ploidy
        items["temps_off"], items["n_temps"] = 0, len(temps)
        items["seed"] = self._seed() if seeds is None else (np.asarray(seeds, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32)
        items["inbreeding"] = np.nan if self.inbreeding is None else float(self.inbreeding)
        items["initial_off"] = -1
        go, lo = int(g_sizes.sum()), n * self.chains * self.steps
        nmax = max(Ns + [1])
        calls = np.concatenate(cs) if cs else np.zeros(0, dtype=np.int8)
        probs = np.concatenate(ps) if ps else np.zeros(0)
        nall = np.concatenate(ns) if ns else np.zeros(0, dtype=np.int8)
        params, keep = self._params(nmax)
        kept = max(self.steps - max(burn, 0), 0) * self.chains
        out = [None] * n
        tally_items, collect = self._tally_helpers(items, burn, n, out)
        everything = np.arange(n)
        table = max(1, min(int(max_unique), max(kept, 1)))
        t, states, counts, first = tally_items(everything, table)
        eres = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        results = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        tres = np.zeros(n, dtype=ITEM_RESULT_DTYPE)
        rc = dev._lib.mchb_encode_assemble_tally_batch(
            dev._h, C.byref(params), _ptr(enc), _ptr(items), _ptr(t), n, _ptr(calls), calls.size, _ptr(probs),
            probs.size, _ptr(nall), nall.size, float(error_factor), go, lo, _ptr(states), states.size, _ptr(counts),
            _ptr(first), counts.size, _ptr(eres), _ptr(results), _ptr(tres))
        dev._check(rc)
        for i in range(n):
            raise_item_status(int(eres["status"][i]), i if n > 1 else None)
            raise_item_status(int(results["status"][i]), i if n > 1 else None)
        over = collect(everything, t, tres, states, counts, first)
        if over and kept <= 8192:
            idx = np.array(over)
            t, states, counts, first = tally_items(idx, kept)
            tres = dev.trace_tally_call(t, None, 0, states, counts, first, mem_in=L.MEM_LAST_TRACE)
            over = collect(idx, t, tres, states, counts, first)
        if over:
            pairs = encode_unique_reads_batch(sub2(calls_list, over), sub2(probs_list, over),
                                              [self.n_alleles] * len(over) if n_alleles_list is None
                                              else sub2(n_alleles_list, over), error_factor, dev)
            traces = self.fit_batch([r for r, _ in pairs], [c for _, c in pairs], None, sub2(n_alleles_list, over),
                                    sub2(seeds, over))
            for i, tr in zip(over, traces):
                out[i] = TraceTally.from_trace(tr.burn(burn))
        return out, eres["n_het"].astype(np.int64)
