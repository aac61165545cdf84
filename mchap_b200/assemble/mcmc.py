"""``DenovoMCMC`` with the reference's constructor and ``fit`` signature
(reference: mchap/assemble/mcmc.py:24-161), executed by the CUDA kernel ``assemble_kernel``.

``fit`` handles one (locus, sample) item like the reference; the ``*_batch`` methods send many
items in one device call (the reason this package exists).  There is no CPU path.

Batch conventions shared by every ``*_batch`` method:

* per-item overrides — ``n_alleles_list``, ``seeds``, ``ploidy_list``, ``inbreeding_list``
  (``None`` entries = flat prior), ``temperatures_list`` — default to the model's own parameters
  for every item, so that samples of different ploidy / prior / temperature ladder (pools, per-sample
  CLI files) share one launch;
* ``errors="raise"`` (default) re-raises the first per-item device status with the reference's
  exception type; ``errors="return"`` puts the exception object in that item's slot of the result
  list instead, so one bad item does not discard the finished results of the others.

Limits of the CUDA kernels (``mchb_get_limits``; items outside them get ``NotImplementedError``, never
an approximation): ploidy <= 16, variable positions x bits per allele <= 64, <= 1024 distinct reads
per item (and tables that fit one CTA's shared memory), <= 8 temperatures.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib as L
from ..api import (ASSEMBLE_ITEM_DTYPE, ITEM_RESULT_DTYPE, TALLY_ITEM_DTYPE, _ptr, default_device,
                   item_status_error, make_assemble_params)
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


_BREAK_TABLES = {}


def break_table(n_pos, alpha=1.0, beta=3.0, n_intervals=None):
    """Row n = break-point distribution used when n positions stay variable (mcmc.py:211-217).
    Tables are memoised: the CLIs build the same one for every block of loci."""
    key = (int(n_pos), float(alpha), float(beta), n_intervals)
    hit = _BREAK_TABLES.get(key)
    if hit is not None:
        return hit
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
    _BREAK_TABLES[key] = (table, lens)
    return table, lens


def sub2(lst, idx):
    return None if lst is None else [lst[i] for i in idx]


def _excl(x):
    """Exclusive prefix sum (element offsets of back-to-back items)."""
    x = np.asarray(x, dtype=np.int64)
    out = np.zeros(len(x), dtype=np.int64)
    if len(x) > 1:
        np.cumsum(x[:-1], out=out[1:])
    return out


def _f64c(a):
    if type(a) is np.ndarray and a.dtype == np.float64 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _flat(arrays, dtype):
    return np.concatenate(arrays, axis=None) if len(arrays) else np.zeros(0, dtype=dtype)


def _sort_temperatures(temperatures):
    temps = np.sort(np.asarray(temperatures, dtype=np.float64).ravel())
    assert temps[0] >= 0.0
    assert temps[-1] == 1.0
    return temps


# Every distinct seed of a call needs its own pre-generated MT19937 stream on the device (a few MB);
# batches with more distinct per-item seeds than this run as consecutive sub-batches.
MAX_DISTINCT_SEEDS = 1024


def split_by_seeds(model, name, n, seeds, lists, kw):
    """Run ``model.<name>`` over consecutive sub-batches of at most MAX_DISTINCT_SEEDS items when the
    per-item ``seeds`` hold more distinct values than that; returns None when no split is needed.
    Results are joined in order: a list, or (list, array, ...)."""
    if seeds is None or n <= MAX_DISTINCT_SEEDS or len(np.unique(np.asarray(seeds))) <= MAX_DISTINCT_SEEDS:
        return None
    outs = []
    for a in range(0, n, MAX_DISTINCT_SEEDS):
        b = min(n, a + MAX_DISTINCT_SEEDS)
        sub = {k: (None if v is None else v[a:b]) for k, v in lists.items()}
        outs.append(getattr(model, name)(**sub, **kw))
    if isinstance(outs[0], tuple):
        joined = []
        for o in outs:
            joined.extend(o[0])
        return (joined,) + tuple(np.concatenate([o[j] for o in outs]) for j in range(1, len(outs[0])))
    joined = []
    for o in outs:
        joined.extend(o)
    return joined


def _settle(out, i, status, n, errors):
    """Handle the device status of item i: returns True when the item is fine; otherwise raises
    (errors='raise') or stores the exception in out[i] (errors='return')."""
    exc = item_status_error(int(status), i if n > 1 else None)
    if exc is None:
        return True
    if errors == "raise":
        raise exc
    out[i] = exc
    return False


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
        return _sort_temperatures(self.temperatures)

    def _seed(self):
        if self.random_seed is not None:
            return int(self.random_seed) & 0xFFFFFFFF
        return int(np.random.randint(0, 2 ** 32, dtype=np.uint64))

    def fit(self, reads, read_counts=None, initial=None):
        """Same contract as the reference's fit: returns a GenotypeMultiTrace with genotypes
        int8[chains, steps, ploidy, n_positions] (haplotypes sorted per step) and llks."""
        res = self.fit_batch([reads], [read_counts], None if initial is None else [initial])
        return res[0]

    # ------------------------------------------------------------------ packing
    def _item_params(self, n, seeds, ploidy_list, inbreeding_list, temperatures_list):
        """Per-item ploidy / inbreeding / temperature ladder / seed columns + the temperature pool."""
        ploidy = (np.full(n, int(self.ploidy), dtype=np.int64) if ploidy_list is None
                  else np.asarray(ploidy_list, dtype=np.int64).reshape(n))
        if inbreeding_list is None:
            inb = np.full(n, np.nan if self.inbreeding is None else float(self.inbreeding))
        else:
            inb = np.array([np.nan if v is None else float(v) for v in inbreeding_list], dtype=np.float64).reshape(n)
        if temperatures_list is None:
            pool = self._temperatures()
            t_off = np.zeros(n, dtype=np.int64)
            t_len = np.full(n, len(pool), dtype=np.int64)
        else:
            ladders, where, t_off, t_len, pos = [], {}, np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64), 0
            for i, t in enumerate(temperatures_list):
                key = tuple(np.asarray(self.temperatures if t is None else t, dtype=np.float64).ravel().tolist())
                hit = where.get(key)
                if hit is None:
                    ladder = _sort_temperatures(key)
                    hit = where[key] = (pos, len(ladder))
                    ladders.append(ladder)
                    pos += len(ladder)
                t_off[i], t_len[i] = hit
            pool = np.concatenate(ladders) if ladders else np.ones(1)
        seed = (np.full(n, self._seed(), dtype=np.uint32) if seeds is None
                else (np.asarray(seeds, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32).reshape(n))
        return ploidy, inb, pool, t_off, t_len, seed

    def _pack(self, reads_list, counts_list, initial_list, n_alleles_list, seeds, ploidy_list=None,
              inbreeding_list=None, temperatures_list=None):
        """Flat input arrays + item descriptors of a batch (outputs laid out item after item).
        Column-wise: per-item Python work is a handful of list comprehensions, the bulk arrays are
        flattened and joined by one ``np.concatenate`` each."""
        n = len(reads_list)
        ploidy, inb, pool, t_off, t_len, seed = self._item_params(n, seeds, ploidy_list, inbreeding_list,
                                                                  temperatures_list)
        arrs = [_f64c(r) for r in reads_list]
        assert all(a.ndim == 3 for a in arrs), "reads must be [n_reads, n_positions, max_allele] per item"
        shp = np.array([a.shape for a in arrs], dtype=np.int64).reshape(n, 3)
        U_, N_, A_ = shp[:, 0], shp[:, 1], shp[:, 2]
        if n_alleles_list is None:
            default_na = np.ascontiguousarray(self.n_alleles, dtype=np.int8)
            assert n == 0 or (N_ == len(default_na)).all()
            nall = np.tile(default_na, n)
        else:
            nas = [np.asarray(a, dtype=np.int8) for a in n_alleles_list]
            assert [len(a) for a in nas] == N_.tolist()
            nall = _flat(nas, np.int8)
        use_counts = counts_list is not None and any(c is not None for c in counts_list)
        counts = None
        if use_counts:
            cs = [np.ones(u, dtype=np.int64) if c is None else np.asarray(c, dtype=np.int64)[:u]
                  for c, u in zip(counts_list, U_.tolist())]
            assert [len(c) for c in cs] == U_.tolist()
            counts = _flat(cs, np.int64)
        use_initial = initial_list is not None and any(i is not None for i in initial_list)
        init_off = np.full(n, -1, dtype=np.int64)
        init_nhet = np.zeros(n, dtype=np.int64)
        initial = None
        if use_initial:
            ins, io = [], 0
            for i, v in enumerate(initial_list):
                if v is None:
                    continue
                ini = np.ascontiguousarray(v, dtype=np.int8)
                assert ini.ndim == 3 and ini.shape[0] == self.chains and ini.shape[1] == ploidy[i]
                init_off[i], init_nhet[i] = io, ini.shape[2]
                ins.append(ini)
                io += ini.size
            initial = _flat(ins, np.int8)
        g_sizes = self.chains * self.steps * ploidy * N_
        l_sizes = np.full(n, self.chains * self.steps, dtype=np.int64)
        items = np.zeros(n, dtype=ASSEMBLE_ITEM_DTYPE)
        items["reads_off"] = _excl(U_ * N_ * A_)
        items["counts_off"] = _excl(U_) if use_counts else 0
        items["nalleles_off"] = _excl(N_)
        items["genotypes_off"], items["llks_off"] = _excl(g_sizes), _excl(l_sizes)
        items["n_reads"], items["n_pos"], items["max_allele"], items["ploidy"] = U_, N_, np.maximum(A_, 1), ploidy
        items["temps_off"], items["n_temps"] = t_off, t_len
        items["seed"], items["inbreeding"] = seed, inb
        items["initial_off"], items["initial_nhet"] = init_off, init_nhet
        go, lo = int(g_sizes.sum()), int(l_sizes.sum())
        reads = _flat(arrs, np.float64)
        return dict(items=items, reads=reads, counts=counts, n_alleles=nall, initial=initial,
                    nmax=max(int(N_.max()) if n else 1, 1), genotypes_len=go, llks_len=lo, temperatures=pool,
                    lens=(reads.size, 0 if counts is None else counts.size, nall.size,
                          0 if initial is None else initial.size))

    def _params(self, nmax, replay_words=None, temperatures=None, sort_haplotypes=False):
        table, lens = break_table(nmax, self.alpha, self.beta, self.n_intervals)
        return make_assemble_params(
            self.steps, self.chains, self.fix_homozygous, self.recombination_step_probability,
            self.partial_dosage_step_probability, self.dosage_step_probability, table, lens,
            self._temperatures() if temperatures is None else temperatures, replay_words=replay_words,
            sort_haplotypes=sort_haplotypes)

    # ------------------------------------------------------------------ full traces
    def fit_batch(self, reads_list, counts_list=None, initial_list=None, n_alleles_list=None,
                  seeds=None, return_results=False, raw=False, replay_words=None, ploidy_list=None,
                  inbreeding_list=None, temperatures_list=None, errors="raise"):
        """Run ``fit`` for many items in one device call.

        raw=True returns the sampler's own (genotypes, llks) arrays (haplotypes in the order the
        chain holds them, as the reference's _denovo_assembler returns them) instead of
        GenotypeMultiTrace objects, whose steps have their haplotypes sorted — on the device, while
        the step is recorded.  The returned arrays are views of one page-locked batch buffer (recycled
        by the Device once every trace of the batch is garbage collected); the other keywords are
        described in the module docstring."""
        assert errors in ("raise", "return")
        dev = self.device or default_device()
        n = len(reads_list)
        split = split_by_seeds(
            self, "fit_batch", n, seeds,
            dict(reads_list=reads_list, counts_list=counts_list, initial_list=initial_list, n_alleles_list=n_alleles_list,
                 seeds=seeds, ploidy_list=ploidy_list, inbreeding_list=inbreeding_list, temperatures_list=temperatures_list),
            dict(return_results=return_results, raw=raw, replay_words=replay_words, errors=errors))
        if split is not None:
            return split
        pk = self._pack(reads_list, counts_list, initial_list, n_alleles_list, seeds, ploidy_list, inbreeding_list,
                        temperatures_list)
        items, go, lo = pk["items"], pk["genotypes_len"], pk["llks_len"]
        out_g = dev.trace_buffer(go, np.int8)
        out_l = dev.trace_buffer(lo, np.float64)
        params, keep = self._params(pk["nmax"], replay_words, pk["temperatures"], sort_haplotypes=not raw)
        results = dev.assemble_call(
            items, params, pk["reads"], pk["counts"], pk["n_alleles"], pk["initial"], out_g, out_l,
            pk["lens"] + (go, lo))
        out = [None] * n
        cs = self.chains * self.steps
        status = results["status"]
        bad = np.flatnonzero(status != 0)
        for i in bad:
            _settle(out, int(i), status[i], n, errors)
        ok = status == 0
        N_, P_ = items["n_pos"].tolist(), items["ploidy"].tolist()
        g0_, l0_ = items["genotypes_off"].tolist(), items["llks_off"].tolist()
        wrap = (lambda g, l: (g, l)) if raw else GenotypeMultiTrace._presorted
        chains, steps = self.chains, self.steps
        for i in range(n):
            if not ok[i]:
                continue
            N, P, g0, l0 = N_[i], P_[i], g0_[i], l0_[i]
            g = out_g[g0: g0 + cs * P * N].reshape(chains, steps, P, N)
            l = out_l[l0: l0 + cs].reshape(chains, steps)
            if N == 0:
                l[:] = np.nan  # no variable position: nothing was sampled (mcmc.py:188-199)
            out[i] = wrap(g, l)
        if return_results:
            return out, results
        return out

    # ------------------------------------------------------------------ tallies
    def _tally_plan(self, items, burn, out, errors):
        """Closures shared by the fit_posterior_* methods: tally descriptors / output arrays for a set of
        items, and the unpacking of the device tallies into TraceTally objects (returns the overflowed)."""
        n = len(items)

        def tally_items(idx, table):
            t = np.zeros(len(idx), dtype=TALLY_ITEM_DTYPE)
            pn = items["ploidy"][idx].astype(np.int64) * items["n_pos"][idx].astype(np.int64)
            t["genotypes_off"] = items["genotypes_off"][idx]
            t["n_pos"], t["ploidy"] = items["n_pos"][idx], items["ploidy"][idx]
            t["chains"], t["steps"], t["burn"], t["max_unique"] = self.chains, self.steps, burn, table
            t["states_off"] = _excl(pn * table)
            t["tallies_off"] = np.arange(len(idx), dtype=np.int64) * table * self.chains
            return (t, np.empty(max(int((pn * table).sum()), 1), dtype=np.int8),
                    np.empty(max(len(idx) * table * self.chains, 1), dtype=np.int32),
                    np.empty(max(len(idx) * table * self.chains, 1), dtype=np.int32))

        def collect(idx, t, tres, states, counts, first):
            over = []
            counts, first = counts.astype(np.int64), first.astype(np.int64)   # once, not per item
            st, nu = tres["status"].tolist(), tres["n_het"].tolist()
            P_, N_ = items["ploidy"][idx].tolist(), items["n_pos"][idx].tolist()
            so_, to_ = t["states_off"].tolist(), t["tallies_off"].tolist()
            C_ = self.chains
            for k, i in enumerate(np.asarray(idx).tolist()):
                if st[k] != 0:
                    if isinstance(out[i], BaseException):
                        continue
                    if st[k] == TALLY_OVERFLOW:
                        over.append(i)
                    else:
                        _settle(out, i, st[k], n, errors)
                    continue
                if out[i] is not None and isinstance(out[i], BaseException):
                    continue
                u, P, N, so, to = nu[k], P_[k], N_[k], so_[k], to_[k]
                out[i] = TraceTally(states[so: so + u * P * N].reshape(u, P, N),
                                    counts[to: to + u * C_].reshape(u, C_), first[to: to + u * C_].reshape(u, C_))
            return over

        return tally_items, collect

    def _second_look(self, dev, over, kept, tally_items, collect):
        """Items with more distinct genotypes than the first table: tally the trace that is still on
        the device again with a table that cannot overflow."""
        if over and kept <= 8192:
            idx = np.array(over)
            t, states, counts, first = tally_items(idx, kept)
            tres = dev.trace_tally_call(t, None, 0, states, counts, first, mem_in=L.MEM_LAST_TRACE)
            over = collect(idx, t, tres, states, counts, first)
        return over

    def fit_posterior_batch(self, reads_list, counts_list=None, burn=0, initial_list=None,
                            n_alleles_list=None, seeds=None, max_unique=128, ploidy_list=None,
                            inbreeding_list=None, temperatures_list=None, errors="raise"):
        """``fit(...).burn(burn)`` for many items with the traces kept on the device: returns one
        TraceTally per item (``.posterior()``, ``.split()``, ``.replicate_incongruence()`` behave
        like the burnt GenotypeMultiTrace of the reference, mchap/application/assemble.py:123-170).

        Only the tallies (distinct genotypes, counts and first occurrences per chain) cross the
        bus.  Items with more than ``max_unique`` distinct genotypes are tallied a second time from
        the trace still held on the device, with a table as large as the trace."""
        assert errors in ("raise", "return")
        dev = self.device or default_device()
        n = len(reads_list)
        burn = int(burn)
        split = split_by_seeds(
            self, "fit_posterior_batch", n, seeds,
            dict(reads_list=reads_list, counts_list=counts_list, initial_list=initial_list, n_alleles_list=n_alleles_list,
                 seeds=seeds, ploidy_list=ploidy_list, inbreeding_list=inbreeding_list, temperatures_list=temperatures_list),
            dict(burn=burn, max_unique=max_unique, errors=errors))
        if split is not None:
            return split
        pk = self._pack(reads_list, counts_list, initial_list, n_alleles_list, seeds, ploidy_list, inbreeding_list,
                        temperatures_list)
        items = pk["items"]
        params, keep = self._params(pk["nmax"], None, pk["temperatures"])
        kept = max(self.steps - max(burn, 0), 0) * self.chains
        out = [None] * n
        tally_items, collect = self._tally_plan(items, burn, out, errors)
        everything = np.arange(n)
        table = max(1, min(int(max_unique), max(kept, 1)))
        t, states, counts, first = tally_items(everything, table)
        results, tres = dev.assemble_tally_call(
            items, t, params, pk["reads"], pk["counts"], pk["n_alleles"], pk["initial"],
            pk["lens"] + (pk["genotypes_len"], pk["llks_len"]), states, counts, first)
        for i in range(n):
            _settle(out, i, results["status"][i], n, errors)
        over = collect(everything, t, tres, states, counts, first)
        over = self._second_look(dev, over, kept, tally_items, collect)
        if over:
            # more distinct genotypes than the device table can hold: bring those traces to the host
            traces = self.fit_batch(sub2(reads_list, over), sub2(counts_list, over), sub2(initial_list, over),
                                    sub2(n_alleles_list, over), sub2(seeds, over), ploidy_list=sub2(ploidy_list, over),
                                    inbreeding_list=sub2(inbreeding_list, over),
                                    temperatures_list=sub2(temperatures_list, over))
            for i, tr in zip(over, traces):
                out[i] = TraceTally.from_trace(tr.burn(burn))
        return out

    def fit_posterior_from_calls_batch(self, calls_list, probs_list, burn=0, n_alleles_list=None, seeds=None,
                                       max_unique=128, error_factor=3, ploidy_list=None, inbreeding_list=None,
                                       temperatures_list=None, errors="raise"):
        """The whole per-sample device path in one call: integer allele calls + P(call correct)
        (mchap/application/baseclass.py:194-209 encodes and de-duplicates them on the host) ->
        de novo assembly -> tallies of the burnt trace.  Equivalent to
        ``encode_unique_reads_batch`` followed by ``fit_posterior_batch``; the encoded reads and the
        traces never leave the device.  Returns (tallies, n_unique_reads)."""
        from ..encoding import ENCODE_ITEM_DTYPE, encode_unique_reads_batch

        assert errors in ("raise", "return")
        dev = self.device or default_device()
        n = len(calls_list)
        burn = int(burn)
        ploidy, inb, pool, t_off, t_len, seed = self._item_params(n, seeds, ploidy_list, inbreeding_list,
                                                                  temperatures_list)
        cs = [np.ascontiguousarray(c, dtype=np.int8) for c in calls_list]
        assert all(c.ndim == 2 for c in cs), "calls must be [n_reads, n_positions] per item"
        shp = np.array([c.shape for c in cs], dtype=np.int64).reshape(n, 2)
        R_, N_ = shp[:, 0], shp[:, 1]
        ps = [np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64), c.shape))
              for p, c in zip(probs_list, cs)]
        if n_alleles_list is None:
            default_na = np.ascontiguousarray(self.n_alleles, dtype=np.int8)
            assert n == 0 or (N_ == len(default_na)).all()
            nas = [default_na] * n
        else:
            nas = [np.asarray(a, dtype=np.int8) for a in n_alleles_list]
            assert [len(a) for a in nas] == N_.tolist()
        A_ = np.array([int(a.max()) if len(a) else 0 for a in nas], dtype=np.int64)
        enc = np.zeros(n, dtype=ENCODE_ITEM_DTYPE)
        enc["calls_off"] = enc["probs_off"] = _excl(R_ * N_)
        enc["nalleles_off"] = _excl(N_)
        enc["reads_off"], enc["counts_off"] = _excl(R_ * N_ * A_), _excl(R_)
        enc["n_reads"], enc["n_pos"], enc["max_allele"] = R_, N_, A_
        items = np.zeros(n, dtype=ASSEMBLE_ITEM_DTYPE)
        g_sizes = self.chains * self.steps * ploidy * N_
        items["genotypes_off"] = _excl(g_sizes)
        items["llks_off"] = np.arange(n, dtype=np.int64) * (self.chains * self.steps)
        items["n_pos"], items["max_allele"], items["ploidy"] = N_, np.maximum(A_, 1), ploidy
        items["temps_off"], items["n_temps"] = t_off, t_len
        items["seed"], items["inbreeding"] = seed, inb
        items["initial_off"] = -1
        go, lo = int(g_sizes.sum()), n * self.chains * self.steps
        nmax = max(int(N_.max()) if n else 1, 1)
        calls, probs, nall = _flat(cs, np.int8), _flat(ps, np.float64), _flat(nas, np.int8)
        params, keep = self._params(nmax, None, pool)
        kept = max(self.steps - max(burn, 0), 0) * self.chains
        out = [None] * n
        tally_items, collect = self._tally_plan(items, burn, out, errors)
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
            if _settle(out, i, eres["status"][i], n, errors):
                _settle(out, i, results["status"][i], n, errors)
        over = collect(everything, t, tres, states, counts, first)
        over = self._second_look(dev, over, kept, tally_items, collect)
        if over:
            pairs = encode_unique_reads_batch(sub2(calls_list, over), sub2(probs_list, over), [nas[i] for i in over],
                                              error_factor, dev)
            traces = self.fit_batch([r for r, _ in pairs], [c for _, c in pairs], None, [nas[i] for i in over],
                                    seed[over], ploidy_list=ploidy[over], inbreeding_list=[
                                        None if np.isnan(inb[i]) else inb[i] for i in over],
                                    temperatures_list=sub2(temperatures_list, over))
            for i, tr in zip(over, traces):
                out[i] = TraceTally.from_trace(tr.burn(burn))
        return out, eres["n_het"].astype(np.int64)
