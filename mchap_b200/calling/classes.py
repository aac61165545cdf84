"""``CallingMCMC`` with the reference's constructor and ``fit`` signature
(mchap/calling/classes.py:14-124), executed by the CUDA kernel ``call_mcmc_kernel``, and the trace
containers it returns (GenotypeAllelesMultiTrace 127-274, PosteriorGenotypeAllelesDistribution
300-368) as vectorised numpy.
"""
from dataclasses import dataclass

import numpy as np

from .. import _lib as L
from ..api import TALLY_ITEM_DTYPE, CallBatch, count_genotypes, default_device, item_status_error
from ..assemble.classes import unique_first_occurrence

__all__ = ["CallingMCMC", "GenotypeAllelesMultiTrace", "PosteriorGenotypeAllelesDistribution", "AllelesTraceTally"]


@dataclass
class PosteriorGenotypeAllelesDistribution(object):
    genotypes: np.ndarray
    probabilities: np.ndarray

    def mode(self, genotype_support=False):
        if not genotype_support:
            i = int(np.argmax(self.probabilities))
            return self.genotypes[i], self.probabilities[i]
        labels = np.zeros(len(self.genotypes), dtype=int)
        seen, totals = {}, {}
        for i, gen in enumerate(self.genotypes):
            key = np.unique(gen).tobytes()  # sorted alleles: unique == first-occurrence order
            label = seen.setdefault(key, i)
            labels[i] = label
            totals[label] = totals.get(label, 0.0) + self.probabilities[i]
        names = list(totals.keys())
        best = names[int(np.argmax([totals[k] for k in names]))]
        keep = labels == best
        gens, probs = self.genotypes[keep], self.probabilities[keep]
        j = int(np.argmax(probs))
        return gens[j], probs[j], probs.sum()

    def as_array(self, n_alleles):
        """Dense VCF-order probability vector (calling/utils.py:60-86)."""
        from ..jitutils import genotype_alleles_as_index

        _, ploidy = self.genotypes.shape
        out = np.zeros(count_genotypes(n_alleles, ploidy), dtype=np.float64)
        # a handful of distinct genotypes per trace: host integers (the array form of the ranking
        # runs on the device, jitutils.genotypes_as_indices)
        for gen, p in zip(self.genotypes, self.probabilities):
            out[genotype_alleles_as_index(gen)] = p
        return out

    def allele_frequencies(self, dosage=False):
        _, ploidy = self.genotypes.shape
        alleles = np.unique(self.genotypes)
        freqs = np.zeros(len(alleles))
        occur = np.zeros(len(alleles))
        for gen, p in zip(self.genotypes, self.probabilities):
            idx = np.searchsorted(alleles, gen)
            np.add.at(freqs, idx, p)
            occur[np.unique(idx)] += p
        if not dosage:
            freqs /= ploidy
        return alleles, freqs, occur


@dataclass
class GenotypeAllelesMultiTrace(object):
    """genotypes int[n_chains, n_steps, ploidy] (sorted allele indices), llks f64[n_chains, n_steps]."""

    genotypes: np.ndarray
    llks: np.ndarray
    n_allele: int

    def relabel(self, labels):
        labels = np.asarray(labels)
        return type(self)(labels[self.genotypes], self.llks, labels.max() + 1)

    def burn(self, n):
        return type(self)(self.genotypes[:, n:], self.llks[:, n:], self.n_allele)

    def posterior(self):
        n_chain, n_step = self.genotypes.shape[:2]
        flat = self.genotypes.reshape((n_chain * n_step,) + self.genotypes.shape[2:])
        states, counts, _, _ = unique_first_occurrence(flat)
        probs = counts / np.sum(counts)
        idx = np.flip(np.argsort(probs))
        return PosteriorGenotypeAllelesDistribution(states[idx], probs[idx])

    def split(self):
        for g, l in zip(self.genotypes, self.llks):
            yield type(self)(g[None, ...], l[None, ...], self.n_allele)

    def replicate_incongruence(self, threshold=0.6):
        return _alleles_incongruence([chain.posterior() for chain in self.split()], threshold)

    def posterior_frequencies(self):
        """(frequencies, counts, occurrence) over all recorded steps (classes.py:258-297)."""
        n_chain, n_step, ploidy = self.genotypes.shape
        g = self.genotypes.reshape(-1, ploidy)
        counts = np.bincount(g.ravel(), minlength=self.n_allele).astype(np.float64)
        first = np.ones_like(g, dtype=bool)
        first[:, 1:] = g[:, 1:] != g[:, :-1]  # rows are sorted: first copy == differs from the left
        if not (np.diff(g, axis=1) >= 0).all():
            first = np.array([[a not in row[:i] for i, a in enumerate(row)] for row in g])
        occurrence = np.bincount(g[first], minlength=self.n_allele).astype(np.float64)
        n_obs = n_chain * n_step
        counts /= n_obs
        occurrence /= n_obs
        return counts / ploidy, counts, occurrence


def _alleles_incongruence(chain_posteriors, threshold):
    """calling/classes.py:221-256 of the reference on the per-chain posteriors."""
    modes = [p.mode(genotype_support=True) for p in chain_posteriors]
    alleles = [m[0] for m in modes if m[-1] >= threshold]
    if len({a.tobytes() for a in alleles}) <= 1:
        return 0
    ploidy = len(alleles[0])
    return 2 if len(set(np.array(alleles).ravel())) > ploidy else 1


@dataclass
class AllelesTraceTally(object):
    """What a burnt GenotypeAllelesMultiTrace is reduced to before its summaries are taken,
    computed on the device by mchb_call_trace_tally_batch: distinct genotypes in order of first
    occurrence in the chain-major flattened trace, occurrences and first step per chain.
    The methods return what the same methods of the burnt trace return (reference:
    mchap/calling/classes.py:147-297)."""

    states: np.ndarray   # int[n_unique, ploidy]
    counts: np.ndarray   # int64[n_unique, n_chains]
    first: np.ndarray    # int64[n_unique, n_chains]
    n_allele: int

    @classmethod
    def from_trace(cls, trace):
        n_chain, n_step, ploidy = trace.genotypes.shape
        flat = trace.genotypes.reshape(n_chain * n_step, ploidy)
        states, _, _, labels = unique_first_occurrence(flat)
        labels = labels.reshape(n_chain, n_step)
        counts = np.zeros((len(states), n_chain), dtype=np.int64)
        first = np.full((len(states), n_chain), -1, dtype=np.int64)
        for c in range(n_chain):
            u, idx, cnt = np.unique(labels[c], return_index=True, return_counts=True)
            counts[u, c] = cnt
            first[u, c] = idx
        return cls(np.array(states), counts, first, trace.n_allele)

    def relabel(self, labels):
        labels = np.asarray(labels)
        return type(self)(labels[self.states], self.counts, self.first, labels.max() + 1)

    def posterior(self):
        totals = self.counts.sum(axis=1)
        probs = totals / np.sum(totals)
        idx = np.flip(np.argsort(probs))
        return PosteriorGenotypeAllelesDistribution(self.states[idx], probs[idx])

    def split(self):
        for c in range(self.counts.shape[1]):
            seen = np.flatnonzero(self.counts[:, c] > 0)
            seen = seen[np.argsort(self.first[seen, c], kind="stable")]
            yield type(self)(self.states[seen], self.counts[seen, c:c + 1], self.first[seen, c:c + 1], self.n_allele)

    def replicate_incongruence(self, threshold=0.6):
        return _alleles_incongruence([t.posterior() for t in self.split()], threshold)

    def posterior_frequencies(self):
        """(frequencies, counts, occurrence): the reference adds 1.0 per recorded allele copy
        (classes.py:277-297); the sums are integers below 2^53, so adding whole tallies is exact."""
        totals = self.counts.sum(axis=1)
        ploidy = self.states.shape[1]
        counts = np.zeros(self.n_allele)
        occurrence = np.zeros(self.n_allele)
        for gen, t in zip(self.states, totals):
            np.add.at(counts, gen, float(t))
            occurrence[np.unique(gen)] += float(t)
        n_obs = np.sum(totals)
        counts /= n_obs
        occurrence /= n_obs
        return counts / ploidy, counts, occurrence


@dataclass
class CallingMCMC(object):
    ploidy: int
    haplotypes: np.ndarray
    prior: tuple = None
    steps: int = 1000
    chains: int = 2
    random_seed: int = None
    step_type: str = "Gibbs"
    device: object = None

    @classmethod
    def parameterize(cls, *args, **kwargs):
        return cls(*args, **kwargs)

    def _step_type(self):
        if self.step_type == "Gibbs":
            return 0
        if self.step_type == "Metropolis-Hastings":
            return 1
        raise ValueError('MCMC step type must be "Gibbs" or "Metropolis-Hastings"')

    def fit(self, reads, read_counts=None, initial=None):
        """Same contract as the reference's fit -> GenotypeAllelesMultiTrace."""
        reads = np.asarray(reads)
        if reads.shape[1] == 0:  # classes.py:75-82: no variants, reference allele only
            assert len(self.haplotypes) == 1
            genotypes = np.zeros((self.chains, self.steps, self.ploidy), dtype=np.int8)
            llks = np.full((self.chains, self.steps), np.nan)
            return GenotypeAllelesMultiTrace(genotypes, llks, len(self.haplotypes))
        return self.fit_batch([reads], [read_counts], None if initial is None else [initial])[0]

    def _prepare(self, reads_list, counts_list, initial_list, haplotypes_list, priors, seeds, ploidy_list):
        """CallBatch + output offsets, seeds and initial genotypes of a batch."""
        n = len(reads_list)
        haps = [self.haplotypes] * n if haplotypes_list is None else haplotypes_list
        prs = ([self.prior] * n if self.prior is not None else None) if priors is None else priors
        ploidy = (np.full(n, int(self.ploidy), dtype=np.int64) if ploidy_list is None
                  else np.asarray(ploidy_list, dtype=np.int64).reshape(n))
        batch = CallBatch(reads_list, haps, ploidy, counts_list, prs)
        seed0 = int(self.random_seed) & 0xFFFFFFFF if self.random_seed is not None else int(
            np.random.randint(0, 2 ** 32, dtype=np.uint64))
        items = batch.items
        per = self.chains * self.steps
        a_off = np.zeros(n, dtype=np.int64)
        if n > 1:
            np.cumsum(per * ploidy[:-1], out=a_off[1:])
        items["gl_off"] = a_off
        items["hap_out_off"] = np.arange(n, dtype=np.int64) * per
        sd = np.full(n, seed0, dtype=np.uint32) if seeds is None else np.asarray(seeds, dtype=np.uint64).astype(np.uint32)
        items["reserved"] = sd.view(np.int32)
        pmax = int(ploidy.max()) if n else 1
        init = None
        if initial_list is not None and any(i is not None for i in initial_list):
            init = np.full((n, pmax), -1, dtype=np.int32)
            for i, v in enumerate(initial_list):
                if v is not None:
                    init[i, :ploidy[i]] = np.asarray(v, dtype=np.int32)
        return batch, haps, prs, ploidy, pmax, init

    def fit_batch(self, reads_list, counts_list=None, initial_list=None, haplotypes_list=None, priors=None,
                  seeds=None, return_results=False, replay_words=None, ploidy_list=None, errors="raise"):
        """``fit`` for many items in one device call.  haplotypes_list / priors / ploidy_list default
        to the model's haplotypes / prior / ploidy for every item; seeds default to random_seed like
        the CLI; errors="return" stores a failing item's exception in its slot instead of raising."""
        assert errors in ("raise", "return")
        dev = self.device or default_device()
        n = len(reads_list)
        from ..assemble.mcmc import split_by_seeds

        split = split_by_seeds(
            self, "fit_batch", n, seeds,
            dict(reads_list=reads_list, counts_list=counts_list, initial_list=initial_list,
                 haplotypes_list=haplotypes_list, priors=priors, seeds=seeds, ploidy_list=ploidy_list),
            dict(return_results=return_results, replay_words=replay_words, errors=errors))
        if split is not None:
            return split
        batch, haps, prs, ploidy, pmax, init = self._prepare(reads_list, counts_list, initial_list, haplotypes_list,
                                                             priors, seeds, ploidy_list)
        out = dev.call_mcmc(batch, self.steps, self.chains, self._step_type(), init, pmax, replay_words)
        per = self.chains * self.steps
        res = [None] * n
        for i in range(n):
            exc = item_status_error(int(out["results"]["status"][i]), i if n > 1 else None)
            if exc is not None:
                if errors == "raise":
                    raise exc
                res[i] = exc
                continue
            P, a0 = int(ploidy[i]), int(batch.items["gl_off"][i])
            g = out["alleles"][a0:a0 + per * P].reshape(self.chains, self.steps, P)
            l = out["llks"][i * per:(i + 1) * per].reshape(self.chains, self.steps)
            res[i] = GenotypeAllelesMultiTrace(g, l, len(haps[i]))
        if return_results:
            return res, out["results"]
        return res

    def fit_posterior_batch(self, reads_list, counts_list=None, burn=0, initial_list=None, haplotypes_list=None,
                            priors=None, seeds=None, max_unique=128, ploidy_list=None, errors="raise"):
        """``fit(...).burn(burn)`` for many items with the traces kept on the device: one
        AllelesTraceTally per item (mchap/application/call.py:134-182 consumes exactly its methods)."""
        assert errors in ("raise", "return")
        dev = self.device or default_device()
        n = len(reads_list)
        st = self._step_type()
        burn = int(burn)
        batch, haps, prs, ploidy, pmax, init = self._prepare(reads_list, counts_list, initial_list, haplotypes_list,
                                                             priors, seeds, ploidy_list)
        items = batch.items
        idx = np.arange(n, dtype=np.int64)
        kept = max(self.steps - max(burn, 0), 0) * self.chains
        out = [None] * n

        def settle(i, status):
            exc = item_status_error(int(status), i if n > 1 else None)
            if exc is None:
                return True
            if errors == "raise":
                raise exc
            out[i] = exc
            return False

        def tally_items(sel, table):
            t = np.zeros(len(sel), dtype=TALLY_ITEM_DTYPE)
            P_ = ploidy[sel]
            t["genotypes_off"] = items["gl_off"][sel]
            t["n_pos"], t["ploidy"] = 1, P_
            t["chains"], t["steps"], t["burn"], t["max_unique"] = self.chains, self.steps, burn, table
            so = np.zeros(len(sel), dtype=np.int64)
            if len(sel) > 1:
                np.cumsum(table * P_[:-1], out=so[1:])
            t["states_off"] = so
            t["tallies_off"] = np.arange(len(sel), dtype=np.int64) * table * self.chains
            return (t, np.empty(max(int((table * P_).sum()), 1), dtype=np.int32),
                    np.empty(max(len(sel) * table * self.chains, 1), dtype=np.int32),
                    np.empty(max(len(sel) * table * self.chains, 1), dtype=np.int32))

        def collect(sel, t, tres, states, counts, first):
            over = []
            counts, first = counts.astype(np.int64), first.astype(np.int64)   # once, not per item
            st, nu = tres["status"].tolist(), tres["n_het"].tolist()
            so_, to_ = t["states_off"].tolist(), t["tallies_off"].tolist()
            P_ = ploidy[np.asarray(sel, dtype=np.int64)].tolist()
            C_ = self.chains
            for k, i in enumerate(np.asarray(sel).tolist()):
                if isinstance(out[i], BaseException):
                    continue
                if st[k] != 0:
                    if st[k] == L.ITEM_TALLY_OVERFLOW:
                        over.append(i)
                    else:
                        settle(i, st[k])
                    continue
                u, P, so, to = nu[k], P_[k], so_[k], to_[k]
                out[i] = AllelesTraceTally(states[so: so + u * P].reshape(u, P), counts[to: to + u * C_].reshape(u, C_),
                                           first[to: to + u * C_].reshape(u, C_), len(haps[i]))
            return over

        table = max(1, min(int(max_unique), max(kept, 1)))
        t, states, counts, first = tally_items(idx, table)
        results, tres = dev.call_mcmc_tally(batch, t, self.steps, self.chains, st, states, counts, first, init, pmax)
        for i in range(n):
            settle(i, results["status"][i])
        over = collect(idx, t, tres, states, counts, first)
        if over and kept <= 8192:
            sel = np.array(over)
            t, states, counts, first = tally_items(sel, kept)
            tres = dev.call_trace_tally_call(t, None, 0, states, counts, first, mem_in=L.MEM_LAST_TRACE)
            over = collect(sel, t, tres, states, counts, first)
        if over:
            pick = lambda lst: None if lst is None else [lst[i] for i in over]
            traces = self.fit_batch(pick(reads_list), pick(counts_list), pick(initial_list), [haps[i] for i in over],
                                    pick(prs), pick(seeds), ploidy_list=ploidy[over])
            for i, tr in zip(over, traces):
                out[i] = AllelesTraceTally.from_trace(tr.burn(burn))
        return out
