"""Trace and posterior containers returned by the assemblers.

Host-side summaries with the behaviour of the reference's ``mchap/assemble/classes.py``
(GenotypeMultiTrace 247-376, PosteriorGenotypeDistribution 55-184,
GenotypeSupportDistribution 187-244), written over vectorised numpy instead of per-step Python
loops.  Orders that decide which of several equally probable genotypes is reported are kept:
unique elements in first-occurrence order (mset.py:242-284), then
``np.flip(np.argsort(probs))`` (classes.py:321-323).
"""
from dataclasses import dataclass

import numpy as np


def _row_keys(array):
    """One hashable bytes key per element of the outer dimension."""
    a = np.ascontiguousarray(array)
    flat = a.reshape(len(a), -1)
    return flat.view(np.dtype((np.void, flat.dtype.itemsize * flat.shape[1]))).ravel() if flat.shape[1] else None


def unique_first_occurrence(array):
    """(unique elements in order of first occurrence, their counts, first indices, inverse labels)."""
    a = np.ascontiguousarray(array)
    n = len(a)
    if n == 0:
        return a, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    keys = _row_keys(a)
    if keys is None:
        return a[:1], np.array([n]), np.array([0]), np.zeros(n, dtype=np.int64)
    _, first, inverse, counts = np.unique(keys, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return a[first[order]], counts[order], first[order], rank[inverse.ravel()]


def sort_haplotypes(genotypes):
    """Lexicographically sort the haplotypes (second-to-last axis) of every genotype
    (reference: encoding/integer/sequence.py:78-110 applied per step in classes.py:274-278)."""
    g = np.asarray(genotypes)
    n_pos = g.shape[-1]
    if n_pos == 0 or g.shape[-2] <= 1:
        return g.copy()
    keys = [g[..., j] for j in range(n_pos - 1, -1, -1)]  # last key is the primary one
    order = np.lexsort(keys, axis=-1)
    return np.take_along_axis(g, order[..., None], axis=-2)


@dataclass
class GenotypeSupportDistribution(object):
    """Genotypes made of the same set of haplotypes at different dosages."""

    genotypes: np.ndarray
    probabilities: np.ndarray

    def alleles(self):
        return unique_first_occurrence(self.genotypes[0])[0]

    def mode_genotype(self):
        i = int(np.argmax(self.probabilities))
        return self.genotypes[i], self.probabilities[i]

    def call_genotype_support(self, threshold=0.95):
        if np.max(self.probabilities) >= threshold:
            return self.mode_genotype()
        _, ploidy, n_pos = self.genotypes.shape
        result = np.full((ploidy, n_pos), -1, dtype=self.genotypes.dtype)
        order = list(np.argsort(-np.asarray(self.probabilities), kind="stable"))
        p = 0.0
        chosen = []
        while p < threshold and order:
            i = order.pop(0)
            p += self.probabilities[i]
            chosen.append(self.genotypes[i])
        # multiset intersection of the haplotypes of the chosen genotypes
        common = None
        for gen in chosen:
            haps, counts, _, _ = unique_first_occurrence(gen)
            tally = {h.tobytes(): (h, c) for h, c in zip(haps, counts)}
            if common is None:
                common = tally
            else:
                common = {k: (h, min(c, tally[k][1])) for k, (h, c) in common.items() if k in tally}
        row = 0
        for h, c in (common or {}).values():
            for _ in range(int(c)):
                result[row] = h
                row += 1
        return result, p


@dataclass
class PosteriorGenotypeDistribution(object):
    """Posterior over the distinct genotypes seen in a trace."""

    genotypes: np.ndarray
    probabilities: np.ndarray

    def mode(self):
        i = int(np.argmax(self.probabilities))
        return self.genotypes[i], self.probabilities[i]

    def mode_genotype_support(self):
        """Genotypes sharing the haplotype set (support) with the largest total probability."""
        n = len(self.genotypes)
        labels = np.zeros(n, dtype=int)
        seen = {}
        totals = {}
        for i, gen in enumerate(self.genotypes):
            key = unique_first_occurrence(gen)[0].tobytes()
            label = seen.setdefault(key, i)
            labels[i] = label
            totals[label] = totals.get(label, 0.0) + self.probabilities[i]
        names = list(totals.keys())
        best = names[int(np.argmax([totals[k] for k in names]))]
        keep = labels == best
        return GenotypeSupportDistribution(self.genotypes[keep], self.probabilities[keep])

    def allele_frequencies(self, dosage=False):
        n_gen, ploidy, n_base = self.genotypes.shape
        haps = self.genotypes.reshape(n_gen * ploidy, n_base)
        uhaps, _, _, inverse = unique_first_occurrence(haps)
        freqs = np.zeros(len(uhaps))
        occur = np.zeros(len(uhaps))
        inverse = inverse.reshape(n_gen, ploidy)
        for g in range(n_gen):
            labs, dose = np.unique(inverse[g], return_counts=True)
            freqs[labs] += self.probabilities[g] * dose
            occur[labs] += self.probabilities[g]
        if not dosage:
            freqs /= ploidy
        return uhaps, freqs, occur


@dataclass
class GenotypeMultiTrace(object):
    """Genotypes int8[n_chains, n_steps, ploidy, n_positions] and llks f64[n_chains, n_steps]."""

    genotypes: np.ndarray
    llks: np.ndarray

    def __post_init__(self):
        if self.genotypes is not None and self.genotypes.shape[-1] != 0:
            assert np.ndim(self.genotypes) == 4
            assert np.ndim(self.llks) == 2
            assert self.genotypes.shape[0:2] == self.llks.shape
            self.genotypes = sort_haplotypes(self.genotypes)
            self.llks = np.array(self.llks, copy=True)

    @classmethod
    def _presorted(cls, genotypes, llks):
        new = cls(None, None)
        new.genotypes = genotypes
        new.llks = llks
        return new

    def burn(self, n):
        return self._presorted(self.genotypes[:, n:], self.llks[:, n:])

    def posterior(self):
        n_chain, n_step, ploidy, n_base = self.genotypes.shape
        flat = self.genotypes.reshape(n_chain * n_step, ploidy, n_base)
        states, counts, _, _ = unique_first_occurrence(flat)
        probs = counts / np.sum(counts)
        idx = np.flip(np.argsort(probs))
        return PosteriorGenotypeDistribution(states[idx], probs[idx])

    def split(self):
        for g, l in zip(self.genotypes, self.llks):
            yield self._presorted(g[None, ...], l[None, ...])

    def replicate_incongruence(self, threshold=0.6):
        """0: chains agree; 1: different modes; 2: more than ploidy haplotypes among the modes."""
        return _replicate_incongruence([t.posterior() for t in self.split()], threshold)


def _replicate_incongruence(chain_posteriors, threshold):
    """classes.py:341-376 of the reference on the per-chain posteriors."""
    modes = [p.mode_genotype_support() for p in chain_posteriors]
    alleles = [m.alleles() for m in modes if m.probabilities.sum() >= threshold]
    if len({a.tobytes() for a in alleles}) <= 1:
        return 0
    ploidy = len(alleles[0])
    union = {}
    for a in alleles:  # multiset union: max multiplicity per haplotype
        haps, counts, _, _ = unique_first_occurrence(a)
        for h, c in zip(haps, counts):
            k = h.tobytes()
            union[k] = max(union.get(k, 0), int(c))
    return 2 if sum(union.values()) > ploidy else 1


@dataclass
class TraceTally(object):
    """What a burnt GenotypeMultiTrace is reduced to before any of its summaries is taken, computed
    on the device by mchb_trace_tally_batch: the distinct genotypes (haplotypes sorted) in order of
    first occurrence in the chain-major flattened trace, how often each occurs in every chain, and
    the step of its first occurrence in every chain (-1: never).

    ``posterior()``, ``split()`` and ``replicate_incongruence()`` return what the same methods of the
    burnt trace return (reference: mchap/assemble/classes.py:307-376)."""

    states: np.ndarray   # int8[n_unique, ploidy, n_positions]
    counts: np.ndarray   # int64[n_unique, n_chains]
    first: np.ndarray    # int64[n_unique, n_chains]

    @classmethod
    def from_trace(cls, trace):
        """Host-side tally of a (sorted, burnt) GenotypeMultiTrace: same content as the device's."""
        n_chain, n_step, ploidy, n_base = trace.genotypes.shape
        flat = trace.genotypes.reshape(n_chain * n_step, ploidy, n_base)
        states, _, _, labels = unique_first_occurrence(flat)
        labels = labels.reshape(n_chain, n_step)
        counts = np.zeros((len(states), n_chain), dtype=np.int64)
        first = np.full((len(states), n_chain), -1, dtype=np.int64)
        for c in range(n_chain):
            u, idx, cnt = np.unique(labels[c], return_index=True, return_counts=True)
            counts[u, c] = cnt
            first[u, c] = idx
        return cls(np.array(states, dtype=np.int8), counts, first)

    def posterior(self):
        totals = self.counts.sum(axis=1)
        probs = totals / np.sum(totals)
        idx = np.flip(np.argsort(probs))
        return PosteriorGenotypeDistribution(self.states[idx], probs[idx])

    def split(self):
        for c in range(self.counts.shape[1]):
            seen = np.flatnonzero(self.counts[:, c] > 0)
            seen = seen[np.argsort(self.first[seen, c], kind="stable")]  # the chain's own first-occurrence order
            yield TraceTally(self.states[seen], self.counts[seen, c:c + 1], self.first[seen, c:c + 1])

    def replicate_incongruence(self, threshold=0.6):
        return _replicate_incongruence([t.posterior() for t in self.split()], threshold)
