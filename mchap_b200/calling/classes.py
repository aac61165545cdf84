"""``CallingMCMC`` with the reference's constructor and ``fit`` signature
(mchap/calling/classes.py:14-124), executed by the CUDA kernel ``call_mcmc_kernel``, and the trace
containers it returns (GenotypeAllelesMultiTrace 127-274, PosteriorGenotypeAllelesDistribution
300-368) as vectorised numpy.
"""
from dataclasses import dataclass

import numpy as np

from ..api import CallBatch, count_genotypes, default_device, raise_item_status
from ..assemble.classes import unique_first_occurrence

__all__ = ["CallingMCMC", "GenotypeAllelesMultiTrace", "PosteriorGenotypeAllelesDistribution"]


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
        from ..jitutils import genotypes_as_indices

        _, ploidy = self.genotypes.shape
        out = np.zeros(count_genotypes(n_alleles, ploidy), dtype=np.float64)
        if len(self.genotypes):
            out[genotypes_as_indices(self.genotypes)] = self.probabilities
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
        modes = [chain.posterior().mode(genotype_support=True) for chain in self.split()]
        alleles = [m[0] for m in modes if m[-1] >= threshold]
        if len({a.tobytes() for a in alleles}) <= 1:
            return 0
        ploidy = len(alleles[0])
        return 2 if len(set(np.array(alleles).ravel())) > ploidy else 1

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

    def fit_batch(self, reads_list, counts_list=None, initial_list=None, haplotypes_list=None, priors=None,
                  seeds=None, return_results=False, replay_words=None):
        """``fit`` for many items in one device call.  haplotypes_list / priors default to the
        model's haplotypes / prior for every item; seeds default to random_seed like the CLI."""
        dev = self.device or default_device()
        n = len(reads_list)
        st = self._step_type()
        haps = [self.haplotypes] * n if haplotypes_list is None else haplotypes_list
        prs = ([self.prior] * n if self.prior is not None else None) if priors is None else priors
        batch = CallBatch(reads_list, haps, self.ploidy, counts_list, prs)
        seed0 = int(self.random_seed) & 0xFFFFFFFF if self.random_seed is not None else int(
            np.random.randint(0, 2 ** 32, dtype=np.uint64))
        items = batch.items
        P = self.ploidy
        per = self.chains * self.steps
        idx = np.arange(n, dtype=np.int64)
        items["gl_off"] = idx * per * P
        items["hap_out_off"] = idx * per
        sd = np.full(n, seed0, dtype=np.uint32) if seeds is None else np.asarray(seeds, dtype=np.uint64).astype(np.uint32)
        items["reserved"] = sd.view(np.int32)
        init = None
        if initial_list is not None and any(i is not None for i in initial_list):
            init = np.full((n, P), -1, dtype=np.int32)
            for i, v in enumerate(initial_list):
                if v is not None:
                    init[i] = np.asarray(v, dtype=np.int32)
        out = dev.call_mcmc(batch, self.steps, self.chains, st, init, P, replay_words)
        res = []
        for i in range(n):
            raise_item_status(int(out["results"]["status"][i]), i if n > 1 else None)
            g = out["alleles"][i * per * P:(i + 1) * per * P].reshape(self.chains, self.steps, P)
            l = out["llks"][i * per:(i + 1) * per].reshape(self.chains, self.steps)
            res.append(GenotypeAllelesMultiTrace(g, l, len(haps[i])))
        if return_results:
            return res, out["results"]
        return res
